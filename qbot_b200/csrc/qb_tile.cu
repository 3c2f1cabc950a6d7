// placeholder until the fused tile engine lands
#include "qb_engine.h"
bool qb_engine_available() { return false; }
void qb_engine_run(qb_state*, const std::vector<QGate>&) { throw qb_error(-4, "fused engine not built"); }
