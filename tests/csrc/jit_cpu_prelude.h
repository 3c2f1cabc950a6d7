// TEST INFRASTRUCTURE: CPU definitions of the macros the sweep specialiser's generated text is
// written over (qbot_b200/csrc/qb_jit.h), plus the per-step harness that runs the stage
// functions over every tile and thread the way the CUDA kernel schedules them (one stage for all
// threads of a tile, then the next -- the barriers of the kernel).
#pragma once
#include <vector>
struct QjC { double x, y; };
#define QJ_C QjC
#define QJ_DEV static inline
#define QJ_LD(p) (*(p))
#define QJ_ST(p, v) (*(p) = (v))
#define QJ_P(i) (P[i])
#define QJ_POOL_PARAM const double* P
#define QJ_RESTRICT
#define QJ_WAR_SYNC()
#define QJ_SYNC()
#define QJ_PREFETCH(psi, nbase, tid)
#define QJ_PRELUDE
#define QJ_WANT_DISPATCH

#define QJ_CPU_HARNESS(NAME)                                                                         \
    extern "C" void NAME(QjC* psi, int nbits, const double* pool) {                                  \
        const unsigned long long ntiles = 1ull << (nbits - QJ_M);                                    \
        std::vector<QjC> buf(QJ_TILE_UNITS);                                                         \
        for (unsigned long long t = 0; t < ntiles; t++) {                                            \
            const unsigned long long tbase = qj_tile_base(t);                                        \
            for (int s = 0; s < QJ_NSTAGES; s++)                                                     \
                for (unsigned tid = 0; tid < QJ_T; tid++) qj_stage(s, tid, tbase, psi, buf.data(), pool); \
        }                                                                                            \
    }
