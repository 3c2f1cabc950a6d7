// TEST INFRASTRUCTURE: CPU emulator of the fused-sweep plan.
// Runs the planner (qbot_b200/csrc/qb_plan.cpp) on a circuit and executes the resulting
// programs on a host ket exactly as the tile kernel does -- same tile addressing (qt_tile_base /
// qt_run_offset / qt_slot), same thread / register mapping, same op semantics
// (qb_tile_ops.h is shared with the kernel).  Built with g++ by tests/test_planner.py; never
// part of the product library.
#include "../../qbot_b200/csrc/qb_gate.h"
#include "../../qbot_b200/csrc/qb_plan.h"
#include "../../qbot_b200/csrc/qb_tile_ops.h"
#include "../../qbot_b200/csrc/qb_jit.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

static std::string g_err;

static void run_program(const uint8_t* prog, qt_c* psi, int nbits) {
    const QtHeader* h = (const QtHeader*)prog;
    const QtStage* stages = (const QtStage*)(prog + h->stages_off);
    const QtOp* ops = (const QtOp*)(prog + h->ops_off);
    const double* pool = (const double*)(prog + h->pool_off);
    const int M = h->M, NH = M - QT_L;
    const uint64_t ntiles = 1ull << (nbits - M);
    const int T = 1 << (M - QT_R);
    std::vector<qt_c> buf(1u << M);
    // the kernel's contract: first and last stage keep the low QT_L bits in the lanes and only
    // free tile bits in registers
    for (int s : {0, (int)h->nstages - 1}) {
        for (int q = 0; q < QT_R; q++) if (stages[s].rb[q] < QT_L) throw std::runtime_error("IO stage holds a low bit in registers");
        for (int q = 0; q < QT_L; q++) if (stages[s].tpos[q] != q) throw std::runtime_error("IO stage lanes are not the low bits");
    }
    if (h->nops > QT_MAX_OPS || h->nstages > QT_MAX_STAGES) throw std::runtime_error("program exceeds the kernel limits");
    if (h->R != QT_R) throw std::runtime_error("the op interpreter (generic kernel) runs programs with 16 amplitudes per thread only");
    for (uint64_t t = 0; t < ntiles; t++) {
        const uint64_t tbase = qt_tile_base(t, h->hbs, NH);
        for (uint32_t j = 0; j < (1u << M); j++) buf[j] = psi[tbase + (j & 31u) + qt_run_offset(j >> QT_L, h->hb, NH)];
        for (int s = 0; s < h->nstages; s++) {
            const QtStage& st = stages[s];
            for (int tid = 0; tid < T; tid++) {
                const uint32_t lbase = qt_thread_lbase(st, (uint32_t)tid, M);
                qt_c a[QT_NR];
                for (int i = 0; i < QT_NR; i++) a[i] = buf[lbase | qt_reg_offset(st, i)];
                for (int o = 0; o < st.nops; o++) {
                    const QtOp& op = ops[st.first_op + o];
                    if (!qt_op_local_ok(op, lbase) || !qt_op_global_ok(op, tbase)) continue;
                    qt_apply_op(a, op, pool, lbase, tbase);
                }
                if (s + 1 == h->nstages && h->scale != 1.0) qt_scale_real(a, h->scale);
                for (int i = 0; i < QT_NR; i++) buf[lbase | qt_reg_offset(st, i)] = a[i];
            }
        }
        for (uint32_t j = 0; j < (1u << M); j++) psi[tbase + (j & 31u) + qt_run_offset(j >> QT_L, h->hb, NH)] = buf[j];
    }
}

static void run_unfused(const QGate& g, qt_c* psi, int nbits) {
    std::vector<cplx> m = qb_dense_of(g);
    const int D = 1 << g.k;
    const uint64_t total = 1ull << nbits, tmask = g.tmask();
    std::vector<uint64_t> off(D);
    for (int j = 0; j < D; j++) {
        uint64_t o = 0;
        for (int b = 0; b < g.k; b++) if ((j >> (g.k - 1 - b)) & 1) o |= 1ull << g.tb[b];
        off[j] = o;
    }
    std::vector<qt_c> x(D);
    for (uint64_t base = 0; base < total; base++) {
        if (base & tmask) continue;
        if ((base & g.cmask) != g.cmask) continue;
        for (int j = 0; j < D; j++) x[j] = psi[base | off[j]];
        for (int i = 0; i < D; i++) {
            qt_c acc = qt_mk(0, 0);
            for (int j = 0; j < D; j++) acc = qt_fma(qt_mk(m[i * D + j].x, m[i * D + j].y), x[j], acc);
            psi[base | off[i]] = acc;
        }
    }
}

extern "C" {

const char* qbt_last_error() { return g_err.c_str(); }

// gates: ks[g], target bits tbs[g*14..], cmasks[g], dense matrices concatenated (4^k complex each)
// stats out: [steps, fused sweeps, unfused steps, stages, ops, program bytes max, gates in fused sweeps]
int qbt_run(int nbits, int ngates, const int* ks, const int* tbs, const uint64_t* cmasks, const double* mats,
            double* psi, int M, int merge, int execute, long long* stats) {
    try {
        std::vector<QGate> gates;
        size_t moff = 0;
        for (int g = 0; g < ngates; g++) {
            QGate q = qb_classify((const cplx*)(mats + moff), ks[g], tbs + (size_t)g * QB_BIG_MAXK, cmasks[g]);
            moff += 2 * ((size_t)1 << (2 * ks[g]));
            if (!qb_is_identity(q)) gates.push_back(q);
        }
        QtPlanOptions opt;
        opt.M = M;
        opt.merge_phases = merge != 0;
        if (const char* e = getenv("QBOT_B200_PLAN_TRIALS")) opt.search_trials = atoi(e);
        std::vector<QGate> planned;
        std::vector<QtPlanStep> steps = qt_plan_best(gates, nbits, opt, &planned);        // as the engine does
        gates.swap(planned);
        long long st[7] = {0, 0, 0, 0, 0, 0, 0};
        int covered = 0;
        for (const QtPlanStep& s : steps) {
            st[0]++;
            covered += s.ngates;
            if (s.fused) {
                const QtHeader* h = (const QtHeader*)s.program.data();
                st[1]++;
                st[3] += h->nstages;
                st[4] += h->nops;
                if ((long long)s.program.size() > st[5]) st[5] = (long long)s.program.size();
                st[6] += s.ngates;
                if (execute) run_program(s.program.data(), (qt_c*)psi, nbits);
            } else {
                st[2]++;
                if (execute) run_unfused(gates[s.gate_index], (qt_c*)psi, nbits);
            }
        }
        if (covered != (int)gates.size()) { g_err = "plan does not cover every gate exactly once"; return -1; }
        if (stats) memcpy(stats, st, sizeof(st));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// plan only: fused flag / gate index per step and the fused steps' programs, concatenated
int qbt_plan(int nbits, int ngates, const int* ks, const int* tbs, const uint64_t* cmasks, const double* mats, int M, int merge,
             int max_steps, int* nsteps, int* step_fused, int* step_gate, long long* prog_off, long long* prog_len,
             unsigned char* prog_buf, long long prog_cap) {
    try {
        std::vector<QGate> gates;
        size_t moff = 0;
        for (int g = 0; g < ngates; g++) {
            QGate q = qb_classify((const cplx*)(mats + moff), ks[g], tbs + (size_t)g * QB_BIG_MAXK, cmasks[g]);
            moff += 2 * ((size_t)1 << (2 * ks[g]));
            if (!qb_is_identity(q)) gates.push_back(q);
        }
        QtPlanOptions opt;
        opt.M = M;
        opt.merge_phases = merge != 0;
        if (const char* e = getenv("QBOT_B200_PLAN_TRIALS")) opt.search_trials = atoi(e);
        if (const char* e = getenv("QBOT_B200_PLAN_R")) opt.R = atoi(e);
        std::vector<QGate> planned;
        std::vector<QtPlanStep> steps = qt_plan_best(gates, nbits, opt, &planned);
        gates.swap(planned);
        if ((int)steps.size() > max_steps) { g_err = "too many steps"; return -1; }
        long long at = 0;
        for (size_t i = 0; i < steps.size(); i++) {
            step_fused[i] = steps[i].fused ? 1 : 0;
            step_gate[i] = steps[i].gate_index;
            prog_off[i] = at;
            prog_len[i] = (long long)steps[i].program.size();
            if (steps[i].fused) {
                if (at + (long long)steps[i].program.size() > prog_cap) { g_err = "program buffer too small"; return -1; }
                memcpy(prog_buf + at, steps[i].program.data(), steps[i].program.size());
                at += ((long long)steps[i].program.size() + 15) & ~15ll;
            }
        }
        *nsteps = (int)steps.size();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}

// specialised source of one program (valid until the next call) and its run-time coefficients
const char* qbt_jit_source(const unsigned char* program) {
    static std::string src;
    src = qj_generate(program, nullptr);
    return src.c_str();
}

int qbt_jit_pool(const unsigned char* program, double* out, int cap) {
    std::vector<double> p = qj_pool(program);
    if ((int)p.size() > cap) return -1;
    memcpy(out, p.data(), p.size() * sizeof(double));
    return (int)p.size();
}
}
