// qbot_b200 -- fusion planner: groups the queued gates into sweeps, picks each sweep's tile
// bits, splits the sweep into register stages and serialises the program the tile kernel runs.
// Pure host code (compiled into the CUDA library and into the CPU plan emulator of the tests).
#include "qb_gate.h"
#include "qb_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

QGate qb_classify(const cplx* m, int k, const int* tb, uint64_t cmask) {
    QGate g;
    g.k = k; g.cmask = cmask;
    for (int i = 0; i < k; i++) g.tb[i] = tb[i];
    const int D = 1 << k;
    bool diag = true, mono = true;
    std::vector<int> src(D, -1), colcnt(D, 0);
    for (int i = 0; i < D; i++) {
        int nz = 0;
        for (int j = 0; j < D; j++) {
            const cplx v = m[i * D + j];
            if (v.x != 0.0 || v.y != 0.0) {
                nz++;
                src[i] = j;
                colcnt[j]++;
                if (i != j) diag = false;
            }
        }
        if (nz != 1) mono = false;
    }
    for (int j = 0; j < D && mono; j++) if (colcnt[j] != 1) mono = false;
    if (diag) {
        g.type = QB_G_DIAG;
        g.m.resize(D);
        for (int i = 0; i < D; i++) g.m[i] = m[i * D + i];
    } else if (mono) {
        g.type = QB_G_MONO;
        g.m.resize(D);
        g.src = src;
        for (int i = 0; i < D; i++) g.m[i] = m[i * D + src[i]];
    } else {
        g.type = QB_G_DENSE;
        g.m.assign(m, m + (size_t)D * D);
    }
    return g;
}

bool qb_is_identity(const QGate& g) {
    if (g.type != QB_G_DIAG) return false;
    for (const cplx& v : g.m) if (v.x != 1.0 || v.y != 0.0) return false;
    return true;
}

std::vector<cplx> qb_dense_of(const QGate& g) {
    const int D = 1 << g.k;
    if (g.type == QB_G_DENSE) return g.m;
    std::vector<cplx> m((size_t)D * D, cplx{0.0, 0.0});
    if (g.type == QB_G_DIAG) for (int i = 0; i < D; i++) m[i * D + i] = g.m[i];
    else for (int i = 0; i < D; i++) m[i * D + g.src[i]] = g.m[i];
    return m;
}

namespace {

struct GInfo {
    uint64_t wmask = 0;     // bits the gate acts on non-diagonally ("writes")
    uint64_t rmask = 0;     // bits it only reads (controls, diagonal targets)
    bool tileable = false;  // can run inside a fused sweep
    bool is_x = false;      // (multi-)controlled Pauli-X: a register renaming when its controls are register bits too
};

GInfo analyse(const QGate& g) {
    GInfo gi;
    if (g.type == QB_G_DIAG) {
        gi.rmask = g.tmask() | g.cmask;
        gi.tileable = g.k <= 3;
    } else {
        gi.wmask = g.tmask();
        gi.rmask = g.cmask;
        gi.tileable = g.k <= 2;
        gi.is_x = g.type == QB_G_MONO && g.k == 1 && g.cmask != 0 && g.src.size() == 2 && g.src[0] == 1 && g.src[1] == 0 &&
                  g.m[0].x == 1 && g.m[0].y == 0 && g.m[1].x == 1 && g.m[1].y == 0;
    }
    return gi;
}

// One order-preserving pass: which of `rem` (gate indices, circuit order) can run now, given
// that a gate needs all its write bits inside `allowed` and must not overtake an earlier
// non-commuting gate that stays behind.
void select_pass(const std::vector<int>& rem, const std::vector<GInfo>& info, uint64_t allowed,
                 size_t window, std::vector<int>* picked, int* count) {
    uint64_t blocked_w = 0, blocked_r = 0;   // bits written / read by a gate left behind
    int c = 0;
    size_t n = std::min(rem.size(), window);
    for (size_t x = 0; x < n; x++) {
        const GInfo& gi = info[rem[x]];
        bool ok = gi.tileable && (gi.wmask & ~allowed) == 0 &&
                  ((gi.wmask | gi.rmask) & blocked_w) == 0 && (gi.wmask & blocked_r) == 0;
        if (ok) {
            c++;
            if (picked) picked->push_back(rem[x]);
        } else {
            blocked_w |= gi.wmask;
            blocked_r |= gi.rmask;
            if (blocked_w == ~0ull) break;
        }
    }
    if (count) *count = c;
}

bool is_hadamard_like(const QGate& g, double* s) {
    if (g.type != QB_G_DENSE || g.k != 1) return false;
    const cplx* m = g.m.data();
    if (m[0].y != 0 || m[1].y != 0 || m[2].y != 0 || m[3].y != 0) return false;
    if (!(m[0].x > 0) || m[1].x != m[0].x || m[2].x != m[0].x || m[3].x != -m[0].x) return false;
    *s = m[0].x;
    return true;
}

bool is_pauli_x(const QGate& g) {
    return g.type == QB_G_MONO && g.k == 1 && g.src[0] == 1 && g.src[1] == 0 && g.m[0].x == 1 && g.m[0].y == 0 &&
           g.m[1].x == 1 && g.m[1].y == 0;
}

std::vector<cplx> dense_matrix(const QGate& g) { return qb_dense_of(g); }

int bank_class(int p) {          // contribution of tile-local bit p to (slot mod 8), see qt_slot
    if (p < 3) return 1 << p;
    if (p < QT_L) return 0;
    return 1 << ((p - QT_L) % 3);
}

struct StageBuild {
    QtStage st;
    std::vector<QtOp> ops;
    std::vector<uint64_t> op_w;                  // write mask (index bits) of each op, for phase merging
    std::vector<std::vector<double>> payload;    // pool data of each op (placed at serialisation)
    void push(const QtOp& op, uint64_t w, const double* v, size_t n) {
        ops.push_back(op);
        op_w.push_back(w);
        payload.emplace_back(v, v + n);
    }
};

struct SweepBuild {
    int M = QT_MAXM;                     // tile bits
    int R = QT_R;                        // register bits per stage
    std::vector<int> hb;                 // free tile bits (index positions), ascending
    int local_of[64];                    // index bit -> tile-local position, -1 if outside the tile
    std::vector<StageBuild> stages;
    int ngates = 0;
    double scale = 1.0;                  // product of the 2^-1/2 factors of the unscaled H ops
    double gph_re = 1.0, gph_im = 0.0;   // product of the d0 factors pulled out of register-bit diagonals
    // uncontrolled 1-qubit diagonals are not emitted where they were picked: they commute with
    // everything except writes of their own bit, so they are PLACED after the stages are known
    struct PendingDiag { int stage, pos, bit; double d[4]; };
    std::vector<PendingDiag> pending;
};

// predicate of a gate's controls (and value-controls) split by where each bit lives in this stage
struct Pred {
    uint32_t regsel;
    uint16_t lmask, lval;
    uint64_t gmask, gval;
};

Pred make_pred(const SweepBuild& sw, const QtStage& st, uint64_t ones_mask, uint64_t zeros_mask) {
    const int R = sw.R;
    Pred p{0, 0, 0, 0, 0};
    uint32_t reg_need1 = 0, reg_need0 = 0;
    for (int b = 0; b < 64; b++) {
        bool one = (ones_mask >> b) & 1ull, zero = (zeros_mask >> b) & 1ull;
        if (!one && !zero) continue;
        int lp = sw.local_of[b];
        if (lp < 0) {
            p.gmask |= 1ull << b;
            if (one) p.gval |= 1ull << b;
            continue;
        }
        int ri = -1;
        for (int q = 0; q < R; q++) if (st.rb[q] == lp) ri = q;
        if (ri >= 0) { if (one) reg_need1 |= 1u << ri; else reg_need0 |= 1u << ri; }
        else { p.lmask |= (uint16_t)(1u << lp); if (one) p.lval |= (uint16_t)(1u << lp); }
    }
    for (int i = 0; i < (1 << R); i++)
        if ((i & reg_need1) == reg_need1 && (i & reg_need0) == 0) p.regsel |= 1u << i;
    return p;
}

int reg_index(const QtStage& st, int lp, int R) {
    for (int q = 0; q < R; q++) if (st.rb[q] == lp) return q;
    return -1;
}

uint32_t all_regs(int R) { return R >= 5 ? 0xffffffffu : (1u << (1 << R)) - 1u; }

void set_pred(QtOp& op, const Pred& p, int R) {
    op.regsel = p.regsel; op.lmask = p.lmask; op.lval = p.lval; op.gmask = p.gmask; op.gval = p.gval;
    op.flags = (uint16_t)((p.gmask ? QT_FLAG_GLOBAL : 0u) | (p.regsel == all_regs(R) ? QT_FLAG_ALLREG : 0u));
}

double phase_code(int loc, int pos) {
    const int64_t c = (int64_t)(loc | (pos << 8));
    double d;
    memcpy(&d, &c, sizeof(d));
    return d;
}

// append one gate to a stage as one or more ops
void emit_gate(SweepBuild& sw, StageBuild& sb, const QGate& g, bool merge_phases) {
    QtOp op;
    memset(&op, 0, sizeof(op));
    const QtStage& st = sb.st;
    if (g.type == QB_G_DIAG) {
        if (g.k == 1 && g.cmask == 0) {
            // uncontrolled 1-qubit diagonal: entry of a PHASE op (merged backwards when it commutes)
            const int bit = g.tb[0];
            const int lp = sw.local_of[bit];
            int loc, pos;
            if (lp < 0) { loc = QT_LOC_GLOBAL; pos = bit; }
            else {
                int ri = reg_index(st, lp, sw.R);
                if (ri >= 0) { loc = QT_LOC_REG; pos = ri; } else { loc = QT_LOC_LOCAL; pos = lp; }
            }
            double ent[5] = {phase_code(loc, pos), g.m[0].x, g.m[0].y, g.m[1].x, g.m[1].y};
            if (merge_phases) {
                sw.pending.push_back({(int)sw.stages.size(), (int)sb.ops.size(), bit, {g.m[0].x, g.m[0].y, g.m[1].x, g.m[1].y}});
                return;
            }
            op.type = QT_OP_PHASE;
            op.nent = 1;
            op.regsel = all_regs(sw.R);
            op.flags = QT_FLAG_ALLREG;
            sb.push(op, 0, ent, 5);
            return;
        }
        // (controlled / multi-qubit) diagonal: one CDIAG per assignment of the leading target bits
        const int k = g.k;
        const int last = g.tb[k - 1];
        for (int v = 0; v < (1 << (k - 1)); v++) {
            cplx d0 = g.m[2 * v], d1 = g.m[2 * v + 1];
            if (d0.x == 1 && d0.y == 0 && d1.x == 1 && d1.y == 0) continue;
            uint64_t ones = g.cmask, zeros = 0;
            for (int j = 0; j < k - 1; j++) {
                if ((v >> (k - 2 - j)) & 1) ones |= 1ull << g.tb[j]; else zeros |= 1ull << g.tb[j];
            }
            QtOp o;
            memset(&o, 0, sizeof(o));
            o.type = QT_OP_CDIAG;
            set_pred(o, make_pred(sw, st, ones, zeros), sw.R);
            const int lp = sw.local_of[last];
            if (lp < 0) { o.t1 = QT_LOC_GLOBAL; o.t0 = (uint8_t)last; }
            else {
                int ri = reg_index(st, lp, sw.R);
                if (ri >= 0) { o.t1 = QT_LOC_REG; o.t0 = (uint8_t)ri; } else { o.t1 = QT_LOC_LOCAL; o.t0 = (uint8_t)lp; }
            }
            double d[4] = {d0.x, d0.y, d1.x, d1.y};
            sb.push(o, 0, d, 4);
        }
        return;
    }
    set_pred(op, make_pred(sw, st, g.cmask, 0), sw.R);
    if (g.k == 1) {
        op.t0 = (uint8_t)reg_index(st, sw.local_of[g.tb[0]], sw.R);
        double s;
        if (g.cmask == 0 && is_hadamard_like(g, &s)) {
            // unscaled butterfly; the factor joins the sweep's scalar
            op.type = QT_OP_H;
            sw.scale *= s;
            sb.push(op, g.tmask(), nullptr, 0);
        }
        else if (is_pauli_x(g)) { op.type = QT_OP_X; sb.push(op, g.tmask(), nullptr, 0); }
        else {
            std::vector<cplx> m = dense_matrix(g);
            op.type = QT_OP_U2;
            sb.push(op, g.tmask(), (const double*)m.data(), 8);
        }
    } else {
        std::vector<cplx> m = dense_matrix(g);
        op.type = QT_OP_U4;
        op.t0 = (uint8_t)reg_index(st, sw.local_of[g.tb[0]], sw.R);
        op.t1 = (uint8_t)reg_index(st, sw.local_of[g.tb[1]], sw.R);
        sb.push(op, g.tmask(), (const double*)m.data(), 32);
    }
}

void choose_thread_bits(QtStage& st, int M, int R, bool io_stage) {
    bool is_reg[QT_MAXM] = {false};
    for (int q = 0; q < R; q++) is_reg[st.rb[q]] = true;
    std::vector<int> freep;
    for (int p = 0; p < M; p++) if (!is_reg[p]) freep.push_back(p);
    std::vector<int> order;
    if (io_stage) {
        // the lanes carry the contiguous low bits: every warp-wide 128-bit access is one 512-byte run
        order = freep;       // ascending: positions 0..4 first (register bits of an IO stage are >= QT_L)
    } else {
        // the three lowest thread bits select the 16-byte bank group of an LDS.128 phase: give them
        // tile bits of three different bank classes when available
        for (int cls : {1, 2, 4}) {
            for (size_t x = 0; x < freep.size(); x++) {
                if (bank_class(freep[x]) == cls) { order.push_back(freep[x]); freep.erase(freep.begin() + x); break; }
            }
        }
        for (int p : freep) order.push_back(p);
    }
    for (int q = 0; q < M - R; q++) st.tpos[q] = (uint8_t)order[q];
}

std::vector<uint8_t> serialise(const SweepBuild& sw, double header_scale) {
    QtHeader h;
    memset(&h, 0, sizeof(h));
    size_t nops = 0, npool = 0;
    for (const auto& s : sw.stages) {
        nops += s.ops.size();
        for (const auto& pl : s.payload) npool += pl.size();
    }
    h.nstages = (uint16_t)sw.stages.size();
    h.nops = (uint16_t)nops;
    h.M = (uint16_t)sw.M;
    h.R = (uint8_t)sw.R;
    h.ngates = (uint16_t)sw.ngates;
    h.scale = header_scale;
    static_assert(sizeof(QtHeader) <= QT_STAGES_OFF, "header too large");
    static_assert(QT_STAGES_OFF + QT_MAX_STAGES * sizeof(QtStage) <= QT_OPS_OFF, "stage table too large");
    static_assert(sizeof(QtOp) == 40, "QtOp must be 40 bytes");
    if (sw.stages.size() > QT_MAX_STAGES || nops > QT_MAX_OPS) return {};
    h.stages_off = QT_STAGES_OFF;
    h.ops_off = QT_OPS_OFF;
    h.pool_off = QT_POOL_OFF;
    h.total_bytes = (uint32_t)((h.pool_off + sizeof(double) * npool + 15) & ~size_t(15));
    for (int i = 0; i < sw.M - QT_L; i++) h.hb[i] = h.hbs[i] = (uint8_t)sw.hb[i];
    std::sort(h.hbs, h.hbs + (sw.M - QT_L));
    std::vector<uint8_t> out(h.total_bytes, 0);
    memcpy(out.data(), &h, sizeof(h));
    size_t first = 0, pool_at = 0;
    double* pool = (double*)(out.data() + h.pool_off);
    for (size_t s = 0; s < sw.stages.size(); s++) {
        QtStage st = sw.stages[s].st;
        st.first_op = (uint16_t)first;
        st.nops = (uint16_t)sw.stages[s].ops.size();
        memcpy(out.data() + h.stages_off + s * sizeof(QtStage), &st, sizeof(st));
        for (size_t x = 0; x < sw.stages[s].ops.size(); x++) {
            QtOp op = sw.stages[s].ops[x];
            const std::vector<double>& pl = sw.stages[s].payload[x];
            op.pool = (uint32_t)pool_at;
            if (!pl.empty()) memcpy(pool + pool_at, pl.data(), sizeof(double) * pl.size());
            pool_at += pl.size();
            memcpy(out.data() + h.ops_off + (first + x) * sizeof(QtOp), &op, sizeof(op));
        }
        first += sw.stages[s].ops.size();
    }
    return out;
}

// greedy choice of a stage's register bits among the tile-local positions [lo, M): returns the
// index-bit mask of the chosen positions (exactly R of them, filled up with the highest unused
// positions when fewer are useful) and how many gates of `rem` the stage can run
uint64_t choose_stage_bits(const std::vector<int>& rem, const std::vector<GInfo>& info, const uint64_t* index_of_local,
                           int M, int R, int lo, int* count_out) {
    uint64_t regmask = 0;
    int cur = 0, npicked = 0;
    select_pass(rem, info, regmask, rem.size(), nullptr, &cur);
    while (npicked < R) {
        int best = -1, best_count = cur;
        for (int lp = lo; lp < M; lp++) {
            const uint64_t bit = 1ull << index_of_local[lp];
            if (regmask & bit) continue;
            int c;
            select_pass(rem, info, regmask | bit, rem.size(), nullptr, &c);
            if (c > best_count) { best_count = c; best = lp; }
        }
        if (best >= 0) {
            regmask |= 1ull << index_of_local[best];
            cur = best_count;
            npicked++;
            continue;
        }
        // no single bit helps (e.g. a two-target gate needs both): try pairs
        int pa = -1, pb = -1, bc = cur;
        if (npicked + 2 <= R) {
            for (int a = lo; a < M; a++) for (int b2 = a + 1; b2 < M; b2++) {
                const uint64_t bits = (1ull << index_of_local[a]) | (1ull << index_of_local[b2]);
                if (regmask & bits) continue;
                int c;
                select_pass(rem, info, regmask | bits, rem.size(), nullptr, &c);
                if (c > bc) { bc = c; pa = a; pb = b2; }
            }
        }
        if (pa >= 0) {
            regmask |= (1ull << index_of_local[pa]) | (1ull << index_of_local[pb]);
            cur = bc;
            npicked += 2;
            continue;
        }
        break;
    }
    // Spare register bits go first to the CONTROLS of the controlled-X gates this stage runs: an X
    // whose controls are register bits too is a pure renaming of registers in the specialised
    // kernel (no instruction), whereas a control on a thread bit costs a predicated swap of half the
    // thread's amplitudes.
    if (npicked < R) {
        std::vector<int> picked;
        select_pass(rem, info, regmask, rem.size(), &picked, nullptr);
        int score[QT_MAXM] = {0};
        for (int gi : picked) {
            if (!info[gi].is_x) continue;
            for (int lp = lo; lp < M; lp++)
                if ((info[gi].rmask >> index_of_local[lp]) & 1ull) score[lp]++;
        }
        while (npicked < R) {
            int best = -1;
            for (int lp = M - 1; lp >= lo; lp--)
                if (!(regmask & (1ull << index_of_local[lp])) && score[lp] > 0 && (best < 0 || score[lp] > score[best])) best = lp;
            if (best < 0) break;
            regmask |= 1ull << index_of_local[best];
            npicked++;
        }
    }
    for (int lp = M - 1; lp >= lo && npicked < R; lp--) {     // fill up
        const uint64_t bit = 1ull << index_of_local[lp];
        if (!(regmask & bit)) { regmask |= bit; npicked++; }
    }
    if (count_out) *count_out = cur;
    return regmask;
}

StageBuild make_stage(uint64_t regmask, const uint64_t* index_of_local, int M, int R, bool io_stage) {
    StageBuild sb;
    memset(&sb.st, 0, sizeof(sb.st));
    int q = 0;
    for (int lp = 0; lp < M; lp++) if (regmask & (1ull << index_of_local[lp])) sb.st.rb[q++] = (uint8_t)lp;
    choose_thread_bits(sb.st, M, R, io_stage);
    return sb;
}

// Placement of the sweep's uncontrolled 1-qubit diagonals (RZ, Z, S, T, ...).  An entry may sit
// anywhere between the previous and the next WRITE of its bit; at the end of a stage it commutes
// with the whole stage as long as no later op of that stage writes the bit.  One PHASE op costs a
// complex multiply per amplitude however many entries it carries, and entries on register bits
// cost a factor table on top -- so every entry goes to the tail of ONE stage inside its window,
// preferably a stage where its bit is not a register bit and that already has a tail PHASE op.
// Only entries sandwiched between two writes inside one stage stay where they were.
void place_diagonals(SweepBuild& sw) {
    if (sw.pending.empty()) return;
    const int ns = (int)sw.stages.size();
    std::vector<std::vector<const SweepBuild::PendingDiag*>> tail(ns);
    std::vector<std::vector<const SweepBuild::PendingDiag*>> inplace(ns);
    // first pass: windows
    struct Win { int lo, hi; };
    std::vector<Win> win(sw.pending.size());
    for (size_t x = 0; x < sw.pending.size(); x++) {
        const auto& e = sw.pending[x];
        const uint64_t b = 1ull << e.bit;
        const int es = std::min(e.stage, ns - 1);            // (a trailing op-less IO stage may have been appended later)
        int ps = -1, nsx = ns;                               // stage of the previous / next write of the bit
        for (int s = es; s >= 0 && ps < 0; s--) {
            const auto& ow = sw.stages[s].op_w;
            const int from = s == e.stage ? std::min(e.pos, (int)ow.size()) : (int)ow.size();
            for (int o = from - 1; o >= 0; o--) if (ow[o] & b) { ps = s; break; }
        }
        for (int s = e.stage; s < ns && nsx == ns; s++) {
            const auto& ow = sw.stages[s].op_w;
            const int from = s == e.stage ? std::min(e.pos, (int)ow.size()) : 0;
            for (int o = from; o < (int)ow.size(); o++) if (ow[o] & b) { nsx = s; break; }
        }
        win[x] = {ps < 0 ? 0 : ps, nsx - 1};
        if (e.stage >= ns) win[x].lo = std::min(win[x].lo, ns - 1);
    }
    // second pass: sandwiched entries stay, the others go to a stage tail.  The tails that carry a PHASE op are
    // chosen first as a MINIMUM set of stages that meets every window (intervals on a line: sort by right end,
    // take the right end of every window that is not met yet) -- every PHASE op is a complex multiply per
    // amplitude, so their number is what counts; each entry then picks, among the chosen stages of its window,
    // one where its bit is not a register bit if there is one (no factor table), the latest on ties.
    const bool pierce = getenv("QBOT_B200_PHASE_GREEDY") == nullptr;
    std::vector<char> chosen(ns, pierce ? 0 : 1);
    if (pierce) {
        std::vector<size_t> order;
        for (size_t x = 0; x < sw.pending.size(); x++) if (win[x].lo <= win[x].hi) order.push_back(x);
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return win[a].hi != win[b].hi ? win[a].hi < win[b].hi : win[a].lo < win[b].lo; });
        // groups of windows met by one point each: the point may sit anywhere in the intersection of its
        // group's windows -- take the stage where the group's PHASE op is cheapest (all entries on register
        // bits: only 32 - 32 / 2^bits amplitudes are multiplied; any other entry: all 32), the latest on ties
        std::vector<char> done(sw.pending.size(), 0);
        for (size_t oi = 0; oi < order.size(); oi++) {
            const size_t x = order[oi];
            if (done[x]) continue;
            bool met = false;
            for (int s = win[x].lo; s <= win[x].hi && !met; s++) met = chosen[s] != 0;
            if (met) { done[x] = 1; continue; }
            const int h = win[x].hi;
            int glo = 0;
            std::vector<size_t> group;
            for (size_t oj = oi; oj < order.size(); oj++) {
                const size_t y = order[oj];
                if (done[y] || win[y].lo > h) continue;
                bool my = false;
                for (int s = win[y].lo; s <= win[y].hi && !my; s++) my = chosen[s] != 0;
                if (my) continue;
                group.push_back(y);
                glo = std::max(glo, win[y].lo);
            }
            int best = h, best_cost = 1 << 30;
            for (int s = h; s >= glo; s--) {
                uint32_t regbits = 0;
                bool other = false;
                for (size_t y : group) {
                    const int lp = sw.local_of[sw.pending[y].bit];
                    const int ri = lp >= 0 ? reg_index(sw.stages[s].st, lp, sw.R) : -1;
                    if (ri >= 0) regbits |= 1u << ri; else other = true;
                }
                const int cost = other ? 32 : 32 - (32 >> __builtin_popcount(regbits));
                if (cost < best_cost) { best_cost = cost; best = s; }
            }
            chosen[best] = 1;
            for (size_t y : group) done[y] = 1;
        }
    }
    std::vector<int> tail_count(ns, 0);
    for (size_t x = 0; x < sw.pending.size(); x++) {
        const auto& e = sw.pending[x];
        if (win[x].lo > win[x].hi) { inplace[std::min(e.stage, ns - 1)].push_back(&e); continue; }
        const int lp = sw.local_of[e.bit];
        int best = -1, best_score = -1;
        for (int s = win[x].lo; s <= win[x].hi; s++) {
            if (!chosen[s]) continue;
            const bool is_reg = lp >= 0 && reg_index(sw.stages[s].st, lp, sw.R) >= 0;
            const int score = (is_reg ? 0 : 2) + (tail_count[s] > 0 ? 1 : 0);
            if (score > best_score || (score == best_score && s > best)) { best_score = score; best = s; }
        }
        tail[best].push_back(&e);
        tail_count[best]++;
    }
    std::vector<const SweepBuild::PendingDiag*> counted;       // entries whose d0 already joined the global factor
    auto entry_of = [&](const StageBuild& sb, const SweepBuild::PendingDiag& e, double* ent) {
        const int lp = sw.local_of[e.bit];
        int loc, pos;
        if (lp < 0) { loc = QT_LOC_GLOBAL; pos = e.bit; }
        else {
            const int ri = reg_index(sb.st, lp, sw.R);
            if (ri >= 0) { loc = QT_LOC_REG; pos = ri; } else { loc = QT_LOC_LOCAL; pos = lp; }
        }
        ent[0] = phase_code(loc, pos);
        for (int i = 0; i < 4; i++) ent[1 + i] = e.d[i];
        if (loc == QT_LOC_REG) {
            // diag(d0, d1) = d0 * diag(1, d1/d0): the scalar joins the sweep's global factor (applied once,
            // where a thread-level factor is multiplied in anyway), and the specialised kernel then has
            // nothing to do for the amplitudes whose register bit is 0
            const double d0r = e.d[0], d0i = e.d[1], d1r = e.d[2], d1i = e.d[3];
            const double n2 = d0r * d0r + d0i * d0i;
            if (n2 > 0.0) {
                ent[1] = 1.0; ent[2] = 0.0;
                ent[3] = (d1r * d0r + d1i * d0i) / n2;
                ent[4] = (d1i * d0r - d1r * d0i) / n2;
                if (std::find(counted.begin(), counted.end(), &e) == counted.end()) {      // (an entry is encoded more than once)
                    counted.push_back(&e);
                    const double gr = sw.gph_re * d0r - sw.gph_im * d0i, gi = sw.gph_re * d0i + sw.gph_im * d0r;
                    sw.gph_re = gr; sw.gph_im = gi;
                }
            }
        }
    };
    auto phase_op = [&](const StageBuild& sb, const std::vector<const SweepBuild::PendingDiag*>& es, QtOp* op, std::vector<double>* pl) {
        memset(op, 0, sizeof(*op));
        op->type = QT_OP_PHASE;
        op->nent = (uint8_t)es.size();
        op->regsel = all_regs(sw.R);
        op->flags = QT_FLAG_ALLREG;
        pl->clear();
        for (const auto* e : es) {
            double ent[5];
            entry_of(sb, *e, ent);
            pl->insert(pl->end(), ent, ent + 5);
        }
    };
    for (int s = 0; s < ns; s++) {
        StageBuild& sb = sw.stages[s];
        if (!inplace[s].empty()) {
            // rebuild the op list with the sandwiched entries at their original positions (entries at
            // the same position share one PHASE op)
            StageBuild nb;
            nb.st = sb.st;
            size_t at = 0;
            for (size_t o = 0; o <= sb.ops.size(); o++) {
                while (at < inplace[s].size() && (size_t)std::min(inplace[s][at]->pos, (int)sb.ops.size()) == o) {
                    const SweepBuild::PendingDiag* e = inplace[s][at++];
                    double ent[5];
                    entry_of(nb, *e, ent);
                    // merge backwards into an earlier PHASE op as long as no write of this bit is crossed
                    bool merged = false;
                    for (int x = (int)nb.ops.size() - 1; x >= 0; x--) {
                        if (nb.ops[x].type == QT_OP_PHASE && nb.ops[x].nent < 200) {
                            nb.payload[x].insert(nb.payload[x].end(), ent, ent + 5);
                            nb.ops[x].nent++;
                            merged = true;
                            break;
                        }
                        if (nb.op_w[x] & (1ull << e->bit)) break;
                    }
                    if (!merged) {
                        QtOp op;
                        std::vector<double> pl;
                        phase_op(nb, {e}, &op, &pl);
                        nb.push(op, 0, pl.data(), pl.size());
                    }
                }
                if (o < sb.ops.size()) nb.push(sb.ops[o], sb.op_w[o], sb.payload[o].data(), sb.payload[o].size());
            }
            sb.ops.swap(nb.ops);
            sb.op_w.swap(nb.op_w);
            sb.payload.swap(nb.payload);
        }
        // tail entries: into an existing PHASE op of the stage when no write of the bit follows it, else
        // into one PHASE op at the end of the stage
        std::vector<const SweepBuild::PendingDiag*> rest;
        for (const auto* e : tail[s]) {
            double ent[5];
            entry_of(sb, *e, ent);
            bool merged = false;
            for (int x = (int)sb.ops.size() - 1; x >= 0; x--) {
                if (sb.ops[x].type == QT_OP_PHASE && sb.ops[x].nent < 200) {
                    sb.payload[x].insert(sb.payload[x].end(), ent, ent + 5);
                    sb.ops[x].nent++;
                    merged = true;
                    break;
                }
                if (sb.op_w[x] & (1ull << e->bit)) break;
            }
            if (!merged) rest.push_back(e);
        }
        for (size_t g0 = 0; g0 < rest.size(); g0 += 200) {
            std::vector<const SweepBuild::PendingDiag*> part(rest.begin() + g0, rest.begin() + std::min(rest.size(), g0 + 200));
            QtOp op;
            std::vector<double> pl;
            phase_op(sb, part, &op, &pl);
            sb.push(op, 0, pl.data(), pl.size());
        }
    }
    sw.pending.clear();
}

// split one sweep's gates into register stages and serialise; returns false if too large.
// The first and the last stage move the tile between HBM and registers, so their register bits
// are free tile bits (positions >= QT_L) and their lanes the contiguous low bits; stages that
// need a low bit in registers sit in between (an op-less IO stage is added when necessary).
bool build_program(const std::vector<QGate>& gates, const std::vector<GInfo>& info, const std::vector<int>& sweep_gates,
                   const std::vector<int>& hb, int M, int R, bool merge_phases, std::vector<uint8_t>* program) {
    SweepBuild sw;
    sw.M = M;
    sw.R = R;
    sw.hb = hb;
    for (int b = 0; b < 64; b++) sw.local_of[b] = -1;
    for (int b = 0; b < QT_L; b++) sw.local_of[b] = b;
    for (int i = 0; i < M - QT_L; i++) sw.local_of[hb[i]] = QT_L + i;
    sw.ngates = (int)sweep_gates.size();
    uint64_t index_of_local[QT_MAXM];
    for (int b = 0; b < 64; b++) if (sw.local_of[b] >= 0) index_of_local[sw.local_of[b]] = (uint64_t)b;
    uint64_t low_index_mask = 0;
    for (int lp = 0; lp < QT_L; lp++) low_index_mask |= 1ull << index_of_local[lp];

    std::vector<int> rem = sweep_gates;
    bool first = true;
    while (!rem.empty()) {
        uint64_t regmask;
        bool io = false;
        int c_all = 0, c_hi = 0;
        const uint64_t m_all = choose_stage_bits(rem, info, index_of_local, M, R, 0, &c_all);
        const uint64_t m_hi = choose_stage_bits(rem, info, index_of_local, M, R, QT_L, &c_hi);
        if (first) {
            // the first stage must be an IO stage; if free bits alone run nothing it stays op-less
            regmask = m_hi; io = true;
        } else if (c_hi >= c_all) { regmask = m_hi; io = true; }     // can serve as the last stage too
        else regmask = m_all;
        if ((regmask & low_index_mask) == 0) io = true;
        StageBuild sb = make_stage(regmask, index_of_local, M, R, io);
        std::vector<int> picked;
        select_pass(rem, info, regmask, rem.size(), &picked, nullptr);
        if (picked.empty() && !first) return false;   // cannot happen: the head gate always fits some stage
        for (int gi : picked) emit_gate(sw, sb, gates[gi], merge_phases);
        std::vector<int> next;
        size_t pi = 0;
        for (int gi : rem) {
            if (pi < picked.size() && picked[pi] == gi) pi++; else next.push_back(gi);
        }
        rem.swap(next);
        const bool empty_first = first && picked.empty();
        first = false;
        if (empty_first && rem.empty()) break;
        sw.stages.push_back(std::move(sb));
        if (sw.stages.size() > QT_MAX_STAGES) return false;
    }
    if (sw.stages.empty()) return false;
    {   // the last stage must be an IO stage
        const QtStage& last = sw.stages.back().st;
        bool low = false;
        for (int q = 0; q < R; q++) if (last.rb[q] < QT_L) low = true;
        if (low) {
            uint64_t regmask = 0;
            for (int lp = M - 1; lp >= M - R; lp--) regmask |= 1ull << index_of_local[lp];
            sw.stages.push_back(make_stage(regmask, index_of_local, M, R, true));
            if (sw.stages.size() > QT_MAX_STAGES) return false;
        } else {
            // make sure its thread map is the IO one (lanes = low bits)
            choose_thread_bits(sw.stages.back().st, M, R, true);
            // predicates of ops already emitted do not depend on the thread map: only on rb
        }
    }
    place_diagonals(sw);
    size_t nops = 0;
    for (const auto& sb : sw.stages) nops += sb.ops.size();
    if (nops > QT_MAX_OPS) return false;
    // the collected 2^-1/2 factors and the d0 factors pulled out of register-bit diagonals ride as
    // ONE constant on a PHASE op -- preferably one that multiplies a thread-level factor into every
    // amplitude anyway (it has an entry that is not on a register bit) -- else on the header
    double header_scale = 1.0;
    const double cre = sw.scale * sw.gph_re, cim = sw.scale * sw.gph_im;
    if (cre != 1.0 || cim != 0.0) {
        StageBuild* best_sb = nullptr;
        size_t best_x = 0;
        int best_score = -1;
        for (auto& sb : sw.stages) {
            for (size_t x = 0; x < sb.ops.size(); x++) {
                if (sb.ops[x].type != QT_OP_PHASE || sb.ops[x].nent >= 200) continue;
                int nonreg = 0, reg = 0;
                for (int e = 0; e < sb.ops[x].nent; e++) {
                    int64_t code;
                    memcpy(&code, &sb.payload[x][5 * e], sizeof(code));
                    if ((code & 0xff) == QT_LOC_REG) reg++; else nonreg++;
                }
                const int score = nonreg > 0 ? 1000 : reg;
                if (score > best_score) { best_score = score; best_sb = &sb; best_x = x; }
            }
        }
        if (best_sb) {
            double ent[5] = {phase_code(QT_LOC_CONST, 0), cre, cim, cre, cim};
            best_sb->payload[best_x].insert(best_sb->payload[best_x].end(), ent, ent + 5);
            best_sb->ops[best_x].nent++;
        } else {
            header_scale = sw.scale;         // no PHASE op => no register-bit diagonal => the factor is the real scale
        }
    }
    std::vector<uint8_t> prog = serialise(sw, header_scale);
    if (prog.empty() || prog.size() > QT_MAX_PROGRAM_BYTES) return false;
    program->swap(prog);
    return true;
}

}  // namespace

namespace {

struct Lcg {            // deterministic: the same circuit always gets the same plan
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
};

// greedy choice of one sweep's free tile bits: repeatedly add the bit (or pair of bits, for
// two-target gates) that admits the most queued gates.  With `rng` the choice is randomised
// among the best two candidates (plan search, see qt_plan).
std::vector<int> choose_tile_bits(const std::vector<int>& rem, const std::vector<GInfo>& info, int nbits, int NH, size_t window, Lcg* rng) {
    const uint64_t low = (1ull << QT_L) - 1ull;
    uint64_t allowed = low;
    int cur;
    select_pass(rem, info, allowed, window, nullptr, &cur);
    std::vector<int> hb;
    while ((int)hb.size() < NH) {
        // the three best improving bits (count, then higher bit first); greedy takes the best, the
        // randomised variant one of them uniformly
        int top_b[3] = {-1, -1, -1}, top_c[3] = {cur, cur, cur};
        for (int b = QT_L; b < nbits; b++) {
            if (allowed & (1ull << b)) continue;
            int c;
            select_pass(rem, info, allowed | (1ull << b), window, nullptr, &c);
            if (c <= cur) continue;
            for (int x = 0; x < 3; x++) {
                if (top_b[x] < 0 || c > top_c[x]) {
                    for (int y = 2; y > x; y--) { top_b[y] = top_b[y - 1]; top_c[y] = top_c[y - 1]; }
                    top_b[x] = b; top_c[x] = c;
                    break;
                }
            }
        }
        int best = top_b[0], best_count = top_c[0];
        if (rng && best >= 0) {
            int k = 1;
            while (k < 3 && top_b[k] >= 0) k++;
            const int pick = (int)(rng->next() % (uint32_t)k);
            best = top_b[pick]; best_count = top_c[pick];
        }
        if (best < 0 && (int)hb.size() + 2 <= NH) {
            // two-target gates need both bits at once
            int ba = -1, bb = -1;
            for (int a = QT_L; a < nbits; a++) for (int b = a + 1; b < nbits; b++) {
                uint64_t bits = (1ull << a) | (1ull << b);
                if (allowed & bits) continue;
                int c;
                select_pass(rem, info, allowed | bits, window, nullptr, &c);
                if (c > best_count) { best_count = c; ba = a; bb = b; }
            }
            if (ba >= 0) {
                allowed |= (1ull << ba) | (1ull << bb);
                hb.push_back(ba); hb.push_back(bb);
                cur = best_count;
                continue;
            }
        }
        if (best < 0) break;
        allowed |= 1ull << best;
        hb.push_back(best);
        cur = best_count;
    }
    // fill up with the lowest unused bits (longer contiguous runs in HBM)
    for (int b = QT_L; b < nbits && (int)hb.size() < NH; b++) {
        if (!(allowed & (1ull << b))) { allowed |= 1ull << b; hb.push_back(b); }
    }
    std::sort(hb.begin(), hb.end());
    return hb;
}

uint64_t allowed_of(const std::vector<int>& hb) {
    uint64_t a = (1ull << QT_L) - 1ull;
    for (int b : hb) a |= 1ull << b;
    return a;
}

void remove_picked(std::vector<int>& rem, const std::vector<int>& picked) {
    std::vector<int> next;
    size_t pi = 0;
    for (int gi : rem) {
        if (pi < picked.size() && picked[pi] == gi) pi++; else next.push_back(gi);
    }
    rem.swap(next);
}

// Beam search over the tile-bit choices: every node is a set of gates still to run; a node is
// expanded by `branch` randomised greedy tile choices (the first one the deterministic greedy),
// all nodes advance one sweep per round and the `width` nodes with the fewest remaining gates
// survive.  Compared with independent randomised trials this finds one sweep less on the
// benchmark circuits (30 q: 12 instead of 13) for ~0.1-0.3 s of planning.
size_t beam_plan(const std::vector<GInfo>& info, int nbits, int NH, size_t window, int width, int branch,
                 std::vector<std::vector<std::vector<int>>>* guides) {
    struct Node {
        std::vector<int> rem;
        std::vector<std::vector<int>> guide;
        size_t steps = 0;
    };
    std::vector<Node> beams(1);
    beams[0].rem.resize(info.size());
    for (size_t i = 0; i < info.size(); i++) beams[0].rem[i] = (int)i;
    Lcg rng{0x9E3779B97F4A7C15ull ^ (uint64_t)info.size()};
    for (;;) {
        // finished nodes: all nodes have run the same number of rounds, unfused steps aside
        const Node* done = nullptr;
        for (const Node& b : beams) if (b.rem.empty() && (!done || b.steps < done->steps)) done = &b;
        if (done) {
            guides->clear();
            for (const Node& b : beams) if (b.rem.empty() && b.steps == done->steps) guides->push_back(b.guide);
            return done->steps;
        }
        std::vector<Node> next;
        for (Node& b : beams) {
            // gates the tile kernel cannot run go one by one, in order, without branching
            while (!b.rem.empty() && !info[b.rem[0]].tileable) { b.rem.erase(b.rem.begin()); b.steps++; }
            if (b.rem.empty()) { next.push_back(b); continue; }
            std::vector<uint64_t> seen_allowed;
            for (int c = 0; c < branch; c++) {
                std::vector<int> hb = choose_tile_bits(b.rem, info, nbits, NH, window, c == 0 ? nullptr : &rng);
                const uint64_t allowed = allowed_of(hb);
                if (std::find(seen_allowed.begin(), seen_allowed.end(), allowed) != seen_allowed.end()) continue;
                seen_allowed.push_back(allowed);
                std::vector<int> picked;
                select_pass(b.rem, info, allowed, window, &picked, nullptr);
                Node nn;
                nn.rem = b.rem;
                nn.guide = b.guide;
                nn.steps = b.steps + 1;
                if (picked.empty()) nn.rem.erase(nn.rem.begin());      // head gate needs more high bits than a tile has: unfused
                else { nn.guide.push_back(hb); remove_picked(nn.rem, picked); }
                next.push_back(std::move(nn));
            }
        }
        std::stable_sort(next.begin(), next.end(), [](const Node& a, const Node& b) {
            return a.rem.size() != b.rem.size() ? a.rem.size() < b.rem.size() : a.steps < b.steps;
        });
        // drop duplicates (same remaining set) and keep the best `width`
        std::vector<Node> kept;
        for (Node& nn : next) {
            bool dup = false;
            for (const Node& k : kept) if (k.rem == nn.rem) { dup = true; break; }
            if (!dup) kept.push_back(std::move(nn));
            if ((int)kept.size() >= width) break;
        }
        beams.swap(kept);
    }
}

}  // namespace

// Peephole rewrite of the queued gate list (same unitary, cheaper gates):
//   X_t (any controls C) ... H_t   ==   ... H_t  Z_t (controls C)       because  H X = Z H
//   H_t ... X_t (controls C)       ==   Z_t (controls C)  H_t ...       because  X H = H Z
// when nothing between the two touches t and nothing acts non-diagonally on a control in C.
// The controlled X would need its target bit inside the tile and costs a (predicated) exchange of
// half of every thread's amplitudes; the controlled Z is diagonal -- a free rider on any sweep,
// wherever its bits live, costing one negation per touched amplitude.  In the random benchmark
// circuits more than half of the CNOT / Toffoli gates have a Hadamard next to them on their target.
std::vector<QGate> qt_peephole(const std::vector<QGate>& in, int level, int* rewritten) {
    std::vector<QGate> g = in;
    int count = 0;
    const size_t LOOKAHEAD = 512;
    if (level <= 0) { if (rewritten) *rewritten = 0; return g; }
    for (size_t i = 0; i < g.size(); i++) {
        if (!is_pauli_x(g[i])) continue;
        const int t = g[i].tb[0];
        const uint64_t C = g[i].cmask, tbit = 1ull << t;
        size_t j = i + 1;
        bool found = false;
        for (; j < g.size() && j <= i + LOOKAHEAD; j++) {
            const QGate& h = g[j];
            const uint64_t touched = h.tmask() | h.cmask;
            if (touched & tbit) {
                double sc;
                found = h.cmask == 0 && h.k == 1 && is_hadamard_like(h, &sc);
                break;
            }
            const uint64_t writes = h.type == QB_G_DIAG ? 0ull : h.tmask();
            if (writes & C) break;                  // the control stops being a plain predicate across h
        }
        if (!found) continue;
        QGate z;
        z.type = QB_G_DIAG;
        z.k = 1;
        z.tb[0] = t;
        z.cmask = C;
        z.m = {cplx{1.0, 0.0}, cplx{-1.0, 0.0}};
        g.insert(g.begin() + (long)j + 1, z);       // after the Hadamard
        g.erase(g.begin() + (long)i);               // the X is gone (indices >= i shift down by one)
        count++;
        i--;                                        // re-examine the gate that moved into slot i
    }
    // the mirror image:  H_t ... X_t (controls C)   ==   Z_t (controls C) H_t ...   (X H = H Z read
    // the other way round), for the X gates the forward rule left
    const bool backward = level >= 2;
    for (size_t i = 0; backward && i < g.size(); i++) {
        if (!is_pauli_x(g[i])) continue;
        const int t = g[i].tb[0];
        const uint64_t C = g[i].cmask, tbit = 1ull << t;
        bool found = false;
        size_t j = i;
        while (j > 0 && i - j < LOOKAHEAD) {
            j--;
            const QGate& h = g[j];
            const uint64_t touched = h.tmask() | h.cmask;
            if (touched & tbit) {
                double sc;
                found = h.cmask == 0 && h.k == 1 && is_hadamard_like(h, &sc);
                break;
            }
            const uint64_t writes = h.type == QB_G_DIAG ? 0ull : h.tmask();
            if (writes & C) break;
        }
        if (!found) continue;
        QGate z;
        z.type = QB_G_DIAG;
        z.k = 1;
        z.tb[0] = t;
        z.cmask = C;
        z.m = {cplx{1.0, 0.0}, cplx{-1.0, 0.0}};
        g.erase(g.begin() + (long)i);
        g.insert(g.begin() + (long)j, z);            // before the Hadamard; slot i now holds the gate that followed the X's predecessor
        count++;
    }
    if (rewritten) *rewritten = count;
    return g;
}

namespace {

// number of stages of a program whose three lowest thread bits do not cover the three bank classes
// (IO stages always do: their lanes are the low index bits)
int conflicting_stages(const std::vector<uint8_t>& program) {
    const QtHeader* h = (const QtHeader*)program.data();
    const QtStage* st = (const QtStage*)(program.data() + h->stages_off);
    int bad = 0;
    for (int s = 0; s < h->nstages; s++) {
        int seen = 0;
        for (int q = 0; q < 3; q++) seen |= bank_class(st[s].tpos[q]);
        if (seen != 7) bad++;
    }
    return bad;
}

// build the steps (sweep programs + unfused gates) following `guide` (tile bits of the fused sweeps in
// order; greedy wherever the guide ends or stops being valid)
std::vector<QtPlanStep> build_steps(const std::vector<QGate>& gates, const std::vector<GInfo>& info, int nbits,
                                    const QtPlanOptions& opt, const std::vector<std::vector<int>>* guide) {
    std::vector<QtPlanStep> steps;
    const int M = opt.M, NH = M - QT_L;
    const bool can_tile = nbits >= M;
    const size_t WINDOW = 512;
    std::vector<int> rem(gates.size());
    for (size_t i = 0; i < gates.size(); i++) rem[i] = (int)i;
    bool have_guide = guide != nullptr;
    size_t guide_at = 0;
    while (!rem.empty()) {
        if (!can_tile || !info[rem[0]].tileable) {
            QtPlanStep st;
            st.fused = false;
            st.gate_index = rem[0];
            st.ngates = 1;
            steps.push_back(std::move(st));
            rem.erase(rem.begin());
            continue;
        }
        // ---- the sweep's free tile bits: from the search, or greedy ----------------------------
        std::vector<int> hb;
        if (have_guide && guide_at < guide->size()) hb = (*guide)[guide_at];
        else hb = choose_tile_bits(rem, info, nbits, NH, WINDOW, nullptr);
        std::vector<int> picked;
        select_pass(rem, info, allowed_of(hb), WINDOW, &picked, nullptr);
        if (picked.empty()) {
            // the head gate is tileable and unblocked, so this only happens if it needs more
            // high bits than the tile has; run it unfused
            QtPlanStep st;
            st.fused = false;
            st.gate_index = rem[0];
            st.ngates = 1;
            steps.push_back(std::move(st));
            rem.erase(rem.begin());
            continue;
        }
        guide_at++;
        // ---- stages + program (shrink the sweep if the program does not fit) -------------------
        std::vector<uint8_t> program;
        while (!build_program(gates, info, picked, hb, M, opt.R, opt.merge_phases, &program)) {
            if (picked.size() <= 1) throw std::runtime_error("planner: cannot build a program for one gate");
            picked.resize(picked.size() / 2);
            have_guide = false;       // the searched plan assumed the whole sweep ran: fall back to greedy
        }
        // Which tile-local position a free tile bit gets decides its shared-memory bank class; if some
        // stage's register bits exhaust a class (2-way conflicts on every transposition access of that
        // stage), try other orders of the same bits.
        if (conflicting_stages(program) > 0) {
            std::vector<int> order = hb, best_order = hb;
            int best_conf = conflicting_stages(program);
            std::sort(order.begin(), order.end());
            int tries = 0;
            while (best_conf > 0 && tries < 200 && std::next_permutation(order.begin(), order.end())) {
                tries++;
                std::vector<uint8_t> cand;
                if (!build_program(gates, info, picked, order, M, opt.R, opt.merge_phases, &cand)) continue;
                const QtHeader* hc = (const QtHeader*)cand.data();
                const QtHeader* hp = (const QtHeader*)program.data();
                if (hc->nstages > hp->nstages) continue;                 // never pay a stage for it
                const int conf = conflicting_stages(cand);
                if (conf < best_conf) { best_conf = conf; program.swap(cand); best_order = order; }
            }
        }
        QtPlanStep st;
        st.fused = true;
        st.gate_index = -1;
        st.ngates = (int)picked.size();
        st.program.swap(program);
        steps.push_back(std::move(st));
        remove_picked(rem, picked);
    }
    return steps;
}

// what a plan costs, lexicographically: passes over the state, then shared-memory transpositions
// (stages), then ops
struct PlanCost {
    size_t steps = 0, stages = 0, ops = 0;
    bool operator<(const PlanCost& o) const {
        if (steps != o.steps) return steps < o.steps;
        if (stages != o.stages) return stages < o.stages;
        return ops < o.ops;
    }
};
PlanCost cost_of(const std::vector<QtPlanStep>& steps) {
    PlanCost c;
    c.steps = steps.size();
    for (const QtPlanStep& st : steps) {
        if (!st.fused) continue;
        const QtHeader* h = (const QtHeader*)st.program.data();
        c.stages += h->nstages;
        c.ops += h->nops;
    }
    return c;
}

}  // namespace

std::vector<QtPlanStep> qt_plan(const std::vector<QGate>& gates, int nbits, const QtPlanOptions& opt) {
    const int M = opt.M;
    if (M < QT_MINM || M > QT_MAXM) throw std::runtime_error("planner: tile bits out of range");
    if (opt.R != QT_R && opt.R != QT_MAXR) throw std::runtime_error("planner: register bits per stage must be 4 or 5");
    const int NH = M - QT_L;
    std::vector<GInfo> info(gates.size());
    for (size_t i = 0; i < gates.size(); i++) info[i] = analyse(gates[i]);
    const bool can_tile = nbits >= M;
    const size_t WINDOW = 512;

    // ---- plan search: the greedy tile-bit choice is a local optimum per sweep.  Every sweep is a
    //      full pass over HBM, so on large states a beam search over the choices (deterministic
    //      seed) is worth its fraction of a second: 14 -> 12 (deepest level: 11) sweeps at 30 q, 9 -> 8 at 34 q.  The
    //      programs of the best few candidates are built and the cheapest plan wins (passes, then
    //      stages, then ops).
    std::vector<QtPlanStep> best = build_steps(gates, info, nbits, opt, nullptr);
    if (can_tile && opt.search_trials > 1 && gates.size() >= 8) {
        // search effort: (beam width, branching) per level; the top level also runs the one below it (the two
        // explore different parts of the tree: the wider beam usually saves a sweep, sometimes only stages)
        struct Level { int width, branch; };
        std::vector<Level> levels;
        if (opt.search_trials >= 128) { levels.push_back({24, 10}); levels.push_back({8, 6}); }
        else if (opt.search_trials >= 32) levels.push_back({8, 6});
        else levels.push_back({4, 4});
        if (const char* e = getenv("QBOT_B200_BEAM")) { int w = 0, b = 0; if (sscanf(e, "%dx%d", &w, &b) == 2 && w > 0 && b > 0) { levels.clear(); levels.push_back({w, b}); } }
        PlanCost best_cost = cost_of(best);
        for (const Level& lv : levels) {
            std::vector<std::vector<std::vector<int>>> guides;
            beam_plan(info, nbits, NH, WINDOW, lv.width, lv.branch, &guides);
            for (const auto& g : guides) {
                std::vector<QtPlanStep> cand = build_steps(gates, info, nbits, opt, &g);
                const PlanCost c = cost_of(cand);
                if (c < best_cost) { best_cost = c; best.swap(cand); }
            }
        }
    }
    return best;
}

// Plan the cheapest of the rewritten variants of a gate list (forward rule only / both rules; env
// QBOT_B200_PEEPHOLE = 0, 1 or 2 pins one): passes over the state first, then stages, then ops.
std::vector<QtPlanStep> qt_plan_best(const std::vector<QGate>& gates, int nbits, const QtPlanOptions& opt, std::vector<QGate>* planned) {
    const char* pin = getenv("QBOT_B200_NO_PEEPHOLE") ? "0" : getenv("QBOT_B200_PEEPHOLE");
    std::vector<int> levels;
    if (pin) levels.push_back(atoi(pin));
    else { levels.push_back(1); levels.push_back(2); }
    std::vector<QtPlanStep> best;
    PlanCost best_cost;
    bool have = false;
    for (int lvl : levels) {
        std::vector<QGate> g = qt_peephole(gates, lvl, nullptr);
        std::vector<QtPlanStep> steps = qt_plan(g, nbits, opt);
        const PlanCost c = cost_of(steps);
        if (!have || c < best_cost) { have = true; best_cost = c; best.swap(steps); planned->swap(g); }
    }
    return best;
}
