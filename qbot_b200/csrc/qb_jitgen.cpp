// qbot_b200 -- sweep specialiser (see qb_jit.h).  Pure host code, no CUDA dependency.
//
// What is resolved at generation time
//   * the 16 or 32 amplitudes of a thread are named locals (a0..a31), so a Pauli-X / CNOT / Toffoli
//     whose controls are register bits is a RENAMING and costs no instruction at all;
//   * an X under a run-time predicate (controls on thread / tile bits) is deferred: it stays a pending
//     flip that the stage's store offsets, the next Hadamard's signs, a diagonal's factor order or a
//     2x2's column order absorb (see `Pend` below); only conflicts fall back to register exchanges;
//   * every shared-memory and HBM offset is an immediate; the tile base and the thread's
//     position are a few shifts with constant amounts;
//   * predicates on thread bits / tile bits are tests of constant masks; ops whose predicate
//     covers the whole tile have none;
//   * uncontrolled diagonals merged into one PHASE op become one factor table (built once per
//     thread and tile) and ONE complex multiply per amplitude.
#include "qb_jit.h"
#include "qb_plan.h"

#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <vector>

namespace {

struct Out {
    std::string s;
    void f(const char* fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        int n = vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        if (n > 0) s.append(buf, (size_t)std::min<int>(n, (int)sizeof(buf) - 1));
    }
};

// expression depositing the bits of `var` (bit q, q = 0..) at positions pos[q]; consecutive runs
// are grouped into one mask-and-shift
std::string deposit_expr(const char* var, const int* pos, int nbits, const char* type_suffix) {
    std::string e;
    int q = 0;
    while (q < nbits) {
        int len = 1;
        while (q + len < nbits && pos[q + len] == pos[q] + len) len++;
        char b[160];
        if (pos[q] >= q)
            snprintf(b, sizeof(b), "((%s)(%s & 0x%xu) << %d)", type_suffix, var, ((1u << len) - 1u) << q, pos[q] - q);
        else
            snprintf(b, sizeof(b), "((%s)(%s & 0x%xu) >> %d)", type_suffix, var, ((1u << len) - 1u) << q, q - pos[q]);
        if (!e.empty()) e += " | ";
        e += b;
        q += len;
    }
    if (e.empty()) e = "0";
    return e;
}

struct Gen {
    const QtHeader* h;
    const QtStage* stages;
    const QtOp* ops;
    const double* pool;
    int M, NH, T;
    int R, NR;                       // register bits per stage, amplitudes per thread
    Out o;
    std::vector<std::string> nm;     // current name of logical register i
    int tmp = 0;
    // Deferred conditional X.  An X on register bit T under a run-time predicate (controls on thread or
    // tile bits) would exchange the selected amplitude pairs of the thread with predicated moves.
    // Instead it is kept as a PENDING flip (predicate variable, register set) that later ops are
    // commuted over:
    //   * the stage's store consumes it as a choice between two store offsets,
    //   * a Hadamard on T consumes it as a sign (H X = Z H: negate the "1" outputs where it holds),
    //   * a diagonal on T swaps its two factors where it holds (D X = X D'),
    //   * a renaming X on another bit whose register set depends on bit T (CNOT controlled by T)
    //     spawns a second pending flip on its own target (X_t[T=1 xor f] = X_t[T=1] X_t^f),
    //   * ops that neither read nor depend on bit T commute,
    // and anything else makes the flip happen first (flush = the predicated exchange, later).
    // Invariant: pending flips commute with each other, so the order in which they are consumed is free.
    struct Pend {
        int T; std::string var; uint32_t sel;
        uint32_t lmask = 0, lval = 0;     // thread-bit part of the predicate (tile-local positions) ...
        bool simple = false;              // ... exact only while the flag is a single op's predicate
    };
    std::vector<Pend> pend;
    bool lazy_x = true;
    bool xor_signs = true;     // conditional sign flips as XORs of the sign bit (op_zsign) instead of branches
    // what the op being emitted has to do for the pending flips it consumes
    std::vector<std::pair<std::string, uint32_t>> h_neg;      // after a butterfly: negate registers (mask) where var holds
    struct PSwap { std::string var; uint32_t sel; };
    std::map<int, PSwap> phase_swap;                          // PHASE: register bit -> (var, register set): d0 / d1 trade places where it holds
    std::string u2_swap;                                      // U2: exchange the matrix columns where it holds
    std::string cdiag_swap;                                   // CDIAG on a register bit: exchange d0 / d1 where it holds
    struct ZCtl { std::string var; int T = 0; uint32_t sel = 0; } zsign_ctl;   // sign flip CONTROLLED by a register bit with a pending flip:
                                                                             // where var holds the flip's pairs trade places in the register set
    std::vector<std::string> post_stmts;                      // statements behind the op's (conditional) block

    static bool sig_partial(const std::vector<Pend>& pd, uint32_t sig, uint32_t all) {
        for (size_t k = 0; k < pd.size(); k++) if (((sig >> k) & 1u) && pd[k].sel != all) return true;
        return false;
    }
    bool bank_aware = false;
    bool l2_late = false;
    bool uniform_issue = true; // bulk copies of the next tile issued per warp from warp-uniform addresses (emit_issue_next)
    // would the store offsets chosen by pending flip p put two lanes of an 8-lane wavefront on one bank?
    bool store_conflicts(const QtStage& st, const Pend& p) const {
        uint32_t phase = 0;
        for (int q = 0; q < 3 && q < M - R; q++) phase |= 1u << st.tpos[q];
        if (!(p.lmask & phase)) return false;          // the same choice for all lanes of a wavefront
        if (!p.simple) return true;
        const uint32_t sT = qt_slot(1u << st.rb[p.T]);
        int used = 0;
        for (int j = 0; j < 8; j++) {
            uint32_t lbj = 0;
            for (int q = 0; q < 3 && q < M - R; q++) if ((j >> q) & 1) lbj |= 1u << st.tpos[q];
            const bool f = ((lbj ^ p.lval) & p.lmask & phase) == 0;
            const int bank = (int)((qt_slot(lbj) + (f ? sT : 0u)) & 7u);
            if ((used >> bank) & 1) return true;
            used |= 1 << bank;
        }
        return false;
    }
    uint32_t all_sel() const { return NR >= 32 ? 0xffffffffu : ((1u << NR) - 1u); }
    uint32_t flip_sel(uint32_t sel, int T) const {
        uint32_t r = 0;
        for (int i = 0; i < NR; i++) if ((sel >> i) & 1u) r |= 1u << (i ^ (1 << T));
        return r;
    }
    uint32_t asym(uint32_t sel, int T) const { return sel ^ flip_sel(sel, T); }
    // do the pair exchanges X_Ta[sa] and X_Tb[sb] commute?  (sufficient conditions)
    bool flips_commute(int Ta, uint32_t sa, int Tb, uint32_t sb) const {
        if (Ta == Tb || !(sa & sb)) return true;
        return !(asym(sa, Tb) & sb) && !(asym(sb, Ta) & sa);
    }
    std::string new_flag(const std::string& init) {
        char nmf[32];
        snprintf(nmf, sizeof(nmf), "fx%d", tmp++);
        o.f("    bool %s = (%s);\n", nmf, init.c_str());
        return nmf;
    }
    void flush(size_t k, const char* why = "") {
        const Pend p = pend[k];
        pend.erase(pend.begin() + (long)k);
        o.f("    if (%s) {    // [flip:flush] %s, %d pairs\n", p.var.c_str(), why, __builtin_popcount(p.sel) / 2);
        for (int i = 0; i < NR; i++) {
            if ((i >> p.T) & 1 || !((p.sel >> i) & 1u)) continue;
            const int j = i | (1 << p.T);
            o.f("      { const QJ_C t_ = %s; %s = %s; %s = t_; }\n", nm[i].c_str(), nm[i].c_str(), nm[j].c_str(), nm[j].c_str());
        }
        o.f("    }\n");
    }
    // add a pending flip (T, value of `init`, sel).  Where it does not commute with a pending flip q, one of
    // the two has to happen now: q (its exchange is emitted before everything pending), or the new one --
    // it is then commuted over q like a renaming X (q's flag spawns a flip on T over the registers whose
    // membership in sel depends on q's bit) and emitted as a predicated exchange.  The cheaper one is taken.
    void add_pending(int T, const std::string& init, uint32_t sel, uint32_t lmask, uint32_t lval, bool simple) {
        if (!sel) return;
        int conflicts = 0, min_q = 64;
        bool can_go_first = true;
        for (size_t k = 0; k < pend.size(); k++) {
            const Pend& q = pend[k];
            if (flips_commute(T, sel, q.T, q.sel)) continue;
            conflicts++;
            min_q = std::min(min_q, __builtin_popcount(q.sel));
            const uint32_t ns = asym(sel, q.T) & q.sel;
            if (asym(q.sel, T) || !admissible(T, ns, pend.size())) can_go_first = false;
        }
        if (conflicts && can_go_first && __builtin_popcount(sel) < min_q) {
            std::vector<std::pair<std::string, uint32_t>> spawned;
            uint32_t spawned_dep = lmask;
            for (const Pend& q : pend)
                if (!flips_commute(T, sel, q.T, q.sel)) {
                    spawned.push_back({"(" + q.var + ") && (" + init + ")", asym(sel, q.T) & q.sel});
                    spawned_dep |= q.lmask;
                }
            o.f("    if (%s) {    // [flip:first]\n", init.c_str());
            for (int i = 0; i < NR; i++) {
                if ((i >> T) & 1 || !((sel >> i) & 1u)) continue;
                const int j = i | (1 << T);
                o.f("      { const QJ_C t_ = %s; %s = %s; %s = t_; }\n", nm[i].c_str(), nm[i].c_str(), nm[j].c_str(), nm[j].c_str());
            }
            o.f("    }\n");
            for (const auto& sp : spawned) add_pending(T, sp.first, sp.second, spawned_dep, 0, false);
            return;
        }
        for (size_t k = 0; k < pend.size();) {
            if (!flips_commute(T, sel, pend[k].T, pend[k].sel)) flush(k, "before a conflicting flip"); else k++;
        }
        for (Pend& q : pend)
            if (q.T == T && q.sel == sel) {
                o.f("    %s ^= (%s);\n", q.var.c_str(), init.c_str());
                q.lmask |= lmask;
                q.simple = false;
                return;
            }
        Pend np;
        np.T = T; np.var = new_flag(init); np.sel = sel; np.lmask = lmask; np.lval = lval; np.simple = simple;
        pend.push_back(np);
    }
    // would a new flip commute with every pending one except index `skip`?
    bool admissible(int T, uint32_t sel, size_t skip) const {
        for (size_t k = 0; k < pend.size(); k++)
            if (k != skip && !flips_commute(T, sel, pend[k].T, pend[k].sel)) return false;
        return true;
    }

    // Commute the op about to be emitted (it comes AFTER the pending flips in time, its code BEFORE theirs)
    // over every pending flip: nothing to do, a consumption recorded for the op's emitter, or a flush.
    void settle(const QtOp& op, const std::string& c) {
        const bool conditional = !c.empty();
        const uint32_t R2 = op.regsel;
        h_neg.clear(); phase_swap.clear(); cdiag_swap.clear(); u2_swap.clear(); post_stmts.clear();
        zsign_ctl = ZCtl();
        struct Spawn { int T; std::string init; uint32_t sel; uint32_t dep; };
        std::vector<Spawn> spawns;
        for (size_t k = 0; k < pend.size();) {
            Pend& p = pend[k];
            bool fl = false;
            // generic test for an op that works on register pairs / quads over targets tg[]:
            auto pairs_commute = [&](std::initializer_list<int> tg) {
                uint32_t a1 = 0;
                for (int t : tg) a1 |= asym(p.sel, t);
                return !(a1 & R2) && !(asym(R2, p.T) & p.sel);
            };
            switch (op.type) {
                case QT_OP_X:       // unconditional: a renaming
                    if (op.t0 == p.T || !(p.sel & (R2 | flip_sel(R2, p.T)))) break;
                    if (pairs_commute({(int)op.t0})) break;
                    if (!asym(p.sel, op.t0)) {
                        const uint32_t ns = asym(R2, p.T) & p.sel;
                        if (admissible(op.t0, ns, k)) { spawns.push_back({(int)op.t0, p.var, ns, p.lmask}); break; }
                    }
                    fl = true;
                    break;
                case QT_OP_H:
                    if (op.t0 == p.T) {
                        const uint32_t ov = p.sel & R2;
                        if (!ov) break;
                        uint32_t ones = 0;
                        for (int i = 0; i < NR; i++) if (((ov >> i) & 1u) && ((i >> p.T) & 1)) ones |= 1u << i;
                        if (!conditional) {
                            const uint32_t rest = p.sel & ~R2;
                            if (rest && !admissible(p.T, rest, k)) { fl = true; break; }
                            h_neg.push_back({p.var, ones});
                            p.sel = rest;
                            if (!rest) { pend.erase(pend.begin() + (long)k); continue; }
                        } else fl = true;      // (the planner emits controlled Hadamards as U2)
                        break;
                    }
                    fl = !pairs_commute({(int)op.t0});
                    break;
                case QT_OP_U2:
                    if (op.t0 == p.T) {
                        // U X = U' (columns exchanged): the flip is absorbed into the matrix where it holds
                        const uint32_t ov = p.sel & R2;
                        if (!ov) break;
                        if (ov != R2 || !u2_swap.empty()) { fl = true; break; }
                        if (!conditional) {
                            const uint32_t rest = p.sel & ~R2;
                            if (rest && !admissible(p.T, rest, k)) { fl = true; break; }
                            u2_swap = p.var;
                            p.sel = rest;
                            if (!rest) { pend.erase(pend.begin() + (long)k); continue; }
                        } else if (ov == p.sel) {
                            // absorbed where the op's predicate holds, still pending where it does not
                            u2_swap = new_flag(p.var);
                            pend[k].lmask |= op.lmask;
                            pend[k].simple = false;
                            post_stmts.push_back("    " + p.var + " = " + p.var + " && !(" + c + ");    // [flip:u2cond]\n");
                        } else fl = true;
                        break;
                    }
                    fl = !pairs_commute({(int)op.t0});
                    break;
                case QT_OP_U4:
                    if (op.t0 == p.T || op.t1 == p.T) fl = (p.sel & R2) != 0; else fl = !pairs_commute({(int)op.t0, (int)op.t1});
                    break;
                case QT_OP_CDIAG:
                    if (op.t1 == QT_LOC_REG && op.t0 == p.T) {
                        const uint32_t ov = p.sel & R2;
                        if (!ov) break;
                        if (ov == R2 && cdiag_swap.empty() && zsign_ctl.var.empty()) { cdiag_swap = p.var; break; }
                        fl = true;
                        break;
                    }
                    if ((asym(R2, p.T) & p.sel) != 0) {
                        // the diagonal is CONTROLLED by the flip's bit (its register set depends on bit T).  A sign
                        // flip commutes over the pending exchange when it is applied to the registers that hold its
                        // amplitudes right now: where the flag holds, the partner registers of the flip's pairs
                        if (xor_signs && is_z_like(pool, op.pool) && zsign_ctl.var.empty() && cdiag_swap.empty()) {
                            zsign_ctl.var = p.var; zsign_ctl.T = p.T; zsign_ctl.sel = p.sel;
                            break;
                        }
                        fl = true;
                    }
                    break;
                case QT_OP_PHASE: {
                    bool has = false;
                    for (int e = 0; e < op.nent; e++) {
                        int64_t code;
                        memcpy(&code, pool + op.pool + 5 * e, sizeof(code));
                        if ((code & 0xff) == QT_LOC_REG && (int)(code >> 8) == p.T) has = true;
                    }
                    if (!has) break;
                    if (!conditional && !phase_swap.count(p.T)) {
                        // one flip (any register set): the op is emitted in two variants; several: only full register
                        // sets, as a run-time exchange of the bit's two factors
                        bool all_full = p.sel == all_sel();
                        for (const auto& e : phase_swap) all_full = all_full && e.second.sel == all_sel();
                        if (phase_swap.empty() ? (all_full || xor_signs) : all_full) { phase_swap[p.T] = PSwap{p.var, p.sel}; break; }
                    }
                    fl = true;
                    break;
                }
                default: fl = true; break;
            }
            if (fl) {
                static const char* const kind[] = {"?", "H", "X", "U2", "U4", "CDIAG", "PHASE"};
                char why[64];
                snprintf(why, sizeof(why), "before %s on bit %d (flip on %d)", kind[op.type <= QT_OP_PHASE ? op.type : 0], (int)op.t0, p.T);
                flush(k, why);
            } else k++;
        }
        for (const Spawn& sp : spawns) {
            o.f("    // [flip:spawn]\n");
            add_pending(sp.T, sp.init, sp.sel, sp.dep, 0, false);
        }
    }

    std::string P(uint32_t i) const {
        char b[32];
        snprintf(b, sizeof(b), "QJ_P(%u)", i);
        return b;
    }

    std::string cond_of(const QtOp& op) const {
        std::string c;
        char b[128];
        if (op.lmask) {
            snprintf(b, sizeof(b), "(lb & 0x%xu) == 0x%xu", (unsigned)op.lmask, (unsigned)op.lval);
            c += b;
        }
        if (op.gmask) {
            snprintf(b, sizeof(b), "(tbase & 0x%llxull) == 0x%llxull", (unsigned long long)op.gmask, (unsigned long long)op.gval);
            if (!c.empty()) c += " && ";
            c += b;
        }
        return c;
    }

    // a <- f * a with f = (fr, fi) expressions
    void cmul_into(const std::string& a, const std::string& fr, const std::string& fi) {
        o.f("    { const double xr = %s.x, xi = %s.y; %s.x = %s * xr - %s * xi; %s.y = %s * xi + %s * xr; }\n",
            a.c_str(), a.c_str(), a.c_str(), fr.c_str(), fi.c_str(), a.c_str(), fr.c_str(), fi.c_str());
    }

    void op_h(const QtOp& op) {
        const int t = op.t0;
        for (int i = 0; i < NR; i++) {
            if ((i >> t) & 1 || !((op.regsel >> i) & 1u)) continue;
            const std::string &A = nm[i], &B = nm[i | (1 << t)];
            o.f("    { const double xr = %s.x, xi = %s.y; %s.x = xr + %s.x; %s.y = xi + %s.y; %s.x = xr - %s.x; %s.y = xi - %s.y; }\n",
                A.c_str(), A.c_str(), A.c_str(), B.c_str(), A.c_str(), B.c_str(), B.c_str(), B.c_str(), B.c_str(), B.c_str());
        }
        for (const auto& hn : h_neg) {
            // pending conditional X on this bit: H X = Z H -> flip the sign of the "1" outputs where it holds
            if (xor_signs) {
                o.f("    if (%s) {    // [flip:h]\n", hn.first.c_str());
                for (int i = 0; i < NR; i++) if ((hn.second >> i) & 1u) xsign(nm[i], "0x80000000u");
                o.f("    }\n");
                continue;
            }
            o.f("    if (%s) {    // [flip:h]\n", hn.first.c_str());
            for (int i = 0; i < NR; i++) if ((hn.second >> i) & 1u) negate(nm[i]);
            o.f("    }\n");
        }
        h_neg.clear();
    }

    void op_x(const QtOp& op, bool conditional) {
        const int t = op.t0;
        for (int i = 0; i < NR; i++) {
            if ((i >> t) & 1 || !((op.regsel >> i) & 1u)) continue;
            const int j = i | (1 << t);
            if (conditional)
                o.f("    { const QJ_C t_ = %s; %s = %s; %s = t_; }\n", nm[i].c_str(), nm[i].c_str(), nm[j].c_str(), nm[j].c_str());
            else
                std::swap(nm[i], nm[j]);
        }
    }

    void op_u2(const QtOp& op) {
        const int t = op.t0;
        const uint32_t p = op.pool;
        if (u2_swap.empty())
            o.f("    { const double m0 = %s, m1 = %s, m2 = %s, m3 = %s, m4 = %s, m5 = %s, m6 = %s, m7 = %s;\n", P(p).c_str(),
                P(p + 1).c_str(), P(p + 2).c_str(), P(p + 3).c_str(), P(p + 4).c_str(), P(p + 5).c_str(), P(p + 6).c_str(), P(p + 7).c_str());
        else {
            o.f("    { const bool f_ = %s;    // [flip:u2]\n", u2_swap.c_str());
            for (int r = 0; r < 2; r++)
                o.f("      const double m%d = f_ ? %s : %s, m%d = f_ ? %s : %s, m%d = f_ ? %s : %s, m%d = f_ ? %s : %s;\n", 4 * r, P(p + 4 * r + 2).c_str(),
                    P(p + 4 * r).c_str(), 4 * r + 1, P(p + 4 * r + 3).c_str(), P(p + 4 * r + 1).c_str(), 4 * r + 2, P(p + 4 * r).c_str(),
                    P(p + 4 * r + 2).c_str(), 4 * r + 3, P(p + 4 * r + 1).c_str(), P(p + 4 * r + 3).c_str());
            u2_swap.clear();
        }
        for (int i = 0; i < NR; i++) {
            if ((i >> t) & 1 || !((op.regsel >> i) & 1u)) continue;
            const std::string &A = nm[i], &B = nm[i | (1 << t)];
            o.f("      { const double xr = %s.x, xi = %s.y, yr = %s.x, yi = %s.y;\n", A.c_str(), A.c_str(), B.c_str(), B.c_str());
            o.f("        %s.x = m0 * xr - m1 * xi + m2 * yr - m3 * yi; %s.y = m0 * xi + m1 * xr + m2 * yi + m3 * yr;\n", A.c_str(), A.c_str());
            o.f("        %s.x = m4 * xr - m5 * xi + m6 * yr - m7 * yi; %s.y = m4 * xi + m5 * xr + m6 * yi + m7 * yr; }\n", B.c_str(), B.c_str());
        }
        o.f("    }\n");
    }

    void op_u4(const QtOp& op) {
        const int t0 = op.t0, t1 = op.t1;
        const uint32_t p = op.pool;
        o.f("    {\n");
        for (int i = 0; i < NR; i++) {
            if ((i >> t0) & 1 || (i >> t1) & 1 || !((op.regsel >> i) & 1u)) continue;
            const int idx[4] = {i, i | (1 << t1), i | (1 << t0), i | (1 << t0) | (1 << t1)};
            o.f("      { const QJ_C x0 = %s, x1 = %s, x2 = %s, x3 = %s;\n", nm[idx[0]].c_str(), nm[idx[1]].c_str(), nm[idx[2]].c_str(),
                nm[idx[3]].c_str());
            for (int r = 0; r < 4; r++) {
                std::string re, im;
                for (int c = 0; c < 4; c++) {
                    const std::string mr = P(p + 8 * r + 2 * c), mi = P(p + 8 * r + 2 * c + 1);
                    char b[256];
                    snprintf(b, sizeof(b), "%s%s * x%d.x - %s * x%d.y", c ? " + " : "", mr.c_str(), c, mi.c_str(), c);
                    re += b;
                    snprintf(b, sizeof(b), "%s%s * x%d.y + %s * x%d.x", c ? " + " : "", mr.c_str(), c, mi.c_str(), c);
                    im += b;
                }
                o.f("        %s.x = %s;\n        %s.y = %s;\n", nm[idx[r]].c_str(), re.c_str(), nm[idx[r]].c_str(), im.c_str());
            }
            o.f("      }\n");
        }
        o.f("    }\n");
    }

    bool pool_is_one(uint32_t p) const { return pool[p] == 1.0 && pool[p + 1] == 0.0; }
    bool pool_is_minus_one(uint32_t p) const { return pool[p] == -1.0 && pool[p + 1] == 0.0; }
    void negate(const std::string& a) { o.f("      %s.x = -%s.x; %s.y = -%s.y;\n", a.c_str(), a.c_str(), a.c_str(), a.c_str()); }
    void xsign(const std::string& a, const char* mask) { o.f("      QJ_XSIGN(%s, %s);\n", a.c_str(), mask); }

    static bool is_z_like(const double* pool, uint32_t p) {
        return pool[p] == 1.0 && pool[p + 1] == 0.0 && pool[p + 2] == -1.0 && pool[p + 3] == 0.0;
    }

    // diag(1, -1) under a RUN-TIME predicate (controls on thread / tile bits, a target that is a thread / tile
    // bit, or a pending flip on a register target): a sign flip where the predicate holds.  Written as an FP64
    // negation it is a DADD per component on the busiest pipe (measured on the benchmark's sweeps: 8 % of the
    // FP64 instructions); here it is one XOR of the sign bit per component on the integer pipe, still skipped
    // as a whole where the predicate is false.
    void op_zsign(const QtOp& op, const std::string& c) {
        std::string pred = c.empty() ? std::string() : "(" + c + ")";
        char b[96];
        b[0] = 0;
        if (op.t1 == QT_LOC_LOCAL) snprintf(b, sizeof(b), "((lb >> %d) & 1u)", (int)op.t0);
        else if (op.t1 == QT_LOC_GLOBAL) snprintf(b, sizeof(b), "((tbase >> %d) & 1ull)", (int)op.t0);
        if (b[0]) pred += (pred.empty() ? "" : " && ") + std::string(b);
        if (pred.empty()) pred = "true";
        if (!zsign_ctl.var.empty()) {
            o.f("    if (%s) {    // [flip:zctl] [zsign]\n      if (%s) {\n", pred.c_str(), zsign_ctl.var.c_str());
            for (int i = 0; i < NR; i++)
                if (((op.regsel >> i) & 1u) && (op.t1 != QT_LOC_REG || ((i >> op.t0) & 1)))
                    xsign(nm[((zsign_ctl.sel >> i) & 1u) ? i ^ (1 << zsign_ctl.T) : i], "0x80000000u");
            o.f("      } else {\n");
            for (int i = 0; i < NR; i++)
                if (((op.regsel >> i) & 1u) && (op.t1 != QT_LOC_REG || ((i >> op.t0) & 1))) xsign(nm[i], "0x80000000u");
            o.f("      }\n    }\n");
            zsign_ctl = ZCtl();
            return;
        }
        if (op.t1 != QT_LOC_REG || cdiag_swap.empty()) {
            o.f("    if (%s) {    // [zsign]\n", pred.c_str());
            for (int i = 0; i < NR; i++)
                if (((op.regsel >> i) & 1u) && (op.t1 != QT_LOC_REG || ((i >> op.t0) & 1))) xsign(nm[i], "0x80000000u");
            o.f("    }\n");
            return;
        }
        // pending conditional X on the target bit: Z X = X (-Z), the minus sign sits on the "0" side where it holds
        o.f("    if (%s) {    // [flip:cdiag] [zsign]\n      if (%s) {\n", pred.c_str(), cdiag_swap.c_str());
        for (int i = 0; i < NR; i++) if (((op.regsel >> i) & 1u) && !((i >> op.t0) & 1)) xsign(nm[i], "0x80000000u");
        o.f("      } else {\n");
        for (int i = 0; i < NR; i++) if (((op.regsel >> i) & 1u) && ((i >> op.t0) & 1)) xsign(nm[i], "0x80000000u");
        o.f("      }\n    }\n");
        cdiag_swap.clear();
    }

    void op_cdiag(const QtOp& op) {
        const uint32_t p = op.pool;
        // exact +1 / -1 factors are structural (part of the source text): controlled-Z and the like cost
        // a sign flip per touched amplitude instead of a complex multiply
        const bool z_like = pool_is_one(p) && pool_is_minus_one(p + 2);
        if (z_like && op.t1 != QT_LOC_REG) {
            if (op.t1 == QT_LOC_LOCAL) o.f("    if ((lb >> %d) & 1u) {\n", (int)op.t0);
            else o.f("    if ((tbase >> %d) & 1ull) {\n", (int)op.t0);
            for (int i = 0; i < NR; i++) if ((op.regsel >> i) & 1u) negate(nm[i]);
            o.f("    }\n");
            return;
        }
        if (op.t1 == QT_LOC_REG) {
            const bool one0 = pool_is_one(p), one1 = pool_is_one(p + 2);
            o.f("    { const double d0r = %s, d0i = %s, d1r = %s, d1i = %s;\n", P(p).c_str(), P(p + 1).c_str(), P(p + 2).c_str(), P(p + 3).c_str());
            // inv: a pending conditional X on the target bit holds -> the two factors trade places (D X = X D')
            auto body = [&](bool inv) {
                for (int i = 0; i < NR; i++) {
                    if (!((op.regsel >> i) & 1u)) continue;
                    const bool hi = (((i >> op.t0) & 1) != 0) != inv;
                    if (hi ? one1 : one0) continue;          // exact unit factor: structural, part of the source text
                    if (pool_is_minus_one(hi ? p + 2 : p)) { negate(nm[i]); continue; }
                    o.f("  ");
                    cmul_into(nm[i], hi ? "d1r" : "d0r", hi ? "d1i" : "d0i");
                }
            };
            if (cdiag_swap.empty()) body(false);
            else {
                o.f("      if (%s) {    // [flip:cdiag]\n", cdiag_swap.c_str());
                body(true);
                o.f("      } else {\n");
                body(false);
                o.f("      }\n");
                cdiag_swap.clear();
            }
            o.f("    }\n");
        } else {
            if (op.t1 == QT_LOC_LOCAL) o.f("    { const bool b_ = (lb >> %d) & 1u;\n", (int)op.t0);
            else o.f("    { const bool b_ = (tbase >> %d) & 1ull;\n", (int)op.t0);
            o.f("      const double fr = b_ ? %s : %s, fi = b_ ? %s : %s;\n", P(p + 2).c_str(), P(p).c_str(), P(p + 3).c_str(), P(p + 1).c_str());
            for (int i = 0; i < NR; i++) {
                if (!((op.regsel >> i) & 1u)) continue;
                o.f("  ");
                cmul_into(nm[i], "fr", "fi");
            }
            o.f("    }\n");
        }
    }

    void op_phase(const QtOp& op) {
        const uint32_t p = op.pool;
        bool variant = false;      // decided below, once the entries are known
        const int u = tmp++;
        o.f("    {\n");
        bool have = false;
        std::vector<int> reg_entries[QT_MAXR];
        for (int e = 0; e < op.nent; e++) {
            const uint32_t q = p + 5 * e;
            int64_t code;
            memcpy(&code, pool + q, sizeof(code));
            const int loc = (int)(code & 0xff), pos = (int)(code >> 8);
            if (loc == QT_LOC_REG) { reg_entries[pos].push_back(e); continue; }
            std::string fr, fi;
            if (loc == QT_LOC_CONST) { fr = P(q + 1); fi = P(q + 2); }
            else {
                if (loc == QT_LOC_LOCAL) o.f("      const bool b%d_%d = (lb >> %d) & 1u;\n", u, e, pos);
                else o.f("      const bool b%d_%d = (tbase >> %d) & 1ull;\n", u, e, pos);
                char b[160];
                snprintf(b, sizeof(b), "(b%d_%d ? %s : %s)", u, e, P(q + 3).c_str(), P(q + 1).c_str());
                fr = b;
                snprintf(b, sizeof(b), "(b%d_%d ? %s : %s)", u, e, P(q + 4).c_str(), P(q + 2).c_str());
                fi = b;
            }
            if (!have) { o.f("      double cr = %s, ci = %s;\n", fr.c_str(), fi.c_str()); have = true; }
            else
                o.f("      { const double fr = %s, fi = %s; const double tr = cr * fr - ci * fi; ci = cr * fi + ci * fr; cr = tr; }\n",
                    fr.c_str(), fi.c_str());
        }
        // factor table over the register bits that carry entries; an entry with an empty name is
        // exactly 1 (register-bit diagonals arrive as diag(1, d1/d0)): nothing to multiply
        if (xor_signs && phase_swap.size() == 1) {
            // the two-variant form pays when some registers keep an exact unit factor (no thread-uniform factor,
            // every register bit's d0 exactly one), and it is the only form that handles a partial register set
            bool unit_possible = !have;
            for (int q = 0; q < R; q++)
                for (int ex : reg_entries[q]) unit_possible = unit_possible && pool_is_one(p + 5 * ex + 1);
            variant = unit_possible || phase_swap.begin()->second.sel != all_sel();
        }
        struct Ent { int mask; std::string r, i; };
        std::vector<Ent> table;
        if (have) table.push_back({0, "cr", "ci"});
        int regbits = 0;
        for (int q = 0; q < R; q++) {
            if (reg_entries[q].empty()) continue;
            regbits |= 1 << q;
            // product of this bit's entries
            char d0r[32], d0i[32], d1r[32], d1i[32];
            snprintf(d0r, sizeof(d0r), "d0r_%d", q); snprintf(d0i, sizeof(d0i), "d0i_%d", q);
            snprintf(d1r, sizeof(d1r), "d1r_%d", q); snprintf(d1i, sizeof(d1i), "d1i_%d", q);
            const uint32_t q0 = p + 5 * reg_entries[q][0];
            o.f("      double %s = %s, %s = %s, %s = %s, %s = %s;\n", d0r, P(q0 + 1).c_str(), d0i, P(q0 + 2).c_str(), d1r, P(q0 + 3).c_str(),
                d1i, P(q0 + 4).c_str());
            for (size_t x = 1; x < reg_entries[q].size(); x++) {
                const uint32_t qx = p + 5 * reg_entries[q][x];
                o.f("      { const double t0 = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = t0; }\n", d0r, P(qx + 1).c_str(), d0i,
                    P(qx + 2).c_str(), d0i, d0r, P(qx + 2).c_str(), d0i, P(qx + 1).c_str(), d0r);
                o.f("      { const double t1 = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = t1; }\n", d1r, P(qx + 3).c_str(), d1i,
                    P(qx + 4).c_str(), d1i, d1r, P(qx + 4).c_str(), d1i, P(qx + 3).c_str(), d1r);
            }
            // is the (product) d0 of this bit exactly one?  (structural: single entries come in as (1, 0))
            bool d0_one = reg_entries[q].size() >= 1;
            for (int ex : reg_entries[q]) d0_one = d0_one && pool_is_one(p + 5 * ex + 1);
            auto sw = phase_swap.find(q);
            if (sw != phase_swap.end() && !variant) {
                // pending conditional X on this bit: D X = X D' with the two factors exchanged where it holds
                o.f("      if (%s) { double t_ = %s; %s = %s; %s = t_; t_ = %s; %s = %s; %s = t_; }    // [flip:phase]\n", sw->second.var.c_str(), d0r, d0r, d1r, d1r, d0i,
                    d0i, d1i, d1i);
                d0_one = false;
            }
            std::vector<Ent> next;
            if (table.empty()) {
                if (d0_one) next.push_back({0, "", ""}); else next.push_back({0, d0r, d0i});
                next.push_back({1 << q, d1r, d1i});
            } else {
                for (const Ent& t : table) {
                    for (int v = 0; v < 2; v++) {
                        char nr[48], ni[48];
                        const int mk = t.mask | (v << q);
                        if (v == 0 && d0_one) { next.push_back({mk, t.r, t.i}); continue; }      // times one
                        if (t.r.empty()) { next.push_back({mk, v ? d1r : d0r, v ? d1i : d0i}); continue; }   // one times d
                        snprintf(nr, sizeof(nr), "f%d_%dr", q, mk);
                        snprintf(ni, sizeof(ni), "f%d_%di", q, mk);
                        o.f("      const double %s = %s * %s - %s * %s, %s = %s * %s + %s * %s;\n", nr, t.r.c_str(), v ? d1r : d0r, t.i.c_str(),
                            v ? d1i : d0i, ni, t.r.c_str(), v ? d1i : d0i, t.i.c_str(), v ? d1r : d0r);
                        next.push_back({mk, nr, ni});
                    }
                }
            }
            table.swap(next);
        }
        auto apply_table = [&](int flip_bit, uint32_t flip_sel) {
            for (int i = 0; i < NR; i++) {
                const int src = (flip_bit >= 0 && ((flip_sel >> i) & 1u)) ? i ^ (1 << flip_bit) : i;
                const int mk = src & regbits;
                for (const Ent& t : table)
                    if (t.mask == mk) {
                        if (!t.r.empty()) { o.f("  "); cmul_into(nm[i], t.r, t.i); }
                        break;
                    }
            }
        };
        if (!table.empty()) {
            if (variant) {
                // one pending conditional X on a bit that carries entries: where it holds, the registers of its set
                // take the factor of their partner (D X = X D').  Two variants of the multiplies instead of a
                // run-time exchange of the factors: exact unit factors stay skipped in both, and a flip over a
                // partial register set needs no exchange of amplitudes first
                const auto& sw = *phase_swap.begin();
                o.f("      if (%s) {    // [flip:phase]\n", sw.second.var.c_str());
                apply_table(sw.first, sw.second.sel);
                o.f("      } else {\n");
                apply_table(-1, 0);
                o.f("      }\n");
            } else apply_table(-1, 0);
        }
        o.f("    }\n");
    }

    void emit_op(const QtOp& op) {
        const std::string c = cond_of(op);
        const bool conditional = !c.empty();
        if (lazy_x) {
            if (op.type == QT_OP_X && conditional) { add_pending(op.t0, c, op.regsel, op.lmask, op.lval, true); return; }
            settle(op, c);
        }
        if (op.type == QT_OP_X && !conditional) { op_x(op, false); return; }
        if (xor_signs && op.type == QT_OP_CDIAG && is_z_like(pool, op.pool) &&
            (conditional || op.t1 != QT_LOC_REG || !cdiag_swap.empty() || !zsign_ctl.var.empty())) {
            op_zsign(op, c);
            for (const std::string& st : post_stmts) o.f("%s", st.c_str());
            post_stmts.clear();
            return;
        }
        if (conditional) o.f("    if (%s) {\n", c.c_str());
        switch (op.type) {
            case QT_OP_H: op_h(op); break;
            case QT_OP_X: op_x(op, true); break;
            case QT_OP_U2: op_u2(op); break;
            case QT_OP_U4: op_u4(op); break;
            case QT_OP_CDIAG: op_cdiag(op); break;
            case QT_OP_PHASE: op_phase(op); break;
            default: break;
        }
        if (conditional) o.f("    }\n");
        for (const std::string& st : post_stmts) o.f("%s", st.c_str());
        post_stmts.clear();
    }

    uint64_t hbm_reg_offset(const QtStage& st, int i) const {
        uint64_t off = 0;
        for (int q = 0; q < R; q++) if ((i >> q) & 1) off |= 1ull << h->hb[st.rb[q] - QT_L];
        return off;
    }
    uint32_t smem_reg_offset(const QtStage& st, int i) const {
        uint32_t off = 0;
        for (int q = 0; q < R; q++) if ((i >> q) & 1) off += qt_slot(1u << st.rb[q]);
        return off;
    }

    // thread-position expressions of a stage (lb = tile-local index with the register bits clear)
    std::string lb_expr(const QtStage& st) const {
        int tpos[QT_MAXM];
        for (int q = 0; q < M - R; q++) tpos[q] = st.tpos[q];
        return deposit_expr("tid", tpos, M - R, "unsigned");
    }
    std::string run_deposit_expr() const {
        int hpos[QT_MAXM];
        for (int i = 0; i < NH; i++) hpos[i] = h->hb[i];
        return deposit_expr("run", hpos, NH, "unsigned long long");
    }

    // The next tile of the CTA is fetched ASYNCHRONOUSLY by the TMA engine (cp.async.bulk, one
    // 512-byte run per copy, completion counted on an mbarrier) into the transposition buffer while
    // the last stage of the current tile computes and stores.  Per-thread loads (LDG or cp.async)
    // cannot do this: the LSU only keeps a few KB of requests in flight per SM, so a warp that
    // asks for its share of a 64 KB tile is stuck at issue for the whole transfer (measured:
    // arithmetic overlaps the stores and the L2 prefetch perfectly but ADDS to the load time).
    // Bulk copies occupy no LSU slot and no register; stage 0 of the next tile waits on the
    // mbarrier and reads its amplitudes out of the buffer like after any transposition.
    // run-index bit that splits the next tile into an EARLY half (landing buffer behind the
    // transposition buffer, fetched a whole tile ahead) and a LATE half (into the transposition
    // buffer once the last stage has emptied it): the run bit of stage 0's highest register bit,
    // so that which half a register comes from is known at generation time
    int early_bit() const { return stages[0].rb[R - 1] - QT_L; }
    static unsigned squeeze(unsigned k, int eb) { return (k & ((1u << eb) - 1u)) | ((k >> (eb + 1)) << eb); }

    void emit_issue_next() {
        const int eb = early_bit();
        o.f("#define QJ_RUNS %d\n#define QJ_HALF_RUNS %d\n", 1 << NH, 1 << (NH - 1));
        // k' enumerates the runs of one half; the full run index gets a 0 (early) / 1 (late) at bit eb.
        // A bulk copy is a warp-level instruction of the UNIFORM datapath (UBLKCP takes its addresses from
        // uniform registers).  With one copy per lane (kp = tid) ptxas has to serialise the lanes in a
        // "waterfall" loop -- ELECT / 3 x R2UR.BROADCAST / UBLKCP / BRA.U.ANY, a dependent chain of ~50 clk per
        // copy -- and with 64 copies per half only warps 0 and 1 of the CTA ran it while the others waited
        // at the next barrier (ncu source view of the heaviest benchmark sweep: 10 % of all warp samples in
        // those two loops, 21 % of the time of warps 0-1).  Now every warp issues its 2^(R-1) copies from
        // one elected lane with addresses derived from warp-uniform values only (the warp number through a
        // shuffle broadcast, the tile base, compile-time run numbers): straight-line UBLKCPs, no waterfall.
        // The run number is k = spread(w * per) | spread(j) (| the late bit): disjoint bits, and both the padded
        // buffer position and the HBM offset of a run are sums over the bits of k, so the part of warp w is computed
        // once (uniform registers) and the part of copy j is an immediate.
        const int per = 1 << (R - 1);       // HALF_RUNS / warps of the CTA
        auto spread = [&](unsigned kp) { return (kp & ((1u << eb) - 1u)) | ((kp >> eb) << (eb + 1)); };
        auto pad = [](unsigned k) { return 32u * k + k + (k >> 3) + (k >> 6); };
        auto run_off = [&](unsigned k) {
            unsigned long long off = 0;
            for (int i = 0; i < NH; i++) if ((k >> i) & 1u) off |= 1ull << h->hb[i];
            return off;
        };
        for (int late = 0; late < 2; late++) {
            o.f("QJ_DEV void %s(const unsigned tid, const unsigned long long nbase, QJ_C* QJ_RESTRICT psi, QJ_C* QJ_RESTRICT %s) {\n",
                late ? "qj_issue_next" : "qj_issue_early", late ? "buf" : "inb");
            o.f("    if (nbase == ~0ull) return;\n");
            if (uniform_issue) {
                o.f("    const unsigned wp = QJ_WARP_ID(tid) * %du;\n", per);
                o.f("    const unsigned kw = (wp & 0x%xu) | ((wp >> %d) << %d);\n", (1u << eb) - 1u, eb, eb + 1);
                if (late) o.f("    QJ_C* const sb = buf + (32u * kw + kw + (kw >> 3) + (kw >> 6));\n");
                else o.f("    QJ_C* const sb = inb + 32u * wp;\n");
                o.f("    const QJ_C* const gb = psi + (nbase + qj_run_offset(kw));\n");
                o.f("    if (QJ_ELECT(tid)) {\n");
                for (int j = 0; j < per; j++) {
                    const unsigned kj = spread((unsigned)j) | (late ? 1u << eb : 0u);
                    o.f("        %s(sb + %uu, gb + 0x%llxull);\n", late ? "QJ_BULK_COPY" : "QJ_BULK_COPY_EARLY", late ? pad(kj) : 32u * (unsigned)j,
                        run_off(kj));
                }
                // experiment (QBOT_B200_JIT_L2_LATE=1): the late half cannot land before the last stage has emptied the
                // transposition buffer; pull it into L2 together with the early half so that its copies find it there
                if (!late && l2_late)
                    for (int j = 0; j < per; j++) o.f("        QJ_L2_PREFETCH(gb + 0x%llxull);\n", run_off(spread((unsigned)j) | (1u << eb)));
                o.f("    }\n}\n\n");
                continue;
            }
            o.f("    for (unsigned kp = tid; kp < QJ_HALF_RUNS; kp += QJ_T) {\n");
            if (late) {
                o.f("        const unsigned k = (kp & 0x%xu) | ((kp >> %d) << %d) | 0x%xu;\n", (1u << eb) - 1u, eb, eb + 1, 1u << eb);
                o.f("        QJ_BULK_COPY(buf + (32u * k + k + (k >> 3) + (k >> 6)), psi + (nbase + qj_run_offset(k)));\n    }\n}\n\n");
            } else {
                o.f("        const unsigned k = (kp & 0x%xu) | ((kp >> %d) << %d);\n", (1u << eb) - 1u, eb, eb + 1);
                o.f("        QJ_BULK_COPY_EARLY(inb + 32u * kp, psi + (nbase + qj_run_offset(k)));\n    }\n}\n\n");
            }
        }
    }

    void emit_stage(int s) {
        const QtStage& st = stages[s];
        const bool first = s == 0, last = s + 1 == h->nstages;
        nm.resize(NR);
        for (int i = 0; i < NR; i++) nm[i] = "a" + std::to_string(i);
        o.f("QJ_DEV void qj_stage%d(const unsigned tid, const unsigned long long tbase, const unsigned long long nbase,\n"
            "                      const unsigned long long pfbase, const unsigned pre,\n"
            "                      QJ_C* QJ_RESTRICT psi, QJ_C* QJ_RESTRICT buf, QJ_POOL_PARAM) {\n", s);
        o.f("    const unsigned lb = %s;\n", lb_expr(st).c_str());
        o.f("    (void)lb; (void)tbase; (void)nbase; (void)pfbase; (void)pre; (void)psi; (void)buf;\n");
        if (first || last) {
            // HBM position of the thread: lane = low bits, run bits deposited at the tile-bit positions
            o.f("    const unsigned run = lb >> %d;\n", QT_L);
            o.f("    QJ_C* const gp = psi + (tbase + (unsigned long long)(lb & 31u) + (%s));\n", run_deposit_expr().c_str());
        }
        o.f("    QJ_C* const sp = buf + (lb + (lb >> 5) + (lb >> 8) + (lb >> 11));\n    (void)sp;\n");
        o.f("    QJ_C a0");
        for (int i = 1; i < NR; i++) o.f(", a%d", i);
        o.f(";\n");
        if (first) {
            // the tile is already in the buffer (bulk copies issued during the previous tile; pre - 1 is the
            // mbarrier phase parity to wait for), or comes straight from HBM (first tile of the CTA)
            // registers whose top register bit is 0 belong to the early half (landing buffer, runs stored
            // back to back in squeezed run order), the others to the late half (transposition buffer)
            const int eb = early_bit();
            // the early half has been in its landing buffer for most of a tile: read it before waiting for the late one
            o.f("    if (pre) {\n        QJ_MBAR_WAIT(0, pre - 1u);\n");
            o.f("        const unsigned kb = lb >> %d;\n", QT_L);
            o.f("        const QJ_C* const ip = buf + QJ_TILE_UNITS + 32u * ((kb & 0x%xu) | ((kb >> %d) << %d)) + (lb & 31u);\n",
                (1u << eb) - 1u, eb + 1, eb);
            for (int i = 0; i < NR / 2; i++) {
                unsigned rr = 0;
                for (int q = 0; q < R; q++) if ((i >> q) & 1) rr |= 1u << (st.rb[q] - QT_L);
                o.f("        a%d = ip[%u];\n", i, 32u * squeeze(rr, eb));
            }
            o.f("        QJ_MBAR_WAIT(1, pre - 1u);\n");
            for (int i = NR / 2; i < NR; i++) o.f("        a%d = sp[%u];\n", i, smem_reg_offset(st, i));
            o.f("    } else {\n");
            for (int i = 0; i < NR; i++) o.f("        a%d = QJ_LD(gp + 0x%llxull);\n", i, (unsigned long long)hbm_reg_offset(st, i));
            o.f("    }\n    QJ_PREFETCH(psi, pfbase, tid);\n");
        } else {
            for (int i = 0; i < NR; i++) o.f("    a%d = sp[%u];\n", i, smem_reg_offset(st, i));
        }
        // last stage: every thread of the CTA has read its amplitudes out of the buffer -> it is free
        // for the next tile's asynchronous copies (this barrier replaces the one before stage 0's stores)
        if (last) o.f("    QJ_ISSUE_NEXT(tid, nbase, psi, buf);\n");
        pend.clear();
        for (int x = 0; x < st.nops; x++) emit_op(ops[st.first_op + x]);
        // pending flips become a choice of store address: register i of a thread whose flag holds goes
        // where register i ^ (1 << T) would have gone.  Per target bit, registers are grouped by the set of
        // pending flips that select them (the XOR of those flags decides).
        if (!last && bank_aware) {
            // A flag that differs between the 8 lanes of one shared-memory wavefront moves some of them by
            // qt_slot(1 << bit) units = 1, 2 or 4 banks (mod 8) onto banks other lanes still use: every STS.128
            // of the stage would take two wavefronts, which costs more than the exchange.  Those flips happen now.
            for (size_t k = 0; k < pend.size();) {
                if (store_conflicts(st, pend[k])) { o.f("    // [flip:bank]\n"); flush(k); } else k++;
            }
        }
        std::vector<int> fmask(NR, 0);
        std::vector<std::string> fterms(NR);
        for (int T = 0; T < R; T++) {
            std::map<uint32_t, int> groups;         // signature (bit k <-> pend[k] selects the register) -> group id
            for (int i = 0; i < NR; i++) {
                uint32_t sig = 0;
                for (size_t k = 0; k < pend.size(); k++) if (pend[k].T == T && ((pend[k].sel >> i) & 1u)) sig |= 1u << k;
                if (!sig) continue;
                auto it = groups.find(sig);
                if (it == groups.end()) {
                    const int g = (int)groups.size();
                    it = groups.emplace(sig, g).first;
                    std::string fe;
                    for (size_t k = 0; k < pend.size(); k++) if ((sig >> k) & 1u) fe += (fe.empty() ? "" : " ^ ") + pend[k].var;
                    o.f("    // [flip:store%s]\n", __builtin_popcount(sig) > 1 ? "-multi" : sig_partial(pend, sig, all_sel()) ? "-partial" : "");
                    if (last) {
                        const unsigned long long hT = 1ull << h->hb[st.rb[T] - QT_L];
                        o.f("    const bool g%d_%d = %s; const unsigned long long e%d_%d_0 = g%d_%d ? 0x%llxull : 0ull, e%d_%d_1 = g%d_%d ? 0ull : 0x%llxull;\n",
                            T, g, fe.c_str(), T, g, T, g, hT, T, g, T, g, hT);
                    } else {
                        const unsigned sT = qt_slot(1u << st.rb[T]);
                        o.f("    const bool g%d_%d = %s; const unsigned e%d_%d_0 = g%d_%d ? %uu : 0u, e%d_%d_1 = g%d_%d ? 0u : %uu;\n", T, g, fe.c_str(), T,
                            g, T, g, sT, T, g, T, g, sT);
                    }
                }
                char b[48];
                snprintf(b, sizeof(b), " + e%d_%d_%d", T, it->second, (i >> T) & 1);
                fterms[i] += b;
                fmask[i] |= 1 << T;
            }
        }
        auto flag_terms = [&](int i) { return fterms[i]; };
        if (last) {
            if (h->scale != 1.0) {
                const uint32_t sidx = (uint32_t)npool_prog;      // the header scale rides behind the program's pool
                o.f("    { const double s_ = %s;\n", P(sidx).c_str());
                for (int i = 0; i < NR; i++) o.f("      %s.x *= s_; %s.y *= s_;\n", nm[i].c_str(), nm[i].c_str());
                o.f("    }\n");
            }
            for (int i = 0; i < NR; i++)
                o.f("    QJ_ST(gp + (0x%llxull%s), %s);\n", (unsigned long long)hbm_reg_offset(st, i & ~fmask[i]), flag_terms(i).c_str(), nm[i].c_str());
        } else {
            // stage 0 stores into the very slots this thread has just read (or, for the CTA's first tile,
            // into a buffer nobody has touched yet): no barrier needed before them
            for (int i = 0; i < NR; i++) o.f("    sp[%uu%s] = %s;\n", smem_reg_offset(st, i & ~fmask[i]), flag_terms(i).c_str(), nm[i].c_str());
        }
        pend.clear();
        o.f("}\n\n");
    }

    int npool_prog = 0;
};

}  // namespace

std::vector<double> qj_pool(const uint8_t* program) {
    const QtHeader* h = (const QtHeader*)program;
    const double* pool = (const double*)(program + h->pool_off);
    const size_t n = (h->total_bytes - h->pool_off) / sizeof(double);
    std::vector<double> out(pool, pool + n);
    out.push_back(h->scale);
    return out;
}

// resident CTAs per SM a kernel is compiled for: threads = 2^(M-R); 16 amplitudes per thread need
// ~128 registers (512 threads per SM); 32 amplitudes per thread spill at 168 registers (measured:
// 10.6 ms per 30-qubit sweep with 384 threads per SM against 7.85 ms with 256 threads at 255 registers)
int qj_default_ctas(int M, int R) {
    const int threads = 1 << (M - R);
    return R >= 5 ? 256 / threads : 512 / threads;
}

uint64_t qj_hash(const std::string& src) {
    uint64_t hsh = 1469598103934665603ull;
    for (unsigned char c : src) { hsh ^= c; hsh *= 1099511628211ull; }
    return hsh;
}

std::string qj_generate(const uint8_t* program, QjSourceInfo* info) {
    Gen g;
    g.h = (const QtHeader*)program;
    g.stages = (const QtStage*)(program + g.h->stages_off);
    g.ops = (const QtOp*)(program + g.h->ops_off);
    g.pool = (const double*)(program + g.h->pool_off);
    g.M = g.h->M;
    g.NH = g.M - QT_L;
    g.R = g.h->R ? g.h->R : QT_R;
    g.NR = 1 << g.R;
    g.T = 1 << (g.M - g.R);
    g.npool_prog = (int)((g.h->total_bytes - g.h->pool_off) / sizeof(double));
    g.lazy_x = getenv("QBOT_B200_EAGER_X") == nullptr;
    g.xor_signs = getenv("QBOT_B200_BRANCHY_SIGNS") == nullptr;        // A/B switch: conditional sign flips as branches + FP64 negations
    g.l2_late = getenv("QBOT_B200_JIT_L2_LATE") != nullptr && atoi(getenv("QBOT_B200_JIT_L2_LATE")) != 0;
    g.uniform_issue = getenv("QBOT_B200_LANE_ISSUE") == nullptr;        // A/B switch: one bulk copy per lane (ptxas waterfall loop)
    g.bank_aware = getenv("QBOT_B200_FLIP_BANK_AWARE") != nullptr;      // measured neutral (5 016 vs 5 023 gates/s): off by default
    const int npool = g.npool_prog + 1;      // + header scale
    Out& o = g.o;
    o.f("// generated by qbot_b200 qj_generate: M=%d R=%d stages=%d ops=%d gates=%d\n", g.M, g.R, (int)g.h->nstages, (int)g.h->nops, (int)g.h->ngates);
    o.f("#define QJ_M %d\n#define QJ_T %d\n#define QJ_NH %d\n#define QJ_NP %d\n#define QJ_NSTAGES %d\n", g.M, g.T, g.NH, npool, (int)g.h->nstages);
    o.f("#define QJ_TILE_UNITS %d\n#ifndef QJ_CTAS\n#define QJ_CTAS %d\n#endif\n", QT_TILE_UNITS(g.M), qj_default_ctas(g.M, g.R));
    o.f("#define QJ_NSTAGES_GT1 %d\n", g.h->nstages > 1 ? 1 : 0);
    o.f("QJ_PRELUDE\n\n");
    // tile base: the tile number's bits deposited into the positions outside the tile
    o.f("QJ_DEV unsigned long long qj_tile_base(const unsigned long long t) {\n    unsigned long long b = t << %d;\n", QT_L);
    for (int i = 0; i < g.NH; i++) {
        const int p = g.h->hbs[i];
        o.f("    b = ((b >> %d) << %d) | (b & 0x%llxull);\n", p, p + 1, (unsigned long long)((1ull << p) - 1ull));
    }
    o.f("    return b;\n}\n");
    {
        int hpos[QT_MAXM];
        for (int i = 0; i < g.NH; i++) hpos[i] = g.h->hb[i];
        o.f("QJ_DEV unsigned long long qj_run_offset(const unsigned k) { return %s; }\n\n",
            deposit_expr("k", hpos, g.NH, "unsigned long long").c_str());
    }
    g.emit_issue_next();
    for (int s = 0; s < g.h->nstages; s++) g.emit_stage(s);
    // after the first barrier of a tile nobody reads the landing buffer any more: the early half of
    // the NEXT tile starts to arrive while stages 1.. of this tile run
    o.f("#define QJ_RUN_STAGES(tid, tbase, nbase, pfbase, pre, psi, buf, P) \\\n");
    for (int s = 0; s < g.h->nstages; s++) {
        if (s == 1) o.f("    QJ_EXPECT_EARLY(tid, nbase); \\\n");
        if (s) o.f("    QJ_SYNC(); \\\n");
        if (s == 1) o.f("    QJ_ISSUE_EARLY(tid, nbase, psi, buf); \\\n");
        o.f("    qj_stage%d(tid, tbase, nbase, pfbase, pre, psi, buf, P); \\\n", s);
    }
    o.f("\n");
    o.f("#ifdef QJ_WANT_DISPATCH\nQJ_DEV void qj_stage(const int s, const unsigned tid, const unsigned long long tbase, const unsigned long long nbase, const unsigned pre, QJ_C* psi, QJ_C* buf, QJ_POOL_PARAM) {\n    switch (s) {\n");
    for (int s = 0; s < g.h->nstages; s++) o.f("        case %d: qj_stage%d(tid, tbase, nbase, ~0ull, pre, psi, buf, P); break;\n", s, s);
    o.f("        default: break;\n    }\n}\n#endif\n");
    if (info) {
        info->M = g.M;
        info->R = g.R;
        info->threads = g.T;
        info->npool = npool;
        info->nstages = g.h->nstages;
        info->tile_units = QT_TILE_UNITS(g.M);
    }
    return o.s;
}
