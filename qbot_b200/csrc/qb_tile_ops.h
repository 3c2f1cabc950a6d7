// qbot_b200 -- semantics of the fused-sweep program on one thread's 16 registers.
// Compiled twice: by nvcc into the tile kernel (qb_tile.cu) and by g++ into the CPU plan
// emulator the tests use to check the planner (tests/csrc/plan_emulator.cpp).
#pragma once
#include "qb_plan.h"
#include <string.h>

#if defined(__CUDACC__)
typedef double2 qt_c;
#define QT_UNROLL _Pragma("unroll")
#else
struct qt_c { double x, y; };
#define QT_UNROLL
#endif

#define QT_NR (1 << QT_R)

QT_HD qt_c qt_mk(double re, double im) { qt_c c; c.x = re; c.y = im; return c; }
QT_HD qt_c qt_mul(qt_c a, qt_c b) { return qt_mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QT_HD qt_c qt_fma(qt_c a, qt_c b, qt_c c) { return qt_mk(c.x + a.x * b.x - a.y * b.y, c.y + a.x * b.y + a.y * b.x); }

// ---- per-target-bit kernels (T = register-bit index of the target, compile time; ALL = the
//      predicate holds for all 16 registers, so no per-pair test is needed) --------------------
template <int T, bool ALL>
QT_HD void qt_h(qt_c (&a)[QT_NR], unsigned regsel) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++) {
        if (!((i >> T) & 1) && (ALL || ((regsel >> i) & 1u))) {
            const int j = i | (1 << T);
            const qt_c x = a[i], y = a[j];
            a[i] = qt_mk(x.x + y.x, x.y + y.y);
            a[j] = qt_mk(x.x - y.x, x.y - y.y);
        }
    }
}

template <int T, bool ALL>
QT_HD void qt_x(qt_c (&a)[QT_NR], unsigned regsel) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++) {
        if (!((i >> T) & 1) && (ALL || ((regsel >> i) & 1u))) {
            const int j = i | (1 << T);
            const qt_c x = a[i];
            a[i] = a[j];
            a[j] = x;
        }
    }
}

template <int T, bool ALL>
QT_HD void qt_u2(qt_c (&a)[QT_NR], unsigned regsel, const double* m) {
    const qt_c m00 = qt_mk(m[0], m[1]), m01 = qt_mk(m[2], m[3]), m10 = qt_mk(m[4], m[5]), m11 = qt_mk(m[6], m[7]);
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++) {
        if (!((i >> T) & 1) && (ALL || ((regsel >> i) & 1u))) {
            const int j = i | (1 << T);
            const qt_c x = a[i], y = a[j];
            a[i] = qt_fma(m01, y, qt_mul(m00, x));
            a[j] = qt_fma(m11, y, qt_mul(m10, x));
        }
    }
}

template <int T0, int T1>          // T0 = most significant matrix bit
QT_HD void qt_u4(qt_c (&a)[QT_NR], unsigned regsel, const double* m) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++) {
        if (!((i >> T0) & 1) && !((i >> T1) & 1) && ((regsel >> i) & 1u)) {
            const int idx[4] = {i, i | (1 << T1), i | (1 << T0), i | (1 << T0) | (1 << T1)};
            const qt_c x[4] = {a[idx[0]], a[idx[1]], a[idx[2]], a[idx[3]]};
            QT_UNROLL
            for (int r = 0; r < 4; r++) {
                qt_c acc = qt_mul(qt_mk(m[8 * r], m[8 * r + 1]), x[0]);
                QT_UNROLL
                for (int c = 1; c < 4; c++) acc = qt_fma(qt_mk(m[8 * r + 2 * c], m[8 * r + 2 * c + 1]), x[c], acc);
                a[idx[r]] = acc;
            }
        }
    }
}

// a[i] *= (bit t of i ? d1 : d0) on the selected registers
template <int T, bool ALL>
QT_HD void qt_diag_reg(qt_c (&a)[QT_NR], unsigned regsel, qt_c d0, qt_c d1) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++)
        if (ALL || ((regsel >> i) & 1u)) a[i] = qt_mul(((i >> T) & 1) ? d1 : d0, a[i]);
}

template <bool ALL>
QT_HD void qt_diag_reg_dyn(qt_c (&a)[QT_NR], unsigned regsel, int t, qt_c d0, qt_c d1) {
    switch (t) {
        case 0: qt_diag_reg<0, ALL>(a, regsel, d0, d1); break;
        case 1: qt_diag_reg<1, ALL>(a, regsel, d0, d1); break;
        case 2: qt_diag_reg<2, ALL>(a, regsel, d0, d1); break;
        default: qt_diag_reg<3, ALL>(a, regsel, d0, d1); break;
    }
}

template <bool ALL>
QT_HD void qt_scale(qt_c (&a)[QT_NR], unsigned regsel, qt_c f) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++)
        if (ALL || ((regsel >> i) & 1u)) a[i] = qt_mul(f, a[i]);
}

QT_HD void qt_scale_real(qt_c (&a)[QT_NR], double s) {
    QT_UNROLL
    for (int i = 0; i < QT_NR; i++) a[i] = qt_mk(a[i].x * s, a[i].y * s);
}

QT_HD int64_t qt_code_of(const double* q) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(q[0]);
#else
    int64_t c;
    memcpy(&c, q, sizeof(c));
    return c;
#endif
}

// ---- predicates ----------------------------------------------------------------------------------
QT_HD bool qt_op_local_ok(const QtOp& op, uint32_t lbase) { return (lbase & op.lmask) == op.lval; }
QT_HD bool qt_op_global_ok(const QtOp& op, uint64_t tbase) { return (tbase & op.gmask) == op.gval; }

// ---- one op on one thread (predicates already checked) -----------------------------------------
// lbase: the thread's tile-local index with its register bits cleared; tbase: index of the tile's
// first amplitude (bits outside the tile; tile bits are zero)
template <bool ALL>
QT_HD void qt_apply_op_sel(qt_c (&a)[QT_NR], const QtOp& op, const double* pool, uint32_t lbase, uint64_t tbase) {
    const double* p = pool + op.pool;
    const unsigned rs = op.regsel;
    switch (op.type) {
        case QT_OP_H:
            switch (op.t0) {
                case 0: qt_h<0, ALL>(a, rs); break;
                case 1: qt_h<1, ALL>(a, rs); break;
                case 2: qt_h<2, ALL>(a, rs); break;
                default: qt_h<3, ALL>(a, rs); break;
            }
            break;
        case QT_OP_X:
            switch (op.t0) {
                case 0: qt_x<0, ALL>(a, rs); break;
                case 1: qt_x<1, ALL>(a, rs); break;
                case 2: qt_x<2, ALL>(a, rs); break;
                default: qt_x<3, ALL>(a, rs); break;
            }
            break;
        case QT_OP_U2:
            switch (op.t0) {
                case 0: qt_u2<0, ALL>(a, rs, p); break;
                case 1: qt_u2<1, ALL>(a, rs, p); break;
                case 2: qt_u2<2, ALL>(a, rs, p); break;
                default: qt_u2<3, ALL>(a, rs, p); break;
            }
            break;
        case QT_OP_U4: {
            const unsigned r4 = ALL ? 0xffffu : rs;
            switch (op.t0 * 4 + op.t1) {
                case 1: qt_u4<0, 1>(a, r4, p); break;
                case 2: qt_u4<0, 2>(a, r4, p); break;
                case 3: qt_u4<0, 3>(a, r4, p); break;
                case 4: qt_u4<1, 0>(a, r4, p); break;
                case 6: qt_u4<1, 2>(a, r4, p); break;
                case 7: qt_u4<1, 3>(a, r4, p); break;
                case 8: qt_u4<2, 0>(a, r4, p); break;
                case 9: qt_u4<2, 1>(a, r4, p); break;
                case 11: qt_u4<2, 3>(a, r4, p); break;
                case 12: qt_u4<3, 0>(a, r4, p); break;
                case 13: qt_u4<3, 1>(a, r4, p); break;
                case 14: qt_u4<3, 2>(a, r4, p); break;
                default: break;
            }
            break;
        }
        case QT_OP_CDIAG: {
            if (op.t1 == QT_LOC_REG) qt_diag_reg_dyn<ALL>(a, rs, op.t0, qt_mk(p[0], p[1]), qt_mk(p[2], p[3]));
            else {
                const int b = op.t1 == QT_LOC_LOCAL ? (int)((lbase >> op.t0) & 1u) : (int)((tbase >> op.t0) & 1ull);
                qt_scale<ALL>(a, rs, qt_mk(p[2 * b], p[2 * b + 1]));
            }
            break;
        }
        case QT_OP_PHASE: {
            // common = product of the factors selected by thread-local / out-of-tile bits (and
            // constants); the first register-bit entry is folded into it, further ones applied alone
            qt_c common = qt_mk(1.0, 0.0);
            bool have = false;
            int first_reg = -1;
            for (int e = 0; e < op.nent; e++) {
                const double* q = p + 5 * e;
                const int code = (int)qt_code_of(q), loc = code & 0xff, pos = code >> 8;
                if (loc == QT_LOC_REG) { if (first_reg < 0) first_reg = e; continue; }
                const int b = loc == QT_LOC_LOCAL ? (int)((lbase >> pos) & 1u)
                            : loc == QT_LOC_GLOBAL ? (int)((tbase >> pos) & 1ull) : 0;
                const qt_c f = qt_mk(q[1 + 2 * b], q[2 + 2 * b]);
                common = have ? qt_mul(common, f) : f;
                have = true;
            }
            if (first_reg < 0) {
                if (have) qt_scale<true>(a, 0xffffu, common);
            } else {
                const double* q = p + 5 * first_reg;
                const int pos = (int)qt_code_of(q) >> 8;
                qt_c d0 = qt_mk(q[1], q[2]), d1 = qt_mk(q[3], q[4]);
                if (have) { d0 = qt_mul(common, d0); d1 = qt_mul(common, d1); }
                qt_diag_reg_dyn<true>(a, 0xffffu, pos, d0, d1);
                for (int e = first_reg + 1; e < op.nent; e++) {
                    const double* q2 = p + 5 * e;
                    const int code = (int)qt_code_of(q2);
                    if ((code & 0xff) == QT_LOC_REG)
                        qt_diag_reg_dyn<true>(a, 0xffffu, code >> 8, qt_mk(q2[1], q2[2]), qt_mk(q2[3], q2[4]));
                }
            }
            break;
        }
        default: break;
    }
}

QT_HD void qt_apply_op(qt_c (&a)[QT_NR], const QtOp& op, const double* pool, uint32_t lbase, uint64_t tbase) {
    if (op.flags & QT_FLAG_ALLREG) qt_apply_op_sel<true>(a, op, pool, lbase, tbase);
    else qt_apply_op_sel<false>(a, op, pool, lbase, tbase);
}

// tile-local index (register bits cleared) of thread `tid` in a stage of an M-bit tile
QT_HD uint32_t qt_thread_lbase(const QtStage& st, uint32_t tid, int M) {
    uint32_t l = 0;
    for (int q = 0; q < M - QT_R; q++) l |= ((tid >> q) & 1u) << st.tpos[q];
    return l;
}

QT_HD uint32_t qt_reg_offset(const QtStage& st, int i) {
    uint32_t o = 0;
    QT_UNROLL
    for (int q = 0; q < QT_R; q++) o |= (uint32_t)((i >> q) & 1) << st.rb[q];
    return o;
}
