// qbot_b200 -- semantics of the fused-sweep program on one thread's registers.
// Compiled twice: by nvcc into the tile kernel (qb_tile.cu) and by g++ into the CPU plan
// emulator the tests use to check the planner (tests/csrc/plan_emulator.cpp).
#pragma once
#include "qb_plan.h"

#if defined(__CUDACC__)
typedef double2 qt_c;
#define QT_UNROLL _Pragma("unroll")
#else
struct qt_c { double x, y; };
#define QT_UNROLL
#endif

QT_HD qt_c qt_mk(double re, double im) { qt_c c; c.x = re; c.y = im; return c; }
QT_HD qt_c qt_mul(qt_c a, qt_c b) { return qt_mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QT_HD qt_c qt_fma(qt_c a, qt_c b, qt_c c) { return qt_mk(c.x + a.x * b.x - a.y * b.y, c.y + a.x * b.y + a.y * b.x); }

// ---- per-target-bit kernels (T = register-bit index of the target, compile time) --------------
template <int R, int T>
QT_HD void qt_h(qt_c (&a)[1 << R], unsigned regsel, double s) {
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++) {
        if (!((i >> T) & 1) && ((regsel >> i) & 1u)) {
            const int j = i | (1 << T);
            qt_c x = a[i], y = a[j];
            a[i] = qt_mk(s * x.x + s * y.x, s * x.y + s * y.y);
            a[j] = qt_mk(s * x.x - s * y.x, s * x.y - s * y.y);
        }
    }
}

template <int R, int T>
QT_HD void qt_x(qt_c (&a)[1 << R], unsigned regsel) {
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++) {
        if (!((i >> T) & 1) && ((regsel >> i) & 1u)) {
            const int j = i | (1 << T);
            qt_c x = a[i];
            a[i] = a[j];
            a[j] = x;
        }
    }
}

template <int R, int T>
QT_HD void qt_u2(qt_c (&a)[1 << R], unsigned regsel, const double* m) {
    const qt_c m00 = qt_mk(m[0], m[1]), m01 = qt_mk(m[2], m[3]), m10 = qt_mk(m[4], m[5]), m11 = qt_mk(m[6], m[7]);
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++) {
        if (!((i >> T) & 1) && ((regsel >> i) & 1u)) {
            const int j = i | (1 << T);
            qt_c x = a[i], y = a[j];
            a[i] = qt_fma(m01, y, qt_mul(m00, x));
            a[j] = qt_fma(m11, y, qt_mul(m10, x));
        }
    }
}

template <int R, int T0, int T1>          // T0 = most significant matrix bit
QT_HD void qt_u4(qt_c (&a)[1 << R], unsigned regsel, const double* m) {
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++) {
        if (!((i >> T0) & 1) && !((i >> T1) & 1) && ((regsel >> i) & 1u)) {
            const int idx[4] = {i, i | (1 << T1), i | (1 << T0), i | (1 << T0) | (1 << T1)};
            qt_c x[4] = {a[idx[0]], a[idx[1]], a[idx[2]], a[idx[3]]};
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

template <int R>
QT_HD void qt_cdiag_reg(qt_c (&a)[1 << R], unsigned regsel, int t, const double* d) {
    const qt_c d0 = qt_mk(d[0], d[1]), d1 = qt_mk(d[2], d[3]);
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++)
        if ((regsel >> i) & 1u) a[i] = qt_mul(((i >> t) & 1) ? d1 : d0, a[i]);
}

template <int R>
QT_HD void qt_scale(qt_c (&a)[1 << R], unsigned regsel, qt_c f) {
    QT_UNROLL
    for (int i = 0; i < (1 << R); i++)
        if ((regsel >> i) & 1u) a[i] = qt_mul(f, a[i]);
}

// ---- one op on one thread ------------------------------------------------------------------------
// lbase: the thread's tile-local index with its register bits cleared; tbase: index of the tile's
// first amplitude (bits outside the tile; tile bits are zero)
template <int R>
QT_HD void qt_apply_op(qt_c (&a)[1 << R], const QtOp& op, const double* pool, uint32_t lbase, uint64_t tbase) {
    if ((tbase & op.gmask) != op.gval) return;
    if ((lbase & op.lmask) != op.lval) return;
    const double* p = pool + op.pool;
    const unsigned rs = op.regsel;
    switch (op.type) {
        case QT_OP_H:
            switch (op.t0) {
                case 0: qt_h<R, 0>(a, rs, p[0]); break;
                case 1: qt_h<R, 1>(a, rs, p[0]); break;
                case 2: qt_h<R, 2>(a, rs, p[0]); break;
                default: if (R > 3) qt_h<R, (R > 3 ? 3 : 0)>(a, rs, p[0]); break;
            }
            break;
        case QT_OP_X:
            switch (op.t0) {
                case 0: qt_x<R, 0>(a, rs); break;
                case 1: qt_x<R, 1>(a, rs); break;
                case 2: qt_x<R, 2>(a, rs); break;
                default: if (R > 3) qt_x<R, (R > 3 ? 3 : 0)>(a, rs); break;
            }
            break;
        case QT_OP_U2:
            switch (op.t0) {
                case 0: qt_u2<R, 0>(a, rs, p); break;
                case 1: qt_u2<R, 1>(a, rs, p); break;
                case 2: qt_u2<R, 2>(a, rs, p); break;
                default: if (R > 3) qt_u2<R, (R > 3 ? 3 : 0)>(a, rs, p); break;
            }
            break;
        case QT_OP_U4: {
            const int key = op.t0 * 4 + op.t1;
            switch (key) {
                case 1: qt_u4<R, 0, 1>(a, rs, p); break;
                case 2: qt_u4<R, 0, 2>(a, rs, p); break;
                case 4: qt_u4<R, 1, 0>(a, rs, p); break;
                case 6: qt_u4<R, 1, 2>(a, rs, p); break;
                case 8: qt_u4<R, 2, 0>(a, rs, p); break;
                case 9: qt_u4<R, 2, 1>(a, rs, p); break;
                case 3: if (R > 3) qt_u4<R, 0, (R > 3 ? 3 : 1)>(a, rs, p); break;
                case 7: if (R > 3) qt_u4<R, 1, (R > 3 ? 3 : 0)>(a, rs, p); break;
                case 11: if (R > 3) qt_u4<R, 2, (R > 3 ? 3 : 0)>(a, rs, p); break;
                case 12: if (R > 3) qt_u4<R, (R > 3 ? 3 : 1), 0>(a, rs, p); break;
                case 13: if (R > 3) qt_u4<R, (R > 3 ? 3 : 0), 1>(a, rs, p); break;
                case 14: if (R > 3) qt_u4<R, (R > 3 ? 3 : 0), 2>(a, rs, p); break;
                default: break;
            }
            break;
        }
        case QT_OP_CDIAG: {
            if (op.t1 == QT_LOC_REG) qt_cdiag_reg<R>(a, rs, op.t0, p);
            else {
                const int b = op.t1 == QT_LOC_LOCAL ? (int)((lbase >> op.t0) & 1u) : (int)((tbase >> op.t0) & 1ull);
                qt_scale<R>(a, rs, qt_mk(p[2 * b], p[2 * b + 1]));
            }
            break;
        }
        case QT_OP_PHASE: {
            qt_c common = qt_mk(1.0, 0.0);
            bool have = false;
            for (int e = 0; e < op.nent; e++) {
                const double* q = p + 5 * e;
                const int code = (int)q[0], loc = code & 0xff, pos = code >> 8;
                if (loc == QT_LOC_REG) continue;
                const int b = loc == QT_LOC_LOCAL ? (int)((lbase >> pos) & 1u) : (int)((tbase >> pos) & 1ull);
                const qt_c f = qt_mk(q[1 + 2 * b], q[2 + 2 * b]);
                common = have ? qt_mul(common, f) : f;
                have = true;
            }
            if (have) qt_scale<R>(a, 0xffffu, common);
            for (int e = 0; e < op.nent; e++) {
                const double* q = p + 5 * e;
                const int code = (int)q[0], loc = code & 0xff, pos = code >> 8;
                if (loc == QT_LOC_REG) qt_cdiag_reg<R>(a, 0xffffu, pos, q + 1);
            }
            break;
        }
        default: break;
    }
}

// tile-local index (register bits cleared) of thread `tid` in a stage
template <int R>
QT_HD uint32_t qt_thread_lbase(const QtStage& st, uint32_t tid) {
    uint32_t l = 0;
    QT_UNROLL
    for (int q = 0; q < QT_M - R; q++) l |= ((tid >> q) & 1u) << st.tpos[q];
    return l;
}

template <int R>
QT_HD uint32_t qt_reg_offset(const QtStage& st, int i) {
    uint32_t o = 0;
    QT_UNROLL
    for (int q = 0; q < R; q++) o |= (uint32_t)((i >> q) & 1) << st.rb[q];
    return o;
}
