// TEST INFRASTRUCTURE: CPU definitions of the macros the sweep specialiser's generated text is
// written over (qbot_b200/csrc/qb_jit.h), plus the per-step harness that runs the stage
// functions over every tile and thread the way the CUDA kernel schedules them (one stage for all
// threads of a tile, then the next -- the barriers of the kernel).
#pragma once
#include <cstring>
#include <vector>
struct QjC { double x, y; };
#define QJ_C QjC
#define QJ_DEV static inline
#define QJ_LD(p) (*(p))
#define QJ_ST(p, v) (*(p) = (v))
#define QJ_P(i) (P[i])
#define QJ_POOL_PARAM const double* P
#define QJ_RESTRICT
static inline double qj_xsign1(double v, unsigned s) { unsigned long long b; memcpy(&b, &v, 8); b ^= (unsigned long long)s << 32; memcpy(&v, &b, 8); return v; }
#define QJ_XSIGN(a, s) do { (a).x = qj_xsign1((a).x, (s)); (a).y = qj_xsign1((a).y, (s)); } while (0)
#define QJ_SYNC()
#define QJ_WARP_ID(tid) ((tid) >> 5)
#define QJ_ELECT(tid) (((tid) & 31u) == 0u)
#define QJ_BULK_COPY(sdst, gsrc) memcpy((sdst), (gsrc), 512)
#define QJ_BULK_COPY_EARLY(sdst, gsrc) memcpy((sdst), (gsrc), 512)
#define QJ_ISSUE_EARLY(tid, nbase, psi, buf)     /* the harness runs qj_issue_early at the end of the tile */
#define QJ_ASYNC_WAIT(parity)
#define QJ_MBAR_WAIT(which, parity)
#define QJ_EXPECT_EARLY(tid, nbase)
#define QJ_L2_PREFETCH(gsrc)
#define QJ_ISSUE_NEXT(tid, nbase, psi, buf)      /* the harness runs qj_issue_next after the last stage (the kernel's barrier) */
#define QJ_PREFETCH(psi, nbase, tid)
#define QJ_PRELUDE
#define QJ_WANT_DISPATCH

#define QJ_CPU_HARNESS(NAME)                                                                         \
    extern "C" void NAME(QjC* psi, int nbits, const double* pool) {                                  \
        const unsigned long long ntiles = 1ull << (nbits - QJ_M);                                    \
        std::vector<QjC> buf(QJ_TILE_UNITS + 32 * QJ_HALF_RUNS);   /* transposition + landing buffer */ \
        /* one "CTA" walks all tiles: the first tile comes from HBM, every later one through the    \
           asynchronous copies its predecessor's last stage issued into the buffer */               \
        for (unsigned long long t = 0; t < ntiles; t++) {                                            \
            const unsigned long long tbase = qj_tile_base(t);                                        \
            const unsigned long long nbase = t + 1 < ntiles ? qj_tile_base(t + 1) : ~0ull;           \
            for (int s = 0; s < QJ_NSTAGES; s++)                                                     \
                for (unsigned tid = 0; tid < QJ_T; tid++) qj_stage(s, tid, tbase, nbase, t > 0, psi, buf.data(), pool); \
            for (unsigned tid = 0; tid < QJ_T; tid++) qj_issue_next(tid, nbase, psi, buf.data());    \
            for (unsigned tid = 0; tid < QJ_T; tid++) qj_issue_early(tid, nbase, psi, buf.data() + QJ_TILE_UNITS); \
        }                                                                                            \
    }
