// Dense k-qubit blocks, k = 6..8, on the FP64 tensor cores (row f3 of SURVEY.md 8(f): the reference's
// qgates.genQFT, qbot/qgates.py:63-74, and any user-supplied 2^k x 2^k unitary reach the state through
// genGateForFullHilbertSpace + applyGate, qbot/qgates.py:161-182, 278-279).
//
// A dense 2^k block costs (8 * 2^k - 2) flops per amplitude against 32 bytes of HBM traffic: 16 flop/B
// at k = 6, 32 at k = 7, 64 at k = 8 -- above B200's FP64 ridge (37 TFLOP/s / 6.47 TB/s = 5.7 flop/B),
// so unlike every other kernel of the state path this one is bound by the FP64 pipes, and it is the
// one real dense contraction of the path: Y[2^k x C] = U[2^k x 2^k] * X[2^k x C] over all C = 2^(n-k)
// assignments of the other index bits.  It runs as a complex GEMM on mma.sync.m8n8k4.f64 (DMMA; no
// tcgen05 FP64 kind exists): four real products per complex one, accumulators in registers.
//
//   tile   = the 2^k target amplitudes x NC = 2^(12-k) columns (the lowest non-target, non-control
//            index bits, so a warp's accesses are whole runs) = 4096 amplitudes = 64 KB, staged in
//            shared memory as separate re / im planes (leading dimension = 4 mod 16 doubles: the
//            8-byte fragment loads of a half-warp hit 16 different bank pairs);
//   U      = split re / im planes in global memory (L2-resident: 64 KB .. 1 MB), streamed through a
//            double-buffered shared-memory chunk of 4 columns with cp.async (small enough for two
//            CTAs per SM: one computes while the other loads / stores its tile);
//   warps  = 8 per CTA, each owning 32 rows x 16 columns of the tile's output: 4 x 2 fragments of 8 x 8,
//            32 DMMA per 4-deep k step against 12 shared-memory fragment loads;
//   in place: a tile holds every amplitude its outputs depend on, so the CTA writes it back where it
//            came from (through the shared-memory planes, coalesced) -- no second buffer.
#include "qb_common.cuh"

namespace {

template <int K>
struct Cfg {
    static constexpr int D = 1 << K;
    static constexpr int NCB = QB_MMA_TILE_BITS - K;
    static constexpr int NC = 1 << NCB;
    static constexpr int LDX = NC + 4;
    static constexpr int KC = 4;                // columns of U per staged chunk = one DMMA k step
    static constexpr int LDA = KC;              // 4 doubles per row: fragment loads of a half-warp cover 16 consecutive doubles
    static constexpr int WC = NC / 16;          // warps along the columns
    static constexpr int WR = 8 / WC;           // warps along the rows (each 32 rows)
    static constexpr int X_DOUBLES = 2 * D * LDX;
    static constexpr int U_DOUBLES = 2 * 2 * D * LDA;      // [buffer][plane][row][LDA]
    static constexpr size_t SMEM = sizeof(double) * (X_DOUBLES + U_DOUBLES) + sizeof(uint64_t) * 64 + sizeof(uint32_t) * 64;
    static_assert(D / WR == 32, "a warp owns 32 rows");
};

__device__ __forceinline__ void dmma(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ double neg(const double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));     // sign bit, off the FP64 pipe
}

template <int K>
__global__ void __launch_bounds__(256) k_dense_mma(DenseMmaArgs a) {
    using C = Cfg<K>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* const xr = (double*)smem_raw;
    double* const xi = xr + C::D * C::LDX;
    double* const us = xi + C::D * C::LDX;
    uint64_t* const hi_off = (uint64_t*)(us + C::U_DOUBLES);        // tile-local bits 6..11 -> address offset
    uint32_t* const hi_sw = (uint32_t*)(hi_off + 64);               //                      -> smem word offset
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) {
        uint64_t o = 0; uint32_t s = 0;
        for (int b = 0; b < 6; b++) if ((tid >> b) & 1) { o |= 1ull << a.apos[6 + b]; s += a.sw[6 + b]; }
        hi_off[tid] = o; hi_sw[tid] = s;
    }
    uint64_t lo_off = 0; uint32_t lo_sw = 0;
    for (int b = 0; b < 6; b++) if ((tid >> b) & 1) { lo_off |= 1ull << a.apos[b]; lo_sw += a.sw[b]; }
    const int wr = warp / C::WC, wc = warp % C::WC;
    const int row0 = wr * 32, col0 = wc * 16;
    const int fr = lane >> 2, fc = lane & 3;                        // fragment row / column of this lane
    __syncthreads();

    for (uint64_t t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        uint64_t base = t;
        for (int i = 0; i < a.nins; i++) base = qb_insert_zero(base, a.ins[i]);
        base |= a.cmask;
        cplx* const g = a.psi + base + lo_off;
        // ---- tile -> shared-memory planes -------------------------------------------------------
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int h = (tid >> 6) + 4 * i;
            const cplx v = g[hi_off[h]];
            const uint32_t s = lo_sw + hi_sw[h];
            xr[s] = v.x; xi[s] = v.y;
        }
        // ---- Y = U X ---------------------------------------------------------------------------
        double cr[4][2][2], ci[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
        auto issue_chunk = [&](int kc) {
            // chunk kc = columns [4 kc, 4 kc + 4) of both planes: D rows x 32 bytes each = 2 x 16 B per row and plane
            double* const dst = us + (kc & 1) * (2 * C::D * C::LDA);
            for (int x = tid; x < 2 * C::D * 2; x += 256) {
                const int plane = x / (C::D * 2), r = (x >> 1) % C::D, q = x & 1;
                cp_async16(dst + plane * (C::D * C::LDA) + r * C::LDA + 2 * q,
                           (plane ? a.ui : a.ur) + (size_t)r * C::D + kc * C::KC + 2 * q);
            }
            cp_async_commit();
        };
        issue_chunk(0);
        constexpr int NCH = C::D / C::KC;
        for (int kc = 0; kc < NCH; kc++) {
            cp_async_wait_all();
            __syncthreads();                   // chunk kc is in, everybody is done with chunk kc - 1 (and, kc = 0, the planes are filled)
            if (kc + 1 < NCH) issue_chunk(kc + 1);
            const double* const ub = us + (kc & 1) * (2 * C::D * C::LDA);
#pragma unroll
            for (int k4 = 0; k4 < C::KC / 4; k4++) {
                double ar[4], ai[4], br[2], bi[2], nbi[2];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int o = (row0 + 8 * i + fr) * C::LDA + 4 * k4 + fc;
                    ar[i] = ub[o];
                    ai[i] = ub[C::D * C::LDA + o];
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int o = (kc * C::KC + 4 * k4 + fc) * C::LDX + col0 + 8 * j + fr;
                    br[j] = xr[o];
                    bi[j] = xi[o];
                    nbi[j] = neg(bi[j]);
                }
#pragma unroll
                for (int i = 0; i < 4; i++)             // 16 independent accumulator chains per round
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        dmma(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                        dmma(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
                    }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        dmma(cr[i][j][0], cr[i][j][1], ai[i], nbi[j]);
                        dmma(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
                    }
            }
        }
        __syncthreads();                       // every warp has read its last fragments of X
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int o = (row0 + 8 * i + fr) * C::LDX + col0 + 8 * j + 2 * fc;
                *(double2*)(xr + o) = make_double2(cr[i][j][0], cr[i][j][1]);      // o is even, the planes are 16-byte aligned
                *(double2*)(xi + o) = make_double2(ci[i][j][0], ci[i][j][1]);
            }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int h = (tid >> 6) + 4 * i;
            const uint32_t s = lo_sw + hi_sw[h];
            g[hi_off[h]] = make_double2(xr[s], xi[s]);
        }
        __syncthreads();                       // the planes are free for the next tile
    }
}

template <int K>
void launch(const LaunchCtx& c, const DenseMmaArgs& a) {
    using C = Cfg<K>;
    static bool configured[64] = {};
    int dev = 0;
    QB_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        QB_CUDA(cudaFuncSetAttribute(k_dense_mma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[dev & 63] = true;
    }
    int per_sm = 1;
    QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dense_mma<K>, 256, C::SMEM));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)c.sms * per_sm;
    const unsigned grid = (unsigned)(a.ntiles < cap ? a.ntiles : cap);
    k_dense_mma<K><<<grid, 256, C::SMEM, c.stream>>>(a);
    QB_CUDA(cudaGetLastError());
}

}  // namespace

void qb_launch_dense_mma(const LaunchCtx& c, int K, const DenseMmaArgs& a) {
    switch (K) {
        case 6: launch<6>(c, a); break;
        case 7: launch<7>(c, a); break;
        case 8: launch<8>(c, a); break;
        default: throw qb_error(-1, "qb_launch_dense_mma: K out of range");
    }
    if (c.launches) ++*c.launches;
}
