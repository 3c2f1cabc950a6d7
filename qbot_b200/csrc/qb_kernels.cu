// qbot_b200 -- per-operation CUDA kernels for the state path (sm_100a).
//
// Every kernel here is HBM-bound strided complex128 work: one amplitude is exactly one
// 128-bit word, so all loads/stores are 128-bit; work items are enumerated so that a warp's
// footprint is contiguous (zero-insertion addressing keeps the low index bits in the lane
// id), grids are sized in multiples of the SM count with 64-bit grid-stride loops.
// The fused multi-gate tile kernel lives in qb_tile.cu; these are the general-case and
// structure (trace / scatter / mix / reduce) kernels.
#include "qb_common.cuh"

static inline int grid_for(const LaunchCtx& c, uint64_t work, int threads, int per_sm = 8) {
    uint64_t blocks = (work + threads - 1) / threads;
    uint64_t cap = (uint64_t)c.sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
#define COUNT_LAUNCH(c) do { if ((c).launches) ++*(c).launches; } while (0)

// ---------------------------------------------------------------------------------------------
// initialisation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_basis(cplx* d, uint64_t per_branch, uint64_t total, uint64_t index) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint64_t l = i % per_branch;
        d[i] = make_double2(l == index ? 1.0 : 0.0, 0.0);
    }
}

void qb_launch_fill_basis(const LaunchCtx& c, cplx* d, uint64_t per_branch, int64_t nbranch, uint64_t index) {
    uint64_t total = per_branch * (uint64_t)nbranch;
    k_fill_basis<<<grid_for(c, total, 256), 256, 0, c.stream>>>(d, per_branch, total, index);
    COUNT_LAUNCH(c);
}

// rho = diag(values): one write of the density matrix
__global__ void __launch_bounds__(256) k_fill_diag(cplx* d, int nq, const double* __restrict__ values) {
    const uint64_t N = 1ull << nq, total = N * N;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t r = i >> nq, c = i & (N - 1);
        __stcs(d + i, make_double2(r == c ? values[r] : 0.0, 0.0));
    }
}

void qb_launch_fill_diag(const LaunchCtx& c, cplx* d, int nq, const double* values_dev) {
    k_fill_diag<<<grid_for(c, 1ull << (2 * nq), 256), 256, 0, c.stream>>>(d, nq, values_dev);
    COUNT_LAUNCH(c);
}

// every branch of dst <- src (fan-out before a batched gate): src is read once from HBM and then from
// L2, the write is the traffic; one launch instead of nbranch device-to-device copies
__global__ void __launch_bounds__(256) k_broadcast(const cplx* __restrict__ src, cplx* dst, uint64_t per, uint64_t total) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        __stcs(dst + i, src[i & (per - 1)]);
}

void qb_launch_broadcast(const LaunchCtx& c, const cplx* src, cplx* dst, uint64_t per, int64_t nbranch) {
    const uint64_t total = per * (uint64_t)nbranch;
    k_broadcast<<<grid_for(c, total, 256), 256, 0, c.stream>>>(src, dst, per, total);
    COUNT_LAUNCH(c);
}

// ket: psi[i] = prod_q v_q[bit_q(i)]   dm: rho[r][c] = prod_q D_q[r_q][c_q]   (qubit 0 first; the
// reference builds these with a kron chain on the host, density.py:7-24).
// The qubits are cut into groups of up to 8 (ket) / 4 (density matrix) consecutive qubits; a
// block builds the 256-entry factor table of every group of ITS branch in shared memory (each
// entry the left-to-right product of the group's factors) and then writes a 2^16-amplitude
// chunk with one complex multiply per group and amplitude -- the kernel is a pure HBM write
// instead of nq dependent loads + multiplies per amplitude.
#define IP_MAXG 5
#define IP_CHUNK_BITS 16
__global__ void __launch_bounds__(256) k_init_product(cplx* d, int kind, int nq, int nbits, uint64_t chunks_per_branch,
                                                      const cplx* __restrict__ vecs, int per_branch_vecs) {
    __shared__ cplx T[IP_MAXG][256];
    const int per = kind == 0 ? 2 : 4;
    const int gq = kind == 0 ? 8 : 4;                       // qubits per group
    const int ng = (nq + gq - 1) / gq;
    const uint64_t b = blockIdx.x / chunks_per_branch, chunk = blockIdx.x % chunks_per_branch;
    const cplx* v = vecs + (per_branch_vecs ? b * (uint64_t)nq * per : 0);
    for (int g = 0; g < ng; g++) {
        const int q0 = g * gq, len = min(gq, nq - q0);
        const int entries = kind == 0 ? (1 << len) : (1 << (2 * len));
        for (int e = threadIdx.x; e < entries; e += blockDim.x) {
            cplx acc = make_double2(1.0, 0.0);
            for (int j = 0; j < len; j++) {
                int sel;
                if (kind == 0) sel = (e >> (len - 1 - j)) & 1;
                else sel = (((e >> (2 * len - 1 - j)) & 1) << 1) | ((e >> (len - 1 - j)) & 1);
                const cplx f = v[(q0 + j) * per + sel];
                acc = j == 0 ? f : qb_cmul(acc, f);
            }
            T[g][e] = acc;
        }
    }
    __syncthreads();
    const uint64_t per_branch = 1ull << nbits;
    const uint64_t lo = chunk << IP_CHUNK_BITS;
    const uint64_t hi = per_branch < lo + (1ull << IP_CHUNK_BITS) ? per_branch : lo + (1ull << IP_CHUNK_BITS);
    cplx* out = d + b * per_branch;
    for (uint64_t l = lo + threadIdx.x; l < hi; l += blockDim.x) {
        cplx acc;
        for (int g = 0; g < ng; g++) {
            const int q0 = g * gq, len = min(gq, nq - q0);
            unsigned e;
            if (kind == 0) e = (unsigned)(l >> (nq - q0 - len)) & ((1u << len) - 1u);
            else e = (((unsigned)(l >> (2 * nq - q0 - len)) & ((1u << len) - 1u)) << len) | ((unsigned)(l >> (nq - q0 - len)) & ((1u << len) - 1u));
            acc = g == 0 ? T[0][e] : qb_cmul(acc, T[g][e]);
        }
        __stcs(out + l, acc);
    }
}

void qb_launch_init_product(const LaunchCtx& c, cplx* d, int kind, int nq, int64_t nbranch, const cplx* vecs_dev, int per_branch) {
    const int nbits = kind == 0 ? nq : 2 * nq;
    QB_REQUIRE(nq <= (kind == 0 ? 8 : 4) * IP_MAXG, "init_product: too many qubits");
    const uint64_t per = 1ull << nbits;
    const uint64_t chunks = per >> IP_CHUNK_BITS ? per >> IP_CHUNK_BITS : 1;
    const uint64_t blocks = chunks * (uint64_t)nbranch;
    QB_REQUIRE(blocks < (1ull << 31), "init_product: state too large");
    k_init_product<<<(unsigned)blocks, 256, 0, c.stream>>>(d, kind, nq, nbits, chunks, vecs_dev, per_branch);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// dense k-qubit gate, in place, registers (K <= 5), optional control mask
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ uint64_t sub_offset(int j, const uint64_t (&tbm)[K]) {
    uint64_t o = 0;
#pragma unroll
    for (int b = 0; b < K; b++)
        if ((j >> (K - 1 - b)) & 1) o |= tbm[b];
    return o;
}

template <int K>
__global__ void __launch_bounds__(K >= 5 ? 128 : 256) k_dense(DenseArgs a) {
    constexpr int D = 1 << K;
    __shared__ cplx sm[D * D];
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sm[i] = (K <= 2) ? a.inl[i] : a.mat[i];
    __syncthreads();
    uint64_t tbm[K];
#pragma unroll
    for (int b = 0; b < K; b++) tbm[b] = 1ull << a.tb[b];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < a.nwork; w += stride) {
        uint64_t base = w;
        for (int i = 0; i < a.nins; i++) base = qb_insert_zero(base, a.ins[i]);
        base |= a.cmask;
        cplx x[D];
#pragma unroll
        for (int j = 0; j < D; j++) x[j] = a.psi[base | sub_offset<K>(j, tbm)];
#pragma unroll(K <= 2 ? D : 1)
        for (int i = 0; i < D; i++) {
            cplx acc = qb_cmul(sm[i * D], x[0]);
#pragma unroll
            for (int j = 1; j < D; j++) acc = qb_cfma(sm[i * D + j], x[j], acc);
            a.psi[base | sub_offset<K>(i, tbm)] = acc;
        }
    }
}

void qb_launch_dense(const LaunchCtx& c, int K, const DenseArgs& a) {
    int threads = K >= 5 ? 128 : 256;
    int grid = grid_for(c, a.nwork, threads);
    switch (K) {
        case 1: k_dense<1><<<grid, threads, 0, c.stream>>>(a); break;
        case 2: k_dense<2><<<grid, threads, 0, c.stream>>>(a); break;
        case 3: k_dense<3><<<grid, threads, 0, c.stream>>>(a); break;
        case 4: k_dense<4><<<grid, threads, 0, c.stream>>>(a); break;
        case 5: k_dense<5><<<grid, threads, 0, c.stream>>>(a); break;
        default: throw qb_error(-1, "qb_launch_dense: K out of range");
    }
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// diagonal gate: psi[idx] *= d[bits(idx)] on the control-satisfying subspace
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag(DiagArgs a) {
    extern __shared__ cplx sd[];
    int D = 1 << a.k;
    for (int i = threadIdx.x; i < D; i += blockDim.x) sd[i] = (a.k <= 2) ? a.inl[i] : a.diag[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < a.nwork; w += stride) {
        uint64_t idx = w;
        for (int i = 0; i < a.nins; i++) idx = qb_insert_zero(idx, a.ins[i]);
        idx |= a.cmask;
        int j = 0;
        for (int b = 0; b < a.k; b++) j |= (int)((idx >> a.tb[b]) & 1) << (a.k - 1 - b);
        a.psi[idx] = qb_cmul(sd[j], a.psi[idx]);
    }
}

void qb_launch_diag(const LaunchCtx& c, const DiagArgs& a) {
    size_t smem = sizeof(cplx) << a.k;
    k_diag<<<grid_for(c, a.nwork, 256), 256, smem, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// qubit transposition: exchange amplitudes whose bits (lo, hi) are (1,0) <-> (0,1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_swap(cplx* d, uint64_t nwork, int lo, int hi) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwork; w += stride) {
        uint64_t base = qb_insert_zero(qb_insert_zero(w, lo), hi);
        uint64_t i1 = base | (1ull << lo), i2 = base | (1ull << hi);
        cplx t = d[i1];
        d[i1] = d[i2];
        d[i2] = t;
    }
}

void qb_launch_swap(const LaunchCtx& c, cplx* d, uint64_t total, int lo, int hi) {
    uint64_t nwork = total >> 2;
    k_swap<<<grid_for(c, nwork, 256), 256, 0, c.stream>>>(d, nwork, lo, hi);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// large dense gate (K > 5): out-of-place, one output amplitude per thread
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_big(BigArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int D = 1 << a.k;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < a.total; idx += stride) {
        if ((idx & a.cmask) != a.cmask) { a.out[idx] = a.in[idx]; continue; }
        int row = 0;
        for (int b = 0; b < a.k; b++) row |= (int)((idx >> a.tb[b]) & 1) << (a.k - 1 - b);
        uint64_t base = idx & ~a.tmask;
        const cplx* mrow = a.mat + (size_t)row * D;
        cplx acc = qb_cmul(mrow[0], a.in[base | a.offs[0]]);
        for (int j = 1; j < D; j++) acc = qb_cfma(mrow[j], a.in[base | a.offs[j]], acc);
        a.out[idx] = acc;
    }
}

void qb_launch_big(const LaunchCtx& c, const BigArgs& a) {
    k_big<<<grid_for(c, a.total, 256), 256, 0, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// batched dense gate: branch b applies its own matrix / targets / controls (piece 5)
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(K >= 5 ? 128 : 256)
k_dense_batched(cplx* psi, int nbits, const cplx* __restrict__ mats, const int* __restrict__ tb,
                const uint64_t* __restrict__ cmasks, const uint8_t* __restrict__ enable) {
    constexpr int D = 1 << K;
    const uint64_t b = blockIdx.y;
    if (enable && !enable[b]) return;
    __shared__ cplx sm[D * D];
    __shared__ int s_ins[QB_MAX_INS];
    __shared__ int s_nins;
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sm[i] = mats[b * (D * D) + i];
    uint64_t tbm[K];
    uint64_t tmask = 0;
#pragma unroll
    for (int q = 0; q < K; q++) { tbm[q] = 1ull << tb[b * K + q]; tmask |= tbm[q]; }
    const uint64_t cmask = cmasks ? cmasks[b] : 0ull;
    if (threadIdx.x == 0) {
        int n = 0;
        uint64_t m = tmask | cmask;
        for (int p = 0; p < nbits; p++) if ((m >> p) & 1) s_ins[n++] = p;
        s_nins = n;
    }
    __syncthreads();
    const int nins = s_nins;
    const uint64_t nwork = 1ull << (nbits - nins);
    cplx* base_ptr = psi + (b << nbits);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // U independent work items per iteration: all their loads are issued before the first multiply
    // (one block walks a whole branch, so memory-level parallelism has to come from the thread)
    constexpr int U = K == 1 ? 4 : K == 2 ? 2 : 1;
    for (uint64_t w0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < nwork; w0 += stride * U) {
        uint64_t base[U];
        cplx x[U][D];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t w = w0 + (uint64_t)u * stride;
            uint64_t bs = w;
            for (int i = 0; i < nins; i++) bs = qb_insert_zero(bs, s_ins[i]);
            base[u] = bs | cmask;
            if (w < nwork) {
#pragma unroll
                for (int j = 0; j < D; j++) x[u][j] = __ldcs(&base_ptr[base[u] | sub_offset<K>(j, tbm)]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (w0 + (uint64_t)u * stride >= nwork) continue;
#pragma unroll(K <= 2 ? D : 1)
            for (int i = 0; i < D; i++) {
                cplx acc = qb_cmul(sm[i * D], x[u][0]);
#pragma unroll
                for (int j = 1; j < D; j++) acc = qb_cfma(sm[i * D + j], x[u][j], acc);
                __stcs(&base_ptr[base[u] | sub_offset<K>(i, tbm)], acc);
            }
        }
    }
}

void qb_launch_dense_batched(const LaunchCtx& c, int K, cplx* psi, int nbits, int64_t nbranch, const cplx* mats,
                             const int* tb, const uint64_t* cmasks, const uint8_t* enable) {
    int threads = K >= 5 ? 128 : 256;
    uint64_t work = 1ull << (nbits - K);
    uint64_t gx = (work + threads - 1) / threads;
    uint64_t cap = (uint64_t)c.sms * 8 / (uint64_t)(nbranch < 1 ? 1 : nbranch) + 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    QB_REQUIRE(nbranch <= 65535, "batched gate: more than 65535 branches per launch");
    dim3 grid((unsigned)gx, (unsigned)nbranch);
    switch (K) {
        case 1: k_dense_batched<1><<<grid, threads, 0, c.stream>>>(psi, nbits, mats, tb, cmasks, enable); break;
        case 2: k_dense_batched<2><<<grid, threads, 0, c.stream>>>(psi, nbits, mats, tb, cmasks, enable); break;
        case 3: k_dense_batched<3><<<grid, threads, 0, c.stream>>>(psi, nbits, mats, tb, cmasks, enable); break;
        case 4: k_dense_batched<4><<<grid, threads, 0, c.stream>>>(psi, nbits, mats, tb, cmasks, enable); break;
        case 5: k_dense_batched<5><<<grid, threads, 0, c.stream>>>(psi, nbits, mats, tb, cmasks, enable); break;
        default: throw qb_error(-1, "batched gate: k must be 1..5");
    }
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// outcome weights: deterministic two-stage binned reduction (no atomics)
//   stage 1: one block folds a chunk of 2^c consecutive values over its non-target bits in smem
//   stage 2: one block per (branch, outcome) sums the matching chunks with a fixed tree
// ---------------------------------------------------------------------------------------------
#define BIN_C 11
__global__ void __launch_bounds__(256) k_bins(BinArgs a) {
    __shared__ double sre[1 << BIN_C];
    __shared__ double sim[1 << BIN_C];
    // block = branch * (nchunks >> ngroup) + group; the group's representative chunk has zeros at
    // the group bits, its 2^ngroup members are summed element-wise in registers (ascending order)
    const uint64_t ngrp = a.nchunks >> a.ngroup;
    const uint64_t branch = blockIdx.x / ngrp;
    uint64_t q = blockIdx.x % ngrp;
    for (int x = 0; x < a.ngroup; x++) q = qb_insert_zero(q, a.groupbits[x]);
    const uint64_t blk = branch * a.nchunks + q;
    const int C = 1 << a.c;
    const cplx* src = a.src + branch * a.branch_stride;
    const int G = 1 << a.ngroup;
    // element offset of every group member, computed once per block (the inner loop is then one
    // add + one load per element instead of re-depositing the group bits)
    __shared__ uint64_t goff[64];
    for (int gi = threadIdx.x; gi < G; gi += blockDim.x) {
        uint64_t qq = q;
        for (int x = 0; x < a.ngroup; x++) qq |= (uint64_t)((gi >> x) & 1) << a.groupbits[x];
        goff[gi] = (qq << a.c) * a.elem_stride;
    }
    __syncthreads();
    // a thread holds the 2^(BIN_C-8) values l = tid + 256 i of the chunk in registers; non-target ("fold") bits are
    // summed away where they live: bits >= 8 inside the thread, bits 0..4 across the lanes with warp shuffles,
    // bits 5..7 (the warp number) through shared memory -- at most 3 barrier rounds instead of one per fold bit.
    // Fixed order, no atomics: the result is deterministic.
    constexpr int PER = (1 << BIN_C) / 256;
    double vre[PER], vim[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int l = (int)threadIdx.x + 256 * i;
        double re = 0.0, im = 0.0;
        if (l < C) {
            const cplx* p = src + (uint64_t)l * a.elem_stride;
            if (a.mode == 0) {
#pragma unroll 8
                for (int gi = 0; gi < G; gi++) {
                    const cplx v = __ldcs(p + goff[gi]);
                    re += v.x * v.x + v.y * v.y;
                }
            } else {
#pragma unroll 8
                for (int gi = 0; gi < G; gi++) {
                    const cplx v = __ldcs(p + goff[gi]);
                    re += v.x; im += v.y;
                }
            }
        }
        vre[i] = re; vim[i] = im;
    }
    unsigned foldmask = 0;
    for (int f = 0; f < a.nfold; f++) foldmask |= 1u << a.foldbits[f];
#pragma unroll
    for (int b = 0; (1 << b) < PER; b++) {
        if ((foldmask >> (8 + b)) & 1u) {
#pragma unroll
            for (int i = 0; i < PER; i++)
                if (!(i & (1 << b))) { vre[i] += vre[i | (1 << b)]; vim[i] += vim[i | (1 << b)]; }
        }
    }
#pragma unroll
    for (int b = 0; b < 5; b++) {
        if ((foldmask >> b) & 1u) {
#pragma unroll
            for (int i = 0; i < PER; i++) {
                vre[i] += __shfl_xor_sync(0xffffffffu, vre[i], 1 << b);
                if (a.mode != 0) vim[i] += __shfl_xor_sync(0xffffffffu, vim[i], 1 << b);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PER; i++) { sre[threadIdx.x + 256 * i] = vre[i]; sim[threadIdx.x + 256 * i] = vim[i]; }
    __syncthreads();
    unsigned folded = foldmask & ~0xe0u;        // the entries with a folded bit set are dead from here on
    for (int bpos = 5; bpos < 8; bpos++) {
        const unsigned bit = 1u << bpos;
        if (!(foldmask & bit)) continue;
        const unsigned dead = folded | bit;
        for (int l = threadIdx.x; l < C; l += blockDim.x) {
            if ((l & dead) == 0) { sre[l] += sre[l | bit]; sim[l] += sim[l | bit]; }
        }
        folded |= bit;
        __syncthreads();
    }
    const int ML = 1 << a.ml;
    for (int jl = threadIdx.x; jl < ML; jl += blockDim.x) {
        unsigned l = 0;
        for (int r = 0; r < a.ml; r++) l |= ((jl >> r) & 1u) << a.lowt[r];
        a.partial[blk * ML + jl] = make_double2(sre[l], sim[l]);
    }
}

void qb_launch_bins(const LaunchCtx& c, const BinArgs& a, int64_t nbranch) {
    uint64_t blocks = (a.nchunks >> a.ngroup) * (uint64_t)nbranch;
    QB_REQUIRE(blocks < (1ull << 31), "probs: too many chunks");
    k_bins<<<(unsigned)blocks, 256, 0, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

__global__ void __launch_bounds__(256) k_bins_final(BinFinalArgs a) {
    __shared__ double sre[256];
    __shared__ double sim[256];
    const uint64_t M = 1ull << a.m;
    const uint64_t branch = blockIdx.x / M, j = blockIdx.x % M;
    // split outcome j into (low part -> slot inside a chunk's partial, high part -> fixed chunk bits)
    uint64_t jl = 0, hfix = 0, hmask = 0;
    for (int t = 0; t < a.m; t++) {
        uint64_t v = (j >> (a.m - 1 - t)) & 1ull;
        if (a.tbits[t] < a.c) jl |= v << a.lowrank[t];
        else { hfix |= v << (a.tbits[t] - a.c); hmask |= 1ull << (a.tbits[t] - a.c); }
    }
    const int qbits = a.nb - a.c > 0 ? a.nb - a.c : 0;
    hmask |= a.groupmask;              // summed by k_bins already: only the representatives (zeros there) hold data
    int nfix = __popcll(hmask);
    const uint64_t nfree = 1ull << (qbits - nfix);
    const uint64_t ML = 1ull << a.ml;
    double re = 0.0, im = 0.0;
    for (uint64_t r = threadIdx.x; r < nfree; r += blockDim.x) {
        uint64_t q = r;
        for (int p = 0; p < qbits; p++) if ((hmask >> p) & 1ull) q = qb_insert_zero(q, p);
        q |= hfix;
        cplx v = a.partial[(branch * a.nchunks + q) * ML + jl];
        re += v.x; im += v.y;
    }
    sre[threadIdx.x] = re; sim[threadIdx.x] = im;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sre[threadIdx.x] += sre[threadIdx.x + s]; sim[threadIdx.x] += sim[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out[blockIdx.x] = make_double2(sre[0], sim[0]);
}

// few partials per outcome (a branch batch: 4096 x 16 outcomes with ONE partial each): a thread per
// (branch, outcome) instead of a 256-thread block -- 65 536 blocks summing one value each took 131 us
__global__ void __launch_bounds__(256) k_bins_final_small(BinFinalArgs a, uint64_t noutcomes) {
    const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= noutcomes) return;
    const uint64_t M = 1ull << a.m;
    const uint64_t branch = o / M, j = o % M;
    uint64_t jl = 0, hfix = 0, hmask = 0;
    for (int t = 0; t < a.m; t++) {
        uint64_t v = (j >> (a.m - 1 - t)) & 1ull;
        if (a.tbits[t] < a.c) jl |= v << a.lowrank[t];
        else { hfix |= v << (a.tbits[t] - a.c); hmask |= 1ull << (a.tbits[t] - a.c); }
    }
    const int qbits = a.nb - a.c > 0 ? a.nb - a.c : 0;
    hmask |= a.groupmask;
    const int nfix = __popcll(hmask);
    const uint64_t nfree = 1ull << (qbits - nfix);
    const uint64_t ML = 1ull << a.ml;
    double re = 0.0, im = 0.0;
    for (uint64_t r = 0; r < nfree; r++) {
        uint64_t q = r;
        for (int p = 0; p < qbits; p++) if ((hmask >> p) & 1ull) q = qb_insert_zero(q, p);
        q |= hfix;
        const cplx v = a.partial[(branch * a.nchunks + q) * ML + jl];
        re += v.x; im += v.y;
    }
    a.out[o] = make_double2(re, im);
}

void qb_launch_bins_final(const LaunchCtx& c, const BinFinalArgs& a, int64_t nbranch) {
    uint64_t blocks = ((uint64_t)nbranch) << a.m;
    QB_REQUIRE(blocks < (1ull << 31), "probs: too many outcomes");
    const int qbits = a.nb - a.c > 0 ? a.nb - a.c : 0;
    uint64_t hmask = a.groupmask;
    for (int t = 0; t < a.m; t++) if (a.tbits[t] >= a.c) hmask |= 1ull << (a.tbits[t] - a.c);
    const int free_bits = qbits - __builtin_popcountll(hmask);
    if (free_bits <= 3 && blocks >= 1024) {
        k_bins_final_small<<<(unsigned)((blocks + 255) / 256), 256, 0, c.stream>>>(a, blocks);
    } else {
        k_bins_final<<<(unsigned)blocks, 256, 0, c.stream>>>(a);
    }
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// partial trace: out[i][j] = sum_t rho[dep(i)|sp(t)][dep(j)|sp(t)]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pt_spread(uint64_t v, const int* pos, int n, bool msb_first) {
    uint64_t o = 0;
    for (int x = 0; x < n; x++) {
        uint64_t bit = msb_first ? (v >> (n - 1 - x)) & 1ull : (v >> x) & 1ull;
        o |= bit << pos[x];
    }
    return o;
}

// small traced space: one thread per output entry, terms added in ascending t (np.trace order)
__global__ void __launch_bounds__(256) k_ptrace_seq(PtraceArgs a) {
    const uint64_t A = 1ull << a.nkeep, total = A * A, T = 1ull << a.ntr;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        uint64_t rb = pt_spread(e >> a.nkeep, a.keepb, a.nkeep, true);
        uint64_t cb = pt_spread(e & (A - 1), a.keepb, a.nkeep, true);
        double re = 0.0, im = 0.0;
        for (uint64_t t = 0; t < T; t++) {
            uint64_t sp = pt_spread(t, a.trb, a.ntr, false);
            cplx v = a.rho[((rb | sp) << a.nq) | (cb | sp)];
            re += v.x; im += v.y;
        }
        a.out[e] = make_double2(re, im);
    }
}

// large traced space: one block per output entry, fixed-shape tree
__global__ void __launch_bounds__(256) k_ptrace_blk(PtraceArgs a) {
    __shared__ double sre[256];
    __shared__ double sim[256];
    const uint64_t A = 1ull << a.nkeep, total = A * A, T = 1ull << a.ntr;
    for (uint64_t e = blockIdx.x; e < total; e += gridDim.x) {
        uint64_t rb = pt_spread(e >> a.nkeep, a.keepb, a.nkeep, true);
        uint64_t cb = pt_spread(e & (A - 1), a.keepb, a.nkeep, true);
        double re = 0.0, im = 0.0;
        for (uint64_t t = threadIdx.x; t < T; t += blockDim.x) {
            uint64_t sp = pt_spread(t, a.trb, a.ntr, false);
            cplx v = a.rho[((rb | sp) << a.nq) | (cb | sp)];
            re += v.x; im += v.y;
        }
        sre[threadIdx.x] = re; sim[threadIdx.x] = im;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) { sre[threadIdx.x] += sre[threadIdx.x + s]; sim[threadIdx.x] += sim[threadIdx.x + s]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) a.out[e] = make_double2(sre[0], sim[0]);
        __syncthreads();
    }
}

// reduced density matrix of a KET: out[i][j] = sum_t psi[dep(i)|sp(t)] * conj(psi[dep(j)|sp(t)])
// (the partial trace of |psi><psi| without ever forming it).  One block per output entry,
// fixed-shape tree; meant for a handful of kept qubits (the measured system of a `peek`).
__global__ void __launch_bounds__(256) k_ket_rdm(PtraceArgs a) {
    __shared__ double sre[256];
    __shared__ double sim[256];
    const uint64_t A = 1ull << a.nkeep, total = A * A, T = 1ull << a.ntr;
    for (uint64_t e = blockIdx.x; e < total; e += gridDim.x) {
        const uint64_t rb = pt_spread(e >> a.nkeep, a.keepb, a.nkeep, true);
        const uint64_t cb = pt_spread(e & (A - 1), a.keepb, a.nkeep, true);
        double re = 0.0, im = 0.0;
        for (uint64_t t = threadIdx.x; t < T; t += blockDim.x) {
            const uint64_t sp = pt_spread(t, a.trb, a.ntr, false);
            const cplx x = a.rho[rb | sp], y = a.rho[cb | sp];
            re += x.x * y.x + x.y * y.y;
            im += x.y * y.x - x.x * y.y;
        }
        sre[threadIdx.x] = re; sim[threadIdx.x] = im;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) { sre[threadIdx.x] += sre[threadIdx.x + s]; sim[threadIdx.x] += sim[threadIdx.x + s]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) a.out[e] = make_double2(sre[0], sim[0]);
        __syncthreads();
    }
}

void qb_launch_ket_rdm(const LaunchCtx& c, const PtraceArgs& a) {
    const uint64_t total = 1ull << (2 * a.nkeep);
    const uint64_t blocks = total < (uint64_t)c.sms * 16 ? total : (uint64_t)c.sms * 16;
    k_ket_rdm<<<(unsigned)blocks, 256, 0, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

void qb_launch_ptrace(const LaunchCtx& c, const PtraceArgs& a) {
    uint64_t total = 1ull << (2 * a.nkeep);
    if (a.ntr <= 6) {
        k_ptrace_seq<<<grid_for(c, total, 256), 256, 0, c.stream>>>(a);
    } else {
        uint64_t blocks = total < (uint64_t)c.sms * 16 ? total : (uint64_t)c.sms * 16;
        k_ptrace_blk<<<(unsigned)blocks, 256, 0, c.stream>>>(a);
    }
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// scatter product: out[r][c] = A[ga(r)][ga(c)] * B[gb(r)][gb(c)]   (interweave / replace)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t sc_gather(uint64_t v, const int* pos, int n) {
    uint64_t o = 0;
    for (int x = 0; x < n; x++) o |= ((v >> pos[x]) & 1ull) << (n - 1 - x);
    return o;
}

// index tables: tA[i] = A-index of the n-bit row / column index i, tB[i] likewise (the same map
// serves rows and columns).  Evaluating the two bit gathers per output entry cost ~100 integer
// instructions per 16-byte store (1.2 TB/s); with the tables the kernel is a plain HBM write.
__global__ void __launch_bounds__(256) k_scatter_tables(ScatterArgs a, uint32_t* tA, uint32_t* tB) {
    const uint64_t N = 1ull << a.n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        tA[i] = (uint32_t)sc_gather(i, a.abits, a.na);
        tB[i] = (uint32_t)sc_gather(i, a.bbits, a.nb);
    }
}

__global__ void __launch_bounds__(256) k_scatter(ScatterArgs a, const uint32_t* __restrict__ tA, const uint32_t* __restrict__ tB) {
    const uint64_t N = 1ull << a.n, total = N * N;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t r = e >> a.n, cidx = e & (N - 1);
        cplx v = a.a[((uint64_t)tA[r] << a.na) | tA[cidx]];
        if (a.b) v = qb_cmul(v, a.b[((uint64_t)tB[r] << a.nb) | tB[cidx]]);
        if (a.has_scale) v = qb_cmul(v, a.scale);
        __stcs(a.out + e, v);
    }
}

void qb_launch_scatter(const LaunchCtx& c, const ScatterArgs& a, uint32_t* tables_dev) {
    const uint64_t N = 1ull << a.n, total = N * N;
    uint32_t* tA = tables_dev;
    uint32_t* tB = tables_dev + N;
    k_scatter_tables<<<grid_for(c, N, 256), 256, 0, c.stream>>>(a, tA, tB);
    COUNT_LAUNCH(c);
    k_scatter<<<grid_for(c, total, 256), 256, 0, c.stream>>>(a, tA, tB);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// ensemble mix: out = sum_i p_i * src_i, list order, round-to-nearest mul then add (no FMA),
// so it is bit-identical to the reference's `result += probs[i] * density` loop
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mix(MixArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.total; e += stride) {
        double re = 0.0, im = 0.0;
        if (a.accumulate) { cplx o = a.out[e]; re = o.x; im = o.y; }
        for (int i = 0; i < a.count; i++) {
            cplx v = a.src[i][e];
            re = __dadd_rn(re, __dmul_rn(a.p[i], v.x));
            im = __dadd_rn(im, __dmul_rn(a.p[i], v.y));
        }
        a.out[e] = make_double2(re, im);
    }
}

void qb_launch_mix(const LaunchCtx& c, const MixArgs& a) {
    k_mix<<<grid_for(c, a.total, 256), 256, 0, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

__global__ void __launch_bounds__(256) k_mix_branches(const cplx* __restrict__ src, const double* __restrict__ p,
                                                      int64_t nbranch, uint64_t per, cplx* out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += stride) {
        // list order, no FMA contraction (bit-identical to the numpy loop); the loads of 8 branches
        // are in flight at once, only the additions are sequential
        double re = 0.0, im = 0.0;
        int64_t b = 0;
        for (; b + 8 <= nbranch; b += 8) {
            cplx v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldcs(&src[(uint64_t)(b + u) * per + e]);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                re = __dadd_rn(re, __dmul_rn(p[b + u], v[u].x));
                im = __dadd_rn(im, __dmul_rn(p[b + u], v[u].y));
            }
        }
        for (; b < nbranch; b++) {
            const cplx v = src[(uint64_t)b * per + e];
            re = __dadd_rn(re, __dmul_rn(p[b], v.x));
            im = __dadd_rn(im, __dmul_rn(p[b], v.y));
        }
        out[e] = make_double2(re, im);
    }
}

void qb_launch_mix_branches(const LaunchCtx& c, const cplx* src, const double* probs_dev, int64_t nbranch, uint64_t per, cplx* out) {
    k_mix_branches<<<grid_for(c, per, 256), 256, 0, c.stream>>>(src, probs_dev, nbranch, per, out);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// rank-1 density from a ket
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_outer(const cplx* __restrict__ ket, cplx* out, int nq, int conj) {
    const uint64_t N = 1ull << nq, total = N * N;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        cplx r = ket[e >> nq], cc = ket[e & (N - 1)];
        if (conj) cc.y = -cc.y;
        out[e] = qb_cmul(r, cc);
    }
}

void qb_launch_outer(const LaunchCtx& c, const cplx* ket, cplx* out, int nq, int conj) {
    uint64_t total = 1ull << (2 * nq);
    k_outer<<<grid_for(c, total, 256), 256, 0, c.stream>>>(ket, out, nq, conj);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// ket projection + renormalisation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project(cplx* psi, uint64_t total, uint64_t mask, uint64_t want, double scale) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        cplx v = psi[i];
        if ((i & mask) == want) psi[i] = make_double2(v.x * scale, v.y * scale);
        else psi[i] = make_double2(0.0, 0.0);
    }
}

void qb_launch_project(const LaunchCtx& c, cplx* psi, uint64_t total, uint64_t mask, uint64_t want, double scale) {
    k_project<<<grid_for(c, total, 256), 256, 0, c.stream>>>(psi, total, mask, want, scale);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// index-bit permutation with per-chunk destinations (pack step of the global-qubit swap; the
// destinations may be peer GPUs' buffers mapped over CUDA IPC, in which case the stores travel
// over NVLink and this kernel *is* the exchange).  One 16-byte amplitude per thread per
// iteration; the planner keeps the low 5 bits in place so that warps read and write whole
// 512-byte runs.
// ---------------------------------------------------------------------------------------------
// Work is cut into UNITS of 2^ub consecutive destination amplitudes (64 KB when the chunk allows).
// Unit w goes to chunk (w mod 2^k) ^ first_chunk, so that (a) a GPU streams to all its peers at
// the same time and (b) ranks that run this kernel concurrently never all target one peer (the
// rotation is the rank's own chunk number: a pairwise-exchange schedule, no incast on a link).
// The source offset of a thread's amplitudes inside a unit does not depend on the unit, so the
// bit permutation is evaluated once per thread (low bits) and once per unit (high bits).
__device__ __forceinline__ uint64_t perm_bits(const PermArgs& a, uint64_t j) {
    uint64_t src = j & a.fixed_mask;
    for (int i = 0; i < a.nmoved; i++) src |= ((j >> a.to[i]) & 1ull) << a.from[i];
    return src;
}

__global__ void __launch_bounds__(256) k_permute_scatter(PermArgs a) {
    const int ub = a.unit_bits, k = a.chunk_bits;
    const uint32_t per_thread = (1u << ub) >> 8;               // amplitudes per thread and unit (ub >= 8): 1..16
    uint64_t soff[16];
#pragma unroll
    for (int t = 0; t < 16; t++) soff[t] = perm_bits(a, (uint64_t)threadIdx.x + 256ull * t);
    const uint64_t nunits = a.total >> ub;
    const uint64_t cmask = (1ull << k) - 1ull;
    for (uint64_t w = blockIdx.x; w < nunits; w += gridDim.x) {
        const uint64_t c = (w & cmask) ^ (uint64_t)a.first_chunk;
        const uint64_t o = (w >> k) << ub;                      // offset of the unit inside its chunk
        const uint64_t ubase = perm_bits(a, (c << a.chunk_shift) | o) | a.src_or;
        cplx* __restrict__ out = a.dst[c] + o + threadIdx.x;
        cplx v[16];
#pragma unroll
        for (int t = 0; t < 16; t++) if ((uint32_t)t < per_thread) v[t] = __ldcs(&a.in[ubase | soff[t]]);
#pragma unroll
        for (int t = 0; t < 16; t++) if ((uint32_t)t < per_thread) __stcs(out + 256 * t, v[t]);
    }
}

void qb_launch_permute_scatter(const LaunchCtx& c, const PermArgs& a, int max_ctas) {
    int grid = grid_for(c, a.total >> 4, 256, 4);
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;      // leave the other SM slots to the sweeps this exchange overlaps with
    k_permute_scatter<<<grid, 256, 0, c.stream>>>(a);
    COUNT_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------
// stream-ordered signals between the GPUs of a box (pieces of a pipelined global-qubit exchange):
// after a piece's stores to the peers have completed (stream order: the scatter kernel has ended, so
// its writes are acknowledged), the sender raises a counter in every receiver's flag block; the
// receiver's compute stream waits (one polling warp, bounded) until all its sources have raised
// theirs, and only then runs the sweeps that read the piece.  The flag blocks are peer-mapped device
// memory (CUDA IPC), like the shard buffers.
// ---------------------------------------------------------------------------------------------
__global__ void k_signal_flags(FlagPtrs f, unsigned long long value) {
    const int i = threadIdx.x;
    if (i < f.n) {
        __threadfence_system();
        *(volatile unsigned long long*)f.p[i] = value;
        __threadfence_system();
    }
}

__global__ void k_wait_flags(FlagPtrs f, unsigned long long value, unsigned long long* timeouts) {
    const int i = threadIdx.x;
    if (i < f.n) {
        const long long t0 = clock64();
        volatile unsigned long long* p = (volatile unsigned long long*)f.p[i];
        bool ok = false;
        while (!(ok = *p >= value)) {
            if (clock64() - t0 > 120000000000ll) break;       // ~60 s (ranks may be seconds apart while one compiles): a peer died; fail (counted) instead of hanging the GPU
            __nanosleep(500);
        }
        __threadfence_system();
        if (!ok && timeouts) atomicAdd(timeouts, 1ull);
    }
}

void qb_launch_signal_flags(cudaStream_t stream, const FlagPtrs& f, unsigned long long value) {
    k_signal_flags<<<1, 32, 0, stream>>>(f, value);
}
void qb_launch_wait_flags(cudaStream_t stream, const FlagPtrs& f, unsigned long long value, unsigned long long* timeouts) {
    k_wait_flags<<<1, 32, 0, stream>>>(f, value, timeouts);
}
