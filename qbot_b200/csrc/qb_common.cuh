// qbot_b200 -- shared definitions for the CUDA state-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <stdexcept>

#include "qb_gate.h"

#define QB_MAX_BITS   48      // index bits of one branch (ket: nq, dm: 2*nq)
#define QB_MAX_INS    48      // zero-insert positions (targets + controls)
#define QB_REG_MAXK   5       // dense gates up to 5 target bits run out of registers
#define QB_DIAG_MAXK  12
#define QB_MMA_MINK   6       // dense gates of 6..8 target bits run on the FP64 tensor cores (qb_dense_mma.cu)
#define QB_MMA_MAXK   8
#define QB_MMA_TILE_BITS 12   // targets + column bits of one CTA tile

struct qb_error : public std::runtime_error {
    int code;
    qb_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define QB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            throw qb_error(e__ == cudaErrorMemoryAllocation ? -3 : -2,                         \
                           std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)

#define QB_REQUIRE(cond, msg)                                                                  \
    do {                                                                                       \
        if (!(cond)) throw qb_error(-1, std::string(msg));                                     \
    } while (0)

__host__ __device__ __forceinline__ uint64_t qb_insert_zero(uint64_t w, int p) {
    return ((w >> p) << (p + 1)) | (w & ((1ull << p) - 1ull));
}

__device__ __forceinline__ cplx qb_cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx qb_cfma(cplx a, cplx b, cplx c) {   // a*b + c
    return make_double2(c.x + a.x * b.x - a.y * b.y, c.y + a.x * b.y + a.y * b.x);
}

// ---- kernel argument blocks (passed by value) ------------------------------------------------
struct DenseArgs {
    cplx* psi;
    const cplx* mat;              // device matrix for K >= 3
    uint64_t nwork;
    uint64_t cmask;
    int nins;
    int ins[QB_MAX_INS];
    int tb[QB_REG_MAXK];
    cplx inl[16];                 // inline matrix for K <= 2
};

struct DiagArgs {
    cplx* psi;
    const cplx* diag;             // device diagonal for k >= 3
    uint64_t nwork;
    uint64_t cmask;
    int nins;
    int ins[QB_MAX_INS];
    int k;
    int tb[QB_DIAG_MAXK];
    cplx inl[4];
};

struct BigArgs {
    const cplx* in;
    cplx* out;
    const cplx* mat;
    const uint64_t* offs;
    uint64_t total;
    uint64_t cmask;
    uint64_t tmask;
    int k;
    int tb[QB_BIG_MAXK];
};

struct DenseMmaArgs {
    cplx* psi;
    const double* ur;             // device matrix, real plane, row-major 2^k x 2^k
    const double* ui;             // imaginary plane
    uint64_t ntiles;
    uint64_t cmask;
    int nins;
    int ins[QB_MAX_INS];          // targets + column bits + controls, ascending
    int apos[QB_MMA_TILE_BITS];   // index bit of tile-local bit b (ascending)
    int sw[QB_MMA_TILE_BITS];     // its contribution to the shared-memory word offset (row * LDX + column)
};

struct BinArgs {
    const cplx* src;
    cplx* partial;
    int mode;                     // 0: |psi_i|^2   1: rho_ii
    int nb;                       // bits of the binned index i
    uint64_t elem_stride;         // ket 1, dm 2^nq + 1
    uint64_t branch_stride;       // amplitudes per branch
    int c;                        // chunk bits
    uint64_t nchunks;             // per branch
    int nfold;
    int foldbits[16];
    int ml;
    int lowt[16];                 // low target bit positions, ascending
    int ngroup;                   // chunks that differ only in these (non-target) chunk-index bits are
    int groupbits[8];             //   summed by ONE block before the fold (positions in the chunk index, ascending)
};

struct BinFinalArgs {
    const cplx* partial;
    cplx* out;                    // [nbranch][2^m]
    int m;
    int c;
    int nb;
    uint64_t nchunks;
    int ml;
    int tbits[QB_MAX_BITS];       // target bits, MSB-first as listed
    int lowrank[QB_MAX_BITS];     // for target t with bit < c: its rank among low targets, else -1
    uint64_t groupmask;           // chunk-index bits already summed by k_bins (only representatives hold data)
};

struct PtraceArgs {
    const cplx* rho;
    cplx* out;
    int nq;
    int nkeep;
    int ntr;
    int keepb[QB_MAX_BITS];       // MSB-first
    int trb[QB_MAX_BITS];         // ascending
};

struct ScatterArgs {
    const cplx* a;
    const cplx* b;                // may be null
    cplx* out;
    int n, na, nb;
    int abits[QB_MAX_BITS];       // MSB-first positions in the n-bit index
    int bbits[QB_MAX_BITS];
    int has_scale;
    cplx scale;
};

#define QB_MIX_MAX 16
struct MixArgs {
    const cplx* src[QB_MIX_MAX];
    double p[QB_MIX_MAX];
    int count;
    int accumulate;               // 1: out already holds a partial sum
    cplx* out;
    uint64_t total;
};

#define QB_PERM_MAXMOVED 40
#define QB_PERM_MAXCHUNK 16
struct PermArgs {
    const cplx* in;
    cplx* dst[QB_PERM_MAXCHUNK];  // destination of chunk c (local buffer or a peer mapping)
    uint64_t total;               // amplitudes of the shard
    uint64_t fixed_mask;          // index bits that keep their position
    uint64_t src_or;              // source index bits that are constant for this launch (a sub-block of the source)
    int chunk_shift;              // nbits - chunk_bits
    int chunk_bits;
    int unit_bits;                // 2^unit_bits consecutive destination amplitudes per work unit (8..12)
    int first_chunk;              // rotation of the chunk visiting order
    int nmoved;
    uint8_t from[QB_PERM_MAXMOVED];   // source bit of moved destination bit to[i]
    uint8_t to[QB_PERM_MAXMOVED];
};

// launch wrappers implemented in qb_kernels.cu ------------------------------------------------
struct LaunchCtx {
    cudaStream_t stream;
    int sms;
    uint64_t* launches;
};

void qb_launch_fill_basis(const LaunchCtx&, cplx* d, uint64_t per_branch, int64_t nbranch, uint64_t index);
void qb_launch_fill_diag(const LaunchCtx&, cplx* d, int nq, const double* values_dev);
void qb_launch_broadcast(const LaunchCtx&, const cplx* src, cplx* dst, uint64_t per_branch, int64_t nbranch);
void qb_launch_init_product(const LaunchCtx&, cplx* d, int kind, int nq, int64_t nbranch, const cplx* vecs_dev, int per_branch);
void qb_launch_dense(const LaunchCtx&, int K, const DenseArgs& a);
void qb_launch_diag(const LaunchCtx&, const DiagArgs& a);
void qb_launch_swap(const LaunchCtx&, cplx* d, uint64_t total, int lo, int hi);
void qb_launch_big(const LaunchCtx&, const BigArgs& a);
void qb_launch_dense_mma(const LaunchCtx&, int K, const DenseMmaArgs& a);     // qb_dense_mma.cu
void qb_launch_dense_batched(const LaunchCtx&, int K, cplx* psi, int nbits, int64_t nbranch, const cplx* mats,
                             const int* tb, const uint64_t* cmasks, const uint8_t* enable);
void qb_launch_bins(const LaunchCtx&, const BinArgs& a, int64_t nbranch);
void qb_launch_bins_final(const LaunchCtx&, const BinFinalArgs& a, int64_t nbranch);
void qb_launch_ptrace(const LaunchCtx&, const PtraceArgs& a);
void qb_launch_ket_rdm(const LaunchCtx&, const PtraceArgs& a);
void qb_launch_scatter(const LaunchCtx&, const ScatterArgs& a, uint32_t* tables_dev);   // tables_dev: 2 * 2^n uint32 of work space
void qb_launch_mix(const LaunchCtx&, const MixArgs& a);
void qb_launch_mix_branches(const LaunchCtx&, const cplx* src, const double* probs_dev, int64_t nbranch, uint64_t per_branch, cplx* out);
void qb_launch_outer(const LaunchCtx&, const cplx* ket, cplx* out, int nq, int conj);
void qb_launch_permute_scatter(const LaunchCtx&, const PermArgs& a, int max_ctas = 0);
#define QB_FLAG_MAXPEERS 16
struct FlagPtrs { unsigned long long* p[QB_FLAG_MAXPEERS]; int n; };
void qb_launch_signal_flags(cudaStream_t stream, const FlagPtrs& f, unsigned long long value);
void qb_launch_wait_flags(cudaStream_t stream, const FlagPtrs& f, unsigned long long value, unsigned long long* timeouts);
void qb_launch_project(const LaunchCtx&, cplx* psi, uint64_t total, uint64_t mask, uint64_t want, double scale);
