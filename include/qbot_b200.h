/*
 * qbot_b200 C ABI -- the drop-in boundary of the B200 state-path backend.
 *
 * The reference (PrinceOfPuppers/qbot) has no FFI layer: its state path is Python calling
 * numpy.  The nearest thing to a plugin interface is the `operations` table
 * (qbot/operators.py:477-506); the six state ops in it (`qset`, `gate`, `disc`, `swap`, `meas`,
 * `peek`) do all their arithmetic through the functions named next to each entry point below.
 * A maintainer binds this library with ctypes and re-registers those six
 * ops; nothing else in the reference changes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative qb_status otherwise, and never throws;
 *     qb_last_error() returns the message of the last failure on the calling thread.
 *   - all matrices / vectors crossing the boundary are HOST pointers to row-major
 *     complex128 (interleaved re, im doubles) unless the name says `_dev`.
 *   - a state is an opaque handle owning `nbranch << nbits` complex128 amplitudes in HBM,
 *     nbits = nqubits (ket) or 2*nqubits (density matrix stored row-major, row index in the
 *     high nqubits bits).  "bit" arguments are positions in the basis-state index of ONE
 *     branch, bit 0 = stride-1.  Reference qubit q of an n-qubit register is bit n-1-q
 *     (qbot/qgates.py:161-182: qubit 0 is the most significant index bit).
 *   - calls are asynchronous on the handle's stream except those that return host data.
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails with
 *     QB_ERR_CUDA.
 */
#ifndef QBOT_B200_H
#define QBOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb_state qb_state;

typedef enum {
    QB_OK = 0,
    QB_ERR_ARG = -1,      /* bad argument (message says which) */
    QB_ERR_CUDA = -2,     /* CUDA runtime error (message = cudaGetErrorString) */
    QB_ERR_ALLOC = -3,    /* out of device memory */
    QB_ERR_STATE = -4     /* operation not valid for this kind of state */
} qb_status;

enum { QB_KET = 0, QB_DM = 1 };

/* counters a caller can read back (bench.py: gpu_launches, passes, bytes) */
typedef struct {
    uint64_t kernel_launches;   /* CUDA kernels launched on behalf of this handle */
    uint64_t gates_applied;     /* ket-level gate operations executed */
    uint64_t state_passes;      /* full read+write sweeps over the state */
    uint64_t fused_passes;      /* sweeps done by the fused tile kernel */
    uint64_t fused_gates;       /* gates executed inside fused sweeps */
    uint64_t bytes_moved;       /* bytes the launched kernels were asked to read + write */
    uint64_t jit_passes;        /* fused sweeps run by a structure-specialised (NVRTC) kernel */
    uint64_t jit_kernel_hash;   /* sum (mod 2^64) of the source hashes of those kernels, one term per launch:
                                   identifies the kernel set a measurement (ncu capture, bench line) was taken on */
} qb_stats;

/* ---- library ------------------------------------------------------------------------ */
const char* qb_version(void);
const char* qb_last_error(void);
int qb_device_count(int* count);
int qb_device_info(int device, char* name, int name_len, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor);

/* ---- lifetime ------------------------------------------------------------------------
 * The register the reference keeps in localNameSpace['state'] (qbot/interpreter.py:218-224). */
int qb_create(qb_state** out, int kind, int nqubits, int64_t nbranch, int device);
/* adopt memory owned by someone else (e.g. a torch tensor); stream may be NULL */
int qb_create_external(qb_state** out, int kind, int nqubits, int64_t nbranch, int device,
                       void* amplitudes_dev, void* cuda_stream);
int qb_destroy(qb_state* s);
int qb_clone(const qb_state* s, qb_state** out);
int qb_info(const qb_state* s, int* kind, int* nqubits, int64_t* nbranch, int* device);
int qb_device_ptr(qb_state* s, void** amplitudes_dev);
int qb_set_stream(qb_state* s, void* cuda_stream);

/* ---- state constructors: density.tensorProd / tensorExp / Basis.__getitem__ results
 *      (qbot/density.py:7-29, qbot/basis.py:27-28) without a host 2^n array ---------------- */
int qb_init_basis(qb_state* s, uint64_t index);                       /* |index><index| or |index> in every branch */
/* per-qubit 2-vectors, qubit 0 first: vecs[(b*nq + q)*2 + {0,1}] complex; per_branch=0 shares one set */
int qb_init_product(qb_state* s, const double* vecs, int per_branch);
/* density matrix <- diag(values) (2^nqubits real weights): the collapsed measured system in the
 * computational frame, measurement.py:160-161, before it is rotated into the basis */
int qb_init_diag(qb_state* s, const double* values);
int qb_upload(qb_state* s, const void* host, size_t bytes);            /* qset with a host array (operators.py:143-153) */
int qb_download(qb_state* s, void* host, size_t bytes);                /* what a user expression reading `state` sees */
int qb_download_range(qb_state* s, uint64_t first_amp, uint64_t count, void* host);

/* ---- gate application ------------------------------------------------------------------
 * replaces qgates.genGateForFullHilbertSpace (qgates.py:161-182), genMultiControlledGate
 * (228-275) and applyGate (278-279) as used by operators._gate / gate (operators.py:255-329):
 * ket: psi <- U_c psi ; density matrix: rho <- U_c rho U_c^dagger, where U_c applies `matrix`
 * to target_bits when every bit of control_mask is 1.  target_bits[0] carries the matrix's
 * most significant index bit.  Gates are queued and fused; qb_flush drains the queue. */
int qb_apply_gate(qb_state* s, const double* matrix, int k, const int* target_bits, uint64_t control_mask);
/* A whole gate list in one call (same semantics as `ngates` calls of qb_apply_gate, in order):
 * ks[g], target_bits[g*14 .. g*14+ks[g]-1] (14 = the largest gate), control_masks[g], matrices back to
 * back (4^ks[g] complex each).  For hosts whose per-call overhead matters (a 20-qubit circuit of
 * 2913 gates is bound by ~4 us of Python per qb_apply_gate call). */
int qb_apply_gates(qb_state* s, int ngates, const int* ks, const int* target_bits, const uint64_t* control_masks,
                   const double* matrices);
/* replaces genSwapGate (qgates.py:77-133) + applyGate as used by operators._swap / swap (364-393) */
int qb_apply_swap(qb_state* s, int bit_a, int bit_b);
/* piece (5): one launch over all branches, each with its own gate (probVal.funcWrapper loop,
 * probVal.py:347-390 + operators.py:313-316).  matrices[b*4^k..], target_bits[b*k..],
 * control_masks[b], enable[b] (0 = leave branch b untouched; NULL = all enabled). */
int qb_apply_gate_batched(qb_state* s, const double* matrices, int k, const int* target_bits,
                          const uint64_t* control_masks, const uint8_t* enable);
/* density matrices only: rho <- R rho C^T with R applied to the row copy and C to the column copy
 * of target_bits (either may be NULL = identity).  qb_apply_gate is the case R = U, C = conj(U).
 * Used for the reference's non-conjugating forms: projector weights phi^T rho phi of
 * basis.density = outer(ket, ket) (qbot/basis.py:24-26, qbot/measurement.py:147-155: R = C = W,
 * rows of W = the basis kets) and the collapsed system sum_i p_i phi_i phi_i^T
 * (qbot/measurement.py:160-161: R = C = W^T on diag(p)). */
int qb_apply_gate_rc(qb_state* s, const double* row_matrix, const double* col_matrix, int k, const int* target_bits);
int qb_flush(qb_state* s);
int qb_sync(qb_state* s);
int qb_set_fusion(qb_state* s, int enabled);
/* Sweep specialisation.  A fused sweep is normally run by the generic tile kernel, which
 * interprets the sweep's program; when the same gate sequence is flushed again (a loop in a .qb
 * script, a benchmark step) its sweeps are compiled once with NVRTC into kernels in which the
 * circuit structure is constant and the gate coefficients are parameters.  mode 0: never,
 * 1: when a plan repeats on a large state (default; env QBOT_B200_JIT), 2: always, -1: default. */
int qb_set_jit(qb_state* s, int mode);
/* specialiser counters of the process: kernels compiled, cache hits, total compile time (ms) */
int qb_jit_info(uint64_t* kernels_compiled, uint64_t* cache_hits, double* compile_ms);
/* Generate the specialised source of every fused sweep of a gate list on an nbits-bit state and
 * compile it for sm_100a WITHOUT running it (works without a GPU: a build / CI check).  Gates as in
 * qb_apply_gate, concatenated: ks[g], target_bits[g*14..], control_masks[g], matrices back to back.
 * Returns the number of kernels compiled in *ncompiled; if cubin_dir is not NULL the cubins are
 * written there as sweep_<i>.cubin (for cuobjdump). */
int qb_jit_check(int nbits, int ngates, const int* ks, const int* target_bits, const uint64_t* control_masks,
                 const double* matrices, int* ncompiled, const char* cubin_dir);

/* ---- measurement ----------------------------------------------------------------------
 * computational-basis outcome weights of `m` listed bits (bits[0] = most significant outcome
 * bit): ket  out[b*2^m + j] = sum |psi|^2 ; density matrix  out = |sum of diagonal entries|.
 * This is the abs(trace(rho_A P_j)) loop of measurement.measureArbitraryMultiState
 * (measurement.py:147-155) for the computational basis; other bases rotate first. */
int qb_probs(qb_state* s, const int* bits, int m, double* out);
/* Outcome weights in a product of measurement bases (measurement.permuteBasis + the outcome loop,
 * qbot/measurement.py:88-101, 147-155): the m listed bits are cut into m/b groups of b bits
 * (bits[0..b-1] = the first group, most significant digit of the outcome index; inside a group the
 * first bit is the most significant index bit of the basis kets); `basis` holds the 2^b basis kets
 * as the rows of a 2^b x 2^b complex matrix W.  out[branch * L^(m/b) + i], L = 2^b:
 *   density matrix  |tr(rho_A P_i)|, P_i = (x)_f outer(ket_{d_f(i)}, ket_{d_f(i)}) -- computed as the
 *                   diagonal of (W..W) rho (W..W)^T on a scratch copy (two gate sweeps + the binned read);
 *   ket             sum |(W..W) psi|^2 over the other bits (identical for real bases, the only kind
 *                   the DSL can name; the reference has no ket path).
 * basis == NULL is qb_probs. The state itself is not modified. */
int qb_probs_basis(qb_state* s, const int* bits, int m, const double* basis, int b, double* out);
int qb_norm2(qb_state* s, double* out);                                /* per branch: <psi|psi> or tr rho */
/* ket only: project the listed bits on `outcome` and renormalise (textbook collapse) */
int qb_project_renorm(qb_state* s, const int* bits, int m, uint64_t outcome);

/* ---- density-matrix structure ops -----------------------------------------------------
 * density.partialTraceArbitrary (density.py:122-148): keep the listed qubit-bits (keep_bits[0]
 * = most significant bit of the result), trace the others.  Returns a new DM handle.  For a ket
 * the result is Tr_rest |psi><psi|, computed from the amplitudes (nkeep <= 13). */
int qb_ptrace(qb_state* s, const int* keep_bits, int nkeep, qb_state** out);
/* density.interweaveDensities / replaceArbitrary (density.py:150-227): out = A (x) B with A's
 * i-th qubit-bit (most significant first) placed at bit a_bits[i] of the result and B's at
 * b_bits[i].  b may be NULL (nb = 0). `scale` multiplies the result (the 1x1 [[tr rho]]
 * factor the reference carries when nothing is left, density.py:203-204). */
int qb_scatter_product(qb_state* a, qb_state* b, const int* a_bits, const int* b_bits,
                       const double* scale, qb_state** out);
/* density.densityEnsambleToDensity (density.py:49-58): out = sum_i probs[i] * states[i],
 * accumulated in list order without FMA contraction (bit-identical to the numpy loop). */
int qb_mix(qb_state* const* states, const double* probs, int count, qb_state** out);
/* same over the branch axis of one batched state -> single-branch state */
int qb_mix_branches(qb_state* s, const double* probs, qb_state** out);
/* ket -> density matrix: conj=0 gives psi psi^T as density.ketToDensity does (density.py:31-32),
 * conj=1 gives psi psi^dagger */
int qb_outer(qb_state* ket, int conj, qb_state** out);
/* copy one state into every branch of a batched state (fan-out before a batched gate) */
int qb_broadcast(qb_state* src, qb_state* dst_batched);

/* ---- sharded kets (one process per GPU) ----------------------------------
 * The reference has no distributed path; these are the device-side pieces of the global-qubit
 * swap that replaces genSwapGate + applyGate (qgates.py:77-133, 278-279) when the exchanged
 * qubit is one of the index bits that select the GPU.  The host side (qbot_b200/sharded.py)
 * owns the qubit map and the rendezvous (torch.distributed); this library owns the memory,
 * the pack kernel and the peer mappings.
 *
 * raw device buffers (cudaMalloc, so that they can be exported over CUDA IPC) */
int qb_buffer_alloc(int device, size_t bytes, void** out_dev);
int qb_buffer_free(int device, void* dev);
/* re-point a handle made by qb_create_external at another buffer of the same size */
int qb_rebind(qb_state* s, void* amplitudes_dev);
/* 64-byte CUDA IPC handle of a qb_buffer_alloc buffer / mapping of a peer's buffer */
int qb_ipc_export(int device, void* dev, void* handle64);
int qb_ipc_open(int device, const void* handle64, void** out_dev);
int qb_ipc_close(int device, void* dev);
/* One pass over a single-branch ket: dst index j takes the amplitude at source index
 * sum_d bit_d(j) << src_bit_of_dst_bit[d]  (a permutation of the index bits), and is written to
 * chunk_dst[j >> (nbits - chunk_bits)][j & (2^(nbits-chunk_bits) - 1)].  The chunk pointers
 * may be local (the second shard buffer, followed by an NCCL all-to-all) or IPC mappings of
 * peer buffers (the exchange then happens inside this kernel, as stores over NVLink).
 * Chunks are visited interleaved (64 KB units round-robin over the chunks) starting with chunk
 * `first_chunk` (unit w goes to chunk (w mod 2^chunk_bits) ^ first_chunk): a rank passes its own
 * chunk number, which makes concurrent ranks follow a pairwise-exchange schedule -- no two of
 * them store to the same peer at the same time. */
int qb_permute_scatter(qb_state* s, const int* src_bit_of_dst_bit, int chunk_bits, void* const* chunk_dst, int first_chunk);
/* The same pass over a SUB-BLOCK of the source, on a stream of the caller's: the source index bits of
 * `src_fixed_mask` are held at `src_fixed_value`, the other ndst_bits = nbits - popcount(mask) bits are
 * permuted as above (src_bit_of_dst_bit has ndst_bits entries; chunks are 2^(ndst_bits - chunk_bits)
 * amplitudes).  Lets a global-qubit exchange be issued in pieces that overlap with the sweeps of the
 * pieces already delivered (qbot_b200/sharded.py).  cuda_stream NULL = the handle's stream; otherwise the
 * caller orders that stream against the handle's work with events (qb_compute_stream). */
int qb_permute_scatter_sub(qb_state* s, const int* src_bit_of_dst_bit, int ndst_bits, uint64_t src_fixed_mask, uint64_t src_fixed_value,
                           int chunk_bits, void* const* chunk_dst, int first_chunk, void* cuda_stream, int max_ctas);
/* The queued gates as a plan whose steps (fused sweeps, one-gate kernels) the CALLER runs, in order, each either over
 * the whole state or sub-block by sub-block -- what lets the pieces of a global-qubit exchange leave (arrive) between the
 * sub-block runs of the last (first) sweeps of a pass:
 *   qb_plan_queue   plans the queue (it is empty afterwards); *nsteps steps; the first *head and the last *tail of them
 *                   may run on 2^park_bits contiguous sub-blocks one at a time: they are specialised sweeps whose tiles
 *                   do not contain the top park_bits index bits (controls / diagonals there are fine: the kernel runs a
 *                   tile range and still sees the true index bits)
 *   qb_run_steps    steps [from, to) on part `part` of `nparts` (nparts = 1: the whole state); sm_limit > 0 sizes the
 *                   persistent grids for that many SMs (the rest is left to an exchange kernel running beside them)
 *   qb_finish_queue checks that every step has run on every part, releases the plan
 * Every step must run exactly once on every amplitude, steps in ascending order per amplitude. */
int qb_plan_queue(qb_state* s, int park_bits, int* nsteps, int* head, int* tail);
int qb_run_steps(qb_state* s, int from, int to, int part, int nparts, int sm_limit);
int qb_finish_queue(qb_state* s);
/* max_ctas > 0 caps the grid (256-thread CTAs), so that the kernel leaves SM slots to sweeps running beside it.
 * qb_set_sm_limit is the other half: the handle's sweeps size their persistent grids for nsms SMs (0 = all). */
int qb_set_sm_limit(qb_state* s, int nsms);
/* Stream-ordered signals between the GPUs of a box.  `flags` are 8-byte counters in device memory, local or
 * peer-mapped (qb_buffer_alloc + qb_ipc_export / qb_ipc_open; zero them once).  qb_signal_flags stores `value`
 * into each of them after everything queued earlier on `cuda_stream` (NULL = the compute stream) has completed;
 * qb_wait_flags holds `cuda_stream` until every listed counter is >= value (bounded: after ~60 s the wait gives
 * up and qb_flag_timeouts counts it -- a lost peer must not hang the GPU). */
int qb_signal_flags(int device, void* cuda_stream, void* const* flags, int n, uint64_t value);
int qb_wait_flags(int device, void* cuda_stream, void* const* flags, int n, uint64_t value);
int qb_flag_timeouts(int device, uint64_t* count);
/* Asynchronous device-to-device copy on `cuda_stream` (NULL = the compute stream); dst / src may be peer mappings
 * (qb_ipc_open): the copy engines carry it over NVLink while the SMs keep running sweeps. */
int qb_copy_async(int device, void* dst, const void* src, size_t bytes, void* cuda_stream);
/* The CUDA stream every handle of `device` works on unless it was given one of its own
 * (qb_create_external / qb_set_stream): for event-ordering foreign streams against the library's work. */
int qb_compute_stream(int device, void** cuda_stream_out);

/* ---- instrumentation ------------------------------------------------------------------- */
int qb_get_stats(const qb_state* s, qb_stats* out);
int qb_reset_stats(qb_state* s);
int qb_timer_start(qb_state* s);                 /* CUDA event on the handle's stream */
int qb_timer_stop(qb_state* s, float* ms);       /* records, synchronises, returns elapsed ms */

#ifdef __cplusplus
}
#endif
#endif /* QBOT_B200_H */
