// qbot_b200 -- fused multi-gate sweep: persistent TMA-staged tile kernel + host engine.
//
// One launch = one sweep = one read + one write of the whole state (32 * 2^n bytes of HBM
// traffic) during which every gate of the sweep's program is applied (qb_plan.h describes
// tiles / stages / ops).  Per CTA (one per SM, persistent over tiles t = blockIdx.x + i*grid):
//
//   HBM --cp.async.bulk (128 x 512 B runs, mbarrier complete_tx)--> smem tile (3-deep ring)
//        for each stage:  smem -> registers (2^R amplitudes / thread), ops, registers -> smem
//   smem --cp.async.bulk.global (bulk_group)--> HBM
//
// Loads are issued two tiles ahead and stores drain asynchronously, so the TMA engine keeps
// HBM busy while all warps compute; the only generic-proxy global accesses are the program
// copy.  Shared-memory placement (qt_slot) keeps every 512-byte run contiguous for the bulk
// copies and skews runs so that any choice of register bits stays (mostly) bank-conflict free.
#include "qb_engine.h"
#include "qb_plan.h"
#include "qb_tile_ops.h"

#include <cstring>
#include <list>

#define QT_TILE_BYTES (QT_TILE_UNITS * 16)
#define QT_NBUF 3
#define QT_SMEM_BYTES (QT_NBUF * QT_TILE_BYTES + QT_MAX_PROGRAM_BYTES + QT_RUNS * 8 + 64)

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- the sweep kernel ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(1 << (QT_M - R), 1)
k_tile_sweep(cplx* __restrict__ psi, const uint8_t* __restrict__ prog_dev, uint64_t ntiles) {
    constexpr int NR = 1 << R;
    constexpr int RUN_BYTES = (1 << QT_L) * 16;
    constexpr int RUNS_PER_LANE = QT_RUNS / 32;
    extern __shared__ __align__(128) uint8_t smem[];
    cplx* bufs = (cplx*)smem;
    uint8_t* prog = smem + QT_NBUF * QT_TILE_BYTES;
    uint64_t* run_off = (uint64_t*)(prog + QT_MAX_PROGRAM_BYTES);
    uint64_t* full = run_off + QT_RUNS;

    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint32_t total = ((const QtHeader*)prog_dev)->total_bytes;
        for (uint32_t i = tid * 16; i < total; i += T * 16) *(uint4*)(prog + i) = *(const uint4*)(prog_dev + i);
        if (tid < QT_NBUF) mbar_init(&full[tid], 1);
    }
    __syncthreads();
    const QtHeader* h = (const QtHeader*)prog;
    const QtStage* stages = (const QtStage*)(prog + h->stages_off);
    const QtOp* ops = (const QtOp*)(prog + h->ops_off);
    const double* pool = (const double*)(prog + h->pool_off);
    if (tid < QT_RUNS) run_off[tid] = qt_run_offset((uint32_t)tid, h->hb);
    fence_mbar_init();
    __syncthreads();

    if (blockIdx.x >= ntiles) return;
    const uint64_t my_n = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int nstages = h->nstages;

    auto issue_load = [&](uint64_t it) {      // warp 0 only
        const int b = (int)(it % QT_NBUF);
        const uint64_t tb = qt_tile_base(blockIdx.x + it * gridDim.x, h->hb);
        if (lane == 0) mbar_arrive_expect_tx(&full[b], QT_RUNS * RUN_BYTES);
        __syncwarp();
        cplx* dst = bufs + (size_t)b * QT_TILE_UNITS;
#pragma unroll
        for (int r = 0; r < RUNS_PER_LANE; r++) {
            const uint32_t k = lane + 32 * r;
            bulk_load(dst + qt_slot(k << QT_L), psi + tb + run_off[k], RUN_BYTES, &full[b]);
        }
    };

    if (warp == 0) {
        issue_load(0);
        if (my_n > 1) issue_load(1);
    }

    for (uint64_t it = 0; it < my_n; it++) {
        const int b = (int)(it % QT_NBUF);
        const uint64_t tbase = qt_tile_base(blockIdx.x + it * gridDim.x, h->hb);
        cplx* buf = bufs + (size_t)b * QT_TILE_UNITS;
        mbar_wait(&full[b], (uint32_t)((it / QT_NBUF) & 1));

        for (int s = 0; s < nstages; s++) {
            const QtStage& st = stages[s];
            const uint32_t lbase = qt_thread_lbase<R>(st, (uint32_t)tid);
            cplx a[NR];
            uint32_t slot[NR];
#pragma unroll
            for (int i = 0; i < NR; i++) {
                slot[i] = qt_slot(lbase | qt_reg_offset<R>(st, i));
                a[i] = buf[slot[i]];
            }
            const int nops = st.nops;
            const QtOp* sop = ops + st.first_op;
            for (int o = 0; o < nops; o++) qt_apply_op<R>(a, sop[o], pool, lbase, tbase);
#pragma unroll
            for (int i = 0; i < NR; i++) buf[slot[i]] = a[i];
            if (s + 1 < nstages) __syncthreads();
        }
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int r = 0; r < RUNS_PER_LANE; r++) {
                const uint32_t k = lane + 32 * r;
                bulk_store(psi + tbase + run_off[k], buf + qt_slot(k << QT_L), RUN_BYTES);
            }
            bulk_commit();
            if (it + 2 < my_n) {
                bulk_wait_read<1>();          // the previous tile's stores no longer read their buffer
                __syncwarp();
                issue_load(it + 2);
            }
        }
    }
    if (warp == 0) bulk_wait<0>();
}

// ---- host engine --------------------------------------------------------------------------------
namespace {

struct CachedPlan {
    uint64_t hash = 0;
    size_t ngates = 0;
    int nbits = 0;
    std::vector<QtPlanStep> steps;
    std::vector<size_t> prog_off;      // per step offset into dev (fused steps)
    uint8_t* dev = nullptr;
};

struct EngineState {
    std::list<CachedPlan> cache;       // most recent first
    bool attr_set[5] = {false, false, false, false, false};
};

uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

uint64_t hash_gates(const std::vector<QGate>& gates, int nbits, int R) {
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, &nbits, sizeof(nbits));
    h = fnv(h, &R, sizeof(R));
    for (const QGate& g : gates) {
        h = fnv(h, &g.type, sizeof(int));
        h = fnv(h, &g.k, sizeof(int));
        h = fnv(h, g.tb, sizeof(int) * g.k);
        h = fnv(h, &g.cmask, sizeof(uint64_t));
        h = fnv(h, g.m.data(), sizeof(cplx) * g.m.size());
        if (!g.src.empty()) h = fnv(h, g.src.data(), sizeof(int) * g.src.size());
    }
    return h;
}

int engine_R() {
    static int r = [] {
        const char* e = getenv("QBOT_B200_TILE_R");
        int v = e ? atoi(e) : 4;
        return (v == 3 || v == 4) ? v : 4;
    }();
    return r;
}

}  // namespace

bool qb_engine_available() { return getenv("QBOT_B200_NO_FUSION") == nullptr; }

void qb_engine_free(qb_state* s) {
    EngineState* es = (EngineState*)s->engine;
    if (!es) return;
    for (auto& c : es->cache) if (c.dev) cudaFree(c.dev);
    delete es;
    s->engine = nullptr;
}

void qb_engine_run(qb_state* s, const std::vector<QGate>& gates) {
    if (!s->engine) s->engine = new EngineState();
    EngineState* es = (EngineState*)s->engine;
    const int R = engine_R();
    const uint64_t hsh = hash_gates(gates, s->nbits, R);
    CachedPlan* plan = nullptr;
    for (auto it = es->cache.begin(); it != es->cache.end(); ++it) {
        if (it->hash == hsh && it->ngates == gates.size() && it->nbits == s->nbits) {
            es->cache.splice(es->cache.begin(), es->cache, it);
            plan = &es->cache.front();
            break;
        }
    }
    if (!plan) {
        CachedPlan cp;
        cp.hash = hsh; cp.ngates = gates.size(); cp.nbits = s->nbits;
        QtPlanOptions opt;
        opt.R = R;
        cp.steps = qt_plan(gates, s->nbits, opt);
        size_t total = 0;
        cp.prog_off.resize(cp.steps.size(), 0);
        for (size_t i = 0; i < cp.steps.size(); i++) {
            if (!cp.steps[i].fused) continue;
            cp.prog_off[i] = total;
            total += (cp.steps[i].program.size() + 255) & ~size_t(255);
        }
        if (total) {
            std::vector<uint8_t> hostbuf(total, 0);
            for (size_t i = 0; i < cp.steps.size(); i++)
                if (cp.steps[i].fused) memcpy(hostbuf.data() + cp.prog_off[i], cp.steps[i].program.data(), cp.steps[i].program.size());
            QB_CUDA(cudaMalloc((void**)&cp.dev, total));
            QB_CUDA(cudaMemcpyAsync(cp.dev, hostbuf.data(), total, cudaMemcpyHostToDevice, s->stream));
            QB_CUDA(cudaStreamSynchronize(s->stream));     // hostbuf dies here
        }
        if (es->cache.size() >= 8) {
            QB_CUDA(cudaStreamSynchronize(s->stream));
            if (es->cache.back().dev) cudaFree(es->cache.back().dev);
            es->cache.pop_back();
        }
        es->cache.push_front(std::move(cp));
        plan = &es->cache.front();
    }

    const uint64_t ntiles = s->total() >> QT_M;
    for (size_t i = 0; i < plan->steps.size(); i++) {
        const QtPlanStep& st = plan->steps[i];
        if (!st.fused) {
            s->run_gate_unfused(gates[st.gate_index]);
            continue;
        }
        const int threads = 1 << (QT_M - R);
        const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)s->sms);
        const uint8_t* prog = plan->dev + plan->prog_off[i];
        if (R == 3) {
            if (!es->attr_set[3]) {
                QB_CUDA(cudaFuncSetAttribute(k_tile_sweep<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, QT_SMEM_BYTES));
                es->attr_set[3] = true;
            }
            k_tile_sweep<3><<<grid, threads, QT_SMEM_BYTES, s->stream>>>(s->d, prog, ntiles);
        } else {
            if (!es->attr_set[4]) {
                QB_CUDA(cudaFuncSetAttribute(k_tile_sweep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, QT_SMEM_BYTES));
                es->attr_set[4] = true;
            }
            k_tile_sweep<4><<<grid, threads, QT_SMEM_BYTES, s->stream>>>(s->d, prog, ntiles);
        }
        QB_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        s->stats.fused_passes++;
        s->stats.state_passes++;
        s->stats.fused_gates += st.ngates;
        s->stats.gates_applied += st.ngates;
        s->stats.bytes_moved += (uint64_t)s->bytes() * 2;
    }
}
