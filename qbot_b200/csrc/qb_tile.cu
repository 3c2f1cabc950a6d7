// qbot_b200 -- fused multi-gate sweep: persistent register-IO tile kernel + host engine.
//
// One launch = one sweep = one read + one write of the whole state (32 * 2^n bytes of HBM
// traffic) during which every gate of the sweep's program is applied (qb_plan.h describes
// tiles / stages / ops).  A CTA of 2^(M-4) threads is persistent over tiles
// t = blockIdx.x + i * gridDim.x and keeps a tile's 2^M amplitudes in REGISTERS, 16 per thread:
//
//   HBM --16 x LDG.128 per thread (each warp access = one 512-byte run)--> registers
//        stage 0 ops | STS.128 -> bar -> LDS.128 | stage 1 ops | ... | last stage ops
//   registers --16 x STG.128 per thread (512-byte runs)--> HBM
//
// Shared memory is touched only by the transposition between two stages (padded layout
// qt_slot, conflict-free for any choice of register bits), never for staging the HBM traffic.
// Two (M = 12) or four (M = 11) CTAs are resident per SM, so that one CTA's loads, barriers and
// stores overlap the other's arithmetic; the next tile of every CTA is prefetched into L2 with
// cp.async.bulk.prefetch while the current one is computed.  Everything that does not depend on
// the tile is computed once per launch: the per-stage thread index / shared-memory slot of every
// thread and a 64-bit mask of the ops whose thread-local predicate holds for this thread.
#include "qb_engine.h"
#include "qb_plan.h"
#include "qb_tile_ops.h"
#include "qb_jit_rt.h"

#include <algorithm>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <thread>

template <int M> struct TileCfg {
    static constexpr int T = 1 << (M - QT_R);
    static constexpr int NH = M - QT_L;
    static constexpr int CTAS_PER_SM = (M == 12) ? 2 : 4;
    static constexpr int TILE_BYTES = QT_TILE_UNITS(M) * 16;
    // per-thread tables: lbase + slot per stage (uint16 each), HBM offset of the thread in the
    // first / last stage (uint64 each)
    static constexpr int TABLE_BYTES = 2 * QT_MAX_STAGES * T * 2 + 2 * T * 8;
    static constexpr int SMEM_BYTES = TILE_BYTES + QT_MAX_PROGRAM_BYTES + TABLE_BYTES + 64;
};

__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// Gray-code walk over the 16 register indices: gray(j) and the bit flipped between j-1 and j
__device__ __forceinline__ constexpr int gray_of(int j) { return j ^ (j >> 1); }
__device__ __forceinline__ constexpr int gray_flip(int j) {        // j >= 1
    const int d = gray_of(j) ^ gray_of(j - 1);
    return d == 1 ? 0 : d == 2 ? 1 : d == 4 ? 2 : 3;
}

template <int M>
__global__ void __launch_bounds__(1 << (M - QT_R), (M == 12) ? 2 : 4)
k_tile_sweep(cplx* __restrict__ psi, const uint8_t* __restrict__ prog_dev, uint64_t ntiles, int prefetch) {
    using Cfg = TileCfg<M>;
    constexpr int T = Cfg::T, NH = Cfg::NH;
    extern __shared__ __align__(128) uint8_t smem[];
    cplx* buf = (cplx*)smem;
    uint8_t* prog = smem + Cfg::TILE_BYTES;
    uint16_t* thr_lbase = (uint16_t*)(prog + QT_MAX_PROGRAM_BYTES);
    uint16_t* thr_slot = thr_lbase + QT_MAX_STAGES * T;
    uint64_t* thr_gin = (uint64_t*)(thr_slot + QT_MAX_STAGES * T);
    uint64_t* thr_gout = thr_gin + T;

    const int tid = threadIdx.x;
    {
        const uint32_t total = ((const QtHeader*)prog_dev)->total_bytes;
        for (uint32_t i = tid * 16; i < total; i += T * 16) *(uint4*)(prog + i) = *(const uint4*)(prog_dev + i);
    }
    __syncthreads();
    const QtHeader* h = (const QtHeader*)prog;
    const QtStage* stages = (const QtStage*)(prog + QT_STAGES_OFF);
    const QtOp* ops = (const QtOp*)(prog + QT_OPS_OFF);
    const double* pool = (const double*)(prog + QT_POOL_OFF);
    const int nstages = h->nstages;

    // ---- tile-independent per-thread state ----------------------------------------------------
    uint64_t lok = 0;                                   // ops whose thread-local predicate holds here
    for (int s = 0; s < nstages; s++) {
        const QtStage& st = stages[s];
        const uint32_t lb = qt_thread_lbase(st, (uint32_t)tid, M);
        thr_lbase[s * T + tid] = (uint16_t)lb;
        thr_slot[s * T + tid] = (uint16_t)qt_slot(lb);
        for (int o = 0; o < st.nops; o++)
            if (qt_op_local_ok(ops[st.first_op + o], lb)) lok |= 1ull << (st.first_op + o);
        // HBM addressing of the first / last stage: lane = low bits, the rest through the run offsets
        if (s == 0) thr_gin[tid] = (lb & 31u) + qt_run_offset(lb >> QT_L, h->hb, NH);
        if (s == nstages - 1) thr_gout[tid] = (lb & 31u) + qt_run_offset(lb >> QT_L, h->hb, NH);
    }
    __syncthreads();

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t tbase = qt_tile_base(tile, h->hbs, NH);
        cplx a[QT_NR];
        {
            const QtStage& s0 = stages[0];
            const cplx* p = psi + tbase + thr_gin[tid];
            a[0] = __ldcs(p);
#pragma unroll
            for (int j = 1; j < QT_NR; j++) {
                const int q = gray_flip(j), g = gray_of(j);
                const int64_t step = (int64_t)1 << h->hb[s0.rb[q] - QT_L];
                p += ((g >> q) & 1) ? step : -step;
                a[g] = __ldcs(p);
            }
        }
        if (prefetch && tid < (1 << NH)) {
            const uint64_t nt = tile + gridDim.x;
            if (nt < ntiles)
                l2_prefetch_bulk(psi + qt_tile_base(nt, h->hbs, NH) + qt_run_offset((uint32_t)tid, h->hb, NH), 512);
        }

        for (int s = 0;; s++) {
            const QtStage& st = stages[s];
            const uint32_t lbase = thr_lbase[s * T + tid];
            const int first = st.first_op, nops = st.nops;
            for (int o = 0; o < nops; o++) {
                const int oi = first + o;
                if (!((lok >> oi) & 1ull)) continue;
                const QtOp& op = ops[oi];
                if ((op.flags & QT_FLAG_GLOBAL) && !qt_op_global_ok(op, tbase)) continue;
                qt_apply_op(a, op, pool, lbase, tbase);
            }
            if (s + 1 == nstages) break;
            // ---- transposition: this stage's register bits out, the next stage's in -------------
            {
                cplx* p = buf + thr_slot[s * T + tid];
                if (s == 0) __syncthreads();          // the previous tile's last reads are done
                p[0] = a[0];
#pragma unroll
                for (int j = 1; j < QT_NR; j++) {
                    const int q = gray_flip(j), g = gray_of(j);
                    const int step = (int)qt_slot(1u << st.rb[q]);
                    p += ((g >> q) & 1) ? step : -step;
                    *p = a[g];
                }
            }
            __syncthreads();
            {
                const QtStage& sn = stages[s + 1];
                const cplx* p = buf + thr_slot[(s + 1) * T + tid];
                a[0] = p[0];
#pragma unroll
                for (int j = 1; j < QT_NR; j++) {
                    const int q = gray_flip(j), g = gray_of(j);
                    const int step = (int)qt_slot(1u << sn.rb[q]);
                    p += ((g >> q) & 1) ? step : -step;
                    a[g] = *p;
                }
            }
        }
        {
            const double scale = h->scale;
            if (scale != 1.0) qt_scale_real(a, scale);
            const QtStage& s1 = stages[nstages - 1];
            cplx* p = psi + tbase + thr_gout[tid];
            __stcs(p, a[0]);
#pragma unroll
            for (int j = 1; j < QT_NR; j++) {
                const int q = gray_flip(j), g = gray_of(j);
                const int64_t step = (int64_t)1 << h->hb[s1.rb[q] - QT_L];
                p += ((g >> q) & 1) ? step : -step;
                __stcs(p, a[g]);
            }
        }
    }
}

// ---- host engine --------------------------------------------------------------------------------
namespace {

struct CachedPlan {
    uint64_t hash = 0;
    uint64_t hash2 = 0;                // second, independent fingerprint of the same gate list (a hit needs both)
    int pins = 0;                      // handles that run this plan step by step right now (qb_plan_queue): not evictable
    size_t ngates = 0;
    int nbits = 0;
    int R = QT_R;                      // register bits per stage the plan was built for (5: specialised kernels only)
    std::vector<QGate> gates;          // the gate list the plan was built from (after the peephole rewrite)
    uint64_t uses = 0;
    bool upgrade_failed = false;       // building / compiling the 32-amplitudes-per-thread variant failed once
    std::vector<QtPlanStep> steps;
    std::vector<size_t> prog_off;      // per step offset into dev (fused steps)
    uint8_t* dev = nullptr;
    // specialised (NVRTC-compiled) kernel of each fused step, resolved when its STRUCTURE has been
    // seen before (in this plan, an earlier plan or another state)
    struct StepJit {
        uint64_t key = 0;              // hash of the specialised source (0: not generated yet)
        QbJitKernel kv;                // the virtual-basis variant (first sweep after qb_init_basis), compiled on first use
        bool kv_ready = false, kv_failed = false;
        bool ready = false, failed = false;
        QbJitKernel k;
        std::vector<double> pool;      // run-time coefficients (host copy, passed as kernel parameters)
        double* pool_dev = nullptr;    // device copy when they do not fit the parameter space
    };
    std::vector<StepJit> jit;
    void release() {
        if (dev) cudaFree(dev);
        dev = nullptr;
        for (auto& j : jit) if (j.pool_dev) { cudaFree(j.pool_dev); j.pool_dev = nullptr; }
    }
};

// One plan cache per DEVICE and process (not per state handle): a program that is run again
// on a fresh register -- every executeTxt call creates one -- finds its plans, device-side sweep
// programs and specialised kernels ready.
struct EngineState {
    std::list<CachedPlan> cache;       // most recent first
    bool attr_set[16] = {false};
};
constexpr size_t kMaxCachedPlans = 256;     // (32 thrashed once one process ran all the configs: a 28-bit plan takes ~0.1 s to search again)
std::mutex g_engine_mu;                // guards the caches and the lazily resolved StepJit records
std::map<int, EngineState> g_engines;

uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

uint64_t hash_gates(const std::vector<QGate>& gates, int nbits, int M, uint64_t seed = 1469598103934665603ull) {
    uint64_t h = seed;
    h = fnv(h, &nbits, sizeof(nbits));
    h = fnv(h, &M, sizeof(M));
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

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int engine_M() {
    static int m = [] {
        int v = env_int("QBOT_B200_TILE_M", 12);
        return (v == 11 || v == 12) ? v : 12;
    }();
    return m;
}

int engine_prefetch() {
    static int p = env_int("QBOT_B200_L2_PREFETCH", 1);
    return p;
}
// specialised sweeps: bit 0 = L2 prefetch of the next tile, bit 1 = asynchronous copy of the next
// tile into the transposition buffer during the last stage (env QBOT_B200_ASYNC_LOAD=0 disables)
int engine_jit_prefetch() {
    // measured at 30 qubits (ms per sweep): bulk copy only 7.44 | bulk copy + L2 prefetch 7.80 |
    // L2 prefetch + LDG 7.96 | plain LDG 10.1
    const int async = env_int("QBOT_B200_ASYNC_LOAD", 1) ? 2 : 0;
    const int l2 = env_int("QBOT_B200_JIT_L2_PREFETCH", async ? 0 : 1) ? 1 : 0;
    return async | l2;
}

// 0: never specialise; 1 (default): specialise a plan when it is run again on a state of at least
// QBOT_B200_JIT_MIN_BITS index bits (compile time is amortised by repetition); 2: always
int engine_jit_default() {
    static int m = env_int("QBOT_B200_JIT", 1);
    return m;
}
int engine_jit_min_bits() {
    static int b = env_int("QBOT_B200_JIT_MIN_BITS", 22);
    return b;
}
// plan shape of specialised sweeps: 5 register bits per stage (32 amplitudes per thread, 128-thread
// CTAs, 3 per SM: fewer shared-memory transpositions, smaller barrier domains) or 4 (same plan as the
// generic kernel's)
int engine_jit_R() {
    return env_int("QBOT_B200_JIT_R", QT_MAXR) == QT_R ? QT_R : QT_MAXR;      // read every time: tests flip it
}

// decide (once per step and run, until resolved) whether fused step i runs specialised
bool step_jit(qb_state* s, CachedPlan* plan, size_t i, int jit_mode) {
    CachedPlan::StepJit& j = plan->jit[i];
    if (j.ready) return true;
    if (j.failed) return false;
    // size that decides whether specialising pays: index bits of a branch plus the branch axis (a 4096 x 16-qubit
    // batch is a 28-bit sweep; counting it as 20 bits kept config 4 on the generic kernel at 0.25 of HBM peak)
    const int total_bits = s->nbits + (s->nbranch > 1 ? 63 - __builtin_clzll((unsigned long long)s->nbranch) : 0);
    if (jit_mode == 1 && total_bits < engine_jit_min_bits()) return false;
    const uint8_t* program = plan->steps[i].program.data();
    try {
        if (!j.key) j.key = qj_hash(qb_jit_full_source(program, nullptr, nullptr));
        const int seen = qb_jit_note(j.key);          // sightings of this structure so far (this one included) or -1 if compiled
        if (jit_mode != 2 && seen >= 0 && seen < 2) return false;
        j.k = qb_jit_get(program, s->device);
        j.pool = qj_pool(program);
        if (j.k.pool_global) {
            QB_CUDA(cudaMalloc((void**)&j.pool_dev, j.pool.size() * sizeof(double)));
            QB_CUDA(cudaMemcpyAsync(j.pool_dev, j.pool.data(), j.pool.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        }
        j.ready = true;
    } catch (const qb_error&) {
        j.failed = true;
        if (jit_mode == 2) throw;       // explicitly requested: fail loudly
    }
    return j.ready;
}

template <int M>
void launch_sweep(qb_state* s, EngineState* es, const uint8_t* prog, uint64_t ntiles) {
    using Cfg = TileCfg<M>;
    if (!es->attr_set[M]) {
        QB_CUDA(cudaFuncSetAttribute(k_tile_sweep<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        QB_CUDA(cudaFuncSetAttribute(k_tile_sweep<M>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        es->attr_set[M] = true;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)s->sms * Cfg::CTAS_PER_SM);
    k_tile_sweep<M><<<grid, Cfg::T, Cfg::SMEM_BYTES, s->stream>>>(s->d, prog, ntiles, engine_prefetch());
}

bool qb_engine_available_impl() { return getenv("QBOT_B200_NO_FUSION") == nullptr; }

// the cached plan of `gates` for R register bits per stage (planning it on a miss)
CachedPlan* get_plan(qb_state* s, EngineState* es, const std::vector<QGate>& gates, int M, int R) {
    uint64_t hsh = hash_gates(gates, s->nbits, M);
    hsh = fnv(hsh, &R, sizeof(R));
    const uint64_t hsh2 = hash_gates(gates, s->nbits, M, 0x9e3779b97f4a7c15ull);
    for (auto it = es->cache.begin(); it != es->cache.end(); ++it) {
        if (it->hash == hsh && it->hash2 == hsh2 && it->ngates == gates.size() && it->nbits == s->nbits && it->R == R) {
            es->cache.splice(es->cache.begin(), es->cache, it);
            return &es->cache.front();
        }
    }
    CachedPlan cp;
    cp.hash = hsh; cp.hash2 = hsh2; cp.ngates = gates.size(); cp.nbits = s->nbits; cp.R = R;
    QtPlanOptions opt;
    opt.M = M;
    opt.R = R;
    // plan search effort grows with the cost of a sweep (env QBOT_B200_PLAN_TRIALS overrides)
    const int total_bits = s->nbits + (s->nbranch > 1 ? 63 - __builtin_clzll((unsigned long long)s->nbranch) : 0);
    // (the deepest level, ~0.2 s, only for plans of the specialiser's shape: those are made in the background or on
    // request, and run many times; the generic first-sight plan must not wait for it)
    opt.search_trials = env_int("QBOT_B200_PLAN_TRIALS", total_bits >= 29 && R == QT_MAXR ? 128 : total_bits >= 27 ? 32 : total_bits >= 23 ? 8 : 1);
    cp.steps = qt_plan_best(gates, s->nbits, opt, &cp.gates);
    size_t total = 0;
    cp.prog_off.resize(cp.steps.size(), 0);
    for (size_t i = 0; i < cp.steps.size(); i++) {
        if (!cp.steps[i].fused) continue;
        cp.prog_off[i] = total;
        total += (cp.steps[i].program.size() + 255) & ~size_t(255);
    }
    if (total && R == QT_R) {          // device copies of the programs: read by the generic kernel only
        std::vector<uint8_t> hostbuf(total, 0);
        for (size_t i = 0; i < cp.steps.size(); i++)
            if (cp.steps[i].fused) memcpy(hostbuf.data() + cp.prog_off[i], cp.steps[i].program.data(), cp.steps[i].program.size());
        QB_CUDA(cudaMalloc((void**)&cp.dev, total));
        QB_CUDA(cudaMemcpyAsync(cp.dev, hostbuf.data(), total, cudaMemcpyHostToDevice, s->stream));
        QB_CUDA(cudaStreamSynchronize(s->stream));     // hostbuf dies here
    }
    if (es->cache.size() >= kMaxCachedPlans) {
        // the least recently used plan that no handle is in the middle of running (qb_plan_queue pins its plan)
        auto victim = es->cache.end();
        for (auto it = es->cache.end(); it != es->cache.begin();) {
            --it;
            if (it->pins == 0) { victim = it; break; }
        }
        if (victim != es->cache.end()) {
            QB_CUDA(cudaStreamSynchronize(s->stream));
            victim->release();
            es->cache.erase(victim);
        }
    }
    cp.jit.assign(cp.steps.size(), CachedPlan::StepJit());
    es->cache.push_front(std::move(cp));
    return &es->cache.front();
}

// First sight of a gate list on a large state (default specialisation mode): the generic kernel
// runs it now; meanwhile a host thread plans the specialiser's variant and NVRTC-compiles its
// kernels into the process-wide cache, so that a second run starts specialised without waiting.
// Pure host work (planner + NVRTC); at most two such jobs at a time; joined at process exit.
struct Warmers {
    std::mutex mu;
    std::vector<std::thread> threads;
    std::vector<uint64_t> started;
    int running = 0;
    ~Warmers() {
        for (auto& t : threads) if (t.joinable()) t.join();
    }
};
Warmers g_warmers;

// the upgrade path waits for a warm-up of the same gate list instead of compiling the same kernels twice
void wait_for_warm(uint64_t plan_hash) {
    std::thread t;
    {
        std::lock_guard<std::mutex> lk(g_warmers.mu);
        for (size_t i = 0; i < g_warmers.started.size() && i < g_warmers.threads.size(); i++)
            if (g_warmers.started[i] == plan_hash && g_warmers.threads[i].joinable()) { t.swap(g_warmers.threads[i]); break; }
    }
    if (t.joinable()) t.join();
}

void warm_in_background(const std::vector<QGate>& gates, int nbits, int total_bits, int M, uint64_t plan_hash, bool virtual_first) {
    if (!qb_jit_available(nullptr)) return;
    std::lock_guard<std::mutex> lk(g_warmers.mu);
    if (g_warmers.running >= 2) return;
    if (std::find(g_warmers.started.begin(), g_warmers.started.end(), plan_hash) != g_warmers.started.end()) return;
    g_warmers.started.push_back(plan_hash);
    g_warmers.running++;
    const int trials = env_int("QBOT_B200_PLAN_TRIALS", total_bits >= 29 ? 128 : total_bits >= 27 ? 32 : total_bits >= 23 ? 8 : 1);   // as get_plan
    g_warmers.threads.emplace_back([gates, nbits, M, trials, virtual_first]() {
        try {
            QtPlanOptions opt;
            opt.M = M;
            opt.R = QT_MAXR;
            opt.search_trials = trials;
            std::vector<QGate> planned;
            const std::vector<QtPlanStep> steps = qt_plan_best(gates, nbits, opt, &planned);
            std::vector<const uint8_t*> progs;
            for (const QtPlanStep& st : steps) if (st.fused) progs.push_back(st.program.data());
            if (progs.size() == 1) qb_jit_compile_cached(progs[0]);
            else if (!progs.empty()) qb_jit_precompile(progs);
            // the gate list came right behind a basis-state initialisation: a run on a fresh register will start its
            // first sweep from the known state (virtual-basis variant of that sweep)
            if (virtual_first && !steps.empty() && steps[0].fused) qb_jit_compile_cached(steps[0].program.data(), true);
        } catch (...) {
        }
        std::lock_guard<std::mutex> lk2(g_warmers.mu);
        g_warmers.running--;
    });
}

// all kernels of a plan that is going to run specialised, compiled in parallel on the host cores
void precompile_plan(CachedPlan* plan) {
    std::vector<const uint8_t*> progs;
    for (size_t i = 0; i < plan->steps.size(); i++)
        if (plan->steps[i].fused && !plan->jit[i].ready && !plan->jit[i].failed) progs.push_back(plan->steps[i].program.data());
    if (progs.size() >= 2) qb_jit_precompile(progs);
}

}  // namespace

bool qb_engine_available() { return qb_engine_available_impl(); }

void qb_engine_free(qb_state* s) {          // plans belong to the device, not to the handle
    if (s->pending_plan) {              // destroyed in the middle of a planned queue: its plan may be evicted again
        std::lock_guard<std::mutex> lk(g_engine_mu);
        ((CachedPlan*)s->pending_plan)->pins--;
        s->pending_plan = nullptr;
    }
    s->engine = nullptr;
}

// plan (or find the cached plan of) `gates` and decide which executor runs it
static CachedPlan* engine_prepare(qb_state* s, const std::vector<QGate>& gates, int* jit_mode_out) {
    EngineState* es = &g_engines[s->device];
    const int M = engine_M();
    int jit_mode = s->jit_mode >= 0 ? s->jit_mode : engine_jit_default();
    const bool tiled = s->nbits >= M;
    const bool big = s->nbits + (s->nbranch > 1 ? 4 : 0) >= engine_jit_min_bits();

    // Which plan runs: the generic 16-amplitudes-per-thread plan (any executor), or -- when the
    // sweeps are going to be specialised anyway -- the specialiser's own 32-amplitudes-per-thread
    // plan.  Mode 2 takes the latter at first sight; mode 1 when the SAME gate list is flushed a
    // second time on a large state (every kernel of the variant is compiled before anything is
    // launched, so a failure leaves the generic plan in charge).
    CachedPlan* plan = nullptr;
    if (tiled && jit_mode == 2 && engine_jit_R() == QT_MAXR) {
        plan = get_plan(s, es, gates, M, QT_MAXR);
        precompile_plan(plan);
    } else {
        plan = get_plan(s, es, gates, M, QT_R);
        plan->uses++;
        if (tiled && jit_mode == 1 && big && engine_jit_R() == QT_MAXR && plan->uses == 1 && env_int("QBOT_B200_JIT_BACKGROUND", 1))
            warm_in_background(gates, s->nbits, s->nbits + (s->nbranch > 1 ? 63 - __builtin_clzll((unsigned long long)s->nbranch) : 0), M,
                               plan->hash, s->virt);     // a second run will find its kernels compiled
        if (tiled && jit_mode == 1 && big && engine_jit_R() == QT_MAXR && plan->uses >= 2 && !plan->upgrade_failed) {
            CachedPlan* base = plan;
            try {
                wait_for_warm(base->hash);
                CachedPlan* p5 = get_plan(s, es, gates, M, QT_MAXR);
                precompile_plan(p5);
                for (size_t i = 0; i < p5->steps.size(); i++)
                    if (p5->steps[i].fused && !step_jit(s, p5, i, 2)) throw qb_error(-2, "specialised variant unavailable");
                plan = p5;
            } catch (const qb_error&) {
                base->upgrade_failed = true;
                plan = base;
            }
        }
    }
    if (plan->R != QT_R) jit_mode = 2;          // only specialised kernels can run this plan shape
    *jit_mode_out = jit_mode;
    return plan;
}

// steps [from, to) of a prepared plan on part `part` of `nparts` equal, contiguous sub-blocks of the state (nparts = 1: all
// of it).  Only specialised sweeps can run on a sub-block (the kernel takes a tile range; its predicates still see the true
// index bits) and only when the bits that select the sub-block -- the top log2(nparts) index bits -- are not tile bits.
static void engine_run_steps(qb_state* s, CachedPlan* plan, int jit_mode, size_t from, size_t to, int part, int nparts, int sms) {
    EngineState* es = &g_engines[s->device];
    const int M = engine_M();
    const uint64_t ntiles = s->total() >> M;
    const uint64_t t0 = ntiles / (uint64_t)nparts * (uint64_t)part, t1 = ntiles / (uint64_t)nparts * (uint64_t)(part + 1);
    const bool counts = part == nparts - 1;          // a pass over the state is complete with its last part
    for (size_t i = from; i < to && i < plan->steps.size(); i++) {
        const QtPlanStep& st = plan->steps[i];
        if (s->virt) {
            // the register is still the untouched basis state: a specialised sweep over the whole state starts from the
            // known amplitudes (no load); every other kind of step needs them in memory first
            bool done = false;
            if (st.fused && nparts == 1 && jit_mode != 0 && step_jit(s, plan, i, jit_mode)) {
                CachedPlan::StepJit& j = plan->jit[i];
                if (!j.kv_ready && !j.kv_failed) {
                    try { j.kv = qb_jit_get(st.program.data(), s->device, true); j.kv_ready = true; }
                    catch (const qb_error&) { j.kv_failed = true; }
                }
                if (j.kv_ready) {
                    qb_jit_launch(j.kv, s->stream, sms, s->d, t0, t1, 0, j.pool.data(), j.pool_dev, s->virt_index);
                    s->virt = false;
                    s->stats.jit_passes++;
                    s->stats.kernel_launches++;
                    s->stats.fused_passes++;
                    s->stats.state_passes++;
                    s->stats.fused_gates += st.ngates;
                    s->stats.gates_applied += st.ngates;
                    s->stats.bytes_moved += (uint64_t)s->bytes();          // written once, never read
                    done = true;
                }
            }
            if (done) continue;
            s->materialize();
        }
        if (!st.fused) {
            if (nparts != 1) throw qb_error(-1, "a one-gate step cannot run on a sub-block");
            s->run_gate_unfused(plan->gates[st.gate_index]);
            continue;
        }
        if (jit_mode != 0 && step_jit(s, plan, i, jit_mode)) {
            const CachedPlan::StepJit& j = plan->jit[i];
            if (plan->R != QT_R && !j.ready) throw qb_error(-2, "no specialised kernel for a 32-amplitudes-per-thread sweep");
            qb_jit_launch(j.k, s->stream, sms, s->d, t0, t1, engine_jit_prefetch(), j.pool.data(), j.pool_dev);
            if (counts) { s->stats.jit_passes++; s->stats.jit_kernel_hash += j.key; }
        } else {
            if (nparts != 1) throw qb_error(-1, "the generic sweep kernel cannot run on a sub-block");
            if (plan->R != QT_R) throw qb_error(-2, "the generic sweep kernel cannot run a 32-amplitudes-per-thread plan");
            const uint8_t* prog = plan->dev + plan->prog_off[i];
            {   // a stale error of an earlier unchecked runtime call must not be blamed on this launch
                const cudaError_t stale = cudaGetLastError();
                if (stale != cudaSuccess) throw qb_error(-2, std::string("stale CUDA error before the sweep launch: ") + cudaGetErrorString(stale));
            }
            if (M == 11) launch_sweep<11>(s, es, prog, ntiles);
            else launch_sweep<12>(s, es, prog, ntiles);
            QB_CUDA(cudaGetLastError());
        }
        s->stats.kernel_launches++;
        if (counts) {
            s->stats.fused_passes++;
            s->stats.state_passes++;
            s->stats.fused_gates += st.ngates;
            s->stats.gates_applied += st.ngates;
            s->stats.bytes_moved += (uint64_t)s->bytes() * 2;
        }
    }
}

void qb_engine_run(qb_state* s, const std::vector<QGate>& gates) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    int jit_mode = 0;
    CachedPlan* plan = engine_prepare(s, gates, &jit_mode);
    engine_run_steps(s, plan, jit_mode, 0, plan->steps.size(), 0, 1, s->sms);
}

// ---- a plan kept pending on the handle, its steps run by the caller range by range (pipelined exchanges) ----------
int qb_engine_plan_pending(qb_state* s, const std::vector<QGate>& gates, int park_bits, int* head, int* tail) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    int jit_mode = 0;
    s->materialize();
    CachedPlan* plan = engine_prepare(s, gates, &jit_mode);
    s->pending_plan = plan;
    plan->pins++;
    s->pending_jit = jit_mode;
    const size_t n = plan->steps.size();
    s->pending_done.assign(n, 0u);
    const int M = engine_M();
    // a step can run sub-block by sub-block when it is a specialised sweep whose tile leaves the parked (top) bits alone
    std::vector<char> ok(n, 0);
    for (size_t i = 0; i < n; i++) {
        const QtPlanStep& st = plan->steps[i];
        if (!st.fused || s->nbranch != 1 || park_bits <= 0 || s->nbits - park_bits < M) continue;
        if (jit_mode == 0 || !step_jit(s, plan, i, jit_mode)) continue;
        const QtHeader* h = (const QtHeader*)st.program.data();
        bool clear = true;
        for (int x = 0; x < (int)h->M - QT_L; x++) if (h->hb[x] >= s->nbits - park_bits) clear = false;
        ok[i] = clear;
    }
    int hd = 0, tl = 0;
    while ((size_t)hd < n && ok[hd]) hd++;
    while ((size_t)tl < n && ok[n - 1 - tl]) tl++;
    if (head) *head = hd;
    if (tail) *tail = tl;
    return (int)n;
}

void qb_engine_run_pending(qb_state* s, int from, int to, int part, int nparts, int sms) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    CachedPlan* plan = (CachedPlan*)s->pending_plan;
    if (!plan) throw qb_error(-1, "no pending plan (qb_plan_queue first)");
    if (from < 0 || to < from || (size_t)to > plan->steps.size() || nparts < 1 || nparts > 16 || part < 0 || part >= nparts)
        throw qb_error(-1, "run_steps: bad step or part range");
    for (int i = from; i < to; i++) {
        const uint32_t bits = nparts == 1 ? 0xffffu : (1u << part);
        if (s->pending_done[i] & bits) throw qb_error(-1, "run_steps: a step would run twice on the same amplitudes");
        s->pending_done[i] |= bits;
        s->pending_parts_of.resize(plan->steps.size(), 1);
        s->pending_parts_of[i] = nparts;
    }
    engine_run_steps(s, plan, s->pending_jit, (size_t)from, (size_t)to, part, nparts, sms > 0 && sms < s->sms ? sms : s->sms);
}

void qb_engine_finish_pending(qb_state* s) {
    std::lock_guard<std::mutex> lk(g_engine_mu);
    CachedPlan* plan = (CachedPlan*)s->pending_plan;
    if (!plan) return;
    s->pending_plan = nullptr;
    plan->pins--;
    for (size_t i = 0; i < plan->steps.size(); i++) {
        const int np = i < s->pending_parts_of.size() ? s->pending_parts_of[i] : 1;
        const uint32_t want = np == 1 ? 0xffffu : ((1u << np) - 1u);
        if (i >= s->pending_done.size() || (s->pending_done[i] & want) != want) {
            s->pending_done.clear(); s->pending_parts_of.clear();
            throw qb_error(-1, "finish_queue: not every step of the plan has run on every part of the state");
        }
    }
    s->pending_done.clear();
    s->pending_parts_of.clear();
}
