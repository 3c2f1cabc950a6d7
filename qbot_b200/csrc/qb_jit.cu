// qbot_b200 -- run time of the sweep specialiser: NVRTC-compiles the text qj_generate emits for a
// sweep program into an sm_100a cubin, loads it with the driver API and launches it.
//
// libnvrtc and libcuda are opened with dlopen, so that the C-ABI library itself has no link-time
// dependency on either (it must load, and export every symbol, on a machine without a driver).
// Compiled kernels are cached per process by the hash of their source text, i.e. by circuit
// STRUCTURE: gate coefficients are kernel parameters (a __grid_constant__ struct, read straight
// from the constant bank by the FP64 instructions).
#include "qb_engine.h"
#include "qb_jit.h"
#include "qb_jit_rt.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <atomic>
#include <mutex>
#include <thread>

namespace {

// ---- CUDA text around the generated stage functions ---------------------------------------------
const char* kPrelude = R"QJ(
#define QJ_C double2
#define QJ_DEV __device__ __forceinline__
#ifdef QJ_DEBUG_NOLOAD
#define QJ_LD(p) make_double2(1.0 + (double)(tid & 7u), 0.5)
#else
#define QJ_LD(p) __ldcs(p)
#endif
#ifdef QJ_DEBUG_NOSTORE
#define QJ_ST(p, v) do { if ((v).x == 1.2345e-300) __stcs(p, v); } while (0)
#else
#define QJ_ST(p, v) __stcs(p, v)
#endif
#define QJ_RESTRICT __restrict__
// conditional sign flip without a branch and off the FP64 pipe: XOR of the sign bits with s = 0 / 0x80000000
#define QJ_XSIGN(a, s) do { (a).x = __hiloint2double(__double2hiint((a).x) ^ (int)(s), __double2loint((a).x)); \
                            (a).y = __hiloint2double(__double2hiint((a).y) ^ (int)(s), __double2loint((a).y)); } while (0)
#define QJ_SYNC() __syncthreads()
// warp number as a value ptxas knows to be warp-uniform (shuffle broadcast), and one elected lane of a converged warp:
// what the bulk copies of the next tile are issued from (uniform-datapath addresses, no per-lane serialisation)
#define QJ_WARP_ID(tid) __shfl_sync(0xffffffffu, (tid) >> 5, 0)
__device__ __forceinline__ bool qj_elect_one() {
    unsigned p_;
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p_));
    return p_ != 0u;
}
#define QJ_ELECT(tid) qj_elect_one()
#define QJ_L2_PREFETCH(gsrc) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(512) : "memory")
// two mbarriers per CTA count the bytes of the next tile's bulk copies: [0] the early half (landing
// buffer behind the transposition buffer), [1] the late half (transposition buffer)
__device__ __forceinline__ unsigned qj_mbar(const int which) {
    __shared__ __align__(8) unsigned long long bar[2];
    return (unsigned)__cvta_generic_to_shared(&bar[which]);
}
#define QJ_MBAR_INIT()                                                                             \
    do {                                                                                           \
        if (threadIdx.x == 0) {                                                                    \
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(qj_mbar(0)) : "memory"); \
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(qj_mbar(1)) : "memory"); \
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                     \
        }                                                                                          \
        __syncthreads();                                                                           \
    } while (0)
#define QJ_BULK_COPY_TO(sdst, gsrc, which)                                                         \
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"( \
                     (unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(qj_mbar(which)) : "memory")
#define QJ_BULK_COPY_EARLY(sdst, gsrc) QJ_BULK_COPY_TO(sdst, gsrc, 0)
#define QJ_BULK_COPY(sdst, gsrc) QJ_BULK_COPY_TO(sdst, gsrc, 1)
#define QJ_MBAR_WAIT(which, parity)                                                                \
    do {                                                                                           \
        unsigned done_;                                                                            \
        do {                                                                                       \
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" \
                         : "=r"(done_) : "r"(qj_mbar(which)), "r"(parity) : "memory");             \
        } while (!done_);                                                                          \
    } while (0)
#define QJ_ASYNC_WAIT(parity)                                                                      \
    do {                                                                                           \
        QJ_MBAR_WAIT(0, parity);                                                                   \
        QJ_MBAR_WAIT(1, parity);                                                                   \
    } while (0)
#define QJ_EXPECT(which)                                                                           \
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(qj_mbar(which)), "r"(QJ_HALF_RUNS * 512) : "memory")
// The expect_tx of a half is posted by thread 0 BEFORE the barrier that releases the copies (its previous phase has
// been waited for by every thread in stage 0, and the new phase cannot complete before its bytes have arrived), so
// that every warp may issue its share of the copies right behind the barrier without racing the expectation.
// early half: issued right after the first barrier of a tile (everybody has read the landing buffer)
#define QJ_EXPECT_EARLY(tid, nbase)                                                                \
    do {                                                                                           \
        if (nbase != ~0ull && tid == 0) QJ_EXPECT(0);                                              \
    } while (0)
#define QJ_ISSUE_EARLY(tid, nbase, psi, buf)                                                       \
    do {                                                                                           \
        if (nbase != ~0ull) {                                                                      \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                           \
            qj_issue_early(tid, nbase, psi, buf + QJ_TILE_UNITS);                                  \
        }                                                                                          \
    } while (0)
// late half: every thread of the CTA has read its amplitudes out of the transposition buffer
// (barrier); order those generic reads before the TMA's writes (proxy fence).  A one-stage sweep
// has no earlier barrier, so its early half goes out here too.
#define QJ_ISSUE_NEXT(tid, nbase, psi, buf)                                                        \
    do {                                                                                           \
        if (nbase != ~0ull && tid == 0) {                                                          \
            QJ_EXPECT(1);                                                                          \
            if (!QJ_NSTAGES_GT1) QJ_EXPECT(0);                                                     \
        }                                                                                          \
        __syncthreads();                                                                           \
        if (nbase != ~0ull) {                                                                      \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                           \
            qj_issue_next(tid, nbase, psi, buf);                                                   \
            if (!QJ_NSTAGES_GT1) qj_issue_early(tid, nbase, psi, buf + QJ_TILE_UNITS);             \
        }                                                                                          \
    } while (0)
#ifdef QJ_POOL_GLOBAL
#define QJ_PRELUDE
#define QJ_POOL_PARAM const double* __restrict__ P
#define QJ_P(i) (__ldg(P + (i)))
#define QJ_POOL_KPARAM const double* __restrict__ P
#else
#define QJ_PRELUDE struct QjPool { double v[QJ_NP]; };
#define QJ_POOL_PARAM const QjPool& P
#define QJ_P(i) (P.v[i])
#define QJ_POOL_KPARAM const __grid_constant__ QjPool P
#endif
#define QJ_PREFETCH(psi, nbase, tid)                                                               \
    do {                                                                                           \
        if (nbase != ~0ull && tid < (1u << QJ_NH))                                                 \
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(psi + nbase + qj_run_offset(tid)), "r"(512) : "memory"); \
    } while (0)
)QJ";

// One CTA of QJ_T threads is persistent over tiles t = blockIdx.x + i * gridDim.x; a tile's 2^M
// amplitudes live in registers (2^R per thread) and pass through shared memory once between two
// stages.  QJ_CTAS CTAs are resident per SM so that one CTA's transfers overlap another's
// arithmetic; the next tile of a CTA arrives by TMA bulk copies while the current one is computed
// (early half: landing buffer, issued after the first barrier; late half: transposition buffer,
// issued in the last stage); only a CTA's first tile is loaded with LDG.
const char* kPostlude = R"QJ(
extern "C" __global__ void __launch_bounds__(QJ_T, QJ_CTAS)
qj_kernel(double2* __restrict__ psi, const unsigned long long tile0, const unsigned long long ntiles, const int prefetch, QJ_POOL_KPARAM) {
    extern __shared__ __align__(16) double2 buf[];
    const unsigned tid = threadIdx.x;
    // prefetch bit 0: L2 prefetch of the CTA's next tile at the start of a tile; bit 1: asynchronous
    // copy of the next tile into the transposition buffer during the last stage
    unsigned pre = 0, phase = 0;
    QJ_MBAR_INIT();
    // tiles [tile0, ntiles): the whole state, or one sub-block of it (a piece of a pipelined exchange)
    for (unsigned long long tile = tile0 + blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned long long tbase = qj_tile_base(tile);
        const unsigned long long nt = tile + gridDim.x;
        const unsigned long long next = nt < ntiles ? qj_tile_base(nt) : ~0ull;
        const unsigned long long pfbase = (prefetch & 1) ? next : ~0ull;
        const unsigned long long nbase = (prefetch & 2) ? next : ~0ull;
        QJ_RUN_STAGES(tid, tbase, nbase, pfbase, pre, psi, buf, P)
        if (nbase != ~0ull) { pre = 1u + phase; phase ^= 1u; } else pre = 0;
    }
}
)QJ";

// Variant of a sweep for a register that still IS a computational basis state |vindex> (qb_init_basis, or a product of
// basis kets, followed directly by gates): the tile is not loaded -- every amplitude is 0 except the one at vindex -- so
// the pass that would write the initial state and the read of it by the first sweep never happen (2 x 16 * 2^n bytes).
// Same stage functions; only the load / prefetch macros and the kernel skeleton differ, appended AFTER the common
// prelude so that the ordinary kernels' sources (and hashes) are untouched.
const char* kVirtualPatch = R"QJ(
#undef QJ_LD
#define QJ_LD(p) (((unsigned long long)((p) - psi)) == nbase ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0))
#undef QJ_ISSUE_EARLY
#define QJ_ISSUE_EARLY(tid, nbase, psi, buf) do { } while (0)
#undef QJ_EXPECT_EARLY
#define QJ_EXPECT_EARLY(tid, nbase) do { } while (0)
#undef QJ_ISSUE_NEXT
#define QJ_ISSUE_NEXT(tid, nbase, psi, buf) __syncthreads()
#undef QJ_PREFETCH
#define QJ_PREFETCH(psi, nbase, tid) do { } while (0)
)QJ";
const char* kVirtualPostlude = R"QJ(
extern "C" __global__ void __launch_bounds__(QJ_T, QJ_CTAS)
qj_kernel(double2* __restrict__ psi, const unsigned long long tile0, const unsigned long long ntiles, const int prefetch, QJ_POOL_KPARAM,
          const unsigned long long vindex) {
    extern __shared__ __align__(16) double2 buf[];
    const unsigned tid = threadIdx.x;
    (void)prefetch;
    for (unsigned long long tile = tile0 + blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned long long tbase = qj_tile_base(tile);
        // `nbase` carries the index of the one non-zero amplitude; pre = 0: every tile takes the "load" path of stage 0
        QJ_RUN_STAGES(tid, tbase, vindex, ~0ull, 0u, psi, buf, P)
    }
}
)QJ";

// ---- dynamically bound NVRTC + driver API ---------------------------------------------------------
struct Nvrtc {
    void* so = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    std::string why;
};

struct Driver {
    void* so = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    std::string why;
};

template <class F> bool bind(void* so, const char* name, F* out) {
    *out = (F)dlsym(so, name);
    return *out != nullptr;
}

Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char* env = getenv("QBOT_B200_NVRTC");
        const char* names[] = {env ? env : "libnvrtc.so.12", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) {
            r.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (r.so) break;
        }
        if (!r.so) { r.why = std::string("libnvrtc not found: ") + dlerror(); return r; }
        bool ok = bind(r.so, "nvrtcCreateProgram", &r.CreateProgram) && bind(r.so, "nvrtcCompileProgram", &r.CompileProgram) &&
                  bind(r.so, "nvrtcDestroyProgram", &r.DestroyProgram) && bind(r.so, "nvrtcGetCUBINSize", &r.GetCUBINSize) &&
                  bind(r.so, "nvrtcGetCUBIN", &r.GetCUBIN) && bind(r.so, "nvrtcGetProgramLogSize", &r.GetProgramLogSize) &&
                  bind(r.so, "nvrtcGetProgramLog", &r.GetProgramLog) && bind(r.so, "nvrtcGetErrorString", &r.GetErrorString);
        if (!ok) { r.why = "libnvrtc lacks a required symbol"; r.so = nullptr; }
        return r;
    }();
    return n;
}

Driver& driver() {
    static Driver d = [] {
        Driver r;
        r.so = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!r.so) { r.why = std::string("libcuda.so.1 not found: ") + dlerror(); return r; }
        bool ok = bind(r.so, "cuModuleLoadData", &r.ModuleLoadData) && bind(r.so, "cuModuleUnload", &r.ModuleUnload) &&
                  bind(r.so, "cuModuleGetFunction", &r.ModuleGetFunction) && bind(r.so, "cuFuncSetAttribute", &r.FuncSetAttribute) &&
                  bind(r.so, "cuFuncGetAttribute", &r.FuncGetAttribute) && bind(r.so, "cuLaunchKernel", &r.LaunchKernel) &&
                  bind(r.so, "cuGetErrorString", &r.GetErrorString);
        if (!ok) { r.why = "libcuda lacks a required symbol"; r.so = nullptr; }
        return r;
    }();
    return d;
}

std::string cu_err(CUresult e) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(e, &s);
    return s ? s : "unknown driver error";
}

// second, independent fingerprint of a generated source: the cache is keyed by a 64-bit FNV-1a hash,
// and a hit whose length or second hash differs must never launch the other structure's kernel
static uint64_t src_hash2(const std::string& src) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)src.size();
    for (unsigned char ch : src) { h ^= ch; h *= 0xff51afd7ed558ccdull; h ^= h >> 29; }
    return h;
}

struct Compiled {
    uint64_t src_len = 0, hash2 = 0;
    std::vector<char> cubin;
    QjSourceInfo info;
    bool pool_global = false;
    double compile_ms = 0;
    std::map<int, std::pair<CUmodule, CUfunction>> per_device;   // loaded module per device
};

// never destroyed: background warm-up threads (qb_tile.cu) may still be compiling into the cache while
// the process runs its static destructors
std::mutex& g_mu = *new std::mutex;
std::map<uint64_t, Compiled>& g_cache = *new std::map<uint64_t, Compiled>;      // by hash of the full source
std::map<uint64_t, int>& g_seen = *new std::map<uint64_t, int>;                // sightings of structures that are not compiled (yet)
QbJitStats& g_stats = *new QbJitStats;

// resident CTAs per SM the kernel is compiled for (register budget = 65536 / (threads * CTAs));
// 0 = the generator's default (2 for 256-thread tiles, 4 for 128-thread tiles)
int jit_ctas_override(int M) {
    const char* e = getenv(M == 12 ? "QBOT_B200_JIT_CTAS12" : "QBOT_B200_JIT_CTAS11");
    const int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 8) ? v : 0;
}

constexpr int kMaxParamPoolDoubles = 480;     // 3840 bytes of coefficients + 24 bytes of scalars < 4 KB of parameters

}  // namespace

std::string qb_jit_full_source(const uint8_t* program, QjSourceInfo* info, bool* pool_global, bool virtual_basis) {
    QjSourceInfo li;
    std::string body = qj_generate(program, &li);
    const bool pg = li.npool > kMaxParamPoolDoubles;
    if (info) *info = li;
    if (pool_global) *pool_global = pg;
    std::string src;
    if (pg) src += "#define QJ_POOL_GLOBAL 1\n";
    // diagnostics only (wrong results!): drop the HBM loads and / or stores of the sweep to time its parts
    if (getenv("QBOT_B200_DEBUG_NOLOAD")) src += "#define QJ_DEBUG_NOLOAD 1\n";
    if (getenv("QBOT_B200_DEBUG_NOSTORE")) src += "#define QJ_DEBUG_NOSTORE 1\n";
    if (getenv("QBOT_B200_DEBUG_NOWAIT")) src += "#define QJ_DEBUG_NOWAIT 1\n";
    if (const int c = jit_ctas_override(li.M)) src += "#define QJ_CTAS " + std::to_string(c) + "\n";
    src += kPrelude;
    if (virtual_basis) src += kVirtualPatch;
    src += body;
    src += virtual_basis ? kVirtualPostlude : kPostlude;
    return src;
}

// compile `src` to an sm_100a cubin; throws qb_error with the NVRTC log on failure
std::vector<char> qb_jit_compile(const std::string& src, std::string* log_out) {
    Nvrtc& n = nvrtc();
    if (!n.so) throw qb_error(-2, "sweep specialiser: " + n.why);
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.CreateProgram(&prog, src.c_str(), "qbot_b200_sweep.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) throw qb_error(-2, std::string("nvrtcCreateProgram: ") + n.GetErrorString(r));
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    r = n.CompileProgram(prog, 4, opts);
    std::string log;
    size_t lsz = 0;
    if (n.GetProgramLogSize(prog, &lsz) == NVRTC_SUCCESS && lsz > 1) {
        log.resize(lsz);
        n.GetProgramLog(prog, &log[0]);
    }
    if (log_out) *log_out = log;
    if (r != NVRTC_SUCCESS) {
        n.DestroyProgram(&prog);
        throw qb_error(-2, std::string("nvrtcCompileProgram: ") + n.GetErrorString(r) + "\n" + log.substr(0, 2000));
    }
    size_t csz = 0;
    r = n.GetCUBINSize(prog, &csz);
    std::vector<char> cubin(csz);
    if (r == NVRTC_SUCCESS) r = n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    if (r != NVRTC_SUCCESS || csz == 0) throw qb_error(-2, std::string("nvrtcGetCUBIN: ") + n.GetErrorString(r));
    return cubin;
}

bool qb_jit_available(std::string* why) {
    if (!nvrtc().so) { if (why) *why = nvrtc().why; return false; }
    if (!driver().so) { if (why) *why = driver().why; return false; }
    return true;
}

int qb_jit_note(uint64_t key) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.count(key)) return -1;
    return ++g_seen[key];
}

QbJitStats qb_jit_stats() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_stats;
}

// Compile the kernels of several sweep programs that are not cached yet, NVRTC running on up to 8
// host threads (one program per thread at a time): a 13-sweep circuit costs one compile latency
// (~0.7 s) instead of thirteen.  Failures are left to qb_jit_get, which reports them.
void qb_jit_precompile(const std::vector<const uint8_t*>& programs) {
    struct Job { std::string src; uint64_t key; QjSourceInfo info; bool pg; std::vector<char> cubin; double ms = 0; bool ok = false; };
    std::vector<Job> jobs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (const uint8_t* p : programs) {
            Job j;
            j.src = qb_jit_full_source(p, &j.info, &j.pg);
            j.key = qj_hash(j.src);
            bool dup = g_cache.count(j.key) != 0;
            for (const Job& o : jobs) dup = dup || o.key == j.key;
            if (!dup) jobs.push_back(std::move(j));
        }
    }
    if (jobs.size() < 2 || !nvrtc().so) return;
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = std::min<size_t>(jobs.size(), std::max(1u, std::min(8u, hw ? hw : 1u)));
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; t++) {
        pool.emplace_back([&] {
            for (size_t i = next++; i < jobs.size(); i = next++) {
                const auto t0 = std::chrono::steady_clock::now();
                try {
                    jobs[i].cubin = qb_jit_compile(jobs[i].src, nullptr);
                    jobs[i].ok = true;
                } catch (...) {
                }
                jobs[i].ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            }
        });
    }
    for (auto& th : pool) th.join();
    std::lock_guard<std::mutex> lk(g_mu);
    for (Job& j : jobs) {
        if (!j.ok || g_cache.count(j.key)) continue;
        Compiled c;
        c.cubin = std::move(j.cubin);
        c.src_len = j.src.size(); c.hash2 = src_hash2(j.src);
        c.info = j.info;
        c.pool_global = j.pg;
        c.compile_ms = j.ms;
        g_stats.kernels_compiled++;
        g_stats.compile_ms += j.ms;
        g_cache.emplace(j.key, std::move(c));
    }
}

// compile one program into the cache without loading it on any device (host work only)
void qb_jit_compile_cached(const uint8_t* program, bool virtual_basis) {
    QjSourceInfo info;
    bool pg = false;
    const std::string src = qb_jit_full_source(program, &info, &pg, virtual_basis);
    const uint64_t key = qj_hash(src);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_cache.count(key)) return;
    }
    const auto t0 = std::chrono::steady_clock::now();
    Compiled c;
    c.cubin = qb_jit_compile(src, nullptr);
    c.src_len = src.size(); c.hash2 = src_hash2(src);
    c.info = info;
    c.pool_global = pg;
    c.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.count(key)) return;
    g_stats.kernels_compiled++;
    g_stats.compile_ms += c.compile_ms;
    g_cache.emplace(key, std::move(c));
}

// the compiled kernel of `program` on `device` (compiling / loading it on first use)
QbJitKernel qb_jit_get(const uint8_t* program, int device, bool virtual_basis) {
    QjSourceInfo info;
    bool pg = false;
    const std::string src = qb_jit_full_source(program, &info, &pg, virtual_basis);
    const uint64_t key = qj_hash(src);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
        const auto t0 = std::chrono::steady_clock::now();
        Compiled c;
        c.cubin = qb_jit_compile(src, nullptr);
        c.src_len = src.size(); c.hash2 = src_hash2(src);
        c.info = info;
        c.pool_global = pg;
        c.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        g_stats.kernels_compiled++;
        g_stats.compile_ms += c.compile_ms;
        it = g_cache.emplace(key, std::move(c)).first;
    } else {
        if (it->second.src_len != src.size() || it->second.hash2 != src_hash2(src))
            throw qb_error(-2, "sweep specialiser: source-hash collision in the kernel cache");
        g_stats.cache_hits++;
    }
    Compiled& c = it->second;
    auto dv = c.per_device.find(device);
    if (dv == c.per_device.end()) {
        Driver& d = driver();
        if (!d.so) throw qb_error(-2, "sweep specialiser: " + d.why);
        CUmodule mod = nullptr;
        CUfunction fn = nullptr;
        CUresult e = d.ModuleLoadData(&mod, c.cubin.data());
        if (e != CUDA_SUCCESS) throw qb_error(-2, "cuModuleLoadData: " + cu_err(e));
        e = d.ModuleGetFunction(&fn, mod, "qj_kernel");
        if (e != CUDA_SUCCESS) throw qb_error(-2, "cuModuleGetFunction: " + cu_err(e));
        const int smem = c.info.tile_units * 16 + (8 << c.info.M);      // transposition buffer + landing buffer of half a tile
        e = d.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
        if (e != CUDA_SUCCESS) throw qb_error(-2, "cuFuncSetAttribute(smem): " + cu_err(e));
        d.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
        dv = c.per_device.emplace(device, std::make_pair(mod, fn)).first;
    }
    QbJitKernel k;
    k.fn = (void*)dv->second.second;
    k.threads = c.info.threads;
    k.smem_bytes = c.info.tile_units * 16 + (8 << c.info.M);
    k.npool = c.info.npool;
    k.pool_global = c.pool_global;
    k.ctas_per_sm = jit_ctas_override(c.info.M) ? jit_ctas_override(c.info.M) : qj_default_ctas(c.info.M, c.info.R);
    k.M = c.info.M;
    return k;
}

// launch: `pool` = qj_pool(program) on the host (parameter variant) or its device copy
void qb_jit_launch(const QbJitKernel& k, cudaStream_t stream, int sms, cplx* psi, uint64_t tile_begin, uint64_t tile_end, int prefetch,
                   const double* pool_host, const double* pool_dev, uint64_t virtual_index) {
    Driver& d = driver();
    unsigned long long t0 = tile_begin, nt = tile_end, vi = virtual_index;
    int pf = prefetch;
    void* psi_arg = (void*)psi;
    const void* pool_ptr = pool_dev;
    void* args[6] = {&psi_arg, &t0, &nt, &pf, k.pool_global ? (void*)&pool_ptr : (void*)pool_host, &vi};     // (vi: the virtual-basis variant only)
    const unsigned grid = (unsigned)std::min<uint64_t>(tile_end - tile_begin, (uint64_t)sms * k.ctas_per_sm);
    CUresult e = d.LaunchKernel((CUfunction)k.fn, grid, 1, 1, (unsigned)k.threads, 1, 1, (unsigned)k.smem_bytes, (CUstream)stream, args, nullptr);
    if (e != CUDA_SUCCESS) throw qb_error(-2, "cuLaunchKernel(sweep): " + cu_err(e));
}
