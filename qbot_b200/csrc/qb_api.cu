// qbot_b200 -- C ABI (include/qbot_b200.h): state handles, gate queue, dispatch to kernels.
#include "../../include/qbot_b200.h"
#include "qb_common.cuh"
#include "qb_engine.h"
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include "qb_jit_rt.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

static thread_local std::string g_err;

#define QB_API_BEGIN try {
#define QB_API_END                                                                             \
    return QB_OK;                                                                              \
    } catch (const qb_error& e) { g_err = e.what(); return e.code; }                           \
    catch (const std::bad_alloc&) { g_err = "host allocation failed"; return QB_ERR_ALLOC; }   \
    catch (const std::exception& e) { g_err = e.what(); return QB_ERR_ARG; }                   \
    catch (...) { g_err = "unknown error"; return QB_ERR_ARG; }

#define QB_REQUIRE_API(cond)                                                                   \
    do {                                                                                       \
        if (!(cond)) { g_err = "bad argument: " #cond; return QB_ERR_ARG; }                    \
    } while (0)

static inline cplx C(double re, double im) { return make_double2(re, im); }
#include "qb_plan.h"

// ---------------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------------
struct DevGuard {
    int prev;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) QB_CUDA(cudaSetDevice(dev)); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int sm_count_of(int dev) {
    static std::mutex mu;
    static std::map<int, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int v = 0;
    QB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    cache[dev] = v;
    return v;
}
static int device_count_cached() {
    static int n = -1;
    if (n < 0) QB_CUDA(cudaGetDeviceCount(&n));
    return n;
}

// Large register buffers are recycled instead of returned to the driver: cudaMalloc / cudaFree of
// a 16 GiB ket cost tens of milliseconds each, and every executeTxt call creates and drops one
// register.  A few buffers per device are kept, together at most a third of the device memory;
// an allocation failure empties the cache and retries.
namespace {
struct BufCache {
    std::mutex mu;
    std::vector<std::tuple<int, size_t, void*>> free_list;      // (device, bytes, pointer): registers
    std::vector<std::tuple<int, size_t, void*>> small_list;     // work buffers in power-of-two size classes
};
BufCache g_bufs;
constexpr size_t kCacheMinBytes = 64u << 20;

// ONE compute stream per device and process.  Every handle works on it unless its creator supplies
// a stream of its own (qb_create_external / qb_set_stream).  Consequences: a handle made from another
// one (partial trace, scatter product, mixture, clone, branch view ...) is ordered after its source
// without any host-side synchronisation, and recycled buffers are handed from one user to the next in
// stream order.  (Creating a stream per handle and synchronising around every hand-over cost ~100 us
// per density-matrix op, several times the kernels themselves at 256 MiB.)
cudaStream_t default_stream(int device) {
    static std::mutex mu;
    static std::map<int, cudaStream_t>& streams = *new std::map<int, cudaStream_t>;
    std::lock_guard<std::mutex> lk(mu);
    auto it = streams.find(device);
    if (it != streams.end()) return it->second;
    cudaStream_t st = nullptr;
    QB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    streams[device] = st;
    return st;
}
// pooled buffers are only ever "free in the order of the default stream": a user on another stream
// synchronises on the way in (with the default stream) and on the way out (with its own)
void pool_enter(int device, cudaStream_t user) {
    if (user != default_stream(device)) cudaStreamSynchronize(default_stream(device));
}
void pool_leave(int device, cudaStream_t user) {
    if (user && user != default_stream(device)) cudaStreamSynchronize(user);
}

// short-lived work buffers (partials of a probability reduction, ...): cudaMalloc / cudaFree cost
// from a few to hundreds of milliseconds next to multi-GiB allocations, so they are recycled too
size_t work_class(size_t bytes) {
    size_t c = 64u << 10;
    while (c < bytes) c <<= 1;
    return c;
}
void* work_alloc(int device, size_t bytes, cudaStream_t user) {
    const size_t c = work_class(bytes);
    pool_enter(device, user);
    {
        std::lock_guard<std::mutex> lk(g_bufs.mu);
        for (size_t i = 0; i < g_bufs.small_list.size(); i++) {
            if (std::get<0>(g_bufs.small_list[i]) == device && std::get<1>(g_bufs.small_list[i]) == c) {
                void* p = std::get<2>(g_bufs.small_list[i]);
                g_bufs.small_list.erase(g_bufs.small_list.begin() + i);
                return p;
            }
        }
    }
    void* p = nullptr;
    QB_CUDA(cudaMalloc(&p, c));
    return p;
}
void work_free(int device, size_t bytes, void* p, cudaStream_t user) {
    const size_t c = work_class(bytes);
    pool_leave(device, user);
    if (c <= (256u << 20)) {
        void* old = nullptr;
        int old_dev = device;
        {
            std::lock_guard<std::mutex> lk(g_bufs.mu);
            g_bufs.small_list.emplace_back(device, c, p);
            if (g_bufs.small_list.size() > 96) {
                old_dev = std::get<0>(g_bufs.small_list.front());
                old = std::get<2>(g_bufs.small_list.front());
                g_bufs.small_list.erase(g_bufs.small_list.begin());
            }
        }
        if (old) { cudaStreamSynchronize(default_stream(old_dev)); cudaFree(old); }      // (the oldest one makes room)
        return;
    }
    cudaStreamSynchronize(default_stream(device));
    cudaFree(p);
}

void* cached_alloc(int device, size_t bytes, cudaStream_t user) {
    if (bytes < kCacheMinBytes) return work_alloc(device, bytes, user);        // small registers: power-of-two size classes
    pool_enter(device, user);
    {
        std::lock_guard<std::mutex> lk(g_bufs.mu);
        for (size_t i = 0; i < g_bufs.free_list.size(); i++) {
            if (std::get<0>(g_bufs.free_list[i]) == device && std::get<1>(g_bufs.free_list[i]) == bytes) {
                void* p = std::get<2>(g_bufs.free_list[i]);
                g_bufs.free_list.erase(g_bufs.free_list.begin() + i);
                return p;
            }
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(g_bufs.mu);
        for (auto& b : g_bufs.free_list) if (std::get<0>(b) == device) cudaFree(std::get<2>(b));
        g_bufs.free_list.erase(std::remove_if(g_bufs.free_list.begin(), g_bufs.free_list.end(),
                                              [&](const std::tuple<int, size_t, void*>& b) { return std::get<0>(b) == device; }),
                               g_bufs.free_list.end());
        e = cudaMalloc(&p, bytes);
    }
    QB_CUDA(e);
    return p;
}

size_t device_total_mem(int device) {           // asked once per device (cudaMemGetInfo is not cheap)
    static std::mutex mu;
    static std::map<int, size_t> total;
    std::lock_guard<std::mutex> lk(mu);
    auto it = total.find(device);
    if (it != total.end()) return it->second;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
    total[device] = total_b;
    return total_b;
}

void cached_free(int device, size_t bytes, void* p, cudaStream_t stream) {
    if (bytes < kCacheMinBytes) {
        work_free(device, bytes, p, stream);
        return;
    }
    if (bytes >= kCacheMinBytes) {
        const size_t total_b = device_total_mem(device);
        if (bytes <= total_b / 3) {
            pool_leave(device, stream);            // nothing on a foreign stream may still be writing into it
            // A few registers are kept (density-matrix ops create a new one per partial trace / scatter / mix), never more
            // than a third of the device memory in total.  The buffer just freed is the likeliest to be asked for again,
            // so it always goes in and the OLDEST ones make room: a program that moves on to another register size (the
            // bench's configs after its 16 GiB headline) is not left allocating and freeing through the driver.
            std::vector<void*> evict;
            {
                std::lock_guard<std::mutex> lk(g_bufs.mu);
                g_bufs.free_list.emplace_back(device, bytes, p);
                for (;;) {
                    int mine = 0;
                    size_t held = 0;
                    for (auto& b : g_bufs.free_list) if (std::get<0>(b) == device) { mine++; held += std::get<1>(b); }
                    if (mine <= 12 && held <= total_b / 3) break;
                    auto it = std::find_if(g_bufs.free_list.begin(), g_bufs.free_list.end(),
                                           [&](const std::tuple<int, size_t, void*>& b) { return std::get<0>(b) == device; });
                    evict.push_back(std::get<2>(*it));
                    g_bufs.free_list.erase(it);
                }
            }
            if (!evict.empty()) {
                if (stream) cudaStreamSynchronize(stream);
                cudaStreamSynchronize(default_stream(device));
                for (void* q : evict) cudaFree(q);
            }
            return;
        }
    }
    if (stream) cudaStreamSynchronize(stream);
    cudaStreamSynchronize(default_stream(device));
    cudaFree(p);
}
}  // namespace

qb_state::~qb_state() {
    qb_engine_free(this);
    if (d && owns) { DevGuard g(device); cached_free(device, bytes(), d, stream); }
    if (scratch) { DevGuard g(device); cached_free(device, bytes(), scratch, stream); }   // same allocator as `d`: the two may have been swapped
    if (stage) { DevGuard g(device); work_free(device, STAGE_BYTES, stage, stream); }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream && owns_stream) cudaStreamDestroy(stream);
}

static qb_state* new_state(int kind, int nq, int64_t nbranch, int device, void* ext, void* ext_stream) {
    QB_REQUIRE(kind == QB_KET || kind == QB_DM, "kind must be QB_KET or QB_DM");
    QB_REQUIRE(nq >= 0 && nq <= 40, "nqubits out of range");
    QB_REQUIRE(nbranch >= 1, "nbranch must be >= 1");
    int nbits = kind == QB_KET ? nq : 2 * nq;
    QB_REQUIRE(nbits <= 40, "state too large (index bits > 40)");
    const int ndev = device_count_cached();
    QB_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    std::unique_ptr<qb_state> s(new qb_state());
    s->kind = kind; s->nq = nq; s->nbits = nbits; s->nbranch = nbranch; s->device = device;
    DevGuard g(device);
    s->sms = sm_count_of(device);
    s->stream = ext_stream ? (cudaStream_t)ext_stream : default_stream(device);
    s->owns_stream = false;
    if (ext) { s->d = (cplx*)ext; s->owns = false; }
    else { s->d = (cplx*)cached_alloc(device, s->bytes(), s->stream); s->owns = true; }
    return s.release();
}

// work queued on `earlier`'s stream must be visible to what `later_stream` does next
static void order_after(cudaStream_t later_stream, const qb_state* earlier) {
    if (earlier->stream != later_stream) QB_CUDA(cudaStreamSynchronize(earlier->stream));
}

LaunchCtx qb_state::ctx() { return LaunchCtx{stream, sms, &stats.kernel_launches}; }

cplx* qb_state::get_scratch() {
    if (!scratch) scratch = (cplx*)cached_alloc(device, bytes(), stream);       // interchangeable with `d` (out-of-place gates swap them)
    return scratch;
}

// ---------------------------------------------------------------------------------------------
// gate classification + queue
// ---------------------------------------------------------------------------------------------
static void fill_ins(uint64_t mask, int nbits_total, int* ins, int& nins) {
    nins = 0;
    for (int p = 0; p < nbits_total; p++) if ((mask >> p) & 1ull) ins[nins++] = p;
}

cplx* qb_state::upload_small(const void* host, size_t bytes) {
    // staging ring for gate matrices: consecutive launches must not overwrite each other
    if (!stage) { stage = work_alloc(device, STAGE_BYTES, stream); stage_off = 0; }
    size_t need = (bytes + 255) & ~size_t(255);
    QB_REQUIRE(need <= STAGE_BYTES, "matrix too large for the staging buffer");
    if (stage_off + need > STAGE_BYTES) { QB_CUDA(cudaStreamSynchronize(stream)); stage_off = 0; }
    char* dst = (char*)stage + stage_off;
    stage_off += need;
    QB_CUDA(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, stream));
    // pageable source: the copy is staged by the runtime before returning, so `host` may die
    return (cplx*)dst;
}

// one ket-level gate, one sweep (no fusion)
void qb_state::run_gate_unfused(const QGate& g) {
    const int K = g.k;
    const uint64_t total_bits_mask_ok = (nbits >= 64) ? ~0ull : ((1ull << nbits) - 1ull);
    QB_REQUIRE((g.tmask() & ~total_bits_mask_ok) == 0 && (g.cmask & ~total_bits_mask_ok) == 0, "gate bit out of range");
    QB_REQUIRE((g.tmask() & g.cmask) == 0, "control overlaps target");
    LaunchCtx c = ctx();
    const int ncontrols = __builtin_popcountll(g.cmask);
    const uint64_t touched = ((uint64_t)nbranch << nbits) >> ncontrols;
    if (g.type == QB_G_DIAG && K <= QB_DIAG_MAXK) {
        DiagArgs a;
        memset(&a, 0, sizeof(a));
        a.psi = d; a.cmask = g.cmask; a.k = K;
        for (int i = 0; i < K; i++) a.tb[i] = g.tb[i];
        fill_ins(g.cmask, nbits, a.ins, a.nins);
        a.nwork = touched;
        if (K <= 2) for (int i = 0; i < (1 << K); i++) a.inl[i] = g.m[i];
        else a.diag = upload_small(g.m.data(), sizeof(cplx) << K);
        qb_launch_diag(c, a);
        stats.bytes_moved += touched * 32;
    } else if (K <= QB_REG_MAXK) {
        std::vector<cplx> m = qb_dense_of(g);
        DenseArgs a;
        memset(&a, 0, sizeof(a));
        a.psi = d; a.cmask = g.cmask;
        for (int i = 0; i < K; i++) a.tb[i] = g.tb[i];
        fill_ins(g.cmask | g.tmask(), nbits, a.ins, a.nins);
        a.nwork = touched >> K;
        if (K <= 2) for (int i = 0; i < (1 << (2 * K)); i++) a.inl[i] = m[i];
        else a.mat = upload_small(m.data(), sizeof(cplx) << (2 * K));
        qb_launch_dense(c, K, a);
        stats.bytes_moved += touched * 32;
    } else if (K >= QB_MMA_MINK && K <= QB_MMA_MAXK && !getenv("QBOT_B200_NO_DMMA") &&
               nbits - K - ncontrols >= QB_MMA_TILE_BITS - K) {
        // dense 6..8-qubit block (qftGate(k), user unitaries): complex GEMM on the FP64 tensor cores, in place
        std::vector<cplx> m = qb_dense_of(g);
        const int D = 1 << K, ncb = QB_MMA_TILE_BITS - K, ldx = (1 << ncb) + 4;
        std::vector<double> planes((size_t)2 * D * D);
        for (size_t i = 0; i < (size_t)D * D; i++) { planes[i] = m[i].x; planes[(size_t)D * D + i] = m[i].y; }
        const size_t pbytes = planes.size() * sizeof(double);
        double* dp = (double*)work_alloc(device, pbytes, stream);
        QB_CUDA(cudaMemcpyAsync(dp, planes.data(), pbytes, cudaMemcpyHostToDevice, stream));     // pageable: staged before returning
        DenseMmaArgs a;
        memset(&a, 0, sizeof(a));
        a.psi = d; a.ur = dp; a.ui = dp + (size_t)D * D; a.cmask = g.cmask;
        uint64_t colmask = 0;
        for (int p = 0, left = ncb; p < nbits && left; p++)
            if (!(((g.tmask() | g.cmask) >> p) & 1ull)) { colmask |= 1ull << p; left--; }
        int nb = 0, ncol = 0;
        for (int p = 0; p < nbits; p++) {
            if ((colmask >> p) & 1ull) { a.apos[nb] = p; a.sw[nb] = 1 << ncol++; nb++; }
            else if ((g.tmask() >> p) & 1ull) {
                int b = 0;
                while (g.tb[b] != p) b++;
                a.apos[nb] = p; a.sw[nb] = (1 << (K - 1 - b)) * ldx; nb++;
            }
        }
        fill_ins(g.cmask | g.tmask() | colmask, nbits, a.ins, a.nins);
        a.ntiles = touched >> QB_MMA_TILE_BITS;
        qb_launch_dense_mma(c, K, a);
        work_free(device, pbytes, dp, stream);
        stats.bytes_moved += touched * 32;
    } else {
        QB_REQUIRE(K <= QB_BIG_MAXK, "gate acts on too many qubits");
        std::vector<cplx> m = qb_dense_of(g);
        const int D = 1 << K;
        std::vector<uint64_t> offs(D);
        for (int j = 0; j < D; j++) {
            uint64_t o = 0;
            for (int b = 0; b < K; b++) if ((j >> (K - 1 - b)) & 1) o |= 1ull << g.tb[b];
            offs[j] = o;
        }
        // the matrix can exceed the staging ring: give it its own allocation
        cplx* dm = nullptr; uint64_t* doffs = nullptr;
        QB_CUDA(cudaMalloc((void**)&dm, sizeof(cplx) * (size_t)D * D));
        QB_CUDA(cudaMalloc((void**)&doffs, sizeof(uint64_t) * D));
        QB_CUDA(cudaMemcpyAsync(dm, m.data(), sizeof(cplx) * (size_t)D * D, cudaMemcpyHostToDevice, stream));
        QB_CUDA(cudaMemcpyAsync(doffs, offs.data(), sizeof(uint64_t) * D, cudaMemcpyHostToDevice, stream));
        BigArgs a;
        memset(&a, 0, sizeof(a));
        a.in = d; a.out = get_scratch(); a.mat = dm; a.offs = doffs;
        a.total = (uint64_t)nbranch << nbits; a.cmask = g.cmask; a.tmask = g.tmask(); a.k = K;
        for (int i = 0; i < K; i++) a.tb[i] = g.tb[i];
        qb_launch_big(c, a);
        if (owns) std::swap(d, scratch);
        else QB_CUDA(cudaMemcpyAsync(d, scratch, bytes(), cudaMemcpyDeviceToDevice, stream));
        QB_CUDA(cudaStreamSynchronize(stream));
        cudaFree(dm); cudaFree(doffs);
        stats.bytes_moved += a.total * 32;
    }
    stats.gates_applied++;
    stats.state_passes++;
}

void qb_state::enqueue(QGate&& g) {
    if (qb_is_identity(g)) return;
    queue.push_back(std::move(g));
    if (!fusion || queue.size() >= 4096) flush();
}

void qb_state::materialize() {
    if (!virt) return;
    virt = false;
    DevGuard gd(device);
    qb_launch_fill_basis(ctx(), d, per_branch(), nbranch, virt_index);
    stats.bytes_moved += bytes();
}

// true when an initial basis state may stay virtual (see qb_state::virt): large single-branch kets only -- small
// registers and density matrices are written at once (their first steps are not specialised sweeps anyway)
static bool lazy_basis_ok(const qb_state* s) {
    static const bool on = [] { const char* e = getenv("QBOT_B200_LAZY_INIT"); return !e || atoi(e) != 0; }();
    return on && s->kind == QB_KET && s->nbranch == 1 && s->nbits >= 22 && s->fusion && qb_engine_available();
}

void qb_state::flush() {
    if (queue.empty()) { materialize(); return; }
    if (pending_plan) throw qb_error(-4, "gates were queued while a planned queue is being run step by step (qb_finish_queue first)");
    DevGuard gd(device);
    std::vector<QGate> q;
    q.swap(queue);
    if (fusion && qb_engine_available()) {
        qb_engine_run(this, q);
        materialize();               // (a plan without steps)
    } else {
        materialize();
        for (const QGate& g : q) run_gate_unfused(g);
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* qb_version(void) { return "qbot_b200 0.1.0 (sm_100a)"; }
const char* qb_last_error(void) { return g_err.c_str(); }

int qb_device_count(int* count) {
    QB_API_BEGIN
    QB_REQUIRE(count, "count is NULL");
    *count = 0;
    QB_CUDA(cudaGetDeviceCount(count));
    QB_API_END
}

int qb_device_info(int device, char* name, int name_len, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
    QB_API_BEGIN
    cudaDeviceProp p;
    QB_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    QB_API_END
}

int qb_create(qb_state** out, int kind, int nqubits, int64_t nbranch, int device) {
    QB_API_BEGIN
    QB_REQUIRE(out, "out is NULL");
    *out = new_state(kind, nqubits, nbranch, device, nullptr, nullptr);
    QB_API_END
}

int qb_create_external(qb_state** out, int kind, int nqubits, int64_t nbranch, int device, void* amplitudes_dev, void* cuda_stream) {
    QB_API_BEGIN
    QB_REQUIRE(out && amplitudes_dev, "NULL argument");
    *out = new_state(kind, nqubits, nbranch, device, amplitudes_dev, cuda_stream);
    QB_API_END
}

int qb_destroy(qb_state* s) {
    QB_API_BEGIN
    if (s) {
        DevGuard g(s->device);
        delete s;                                   // the destructor hands stage / scratch / the register buffer back in stream order
    }
    QB_API_END
}

int qb_clone(const qb_state* cs, qb_state** out) {
    QB_API_BEGIN
    qb_state* s = const_cast<qb_state*>(cs);
    QB_REQUIRE(s && out, "NULL argument");
    s->flush();
    DevGuard g(s->device);
    qb_state* n = new_state(s->kind, s->nq, s->nbranch, s->device, nullptr, s->owns_stream ? nullptr : (void*)s->stream);
    n->fusion = s->fusion;
    QB_CUDA(cudaMemcpyAsync(n->d, s->d, s->bytes(), cudaMemcpyDeviceToDevice, s->stream));      // n shares s's stream
    *out = n;
    QB_API_END
}

int qb_info(const qb_state* s, int* kind, int* nqubits, int64_t* nbranch, int* device) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    if (kind) *kind = s->kind;
    if (nqubits) *nqubits = s->nq;
    if (nbranch) *nbranch = s->nbranch;
    if (device) *device = s->device;
    QB_API_END
}

int qb_device_ptr(qb_state* s, void** p) {
    QB_API_BEGIN
    QB_REQUIRE(s && p, "NULL argument");
    s->flush();
    *p = s->d;
    QB_API_END
}

int qb_set_stream(qb_state* s, void* cuda_stream) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    DevGuard g(s->device);
    QB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->owns_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = cuda_stream ? (cudaStream_t)cuda_stream : default_stream(s->device);
    s->owns_stream = false;
    QB_CUDA(cudaStreamSynchronize(default_stream(s->device)));     // the buffer changes stream: nothing of the old order may be pending
    QB_API_END
}

int qb_init_basis(qb_state* s, uint64_t index) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->queue.clear();
    DevGuard g(s->device);
    uint64_t per = s->per_branch();
    uint64_t flat = index;
    if (s->kind == QB_DM) {
        QB_REQUIRE(index < (1ull << s->nq), "basis index out of range");
        flat = (index << s->nq) | index;
    } else QB_REQUIRE(index < per, "basis index out of range");
    if (lazy_basis_ok(s)) { s->virt = true; s->virt_index = flat; return QB_OK; }      // written by whoever needs it first
    s->virt = false;
    qb_launch_fill_basis(s->ctx(), s->d, per, s->nbranch, flat);
    s->stats.bytes_moved += s->bytes();
    QB_API_END
}

int qb_init_product(qb_state* s, const double* vecs, int per_branch) {
    QB_API_BEGIN
    QB_REQUIRE(s && vecs, "NULL argument");
    s->queue.clear();
    s->virt = false;
    if (!per_branch && lazy_basis_ok(s)) {
        // a product of computational basis kets (tensorExp(comp.kets[0], n), ...) is a basis state
        uint64_t index = 0;
        bool basis = true;
        for (int q = 0; q < s->nq && basis; q++) {
            const double* f = vecs + 4 * q;          // (re0, im0, re1, im1)
            if (f[0] == 1.0 && f[1] == 0.0 && f[2] == 0.0 && f[3] == 0.0) { }
            else if (f[0] == 0.0 && f[1] == 0.0 && f[2] == 1.0 && f[3] == 0.0) index |= 1ull << (s->nq - 1 - q);
            else basis = false;
        }
        if (basis) { s->virt = true; s->virt_index = index; return QB_OK; }
    }
    DevGuard g(s->device);
    size_t per = (s->kind == QB_KET ? 2 : 4) * (size_t)s->nq;
    size_t count = per * (per_branch ? (size_t)s->nbranch : 1);
    const size_t dv_bytes = sizeof(cplx) * std::max<size_t>(count, 1);
    cplx* dv = (cplx*)work_alloc(s->device, dv_bytes, s->stream);
    QB_CUDA(cudaMemcpyAsync(dv, vecs, sizeof(cplx) * count, cudaMemcpyHostToDevice, s->stream));   // pageable source: staged before the call returns
    qb_launch_init_product(s->ctx(), s->d, s->kind, s->nq, s->nbranch, dv, per_branch);
    work_free(s->device, dv_bytes, dv, s->stream);
    s->stats.bytes_moved += s->bytes();
    QB_API_END
}

int qb_init_diag(qb_state* s, const double* values) {
    QB_API_BEGIN
    QB_REQUIRE(s && values, "NULL argument");
    QB_REQUIRE(s->kind == QB_DM && s->nbranch == 1, "init_diag: needs a single-branch density matrix");
    s->queue.clear();
    s->virt = false;
    DevGuard g(s->device);
    const size_t vb = sizeof(double) << s->nq;
    double* dv = (double*)work_alloc(s->device, vb, s->stream);
    QB_CUDA(cudaMemcpyAsync(dv, values, vb, cudaMemcpyHostToDevice, s->stream));
    qb_launch_fill_diag(s->ctx(), s->d, s->nq, dv);
    work_free(s->device, vb, dv, s->stream);
    s->stats.bytes_moved += s->bytes();
    QB_API_END
}

int qb_upload(qb_state* s, const void* host, size_t bytes) {
    QB_API_BEGIN
    QB_REQUIRE(s && host, "NULL argument");
    QB_REQUIRE(bytes == s->bytes(), "upload: byte count does not match the state size");
    s->queue.clear();
    s->virt = false;
    DevGuard g(s->device);
    QB_CUDA(cudaMemcpyAsync(s->d, host, bytes, cudaMemcpyHostToDevice, s->stream));
    QB_CUDA(cudaStreamSynchronize(s->stream));
    QB_API_END
}

int qb_download(qb_state* s, void* host, size_t bytes) {
    QB_API_BEGIN
    QB_REQUIRE(s && host, "NULL argument");
    QB_REQUIRE(bytes == s->bytes(), "download: byte count does not match the state size");
    s->flush();
    DevGuard g(s->device);
    QB_CUDA(cudaMemcpyAsync(host, s->d, bytes, cudaMemcpyDeviceToHost, s->stream));
    QB_CUDA(cudaStreamSynchronize(s->stream));
    QB_API_END
}

int qb_download_range(qb_state* s, uint64_t first, uint64_t count, void* host) {
    QB_API_BEGIN
    QB_REQUIRE(s && host, "NULL argument");
    QB_REQUIRE(first + count <= s->total(), "download_range: out of bounds");
    s->flush();
    DevGuard g(s->device);
    QB_CUDA(cudaMemcpyAsync(host, s->d + first, count * sizeof(cplx), cudaMemcpyDeviceToHost, s->stream));
    QB_CUDA(cudaStreamSynchronize(s->stream));
    QB_API_END
}

int qb_apply_gate(qb_state* s, const double* matrix, int k, const int* target_bits, uint64_t control_mask) {
    QB_API_BEGIN
    QB_REQUIRE(s && matrix && target_bits, "NULL argument");
    QB_REQUIRE(k >= 1 && k <= QB_BIG_MAXK, "gate size out of range");
    QB_REQUIRE(k <= s->nq, "gate has more qubits than the register");
    const cplx* m = (const cplx*)matrix;
    uint64_t tmask = 0;
    for (int i = 0; i < k; i++) {
        QB_REQUIRE(target_bits[i] >= 0 && target_bits[i] < s->nq, "target bit out of range");
        QB_REQUIRE(!((tmask >> target_bits[i]) & 1ull), "duplicate target bit");
        tmask |= 1ull << target_bits[i];
    }
    QB_REQUIRE(s->nq >= 64 || (control_mask >> s->nq) == 0, "control bit out of range");
    QB_REQUIRE((tmask & control_mask) == 0, "control overlaps target");
    if (s->kind == QB_KET) {
        s->enqueue(qb_classify(m, k, target_bits, control_mask));
    } else {
        // rho <- U rho U^dagger on vec(rho): U on the row bits, conj(U) on the column bits
        int tb_row[QB_BIG_MAXK];
        for (int i = 0; i < k; i++) tb_row[i] = target_bits[i] + s->nq;
        s->enqueue(qb_classify(m, k, tb_row, control_mask << s->nq));
        const size_t D2 = (size_t)1 << (2 * k);
        std::vector<cplx> mc(D2);
        for (size_t i = 0; i < D2; i++) mc[i] = C(m[i].x, -m[i].y);
        s->enqueue(qb_classify(mc.data(), k, target_bits, control_mask));
    }
    QB_API_END
}

int qb_apply_gates(qb_state* s, int ngates, const int* ks, const int* target_bits, const uint64_t* control_masks,
                   const double* matrices) {
    QB_REQUIRE_API(s && ks && target_bits && control_masks && matrices && ngates >= 0);
    size_t moff = 0;
    for (int g = 0; g < ngates; g++) {
        if (ks[g] < 1 || ks[g] > QB_BIG_MAXK) { g_err = "gate size out of range"; return QB_ERR_ARG; }
        const int rc = qb_apply_gate(s, matrices + moff, ks[g], target_bits + (size_t)g * QB_BIG_MAXK, control_masks[g]);
        if (rc != QB_OK) return rc;
        moff += 2 * ((size_t)1 << (2 * ks[g]));
    }
    return QB_OK;
}

int qb_apply_gate_rc(qb_state* s, const double* row_matrix, const double* col_matrix, int k, const int* target_bits) {
    QB_API_BEGIN
    QB_REQUIRE(s && target_bits, "NULL argument");
    QB_REQUIRE(s->kind == QB_DM, "apply_gate_rc is defined for density matrices");
    QB_REQUIRE(k >= 1 && k <= QB_BIG_MAXK && k <= s->nq, "gate size out of range");
    uint64_t tmask = 0;
    int tb_row[QB_BIG_MAXK];
    for (int i = 0; i < k; i++) {
        QB_REQUIRE(target_bits[i] >= 0 && target_bits[i] < s->nq, "target bit out of range");
        QB_REQUIRE(!((tmask >> target_bits[i]) & 1ull), "duplicate target bit");
        tmask |= 1ull << target_bits[i];
        tb_row[i] = target_bits[i] + s->nq;
    }
    if (row_matrix) s->enqueue(qb_classify((const cplx*)row_matrix, k, tb_row, 0));
    if (col_matrix) s->enqueue(qb_classify((const cplx*)col_matrix, k, target_bits, 0));
    QB_API_END
}

int qb_apply_swap(qb_state* s, int bit_a, int bit_b) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    QB_REQUIRE(bit_a >= 0 && bit_a < s->nq && bit_b >= 0 && bit_b < s->nq, "swap bit out of range");
    if (bit_a == bit_b) return QB_OK;
    s->flush();
    DevGuard g(s->device);
    int lo = std::min(bit_a, bit_b), hi = std::max(bit_a, bit_b);
    qb_launch_swap(s->ctx(), s->d, s->total(), lo, hi);
    s->stats.bytes_moved += s->bytes();
    s->stats.state_passes++;
    if (s->kind == QB_DM) {
        qb_launch_swap(s->ctx(), s->d, s->total(), lo + s->nq, hi + s->nq);
        s->stats.bytes_moved += s->bytes();
        s->stats.state_passes++;
    }
    s->stats.gates_applied++;
    QB_API_END
}

int qb_apply_gate_batched(qb_state* s, const double* matrices, int k, const int* target_bits,
                          const uint64_t* control_masks, const uint8_t* enable) {
    QB_API_BEGIN
    QB_REQUIRE(s && matrices && target_bits, "NULL argument");
    QB_REQUIRE(k >= 1 && k <= QB_REG_MAXK, "batched gate: k must be 1..5");
    QB_REQUIRE(k <= s->nq, "gate has more qubits than the register");
    QB_REQUIRE(s->nbranch <= 65535, "batched gate: more than 65535 branches per launch");   // before anything is allocated
    s->flush();
    DevGuard g(s->device);
    const int64_t B = s->nbranch;
    const size_t D2 = (size_t)1 << (2 * k);
    for (int64_t b = 0; b < B; b++) {
        uint64_t tmask = 0;
        for (int i = 0; i < k; i++) {
            int t = target_bits[b * k + i];
            QB_REQUIRE(t >= 0 && t < s->nq, "batched gate: target bit out of range");
            QB_REQUIRE(!((tmask >> t) & 1ull), "batched gate: duplicate target bit");
            tmask |= 1ull << t;
        }
        uint64_t cm = control_masks ? control_masks[b] : 0;
        QB_REQUIRE((cm >> s->nq) == 0 && (cm & tmask) == 0, "batched gate: bad control mask");
    }
    const int passes = s->kind == QB_KET ? 1 : 2;
    // device descriptor tables
    size_t mat_bytes = sizeof(cplx) * D2 * B, tb_bytes = sizeof(int) * k * B, cm_bytes = sizeof(uint64_t) * B;
    char* buf = nullptr;
    size_t off_tb = (mat_bytes + 255) & ~size_t(255), off_cm = (off_tb + tb_bytes + 255) & ~size_t(255),
           off_en = (off_cm + cm_bytes + 255) & ~size_t(255), tot = off_en + B + 256;
    buf = (char*)work_alloc(s->device, tot, s->stream);
    std::vector<cplx> hm((const cplx*)matrices, (const cplx*)matrices + D2 * B);
    std::vector<int> htb(target_bits, target_bits + (size_t)k * B);
    std::vector<uint64_t> hcm(B, 0);
    if (control_masks) hcm.assign(control_masks, control_masks + B);
    for (int pass = 0; pass < passes; pass++) {
        if (s->kind == QB_DM && pass == 0) {
            for (auto& t : htb) t += s->nq;
            for (auto& c : hcm) c <<= s->nq;
        } else if (s->kind == QB_DM && pass == 1) {
            htb.assign(target_bits, target_bits + (size_t)k * B);
            if (control_masks) hcm.assign(control_masks, control_masks + B); else std::fill(hcm.begin(), hcm.end(), 0);
            for (auto& v : hm) v.y = -v.y;
        }
        QB_CUDA(cudaMemcpyAsync(buf, hm.data(), mat_bytes, cudaMemcpyHostToDevice, s->stream));
        QB_CUDA(cudaMemcpyAsync(buf + off_tb, htb.data(), tb_bytes, cudaMemcpyHostToDevice, s->stream));
        QB_CUDA(cudaMemcpyAsync(buf + off_cm, hcm.data(), cm_bytes, cudaMemcpyHostToDevice, s->stream));
        if (enable) QB_CUDA(cudaMemcpyAsync(buf + off_en, enable, B, cudaMemcpyHostToDevice, s->stream));
        qb_launch_dense_batched(s->ctx(), k, s->d, s->nbits, B, (const cplx*)buf, (const int*)(buf + off_tb),
                                (const uint64_t*)(buf + off_cm), enable ? (const uint8_t*)(buf + off_en) : nullptr);
        if (pass + 1 < passes) QB_CUDA(cudaStreamSynchronize(s->stream));       // the tables are rewritten for the column pass
        s->stats.bytes_moved += s->bytes() * 2;
        s->stats.state_passes++;
    }
    work_free(s->device, tot, buf, s->stream);
    s->stats.gates_applied += B;
    QB_API_END
}

int qb_flush(qb_state* s) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    QB_API_END
}

int qb_sync(qb_state* s) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    DevGuard g(s->device);
    QB_CUDA(cudaStreamSynchronize(s->stream));
    QB_API_END
}

int qb_set_fusion(qb_state* s, int enabled) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    s->fusion = enabled != 0;
    QB_API_END
}

int qb_set_jit(qb_state* s, int mode) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    QB_REQUIRE(mode >= -1 && mode <= 2, "jit mode must be -1, 0, 1 or 2");
    s->flush();
    s->jit_mode = mode;
    QB_API_END
}

int qb_jit_info(uint64_t* kernels_compiled, uint64_t* cache_hits, double* compile_ms) {
    QB_API_BEGIN
    const QbJitStats st = qb_jit_stats();
    if (kernels_compiled) *kernels_compiled = st.kernels_compiled;
    if (cache_hits) *cache_hits = st.cache_hits;
    if (compile_ms) *compile_ms = st.compile_ms;
    QB_API_END
}

int qb_jit_check(int nbits, int ngates, const int* ks, const int* target_bits, const uint64_t* control_masks,
                 const double* matrices, int* ncompiled, const char* cubin_dir) {
    QB_API_BEGIN
    QB_REQUIRE(ks && target_bits && control_masks && matrices, "NULL argument");
    std::vector<QGate> gates;
    size_t moff = 0;
    for (int g = 0; g < ngates; g++) {
        QB_REQUIRE(ks[g] >= 1 && ks[g] <= QB_BIG_MAXK, "gate size out of range");
        QGate q = qb_classify((const cplx*)(matrices + moff), ks[g], target_bits + (size_t)g * QB_BIG_MAXK, control_masks[g]);
        moff += 2 * ((size_t)1 << (2 * ks[g]));
        if (!qb_is_identity(q)) gates.push_back(q);
    }
    QtPlanOptions opt;
    if (const char* e = getenv("QBOT_B200_TILE_M")) { const int v = atoi(e); if (v == 11 || v == 12) opt.M = v; }
    opt.R = QT_MAXR;          // the specialiser's own plan shape (32 amplitudes per thread) unless overridden
    if (const char* e = getenv("QBOT_B200_JIT_R")) { const int v = atoi(e); if (v == 4 || v == 5) opt.R = v; }
    opt.search_trials = nbits >= 29 ? 128 : nbits >= 27 ? 32 : nbits >= 23 ? 8 : 1;      // the engine's search effort (qb_tile.cu get_plan)
    if (const char* e = getenv("QBOT_B200_PLAN_TRIALS")) opt.search_trials = atoi(e);
    std::vector<QGate> planned;
    std::vector<QtPlanStep> steps = qt_plan_best(gates, nbits, opt, &planned);
    int n = 0;
    if (!cubin_dir) {
        // no artefacts wanted: compile through the parallel path the engine uses
        std::vector<const uint8_t*> progs;
        for (const QtPlanStep& st : steps) if (st.fused) progs.push_back(st.program.data());
        const uint64_t before = qb_jit_stats().kernels_compiled;
        if (progs.size() == 1) qb_jit_compile(qb_jit_full_source(progs[0], nullptr, nullptr), nullptr);      // the pool takes >= 2
        else qb_jit_precompile(progs);
        if (ncompiled) *ncompiled = progs.size() == 1 ? 1 : (int)(qb_jit_stats().kernels_compiled - before);
        return QB_OK;
    }
    for (const QtPlanStep& st : steps) {
        if (!st.fused) continue;
        // (QBOT_B200_JIT_CHECK_VIRTUAL=1: the first sweep as its virtual-basis variant -- what runs right after qb_init_basis)
        const bool virt = n == 0 && getenv("QBOT_B200_JIT_CHECK_VIRTUAL") != nullptr;
        const std::string src = qb_jit_full_source(st.program.data(), nullptr, nullptr, virt);
        const std::vector<char> cubin = qb_jit_compile(src, nullptr);
        if (cubin_dir) {
            const std::string base = std::string(cubin_dir) + "/sweep_" + std::to_string(n);
            FILE* f = fopen((base + ".cubin").c_str(), "wb");
            QB_REQUIRE(f, "cannot write cubin");
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
            f = fopen((base + ".cu").c_str(), "w");
            if (f) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
        }
        n++;
    }
    if (ncompiled) *ncompiled = n;
    QB_API_END
}

// outcome weights ---------------------------------------------------------------------------
static void run_bins(qb_state* s, const int* bits, int m, std::vector<cplx>& host_out) {
    const int nb = s->nq;     // binned index: ket -> amplitude index, dm -> diagonal index
    QB_REQUIRE(m >= 0 && m <= nb, "probs: bad number of target bits");
    QB_REQUIRE(m <= 26, "probs: too many outcome bits");
    uint64_t seen = 0;
    for (int t = 0; t < m; t++) {
        QB_REQUIRE(bits[t] >= 0 && bits[t] < nb, "probs: bit out of range");
        QB_REQUIRE(!((seen >> bits[t]) & 1ull), "probs: duplicate bit");
        seen |= 1ull << bits[t];
    }
    s->flush();
    DevGuard g(s->device);
    BinArgs a;
    memset(&a, 0, sizeof(a));
    a.src = s->d;
    a.mode = s->kind == QB_KET ? 0 : 1;
    a.nb = nb;
    a.elem_stride = s->kind == QB_KET ? 1 : ((1ull << s->nq) + 1ull);
    a.branch_stride = s->per_branch();
    a.c = std::min(nb, 11);
    a.nchunks = 1ull << (nb - a.c);
    a.nfold = 0; a.ml = 0;
    for (int p = 0; p < a.c; p++) {
        if ((seen >> p) & 1ull) a.lowt[a.ml++] = p;
        else a.foldbits[a.nfold++] = p;
    }
    // chunks that differ only in non-target index bits feed the same outcomes: up to 64 of them
    // are summed by one block (fewer folds, fewer partials, 2 MB of loads per block)
    for (int p = a.c; p < nb && a.ngroup < 6; p++)
        if (!((seen >> p) & 1ull)) a.groupbits[a.ngroup++] = p - a.c;
    BinFinalArgs f;
    memset(&f, 0, sizeof(f));
    f.m = m; f.c = a.c; f.nb = nb; f.nchunks = a.nchunks; f.ml = a.ml;
    for (int x = 0; x < a.ngroup; x++) f.groupmask |= 1ull << a.groupbits[x];
    for (int t = 0; t < m; t++) {
        f.tbits[t] = bits[t];
        f.lowrank[t] = -1;
        if (bits[t] < a.c) for (int r = 0; r < a.ml; r++) if (a.lowt[r] == bits[t]) f.lowrank[t] = r;
    }
    size_t npartial = (size_t)s->nbranch * a.nchunks << a.ml;
    size_t nout = (size_t)s->nbranch << m;
    cplx* dpart = (cplx*)work_alloc(s->device, sizeof(cplx) * npartial, s->stream);
    cplx* dout = (cplx*)work_alloc(s->device, sizeof(cplx) * nout, s->stream);
    a.partial = dpart; f.partial = dpart; f.out = dout;
    qb_launch_bins(s->ctx(), a, s->nbranch);
    qb_launch_bins_final(s->ctx(), f, s->nbranch);
    host_out.resize(nout);
    QB_CUDA(cudaMemcpyAsync(host_out.data(), dout, sizeof(cplx) * nout, cudaMemcpyDeviceToHost, s->stream));
    QB_CUDA(cudaStreamSynchronize(s->stream));
    work_free(s->device, sizeof(cplx) * npartial, dpart, s->stream);
    work_free(s->device, sizeof(cplx) * nout, dout, s->stream);
    s->stats.bytes_moved += (s->kind == QB_KET ? s->bytes() : (sizeof(cplx) * (size_t)s->nbranch << s->nq));
}

int qb_probs(qb_state* s, const int* bits, int m, double* out) {
    QB_API_BEGIN
    QB_REQUIRE(s && out && (bits || m == 0), "NULL argument");
    std::vector<cplx> h;
    run_bins(s, bits, m, h);
    for (size_t i = 0; i < h.size(); i++) out[i] = s->kind == QB_KET ? h[i].x : std::hypot(h[i].x, h[i].y);
    QB_API_END
}

int qb_probs_basis(qb_state* s, const int* bits, int m, const double* basis, int b, double* out) {
    if (!basis) return qb_probs(s, bits, m, out);
    QB_API_BEGIN
    QB_REQUIRE(s && out && bits, "NULL argument");
    QB_REQUIRE(b >= 1 && b <= 5 && m >= b && m % b == 0, "probs_basis: the basis size must divide the number of bits");
    QB_REQUIRE(m <= s->nq && m <= 26, "probs_basis: too many outcome bits");
    {   // the bits become gate targets below: check them before anything is queued (run_bins checks again, too late for that)
        uint64_t seen = 0;
        for (int t = 0; t < m; t++) {
            QB_REQUIRE(bits[t] >= 0 && bits[t] < s->nq, "probs_basis: bit out of range");
            QB_REQUIRE(!((seen >> bits[t]) & 1ull), "probs_basis: duplicate bit");
            seen |= 1ull << bits[t];
        }
    }
    s->flush();
    DevGuard g(s->device);
    // rotate a scratch copy so that basis ket j of every group sits at group index j, then read the
    // computational-frame weights; the register itself stays as it is (peek must not change it)
    std::unique_ptr<qb_state> tmp(new_state(s->kind, s->nq, s->nbranch, s->device, nullptr, (void*)s->stream));
    tmp->fusion = s->fusion;
    tmp->jit_mode = 0;                 // one-off gate list: never worth a specialised kernel
    QB_CUDA(cudaMemcpyAsync(tmp->d, s->d, s->bytes(), cudaMemcpyDeviceToDevice, s->stream));
    const cplx* W = (const cplx*)basis;
    for (int f = 0; f < m / b; f++) {
        const int* tb = bits + f * b;
        if (s->kind == QB_KET) {
            tmp->enqueue(qb_classify(W, b, tb, 0));
        } else {
            int tb_row[8];
            for (int i = 0; i < b; i++) tb_row[i] = tb[i] + s->nq;
            tmp->enqueue(qb_classify(W, b, tb_row, 0));
            tmp->enqueue(qb_classify(W, b, tb, 0));
        }
    }
    std::vector<cplx> h;
    run_bins(tmp.get(), bits, m, h);
    s->stats.kernel_launches += tmp->stats.kernel_launches;
    s->stats.bytes_moved += tmp->stats.bytes_moved + 2 * s->bytes();
    for (size_t i = 0; i < h.size(); i++) out[i] = s->kind == QB_KET ? h[i].x : std::hypot(h[i].x, h[i].y);
    QB_API_END
}

int qb_norm2(qb_state* s, double* out) {
    QB_API_BEGIN
    QB_REQUIRE(s && out, "NULL argument");
    std::vector<cplx> h;
    run_bins(s, nullptr, 0, h);
    for (size_t i = 0; i < h.size(); i++) out[i] = h[i].x;
    QB_API_END
}

int qb_project_renorm(qb_state* s, const int* bits, int m, uint64_t outcome) {
    QB_API_BEGIN
    QB_REQUIRE(s && bits, "NULL argument");
    QB_REQUIRE(s->kind == QB_KET, "project_renorm is defined for kets only");
    QB_REQUIRE(s->nbranch == 1, "project_renorm: single-branch states only");
    QB_REQUIRE(m >= 1 && m <= s->nq && outcome < (1ull << m), "project_renorm: bad outcome");
    std::vector<cplx> h;
    run_bins(s, bits, m, h);
    double p = h[outcome].x;
    QB_REQUIRE(p > 0.0, "project_renorm: outcome has zero probability");
    uint64_t mask = 0, want = 0;
    for (int t = 0; t < m; t++) {
        mask |= 1ull << bits[t];
        if ((outcome >> (m - 1 - t)) & 1ull) want |= 1ull << bits[t];
    }
    DevGuard g(s->device);
    qb_launch_project(s->ctx(), s->d, s->total(), mask, want, 1.0 / std::sqrt(p));
    s->stats.bytes_moved += s->bytes() * 2;
    s->stats.state_passes++;
    QB_API_END
}

// density-matrix structure ops ------------------------------------------------------------------
int qb_ptrace(qb_state* s, const int* keep_bits, int nkeep, qb_state** out) {
    QB_API_BEGIN
    QB_REQUIRE(s && out && (keep_bits || nkeep == 0), "NULL argument");
    QB_REQUIRE(s->nbranch == 1, "ptrace needs a single-branch state");
    QB_REQUIRE(nkeep >= 0 && nkeep <= s->nq, "ptrace: bad keep count");
    QB_REQUIRE(s->kind == QB_DM || nkeep <= 13, "ptrace of a ket: at most 13 kept qubits");
    s->flush();
    DevGuard g(s->device);
    PtraceArgs a;
    memset(&a, 0, sizeof(a));
    uint64_t seen = 0;
    for (int i = 0; i < nkeep; i++) {
        QB_REQUIRE(keep_bits[i] >= 0 && keep_bits[i] < s->nq, "ptrace: bit out of range");
        QB_REQUIRE(!((seen >> keep_bits[i]) & 1ull), "ptrace: duplicate bit");
        seen |= 1ull << keep_bits[i];
        a.keepb[i] = keep_bits[i];
    }
    a.nq = s->nq; a.nkeep = nkeep; a.ntr = 0;
    for (int p = 0; p < s->nq; p++) if (!((seen >> p) & 1ull)) a.trb[a.ntr++] = p;
    qb_state* o = new_state(QB_DM, nkeep, 1, s->device, nullptr, s->owns_stream ? nullptr : (void*)s->stream);
    o->fusion = s->fusion;
    a.rho = s->d; a.out = o->d;
    // order the new handle's stream after ours
    if (s->kind == QB_DM) qb_launch_ptrace(s->ctx(), a);
    else qb_launch_ket_rdm(s->ctx(), a);          // Tr_rest |psi><psi| straight from the amplitudes
    s->stats.bytes_moved += (sizeof(cplx) << (s->nq + nkeep)) + o->bytes();
    *out = o;
    QB_API_END
}

int qb_scatter_product(qb_state* a, qb_state* b, const int* a_bits, const int* b_bits, const double* scale, qb_state** out) {
    QB_API_BEGIN
    QB_REQUIRE(a && out && a_bits, "NULL argument");
    QB_REQUIRE(a->kind == QB_DM && a->nbranch == 1, "scatter_product: A must be a single-branch density matrix");
    if (b) QB_REQUIRE(b->kind == QB_DM && b->nbranch == 1 && b->device == a->device && b_bits, "scatter_product: bad B");
    a->flush();
    if (b) b->flush();
    DevGuard g(a->device);
    const int na = a->nq, nb = b ? b->nq : 0, n = na + nb;
    ScatterArgs sa;
    memset(&sa, 0, sizeof(sa));
    uint64_t seen = 0;
    for (int i = 0; i < na; i++) {
        QB_REQUIRE(a_bits[i] >= 0 && a_bits[i] < n && !((seen >> a_bits[i]) & 1ull), "scatter_product: bad A position");
        seen |= 1ull << a_bits[i]; sa.abits[i] = a_bits[i];
    }
    for (int i = 0; i < nb; i++) {
        QB_REQUIRE(b_bits[i] >= 0 && b_bits[i] < n && !((seen >> b_bits[i]) & 1ull), "scatter_product: bad B position");
        seen |= 1ull << b_bits[i]; sa.bbits[i] = b_bits[i];
    }
    qb_state* o = new_state(QB_DM, n, 1, a->device, nullptr, a->owns_stream ? nullptr : (void*)a->stream);
    o->fusion = a->fusion;
    sa.a = a->d; sa.b = (b && nb > 0) ? b->d : nullptr; sa.out = o->d; sa.n = n; sa.na = na; sa.nb = nb;
    cplx extra = C(1, 0);
    bool has = false;
    if (b && nb == 0) {   // 1x1 factor
        cplx h;
        QB_CUDA(cudaStreamSynchronize(b->stream));
        QB_CUDA(cudaMemcpy(&h, b->d, sizeof(cplx), cudaMemcpyDeviceToHost));
        extra = h; has = true;
    }
    if (scale) {
        cplx sc = C(scale[0], scale[1]);
        extra = has ? C(extra.x * sc.x - extra.y * sc.y, extra.x * sc.y + extra.y * sc.x) : sc;
        has = true;
    }
    sa.has_scale = has ? 1 : 0; sa.scale = extra;
    if (b) order_after(a->stream, b);
    const size_t tab_bytes = sizeof(uint32_t) * 2 * ((size_t)1 << n);
    uint32_t* tabs = (uint32_t*)work_alloc(a->device, tab_bytes, a->stream);
    qb_launch_scatter(a->ctx(), sa, tabs);
    work_free(a->device, tab_bytes, tabs, a->stream);
    a->stats.bytes_moved += o->bytes();
    *out = o;
    QB_API_END
}

int qb_mix(qb_state* const* states, const double* probs, int count, qb_state** out) {
    QB_API_BEGIN
    QB_REQUIRE(states && probs && out && count >= 1, "bad argument");
    qb_state* s0 = states[0];
    QB_REQUIRE(s0, "NULL state");
    // an ensemble sum_i p_i rho_i is a density-matrix notion (density.py:49-58); sum_i p_i psi_i is not a state
    if (s0->kind != QB_DM) throw qb_error(QB_ERR_STATE, "mix: needs density matrices (qb_outer turns a ket into one)");
    for (int i = 0; i < count; i++) {
        QB_REQUIRE(states[i] && states[i]->kind == s0->kind && states[i]->nq == s0->nq &&
                   states[i]->nbranch == s0->nbranch && states[i]->device == s0->device, "mix: states differ in shape");
        states[i]->flush();
    }
    DevGuard g(s0->device);
    for (int i = 1; i < count; i++) order_after(s0->stream, states[i]);
    qb_state* o = new_state(s0->kind, s0->nq, s0->nbranch, s0->device, nullptr, s0->owns_stream ? nullptr : (void*)s0->stream);
    o->fusion = s0->fusion;
    for (int first = 0; first < count; first += QB_MIX_MAX) {
        MixArgs a;
        memset(&a, 0, sizeof(a));
        a.count = std::min(QB_MIX_MAX, count - first);
        for (int i = 0; i < a.count; i++) { a.src[i] = states[first + i]->d; a.p[i] = probs[first + i]; }
        a.accumulate = first > 0; a.out = o->d; a.total = s0->total();
        qb_launch_mix(s0->ctx(), a);
    }
    s0->stats.bytes_moved += s0->bytes() * (count + 1);
    *out = o;
    QB_API_END
}

int qb_mix_branches(qb_state* s, const double* probs, qb_state** out) {
    QB_API_BEGIN
    QB_REQUIRE(s && probs && out, "NULL argument");
    if (s->kind != QB_DM) throw qb_error(QB_ERR_STATE, "mix_branches: needs a batch of density matrices");
    s->flush();
    DevGuard g(s->device);
    qb_state* o = new_state(s->kind, s->nq, 1, s->device, nullptr, s->owns_stream ? nullptr : (void*)s->stream);
    o->fusion = s->fusion;
    double* dp = (double*)work_alloc(s->device, sizeof(double) * s->nbranch, s->stream);
    QB_CUDA(cudaMemcpyAsync(dp, probs, sizeof(double) * s->nbranch, cudaMemcpyHostToDevice, s->stream));
    qb_launch_mix_branches(s->ctx(), s->d, dp, s->nbranch, s->per_branch(), o->d);
    work_free(s->device, sizeof(double) * s->nbranch, dp, s->stream);
    s->stats.bytes_moved += s->bytes() + o->bytes();
    *out = o;
    QB_API_END
}

int qb_outer(qb_state* ket, int conj, qb_state** out) {
    QB_API_BEGIN
    QB_REQUIRE(ket && out, "NULL argument");
    QB_REQUIRE(ket->kind == QB_KET && ket->nbranch == 1, "outer: needs a single-branch ket");
    QB_REQUIRE(2 * ket->nq <= 40, "outer: density matrix would be too large");
    ket->flush();
    DevGuard g(ket->device);
    qb_state* o = new_state(QB_DM, ket->nq, 1, ket->device, nullptr, ket->owns_stream ? nullptr : (void*)ket->stream);
    o->fusion = ket->fusion;
    qb_launch_outer(ket->ctx(), ket->d, o->d, ket->nq, conj);
    ket->stats.bytes_moved += o->bytes();
    *out = o;
    QB_API_END
}

int qb_broadcast(qb_state* src, qb_state* dst) {
    QB_API_BEGIN
    QB_REQUIRE(src && dst, "NULL argument");
    QB_REQUIRE(src->nbranch == 1 && src->kind == dst->kind && src->nq == dst->nq && src->device == dst->device,
               "broadcast: shapes differ");
    src->flush();
    dst->queue.clear();
    DevGuard g(src->device);
    order_after(dst->stream, src);
    qb_launch_broadcast(dst->ctx(), src->d, dst->d, src->per_branch(), dst->nbranch);
    // single-branch views of dst (qb_create_external without a stream) work on the device's default
    // stream like dst itself, so they are ordered after this copy; a dst on a foreign stream is
    // synchronised here because its views cannot be
    if (dst->stream != default_stream(dst->device)) QB_CUDA(cudaStreamSynchronize(dst->stream));
    dst->stats.bytes_moved += dst->bytes() + src->bytes();
    QB_API_END
}

// sharded kets ------------------------------------------------------------------------------
int qb_buffer_alloc(int device, size_t bytes, void** out_dev) {
    QB_API_BEGIN
    QB_REQUIRE(out_dev && bytes > 0, "bad argument");
    DevGuard g(device);
    QB_CUDA(cudaMalloc(out_dev, bytes));
    QB_API_END
}

int qb_buffer_free(int device, void* dev) {
    QB_API_BEGIN
    if (dev) {
        DevGuard g(device);
        QB_CUDA(cudaDeviceSynchronize());
        QB_CUDA(cudaFree(dev));
    }
    QB_API_END
}

int qb_rebind(qb_state* s, void* amplitudes_dev) {
    QB_API_BEGIN
    QB_REQUIRE(s && amplitudes_dev, "NULL argument");
    QB_REQUIRE(!s->owns, "rebind: the handle owns its memory (use qb_create_external)");
    s->flush();
    s->d = (cplx*)amplitudes_dev;
    QB_API_END
}

int qb_ipc_export(int device, void* dev, void* handle64) {
    QB_API_BEGIN
    QB_REQUIRE(dev && handle64, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DevGuard g(device);
    cudaIpcMemHandle_t h;
    QB_CUDA(cudaIpcGetMemHandle(&h, dev));
    memcpy(handle64, &h, 64);
    QB_API_END
}

int qb_ipc_open(int device, const void* handle64, void** out_dev) {
    QB_API_BEGIN
    QB_REQUIRE(handle64 && out_dev, "NULL argument");
    DevGuard g(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    QB_CUDA(cudaIpcOpenMemHandle(out_dev, h, cudaIpcMemLazyEnablePeerAccess));
    QB_API_END
}

int qb_ipc_close(int device, void* dev) {
    QB_API_BEGIN
    if (dev) {
        DevGuard g(device);
        QB_CUDA(cudaIpcCloseMemHandle(dev));
    }
    QB_API_END
}

int qb_permute_scatter_sub(qb_state* s, const int* src_bit_of_dst_bit, int ndst_bits, uint64_t src_fixed_mask, uint64_t src_fixed_value,
                           int chunk_bits, void* const* chunk_dst, int first_chunk, void* cuda_stream, int max_ctas) {
    QB_API_BEGIN
    QB_REQUIRE(s && src_bit_of_dst_bit && chunk_dst, "NULL argument");
    QB_REQUIRE(first_chunk >= 0 && first_chunk < (1 << chunk_bits), "permute_scatter: first_chunk out of range");
    QB_REQUIRE(ndst_bits >= 0 && ndst_bits <= s->nbits, "permute_scatter: bad number of destination bits");
    QB_REQUIRE(ndst_bits - chunk_bits >= 8, "permute_scatter: chunks must hold at least 256 amplitudes");
    QB_REQUIRE(s->kind == QB_KET && s->nbranch == 1, "permute_scatter: single-branch kets only");
    QB_REQUIRE(chunk_bits >= 0 && chunk_bits <= 4 && chunk_bits <= ndst_bits, "permute_scatter: chunk_bits must be 0..4");
    QB_REQUIRE(__builtin_popcountll(src_fixed_mask) == s->nbits - ndst_bits && (src_fixed_value & ~src_fixed_mask) == 0 &&
               (s->nbits >= 64 || (src_fixed_mask >> s->nbits) == 0), "permute_scatter: the fixed source bits do not complete the destination bits");
    s->flush();
    DevGuard g(s->device);
    PermArgs a;
    memset(&a, 0, sizeof(a));
    uint64_t seen = src_fixed_mask;
    for (int d = 0; d < ndst_bits; d++) {
        const int f = src_bit_of_dst_bit[d];
        QB_REQUIRE(f >= 0 && f < s->nbits && !((seen >> f) & 1ull), "permute_scatter: not a permutation of the index bits");
        seen |= 1ull << f;
        if (f == d) a.fixed_mask |= 1ull << d;
        else {
            QB_REQUIRE(a.nmoved < QB_PERM_MAXMOVED, "permute_scatter: too many moved bits");
            a.from[a.nmoved] = (uint8_t)f; a.to[a.nmoved] = (uint8_t)d; a.nmoved++;
        }
    }
    a.in = s->d;
    a.src_or = src_fixed_value;
    a.total = 1ull << ndst_bits;
    a.chunk_shift = ndst_bits - chunk_bits;
    a.chunk_bits = chunk_bits;
    a.unit_bits = std::min(12, a.chunk_shift);
    a.first_chunk = first_chunk;
    for (int c = 0; c < (1 << chunk_bits); c++) {
        QB_REQUIRE(chunk_dst[c], "permute_scatter: NULL chunk destination");
        QB_REQUIRE(chunk_dst[c] != (void*)s->d, "permute_scatter: destination aliases the source");
        a.dst[c] = (cplx*)chunk_dst[c];
    }
    LaunchCtx ctx = s->ctx();
    if (cuda_stream) ctx.stream = (cudaStream_t)cuda_stream;      // the caller orders this stream against the handle's (events)
    qb_launch_permute_scatter(ctx, a, max_ctas);
    QB_CUDA(cudaGetLastError());
    s->stats.bytes_moved += (16ull << ndst_bits) * 2;
    if (ndst_bits == s->nbits) s->stats.state_passes++;
    QB_API_END
}

int qb_permute_scatter(qb_state* s, const int* src_bit_of_dst_bit, int chunk_bits, void* const* chunk_dst, int first_chunk) {
    if (!s) { g_err = "NULL argument"; return QB_ERR_ARG; }
    return qb_permute_scatter_sub(s, src_bit_of_dst_bit, s->nbits, 0, 0, chunk_bits, chunk_dst, first_chunk, nullptr, 0);
}

int qb_plan_queue(qb_state* s, int park_bits, int* nsteps, int* head, int* tail) {
    QB_API_BEGIN
    QB_REQUIRE(s && nsteps && head && tail, "NULL argument");
    QB_REQUIRE(!s->pending_plan, "plan_queue: the previous plan has not been finished (qb_finish_queue)");
    *nsteps = *head = *tail = 0;
    if (s->queue.empty()) return QB_OK;
    if (!s->fusion || !qb_engine_available()) { s->flush(); return QB_OK; }
    DevGuard gd(s->device);
    std::vector<QGate> q;
    q.swap(s->queue);
    *nsteps = qb_engine_plan_pending(s, q, park_bits, head, tail);
    QB_API_END
}

int qb_run_steps(qb_state* s, int from, int to, int part, int nparts, int sm_limit) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    DevGuard gd(s->device);
    qb_engine_run_pending(s, from, to, part, nparts, sm_limit);
    QB_API_END
}

int qb_finish_queue(qb_state* s) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    qb_engine_finish_pending(s);
    QB_API_END
}

int qb_set_sm_limit(qb_state* s, int nsms) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    const int all = sm_count_of(s->device);
    s->sms = (nsms > 0 && nsms < all) ? nsms : all;
    QB_API_END
}

static unsigned long long* flag_timeouts(int device) {
    static std::mutex mu;
    static std::map<int, unsigned long long*>& m = *new std::map<int, unsigned long long*>;
    std::lock_guard<std::mutex> lk(mu);
    auto it = m.find(device);
    if (it != m.end()) return it->second;
    unsigned long long* p = nullptr;
    QB_CUDA(cudaMalloc((void**)&p, sizeof(unsigned long long)));
    QB_CUDA(cudaMemset(p, 0, sizeof(unsigned long long)));
    m[device] = p;
    return p;
}

int qb_signal_flags(int device, void* cuda_stream, void* const* flags, int n, uint64_t value) {
    QB_API_BEGIN
    QB_REQUIRE(flags && n >= 0 && n <= QB_FLAG_MAXPEERS, "signal_flags: 0..16 flag pointers");
    DevGuard g(device);
    FlagPtrs f;
    memset(&f, 0, sizeof(f));
    f.n = n;
    for (int i = 0; i < n; i++) { QB_REQUIRE(flags[i], "signal_flags: NULL flag"); f.p[i] = (unsigned long long*)flags[i]; }
    if (n) qb_launch_signal_flags(cuda_stream ? (cudaStream_t)cuda_stream : default_stream(device), f, value);
    QB_CUDA(cudaGetLastError());
    QB_API_END
}

int qb_wait_flags(int device, void* cuda_stream, void* const* flags, int n, uint64_t value) {
    QB_API_BEGIN
    QB_REQUIRE(flags && n >= 0 && n <= QB_FLAG_MAXPEERS, "wait_flags: 0..16 flag pointers");
    DevGuard g(device);
    FlagPtrs f;
    memset(&f, 0, sizeof(f));
    f.n = n;
    for (int i = 0; i < n; i++) { QB_REQUIRE(flags[i], "wait_flags: NULL flag"); f.p[i] = (unsigned long long*)flags[i]; }
    if (n) qb_launch_wait_flags(cuda_stream ? (cudaStream_t)cuda_stream : default_stream(device), f, value, flag_timeouts(device));
    QB_CUDA(cudaGetLastError());
    QB_API_END
}

int qb_flag_timeouts(int device, uint64_t* count) {
    QB_API_BEGIN
    QB_REQUIRE(count, "NULL argument");
    DevGuard g(device);
    unsigned long long v = 0;
    QB_CUDA(cudaMemcpy(&v, flag_timeouts(device), sizeof(v), cudaMemcpyDeviceToHost));
    *count = v;
    QB_API_END
}

int qb_copy_async(int device, void* dst, const void* src, size_t bytes, void* cuda_stream) {
    QB_API_BEGIN
    QB_REQUIRE(dst && src, "NULL argument");
    DevGuard g(device);
    // device-to-device (local or a peer mapping): the copy engines move it, no SM is involved
    QB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, cuda_stream ? (cudaStream_t)cuda_stream : default_stream(device)));
    QB_API_END
}

int qb_compute_stream(int device, void** cuda_stream_out) {
    QB_API_BEGIN
    QB_REQUIRE(cuda_stream_out, "NULL argument");
    QB_REQUIRE(device >= 0 && device < device_count_cached(), "no such CUDA device");
    DevGuard g(device);
    *cuda_stream_out = (void*)default_stream(device);
    QB_API_END
}

int qb_get_stats(const qb_state* s, qb_stats* out) {
    QB_API_BEGIN
    QB_REQUIRE(s && out, "NULL argument");
    *out = s->stats;
    QB_API_END
}

int qb_reset_stats(qb_state* s) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    memset(&s->stats, 0, sizeof(s->stats));
    QB_API_END
}

int qb_timer_start(qb_state* s) {
    QB_API_BEGIN
    QB_REQUIRE(s, "NULL state");
    s->flush();
    DevGuard g(s->device);
    if (!s->ev0) { QB_CUDA(cudaEventCreate(&s->ev0)); QB_CUDA(cudaEventCreate(&s->ev1)); }
    QB_CUDA(cudaEventRecord(s->ev0, s->stream));
    QB_API_END
}

int qb_timer_stop(qb_state* s, float* ms) {
    QB_API_BEGIN
    QB_REQUIRE(s && ms && s->ev0, "timer not started");
    s->flush();
    DevGuard g(s->device);
    QB_CUDA(cudaEventRecord(s->ev1, s->stream));
    QB_CUDA(cudaEventSynchronize(s->ev1));
    QB_CUDA(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    QB_API_END
}

}  // extern "C"
