// engine.cu — runtime, column registry and kernel orchestration.
// See engine.cuh for the interface and kernels.cuh for the reference loops
// (file:line) each kernel replaces.
#include "engine.cuh"
#include "kernels.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace b200 {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string t_last_error;

void set_last_error(const std::string &msg) { t_last_error = msg; }
const std::string &last_error_string() { return t_last_error; }

[[noreturn]] void fatal(const char *file, int line, const char *what, const char *detail) {
    // reference convention: diagnostics on stderr, exit(2) (rhjoin.c:285-286)
    fprintf(stderr, "b200join fatal: %s:%d: %s: %s\n", file, line, what, detail ? detail : "");
    fflush(stderr);
    exit(2);
}

// ---------------------------------------------------------------------------
// device + per-thread context
// ---------------------------------------------------------------------------
static std::once_flag g_init_once;
static int            g_device   = -1;
static int            g_sm_count = 0;
static bool           g_profiling = false;
std::atomic<uint64_t> g_launches{0};
thread_local uint64_t t_launches = 0;

// keep freed blocks in the stream-ordered pool instead of returning them
static void configure_device_pool(int device) {
    {
        // Keep the per-thread local-memory reservation at its high-water mark.  A few scatter instances spill some
        // tens of bytes; by default the driver shrinks the reservation again after such a kernel, and the next launch
        // of one resizes it under a device-wide synchronisation: config 5 showed 50-900 ms stalls in queries whose
        // kernels sum to a few milliseconds, in some runs and not in others.  (Refused when another library already
        // fixed the flags of an active context: then it stays as it is.)
        unsigned flags = 0;
        cudaSetDevice(device);
        if (cudaGetDeviceFlags(&flags) != cudaSuccess) flags = 0;
        if (cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax) != cudaSuccess) cudaGetLastError();
        cudaGetLastError();
    }
    cudaMemPool_t pool;
    B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
}

static void init_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        fatal(__FILE__, __LINE__, "cudaGetDeviceCount",
              "no CUDA device: this library has no CPU fallback");
    if (device < 0) {
        const char *env = getenv("B200_DEVICE");
        device          = env ? atoi(env) : 0;
    }
    B200_REQUIRE(device < count, "device index out of range");
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    g_sm_count = prop.multiProcessorCount;
    g_device   = device;
    configure_device_pool(device);
    Tuning &t = tuning();
    if (const char *v = getenv("B200_RADIX_BITS")) t.radix_bits = atoi(v);
    if (const char *v = getenv("B200_FORCE_KEY64")) t.force_key64 = atoi(v);
    if (const char *v = getenv("B200_CAP32")) t.cap32 = (uint32_t)atoi(v);
    if (const char *v = getenv("B200_CAP64")) t.cap64 = (uint32_t)atoi(v);
    if (const char *v = getenv("B200_TAG64")) t.tag64 = atoi(v);
    if (const char *v = getenv("B200_SLICE")) t.slice = (uint32_t)atoi(v);
    if (const char *v = getenv("B200_DEBUG")) t.debug = atoi(v);
    if (const char *v = getenv("B200_SCATTER_CFG")) t.scatter_cfg = atoi(v);
    if (const char *v = getenv("B200_EARLY_MAT")) t.early_mat = atoi(v);
    if (const char *v = getenv("B200_OPT_PARTITION")) t.opt_partition = atoi(v);
    if (const char *v = getenv("B200_CARRY32")) t.carry32 = atoi(v);
    if (const char *v = getenv("B200_CARRY_PROBE")) t.carry_probe = atoi(v);
    if (const char *v = getenv("B200_L2_FETCH")) {
        // granularity hint for L2 fills of the random payload gathers (32, 64 or 128)
        B200_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(v)));
    }
}

static int g_requested_device = -1;
void request_device(int device) { g_requested_device = device; }

void ensure_init() {
    std::call_once(g_init_once, [] { init_device(g_requested_device); });
}
int  device_index() { return g_device; }
int  sm_count() { return g_sm_count; }
int g_slowlog_ms = [] {
    const char *v = getenv("B200_SLOWLOG");
    return v ? atoi(v) : 0;
}();
void slow_checkpoint(const char *file, int line, const char *what) {
    static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    static thread_local const char *last_what = "(start)";
    static thread_local int         last_line = 0;
    const auto   now = std::chrono::steady_clock::now();
    const double ms  = std::chrono::duration<double, std::milli>(now - last).count();
    if (ms > (double)g_slowlog_ms)
        fprintf(stderr, "b200 slow: %.1f ms between line %d [%.60s] and %s:%d [%.60s]\n", ms, last_line, last_what, file, line, what);
    last      = now;
    last_what = what;
    last_line = line;
}
bool profiling_enabled() { return g_profiling; }
void set_profiling(bool on) { g_profiling = on; }

// The reference's unmodified handler.o never calls b200_init: warm CUDA up (primary context, eager module
// load, memory pool) from a helper thread as soon as the library is loaded, i.e. while the process is still
// reading relation names (the contest's untimed preparation phase), instead of inside the first query.
// A thread, not the constructor itself: the constructor may run before this library's kernels are
// registered with the runtime.  B200_EAGER_INIT=0 disables it.
static std::thread *g_warmup = nullptr;
static void join_warmup();
__attribute__((constructor)) static void eager_init() {
    const char *v = getenv("B200_EAGER_INIT");
    if (v && atoi(v) == 0) return;
    if (const char *m = getenv("B200_MODULE_LOADING")) setenv("CUDA_MODULE_LOADING", m, 1);
    else setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    atexit(join_warmup);   // never tear the process down underneath a half-initialised runtime
    g_warmup = new std::thread([] {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            return;   // no device: the first operator reports it
        }
        ensure_init();
        void *p = nullptr;
        if (cudaMalloc(&p, 1 << 20) == cudaSuccess) cudaFree(p);
        cudaGetLastError();
    });
}
static void join_warmup() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (g_warmup && g_warmup->joinable() && std::this_thread::get_id() != g_warmup->get_id()) g_warmup->join();
    });
}

Tuning &tuning() {
    static Tuning t;
    return t;
}

Context::~Context() {
    // process teardown may already have destroyed the CUDA context: ignore errors
    for (auto &kv : timers) {
        if (kv.second.start) cudaEventDestroy(kv.second.start);
        if (kv.second.stop) cudaEventDestroy(kv.second.stop);
    }
    if (h_scratch) cudaFreeHost(h_scratch);
    if (d_scratch) cudaFree(d_scratch);
    if (own_stream) cudaStreamDestroy(own_stream);
}

// Contexts are pooled: a host that creates fresh worker threads for every batch (host/b200_engine.c before its
// persistent pool, a ThreadPoolExecutor per execute_batch call) would otherwise leak a stream plus pinned and
// device scratch per thread, and pay their (device-synchronising) allocation inside the batch.  A thread takes a
// context of its device from the free list and hands it back when it exits; no CUDA call is made at thread exit
// (destroying CUDA objects from a thread_local destructor races with runtime teardown).
static std::mutex              g_ctx_mu;
static std::vector<Context *>  g_ctx_free;
static thread_local int        t_device = -1;   // device this thread works on (-1: the process default)

struct ContextHolder {
    Context *c = nullptr;
    ~ContextHolder() {
        if (!c) return;
        c->stream = c->own_stream;   // an adopted caller stream is not ours to keep
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        g_ctx_free.push_back(c);
    }
};

void set_thread_device(int device) { t_device = device; }
int  thread_device() { return t_device >= 0 ? t_device : g_device; }

Context &ctx() {
    static thread_local ContextHolder h;
    const int want = t_device;
    if (h.c && (want < 0 || h.c->device == want)) return *h.c;
    join_warmup();
    ensure_init();
    const int device = want >= 0 ? want : g_device;
    B200_CUDA(cudaSetDevice(device));
    if (h.c) {   // the thread moved to another device: hand the old context back
        h.c->stream = h.c->own_stream;
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        g_ctx_free.push_back(h.c);
        h.c = nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (size_t i = 0; i < g_ctx_free.size(); ++i)
            if (g_ctx_free[i]->device == device) {
                h.c = g_ctx_free[i];
                g_ctx_free.erase(g_ctx_free.begin() + (long)i);
                break;
            }
    }
    if (!h.c) {
        Context *c = new Context();
        c->device  = device;
        if (device != g_device) configure_device_pool(device);
        B200_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        B200_CUDA(cudaMallocHost(&c->h_scratch, 64 * sizeof(unsigned long long)));
        B200_CUDA(cudaMalloc(&c->d_scratch, 64 * sizeof(unsigned long long)));
        h.c = c;
    }
    return *h.c;
}

static bool nvtx_on() {
    static const bool on = [] {
        const char *v = getenv("B200_NVTX");
        return v && atoi(v) != 0;
    }();
    return on;
}
TimedScope::TimedScope(const char *name) {
    if (nvtx_on()) {
        nvtxRangePushA(name);
        nvtx = true;
    }
    if (!g_profiling) return;
    Context &c = ctx();
    t          = &c.timers[name];
    if (!t->start) {
        B200_CUDA(cudaEventCreate(&t->start));
        B200_CUDA(cudaEventCreate(&t->stop));
    }
    if (t->used) {   // a scope of this name ran before: keep its time (profiling mode may synchronise)
        float ms = 0.f;
        if (cudaEventSynchronize(t->stop) == cudaSuccess && cudaEventElapsedTime(&ms, t->start, t->stop) == cudaSuccess)
            t->earlier_ms += ms;
    }
    t->used = true;
    ++t->scopes;
    B200_CUDA(cudaEventRecord(t->start, c.stream));
}
TimedScope::~TimedScope() {
    if (t) cudaEventRecord(t->stop, ctx().stream);
    if (nvtx) nvtxRangePop();
}

constexpr size_t kBigBuf       = 64ull << 20;    // blocks from this size up are recycled through the context
constexpr size_t kBigCacheMax  = 24ull << 30;    // ... up to this many bytes per context

DevBuf::DevBuf(size_t nbytes, cudaStream_t s) : bytes(nbytes), capacity(nbytes ? nbytes : 16), stream(s) {
    if (nbytes >= kBigBuf) {
        Context &c    = ctx();
        int      best = -1;
        for (size_t i = 0; i < c.big_cache.size(); ++i) {
            const auto &b = c.big_cache[i];
            if (b.stream == s && b.bytes >= nbytes && b.bytes <= 2 * nbytes &&
                (best < 0 || b.bytes < c.big_cache[(size_t)best].bytes))
                best = (int)i;
        }
        if (best >= 0) {
            ptr      = c.big_cache[(size_t)best].ptr;
            capacity = c.big_cache[(size_t)best].bytes;
            c.big_cached -= capacity;
            c.big_cache.erase(c.big_cache.begin() + best);
            return;
        }
    }
    // never hand out NULL: a 0-row table must still read as "active"
    B200_CUDA(cudaMallocAsync(&ptr, capacity, s));
}
DevBuf::~DevBuf() {
    if (!ptr) return;
    if (capacity >= kBigBuf) {
        // stream order makes the hand-over safe: whoever takes the block enqueues its work on the same stream, behind
        // whatever still uses it
        Context &c = ctx();
        if (c.device == thread_device() && c.big_cached + capacity <= kBigCacheMax) {
            c.big_cache.push_back(Context::BigBlock{ptr, capacity, stream});
            c.big_cached += capacity;
            return;
        }
    }
    cudaFreeAsync(ptr, stream);
}

uint64_t read_counter(const unsigned long long *d_ptr) {
    Context &c = ctx();
    B200_CUDA(cudaMemcpyAsync(c.h_scratch, d_ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                              c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return c.h_scratch[0];
}

// SMs the streaming kernels of the calling thread leave free (multi-GPU broadcast plan, pull variant: the probe-side
// scatter is a persistent grid that would otherwise own every SM, and the fetch kernel beside it must stay resident)
static thread_local int t_reserved_sms = 0;
void set_reserved_sms(int n) { t_reserved_sms = n < 0 ? 0 : n; }

int grid_for(uint64_t work_items, int per_block, int max_blocks_per_sm) {
    uint64_t blocks = (work_items + per_block - 1) / per_block;
    uint64_t cap    = (uint64_t)std::max(1, sm_count() - t_reserved_sms) * max_blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------
// column registry (the hook after relation_map.c:InitRelationMap)
// ---------------------------------------------------------------------------
struct ColumnEntry {
    uint64_t *owned   = nullptr;   // cudaMalloc'ed copy (nullptr for external)
    DevColumn col;
};
static std::mutex                                       g_col_mu;
static std::unordered_map<const uint64_t *, ColumnEntry> g_columns;

static uint64_t device_column_max(const uint64_t *d, uint64_t n) {
    Context &c = ctx();
    B200_CUDA(cudaMemsetAsync(c.d_scratch + 8, 0, sizeof(unsigned long long), c.stream));
    if (n) {
        column_max_kernel<<<grid_for(n, 256 * 8, 8), 256, 0, c.stream>>>(d, n, c.d_scratch + 8);
        B200_LAUNCH_CHECK();
    }
    return read_counter(c.d_scratch + 8);
}

void device_column_stats(const uint64_t *d_col, uint64_t n, uint64_t *out_min, uint64_t *out_max, uint64_t *out_distinct) {
    Context &c = ctx();
    *out_min = *out_max = *out_distinct = 0;
    if (n == 0) return;
    unsigned long long init[3] = {~0ull, 0ull, 0ull};
    B200_CUDA(cudaMemcpyAsync(c.d_scratch + 40, init, sizeof(init), cudaMemcpyHostToDevice, c.stream));
    column_minmax_kernel<<<grid_for(n, 256 * 8, 8), 256, 0, c.stream>>>(d_col, n, c.d_scratch + 40, c.d_scratch + 41);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(c.h_scratch + 40, c.d_scratch + 40, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                              c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    const uint64_t lo = c.h_scratch[40], hi = c.h_scratch[41];
    // relation_map.c:64-83: min(u - l + 1, 50 000 000) entries; below the cap entry v - l, else (v - l) % 5 000 000
    uint64_t size = hi - lo + 1;
    if (size > 50000000ull || size == 0) size = 50000000ull;
    const uint64_t modulus = size < 50000000ull ? 0ull : 5000000ull;
    const uint64_t nwords  = (size + 31) / 32;
    DevBufPtr      bits    = dev_alloc(nwords * sizeof(uint32_t));
    B200_CUDA(cudaMemsetAsync(bits->ptr, 0, nwords * sizeof(uint32_t), c.stream));
    column_mark_kernel<<<grid_for(n, 256 * 8, 8), 256, 0, c.stream>>>(d_col, n, lo, modulus, bits->as<uint32_t>());
    B200_LAUNCH_CHECK();
    bitmap_count_kernel<<<grid_for(nwords, 256 * 8, 8), 256, 0, c.stream>>>(bits->as<uint32_t>(), nwords, c.d_scratch + 42);
    B200_LAUNCH_CHECK();
    *out_min      = lo;
    *out_max      = hi;
    *out_distinct = read_counter(c.d_scratch + 42);
}

static ColumnEntry upload_entry(const uint64_t *host_col, uint64_t n) {
    B200_REQUIRE(n <= kMaxRows, "relation has more than 2^32-1 rows (32-bit device row ids)");
    Context    &c = ctx();
    ColumnEntry e;
    B200_CUDA(cudaMalloc(&e.owned, n ? n * sizeof(uint64_t) : 16));
    if (n)
        B200_CUDA(cudaMemcpyAsync(e.owned, host_col, n * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                  c.stream));
    e.col.d       = e.owned;
    e.col.n       = n;
    e.col.max_val = device_column_max(e.owned, n);   // also synchronises the copy
    if (n) {
        uint64_t mx = 0;
        device_column_stats(e.owned, n, &e.col.min_val, &mx, &e.col.distinct);
    }
    return e;
}

void register_host_column(const uint64_t *host_col, uint64_t n, bool replace) {
    ensure_init();
    {
        std::unique_lock<std::mutex> lk(g_col_mu);
        auto it = g_columns.find(host_col);
        if (it != g_columns.end()) {
            if (!replace) return;
            if (it->second.owned && it->second.col.n == n) {
                // refresh in place (synchronous): new data may exceed the old maximum, which selects the 32-bit
                // kernels and the carried payloads — recompute it; device_column_max also synchronises the copy
                uint64_t *dst = it->second.owned;
                lk.unlock();
                Context &c = ctx();
                if (n) B200_CUDA(cudaMemcpyAsync(dst, host_col, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
                const uint64_t mx = device_column_max(dst, n);
                lk.lock();
                auto again = g_columns.find(host_col);
                if (again != g_columns.end() && again->second.owned == dst) {
                    again->second.col.max_val  = mx;
                    again->second.col.min_val  = 0;   // unknown until asked for again
                    again->second.col.distinct = 0;
                }
                return;
            }
            if (it->second.owned) cudaFree(it->second.owned);
            g_columns.erase(it);
        }
    }
    ColumnEntry e = upload_entry(host_col, n);
    std::lock_guard<std::mutex> lk(g_col_mu);
    // two threads may have uploaded the same column concurrently (first use from worker threads): keep the
    // entry that is already in the map and free the loser's copy
    auto ins = g_columns.emplace(host_col, e);
    if (!ins.second) {
        if (replace || ins.first->second.col.n != n) {
            if (ins.first->second.owned) cudaFree(ins.first->second.owned);
            ins.first->second = e;
        } else if (e.owned) {
            cudaFree(e.owned);
        }
    }
}

void register_device_column(const uint64_t *host_key, const uint64_t *dev, uint64_t n, uint64_t max_val) {
    ensure_init();
    B200_REQUIRE(n <= kMaxRows, "relation has more than 2^32-1 rows (32-bit device row ids)");
    ColumnEntry e;
    e.col.d       = dev;
    e.col.n       = n;
    e.col.max_val = max_val;
    std::lock_guard<std::mutex> lk(g_col_mu);
    auto it = g_columns.find(host_key);
    if (it != g_columns.end() && it->second.owned) cudaFree(it->second.owned);
    g_columns[host_key] = e;
}

void unregister_column(const uint64_t *host_col) {
    std::lock_guard<std::mutex> lk(g_col_mu);
    auto it = g_columns.find(host_col);
    if (it == g_columns.end()) return;
    if (it->second.owned) cudaFree(it->second.owned);
    g_columns.erase(it);
}

DevColumn lookup_column(const uint64_t *host_col, uint64_t n) {
    {
        std::lock_guard<std::mutex> lk(g_col_mu);
        auto it = g_columns.find(host_col);
        if (it != g_columns.end()) {
            // a different length under the same host pointer is a stale entry
            // (the caller freed and re-used the memory): upload again
            if (it->second.col.n == n) return it->second.col;
            if (it->second.owned) cudaFree(it->second.owned);
            g_columns.erase(it);
        }
    }
    // never registered (the reference's unmodified handler.o): upload on first use
    register_host_column(host_col, n, false);
    std::lock_guard<std::mutex> lk(g_col_mu);
    return g_columns[host_col].col;
}

// largest value of a column registered under this DEVICE pointer, or UINT64_MAX when unknown
uint64_t known_column_max(const uint64_t *dev_ptr) {
    std::lock_guard<std::mutex> lk(g_col_mu);
    for (const auto &kv : g_columns)
        if (kv.second.col.d == dev_ptr) return kv.second.col.max_val;
    return UINT64_MAX;
}

void unregister_all_columns() {
    std::lock_guard<std::mutex> lk(g_col_mu);
    for (auto &kv : g_columns)
        if (kv.second.owned) cudaFree(kv.second.owned);
    g_columns.clear();
}

// ---------------------------------------------------------------------------
// kernel launch helpers
// ---------------------------------------------------------------------------
static cudaStream_t launch_stream() { return ctx().stream; }

// Dynamic shared memory above 48 KB is opt-in per kernel function.  The attribute is raised ONCE per (device, kernel)
// to everything the SM offers beyond the kernel's static shared memory: worker threads launch the same instance with
// different sizes concurrently, and re-setting the attribute to each launch's own size let one thread lower it under
// another thread's launch ("invalid argument" at 4 workers on config 5 x100).
static std::mutex                               g_smem_mu;
static std::map<std::pair<int, const void *>, size_t> g_smem_max;
template <typename KernelT>
static void allow_smem(KernelT kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    const int    dev = ctx().device;
    const void  *fn  = reinterpret_cast<const void *>(kernel);
    size_t       max_dyn;
    {
        std::lock_guard<std::mutex> lk(g_smem_mu);
        auto it = g_smem_max.find({dev, fn});
        if (it == g_smem_max.end()) {
            cudaFuncAttributes fa{};
            int                optin = 0;
            B200_CUDA(cudaFuncGetAttributes(&fa, kernel));
            B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            const size_t m = (size_t)optin > fa.sharedSizeBytes ? (size_t)optin - fa.sharedSizeBytes : 0;
            B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m));
            it = g_smem_max.emplace(std::make_pair(dev, fn), m).first;
        }
        max_dyn = it->second;
    }
    if (bytes > max_dyn) {
        char msg[160];
        snprintf(msg, sizeof(msg), "a kernel needs %zu bytes of shared memory, the SM offers %zu", bytes, max_dyn);
        fatal(__FILE__, __LINE__, "shared memory", msg);
    }
}
// shared memory one CTA may use (dynamic + static), with room for the kernels' few static words
static size_t smem_budget() {
    int optin = 0;
    B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx().device));
    return (size_t)optin - 2048;
}

constexpr int kPartNT = 512;

// scatter configurations (Tuning::scatter_cfg): threads, keys per thread, CTAs per SM
//   0: 512 x 16 (32-bit keys) / 512 x 8 (64-bit), 2 CTAs/SM  -> 64 KB stage each
//   1: 512 x 32 / 512 x 16, 1 CTA/SM                          -> 128 KB stage
//   2: 1024 x 16 / 1024 x 8, 1 CTA/SM                         -> 128 KB stage, 32 warps
//   3: 256 x 48 / 256 x 24, 2 CTAs/SM                         -> 96 KB stage each: two independent phase streams
//      per SM (one CTA copies out while the other ranks/stages) with tiles of 12 K tuples; not yet measured
template <typename KeyT, int CFG> struct PartCfg;
template <> struct PartCfg<uint32_t, 0> { static constexpr int NT = 512, U = 16, MINB = 2; };
template <> struct PartCfg<uint64_t, 0> { static constexpr int NT = 512, U = 8, MINB = 2; };
template <> struct PartCfg<uint32_t, 1> { static constexpr int NT = 512, U = 32, MINB = 1; };
template <> struct PartCfg<uint64_t, 1> { static constexpr int NT = 512, U = 16, MINB = 1; };
template <> struct PartCfg<uint32_t, 2> { static constexpr int NT = 1024, U = 16, MINB = 1; };
template <> struct PartCfg<uint64_t, 2> { static constexpr int NT = 1024, U = 8, MINB = 1; };
template <> struct PartCfg<uint32_t, 3> { static constexpr int NT = 256, U = 48, MINB = 2; };
template <> struct PartCfg<uint64_t, 3> { static constexpr int NT = 256, U = 24, MINB = 2; };

template <typename KeyT>
static void launch_hist(const KeySrc &src, int bits, uint32_t *ghist, int ctas_per_sm = 4, const PredSet *ps = nullptr) {
    constexpr int U    = PartCfg<KeyT, 0>::U;
    const size_t  smem = (size_t)(1u << bits) * sizeof(uint32_t);
    if (ps && ps->npred) {
        B200_REQUIRE(src.ids == nullptr, "fused predicates apply to base relations");
        auto k = radix_hist_kernel<kPartNT, U, KeyT, true>;
        allow_smem(k, smem);
        k<<<grid_for(src.n, kPartNT * U, ctas_per_sm), kPartNT, smem, launch_stream()>>>(src, (uint32_t)bits, ghist, *ps);
    } else {
        auto k = radix_hist_kernel<kPartNT, U, KeyT>;
        allow_smem(k, smem);
        k<<<grid_for(src.n, kPartNT * U, ctas_per_sm), kPartNT, smem, launch_stream()>>>(src, (uint32_t)bits, ghist,
                                                                                        PredSet{});
    }
    B200_LAUNCH_CHECK();
}

// the tuned scatter carrying a 32-bit payload, skipping the rows opt.pred rules out (hot keys of the exchange plan)
static void launch_scatter_pred_carry(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    constexpr int NT   = PartCfg<uint32_t, 1>::NT;
    constexpr int U    = PartCfg<uint32_t, 1>::U;
    constexpr int MINB = PartCfg<uint32_t, 1>::MINB;
    const size_t  smem = (size_t)NT * U * sizeof(Tup32) + 3 * (size_t)(1u << bits) * sizeof(uint32_t) +
                        (opt.pred.hot_keys ? kHotBitmapWords * sizeof(uint32_t) : 0);
    auto          k    = radix_scatter_kernel<NT, U, MINB, uint32_t, false, true, true>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, MINB), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor, static_cast<Tup32 *>(out),
                                                                 opt);
    B200_LAUNCH_CHECK();
}

// scatter with the relation's filter predicates folded into its load stage (plain and histogram-free instance)
template <typename KeyT, bool OPT>
static void launch_scatter_pred(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    using TupT         = typename TupOf<KeyT>::type;
    constexpr int NT   = PartCfg<KeyT, 1>::NT;
    constexpr int U    = PartCfg<KeyT, 1>::U;
    constexpr int MINB = PartCfg<KeyT, 1>::MINB;
    B200_REQUIRE(src.ids == nullptr && (opt.pred.npred > 0 || opt.pred.hot_keys), "fused predicates apply to base relations");
    const size_t smem = (size_t)NT * U * sizeof(TupT) + (OPT ? 4 : 3) * (size_t)(1u << bits) * sizeof(uint32_t) +
                        (opt.pred.hot_keys ? kHotBitmapWords * sizeof(uint32_t) : 0);
    auto         k    = radix_scatter_kernel<NT, U, MINB, KeyT, OPT, false, true>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, MINB), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor, static_cast<TupT *>(out),
                                                                 opt);
    B200_LAUNCH_CHECK();
}

template <typename KeyT, int CFG, bool OPT>
static void launch_scatter_c(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    using TupT            = typename TupOf<KeyT>::type;
    constexpr int NT      = PartCfg<KeyT, CFG>::NT;
    constexpr int U       = PartCfg<KeyT, CFG>::U;
    constexpr int MINB    = PartCfg<KeyT, CFG>::MINB;
    const size_t  smem    = (size_t)NT * U * sizeof(TupT) + (OPT ? 4 : 3) * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k       = radix_scatter_kernel<NT, U, MINB, KeyT, OPT>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, MINB), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor,
                                                                 static_cast<TupT *>(out), opt);
    B200_LAUNCH_CHECK();
}

template <typename KeyT>
static void launch_scatter(const KeySrc &src, int bits, uint32_t *cursor, void *out) {
    const OptArgs none{0, nullptr, nullptr, nullptr, PredSet{}};
    switch (tuning().scatter_cfg) {
        case 0: launch_scatter_c<KeyT, 0, false>(src, bits, cursor, out, none); break;
        case 2: launch_scatter_c<KeyT, 2, false>(src, bits, cursor, out, none); break;
        case 3: launch_scatter_c<KeyT, 3, false>(src, bits, cursor, out, none); break;
        default: launch_scatter_c<KeyT, 1, false>(src, bits, cursor, out, none); break;
    }
}

// the tuned scatter with a 32-bit payload column carried in the row-id slot (base columns only)
static void launch_scatter_carry_tuned(const KeySrc &src, int bits, uint32_t *cursor, void *out,
                                       const uint64_t *carry_col) {
    constexpr int NT   = PartCfg<uint32_t, 1>::NT;
    constexpr int U    = PartCfg<uint32_t, 1>::U;
    constexpr int MINB = PartCfg<uint32_t, 1>::MINB;
    const size_t  smem = (size_t)NT * U * sizeof(Tup32) + 3 * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k    = radix_scatter_kernel<NT, U, MINB, uint32_t, false, true>;
    allow_smem(k, smem);
    const OptArgs carry{0, nullptr, nullptr, carry_col, PredSet{}};
    k<<<grid_for(src.n, NT * U, MINB), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor,
                                                                 static_cast<Tup32 *>(out), carry);
    B200_LAUNCH_CHECK();
}

// histogram-free probe-side scatter: fixed regions of opt.opt_cap tuples + overflow
static void launch_scatter_opt(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    switch (tuning().scatter_cfg) {
        case 0: launch_scatter_c<uint32_t, 0, true>(src, bits, cursor, out, opt); break;
        case 2: launch_scatter_c<uint32_t, 2, true>(src, bits, cursor, out, opt); break;
        case 3: launch_scatter_c<uint32_t, 3, true>(src, bits, cursor, out, opt); break;
        default: launch_scatter_c<uint32_t, 1, true>(src, bits, cursor, out, opt); break;
    }
}
static void launch_scatter_opt64(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    launch_scatter_c<uint64_t, 1, true>(src, bits, cursor, out, opt);
}

// ... the same with a 32-bit probe-side SUM column carried in the row-id slot of the probe tuples
template <typename KeyT, int CFG, bool OPT>
static void launch_scatter_carry_c(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    using TupT         = typename TupOf<KeyT>::type;
    constexpr int NT   = PartCfg<KeyT, CFG>::NT;
    constexpr int U    = PartCfg<KeyT, CFG>::U;
    constexpr int MINB = PartCfg<KeyT, CFG>::MINB;
    const size_t  smem = (size_t)NT * U * sizeof(TupT) + (OPT ? 4 : 3) * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k    = radix_scatter_kernel<NT, U, MINB, KeyT, OPT, true>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, MINB), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor,
                                                                 static_cast<TupT *>(out), opt);
    B200_LAUNCH_CHECK();
}
static void launch_scatter_opt_carry(const KeySrc &src, int bits, uint32_t *cursor, void *out, const OptArgs &opt) {
    switch (tuning().scatter_cfg) {
        case 0: launch_scatter_carry_c<uint32_t, 0, true>(src, bits, cursor, out, opt); break;
        case 2: launch_scatter_carry_c<uint32_t, 2, true>(src, bits, cursor, out, opt); break;
        case 3: launch_scatter_carry_c<uint32_t, 3, true>(src, bits, cursor, out, opt); break;
        default: launch_scatter_carry_c<uint32_t, 1, true>(src, bits, cursor, out, opt); break;
    }
}

// build-side scatter with early-materialised projections (radix_scatter_pay_kernel)
template <typename KeyT, int NPAY>
static void launch_scatter_pay_n(const KeySrc &src, int bits, uint32_t *cursor, void *out, const PayArgs &pay) {
    using TupT         = typename TupOf<KeyT>::type;
    constexpr int NT   = 1024;   // 8192-tuple tiles, one CTA per SM
    constexpr int U    = 8;
    const size_t  smem = (size_t)NT * U * (sizeof(TupT) + 8 * NPAY) + 3 * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k    = radix_scatter_pay_kernel<NT, U, KeyT, NPAY, false>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, 1), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor, static_cast<TupT *>(out),
                                                               pay);
    B200_LAUNCH_CHECK();
}
// overflow tuples of an OPT scatter -> partition order (input = packed tuples, row id kept)
static void launch_scatter_tuples(const uint64_t *tuples, uint32_t n, int bits, uint32_t *cursor, void *out) {
    constexpr int NT   = 1024;
    constexpr int U    = 8;
    const size_t  smem = (size_t)NT * U * sizeof(Tup32) + 3 * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k    = radix_scatter_pay_kernel<NT, U, uint32_t, 0, true>;
    allow_smem(k, smem);
    KeySrc  src{tuples, nullptr, n};
    PayArgs pay{};
    k<<<grid_for(n, NT * U, 1), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor, static_cast<Tup32 *>(out), pay);
    B200_LAUNCH_CHECK();
}

// build-side scatter whose tuples carry a 32-bit payload in the row-id slot
template <typename KeyT>
static void launch_scatter_carry(const KeySrc &src, int bits, uint32_t *cursor, void *out, const PayArgs &pay) {
    using TupT         = typename TupOf<KeyT>::type;
    constexpr int NT   = 1024;
    constexpr int U    = 8;
    const size_t  smem = (size_t)NT * U * sizeof(TupT) + 3 * (size_t)(1u << bits) * sizeof(uint32_t);
    auto          k    = radix_scatter_pay_kernel<NT, U, KeyT, 0, false, true>;
    allow_smem(k, smem);
    k<<<grid_for(src.n, NT * U, 1), NT, smem, launch_stream()>>>(src, (uint32_t)bits, cursor, static_cast<TupT *>(out),
                                                                pay);
    B200_LAUNCH_CHECK();
}

template <typename KeyT>
static void launch_scatter_pay(const KeySrc &src, int bits, uint32_t *cursor, void *out, const PayArgs &pay, int npay) {
    if (pay.carry32) {
        launch_scatter_carry<KeyT>(src, bits, cursor, out, pay);
        return;
    }
    if constexpr (sizeof(KeyT) == 8) {
        B200_REQUIRE(npay == 0, "64-bit keys carry their one build-side SUM column in the tuple");
        launch_scatter_pay_n<KeyT, 0>(src, bits, cursor, out, pay);
        return;
    }
    if (npay == 0)
        launch_scatter_pay_n<KeyT, 0>(src, bits, cursor, out, pay);
    else if (npay == 1)
        launch_scatter_pay_n<KeyT, 1>(src, bits, cursor, out, pay);
    else
        launch_scatter_pay_n<KeyT, 2>(src, bits, cursor, out, pay);
}

constexpr int kJoinNT = 512;
constexpr int kJoinU  = 8;

template <typename KernelT>
static void launch_persistent_join(KernelT k, const JoinArgs &a, bool direct, size_t smem) {
    allow_smem(k, smem);
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kJoinNT, smem));
    B200_REQUIRE(occ >= 1, "join kernel does not fit on an SM");
    int grid = sm_count() * occ;
    if (direct && (int)a.n_items_direct < grid) grid = (int)a.n_items_direct;
    if (grid < 1) grid = 1;
    k<<<grid, kJoinNT, smem, ctx().stream>>>(a);
    B200_LAUNCH_CHECK();
}

// 64-bit keys: chained table (hash_join_kernel)
template <bool DIRECT>
static void launch_join64(const JoinArgs &a, int mode, size_t smem) {
    if (mode == MODE_COUNT)
        launch_persistent_join(hash_join_kernel<kJoinNT, kJoinU, uint64_t, DIRECT, MODE_COUNT, 0>, a, DIRECT, smem);
    else if (mode == MODE_WRITE)
        launch_persistent_join(hash_join_kernel<kJoinNT, kJoinU, uint64_t, DIRECT, MODE_WRITE, 0>, a, DIRECT, smem);
    else if (a.nproj <= 2)
        launch_persistent_join(hash_join_kernel<kJoinNT, kJoinU, uint64_t, DIRECT, MODE_SUM, 2>, a, DIRECT, smem);
    else if (a.nproj <= 4)
        launch_persistent_join(hash_join_kernel<kJoinNT, kJoinU, uint64_t, DIRECT, MODE_SUM, 4>, a, DIRECT, smem);
    else
        launch_persistent_join(hash_join_kernel<kJoinNT, kJoinU, uint64_t, DIRECT, MODE_SUM, kMaxProj>, a, DIRECT, smem);
}

// 32-bit keys, partitioned: tag table (tag_join_kernel)
constexpr int kJoinG = 4;
static uint32_t tag_slots_log2_for(uint32_t cap) {
    uint32_t l = 4;   // slots >= 1.75 x cap: load factor <= 0.57
    while ((uint64_t)4 << l < (uint64_t)7 * cap) ++l;
    return l;
}
template <typename KernelT>
static void launch_persistent_join_nt(KernelT k, const JoinArgs &a, size_t smem, int nt) {
    allow_smem(k, smem);
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nt, smem));
    B200_REQUIRE(occ >= 1, "join kernel does not fit on an SM");
    k<<<sm_count() * occ, nt, smem, ctx().stream>>>(a);
    B200_LAUNCH_CHECK();
}
template <int NT, int MINB>
static void launch_join32_cfg(const JoinArgs &a, int mode, size_t smem) {
    if (mode == MODE_COUNT)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_COUNT, 0>, a, smem, NT);
    else if (mode == MODE_WRITE)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_WRITE, 0>, a, smem, NT);
    else if (a.nproj <= 2)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, 2>, a, smem, NT);
    else if (a.nproj <= 4)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, 4>, a, smem, NT);
    else
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, kMaxProj>, a, smem, NT);
}
static bool tag_join_big(uint32_t cap, uint32_t slots_log2) {
    // two 512-thread CTAs per SM while the table allows it, else one 1024-thread CTA
    return ((size_t)4 << slots_log2) + (size_t)4 * cap + (size_t)8 * kTagQueue * 16 > 113 * 1024;
}
static size_t tag_join_smem(uint32_t cap, uint32_t slots_log2) {
    const int nw = tag_join_big(cap, slots_log2) ? 32 : 16;
    return ((size_t)4 << slots_log2) + (size_t)4 * cap + (size_t)8 * kTagQueue * nw;
}
// segmented build side (multi-GPU rank-major layout): fused SUM only
template <int NT, int MINB>
static void launch_join32_seg(const JoinArgs &a, size_t smem) {
    if (a.nproj <= 2)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, 2, true>, a, smem, NT);
    else if (a.nproj <= 4)
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, 4, true>, a, smem, NT);
    else
        launch_persistent_join_nt(tag_join_kernel<NT, MINB, kJoinG, MODE_SUM, kMaxProj, true>, a, smem, NT);
}
static void launch_join32(const JoinArgs &a, int mode, size_t smem) {
    if (a.nseg > 0) {
        B200_REQUIRE(mode == MODE_SUM, "a segmented build side is only joined into SUMs");
        if (tag_join_big(a.cap, a.slots_log2))
            launch_join32_seg<1024, 1>(a, smem);
        else
            launch_join32_seg<512, 2>(a, smem);
        return;
    }
    if (tag_join_big(a.cap, a.slots_log2))
        launch_join32_cfg<1024, 1>(a, mode, smem);
    else
        launch_join32_cfg<512, 2>(a, mode, smem);
}

// 64-bit keys, partitioned: the tag table with verified candidates (tag_join_kernel<K64>), 768 threads per CTA so
// that slots (128 KB) + links (68 KB) + the 16-byte queue entries of 24 warps fit one SM's shared memory
constexpr int kJoin64NT = 768;
constexpr int kJoin64G  = 2;
static size_t tag_join64_smem(uint32_t cap, uint32_t slots_log2) {
    return ((size_t)4 << slots_log2) + (size_t)4 * cap + (size_t)16 * kTagQueue * (kJoin64NT / 32);
}
static void launch_join_tag64(const JoinArgs &a, int mode) {
    const size_t smem = tag_join64_smem(a.cap, a.slots_log2);
    if (mode == MODE_COUNT)
        launch_persistent_join_nt(tag_join_kernel<kJoin64NT, 1, kJoin64G, MODE_COUNT, 0, false, true>, a, smem, kJoin64NT);
    else if (mode == MODE_WRITE)
        launch_persistent_join_nt(tag_join_kernel<kJoin64NT, 1, kJoin64G, MODE_WRITE, 0, false, true>, a, smem, kJoin64NT);
    else if (a.nproj <= 2)
        launch_persistent_join_nt(tag_join_kernel<kJoin64NT, 1, kJoin64G, MODE_SUM, 2, false, true>, a, smem, kJoin64NT);
    else
        launch_persistent_join_nt(tag_join_kernel<kJoin64NT, 1, kJoin64G, MODE_SUM, kMaxProj, false, true>, a, smem, kJoin64NT);
}

static void launch_join(const JoinArgs &a, bool key64, bool direct, int mode, bool chained64 = false) {
    if (key64 && !direct && chained64) {
        launch_join64<false>(a, mode, TableView<uint64_t>::bytes(a.cap, a.slots_log2));
    } else if (key64 && !direct) {
        launch_join_tag64(a, mode);
    } else if (key64) {
        const size_t smem = TableView<uint64_t>::bytes(a.cap, a.slots_log2);
        launch_join64<true>(a, mode, smem);
    } else {
        B200_REQUIRE(!direct, "32-bit keys are joined partitioned");
        launch_join32(a, mode, tag_join_smem(a.cap, a.slots_log2));
    }
}

uint32_t opt_region_cap(uint64_t n_probe, int bits);

static uint32_t ceil_log2(uint64_t v) {
    uint32_t l = 0;
    while ((1ull << l) < v) ++l;
    return l;
}

// ---------------------------------------------------------------------------
// radix partition (K3-K5), standalone for the parity tests
// ---------------------------------------------------------------------------
PartitionOut run_partition(const KeyVec &kv, int bits) {
    Context &c = ctx();
    B200_REQUIRE(bits >= 0 && bits <= tuning().max_bits, "radix bits out of range");
    PartitionOut  o;
    o.key64               = tuning().force_key64 || kv.max_val > 0xFFFFFFFFull;
    const uint32_t nparts = 1u << bits;
    o.hist                = dev_alloc((size_t)nparts * sizeof(uint32_t));
    DevBufPtr zeros       = dev_alloc((size_t)nparts * sizeof(uint32_t));
    DevBufPtr plan        = dev_alloc((size_t)(5 * (nparts + 1)) * sizeof(uint32_t));
    uint32_t *off_b = plan->as<uint32_t>(), *off_p = off_b + nparts + 1, *cur_b = off_p + nparts + 1,
             *cur_p = cur_b + nparts + 1, *items = cur_p + nparts + 1;
    B200_CUDA(cudaMemsetAsync(o.hist->ptr, 0, (size_t)nparts * sizeof(uint32_t), c.stream));
    B200_CUDA(cudaMemsetAsync(zeros->ptr, 0, (size_t)nparts * sizeof(uint32_t), c.stream));
    o.tuples = dev_alloc((size_t)kv.src.n * (o.key64 ? sizeof(Tup64) : sizeof(Tup32)));
    if (kv.src.n == 0) return o;
    if (o.key64)
        launch_hist<uint64_t>(kv.src, bits, o.hist->as<uint32_t>());
    else
        launch_hist<uint32_t>(kv.src, bits, o.hist->as<uint32_t>());
    partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(o.hist->as<uint32_t>(), zeros->as<uint32_t>(),
                                                          nparts, 1u, 1u, off_b, off_p, cur_b, cur_p, items,
                                                          zeros->as<uint32_t>(), 0u);
    B200_LAUNCH_CHECK();
    if (o.key64)
        launch_scatter<uint64_t>(kv.src, bits, cur_b, o.tuples->ptr);
    else
        launch_scatter<uint32_t>(kv.src, bits, cur_b, o.tuples->ptr);
    return o;
}

// Radix bits of a partitioned join: the fewest partitions whose build side — expected largest partition under a
// uniform split, mean + 5 sigma; a larger one is split into build chunks — fits one shared-memory table.  Fewer
// partitions mean longer runs per scatter tile.  The tag table (32-bit keys) needs radix_bits + slots_log2 >= 17.
int auto_radix_bits(uint64_t n_build, bool key64, bool chained64) {
    // (both key widths use the tag table, cap32 build tuples per table; only 32-bit keys need radix_bits +
    // slots_log2 >= 17 so that slot and tag identify the key.  chained64: the chained table of 64-bit keys, cap64)
    const Tuning  &t          = tuning();
    const uint32_t cap        = chained64 ? t.cap64 : t.cap32;
    const uint32_t slots_log2 = tag_slots_log2_for(cap);
    const int      min_bits   = key64 ? 1 : std::max(2, 17 - (int)slots_log2);
    if (t.radix_bits > 0) return std::max(min_bits, std::min(t.radix_bits, t.max_bits));
    auto fits = [&](int b) {
        const double mean = (double)n_build / (double)(1ull << b);
        return mean + 5.0 * sqrt(mean) <= (double)cap;
    };
    int bits = min_bits;
    while (bits < t.max_bits && !fits(bits)) ++bits;
    return bits;
}

// ---------------------------------------------------------------------------
// the join: RadixHashJoin (rhjoin.c:13-111) + optional fused SUM
// ---------------------------------------------------------------------------
JoinResult run_join(const KeyVec &R, const KeyVec &S, JoinOut mode, int nproj, const ProjDesc *proj) {
    Context    &c = ctx();
    Tuning     &t = tuning();
    JoinResult  res;
    B200_REQUIRE(R.src.n > 0 && S.src.n > 0, "run_join called with an empty side");
    B200_REQUIRE(nproj >= 0 && nproj <= kMaxProj, "too many fused projections");

    // rhjoin.c:118-134 indexes the smaller bucket side; here the smaller
    // relation is the build side for every partition.
    const bool    swapped = S.src.n < R.src.n;
    const KeyVec &B       = swapped ? S : R;
    const KeyVec &P       = swapped ? R : S;
    // A build side that fits one shared-memory table is joined without a
    // partition pass (DIRECT), by the chained-table kernel on 64-bit keys.
    const bool pred_b = B.preds.npred > 0, pred_p = P.preds.npred > 0;
    // (a relation with fused predicates is a large base relation — operators.cu fuses from 2^18 rows up — so it is
    // never the build side of an unpartitioned join)
    const bool direct = t.radix_bits <= 0 && B.src.n <= t.cap64 && !pred_b;
    // 32-bit keys (8-byte partition tuples, tag-table kernel) when every key fits
    const bool key64 = direct || t.force_key64 || B.max_val > 0xFFFFFFFFull || P.max_val > 0xFFFFFFFFull;
    // unpartitioned joins run the chained table (16-bit links, keys in shared memory), partitioned ones the tag table
    const bool chained64 = key64 && !direct && !t.tag64;   // the chained table for partitioned 64-bit-key joins
    const bool chained   = direct || chained64;
    const uint32_t cap = chained ? t.cap64 : t.cap32;
    B200_REQUIRE(cap >= 32 && cap <= (chained ? 65534u : 32766u), "table capacity out of range");
    const uint32_t slots_log2 = chained ? TableView<uint64_t>::slots_log2_for(cap) : tag_slots_log2_for(cap);
    const int bits = direct ? 0 : auto_radix_bits(B.src.n, key64, chained64);

    JoinArgs a;
    memset(&a, 0, sizeof(a));
    a.cap        = cap;
    a.slots_log2 = slots_log2;
    a.radix_bits = (uint32_t)bits;

    // one allocation for the small control arrays
    const uint32_t nparts = 1u << bits;
    DevBufPtr ctrl = dev_alloc((64 + 8 * (size_t)(nparts + 1)) * sizeof(uint32_t) + 64 * sizeof(unsigned long long));
    B200_CUDA(cudaMemsetAsync(ctrl->ptr, 0, ctrl->bytes, c.stream));
    unsigned long long *d_u64   = ctrl->as<unsigned long long>();   // [0] total, [1] out_cursor, [8..16) sums
    uint32_t           *d_u32   = reinterpret_cast<uint32_t *>(d_u64 + 64);
    uint32_t           *d_work  = d_u32;   // work counter
    uint32_t           *hist_b  = d_u32 + 64;
    uint32_t           *hist_p  = hist_b + nparts + 1;
    uint32_t           *off_b   = hist_p + nparts + 1;
    uint32_t           *off_p   = off_b + nparts + 1;
    uint32_t           *cur_b   = off_p + nparts + 1;
    uint32_t           *cur_p   = cur_b + nparts + 1;
    uint32_t           *items   = cur_p + nparts + 1;
    uint32_t           *cnt_p   = items + nparts + 1;
    uint32_t           *d_ovcnt = d_u32 + 1;   // overflow counter of the histogram-free scatter
    a.work_counter = d_work;
    a.total        = d_u64;
    a.out_cursor   = d_u64 + 1;
    a.sums         = d_u64 + 8;
    unsigned long long *d_valid_p = d_u64 + 2;   // DIRECT: probe rows that passed their fused predicates

    DevBufPtr tup_b, tup_p, ov_tup;
    bool      opt     = false;
    uint32_t  opt_cap = 0;
    DevBufPtr part_vals[kMaxProj];
    int       carry_k = -1, carry_p = -1;
    uint64_t  n_items = 0;
    if (direct) {
        a.src_b = B.src;
        a.src_p = P.src;
        if (pred_p) {
            a.pred_p  = P.preds;
            a.valid_p = d_valid_p;
        }
        // slices big enough to amortise rebuilding the table, small enough to
        // spread over the SMs
        uint64_t slice = std::max<uint64_t>({4096, 4ull * std::min<uint64_t>(B.src.n, cap),
                                             (P.src.n + 2ull * sm_count() - 1) / (2ull * sm_count())});
        slice          = std::min<uint64_t>(slice, t.slice);
        a.slice        = (uint32_t)slice;
        const uint64_t rc = (B.src.n + cap - 1) / cap;
        const uint64_t sc = (P.src.n + slice - 1) / slice;
        B200_REQUIRE(rc * sc < (1ull << 31), "too many join work items");
        a.sc_direct      = (uint32_t)sc;
        a.n_items_direct = (uint32_t)(rc * sc);
        n_items          = rc * sc;
    } else {
        // probe tuples per work item: the configured slice, but never so large that a small join is a handful of
        // items (config 3's last join: 25 M probe rows against 4 partitions were 96 items for 148 SMs) — about three
        // items per SM, each still long enough to amortise building its table
        a.slice = (uint32_t)std::min<uint64_t>(t.slice, std::max<uint64_t>(32768, (P.src.n / (3ull * sm_count()) + 1023) & ~1023ull));
        const size_t tsz = key64 ? sizeof(Tup64) : sizeof(Tup32);
        // Histogram-free probe side (fused SUM, 32-bit keys): every partition owns a region a few percent
        // above the uniform expectation; what does not fit overflows and is partitioned exactly afterwards.
        opt = t.opt_partition && P.src.n >= (1u << 20) && P.src.n <= (1u << 30) && !chained64;
        if (opt) opt_cap = opt_region_cap(P.src.n, bits);
        tup_b = dev_alloc((size_t)B.src.n * tsz);
        tup_p = dev_alloc(opt ? (size_t)opt_cap * nparts * tsz : (size_t)P.src.n * tsz);
        // early materialisation: the first two build-side projections of a fused
        // SUM travel with the build tuples into partition order (32-bit-key path)
        PayArgs pay{};
        int     npay = 0;
        if (mode == JoinOut::Sum && t.early_mat && !pred_b) {
            const int side_b  = swapped ? 1 : 0;   // proj[].side is relative to (R, S)
            int       n_build = 0, first = -1;
            for (int k = 0; k < nproj; ++k)
                if (proj[k].side == side_b) {
                    if (first < 0) first = k;
                    ++n_build;
                }
            // one build-side projection whose column is known to hold 32-bit values: carry it in the row-id slot
            // (no other projection reads the build row id then)
            if (n_build == 1 && t.carry32 && !chained64 && known_column_max(proj[first].col) <= 0xFFFFFFFFull) {
                carry_k      = first;
                pay.carry32  = 1;
                pay.col[0]   = proj[first].col;
                pay.ids[0]   = proj[first].ids;
                npay         = 1;   // selects the payload-aware scatter
            } else if (!key64) {   // (16-byte tuples leave no room to stage payload columns beside them)
                // 8192-tuple tiles of radix_scatter_pay_kernel: 8 + 8 per staged column bytes each, beside three bin
                // arrays (2^12 partitions leave room for one column, fewer partitions for two)
                int max_pay = 2;
                while (max_pay > 0 && 8192 * (sizeof(Tup32) + 8 * (size_t)max_pay) + 3 * (size_t)nparts * 4 > smem_budget())
                    --max_pay;
                for (int k = 0; k < nproj && npay < max_pay; ++k) {
                    if (proj[k].side != side_b) continue;
                    part_vals[k]  = dev_alloc((size_t)B.src.n * sizeof(uint64_t));
                    pay.col[npay] = proj[k].col;
                    pay.ids[npay] = proj[k].ids;
                    pay.out[npay] = part_vals[k]->as<uint64_t>();
                    ++npay;
                }
            }
        }
        // Probe-side early materialisation: ONE probe-side SUM column over the base relation whose values fit
        // 32 bits rides in the row-id slot of the probe tuples (nothing else reads the probe row id in a fused
        // SUM).  The scatter then streams that column (8 B per probe row, hidden under its L1-bound phases)
        // instead of the join gathering it at random (8 B per MATCH at ~38 G gathers/s), which pays
        // when matches are not rare: expected matches per probe row = build rows / key domain for a unique
        // build side (stats.c estimates join cardinalities the same way from the column ranges).
        if (mode == JoinOut::Sum && opt && t.carry32 && t.carry_probe && P.src.ids == nullptr && !pred_p) {
            const int side_p  = swapped ? 0 : 1;
            int       n_probe = 0, first = -1;
            for (int k = 0; k < nproj; ++k)
                if (proj[k].side == side_p) {
                    if (first < 0) first = k;
                    ++n_probe;
                }
            const double domain = (double)std::max(B.max_val, P.max_val) + 1.0;
            const double sel    = std::min(1.0, (double)B.src.n / domain);
            if (n_probe == 1 && proj[first].ids == nullptr && sel >= 1.0 / 24.0 &&
                (reinterpret_cast<uintptr_t>(proj[first].col) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(P.src.col) & 15) == 0 &&
                known_column_max(proj[first].col) <= 0xFFFFFFFFull)
                carry_p = first;
        }
        auto scatter_build = [&](uint32_t *cursor) {
            TimedScope ts("scatter_b");
            if (pred_b) {
                const OptArgs o{0, nullptr, nullptr, nullptr, B.preds, B.src.n};
                if (key64) launch_scatter_pred<uint64_t, false>(B.src, bits, cursor, tup_b->ptr, o);
                else launch_scatter_pred<uint32_t, false>(B.src, bits, cursor, tup_b->ptr, o);
            } else if (pay.carry32 && !key64 && B.src.ids == nullptr && pay.ids[0] == nullptr && B.src.n >= (1u << 18) &&
                       (reinterpret_cast<uintptr_t>(pay.col[0]) & 15) == 0)
                // a large base relation with its one SUM column carried: the tuned instance (16 K-tuple tiles)
                launch_scatter_carry_tuned(B.src, bits, cursor, tup_b->ptr, pay.col[0]);
            else if (npay > 0 && key64)
                launch_scatter_pay<uint64_t>(B.src, bits, cursor, tup_b->ptr, pay, npay);
            else if (npay > 0)
                launch_scatter_pay<uint32_t>(B.src, bits, cursor, tup_b->ptr, pay, npay);
            else if (key64)
                launch_scatter<uint64_t>(B.src, bits, cursor, tup_b->ptr);
            else
                launch_scatter<uint32_t>(B.src, bits, cursor, tup_b->ptr);
        };
        {
            TimedScope ts("hist_b");
            if (key64)
                launch_hist<uint64_t>(B.src, bits, hist_b, 4, &B.preds);
            else
                launch_hist<uint32_t>(B.src, bits, hist_b, 4, &B.preds);
        }
        if (opt) {
            // build side: exact (its histogram is cheap); hist_p is still all zero here
            partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, hist_p, nparts, cap, a.slice, off_b, off_p,
                                                                  cur_b, cur_p, items, cnt_p, 0u);
            B200_LAUNCH_CHECK();
            scatter_build(cur_b);
            init_opt_cursors_kernel<<<(nparts + 255) / 256, 256, 0, c.stream>>>(cur_p, nparts, opt_cap);
            B200_LAUNCH_CHECK();
            ov_tup = dev_alloc((size_t)P.src.n * tsz);
            {
                TimedScope ts(carry_p >= 0 ? "scatter_pc" : pred_p ? "filter_fused" : "scatter_p");   // pc: carried SUM column
                const OptArgs oa{opt_cap, d_ovcnt, ov_tup->ptr, carry_p >= 0 ? proj[carry_p].col : nullptr,
                                 pred_p ? P.preds : PredSet{}, (uint64_t)opt_cap * nparts};
                if (pred_p && key64) launch_scatter_pred<uint64_t, true>(P.src, bits, cur_p, tup_p->ptr, oa);
                else if (pred_p) launch_scatter_pred<uint32_t, true>(P.src, bits, cur_p, tup_p->ptr, oa);
                else if (carry_p >= 0 && key64) launch_scatter_carry_c<uint64_t, 1, true>(P.src, bits, cur_p, tup_p->ptr, oa);
                else if (carry_p >= 0) launch_scatter_opt_carry(P.src, bits, cur_p, tup_p->ptr, oa);
                else if (key64) launch_scatter_opt64(P.src, bits, cur_p, tup_p->ptr, oa);
                else launch_scatter_opt(P.src, bits, cur_p, tup_p->ptr, oa);
            }
            {
                TimedScope ts("scan");
                partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, cur_p, nparts, cap, a.slice, off_b,
                                                                      off_p, cur_b, cur_p, items, cnt_p, opt_cap);
                B200_LAUNCH_CHECK();
            }
        } else {
            {
                TimedScope ts("hist_p");
                if (key64)
                    launch_hist<uint64_t>(P.src, bits, hist_p, 4, &P.preds);
                else
                    launch_hist<uint32_t>(P.src, bits, hist_p, 4, &P.preds);
            }
            {
                TimedScope ts("scan");
                partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, hist_p, nparts, cap, a.slice, off_b,
                                                                      off_p, cur_b, cur_p, items, cnt_p, 0u);
                B200_LAUNCH_CHECK();
            }
            scatter_build(cur_b);
            {
                TimedScope ts(pred_p ? "filter_fused" : "scatter_p");
                if (pred_p) {
                    const OptArgs o{0, nullptr, nullptr, nullptr, P.preds, P.src.n};
                    if (key64) launch_scatter_pred<uint64_t, false>(P.src, bits, cur_p, tup_p->ptr, o);
                    else launch_scatter_pred<uint32_t, false>(P.src, bits, cur_p, tup_p->ptr, o);
                } else if (key64)
                    launch_scatter<uint64_t>(P.src, bits, cur_p, tup_p->ptr);
                else
                    launch_scatter<uint32_t>(P.src, bits, cur_p, tup_p->ptr);
            }
        }
        a.tup_b      = tup_b->ptr;
        a.tup_p      = tup_p->ptr;
        a.off_b      = off_b;
        a.off_p      = off_p;
        a.cnt_p      = cnt_p;
        a.item_start = items;
        a.nparts     = nparts;
    }

    if (mode == JoinOut::Sum) {
        a.nproj = nproj;
        a.need_brid = 0;
        for (int k = 0; k < nproj; ++k) {
            a.proj[k] = proj[k];
            // sides are given relative to (R, S); the kernel wants (build, probe)
            a.proj[k].side      = swapped ? 1 - proj[k].side : proj[k].side;
            a.proj[k].part_vals = (k == carry_k || k == carry_p) ? B200_PROJ_IN_RID
                                  : part_vals[k]                 ? part_vals[k]->as<uint64_t>()
                                                                 : nullptr;
            if (a.proj[k].side == 0 && (!a.proj[k].part_vals || k == carry_k)) a.need_brid = 1;
        }
        {
            TimedScope ts("join");
            launch_join(a, key64, direct, MODE_SUM, chained64);
        }
        auto read_back = [&]() {
            B200_CUDA(cudaMemcpyAsync(c.h_scratch, d_u64, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                      c.stream));
            B200_CUDA(cudaMemcpyAsync(c.h_scratch + 16, d_ovcnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
            if (!direct && (pred_b || pred_p)) {   // tuples that reached the partition buffers = rows that passed
                B200_CUDA(cudaMemcpyAsync(c.h_scratch + 17, off_b + nparts, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
                B200_CUDA(cudaMemcpyAsync(c.h_scratch + 18, off_p + nparts, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
            }
            B200_CUDA(cudaStreamSynchronize(c.stream));
        };
        read_back();
        const uint32_t n_over = opt ? *reinterpret_cast<uint32_t *>(c.h_scratch + 16) : 0u;
        uint64_t valid_b = UINT64_MAX, valid_p = UINT64_MAX;
        if (direct) {
            if (pred_p) valid_p = c.h_scratch[2];
        } else {
            if (pred_b) valid_b = *reinterpret_cast<uint32_t *>(c.h_scratch + 17);
            if (pred_p) valid_p = (uint64_t)*reinterpret_cast<uint32_t *>(c.h_scratch + 18) + n_over;
        }
        res.valid_r = swapped ? valid_p : valid_b;
        res.valid_s = swapped ? valid_b : valid_p;
        if (n_over && key64) {
            // 64-bit keys: the probe side is partitioned again, exactly, and joined from scratch
            TimedScope ts("overflow");
            B200_CUDA(cudaMemsetAsync(d_u64, 0, 16 * sizeof(unsigned long long), c.stream));
            B200_CUDA(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), c.stream));
            B200_CUDA(cudaMemsetAsync(hist_p, 0, (size_t)nparts * sizeof(uint32_t), c.stream));
            launch_hist<uint64_t>(P.src, bits, hist_p, 4, &P.preds);
            partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, hist_p, nparts, cap, a.slice, off_b, off_p,
                                                                  cur_b, cur_p, items, cnt_p, 0u);
            B200_LAUNCH_CHECK();
            tup_p = dev_alloc((size_t)P.src.n * sizeof(Tup64));
            const OptArgs oa{0, nullptr, nullptr, carry_p >= 0 ? proj[carry_p].col : nullptr, pred_p ? P.preds : PredSet{},
                             P.src.n};
            if (pred_p) launch_scatter_pred<uint64_t, false>(P.src, bits, cur_p, tup_p->ptr, oa);
            else if (carry_p >= 0) launch_scatter_carry_c<uint64_t, 1, false>(P.src, bits, cur_p, tup_p->ptr, oa);
            else launch_scatter<uint64_t>(P.src, bits, cur_p, tup_p->ptr);
            a.tup_p = tup_p->ptr;
            launch_join(a, key64, direct, MODE_SUM, chained64);
            read_back();
            if (pred_p) valid_p = *reinterpret_cast<uint32_t *>(c.h_scratch + 18);
            res.valid_r = swapped ? valid_p : valid_b;
            res.valid_s = swapped ? valid_b : valid_p;
        } else if (n_over) {
            // second pass over the overflow only: exact histogram, scatter, join against the same build partitions
            // (matches and sums keep accumulating in the same device counters)
            TimedScope ts("overflow");
            B200_CUDA(cudaMemsetAsync(hist_p, 0, (size_t)nparts * sizeof(uint32_t), c.stream));
            KeySrc ov_src{ov_tup->as<uint64_t>(), nullptr, n_over};   // low half of a packed tuple is its key
            launch_hist<uint32_t>(ov_src, bits, hist_p);
            partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, hist_p, nparts, cap, a.slice, off_b, off_p,
                                                                  cur_b, cur_p, items, cnt_p, 0u);
            B200_LAUNCH_CHECK();
            DevBufPtr ov_part = dev_alloc((size_t)n_over * sizeof(Tup32));
            launch_scatter_tuples(ov_tup->as<uint64_t>(), n_over, bits, cur_p, ov_part->ptr);
            B200_CUDA(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), c.stream));
            a.tup_p = ov_part->ptr;
            launch_join(a, key64, direct, MODE_SUM, chained64);
            read_back();
        }
        res.m = c.h_scratch[0];
        for (int k = 0; k < nproj; ++k) res.sums[k] = c.h_scratch[8 + k];
        return res;
    }

    // Pairs: count pass sizes the output exactly, write pass fills it.
    if (!direct) {
        auto read_items = [&]() {
            B200_CUDA(cudaMemcpyAsync(c.h_scratch, items + nparts, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      c.stream));
            B200_CUDA(cudaMemcpyAsync(c.h_scratch + 1, d_ovcnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
            B200_CUDA(cudaMemcpyAsync(c.h_scratch + 17, off_b + nparts, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
            B200_CUDA(cudaMemcpyAsync(c.h_scratch + 18, off_p + nparts, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
            B200_CUDA(cudaStreamSynchronize(c.stream));
            n_items = *reinterpret_cast<uint32_t *>(c.h_scratch);
        };
        read_items();
        if (opt && *reinterpret_cast<uint32_t *>(c.h_scratch + 1) != 0) {
            // the histogram-free scatter overflowed its regions (skewed keys): the pair-materialising path
            // simply partitions the probe side again, exactly
            TimedScope ts("overflow");
            const size_t tsz = key64 ? sizeof(Tup64) : sizeof(Tup32);
            B200_CUDA(cudaMemsetAsync(hist_p, 0, (size_t)nparts * sizeof(uint32_t), c.stream));
            B200_CUDA(cudaMemsetAsync(d_ovcnt, 0, sizeof(uint32_t), c.stream));
            if (key64) launch_hist<uint64_t>(P.src, bits, hist_p, 4, &P.preds);
            else launch_hist<uint32_t>(P.src, bits, hist_p, 4, &P.preds);
            partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(hist_b, hist_p, nparts, cap, a.slice, off_b, off_p,
                                                                  cur_b, cur_p, items, cnt_p, 0u);
            B200_LAUNCH_CHECK();
            tup_p = dev_alloc((size_t)P.src.n * tsz);
            const OptArgs oa{0, nullptr, nullptr, nullptr, P.preds, P.src.n};
            if (pred_p && key64) launch_scatter_pred<uint64_t, false>(P.src, bits, cur_p, tup_p->ptr, oa);
            else if (pred_p) launch_scatter_pred<uint32_t, false>(P.src, bits, cur_p, tup_p->ptr, oa);
            else if (key64) launch_scatter<uint64_t>(P.src, bits, cur_p, tup_p->ptr);
            else launch_scatter<uint32_t>(P.src, bits, cur_p, tup_p->ptr);
            a.tup_p = tup_p->ptr;
            opt     = false;
            read_items();
        }
        uint64_t valid_b = UINT64_MAX, valid_p = UINT64_MAX;
        if (pred_b) valid_b = *reinterpret_cast<uint32_t *>(c.h_scratch + 17);
        if (pred_p) valid_p = *reinterpret_cast<uint32_t *>(c.h_scratch + 18);   // (an overflowing pass was redone exactly)
        res.valid_r = swapped ? valid_p : valid_b;
        res.valid_s = swapped ? valid_b : valid_p;
    }
    DevBufPtr item_count = dev_alloc((n_items + 1) * sizeof(unsigned long long));
    a.item_count         = item_count->as<unsigned long long>();
    {
        TimedScope ts("join");
        launch_join(a, key64, direct, MODE_COUNT, chained64);
    }
    res.m = read_counter(a.total);
    if (direct && pred_p) {
        const uint64_t v = read_counter(d_valid_p);
        (swapped ? res.valid_r : res.valid_s) = v;
        a.valid_p = nullptr;   // the write pass loads the same rows again: count them once
    }
    B200_REQUIRE(res.m <= kMaxRows, "join output exceeds 2^32-1 pairs (32-bit device positions)");
    DevBufPtr out_b = dev_alloc(res.m * sizeof(uint32_t));
    DevBufPtr out_p = dev_alloc(res.m * sizeof(uint32_t));
    if (res.m) {
        a.out_b = out_b->as<uint32_t>();
        a.out_p = out_p->as<uint32_t>();
        B200_CUDA(cudaMemsetAsync(d_work, 0, sizeof(uint32_t), c.stream));
        TimedScope ts("join_write");
        launch_join(a, key64, direct, MODE_WRITE, chained64);
    }
    res.r_ids = swapped ? out_p : out_b;
    res.s_ids = swapped ? out_b : out_p;
    return res;
}

// ---------------------------------------------------------------------------
// compaction scans, gathers, checksums
// ---------------------------------------------------------------------------
constexpr int kScanNT = 256;
constexpr int kScanU  = 8;

IdList run_filter(const KeySrc &src, char cmp, int value) {
    int code;
    switch (cmp) {
        case '<': code = 0; break;
        case '>': code = 1; break;
        case '=': code = 2; break;
        default:
            // filter.c:184-186
            fprintf(stderr, "Wrong comperator in filter function\n");
            exit(2);
    }
    // `uint64_t ⋄ int`: the int is converted to uint64_t (sign-extended)
    return run_filter_u64(src, code, (uint64_t)(int64_t)value);
}

IdList run_filter_u64(const KeySrc &src, int code, uint64_t constant) {
    Context &c = ctx();
    IdList   out;
    out.ids = dev_alloc((size_t)src.n * sizeof(uint32_t));
    if (src.n == 0) return out;
    TimedScope ts("filter");
    B200_CUDA(cudaMemsetAsync(c.d_scratch, 0, sizeof(unsigned long long), c.stream));
    scan_filter_kernel<kScanNT, kScanU>
        <<<grid_for(src.n, kScanNT * kScanU, 8), kScanNT, 0, c.stream>>>(src, code, constant,
                                                                         out.ids->as<uint32_t>(), c.d_scratch);
    B200_LAUNCH_CHECK();
    out.n = read_counter(c.d_scratch);
    return out;
}

IdList run_inter_equal(const uint64_t *col_a, const uint32_t *ta, const uint64_t *col_b, const uint32_t *tb,
                       uint64_t n) {
    Context &c = ctx();
    IdList   out;
    out.ids = dev_alloc((size_t)n * sizeof(uint32_t));
    if (n == 0) return out;
    B200_CUDA(cudaMemsetAsync(c.d_scratch, 0, sizeof(unsigned long long), c.stream));
    inter_equal_kernel<kScanNT, kScanU><<<grid_for(n, kScanNT * kScanU, 8), kScanNT, 0, c.stream>>>(
        col_a, ta, col_b, tb, (uint32_t)n, out.ids->as<uint32_t>(), c.d_scratch);
    B200_LAUNCH_CHECK();
    out.n = read_counter(c.d_scratch);
    return out;
}

std::vector<DevBufPtr> run_gather(const uint32_t *pos, uint64_t m, const std::vector<const uint32_t *> &in) {
    Context               &c = ctx();
    std::vector<DevBufPtr> out;
    for (size_t i = 0; i < in.size(); ++i) out.push_back(dev_alloc(m * sizeof(uint32_t)));
    if (m == 0) return out;
    for (size_t first = 0; first < in.size(); first += kMaxGather) {
        GatherArgs g;
        g.pos   = pos;
        g.m     = (uint32_t)m;
        g.ncols = (int)std::min<size_t>(kMaxGather, in.size() - first);
        for (int k = 0; k < g.ncols; ++k) {
            g.in[k]  = in[first + k];
            g.out[k] = out[first + k]->as<uint32_t>();
        }
        gather_columns_kernel<<<grid_for(m, 256 * 4, 8), 256, 0, c.stream>>>(g);
        B200_LAUNCH_CHECK();
    }
    return out;
}

void run_checksum(uint64_t m, int nproj, const uint64_t *const *cols, const uint32_t *const *ids,
                  uint64_t *out_sums) {
    Context &c = ctx();
    for (int first = 0; first < nproj; first += kMaxProj) {
        const int k = std::min(kMaxProj, nproj - first);
        if (m == 0) {
            for (int i = 0; i < k; ++i) out_sums[first + i] = 0;
            continue;
        }
        ChecksumArgs a;
        a.m     = (uint32_t)m;
        a.nproj = k;
        for (int i = 0; i < k; ++i) {
            a.col[i] = cols[first + i];
            a.ids[i] = ids[first + i];
        }
        a.sums = c.d_scratch + 16;
        B200_CUDA(cudaMemsetAsync(a.sums, 0, kMaxProj * sizeof(unsigned long long), c.stream));
        checksum_kernel<<<grid_for(m, 256 * 4, 8), 256, 0, c.stream>>>(a);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemcpyAsync(c.h_scratch + 16, a.sums, kMaxProj * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, c.stream));
        B200_CUDA(cudaStreamSynchronize(c.stream));
        for (int i = 0; i < k; ++i) out_sums[first + i] = c.h_scratch[16 + i];
    }
}

void run_cartesian(const uint32_t *in, uint64_t n1, uint64_t n2, bool from_first, uint32_t *out) {
    const uint64_t total = n1 * n2;
    if (total == 0) return;
    cartesian_kernel<<<grid_for(total, 256 * 4, 8), 256, 0, ctx().stream>>>(in, (uint32_t)n1, (uint32_t)n2,
                                                                            from_first ? 1 : 0, out);
    B200_LAUNCH_CHECK();
}

void run_synth_column(uint64_t *d_out, uint64_t first, uint64_t n, int kind, uint64_t k, uint64_t seed) {
    if (n == 0) return;
    synth_column_kernel<<<grid_for(n, 256 * 4, 8), 256, 0, ctx().stream>>>(d_out, first, n, kind, k, seed);
    B200_LAUNCH_CHECK();
}

void widen_ids(const uint32_t *d_in, uint64_t n, uint64_t *d_out) {
    if (n == 0) return;
    widen_u32_kernel<<<grid_for(n, 256 * 4, 8), 256, 0, ctx().stream>>>(d_in, n, d_out);
    B200_LAUNCH_CHECK();
}

void narrow_ids(const uint64_t *d_in, uint64_t n, uint32_t *d_out) {
    if (n == 0) return;
    narrow_u64_kernel<<<grid_for(n, 256 * 4, 8), 256, 0, ctx().stream>>>(d_in, n, d_out);
    B200_LAUNCH_CHECK();
}

void unpack_partition(const PartitionOut &p, uint64_t n, uint64_t *d_keys, uint64_t *d_rids) {
    if (n == 0) return;
    if (p.key64)
        unpack_tuples_kernel<Tup64><<<grid_for(n, 256 * 4, 8), 256, 0, ctx().stream>>>(p.tuples->as<Tup64>(), n,
                                                                                       d_keys, d_rids);
    else
        unpack_tuples_kernel<Tup32><<<grid_for(n, 256 * 4, 8), 256, 0, ctx().stream>>>(p.tuples->as<Tup32>(), n,
                                                                                       d_keys, d_rids);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------
// staged join on caller-owned device buffers (32-bit keys): the multi-GPU plans
// of sharding.py call the phases separately so that the exchange can sit
// between them.  Every phase runs on the calling thread's stream.
// ---------------------------------------------------------------------------
void stage_hist(const uint64_t *d_keys, uint64_t n, int bits, uint32_t *d_hist) {
    Context &c = ctx();
    B200_REQUIRE(bits >= 2 && bits <= tuning().max_bits, "radix bits out of range");
    B200_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) << bits, c.stream));
    if (n == 0) return;
    KeySrc src{d_keys, nullptr, (uint32_t)n};
    TimedScope ts("hist");
    launch_hist<uint32_t>(src, bits, d_hist);
}

// phase 0: partition + broadcast; phase 1: only the local partition pass (the staging buffers stay alive
// in the calling thread's context); phase 2: only the broadcast of what phase 1 staged.
struct BuildStaging {
    DevBufPtr ctrl, tup, pay[2];
};
static thread_local BuildStaging t_build_staging;

void stage_scatter_build(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int bits, const uint32_t *d_hist_local,
                         const uint32_t *d_dst_start, int ndst, void *const *tup_dst, int npay,
                         const uint64_t *const *pay_cols, uint64_t *const *pay_dst, int phase) {
    B200_REQUIRE(ndst >= 1 && ndst <= kMaxPeers && npay >= 0 && npay <= 2, "bad destination / payload count");
    if (n == 0) return;
    Context       &c      = ctx();
    const uint32_t nparts = 1u << bits;
    BuildStaging  &st     = t_build_staging;
    if (phase == 2) {
        B200_REQUIRE(st.tup != nullptr, "broadcast phase without a staged partition pass");
    } else {
    // 1. partition the local shard into a staging buffer (local offsets from the local histogram)
    st.ctrl = dev_alloc(5 * (size_t)(nparts + 1) * sizeof(uint32_t));
    DevBufPtr &ctrl = st.ctrl;
    uint32_t *off_l = ctrl->as<uint32_t>(), *off_x = off_l + nparts + 1, *cur_l = off_x + nparts + 1,
             *cur_x = cur_l + nparts + 1, *items = cur_x + nparts + 1;
    partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_local, d_hist_local, nparts, 1u, 1u, off_l, off_x,
                                                          cur_l, cur_x, items, items, 0u);
    B200_LAUNCH_CHECK();
    st.tup = dev_alloc(n * sizeof(Tup32));
    PayArgs   pay{};
    pay.ndst     = 0;
    pay.rid_base = rid_base;
    const bool carry = npay == 1 && pay_dst == nullptr;   // the one payload column travels in the row-id slot
    if (carry) {
        pay.carry32 = 1;
        pay.col[0]  = pay_cols[0];
        pay.ids[0]  = nullptr;
    }
    for (int k = 0; k < (carry ? 0 : npay); ++k) {
        st.pay[k]  = dev_alloc(n * sizeof(uint64_t));
        pay.col[k] = pay_cols[k];
        pay.ids[k] = nullptr;
        pay.out[k] = st.pay[k]->as<uint64_t>();
    }
    KeySrc src{d_keys, nullptr, (uint32_t)n};
    {
        TimedScope ts("scatter_b");
        launch_scatter_pay<uint32_t>(src, bits, cur_l, st.tup->ptr, pay, npay);
    }
    }
    if (phase == 1) return;
    // 2. copy every partition segment to its place in the global layout of all destinations
    DevBufPtr stage_tup = st.tup;
    DevBufPtr stage_pay[2] = {st.pay[0], st.pay[1]};
    const uint32_t *off_l = st.ctrl->as<uint32_t>();
    SegCopyArgs s{};
    s.src_tup   = stage_tup->as<uint64_t>();
    s.src_off   = off_l;
    s.dst_start = d_dst_start;
    s.ndst      = ndst;
    s.npay      = pay_dst == nullptr ? 0 : npay;
    s.first     = ndst > 1 ? (int)(rid_base / (uint32_t)n) % ndst : 0;   // = this rank (equal shards)
    for (int d = 0; d < ndst; ++d) s.dst_tup[d] = static_cast<uint64_t *>(tup_dst[d]);
    for (int k = 0; k < s.npay; ++k) {
        s.src_pay[k] = stage_pay[k]->as<uint64_t>();
        for (int d = 0; d < ndst; ++d) s.dst_pay[k][d] = pay_dst[k * ndst + d];
    }
    {
        TimedScope ts("broadcast");
        segment_broadcast_kernel<<<nparts, 256, 0, c.stream>>>(s);
        B200_LAUNCH_CHECK();
    }
    st = BuildStaging{};   // released in stream order
}

// partition the local build shard straight into a caller-owned region (rank-major layout): tuples, and either
// up to two payload columns (pay_out) or one 32-bit payload carried in the row-id slot (npay == 1, pay_out == NULL)
size_t stage_scratch_bytes(int bits, int nseg) {
    const size_t nparts = (size_t)1 << bits;
    return (64 + 6 * (nparts + 1)) * sizeof(uint32_t) + 64 * sizeof(unsigned long long) +
           ((size_t)nseg + 1) * nparts * sizeof(uint32_t) + 256;
}

void stage_scatter_build_local(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int bits,
                               const uint32_t *d_hist_local, void *d_tup_out, int npay, const uint64_t *const *pay_cols,
                               uint64_t *const *pay_out, const StageScratch *scr, uint32_t *d_off_out,
                               const PredSet *skip) {
    B200_REQUIRE(npay >= 0 && npay <= 2, "bad payload count");
    if (n == 0 && !d_off_out) return;
    Context       &c      = ctx();
    const uint32_t nparts = 1u << bits;
    DevBufPtr ctrl;
    uint32_t *off_l;
    if (scr) {
        B200_REQUIRE(scr->bytes >= 5 * (size_t)(nparts + 1) * sizeof(uint32_t), "stage scratch too small");
        off_l = static_cast<uint32_t *>(scr->ptr);
    } else {
        ctrl  = dev_alloc(5 * (size_t)(nparts + 1) * sizeof(uint32_t));
        off_l = ctrl->as<uint32_t>();
    }
    if (d_off_out) off_l = d_off_out;   // the caller keeps the local offsets [nparts + 1] (exchange plan)
    uint32_t *base5 = scr ? static_cast<uint32_t *>(scr->ptr) : ctrl->as<uint32_t>();
    uint32_t *off_x = base5 + nparts + 1, *cur_l = off_x + nparts + 1,
             *cur_x = cur_l + nparts + 1, *items = cur_x + nparts + 1;
    partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_local, d_hist_local, nparts, 1u, 1u, off_l, off_x,
                                                          cur_l, cur_x, items, items, 0u);
    B200_LAUNCH_CHECK();
    if (n == 0) return;
    PayArgs pay{};
    pay.rid_base = rid_base;
    const bool carry = npay == 1 && pay_out == nullptr;
    if (carry) {
        pay.carry32 = 1;
        pay.col[0]  = pay_cols[0];
    }
    for (int k = 0; k < (carry ? 0 : npay); ++k) {
        pay.col[k] = pay_cols[k];
        pay.out[k] = pay_out[k];
    }
    KeySrc src{d_keys, nullptr, (uint32_t)n};
    TimedScope ts("scatter_b");
    if (skip && skip->hot_keys) {
        // rows the caller's histogram did not count (hot keys, joined on the spot) are skipped here as well
        B200_REQUIRE(carry || npay == 0, "the skipping scatter carries at most one 32-bit payload");
        const OptArgs o{0, nullptr, nullptr, carry ? pay_cols[0] : nullptr, *skip};
        if (carry) launch_scatter_pred_carry(src, bits, cur_l, d_tup_out, o);
        else launch_scatter_pred<uint32_t, false>(src, bits, cur_l, d_tup_out, o);
        return;
    }
    // large shards with a carried payload (the probe side of the exchange plan) take the tuned scatter instance
    if (carry && n >= (1u << 18))
        launch_scatter_carry_tuned(src, bits, cur_l, d_tup_out, pay_cols[0]);
    else
        launch_scatter_pay<uint32_t>(src, bits, cur_l, d_tup_out, pay, npay);
}

// Two-pass variant of stage_scatter_build_local for MANY partitions (the exchange plan at 2^11 .. 2^12): a 16 K-tuple
// tile spread over 4096 partitions leaves runs of four tuples — one 32-byte sector each, and 4096 run reservations per
// tile.  Pass 1 partitions the column(s) by the LOW half of the partition bits into d_tmp (packed tuples, the carried
// value in the slot; rows with a hot key skipped), pass 2 partitions d_tmp by all the bits: its tiles hold tuples of
// one coarse partition, i.e. of 2^(bits - cbits) fine ones, so the runs are long again.  8 more bytes per row read and
// written, both passes at the rate of the well-behaved scatter.
void stage_scatter_two_pass(const uint64_t *d_keys, uint64_t n, int bits, const uint32_t *d_hist_local, void *d_tup_out,
                            const uint64_t *carry_col, void *d_tmp, const StageScratch *scr, uint32_t *d_off_out,
                            const PredSet *skip) {
    B200_REQUIRE(scr && carry_col && d_tmp && d_off_out, "two-pass scatter: scratch, carried column, buffer and offsets");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(d_keys) & 15) == 0 && (reinterpret_cast<uintptr_t>(carry_col) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(d_tmp) & 15) == 0, "two-pass scatter: 16-byte aligned inputs");
    Context       &c      = ctx();
    const uint32_t nparts = 1u << bits;
    const int      cbits  = bits / 2;
    const uint32_t nc     = 1u << cbits;
    B200_REQUIRE(scr->bytes >= (5 * (size_t)(nparts + 1) + 6 * (size_t)(nc + 1)) * sizeof(uint32_t), "stage scratch too small");
    uint32_t *base5 = static_cast<uint32_t *>(scr->ptr);
    uint32_t *off_x = base5 + nparts + 1, *cur_l = off_x + nparts + 1, *cur_x = cur_l + nparts + 1, *items = cur_x + nparts + 1;
    uint32_t *coarse = items + nparts + 1, *c_off = coarse + nc + 1, *c_offx = c_off + nc + 1, *c_cur = c_offx + nc + 1,
             *c_curx = c_cur + nc + 1, *c_items = c_curx + nc + 1;
    // fine plan: offsets the caller keeps, cursors of pass 2
    partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_local, d_hist_local, nparts, 1u, 1u, d_off_out, off_x, cur_l,
                                                          cur_x, items, items, 0u);
    B200_LAUNCH_CHECK();
    if (n == 0) return;
    coarse_hist_kernel<<<1, 1024, 0, c.stream>>>(d_hist_local, nparts, (uint32_t)cbits, coarse);
    B200_LAUNCH_CHECK();
    partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(coarse, coarse, nc, 1u, 1u, c_off, c_offx, c_cur, c_curx, c_items,
                                                          c_items, 0u);
    B200_LAUNCH_CHECK();
    TimedScope ts("scatter_b");
    KeySrc     src{d_keys, nullptr, (uint32_t)n};
    if (skip && skip->hot_keys) {
        const OptArgs o{0, nullptr, nullptr, carry_col, *skip};
        launch_scatter_pred_carry(src, cbits, c_cur, d_tmp, o);
    } else {
        launch_scatter_carry_tuned(src, cbits, c_cur, d_tmp, carry_col);
    }
    {
        constexpr int NT   = PartCfg<uint32_t, 1>::NT;
        constexpr int U    = PartCfg<uint32_t, 1>::U;
        const size_t  smem = (size_t)NT * U * sizeof(Tup32) + 3 * (size_t)nparts * sizeof(uint32_t);
        auto          k    = radix_scatter_kernel<NT, U, 1, uint32_t, false, false, false, true>;
        allow_smem(k, smem);
        KeySrc  tsrc{static_cast<const uint64_t *>(d_tmp), nullptr, (uint32_t)n};
        OptArgs o{0, nullptr, nullptr, nullptr, PredSet{}};
        o.n_dev = d_off_out + nparts;   // rows that survived pass 1 = the total of the fine histogram
        k<<<grid_for(n, NT * U, 1), NT, smem, c.stream>>>(tsrc, (uint32_t)bits, cur_l, static_cast<Tup32 *>(d_tup_out), o);
        B200_LAUNCH_CHECK();
    }
}

void stage_scatter_probe(const uint64_t *d_keys, uint64_t n, int bits, uint32_t *d_cursor, void *d_tup_out) {
    if (n == 0) return;
    KeySrc src{d_keys, nullptr, (uint32_t)n};
    TimedScope ts("scatter_p");
    launch_scatter<uint32_t>(src, bits, d_cursor, d_tup_out);
}

void stage_build_cursors(const uint32_t *d_hist_all, int world, int rank, int bits, uint32_t *d_total,
                         uint32_t *d_my_start) {
    build_cursors_kernel<1024><<<1, 1024, 0, ctx().stream>>>(d_hist_all, (uint32_t)world, (uint32_t)rank, 1u << bits,
                                                              d_total, d_my_start);
    B200_LAUNCH_CHECK();
}

// Radix-sharded exchange (SURVEY §8e all-to-all): layout of the exchange from the all-gathered histograms
void stage_exchange_cursors(const uint32_t *d_hist_all, int world, int rank, int bits, uint32_t cap,
                            uint32_t *d_src_off, uint32_t *d_dst_start, uint32_t *d_own_total, uint32_t *d_need) {
    B200_REQUIRE(world >= 1 && world <= kMaxPeers && (1u << bits) >= (uint32_t)world, "bad world for the exchange");
    exchange_cursors_kernel<1024><<<1, 1024, 0, ctx().stream>>>(d_hist_all, (uint32_t)world, (uint32_t)rank,
                                                                (uint32_t)bits, cap, d_src_off, d_dst_start,
                                                                d_own_total, d_need);
    B200_LAUNCH_CHECK();
}

// ... and the exchange: every staged tuple (and payload value) goes to the owner of its partition
void stage_exchange_segments(const void *d_src_tup, int npay, const uint64_t *const *src_pay, uint64_t n, int bits,
                             int world, const uint32_t *d_src_off, const uint32_t *d_dst_start, uint32_t cap,
                             int rewrite_rid, void *const *tup_dst, uint64_t *const *pay_dst) {
    B200_REQUIRE(world >= 1 && world <= kMaxPeers && npay >= 0 && npay <= 2, "bad destination / payload count");
    if (n == 0) return;
    ExchangeArgs x{};
    x.src_tup     = static_cast<const uint64_t *>(d_src_tup);
    x.src_off     = d_src_off;
    x.dst_start   = d_dst_start;
    x.n           = (uint32_t)n;
    x.radix_bits  = (uint32_t)bits;
    x.world       = (uint32_t)world;
    x.cap         = cap;
    x.npay        = npay;
    x.rewrite_rid = rewrite_rid;
    for (int d = 0; d < world; ++d) x.dst_tup[d] = static_cast<uint64_t *>(tup_dst[d]);
    for (int k = 0; k < npay; ++k) {
        x.src_pay[k] = src_pay[k];
        for (int d = 0; d < world; ++d) x.dst_pay[k][d] = pay_dst[k * world + d];
    }
    TimedScope ts("exchange");
    const uint32_t blocks = (uint32_t)((n + 1023) / 1024);
    segment_exchange_kernel<<<blocks, 256, 0, ctx().stream>>>(x);
    B200_LAUNCH_CHECK();
}

uint32_t opt_region_cap(uint64_t n_probe, int bits) {
    const uint64_t nparts = 1ull << bits;
    const uint64_t mean   = (n_probe + nparts - 1) / nparts;
    return (uint32_t)((mean + mean / 32 + 6 * (uint64_t)sqrt((double)mean) + 64 + 3) & ~3ull);
}

// histogram-free probe-side scatter into caller-owned buffers: d_tup_out holds 2^bits regions of opt_cap tuples,
// d_ov (n tuples) and d_ovcnt (one u32) receive what does not fit
// carry_col != nullptr: the row-id slot of every probe tuple carries (uint32)carry_col[row] (a probe-side SUM column
// with 32-bit values; pass that projection to the join with part_vals = B200_PROJ_IN_RID)
void stage_scatter_probe_opt(const uint64_t *d_keys, uint64_t n, int bits, uint32_t opt_cap, uint32_t *d_cursor,
                             void *d_tup_out, void *d_ov, uint32_t *d_ovcnt, const uint64_t *carry_col) {
    Context       &c      = ctx();
    const uint32_t nparts = 1u << bits;
    B200_REQUIRE(n <= (1u << 30), "histogram-free scatter is limited to 2^30 probe rows");
    init_opt_cursors_kernel<<<(nparts + 255) / 256, 256, 0, c.stream>>>(d_cursor, nparts, opt_cap);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemsetAsync(d_ovcnt, 0, sizeof(uint32_t), c.stream));
    if (n == 0) return;
    KeySrc src{d_keys, nullptr, (uint32_t)n};
    TimedScope ts(carry_col ? "scatter_pc" : "scatter_p");
    if (carry_col)
        launch_scatter_opt_carry(src, bits, d_cursor, d_tup_out, OptArgs{opt_cap, d_ovcnt, d_ov, carry_col, PredSet{}});
    else
        launch_scatter_opt(src, bits, d_cursor, d_tup_out, OptArgs{opt_cap, d_ovcnt, d_ov, nullptr, PredSet{}});
}

// d_hist_p: probe-side histogram (opt_cap == 0) or the cursor array stage_scatter_probe_opt left behind
// d_result != nullptr: asynchronous — matches, the nproj sums and the overflow count are copied to
// d_result[0], d_result[1..nproj], d_result[nproj + 1] (u64 each) on the stream and nothing is read back;
// the caller runs stage_join_overflow when it later finds a non-zero overflow count.
// nseg > 0: rank-major build layout — d_hist_b is then hist_all[nseg][2^bits] and region r of d_tup_b (and of the
// part_vals arrays) starts at r * seg_rows and holds rank r's shard in partition order.
JoinResult stage_join_sum(const void *d_tup_b, const uint32_t *d_hist_b, const void *d_tup_p, const uint32_t *d_hist_p,
                          int bits, int nproj, const ProjDesc *proj, uint32_t opt_cap, const void *d_ov,
                          const uint32_t *d_ovcnt, unsigned long long *d_result, int nseg, uint32_t seg_rows,
                          const StageScratch *scr, const JoinWait *wait, uint32_t seg_head) {
    Context   &c = ctx();
    Tuning    &t = tuning();
    JoinResult res;
    B200_REQUIRE(nproj >= 0 && nproj <= kMaxProj, "too many fused projections");
    const uint32_t nparts = 1u << bits;
    const uint32_t cap    = t.cap32;
    JoinArgs a;
    memset(&a, 0, sizeof(a));
    a.cap        = cap;
    a.slots_log2 = tag_slots_log2_for(cap);
    a.radix_bits = (uint32_t)bits;
    a.slice      = t.slice;
    B200_REQUIRE(bits + (int)a.slots_log2 >= 17, "radix bits too small for the tag table");
    const size_t ctrl_bytes = (64 + 6 * (size_t)(nparts + 1)) * sizeof(uint32_t) + 64 * sizeof(unsigned long long);
    DevBufPtr    ctrl;
    void        *ctrl_ptr;
    if (scr) {   // caller-owned control memory: nothing is allocated here (the step can be captured in a CUDA graph)
        B200_REQUIRE(scr->bytes >= stage_scratch_bytes(bits, nseg), "stage scratch too small");
        B200_REQUIRE(d_result != nullptr, "a stage with caller-owned scratch is asynchronous");
        ctrl_ptr = scr->ptr;
    } else {
        ctrl     = dev_alloc(ctrl_bytes);
        ctrl_ptr = ctrl->ptr;
    }
    B200_CUDA(cudaMemsetAsync(ctrl_ptr, 0, ctrl_bytes, c.stream));
    unsigned long long *d_u64 = static_cast<unsigned long long *>(ctrl_ptr);
    uint32_t           *d_u32 = reinterpret_cast<uint32_t *>(d_u64 + 64);
    uint32_t *off_b = d_u32 + 64, *off_p = off_b + nparts + 1, *cur_b = off_p + nparts + 1,
             *cur_p = cur_b + nparts + 1, *items = cur_p + nparts + 1, *cnt_p = items + nparts + 1;
    a.work_counter = d_u32;
    a.total        = d_u64;
    a.out_cursor   = d_u64 + 1;
    a.sums         = d_u64 + 8;
    DevBufPtr seg;
    if (nseg > 0) {
        B200_REQUIRE(nseg <= kMaxPeers, "at most 8 build segments");
        uint32_t *seg_off;
        if (scr) {
            seg_off = reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(ctrl_ptr) + ((ctrl_bytes + 255) & ~(size_t)255));
        } else {
            seg     = dev_alloc(((size_t)nseg + 1) * nparts * sizeof(uint32_t));
            seg_off = seg->as<uint32_t>();
        }
        uint32_t *total = seg_off + (size_t)nseg * nparts;
        segment_offsets_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_b, (uint32_t)nseg, nparts, seg_rows, seg_off,
                                                                total, seg_head);
        B200_LAUNCH_CHECK();
        a.nseg    = nseg;
        a.seg_off = seg_off;
        a.seg_cnt = d_hist_b;
        a.seg_rows = seg_rows;
        d_hist_b  = total;   // the plan works on the global histogram (virtual concatenation of the runs)
        if (wait) {
            a.wait_flags = wait->flags;
            a.wait_epoch = wait->epoch;
            a.chunk_rows = wait->chunk_rows;
            a.wait_error = wait->error;
        }
    }
    {
        TimedScope ts("scan");
        partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_b, d_hist_p, nparts, cap, a.slice, off_b, off_p,
                                                              cur_b, cur_p, items, cnt_p, opt_cap);
        B200_LAUNCH_CHECK();
    }
    a.tup_b      = d_tup_b;
    a.tup_p      = d_tup_p;
    a.off_b      = off_b;
    a.off_p      = off_p;
    a.cnt_p      = cnt_p;
    a.item_start = items;
    a.nparts     = nparts;
    a.nproj      = nproj;
    a.need_brid  = 0;
    for (int k = 0; k < nproj; ++k) {
        a.proj[k] = proj[k];
        if (a.proj[k].side == 0 && (!a.proj[k].part_vals || a.proj[k].part_vals == B200_PROJ_IN_RID)) a.need_brid = 1;
    }
    {
        TimedScope ts("join");
        launch_join(a, false, false, MODE_SUM);
    }
    if (d_result) {
        B200_CUDA(cudaMemcpyAsync(d_result, d_u64, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream));
        if (nproj)
            B200_CUDA(cudaMemcpyAsync(d_result + 1, d_u64 + 8, (size_t)nproj * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToDevice, c.stream));
        B200_CUDA(cudaMemsetAsync(d_result + nproj + 1, 0, sizeof(unsigned long long), c.stream));
        if (opt_cap)
            B200_CUDA(cudaMemcpyAsync(d_result + nproj + 1, d_ovcnt, sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                      c.stream));
        return res;
    }
    auto read_back = [&]() {
        B200_CUDA(cudaMemcpyAsync(c.h_scratch, d_u64, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  c.stream));
        if (opt_cap)
            B200_CUDA(cudaMemcpyAsync(c.h_scratch + 16, d_ovcnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
        B200_CUDA(cudaStreamSynchronize(c.stream));
    };
    read_back();
    const uint32_t n_over = opt_cap ? *reinterpret_cast<uint32_t *>(c.h_scratch + 16) : 0u;
    if (n_over) {
        // exact second pass over the overflow (see run_join)
        TimedScope ts("overflow");
        DevBufPtr hist_ov = dev_alloc((size_t)nparts * sizeof(uint32_t));
        B200_CUDA(cudaMemsetAsync(hist_ov->ptr, 0, hist_ov->bytes, c.stream));
        KeySrc ov_src{static_cast<const uint64_t *>(d_ov), nullptr, n_over};
        launch_hist<uint32_t>(ov_src, bits, hist_ov->as<uint32_t>());
        partition_plan_kernel<1024><<<1, 1024, 0, c.stream>>>(d_hist_b, hist_ov->as<uint32_t>(), nparts, cap, a.slice,
                                                              off_b, off_p, cur_b, cur_p, items, cnt_p, 0u);
        B200_LAUNCH_CHECK();
        DevBufPtr ov_part = dev_alloc((size_t)n_over * sizeof(Tup32));
        launch_scatter_tuples(static_cast<const uint64_t *>(d_ov), n_over, bits, cur_p, ov_part->ptr);
        B200_CUDA(cudaMemsetAsync(d_u32, 0, sizeof(uint32_t), c.stream));
        a.tup_p = ov_part->ptr;
        launch_join(a, false, false, MODE_SUM);
        read_back();
    }
    res.m = c.h_scratch[0];
    for (int k = 0; k < nproj; ++k) res.sums[k] = c.h_scratch[8 + k];
    return res;
}

}  // namespace b200
