// common.cuh — runtime plumbing shared by every translation unit of
// libb200join.so: error handling, the per-thread stream context, stream-ordered
// device buffers, the launch counter and optional CUDA-event kernel timing.
//
// Error behaviour mirrors the reference's "print and exit(2)" convention
// (rhjoin.c:285-286, filter.c:185-186, query.c:424-425): a CUDA failure inside
// an operator is fatal.  There is no CPU fallback anywhere in this library.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace b200 {

[[noreturn]] void fatal(const char *file, int line, const char *what, const char *detail);
void set_last_error(const std::string &msg);

// B200_SLOWLOG=<ms>: every checked CUDA call and kernel launch is a checkpoint; when more than <ms> of host time
// passed on the calling thread since its previous checkpoint, the two sites and the gap are printed to stderr — finds
// the host-side stall (an allocation, a synchronisation, a launch that blocks) nsys would show, without nsys.
extern int g_slowlog_ms;
void       slow_checkpoint(const char *file, int line, const char *what);

#define B200_CUDA(expr)                                                               \
    do {                                                                              \
        cudaError_t err__ = (expr);                                                   \
        if (err__ != cudaSuccess)                                                     \
            ::b200::fatal(__FILE__, __LINE__, #expr, cudaGetErrorString(err__));      \
        if (::b200::g_slowlog_ms) ::b200::slow_checkpoint(__FILE__, __LINE__, #expr); \
    } while (0)

#define B200_REQUIRE(cond, msg)                                                       \
    do {                                                                              \
        if (!(cond)) ::b200::fatal(__FILE__, __LINE__, #cond, msg);                   \
    } while (0)

// Checked build (make checked -> lib/libb200join_checked.so, -DB200_CHECKED): device-side bounds checks on the
// hand-rolled shared-memory structures of the hot kernels (stage buffers, tag tables, warp queues) and on the
// positions they write to.  compute-sanitizer is closed on the GPU pool this was developed on
// (profiles/r2_sanitizer_unavailable.txt), so tests/test_checked_build_gpu.py runs a subset of the GPU tests over
// this library instead; the shipped library compiles the checks out.
#ifdef B200_CHECKED
#include <cassert>
#define B200_DCHECK(cond) assert(cond)
#else
#define B200_DCHECK(cond) ((void)0)
#endif

// Row ids and positions travel as 32-bit integers on the device; every length
// that becomes a row id is checked against this bound at the API boundary.
constexpr uint64_t kMaxRows = 0xFFFFFFFFull;

struct KernelTimer {
    cudaEvent_t start = nullptr, stop = nullptr;
    bool        used  = false;
    double      earlier_ms = 0;   // scopes of this name that ended before the current one (since profiling was enabled)
    int         scopes     = 0;
};

// One context per host thread: the reference fans a join out over pthreads
// (scheduler.c); here every calling thread owns a CUDA stream and all of an
// operator's kernels are enqueued on it in order.
struct Context {
    cudaStream_t stream       = nullptr;   // the stream operators run on: own_stream, or one adopted by b200_set_stream
    cudaStream_t own_stream   = nullptr;   // created with the context, never destroyed while the process lives
    int          device       = 0;
    // pinned scratch for counters read back after a stream synchronise
    unsigned long long *h_scratch = nullptr;   // 64 x u64, pinned
    unsigned long long *d_scratch = nullptr;   // 64 x u64, device
    std::map<std::string, KernelTimer> timers;
    // large transient buffers (partition buffers, row-id lists of big relations) are kept here when they are
    // released and handed out again to the next request of a similar size on the same stream: the stream-ordered
    // pool splits its big free blocks for small requests, so a query that needs two 1.6 GB buffers per join paid a
    // fresh physical allocation (milliseconds) every time
    struct BigBlock {
        void        *ptr;
        size_t       bytes;
        cudaStream_t stream;
    };
    std::vector<BigBlock> big_cache;
    size_t                big_cached = 0;
    ~Context();
};

Context &ctx();                 // thread-local, created on first use (pooled across threads)
void     set_thread_device(int device);   // device the calling thread's context lives on (-1 = process default)
int      thread_device();
void     ensure_init();         // device + pool, idempotent
int      sm_count();
bool     profiling_enabled();
extern std::atomic<uint64_t> g_launches;

// RAII timing scope: records events on the context stream when profiling is on.
// ... and, with B200_NVTX=1, opens an NVTX range of the same name (header-only nvtx3: a timeline tool attached to the
// process shows hist_b / scatter_pc / join / exchange ... per operator; no cost when no tool is attached).
struct TimedScope {
    KernelTimer *t = nullptr;
    bool         nvtx = false;
    explicit TimedScope(const char *name);
    ~TimedScope();
};

extern thread_local uint64_t t_launches;   // the calling thread's share of g_launches
inline void count_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ++t_launches;
}

#define B200_LAUNCH_CHECK()                                                           \
    do {                                                                              \
        ::b200::count_launch();                                                       \
        B200_CUDA(cudaGetLastError());                                                \
    } while (0)

// Stream-ordered device buffer (cudaMallocAsync on the owning thread's stream).
struct DevBuf {
    void        *ptr   = nullptr;
    size_t       bytes = 0;      // what was asked for
    size_t       capacity = 0;   // what the block holds (a recycled big block may be larger)
    cudaStream_t stream = nullptr;
    DevBuf(size_t nbytes, cudaStream_t s);
    ~DevBuf();
    DevBuf(const DevBuf &)            = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <typename T> T *as() const { return static_cast<T *>(ptr); }
};
using DevBufPtr = std::shared_ptr<DevBuf>;

inline DevBufPtr dev_alloc(size_t nbytes) {
    return std::make_shared<DevBuf>(nbytes, ctx().stream);
}

}  // namespace b200
