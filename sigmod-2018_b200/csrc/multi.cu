// multi.cu — the join over the GPUs of one box, behind the C ABI (include/b200_join.h, "multi-GPU plans").
//
// The reference is one process with pthreads (scheduler.c); its join shards by bucket (rhjoin.c:42-57: one
// JoinJob per bucket pair, nothing shared between buckets).  One rank drives one GPU — a process per GPU
// (bench.py under torchrun: peers' memory through CUDA IPC) or a host thread per GPU inside one process
// (host/b200_engine -g N, b200_join_sum_multi: cudaDeviceEnablePeerAccess).  No torch, no NCCL: ranks
// synchronise through epoch flags in each other's memory (multi_kernels.cuh), data moves with copy engines or
// with stores from the exchange kernel over NVLink, and the k + 1 result words are summed by every rank from
// slots its peers wrote.  A step enqueues device work only (no allocation, no host round trip), so it can be
// captured in a CUDA graph; b200_multi_finish is the one synchronisation.
//
// Plans:
//   broadcast  small build side (config 2): every rank partitions its build shard ONCE into region `rank` of
//              its build buffer (rank-major layout) and its copy engines push that region, chunk by chunk with
//              a flag behind each chunk, into the same region of every peer; the probe shard is partitioned
//              locally meanwhile and never moves; the join reads a partition as `world` runs and waits, per
//              partition, only for the chunks that hold its runs — it overlaps the tail of the broadcast.
//   exchange   radix-sharded all-to-all (config 4): both shards are partitioned locally (the probe shard in
//              chunks), every partition is stored into the receive buffer of its owner by the exchange kernel
//              (the exchange of chunk c runs under the partition pass of chunk c + 1), owners are contiguous
//              partition ranges cut on the global histogram so that skew does not overload one GPU, and every
//              owner joins what it received.
#include "../../include/b200_join.h"
#include "engine.cuh"
#include "multi_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace b200 {

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {   // carve one allocation into aligned pieces
    unsigned char *base;
    size_t         off = 0;
    explicit Carver(void *p) : base(static_cast<unsigned char *>(p)) {}
    template <typename T> T *take(size_t count) {
        off      = align_up(off, 256);
        T *p     = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

// run the stages of engine.cu on another stream of the calling thread's context
struct StreamSwap {
    Context     &c;
    cudaStream_t saved;
    StreamSwap(Context &ctx_, cudaStream_t s) : c(ctx_), saved(ctx_.stream) { c.stream = s; }
    ~StreamSwap() { c.stream = saved; }
};

}  // namespace

struct MultiPlan {
    b200_multi_config cfg{};
    int      rank = 0, world = 1, device = 0;
    int      bits = 0, K = 1;
    uint32_t P = 1, seg_rows = 0, seg_head = 0, chunk_rows = 0, opt_cap = 0, cap_b = 0, cap_p = 0;   // seg_rows: region stride
    int      nproj = 0;
    // shared region (identical layout on every rank)
    unsigned char *shared = nullptr;
    size_t         shared_bytes = 0, off_hist_b = 0, off_hist_p = 0, off_build = 0, off_recv_p = 0;
    unsigned char *peer[kMaxPeers] = {nullptr};
    bool           peer_ipc[kMaxPeers] = {false};
    // local device memory (one allocation)
    unsigned char *local = nullptr;
    uint32_t      *d_epoch = nullptr, *d_error = nullptr, *cur_p = nullptr, *ovcnt = nullptr, *pull_work = nullptr;
    int            pull = 0, pull_sms = 12;   // broadcast plan: fetch the peers' regions with a kernel instead of pushing
    unsigned long long *d_result = nullptr, *d_final = nullptr;
    void          *tup_p = nullptr, *ov_p = nullptr, *stage_b = nullptr, *stage_p = nullptr, *stage_tmp = nullptr;
    int            two_pass_bits = 99;   // B200_TWO_PASS_BITS: probe chunks take two partition passes from this many radix
                                         // bits up (off by default: measured no faster than one pass)
    uint32_t      *src_off_b = nullptr, *src_off_p = nullptr, *dst_start_b = nullptr, *dst_start_p = nullptr;
    uint32_t      *own_total = nullptr, *total = nullptr, *cut = nullptr, *need = nullptr;
    // hot keys (exchange plan): sampling table, the agreed table, build-side aggregates, the hot rows' result
    int            hot = 0;
    uint32_t      *hs_keys = nullptr, *hs_cnt = nullptr, *hot_keys = nullptr, *hot_n = nullptr;
    unsigned long long *hot_agg = nullptr, *hot_acc = nullptr;
    size_t         off_cand = 0, off_agg = 0;
    StageScratch   scr_a, scr_b;
    unsigned long long *h_final = nullptr;   // pinned
    cudaStream_t   copy_stream[kMaxPeers] = {nullptr}, xstream = nullptr, gstream = nullptr;
    cudaEvent_t    ev_build = nullptr, ev_copy[kMaxPeers] = {nullptr}, ev_chunk[kMaxChunks] = {nullptr}, ev_x = nullptr,
                   ev_hist = nullptr;
    // inputs of the step in flight (the overflow pass of finish() needs them)
    const uint64_t *in_bk = nullptr, *in_bp = nullptr, *in_pk = nullptr, *in_pp = nullptr;
    bool            pending = false;
    // CUDA graph of the step (on by default with several GPUs, B200_MULTI_GRAPH=0/1): valid while the input pointers
    // stay the same
    cudaGraphExec_t graph = nullptr;
    const uint64_t *g_in[4] = {nullptr, nullptr, nullptr, nullptr};
    int             use_graph = 0;
    uint64_t        graph_kernels = 0;   // kernel launches one replay of the graph stands for (b200_kernel_launches)
    bool            pull_smem_set = false;

    SharedHeader *hdr(int r) const { return reinterpret_cast<SharedHeader *>(peer[r]); }
    uint32_t     *hist_b(int r) const { return reinterpret_cast<uint32_t *>(peer[r] + off_hist_b); }
    uint32_t     *hist_p(int r) const { return reinterpret_cast<uint32_t *>(peer[r] + off_hist_p); }
    unsigned char *build(int r) const { return peer[r] + off_build; }
    unsigned char *recv_p(int r) const { return peer[r] + off_recv_p; }
    uint32_t      *cand(int r) const { return reinterpret_cast<uint32_t *>(peer[r] + off_cand); }   // [world][kHotCandWords]
    unsigned long long *agg(int r) const { return reinterpret_cast<unsigned long long *>(peer[r] + off_agg); }
    PeerPtrs peers() const {
        PeerPtrs p{};
        for (int r = 0; r < world; ++r) p.hdr[r] = hdr(r);
        return p;
    }
};

// broadcast: header | hist_b[world][P] | build[world][seg_rows]
// exchange:  header | hist_b[world][P] | hist_p[world * K][P] | cand[world][1 + 2048] | agg[world][8192][2] |
//            recv_b[cap_b] | recv_p[cap_p]   (build() = recv_b)
static void layout_shared(MultiPlan &m) {
    m.off_hist_b = align_up(sizeof(SharedHeader), 256);
    if (m.cfg.plan == B200_PLAN_BROADCAST) {
        m.off_hist_p   = m.off_hist_b;
        m.off_build    = align_up(m.off_hist_b + (size_t)m.world * m.P * 4, 256);
        m.off_recv_p   = m.off_build;
        m.shared_bytes = align_up(m.off_build + ((size_t)m.world * m.seg_rows + 16) * 8, 256);
    } else {
        m.off_hist_p   = align_up(m.off_hist_b + (size_t)m.world * m.P * 4, 256);
        m.off_cand     = align_up(m.off_hist_p + (size_t)m.world * m.K * m.P * 4, 256);
        m.off_agg      = align_up(m.off_cand + (size_t)m.world * kHotCandWords * 4, 256);
        m.off_build    = align_up(m.off_agg + (size_t)m.world * kHotSlots * 16, 256);
        m.off_recv_p   = align_up(m.off_build + ((size_t)m.cap_b + 16) * 8, 256);
        m.shared_bytes = align_up(m.off_recv_p + ((size_t)m.cap_p + 16) * 8, 256);
    }
}

static void layout_local(MultiPlan &m, void *base, size_t *bytes) {
    Carver c(base);
    const size_t P = m.P, np = m.cfg.n_probe_local, nb = m.cfg.n_build_local;
    m.d_epoch  = c.take<uint32_t>(4);
    m.d_error  = c.take<uint32_t>(4);
    m.ovcnt    = c.take<uint32_t>(4);
    m.pull_work = c.take<uint32_t>(1 + kMaxChunks * kMaxPeers);
    m.d_result = c.take<unsigned long long>(8);
    m.d_final  = c.take<unsigned long long>(8);
    m.cur_p    = c.take<uint32_t>(P + 1);
    const size_t sb = stage_scratch_bytes(m.bits, m.world);
    m.scr_a = StageScratch{c.take<unsigned char>(sb), sb};
    m.scr_b = StageScratch{c.take<unsigned char>(sb), sb};
    if (m.cfg.plan == B200_PLAN_BROADCAST) {
        m.tup_p = c.take<uint64_t>(std::max<size_t>(m.opt_cap ? (size_t)m.opt_cap * P : np, 1));
        m.ov_p  = c.take<uint64_t>(std::max<size_t>(m.opt_cap ? np : 1, 1));
    } else {
        m.stage_b     = c.take<uint64_t>(std::max<size_t>(nb, 1));
        m.stage_p     = c.take<uint64_t>(std::max<size_t>(np, 1));
        m.stage_tmp   = c.take<uint64_t>((np + (size_t)m.K - 1) / (size_t)m.K + 16);   // one chunk, between the two passes
        m.src_off_b   = c.take<uint32_t>(P + 1);
        m.src_off_p   = c.take<uint32_t>((size_t)m.K * (P + 1));
        m.dst_start_b = c.take<uint32_t>(P);
        m.dst_start_p = c.take<uint32_t>((size_t)m.K * P);
        m.own_total   = c.take<uint32_t>(2 * P);
        m.total       = c.take<uint32_t>(2 * P);
        m.cut         = c.take<uint32_t>(kMaxPeers + 1);
        m.need        = c.take<uint32_t>(4);
        m.hs_keys     = c.take<uint32_t>(kHotSample);
        m.hs_cnt      = c.take<uint32_t>(kHotSample);
        m.hot_keys    = c.take<uint32_t>(kHotSlots);
        m.hot_n       = c.take<uint32_t>(4);
        m.hot_agg     = c.take<unsigned long long>(2 * kHotSlots);
        m.hot_acc     = c.take<unsigned long long>(4);
    }
    *bytes = align_up(c.off, 256);
}

// first row of chunk c of n rows (chunk K starts at n); even, so that 16-byte aligned shards give 16-byte aligned chunks
static uint32_t chunk_first(const MultiPlan &m, uint64_t n, int c) {
    return c >= m.K ? (uint32_t)n : (uint32_t)((n * (uint64_t)c / (uint64_t)m.K) & ~1ull);
}

// ---------------------------------------------------------------------------
// steps
// ---------------------------------------------------------------------------
static void wait_signal_row(MultiPlan &m, cudaStream_t stream, int sig) {
    wait_peers_kernel<<<1, 32, 0, stream>>>(m.hdr(m.rank)->sig[sig], m.world, m.d_epoch, m.d_error);
    B200_LAUNCH_CHECK();
}

static void enqueue_broadcast(MultiPlan &m, int phases) {
    Context     &c    = ctx();
    cudaStream_t main = c.stream;
    const int    rank = m.rank, world = m.world;
    const uint32_t P  = m.P;
    const uint64_t nb = m.cfg.n_build_local, np = m.cfg.n_probe_local;
    const size_t   region_bytes = (size_t)m.seg_rows * 8;
    if (phases & 1) {
        bump_epoch_kernel<<<1, 1, 0, main>>>(m.d_epoch);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemsetAsync(m.d_error, 0, 16, main));
        B200_CUDA(cudaMemsetAsync(m.d_result, 0, 64, main));
        // ---- build shard: histogram into the head of my region, ONE local partition pass into the rest of it ----
        unsigned char *my_region = m.build(rank) + (size_t)rank * region_bytes;
        uint32_t      *my_hist   = reinterpret_cast<uint32_t *>(my_region);
        unsigned char *my_tuples = my_region + (size_t)m.seg_head * 8;
        stage_hist(m.in_bk, nb, m.bits, my_hist);
        const uint64_t *pay_cols[1] = {m.in_bp};
        stage_scatter_build_local(m.in_bk, nb, 0, m.bits, my_hist, my_tuples, m.cfg.has_build_sum ? 1 : 0, pay_cols, nullptr,
                                  &m.scr_a);
        B200_CUDA(cudaEventRecord(m.ev_build, main));
        if (m.pull) {
            // ---- pull variant: announce my region, then fetch the peers' regions with a small persistent kernel that
            //      runs beside the probe-side scatter (on SMs that scatter leaves free) and raises the chunk flags ----
            signal_peers_kernel<<<1, 32, 0, main>>>(m.peers(), world, rank, SIG_READY, m.d_epoch);
            B200_LAUNCH_CHECK();
            B200_CUDA(cudaEventRecord(m.ev_build, main));
            B200_CUDA(cudaStreamWaitEvent(m.xstream, m.ev_build, 0));
            B200_CUDA(cudaMemsetAsync(m.pull_work, 0, (1 + kMaxChunks * kMaxPeers) * sizeof(uint32_t), m.xstream));
            PullArgs a{};
            for (int d = 0; d < world; ++d) a.src_build[d] = m.build(d);
            a.dst_build   = m.build(rank);
            a.hdr         = m.hdr(rank);
            a.region_rows = m.seg_rows;
            a.chunk_rows  = m.chunk_rows;
            a.nchunks     = (uint32_t)m.K;
            a.rank        = rank;
            a.world       = world;
            a.done        = m.pull_work + 1;
            a.epoch       = m.d_epoch;
            a.error       = m.d_error;
            if (world > 1) {
                StreamSwap sw(c, m.xstream);
                TimedScope ts("broadcast");
                if (!m.pull_smem_set) {   // per plan, i.e. per device: function attributes belong to the device's context
                    B200_CUDA(cudaFuncSetAttribute(pull_regions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)kPullSmem));
                    m.pull_smem_set = true;
                }
                pull_regions_kernel<<<m.pull_sms, 32, kPullSmem, c.stream>>>(a);
                B200_LAUNCH_CHECK();
            }
            B200_CUDA(cudaEventRecord(m.ev_x, m.xstream));
            set_reserved_sms(world > 1 ? m.pull_sms : 0);
        }
        // ---- probe shard: partitioned locally (histogram-free regions + overflow), never moves.  Enqueued before the
        //      copies so that the host's enqueue time of those does not delay it ----
        if (m.opt_cap) {
            stage_scatter_probe_opt(m.in_pk, np, m.bits, m.opt_cap, m.cur_p, m.tup_p, m.ov_p, m.ovcnt,
                                    m.cfg.has_probe_sum ? m.in_pp : nullptr);
        } else {
            // small shards: exact histogram + the payload-aware scatter
            stage_hist(m.in_pk, np, m.bits, m.cur_p);
            const uint64_t *pp[1] = {m.in_pp};
            stage_scatter_build_local(m.in_pk, np, 0, m.bits, m.cur_p, m.tup_p, m.cfg.has_probe_sum ? 1 : 0, pp, nullptr,
                                      &m.scr_b);
            B200_CUDA(cudaMemsetAsync(m.ovcnt, 0, 4, main));
        }
        set_reserved_sms(0);
        // ---- broadcast on the copy engines: my region (histogram + tuples) into the same place of every peer's build
        //      buffer, every rank starting at a different peer, in K chunks with a 4-byte flag behind each ----
        const uint64_t used = m.seg_head + nb;   // tuple slots of my region that hold something
        for (int j = 1; j < world && !m.pull; ++j) {
            const int    d = (rank + j) % world;
            cudaStream_t s = m.copy_stream[j - 1];
            B200_CUDA(cudaStreamWaitEvent(s, m.ev_build, 0));
            for (int k = 0; k < m.K; ++k) {
                const uint64_t first = (uint64_t)k * m.chunk_rows;
                const uint64_t rows  = first < used ? std::min<uint64_t>(m.chunk_rows, used - first) : 0;
                if (rows)
                    B200_CUDA(cudaMemcpyAsync(m.build(d) + (size_t)rank * region_bytes + first * 8, my_region + first * 8,
                                              rows * 8, cudaMemcpyDeviceToDevice, s));
                B200_CUDA(cudaMemcpyAsync(&m.hdr(d)->sig[SIG_CHUNK0 + k][rank], m.d_epoch, 4, cudaMemcpyDeviceToDevice, s));
            }
            B200_CUDA(cudaEventRecord(m.ev_copy[j - 1], s));
        }
        // my own region needs no transfer: raise my own flags
        signal_self_chunks_kernel<<<1, 32, 0, main>>>(m.hdr(rank), rank, m.K, m.d_epoch);
        B200_LAUNCH_CHECK();
    }
    if (phases & 2) {
        // the histograms travel at the head of the regions: chunk 0 of every rank, then one contiguous copy of them
        wait_signal_row(m, main, SIG_CHUNK0);
        collect_hist_kernel<<<world, 256, 0, main>>>(m.build(rank), region_bytes, P, m.hist_b(rank));
        B200_LAUNCH_CHECK();
        ProjDesc pd[2];
        int      k = 0;
        if (m.cfg.has_build_sum) pd[k++] = ProjDesc{m.in_bp, nullptr, 0, B200_PROJ_IN_RID};
        if (m.cfg.has_probe_sum) pd[k++] = ProjDesc{m.in_pp, nullptr, 1, B200_PROJ_IN_RID};
        JoinWait w{&m.hdr(rank)->sig[SIG_CHUNK0][0], m.d_epoch, m.chunk_rows, m.d_error};
        stage_join_sum(m.build(rank), m.hist_b(rank), m.tup_p, m.cur_p, m.bits, k, pd, m.opt_cap, m.ov_p, m.ovcnt,
                       m.d_result, world, m.seg_rows, &m.scr_b, m.K > 1 ? &w : nullptr, m.seg_head);
        push_result_kernel<<<1, 32, 0, main>>>(m.peers(), world, rank, m.d_result, m.d_error, m.d_epoch);
        B200_LAUNCH_CHECK();
        // the next step must not overwrite my build buffer under transfers still in flight
        if (m.pull) B200_CUDA(cudaStreamWaitEvent(main, m.ev_x, 0));
        else
            for (int j = 1; j < world; ++j) B200_CUDA(cudaStreamWaitEvent(main, m.ev_copy[j - 1], 0));
    }
    if (phases & 4) {
        wait_signal_row(m, main, SIG_RESULT);
        reduce_result_kernel<<<1, 32, 0, main>>>(m.hdr(rank), world, m.d_final);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemcpyAsync(m.h_final, m.d_final, 64, cudaMemcpyDeviceToHost, main));
    }
}

static void push_block(MultiPlan &m, cudaStream_t stream, const void *src, size_t bytes, size_t shared_offset, int sig) {
    PushArgs a{};
    a.src   = static_cast<const uint32_t *>(src);
    a.words = (uint32_t)(bytes / 4);
    a.rank  = m.rank;
    a.world = m.world;
    a.sig   = sig;
    a.epoch = m.d_epoch;
    for (int d = 0; d < m.world; ++d) {
        a.dst[d] = reinterpret_cast<uint32_t *>(m.peer[d] + shared_offset);
        a.hdr[d] = m.hdr(d);
    }
    push_to_peers_kernel<<<m.world, 256, 0, stream>>>(a);
    B200_LAUNCH_CHECK();
}

static void wait_signal(MultiPlan &m, cudaStream_t stream, int sig) {
    wait_peers_kernel<<<1, 32, 0, stream>>>(m.hdr(m.rank)->sig[sig], m.world, m.d_epoch, m.d_error);
    B200_LAUNCH_CHECK();
}

// exchange plan, phases (bits): 1 build histogram + hot-key candidates, 2 hot table + build-side aggregates,
// 4 probe histograms (hot rows joined on the spot), 8 cuts + partition + exchange, 16 join + publish, 32 reduce
static void enqueue_exchange(MultiPlan &m, int phases) {
    Context     &c    = ctx();
    cudaStream_t main = c.stream;
    const int    rank = m.rank, world = m.world, K = m.K;
    const uint32_t P  = m.P;
    const uint64_t nb = m.cfg.n_build_local, np = m.cfg.n_probe_local;
    uint32_t *my_hb = m.hist_b(rank) + (size_t)rank * P;
    uint32_t *my_hp = m.hist_p(rank) + (size_t)rank * K * P;
    const size_t cand_words = kHotCandWords;
    if (phases & 1) {
        bump_epoch_kernel<<<1, 1, 0, main>>>(m.d_epoch);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemsetAsync(m.d_error, 0, 16, main));
        B200_CUDA(cudaMemsetAsync(m.d_result, 0, 64, main));
        B200_CUDA(cudaMemsetAsync(m.hot_acc, 0, 32, main));
        B200_CUDA(cudaMemsetAsync(m.hot_n, 0, 16, main));
        stage_hist(m.in_bk, nb, m.bits, my_hb);
        push_block(m, main, my_hb, (size_t)P * 4, m.off_hist_b + (size_t)rank * P * 4, SIG_HIST);
        if (m.hot) {
            // ---- candidates: keys that at least 1 / 16384 of a sample of my probe rows carry ----
            uint32_t *my_cand = m.cand(rank) + (size_t)rank * cand_words;
            B200_CUDA(cudaMemsetAsync(m.hs_keys, 0xFF, kHotSample * 4, main));
            B200_CUDA(cudaMemsetAsync(m.hs_cnt, 0, kHotSample * 4, main));
            B200_CUDA(cudaMemsetAsync(my_cand, 0, 4, main));
            const uint64_t nsample = std::min<uint64_t>(np, 1u << 20);
            if (nsample) {
                TimedScope ts("hot_sample");
                hot_sample_kernel<<<grid_for(nsample, 256, 8), 256, 0, main>>>(m.in_pk, np, nsample, m.hs_keys, m.hs_cnt);
                B200_LAUNCH_CHECK();
                const uint32_t threshold = (uint32_t)std::max<uint64_t>(8, nsample >> 14);
                hot_select_kernel<<<kHotSample / 256, 256, 0, main>>>(m.hs_keys, m.hs_cnt, threshold, my_cand);
                B200_LAUNCH_CHECK();
            }
            push_block(m, main, my_cand, cand_words * 4, m.off_cand + (size_t)rank * cand_words * 4, SIG_HOT);
        }
    }
    if ((phases & 2) && m.hot) {
        wait_signal(m, main, SIG_HOT);
        {
            TimedScope ts("hot_table");
            hot_table_kernel<<<1, 1024, 0, main>>>(m.cand(rank), world, m.hot_keys, m.hot_n);
            B200_LAUNCH_CHECK();
        }
        // ---- how many of my build rows carry each hot key, and the sum of their SUM column ----
        unsigned long long *my_agg = m.agg(rank) + (size_t)rank * 2 * kHotSlots;
        B200_CUDA(cudaMemsetAsync(my_agg, 0, (size_t)kHotSlots * 16, main));
        if (nb) {
            TimedScope ts("hot_build");
            hot_build_kernel<<<grid_for(nb, 256 * 4, 8), 256, 0, main>>>(m.in_bk, m.cfg.has_build_sum ? m.in_bp : nullptr, nb,
                                                                        m.hot_keys, m.hot_n, my_agg);
            B200_LAUNCH_CHECK();
        }
        push_block(m, main, my_agg, (size_t)kHotSlots * 16, m.off_agg + (size_t)rank * kHotSlots * 16, SIG_AGG);
    }
    if (phases & 4) {
        if (m.hot) {
            wait_signal(m, main, SIG_AGG);
            hot_reduce_kernel<<<2 * kHotSlots / 256, 256, 0, main>>>(m.agg(rank), world, m.hot_agg);
            B200_LAUNCH_CHECK();
        }
        // ---- histograms of the probe shard per chunk; rows with a hot key are joined here and not counted ----
        for (int k = 0; k < K; ++k) {
            const uint32_t a = chunk_first(m, np, k), b = chunk_first(m, np, k + 1);
            uint32_t      *h = my_hp + (size_t)k * P;
            if (!m.hot) {
                stage_hist(m.in_pk + a, b - a, m.bits, h);
                continue;
            }
            B200_CUDA(cudaMemsetAsync(h, 0, (size_t)P * 4, main));
            if (b > a) {
                constexpr int NT   = 512;
                const size_t  smem = ((size_t)P + kHotSlots) * 4;
                auto          kern = hot_hist_kernel<NT>;
                if (smem > 48 * 1024)
                    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                TimedScope ts("hist");
                kern<<<grid_for(b - a, NT * 8, 2), NT, smem, main>>>(m.in_pk + a, m.cfg.has_probe_sum ? m.in_pp + a : nullptr,
                                                                      b - a, (uint32_t)m.bits, m.hot_keys, m.hot_n, m.hot_agg,
                                                                      h, m.hot_acc);
                B200_LAUNCH_CHECK();
            }
        }
        push_block(m, main, my_hp, (size_t)K * P * 4, m.off_hist_p + (size_t)rank * K * P * 4, SIG_HIST2);
    }
    if (phases & 8) {
        wait_signal(m, main, SIG_HIST);
        wait_signal(m, main, SIG_HIST2);
        // ---- ownership cuts on the global histogram, then where my segments go ----
        uint32_t *total_b = m.total, *total_p = m.total + P;
        {
        TimedScope ts_layout("layout");
        balanced_cuts_kernel<1024><<<1, 1024, 0, main>>>(m.hist_b(rank), (uint32_t)world, m.hist_p(rank),
                                                         (uint32_t)(world * K), P, (uint32_t)world, m.cut, total_b, total_p);
        B200_LAUNCH_CHECK();
        exchange_layout_kernel<1024><<<1, 1024, 0, main>>>(m.hist_b(rank), (uint32_t)world, (uint32_t)rank, 1u, P, m.cut,
                                                           (uint32_t)world, (uint32_t)rank, m.cap_b, total_b,
                                                           m.dst_start_b, m.own_total, m.need, m.d_error);
        B200_LAUNCH_CHECK();
        exchange_layout_kernel<1024><<<1, 1024, 0, main>>>(m.hist_p(rank), (uint32_t)(world * K), (uint32_t)(rank * K),
                                                           (uint32_t)K, P, m.cut, (uint32_t)world, (uint32_t)rank, m.cap_p,
                                                           total_p, m.dst_start_p, m.own_total + P, m.need + 2, m.d_error);
        B200_LAUNCH_CHECK();
        }
        auto exchange = [&](const void *staged, uint64_t n, const uint32_t *src_off, const uint32_t *dst_start, uint32_t cap,
                            bool build_side) {
            if (n == 0) return;
            ExchangeArgs2 x{};
            x.src_tup   = static_cast<const uint64_t *>(staged);
            x.src_off   = src_off;
            x.dst_start = dst_start;
            x.cut       = m.cut;
            x.n         = (uint32_t)n;   // upper bound: the kernel reads the staged count from src_off[nparts]
            x.nparts    = P;
            x.world     = (uint32_t)world;
            x.cap       = cap;
            for (int d = 0; d < world; ++d)
                x.dst_tup[d] = reinterpret_cast<uint64_t *>(build_side ? m.build(d) : m.recv_p(d));
            StreamSwap sw(c, m.xstream);
            TimedScope ts("exchange");
            segment_exchange2_kernel<<<(unsigned)((n + 2047) / 2048), 256, 0, c.stream>>>(x);
            B200_LAUNCH_CHECK();
        };
        // ---- build shard: partition locally, exchange ----
        const uint64_t *bp[1] = {m.in_bp};
        stage_scatter_build_local(m.in_bk, nb, 0, m.bits, my_hb, m.stage_b, m.cfg.has_build_sum ? 1 : 0, bp, nullptr, &m.scr_a,
                                  m.src_off_b);
        B200_CUDA(cudaEventRecord(m.ev_build, main));
        B200_CUDA(cudaStreamWaitEvent(m.xstream, m.ev_build, 0));
        exchange(m.stage_b, nb, m.src_off_b, m.dst_start_b, m.cap_b, true);
        // ---- probe shard, chunk by chunk: the exchange of chunk k runs under the partition pass of chunk k + 1 ----
        PredSet skip{};
        if (m.hot) {
            skip.hot_keys = m.hot_keys;
            skip.hot_n    = m.hot_n;
        }
        for (int k = 0; k < K; ++k) {
            const uint32_t a = chunk_first(m, np, k), b = chunk_first(m, np, k + 1);
            const uint64_t *pp[1] = {m.in_pp ? m.in_pp + a : nullptr};
            uint64_t *staged = static_cast<uint64_t *>(m.stage_p) + a;
            // (chunk starts are even, so a 16-byte aligned shard gives 16-byte aligned chunks)
            const bool two_pass = m.bits >= m.two_pass_bits && m.cfg.has_probe_sum && b - a >= (1u << 20) && (a & 1u) == 0 &&
                                  ((reinterpret_cast<uintptr_t>(m.in_pk) | reinterpret_cast<uintptr_t>(m.in_pp)) & 15) == 0;
            if (two_pass)
                stage_scatter_two_pass(m.in_pk + a, b - a, m.bits, my_hp + (size_t)k * P, staged, m.in_pp + a, m.stage_tmp,
                                       &m.scr_a, m.src_off_p + (size_t)k * (P + 1), m.hot ? &skip : nullptr);
            else
                stage_scatter_build_local(m.in_pk + a, b - a, a, m.bits, my_hp + (size_t)k * P, staged,
                                          m.cfg.has_probe_sum ? 1 : 0, pp, nullptr, &m.scr_a,
                                          m.src_off_p + (size_t)k * (P + 1), m.hot ? &skip : nullptr);
            B200_CUDA(cudaEventRecord(m.ev_chunk[k], main));
            B200_CUDA(cudaStreamWaitEvent(m.xstream, m.ev_chunk[k], 0));
            exchange(staged, b - a, m.src_off_p + (size_t)k * (P + 1), m.dst_start_p + (size_t)k * P, m.cap_p, false);
        }
        signal_peers_kernel<<<1, 32, 0, m.xstream>>>(m.peers(), world, rank, SIG_DATA, m.d_epoch);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaEventRecord(m.ev_x, m.xstream));
        B200_CUDA(cudaStreamWaitEvent(main, m.ev_x, 0));
    }
    if (phases & 16) {
        wait_signal(m, main, SIG_DATA);
        ProjDesc pd[2];
        int      k = 0;
        if (m.cfg.has_build_sum) pd[k++] = ProjDesc{m.in_bp, nullptr, 0, B200_PROJ_IN_RID};
        if (m.cfg.has_probe_sum) pd[k++] = ProjDesc{m.in_pp, nullptr, 1, B200_PROJ_IN_RID};
        stage_join_sum(m.build(rank), m.own_total, m.recv_p(rank), m.own_total + P, m.bits, k, pd, 0, nullptr, nullptr,
                       m.d_result, 0, 0, &m.scr_b, nullptr);
        if (m.hot) {
            add_hot_result_kernel<<<1, 32, 0, main>>>(m.d_result, m.hot_acc, m.cfg.has_build_sum, m.cfg.has_probe_sum);
            B200_LAUNCH_CHECK();
        }
        push_result_kernel<<<1, 32, 0, main>>>(m.peers(), world, rank, m.d_result, m.d_error, m.d_epoch);
        B200_LAUNCH_CHECK();
    }
    if (phases & 32) {
        wait_signal(m, main, SIG_RESULT);
        reduce_result_kernel<<<1, 32, 0, main>>>(m.hdr(rank), world, m.d_final);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemcpyAsync(m.h_final, m.d_final, 64, cudaMemcpyDeviceToHost, main));
    }
}

static void enqueue(MultiPlan &m, int phases) {
    if (m.cfg.plan == B200_PLAN_BROADCAST) enqueue_broadcast(m, phases);
    else enqueue_exchange(m, phases);
}

static int all_phases(const MultiPlan &m) { return m.cfg.plan == B200_PLAN_BROADCAST ? 7 : 63; }

}  // namespace b200

using namespace b200;

namespace {
int mfail(const char *msg) {
    set_last_error(msg);
    return 1;
}
}  // namespace

extern "C" {

b200_multi *b200_multi_create(const b200_multi_config *cfg) {
    if (!cfg || cfg->world < 1 || cfg->world > kMaxPeers || cfg->rank < 0 || cfg->rank >= cfg->world) {
        set_last_error("b200_multi_create: bad rank / world (1..8 ranks)");
        return nullptr;
    }
    if (cfg->n_build_local > cfg->n_build_local_max || cfg->n_probe_local > cfg->n_probe_local_max ||
        cfg->n_build_total > kMaxRows || cfg->n_probe_local_max > (1ull << 31)) {
        set_last_error("b200_multi_create: inconsistent or too large row counts");
        return nullptr;
    }
    set_thread_device(cfg->device);
    ensure_init();
    Context &c = ctx();
    (void)c;
    auto *m   = new MultiPlan();
    m->cfg    = *cfg;
    m->rank   = cfg->rank;
    m->world  = cfg->world;
    m->device = cfg->device;
    m->nproj  = (cfg->has_build_sum ? 1 : 0) + (cfg->has_probe_sum ? 1 : 0);
    m->bits   = cfg->radix_bits > 0 ? cfg->radix_bits : auto_radix_bits(cfg->n_build_total ? cfg->n_build_total : 1, false);
    if (cfg->plan == B200_PLAN_EXCHANGE) {
        int min_bits = 2;
        while ((1 << min_bits) < 4 * cfg->world) ++min_bits;   // at least four partitions per rank
        m->bits = std::max(m->bits, min_bits);
    }
    m->P = 1u << m->bits;
    if (cfg->plan == B200_PLAN_BROADCAST) {
        // a region = [histogram: P u32 = P / 2 tuple slots][this rank's tuples]; seg_rows is the region stride in tuples
        m->seg_head   = m->P / 2;
        m->seg_rows   = m->seg_head + (uint32_t)((std::max<uint64_t>(cfg->n_build_local_max, 2) + 1) & ~1ull);
        // the copy engines move it.  Measured on 2 and 8 B200 (profiles/r2_broadcast_variants.txt): stores from an SM
        // kernel slow the concurrent probe scatter by a third even on reserved SMs, and every copy-engine operation
        // costs about 5 us whatever its size, serialised across streams — 70 operations (4 chunks + flags + histogram
        // per peer) cost more than the data at 8 GPUs.  So: one region copy and one flag per peer.
        if (const char *e = getenv("B200_BCAST")) m->pull = !strcmp(e, "pull");
        if (const char *e = getenv("B200_BCAST_SMS")) m->pull_sms = std::max(1, std::min(atoi(e), 64));
        m->K = cfg->chunks > 0 ? std::min(cfg->chunks, kMaxChunks) : (m->pull ? kMaxChunks : 1);
        if (const char *e = getenv("B200_BCAST_CHUNKS")) m->K = std::max(1, std::min(atoi(e), kMaxChunks));
        m->chunk_rows = (((m->seg_rows + (uint32_t)m->K - 1) / (uint32_t)m->K) + 1u) & ~1u;
        m->opt_cap    = (cfg->n_probe_local_max >= (1u << 20) && cfg->n_probe_local_max <= (1u << 30))
                            ? opt_region_cap(cfg->n_probe_local_max, m->bits) : 0;
    } else {
        m->K = cfg->chunks > 0 ? std::min(cfg->chunks, kMaxChunks) : (cfg->n_probe_local_max >= (1u << 24) ? 8 : 1);
        const uint64_t nb_tot = cfg->n_build_total, np_tot = cfg->n_probe_total;
        // receive capacity: the mean plus an eighth (cuts balance owners to within one partition's weight), never
        // less than one chunk table's worth — or what the caller measured
        uint64_t cb = cfg->recv_rows_build ? cfg->recv_rows_build : nb_tot / cfg->world + nb_tot / (8 * cfg->world) + 65536;
        uint64_t cp = cfg->recv_rows_probe ? cfg->recv_rows_probe : np_tot / cfg->world + np_tot / (8 * cfg->world) + 65536;
        cb = std::min<uint64_t>(cb, std::max<uint64_t>(nb_tot, 1));
        cp = std::min<uint64_t>(cp, std::max<uint64_t>(np_tot, 1));
        if (cb > kMaxRows || cp > kMaxRows) {
            delete m;
            set_last_error("b200_multi_create: receive buffers beyond 2^32-1 rows");
            return nullptr;
        }
        m->cap_b = (uint32_t)cb;
        m->cap_p = (uint32_t)cp;
        m->hot   = cfg->hot_keys >= 0 ? 1 : 0;
        if (const char *e = getenv("B200_HOT_KEYS")) m->hot = atoi(e) != 0;
    }
    layout_shared(*m);
    size_t local_bytes = 0;
    layout_local(*m, nullptr, &local_bytes);
    // (out of memory is the caller's to handle, not a reason to end the process)
    if (cudaMalloc(&m->shared, m->shared_bytes) != cudaSuccess || cudaMalloc(&m->local, local_bytes) != cudaSuccess) {
        cudaGetLastError();
        if (m->shared) cudaFree(m->shared);
        char msg[160];
        snprintf(msg, sizeof(msg), "b200_multi_create: cannot allocate %zu + %zu bytes of device memory", m->shared_bytes,
                 local_bytes);
        delete m;
        set_last_error(msg);
        return nullptr;
    }
    B200_CUDA(cudaMemset(m->shared, 0, m->off_build));   // flags, result slots and histograms start at zero
    layout_local(*m, m->local, &local_bytes);
    B200_CUDA(cudaMemset(m->local, 0, 4096));
    B200_CUDA(cudaMallocHost(&m->h_final, 64));
    for (int j = 0; j < m->world; ++j) {
        B200_CUDA(cudaStreamCreateWithFlags(&m->copy_stream[j], cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&m->ev_copy[j], cudaEventDisableTiming));
    }
    B200_CUDA(cudaStreamCreateWithFlags(&m->xstream, cudaStreamNonBlocking));
    B200_CUDA(cudaStreamCreateWithFlags(&m->gstream, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&m->ev_build, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&m->ev_hist, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&m->ev_x, cudaEventDisableTiming));
    for (int k = 0; k < kMaxChunks; ++k) B200_CUDA(cudaEventCreateWithFlags(&m->ev_chunk[k], cudaEventDisableTiming));
    m->peer[m->rank] = m->shared;
    if (const char *e = getenv("B200_TWO_PASS_BITS")) m->two_pass_bits = atoi(e);
    m->use_graph = m->world > 1;   // measured at 8 GPUs: 0.612 -> 0.579 ms per config-2 step
    if (const char *g = getenv("B200_MULTI_GRAPH")) m->use_graph = atoi(g);
    B200_CUDA(cudaDeviceSynchronize());
    return reinterpret_cast<b200_multi *>(m);
}

void b200_multi_destroy(b200_multi *plan) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    if (!m) return;
    set_thread_device(m->device);
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    if (m->graph) cudaGraphExecDestroy(m->graph);
    for (int r = 0; r < m->world; ++r)
        if (r != m->rank && m->peer[r] && m->peer_ipc[r]) cudaIpcCloseMemHandle(m->peer[r]);
    for (int j = 0; j < m->world; ++j) {
        if (m->copy_stream[j]) cudaStreamDestroy(m->copy_stream[j]);
        if (m->ev_copy[j]) cudaEventDestroy(m->ev_copy[j]);
    }
    if (m->xstream) cudaStreamDestroy(m->xstream);
    if (m->gstream) cudaStreamDestroy(m->gstream);
    if (m->ev_build) cudaEventDestroy(m->ev_build);
    if (m->ev_hist) cudaEventDestroy(m->ev_hist);
    if (m->ev_x) cudaEventDestroy(m->ev_x);
    for (int k = 0; k < kMaxChunks; ++k)
        if (m->ev_chunk[k]) cudaEventDestroy(m->ev_chunk[k]);
    if (m->h_final) cudaFreeHost(m->h_final);
    if (m->local) cudaFree(m->local);
    if (m->shared) cudaFree(m->shared);
    delete m;
}

int b200_multi_export(b200_multi *plan, unsigned char *out_handle64) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, m->shared) != cudaSuccess) {
        cudaGetLastError();
        return mfail("cudaIpcGetMemHandle failed");
    }
    memcpy(out_handle64, &h, 64);
    return 0;
}

void *b200_multi_shared_ptr(b200_multi *plan) { return reinterpret_cast<MultiPlan *>(plan)->shared; }

int b200_multi_connect_ipc(b200_multi *plan, int peer, const unsigned char *handle64) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    if (peer < 0 || peer >= m->world) return mfail("bad peer");
    if (peer == m->rank) return 0;
    set_thread_device(m->device);
    (void)ctx();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return mfail("cudaIpcOpenMemHandle failed (no peer access between the two GPUs?)");
    }
    m->peer[peer]     = static_cast<unsigned char *>(p);
    m->peer_ipc[peer] = true;
    return 0;
}

int b200_multi_connect_ptr(b200_multi *plan, int peer, void *peer_shared, int peer_device) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    if (peer < 0 || peer >= m->world) return mfail("bad peer");
    if (peer == m->rank) return 0;
    set_thread_device(m->device);
    (void)ctx();
    if (peer_device != m->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            return mfail("cudaDeviceEnablePeerAccess failed");
        }
        cudaGetLastError();
    }
    m->peer[peer] = static_cast<unsigned char *>(peer_shared);
    return 0;
}

int b200_multi_radix_bits(b200_multi *plan) { return reinterpret_cast<MultiPlan *>(plan)->bits; }

int b200_multi_enqueue(b200_multi *plan, const uint64_t *d_build_keys, const uint64_t *d_build_sum,
                       const uint64_t *d_probe_keys, const uint64_t *d_probe_sum, int phases) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    for (int r = 0; r < m->world; ++r)
        if (!m->peer[r]) return mfail("b200_multi_enqueue: not every peer is connected");
    if ((m->cfg.has_build_sum && !d_build_sum) || (m->cfg.has_probe_sum && !d_probe_sum)) return mfail("missing SUM column");
    set_thread_device(m->device);
    Context &c = ctx();
    m->in_bk = d_build_keys;
    m->in_bp = d_build_sum;
    m->in_pk = d_probe_keys;
    m->in_pp = d_probe_sum;
    if (phases <= 0) phases = all_phases(*m);
    if (m->use_graph && phases == all_phases(*m) && !profiling_enabled()) {
        const uint64_t *in[4] = {d_build_keys, d_build_sum, d_probe_keys, d_probe_sum};
        if (!m->graph || memcmp(in, m->g_in, sizeof(in)) != 0) {
            if (m->graph) {
                cudaGraphExecDestroy(m->graph);
                m->graph = nullptr;
            }
            // captured on a stream of the plan's own: the caller's may be the legacy default stream, which cannot be
            // captured; the copy / exchange streams fork from it and join it again through the step's events
            cudaGraph_t g = nullptr;
            {
                StreamSwap sw(c, m->gstream);
                const uint64_t before = t_launches;
                B200_CUDA(cudaStreamBeginCapture(m->gstream, cudaStreamCaptureModeThreadLocal));
                enqueue(*m, phases);
                B200_CUDA(cudaStreamEndCapture(m->gstream, &g));
                // the launch counter counts kernels that RUN: nothing ran during the capture, every replay runs them all
                m->graph_kernels = t_launches - before;
                g_launches.fetch_sub(m->graph_kernels);
            }
            B200_CUDA(cudaGraphInstantiate(&m->graph, g, 0));
            cudaGraphDestroy(g);
            memcpy(m->g_in, in, sizeof(in));
        }
        B200_CUDA(cudaGraphLaunch(m->graph, c.stream));
        g_launches.fetch_add(m->graph_kernels);
    } else {
        enqueue(*m, phases);
    }
    m->pending = true;
    return 0;
}

int b200_multi_finish(b200_multi *plan, uint64_t *out_sums, uint64_t *out_matches) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    set_thread_device(m->device);
    Context &c = ctx();
    B200_CUDA(cudaStreamSynchronize(c.stream));
    m->pending = false;
    const int k = m->nproj;
    if (m->h_final[7] != 0) {
        return mfail((m->h_final[7] & 1ull) || m->cfg.plan == B200_PLAN_BROADCAST
                         ? "multi-GPU step failed: a peer's signal did not arrive (time-out)"
                         : "multi-GPU step failed: exchange receive buffers too small for this input "
                           "(raise recv_rows_build / recv_rows_probe)");
    }
    if (m->cfg.plan == B200_PLAN_BROADCAST && m->h_final[k + 1] != 0) {
        // some rank's histogram-free probe scatter overflowed its regions (skewed keys): every rank sees the same
        // total, redoes its join with the exact overflow pass, and the results are exchanged once more
        ProjDesc pd[2];
        int      n = 0;
        if (m->cfg.has_build_sum) pd[n++] = ProjDesc{m->in_bp, nullptr, 0, B200_PROJ_IN_RID};
        if (m->cfg.has_probe_sum) pd[n++] = ProjDesc{m->in_pp, nullptr, 1, B200_PROJ_IN_RID};
        JoinResult j = stage_join_sum(m->build(m->rank), m->hist_b(m->rank), m->tup_p, m->cur_p, m->bits, n, pd, m->opt_cap,
                                      m->ov_p, m->ovcnt, nullptr, m->world, m->seg_rows, nullptr, nullptr, m->seg_head);
        unsigned long long local[8] = {j.m, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i) local[1 + i] = j.sums[i];
        B200_CUDA(cudaMemcpyAsync(m->d_result, local, 64, cudaMemcpyHostToDevice, c.stream));
        bump_epoch_kernel<<<1, 1, 0, c.stream>>>(m->d_epoch);
        B200_LAUNCH_CHECK();
        push_result_kernel<<<1, 32, 0, c.stream>>>(m->peers(), m->world, m->rank, m->d_result, m->d_error, m->d_epoch);
        B200_LAUNCH_CHECK();
        wait_peers_kernel<<<1, 32, 0, c.stream>>>(m->hdr(m->rank)->sig[SIG_RESULT], m->world, m->d_epoch, m->d_error);
        B200_LAUNCH_CHECK();
        reduce_result_kernel<<<1, 32, 0, c.stream>>>(m->hdr(m->rank), m->world, m->d_final);
        B200_LAUNCH_CHECK();
        B200_CUDA(cudaMemcpyAsync(m->h_final, m->d_final, 64, cudaMemcpyDeviceToHost, c.stream));
        B200_CUDA(cudaStreamSynchronize(c.stream));
        if (m->h_final[7] != 0) return mfail("multi-GPU step failed: a peer's signal did not arrive (time-out)");
    }
    if (out_matches) *out_matches = m->h_final[0];
    for (int i = 0; i < k; ++i) out_sums[i] = m->h_final[1 + i];
    return 0;
}

/* rows this rank received in the last exchange step: {build, probe} (after b200_multi_finish) */
int b200_multi_received(b200_multi *plan, uint64_t *out2) {
    auto *m = reinterpret_cast<MultiPlan *>(plan);
    if (m->cfg.plan != B200_PLAN_EXCHANGE) return mfail("not an exchange plan");
    set_thread_device(m->device);
    uint32_t h[4];
    B200_CUDA(cudaMemcpy(h, m->need, sizeof(h), cudaMemcpyDeviceToHost));
    out2[0] = h[0];
    out2[1] = h[2];
    return 0;
}

/* One process, one host thread per GPU: the whole join from DEVICE-resident position shards.
 * shard pointers: [g] = this GPU's shard of the column (NULL sum columns = no SUM on that side). */
int b200_join_sum_multi(int n_gpus, int plan_kind, const uint64_t *const *d_build_keys, const uint64_t *const *d_build_sum,
                        const uint64_t *n_build, const uint64_t *const *d_probe_keys, const uint64_t *const *d_probe_sum,
                        const uint64_t *n_probe, int steps, uint64_t *out_sums, uint64_t *out_matches, double *out_ms) {
    if (n_gpus < 1 || n_gpus > kMaxPeers) return mfail("1..8 GPUs");
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < n_gpus) return mfail("not enough visible GPUs");
    uint64_t nb_tot = 0, np_tot = 0, nb_max = 0, np_max = 0;
    for (int g = 0; g < n_gpus; ++g) {
        nb_tot += n_build[g];
        np_tot += n_probe[g];
        nb_max = std::max(nb_max, n_build[g]);
        np_max = std::max(np_max, n_probe[g]);
    }
    std::vector<b200_multi *> plans((size_t)n_gpus, nullptr);
    std::vector<int>          rc((size_t)n_gpus, 0);
    std::vector<std::string>  err((size_t)n_gpus);
    auto on_all = [&](auto fn) {
        std::vector<std::thread> th;
        for (int g = 0; g < n_gpus; ++g) th.emplace_back([&, g] { fn(g); });
        for (auto &t : th) t.join();
    };
    on_all([&](int g) {
        b200_multi_config cfg{};
        cfg.plan = plan_kind;
        cfg.rank = g;
        cfg.world = n_gpus;
        cfg.device = g;
        cfg.n_build_total = nb_tot;
        cfg.n_probe_total = np_tot;
        cfg.n_build_local = n_build[g];
        cfg.n_probe_local = n_probe[g];
        cfg.n_build_local_max = nb_max;
        cfg.n_probe_local_max = np_max;
        cfg.has_build_sum = d_build_sum && d_build_sum[g];
        cfg.has_probe_sum = d_probe_sum && d_probe_sum[g];
        plans[(size_t)g] = b200_multi_create(&cfg);
        if (!plans[(size_t)g]) { rc[(size_t)g] = 1; err[(size_t)g] = last_error_string(); }
    });
    for (int g = 0; g < n_gpus; ++g)
        if (rc[(size_t)g]) {
            for (auto *p : plans) b200_multi_destroy(p);
            return mfail(err[(size_t)g].c_str());
        }
    on_all([&](int g) {
        for (int r = 0; r < n_gpus; ++r)
            if (b200_multi_connect_ptr(plans[(size_t)g], r, b200_multi_shared_ptr(plans[(size_t)r]), r)) {
                rc[(size_t)g] = 1;
                err[(size_t)g] = last_error_string();
            }
    });
    std::vector<double> ms((size_t)n_gpus, 0.0);
    if (!std::any_of(rc.begin(), rc.end(), [](int v) { return v != 0; })) {
        on_all([&](int g) {
            b200_multi *p = plans[(size_t)g];
            uint64_t    sums[2] = {0, 0}, mm = 0;
            auto step = [&]() {
                int r = b200_multi_enqueue(p, d_build_keys[g], d_build_sum ? d_build_sum[g] : nullptr, d_probe_keys[g],
                                           d_probe_sum ? d_probe_sum[g] : nullptr, 0);
                if (!r) r = b200_multi_finish(p, sums, &mm);
                if (r) { rc[(size_t)g] = 1; err[(size_t)g] = last_error_string(); }
                return r;
            };
            if (step()) return;   // warm-up (and the answer, when steps == 0)
            if (steps > 0) {
                const auto t0 = std::chrono::steady_clock::now();
                for (int s = 0; s < steps; ++s)
                    if (step()) return;
                ms[(size_t)g] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
            }
            if (g == 0) {   // every rank holds the same totals: {SUM(build column)?, SUM(probe column)?}
                const int k = ((d_build_sum && d_build_sum[0]) ? 1 : 0) + ((d_probe_sum && d_probe_sum[0]) ? 1 : 0);
                for (int i = 0; i < k; ++i) out_sums[i] = sums[i];
                if (out_matches) *out_matches = mm;
            }
        });
    }
    on_all([&](int g) { b200_multi_destroy(plans[(size_t)g]); });
    for (int g = 0; g < n_gpus; ++g)
        if (rc[(size_t)g]) return mfail(err[(size_t)g].c_str());
    if (out_ms) *out_ms = *std::max_element(ms.begin(), ms.end());
    return 0;
}

}  // extern "C"
