// multi_kernels.cuh — the small kernels of the multi-GPU plans (multi.cu): synchronisation between GPUs through
// flags in peer memory instead of NCCL collectives, the layout of the radix-sharded exchange, and the exchange
// itself (the scatter's second half, fused with the all-to-all: stores straight into the owners' memory).
//
// The reference is a single process (no collective anywhere); what these kernels shard is its bucket
// independence, rhjoin.c:42-57 — one JoinJob per bucket, no cross-bucket state.
//
// Flag protocol.  Every rank owns a SharedHeader at the start of its shared region.  sig[s][r] is written only by
// rank r (a 4-byte copy-engine write behind the data it announces, or a st.release.sys from a kernel) and holds
// the epoch (step number) of r's latest signal s; waiting = spinning until it reaches this rank's own epoch.
// Epochs only grow, so nothing is ever reset, and they live in device memory (not in kernel arguments), so a
// captured CUDA graph of a step stays valid for the next step.
#pragma once

#include "kernels.cuh"

namespace b200 {

constexpr int kNSig      = 16;
constexpr int kMaxChunks = 8;
enum MultiSig {
    SIG_HIST   = 0,   // my (build-side) histogram is in your hist_all
    SIG_RESULT = 1,   // my {matches, sums, flags} are in your result[]
    SIG_DATA   = 2,   // exchange plan: every row I owe you has landed
    SIG_HIST2  = 3,   // exchange plan: my probe-side histograms (per chunk) are in your hist_all
    SIG_HOT    = 4,   // exchange plan: my hot-key candidates are in your cand[]
    SIG_AGG    = 5,   // exchange plan: my build-side aggregates of the hot keys are in your agg[]
    SIG_READY  = 6,   // broadcast plan, pull variant: my region of my build buffer is complete — fetch it
    SIG_CHUNK0 = 8,   // broadcast plan: chunk c of my build region has landed (SIG_CHUNK0 + c)
};

struct SharedHeader {
    uint32_t           sig[kNSig][kMaxPeers];
    unsigned long long result[kMaxPeers][8];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void bump_epoch_kernel(uint32_t *epoch) { *epoch += 1u; }

// one lane per peer: wait for signal row `row` (kMaxPeers flags) of every rank
__global__ void wait_peers_kernel(const uint32_t *row, int world, const uint32_t *epoch, uint32_t *error) {
    const int r = threadIdx.x;
    if (r < world && !spin_until_epoch(row + r, *epoch)) *error = 1u;
}

struct PeerPtrs {
    SharedHeader *hdr[kMaxPeers];
};

// announce `sig` to every rank (stream-ordered behind this rank's stores into their memory)
__global__ void signal_peers_kernel(PeerPtrs peers, int world, int rank, int sig, const uint32_t *epoch) {
    const int d = threadIdx.x;
    if (d < world) {
        __threadfence_system();
        st_release_sys_u32(&peers.hdr[d]->sig[sig][rank], *epoch);
    }
}

// my own region needs no transfer: raise all of its chunk flags in my own header
static __global__ void signal_self_chunks_kernel(SharedHeader *hdr, int rank, int nchunks, const uint32_t *epoch) {
    if ((int)threadIdx.x < nchunks) st_release_sys_u32(&hdr->sig[SIG_CHUNK0 + threadIdx.x][rank], *epoch);
}

// {matches, sums..., flags} of this rank into slot `rank` of every rank's header, then SIG_RESULT
__global__ void push_result_kernel(PeerPtrs peers, int world, int rank, const unsigned long long *local,
                                   const uint32_t *extra_flag, const uint32_t *epoch) {
    const int d = threadIdx.x;
    if (d < world) {
        volatile unsigned long long *dst = peers.hdr[d]->result[rank];
        for (int k = 0; k < 8; ++k) dst[k] = local[k] + (k == 7 && extra_flag ? (unsigned long long)*extra_flag : 0ull);
        __threadfence_system();
        st_release_sys_u32(&peers.hdr[d]->sig[SIG_RESULT][rank], *epoch);
    }
}

// sum of all ranks' result slots (u64 sums wrap mod 2^64 exactly as inter_res.c:332-333's accumulation does)
__global__ void reduce_result_kernel(const SharedHeader *hdr, int world, unsigned long long *final8) {
    const int k = threadIdx.x;
    if (k < 8) {
        unsigned long long s = 0;
        for (int r = 0; r < world; ++r) s += reinterpret_cast<const volatile unsigned long long *>(hdr->result[r])[k];
        final8[k] = s;
    }
}

// ---------------------------------------------------------------------------
// Pull variant of the broadcast (B200_BCAST=pull): instead of every rank pushing its region to seven peers with
// the copy engines, every rank FETCHES the seven regions it needs with the TMA unit: one thread per CTA drives a ring
// of 32 KB stages in shared memory — cp.async.bulk global(peer, over NVLink) -> shared, completion on an mbarrier,
// then cp.async.bulk shared -> global(local HBM) — so that 160 KB per SM are in flight with no register or LSU
// cost (a fetch kernel built on ordinary loads was latency-bound: 117 MB took 0.9 ms on 12 SMs,
// profiles/r2_broadcast_variants.txt).  Remote LOADS do not disturb the concurrent probe scatter the way remote
// stores do, and the kernel raises the per-chunk flags the join waits on itself, locally, so the join overlaps the
// transfer chunk by chunk.  A small grid on SMs the probe-side scatter leaves free (a CTA takes nearly all of an
// SM's shared memory, so it owns its SM); a piece of a peer's region is fetched only after that peer has announced
// (SIG_READY) that the region is complete.
// ---------------------------------------------------------------------------
struct PullArgs {
    const unsigned char *src_build[kMaxPeers];   // peer d's build buffer (region d of it is what d produced)
    unsigned char       *dst_build;              // my build buffer
    SharedHeader        *hdr;                    // my header: SIG_READY[d] written by peer d, SIG_CHUNK0+c[d] by this kernel
    uint32_t             region_rows, chunk_rows, nchunks;
    int                  rank, world;
    uint32_t            *done;                   // [kMaxChunks][kMaxPeers] pieces completed, zero at launch
    const uint32_t      *epoch;
    uint32_t            *error;
};
constexpr uint32_t kPullStage  = 32 * 1024;   // bytes per ring stage = one piece
constexpr int      kPullStages = 6;
constexpr size_t   kPullSmem   = (size_t)kPullStage * kPullStages;

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

static __global__ void __launch_bounds__(32) pull_regions_kernel(const PullArgs a) {
    extern __shared__ __align__(128) unsigned char pull_ring[];
    __shared__ __align__(8) unsigned long long pull_full[kPullStages];
    if (threadIdx.x != 0) return;   // one thread drives the whole pipeline: the TMA unit does the moving
    const uint32_t epoch  = *a.epoch;
    const uint32_t ring   = (uint32_t)__cvta_generic_to_shared(pull_ring);
    const uint32_t mbar0  = (uint32_t)__cvta_generic_to_shared(pull_full);
    for (int s = 0; s < kPullStages; ++s) mbar_init(mbar0 + 8u * s, 1u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    // pieces in (chunk, peer, position) order — the first chunks of every region arrive first — dealt round-robin to
    // the CTAs; piece = (chunk c, j-th peer after me, k-th 32 KB of that chunk)
    const uint32_t npeer = (uint32_t)a.world - 1u;
    auto chunk_bytes = [&](uint32_t c) -> uint32_t {
        const uint64_t first = (uint64_t)c * a.chunk_rows;
        return first < a.region_rows ? (uint32_t)(min((uint64_t)a.chunk_rows, (uint64_t)a.region_rows - first) * 8u) : 0u;
    };
    auto pieces_of = [&](uint32_t c) -> uint32_t { return (chunk_bytes(c) + kPullStage - 1) / kPullStage; };
    uint32_t total = 0;
    for (uint32_t c = 0; c < a.nchunks; ++c) total += pieces_of(c) * npeer;
    struct Piece { uint32_t c, d, bytes; size_t off; };
    auto piece_at = [&](uint32_t idx) -> Piece {
        uint32_t c = 0;
        while (idx >= pieces_of(c) * npeer) idx -= pieces_of(c++) * npeer;
        const uint32_t j = idx / pieces_of(c), k = idx % pieces_of(c);
        Piece          p;
        p.c     = c;
        p.d     = (uint32_t)((a.rank + 1 + (int)j) % a.world);
        p.off   = ((size_t)p.d * a.region_rows + (size_t)c * a.chunk_rows) * 8 + (size_t)k * kPullStage;
        p.bytes = min(kPullStage, chunk_bytes(c) - k * kPullStage);
        return p;
    };
    uint32_t ready = 0;   // bit d: peer d has announced its region
    auto issue_load = [&](uint32_t n /* my n-th piece */) {
        const Piece p = piece_at(blockIdx.x + n * gridDim.x);
        if (!((ready >> p.d) & 1u)) {
            if (!spin_until_epoch(&a.hdr->sig[SIG_READY][p.d], epoch)) *a.error = 1u;
            ready |= 1u << p.d;
        }
        const uint32_t slot = n % kPullStages;
        mbar_expect_tx(mbar0 + 8u * slot, p.bytes);
        bulk_g2s(ring + slot * kPullStage, a.src_build[p.d] + p.off, p.bytes, mbar0 + 8u * slot);
    };
    auto account = [&](uint32_t n) {   // my n-th piece is in local memory
        const Piece    p        = piece_at(blockIdx.x + n * gridDim.x);
        const uint32_t finished = atomicAdd(&a.done[p.c * kMaxPeers + p.d], 1u) + 1u;
        if (finished == pieces_of(p.c)) {
            __threadfence();
            st_release_sys_u32(&a.hdr->sig[SIG_CHUNK0 + p.c][p.d], epoch);
        }
    };
    const uint32_t mine = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    for (uint32_t n = 0; n < mine && n < (uint32_t)kPullStages; ++n) issue_load(n);   // piece p lives in stage p % kPullStages
    for (uint32_t n = 0; n < mine; ++n) {
        const uint32_t slot = n % kPullStages, parity = (n / kPullStages) & 1u;
        const long long t0  = clock64();
        while (!mbar_try_wait(mbar0 + 8u * slot, parity))
            if (clock64() - t0 > 4000000000ll) { *a.error = 1u; return; }
        const Piece p = piece_at(blockIdx.x + n * gridDim.x);
        bulk_s2g(a.dst_build + p.off, ring + slot * kPullStage, p.bytes);
        // every store but the one just issued is complete: piece n - 1 is in local memory and its stage is free
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
        if (n >= 1) {
            __threadfence();
            account(n - 1);
            if (n - 1 + kPullStages < mine) issue_load(n - 1 + kPullStages);   // into the stage piece n - 1 just left
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (mine) {
        __threadfence();
        account(mine - 1);
    }
}

// hist_all[r][i] = the histogram at the head of region r of the build buffer (broadcast plan: a region travels with its
// histogram in front of its tuples, so that one copy per peer moves both)
static __global__ void __launch_bounds__(256)
collect_hist_kernel(const unsigned char *__restrict__ build, size_t region_bytes, uint32_t nparts, uint32_t *__restrict__ hist_all) {
    const uint32_t *h = reinterpret_cast<const uint32_t *>(build + (size_t)blockIdx.x * region_bytes);
    for (uint32_t i = threadIdx.x; i < nparts; i += 256) hist_all[(size_t)blockIdx.x * nparts + i] = __ldcg(h + i);
}

// a block of this rank's memory into the same place of every peer's shared region, then `sig` (small blocks:
// histograms, candidate lists, aggregates — a few KB to 128 KB; one CTA per peer)
struct PushArgs {
    const uint32_t *src;
    uint32_t        words;
    uint32_t       *dst[kMaxPeers];
    SharedHeader   *hdr[kMaxPeers];
    int             rank, world, sig;
    const uint32_t *epoch;
};
static __global__ void __launch_bounds__(256) push_to_peers_kernel(const PushArgs a) {
    const int d = blockIdx.x;
    if (d >= a.world) return;
    if (d != a.rank)
        for (uint32_t i = threadIdx.x; i < a.words; i += 256) a.dst[d][i] = a.src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys_u32(&a.hdr[d]->sig[a.sig][a.rank], *a.epoch);
}

// ---------------------------------------------------------------------------
// Hot keys (config 4: Zipf probe keys).  A key that a large share of the probe rows carries would send all of them
// to one owner and hammer one histogram bin and one hash-table slot on the way.  Instead the ranks agree on a
// small set of hot keys, every rank learns how many build rows carry each of them and the sum of their SUM column
// (replicating the hot build keys), and a probe row with a hot key is joined where it is, while the histogram pass
// streams by: matches += cnt[key], SUM(build) += sum[key], SUM(probe) += cnt[key] * value.  Such rows are neither
// partitioned nor exchanged nor probed.
//   hot_sample_kernel    counts a sample of the local probe keys in an open-addressing table
//   hot_select_kernel    keys seen at least `threshold` times -> this rank's candidate list {n, keys..., sample counts...}
//   hot_table_kernel     every rank, from the SAME gathered candidate lists: a direct-mapped table of kHotSlots keys
//                        (of the candidates that share a slot the one with the largest sample count is hot)
//   hot_build_kernel     local build rows with a hot key: agg[slot] += {1, value}
//   hot_reduce_kernel    sum of all ranks' aggregates
//   hot_hist_kernel      the probe-side histogram pass: hot rows are accumulated, the others counted
// ---------------------------------------------------------------------------
// (kHotSlots = 8192 direct-mapped slots, hot_slot(), kHotEmpty: kernels.cuh — the scatter skips hot rows too)
constexpr uint32_t kHotMaxCand  = 2048;                  // candidates per rank
constexpr uint32_t kHotCandWords = 1 + 2 * kHotMaxCand;  // a rank's list: count, keys[kHotMaxCand], sample counts[kHotMaxCand]
constexpr uint32_t kHotSample  = 1u << 16;   // slots of the sampling table

static __global__ void __launch_bounds__(256)
hot_sample_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint64_t nsample, uint32_t *__restrict__ t_keys,
                  uint32_t *__restrict__ t_cnt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nsample; s += stride) {
        const uint32_t key = (uint32_t)ld_stream_u64(keys + (s * n) / nsample);
        if (key == kHotEmpty) continue;
        uint32_t slot = (key * 0x85EBCA6Bu) >> 16;
        for (int probe = 0; probe < 16; ++probe, slot = (slot + 1) & (kHotSample - 1)) {
            const uint32_t old = atomicCAS(&t_keys[slot], kHotEmpty, key);
            if (old == kHotEmpty || old == key) {
                atomicAdd(&t_cnt[slot], 1u);
                break;
            }
        }
    }
}
static __global__ void __launch_bounds__(256)
hot_select_kernel(const uint32_t *__restrict__ t_keys, const uint32_t *__restrict__ t_cnt, uint32_t threshold,
                  uint32_t *__restrict__ cand /* [kHotCandWords] */) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < kHotSample && t_keys[slot] != kHotEmpty && t_cnt[slot] >= threshold) {
        const uint32_t at = atomicAdd(&cand[0], 1u);
        if (at < kHotMaxCand) {
            cand[1 + at]               = t_keys[slot];
            cand[1 + kHotMaxCand + at] = t_cnt[slot];
        }
    }
}
static __global__ void __launch_bounds__(1024)
hot_table_kernel(const uint32_t *__restrict__ cand_all /* [world][kHotCandWords] */, int world,
                 uint32_t *__restrict__ hot_keys /* [kHotSlots] */, uint32_t *__restrict__ hot_n) {
    // Every rank must end up with the same table.  A slot belongs to the candidate with the LARGEST sample count that
    // maps to it (the table is direct-mapped: losing the slot means being exchanged like any other key, which the
    // hottest keys must not be), ties to the first in rank / position order: an atomicMin over
    // [inverted count | rank, position], so all 1024 threads work on the lists (one thread walking up to
    // world x 2048 candidates through dependent global loads took ~1 ms at 8 ranks).
    static_assert(kMaxPeers * kHotMaxCand <= (1u << 14), "rank and position in 14 bits");
    __shared__ uint32_t s_ord[kHotSlots];
    __shared__ uint32_t s_n;
    for (uint32_t i = threadIdx.x; i < kHotSlots; i += 1024) s_ord[i] = 0xFFFFFFFFu;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int r = 0; r < world; ++r) {
        const uint32_t *c   = cand_all + (size_t)r * kHotCandWords;
        const uint32_t  cnt = min(c[0], kHotMaxCand);
        for (uint32_t i = threadIdx.x; i < cnt; i += 1024) {
            const uint32_t weight = min(c[1 + kHotMaxCand + i], 0x3FFFEu);
            atomicMin(&s_ord[hot_slot(c[1 + i])], ((0x3FFFEu - weight) << 14) | ((uint32_t)r * kHotMaxCand + i));
        }
    }
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t slot = threadIdx.x; slot < kHotSlots; slot += 1024) {
        const uint32_t ord = s_ord[slot];
        uint32_t       key = kHotEmpty;
        if (ord != 0xFFFFFFFFu) {
            const uint32_t at = ord & 0x3FFFu;
            key = cand_all[(size_t)(at / kHotMaxCand) * kHotCandWords + 1 + at % kHotMaxCand];
            ++mine;
        }
        hot_keys[slot] = key;
    }
    if (mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (threadIdx.x == 0) *hot_n = s_n;
}
static __global__ void __launch_bounds__(256)
hot_build_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, uint64_t n,
                 const uint32_t *__restrict__ hot_keys, const uint32_t *__restrict__ hot_n,
                 unsigned long long *__restrict__ agg /* [kHotSlots][2] */) {
    if (*hot_n == 0) return;
    __shared__ uint32_t s_tab[kHotSlots];
    for (uint32_t i = threadIdx.x; i < kHotSlots; i += 256) s_tab[i] = hot_keys[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t key = (uint32_t)ld_stream_u64(keys + i), slot = hot_slot(key);
        if (s_tab[slot] == key && key != kHotEmpty) {
            atomicAdd(&agg[2 * slot], 1ull);
            if (vals) atomicAdd(&agg[2 * slot + 1], (unsigned long long)ld_stream_u64(vals + i));
        }
    }
}
static __global__ void __launch_bounds__(256)
hot_reduce_kernel(const unsigned long long *__restrict__ agg_all /* [world][kHotSlots][2] */, int world,
                  unsigned long long *__restrict__ hot_agg /* [kHotSlots][2] */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * kHotSlots) {
        unsigned long long s = 0;
        for (int r = 0; r < world; ++r) s += agg_all[(size_t)r * 2 * kHotSlots + i];
        hot_agg[i] = s;
    }
}
// probe-side histogram of key & mask over rows [0, n) with the hot rows taken out and joined on the spot:
// acc[0] += cnt, acc[1] += build sum, acc[2] += cnt * probe value
template <int NT>
__global__ void __launch_bounds__(NT)
hot_hist_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, uint64_t n, uint32_t radix_bits,
                const uint32_t *__restrict__ hot_keys, const uint32_t *__restrict__ hot_n,
                const unsigned long long *__restrict__ hot_agg, uint32_t *__restrict__ ghist,
                unsigned long long *__restrict__ acc) {
    extern __shared__ uint32_t sh_dyn[];
    constexpr int  U       = 8;                       // rows per thread and tile: all loads of a tile in flight together
    uint32_t      *sh_hist = sh_dyn;
    const uint32_t nbins   = 1u << radix_bits, mask = nbins - 1u;
    uint32_t      *s_tab   = sh_dyn + nbins;
    const bool     hot     = *hot_n != 0;
    for (uint32_t b = threadIdx.x; b < nbins; b += NT) sh_hist[b] = 0;
    if (hot)
        for (uint32_t i = threadIdx.x; i < kHotSlots; i += NT) s_tab[i] = hot_keys[i];
    __syncthreads();
    unsigned long long m = 0, sb = 0, sp = 0;
    const bool     vec    = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    const uint64_t ntiles = (n + (uint64_t)NT * U - 1) / ((uint64_t)NT * U);
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t base = tile * NT * U;
        uint64_t       k[U], v[U];
        const bool     full = vec && base + (uint64_t)NT * U <= n;
        if (full) {   // thread t holds rows base + ((j / 2) * NT + t) * 2 + (j & 1): 128-bit loads
#pragma unroll
            for (int j = 0; j < U; j += 2) {
                const uint64_t   r  = base + ((uint64_t)(j >> 1) * NT + threadIdx.x) * 2u;
                const ulonglong2 kk = ld_stream_u64x2(keys + r);
                k[j]                = kk.x;
                k[j + 1]            = kk.y;
                if (vals) {
                    const ulonglong2 vv = ld_stream_u64x2(vals + r);
                    v[j]                = vv.x;
                    v[j + 1]            = vv.y;
                } else {
                    v[j] = v[j + 1] = 0;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint64_t r = base + (uint64_t)j * NT + threadIdx.x;
                k[j]             = r < n ? ld_stream_u64(keys + r) : 0ull;
                v[j]             = (r < n && vals) ? ld_stream_u64(vals + r) : 0ull;
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (!full && base + (uint64_t)j * NT + threadIdx.x >= n) continue;
            const uint32_t key    = (uint32_t)k[j];
            bool           is_hot = false;
            if (hot) {
                const uint32_t slot = hot_slot(key);
                if (s_tab[slot] == key && key != kHotEmpty) {
                    is_hot                       = true;
                    const unsigned long long cnt = hot_agg[2 * slot];
                    m += cnt;
                    sb += hot_agg[2 * slot + 1];
                    sp += cnt * (unsigned long long)v[j];
                }
            }
            if (!is_hot) atomicAdd(&sh_hist[key & mask], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += NT) {
        const uint32_t c = sh_hist[b];
        if (c) atomicAdd(&ghist[b], c);
    }
    m  = warp_sum_u64(m);
    sb = warp_sum_u64(sb);
    sp = warp_sum_u64(sp);
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(acc + 0, m);
        atomicAdd(acc + 1, sb);
        atomicAdd(acc + 2, sp);
    }
}
// {matches, SUM(build), SUM(probe)} of the hot rows into the join's result slots
static __global__ void add_hot_result_kernel(unsigned long long *result, const unsigned long long *acc, int has_build_sum,
                                             int has_probe_sum) {
    if (threadIdx.x == 0) {
        result[0] += acc[0];
        int k = 1;
        if (has_build_sum) result[k++] += acc[1];
        if (has_probe_sum) result[k] += acc[2];
    }
}

// ---------------------------------------------------------------------------
// Radix-sharded exchange (SURVEY §8e "all-to-all"; config 4).
//
// Ownership: rank g owns the contiguous partition range [cut[g], cut[g+1]).  The cuts are placed on the GLOBAL
// histogram (build + probe rows per partition, all ranks, all chunks) so that every owner receives about the same
// number of rows: with Zipf(1.0) probe keys the partition of the hottest key alone holds 1/k of all probe rows
// (3.7 % for k = 27), and equal-width ranges would hand its owner 1.3x the mean; cut placement keeps max/mean
// within a partition's weight of 1.
// hist_b[Vb][P], hist_p[Vp][P]: all-gathered histograms per virtual rank (a virtual rank is one (rank, chunk)
// pair: the probe shard is partitioned and exchanged chunk by chunk so that the exchange of chunk c overlaps the
// partition pass of chunk c + 1).  One CTA.
//   cut[world + 1], total_b[P], total_p[P]
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
balanced_cuts_kernel(const uint32_t *__restrict__ hist_b, uint32_t vb, const uint32_t *__restrict__ hist_p, uint32_t vp,
                     uint32_t nparts, uint32_t world, uint32_t *__restrict__ cut, uint32_t *__restrict__ total_b,
                     uint32_t *__restrict__ total_p) {
    __shared__ unsigned long long wsum[NT / 32 + 1];
    const uint32_t per   = (nparts + NT - 1) / NT;
    const uint32_t first = threadIdx.x * per;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    unsigned long long s = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t p = first + k;
        if (p < nparts) {
            uint32_t tb = 0, tp = 0;
            for (uint32_t v = 0; v < vb; ++v) tb += hist_b[(size_t)v * nparts + p];
            for (uint32_t v = 0; v < vp; ++v) tp += hist_p[(size_t)v * nparts + p];
            total_b[p] = tb;
            total_p[p] = tp;
            s += (unsigned long long)tb + tp;
        }
    }
    unsigned long long incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int w = 0; w < NT / 32; ++w) {
            const unsigned long long t = wsum[w];
            wsum[w] = run;
            run += t;
        }
        wsum[NT / 32] = run;
        cut[0]        = 0;
        cut[world]    = nparts;
    }
    __syncthreads();
    const unsigned long long grand = wsum[NT / 32];
    unsigned long long       run   = wsum[wid] + incl - s;   // weight of all partitions before `first`
    // cut[g] (0 < g < world) = first partition whose preceding weight reaches g * grand / world
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t p = first + k;
        if (p < nparts) {
            const unsigned long long w = (unsigned long long)total_b[p] + total_p[p];
            for (uint32_t g = 1; g < world; ++g) {
                const unsigned long long target = (grand * g + world - 1) / world;
                // the prefix crosses the target inside partition p: cut at whichever of its two ends is nearer
                if (run < target && run + w >= target) cut[g] = (run + w - target <= target - run) ? p + 1u : p;
            }
            run += w;
        }
    }
    if (threadIdx.x == 0 && grand == 0) {
        for (uint32_t g = 1; g < world; ++g) cut[g] = (uint32_t)(((unsigned long long)nparts * g) / world);
    }
}

__device__ __forceinline__ uint32_t owner_of(uint32_t p, const uint32_t *cut, uint32_t world) {
    uint32_t g = 0;
    while (g + 1 < world && p >= cut[g + 1]) ++g;
    return g;
}

// Layout of one side of the exchange for this rank's virtual ranks [v0, v0 + nv) out of `vtot`:
//   dst_start[c][p]  where the segment of partition p of my virtual rank v0 + c starts in owner(p)'s receive
//                    buffer (partition-major, virtual-rank-minor: an owned partition is contiguous)
//   own_total[p]     global size of p if this rank owns it, else 0 (the histogram the local join runs on; all
//                    zero when the rows this rank receives exceed `cap`: the join then reads nothing, the
//                    exchange drops what does not fit and need[1] = 1 fails the step)
//   need[0]          rows this rank receives
// One CTA.
template <int NT>
__global__ void __launch_bounds__(NT)
exchange_layout_kernel(const uint32_t *__restrict__ hist, uint32_t vtot, uint32_t v0, uint32_t nv, uint32_t nparts,
                       const uint32_t *__restrict__ cut, uint32_t world, uint32_t rank, uint32_t cap,
                       const uint32_t *__restrict__ total, uint32_t *__restrict__ dst_start,
                       uint32_t *__restrict__ own_total, uint32_t *__restrict__ need, uint32_t *__restrict__ error) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    __shared__ uint32_t owner_base[kMaxPeers + 1];
    const uint32_t per   = (nparts + NT - 1) / NT;
    const uint32_t first = threadIdx.x * per;
    uint32_t       s     = 0;
    for (uint32_t k = 0; k < per; ++k)
        if (first + k < nparts) s += total[first + k];
    const uint32_t start = block_exclusive_scan<NT>(s, warp_sums);
    const uint32_t grand = warp_sums[NT / 32];
    uint32_t       run   = start;
    // owner_base[g] = rows of all partitions before cut[g]: where owner g's receive buffer starts in the global order
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t p = first + k;
        if (p < nparts) {
            for (uint32_t g = 0; g < world; ++g)
                if (cut[g] == p) owner_base[g] = run;
            run += total[p];
        }
    }
    if (threadIdx.x == 0)
        for (uint32_t g = 0; g <= world; ++g)
            if (cut[g] >= nparts) owner_base[g] = grand;   // empty tail ranges, and the end sentinel
    __syncthreads();
    const uint32_t mine = owner_base[rank + 1] - owner_base[rank];
    const bool     over = mine > cap;
    run = start;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t p = first + k;
        if (p < nparts) {
            const uint32_t o = owner_of(p, cut, world);
            uint32_t before = 0;
            for (uint32_t v = 0; v < v0; ++v) before += hist[(size_t)v * nparts + p];
            for (uint32_t c = 0; c < nv; ++c) {
                dst_start[(size_t)c * nparts + p] = run - owner_base[o] + before;
                before += hist[(size_t)(v0 + c) * nparts + p];
            }
            own_total[p] = (o == rank && !over) ? total[p] : 0u;
            run += total[p];
        }
    }
    if (threadIdx.x == 0) {
        need[0] = mine;
        need[1] = over ? 1u : 0u;
        if (over) *error = 2u;
    }
    (void)vtot;
}

// The exchange of one staged chunk: element e of partition p (src_off[p] <= e < src_off[p + 1]) goes to owner(p)'s
// receive buffer at dst_start[p] + (e - src_off[p]).  Flat decomposition (a CTA takes 2048 consecutive staged
// tuples whatever partition they belong to: a skewed partition is copied by as many CTAs as it has kilo-tuples);
// consecutive lanes copy consecutive tuples, so a warp's stores are 256 contiguous bytes in the peer's memory
// (full NVLink packets) except where a partition ends.
struct ExchangeArgs2 {
    const uint64_t *src_tup;
    const uint32_t *src_off;     // [nparts + 1]
    const uint32_t *dst_start;   // [nparts]
    const uint32_t *cut;         // [world + 1]
    uint32_t        n, nparts, world, cap;
    uint64_t       *dst_tup[kMaxPeers];
};
__global__ void __launch_bounds__(256) segment_exchange2_kernel(ExchangeArgs2 x) {
    constexpr int  UN   = 8;
    const uint32_t base = blockIdx.x * (256u * UN);
    uint32_t       e    = base + threadIdx.x;
    x.n                 = min(x.n, __ldg(x.src_off + x.nparts));   // staged tuples (rows taken out as hot are not staged)
    uint32_t       lo = 0, hi = x.nparts;   // invariant: src_off[lo] <= e < src_off[hi]
    if (e < x.n) {
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(x.src_off + mid) <= e) lo = mid; else hi = mid;
        }
    }
    uint32_t p = lo;
    uint32_t cuts[kMaxPeers];   // cut[1 .. world - 1] in registers: owner(p) = number of cuts <= p
#pragma unroll
    for (int g = 0; g < kMaxPeers; ++g) cuts[g] = (uint32_t)(g + 1) < x.world ? __ldg(x.cut + g + 1) : 0xFFFFFFFFu;
    uint64_t t[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const uint32_t eu = e + (uint32_t)u * 256u;
        t[u]              = eu < x.n ? ld_stream_u64(x.src_tup + eu) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u, e += 256u) {
        if (e >= x.n) break;
        while (__ldg(x.src_off + p + 1) <= e) ++p;   // empty partitions are skipped as well
        uint32_t d = 0;
#pragma unroll
        for (int g = 0; g < kMaxPeers; ++g) d += p >= cuts[g] ? 1u : 0u;
        const uint32_t pos = __ldg(x.dst_start + p) + (e - __ldg(x.src_off + p));
        if (pos < x.cap) x.dst_tup[d][pos] = t[u];   // too small a receive buffer is flagged by the layout kernel
    }
}

}  // namespace b200
