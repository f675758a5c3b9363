// kernels.cuh — hand-written sm_100a kernels of the join hot path.
//
// Reference loops replaced (file:line in /root/reference):
//   K1  filter.c:115-170            -> scan_filter_kernel
//   K2  inter_res.c:200-204,223-227 -> fused away: KeySrc is a lazy (column,
//                                      row-id list) view, never an AoS copy
//   K3  preprocess.c:189-192        -> radix_hist_kernel
//   K4  preprocess.c:83-102         -> partition_plan_kernel
//   K5  preprocess.c:262-296,350-359-> radix_scatter_kernel (+ its instances: OPT
//                                      histogram-free, CARRY a SUM value in the
//                                      row-id slot, PRED filter predicates / hot
//                                      keys evaluated in the load stage, TUPIN
//                                      packed-tuple input), radix_scatter_pay_kernel
//   K6  rhjoin.c:227-248,270-271    -> tag_join_kernel (both key widths; K64
//   K7  rhjoin.c:154-216               verifies candidates, SEG reads a build side
//                                      made of one run per source GPU) /
//                                      hash_join_kernel (small unpartitioned
//                                      builds): build, probe
//   K8  inter_res.c:79-98,119-137;
//       filter.c:60-76; inter_res.c:304-313 -> gather_columns_kernel
//   K9  inter_res.c:332-333         -> checksum_kernel / the MODE_SUM instances
//                                      of the join kernels
//   K11 inter_res.c:376-385         -> inter_equal_kernel
//   loader statistics relation_map.c:53-83 -> column_minmax_kernel,
//                                      column_mark_kernel, bitmap_count_kernel
//   multi-GPU (absent in the reference): segment_offsets_kernel, coarse_hist_kernel
//                                      here; flags, pushes, hot keys, ownership
//                                      cuts, exchange and fetch kernels in
//                                      multi_kernels.cuh.  build_cursors_kernel,
//                                      exchange_cursors_kernel and
//                                      segment_exchange_kernel serve the staged
//                                      entry points (b200_stage_*) of round 1
//
// Integer/byte work only: no tensor cores.  Row ids and positions are 32-bit
// on the device; keys are 32-bit when the column maxima allow it (8-byte
// partition tuples) and 64-bit otherwise (16-byte tuples).  What bounds each
// kernel (HBM, or the SM's L1/shared-memory pipe) is in DESIGN.md §4-5.
#pragma once

#include "types.cuh"
#include "../../include/b200_synth.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

namespace b200 {

constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr uint32_t kEmpty16  = 0xFFFFu;

__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const uint64_t *p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];"
                 : "=l"(v.x), "=l"(v.y)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- predicates folded into the load stage (types.cuh PredSet; filter.c:115-170's comparisons) ----
__device__ __forceinline__ bool pred_holds(int cmp, uint64_t v, uint64_t k) {
    return cmp == 0 ? v < k : cmp == 1 ? v > k : v == k;
}
__device__ __forceinline__ bool preds_hold(const PredSet &ps, const uint64_t (&v)[kMaxPredCols]) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < kMaxPred; ++i) {
        if (i < ps.npred) {
            const int      c  = ps.p[i].col;
            const uint64_t vv = c == 0 ? v[0] : c == 1 ? v[1] : v[2];
            ok                = ok && pred_holds(ps.p[i].cmp, vv, ps.p[i].k);
        }
    }
    return ok;
}
__device__ __forceinline__ bool preds_hold_row(const PredSet &ps, uint64_t row) {
    uint64_t v[kMaxPredCols];
#pragma unroll
    for (int c = 0; c < kMaxPredCols; ++c) v[c] = c < ps.ncols ? ld_stream_u64(ps.col[c] + row) : 0ull;
    return preds_hold(ps, v);
}

// Exclusive block scan of one value per thread.  warp_sums needs NT/32+1
// slots; slot NT/32 receives the block total.  Ends with a barrier so the
// scratch can be reused immediately.
template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums) {
    constexpr int NW   = NT / 32;
    const int     lane = threadIdx.x & 31;
    const int     wid  = threadIdx.x >> 5;
    uint32_t      incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w  = lane < NW ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(kFullMask, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < NW) warp_sums[lane] = wi - w;
        if (lane == 31) warp_sums[NW] = wi;
    }
    __syncthreads();
    uint32_t res = warp_sums[wid] + incl - v;
    __syncthreads();
    return res;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

// ---------------------------------------------------------------------------
// Tile loader shared by the histogram and scatter kernels.  A tile is NT*U
// consecutive positions of the key vector.  For a full tile of a base column
// whose pointer is 16-byte aligned the loads are 128-bit (two keys each);
// local position of register j is then ((j/2)*NT + tid)*2 + (j&1), otherwise
// j*NT + tid.  Either way a warp touches consecutive memory.
// ---------------------------------------------------------------------------
template <int NT, int U>
__device__ __forceinline__ uint32_t tile_local_index(int j, bool vec) {
    return vec ? ((uint32_t)((j >> 1) * NT + (int)threadIdx.x) * 2u + (uint32_t)(j & 1))
               : (uint32_t)(j * NT + (int)threadIdx.x);
}

// FOLD (32-bit-key instances over a genuine column, never over packed tuples): see below.
template <int NT, int U, typename KeyT, bool FOLD = false>
__device__ __forceinline__ void load_tile_keys(const KeySrc &src, uint64_t base, uint32_t count,
                                               bool vec, KeyT (&keys)[U]) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < U; j += 2) {
            const uint32_t li = ((uint32_t)((j >> 1) * NT + (int)threadIdx.x)) * 2u;
            ulonglong2     v  = ld_stream_u64x2(src.col + base + li);
            if constexpr (sizeof(KeyT) == 4 && FOLD) {
                // 32-bit-key instances run only on columns whose maximum is < 2^32, so the high halves are zero;
                // folding them in keeps ptxas from narrowing the 128-bit load into two 32-bit loads (which costs
                // twice the L1 wavefronts in kernels bound by that pipe)
                keys[j]     = (KeyT)v.x ^ (KeyT)(v.x >> 32);
                keys[j + 1] = (KeyT)v.y ^ (KeyT)(v.y >> 32);
            } else {
                keys[j]     = (KeyT)v.x;
                keys[j + 1] = (KeyT)v.y;
            }
        }
    } else if (src.ids == nullptr) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            keys[j]           = li < count ? (KeyT)ld_stream_u64(src.col + base + li) : (KeyT)0;
        }
    } else {
        uint32_t id[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            id[j]             = li < count ? ld_stream_u32(src.ids + base + li) : 0u;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            keys[j]           = li < count ? (KeyT)__ldg(src.col + id[j]) : (KeyT)0;
        }
    }
}

// ---------------------------------------------------------------------------
// K3: histogram of key & mask (preprocess.c:189-192 with N_LSB = radix bits).
// Shared-memory bins per CTA, one global atomicAdd per non-empty bin at exit.
// ---------------------------------------------------------------------------
// PRED: rows that fail the predicates are not counted (they will not be scattered either).
template <int NT, int U, typename KeyT, bool PRED = false>
__global__ void __launch_bounds__(NT)
radix_hist_kernel(KeySrc src, uint32_t radix_bits, uint32_t *__restrict__ ghist, const PredSet ps) {
    extern __shared__ uint32_t sh_hist[];
    constexpr uint32_t TILE  = NT * U;
    const uint32_t     nbins = 1u << radix_bits;
    const uint32_t     mask  = nbins - 1u;
    for (uint32_t b = threadIdx.x; b < nbins; b += NT) sh_hist[b] = 0;
    __syncthreads();
    const uint64_t n      = src.n;
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    const bool     vec_ok = src.ids == nullptr && ((reinterpret_cast<uintptr_t>(src.col) & 15) == 0);
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t base  = tile * TILE;
        const uint32_t count = (uint32_t)min((uint64_t)TILE, n - base);
        const bool     vec   = vec_ok && count == TILE;
        KeyT           keys[U];
        load_tile_keys<NT, U, KeyT>(src, base, count, vec, keys);
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = tile_local_index<NT, U>(j, vec);
            bool           on = li < count;
            if constexpr (PRED) on = on && preds_hold_row(ps, base + li);
            if (on) atomicAdd(&sh_hist[(uint32_t)keys[j] & mask], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += NT) {
        const uint32_t c = sh_hist[b];
        if (c) atomicAdd(&ghist[b], c);
    }
}

// ---------------------------------------------------------------------------
// K4: merge + exclusive prefix sum (preprocess.c:83-102) for both sides, the
// scatter cursors, and the virtual work list of the join kernel: partition p
// contributes ceil(b_p/cap) * ceil(p_p/slice) items when both sides are
// non-empty (rhjoin.c:31-34 counts the same bucket pairs).  One CTA.
// ---------------------------------------------------------------------------
// opt_cap > 0 selects the histogram-free probe side (see radix_scatter_kernel,
// OPT): partition p owns the fixed region [p*opt_cap, (p+1)*opt_cap) and
// `hist_p` is the cursor array the scatter left behind, so the number of tuples
// that landed in the region is min(cursor - p*opt_cap, opt_cap).
// cnt_p[p] always receives the probe-side count of partition p.
__device__ __forceinline__ uint32_t probe_count(const uint32_t *hist_p, uint32_t b, uint32_t opt_cap) {
    return opt_cap ? min(hist_p[b] - b * opt_cap, opt_cap) : hist_p[b];
}
template <int NT>
__global__ void __launch_bounds__(NT)
partition_plan_kernel(const uint32_t *__restrict__ hist_b, const uint32_t *__restrict__ hist_p,
                      uint32_t nparts, uint32_t cap, uint32_t slice, uint32_t *__restrict__ off_b,
                      uint32_t *__restrict__ off_p, uint32_t *__restrict__ cur_b,
                      uint32_t *__restrict__ cur_p, uint32_t *__restrict__ item_start,
                      uint32_t *__restrict__ cnt_p, uint32_t opt_cap) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    const uint32_t      per   = (nparts + NT - 1) / NT;
    const uint32_t      first = threadIdx.x * per;
    uint32_t            sb = 0, sp = 0, si = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            const uint32_t cb = hist_b[b], cp = probe_count(hist_p, b, opt_cap);
            sb += cb;
            sp += cp;
            if (cb && cp) si += ((cb + cap - 1) / cap) * ((cp + slice - 1) / slice);
        }
    }
    uint32_t eb = block_exclusive_scan<NT>(sb, warp_sums);
    const uint32_t tb = warp_sums[NT / 32];
    __syncthreads();
    uint32_t ep = block_exclusive_scan<NT>(sp, warp_sums);
    const uint32_t tp = warp_sums[NT / 32];
    __syncthreads();
    uint32_t ei = block_exclusive_scan<NT>(si, warp_sums);
    const uint32_t ti = warp_sums[NT / 32];
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            const uint32_t cb = hist_b[b], cp = probe_count(hist_p, b, opt_cap);
            off_b[b] = eb;
            cur_b[b] = eb;
            off_p[b] = opt_cap ? b * opt_cap : ep;
            if (!opt_cap) cur_p[b] = ep;   // in OPT mode hist_p may alias cur_p
            cnt_p[b] = cp;
            item_start[b] = ei;
            eb += cb;
            ep += cp;
            if (cb && cp) ei += ((cb + cap - 1) / cap) * ((cp + slice - 1) / slice);
        }
    }
    if (threadIdx.x == 0) {
        off_b[nparts]      = tb;
        off_p[nparts]      = tp;
        item_start[nparts] = ti;
    }
}

// ---------------------------------------------------------------------------
// K5: scatter into the partition-contiguous copy (the counting-sort scatter of
// preprocess.c:350-359; the reference's parallel variant re-scans the input
// once per bucket, preprocess.c:262-296).
//
// Per tile of NT*U keys: (1) the keys are already in registers (prefetched
// while the previous tile was being copied out), (2) shared-memory atomics
// give every tuple its rank inside its partition, (3) a block scan turns the
// tile histogram into local offsets and reserves the tile's run in every
// partition with ONE global atomicAdd per non-empty partition, (4) tuples are
// written to shared memory in partition order, (5) copied out so that a warp
// stores consecutive addresses inside each run.  All CTAs advance the same
// 2^bits cursors, so the write frontier is a few hundred KB and partially
// written sectors are completed in L2 before they reach HBM (ncu: DRAM write
// bytes = 1.0 x the tuple bytes).
// The kernel is bound by the SM's load/store pipe, not by HBM (ncu: l1tex is
// the busiest unit), so full, aligned tiles run a predicate-free instance
// (FULL) and ranks are packed two per register.
// Order inside a partition is not the reference's (stable) order; only the
// multiset matters downstream.
// ---------------------------------------------------------------------------
// OPT = histogram-free ("optimistic") partitioning of the probe side: partition
// b owns the fixed region [b*opt_cap, (b+1)*opt_cap) of `out`, sized a few
// percent above the uniform expectation, and its cursor starts at b*opt_cap.
// Tuples whose position falls beyond the region are appended to `ov_out`
// (one global counter) and are partitioned exactly, with a histogram, in a
// second pass over that (normally empty) overflow only.  This removes the
// histogram read of the whole probe column (8 B/row) from the common case.
struct OptArgs {
    uint32_t        opt_cap;
    uint32_t       *ov_cursor;
    void           *ov_out;
    const uint64_t *carry_col;   // CARRY instances only
    PredSet         pred;        // PRED instances only
    uint64_t        out_rows = 0;   // capacity of `out` in tuples when the caller knows it (checked build only)
    const uint32_t *n_dev    = nullptr;   // when set: the number of input rows, read on the device (src.n is its bound)
};
// CARRY: the row-id slot of a tuple carries (uint32)carry_col[row] instead of the row id (a SUM column whose
// values fit 32 bits travels inside the tuple: the probe side of the multi-GPU exchange plan).  The column is
// read where the tuples are staged; its tile was pulled into L2 (prefetch.global.L2, one 128-byte line per
// thread) while the previous tile was processed, so those loads do not wait on DRAM.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int NT, int U>
__device__ __forceinline__ void prefetch_carry_tile(const uint64_t *col, uint64_t base, uint32_t count) {
    for (uint32_t e = threadIdx.x * 16u; e < count; e += NT * 16u) prefetch_l2(col + base + e);
}
// direct-mapped table of hot keys (multi-GPU exchange plan, multi_kernels.cuh): rows carrying one are skipped here
constexpr uint32_t kHotSlots = 8192;
constexpr uint32_t kHotEmpty = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t hot_slot(uint32_t key) { return (key * 0x9E3779B1u) >> (32 - 13); }
static_assert(kHotSlots == (1u << 13), "hot_slot produces 13 bits");
__device__ __forceinline__ bool key_is_hot(const PredSet &ps, bool hot_on, uint32_t key) {
    return hot_on && key != kHotEmpty && __ldg(ps.hot_keys + hot_slot(key)) == key;
}
// Shared-memory pre-filter of the scatter: 2^16 bits, one set per hot key (bit = the top 16 bits of the hash whose top
// 13 bits are the table slot).  A probe of the table is a random 4-byte global load per key — 32 sectors per warp
// instruction through a load/store pipe that is the scatter's bottleneck; the bitmap answers "not hot" for all but a
// few percent of the keys with one shared-memory load.
constexpr uint32_t kHotBitmapWords = (1u << 16) / 32;
__device__ __forceinline__ uint32_t hot_bit(uint32_t key) { return (key * 0x9E3779B1u) >> 16; }
__device__ __forceinline__ bool key_maybe_hot(const uint32_t *hotbits, uint32_t key) {
    const uint32_t h = hot_bit(key);
    return ((hotbits[h >> 5] >> (h & 31u)) & 1u) != 0u;
}

// Raw tile loads of the scatter: 64-bit values exactly as the column holds them (two per 128-bit load on the vector
// path).  They stay raw in registers across the copy-out of the previous tile and are narrowed to KeyT only at the
// top of their own tile: ncu (profiles/r1_g_probe_carry_summary.txt, source page) showed 11 % of all warp samples
// on the narrowing XOR that round 1 placed right behind the loads — every warp sat on the DRAM latency of the NEXT
// tile before it stored a single tuple of the current one.
template <int NT, int U>
__device__ __forceinline__ void load_tile_raw(const KeySrc &src, uint64_t base, uint32_t count, bool vec,
                                              uint64_t (&raw)[U]) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < U; j += 2) {
            const uint32_t   li = ((uint32_t)((j >> 1) * NT + (int)threadIdx.x)) * 2u;
            const ulonglong2 v  = ld_stream_u64x2(src.col + base + li);
            raw[j]              = v.x;
            raw[j + 1]          = v.y;
        }
    } else if (src.ids == nullptr) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            raw[j]            = li < count ? ld_stream_u64(src.col + base + li) : 0ull;
        }
    } else {
        uint32_t id[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            id[j]             = li < count ? ld_stream_u32(src.ids + base + li) : 0u;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT + (int)threadIdx.x);
            raw[j]            = li < count ? __ldg(src.col + id[j]) : 0ull;
        }
    }
}
// 32-bit-key instances run only on columns whose maximum is < 2^32, so the high halves are zero; folding them in
// keeps ptxas from narrowing the 128-bit loads into two 32-bit loads (twice the L1 wavefronts, r1 finding)
template <typename KeyT>
__device__ __forceinline__ KeyT narrow_key(uint64_t v) {
    if constexpr (sizeof(KeyT) == 4) return (KeyT)((uint32_t)v ^ (uint32_t)(v >> 32));
    else return (KeyT)v;
}

// PRED (never FULL): rows that fail opt.pred are skipped — a filtered base relation feeds the join without a row-id
// list, a host round trip or a compaction gather.  The predicate columns of a tile are pulled into L2
// one tile ahead like a carried column and read with 128-bit loads at the top of the tile.
// TUPIN: the input is an array of packed 32-bit-key tuples {key32, slot32} (the first pass of a two-pass partition):
// the key is the low half of the raw value, the slot travels on unchanged.
template <int NT, int U, typename KeyT, bool FULL, bool OPT, bool CARRY = false, bool PRED = false, bool TUPIN = false>
__device__ __forceinline__ void scatter_tile(const KeySrc &src, uint64_t (&raw)[U], uint64_t base, uint32_t count,
                                             bool vec, uint64_t nbase, uint32_t ncount, bool nvec, bool has_next,
                                             uint32_t nbins, uint32_t mask, uint32_t per,
                                             typename TupOf<KeyT>::type *stage, uint32_t *cnt, uint32_t *loc,
                                             uint32_t *gdelta, uint32_t *warp_sums, uint32_t *__restrict__ cursor,
                                             typename TupOf<KeyT>::type *__restrict__ out, const OptArgs &opt,
                                             uint32_t *ovdelta, uint32_t *s_over, const uint32_t *hotbits = nullptr) {
    using TupT = typename TupOf<KeyT>::type;
    constexpr int  NW   = NT / 32;
    constexpr int  PER  = 2;   // bins per thread whose reservation stays in flight across the staging phase
    const uint32_t tid  = threadIdx.x;
    const uint32_t lane = tid & 31u, wid = tid >> 5;
    KeyT           keys[U];
#pragma unroll
    for (int j = 0; j < U; ++j) keys[j] = TUPIN ? (KeyT)(uint32_t)raw[j] : narrow_key<KeyT>(raw[j]);
    static_assert(U <= 64, "one validity bit per key");
    static_assert(!TUPIN || (sizeof(KeyT) == 4 && !CARRY && !PRED && !OPT), "packed tuples: plain 32-bit-key instance");
    [[maybe_unused]] uint64_t valid = 0;   // PRED: bit j = row of register j passes every predicate
    if constexpr (PRED) {
        const PredSet &ps     = opt.pred;
        const bool     hot_on = ps.hot_keys != nullptr && __ldg(ps.hot_n) != 0u;
        if (vec) {
            // One pass per predicate column: all U/2 128-bit loads of the column are issued back to back and only
            // then compared, so a tile exposes ONE memory latency per column.  (Evaluating row by row — every
            // column of rows j, j+1, then the comparison — exposed U/2 of them in sequence: 2.06 ms for config 3's
            // 200 M-row fact relation with two predicate columns, 12 K tiles of 47 K cycles each.)
            valid = U == 64 ? ~0ull : (1ull << U) - 1ull;
#pragma unroll
            for (int cc = 0; cc < kMaxPredCols; ++cc) {
                if (cc < ps.ncols) {
                    ulonglong2 t[U / 2];
#pragma unroll
                    for (int j = 0; j < U; j += 2)
                        t[j >> 1] = ld_stream_u64x2(ps.col[cc] + base + ((uint32_t)((j >> 1) * NT) + tid) * 2u);
#pragma unroll
                    for (int i = 0; i < kMaxPred; ++i) {
                        if (i < ps.npred && ps.p[i].col == cc) {
                            const int      cmp = ps.p[i].cmp;
                            const uint64_t k   = ps.p[i].k;
#pragma unroll
                            for (int j = 0; j < U; j += 2) {
                                if (!pred_holds(cmp, t[j >> 1].x, k)) valid &= ~(1ull << j);
                                if (!pred_holds(cmp, t[j >> 1].y, k)) valid &= ~(1ull << (j + 1));
                            }
                        }
                    }
                }
            }
            if (hot_on) {
#pragma unroll
                for (int j = 0; j < U; ++j)
                    if (((valid >> j) & 1ull) && key_maybe_hot(hotbits, (uint32_t)keys[j]) && key_is_hot(ps, true, (uint32_t)keys[j]))
                        valid &= ~(1ull << j);
            }
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t li = (uint32_t)(j * NT) + tid;
                if (li < count && preds_hold_row(ps, base + li) &&
                    !(hot_on && key_maybe_hot(hotbits, (uint32_t)keys[j]) && key_is_hot(ps, true, (uint32_t)keys[j])))
                    valid |= 1ull << j;
            }
        }
    }
    auto row_on = [&](int j, uint32_t li) -> bool {
        if constexpr (PRED) return ((valid >> j) & 1ull) != 0ull;
        else if constexpr (FULL) return true;
        else return li < count;
    };
    uint32_t rank2[(U + 1) / 2];   // two 16-bit ranks per register
#pragma unroll
    for (int j = 0; j < (U + 1) / 2; ++j) rank2[j] = 0;
    // ---- (1) rank of every tuple inside its partition: shared-memory atomics ----
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const uint32_t li = FULL ? ((uint32_t)((j >> 1) * NT) + tid) * 2u + (uint32_t)(j & 1)
                                 : tile_local_index<NT, U>(j, vec);
        if (row_on(j, li)) {
            const uint32_t r = atomicAdd(&cnt[(uint32_t)keys[j] & mask], 1u);
            rank2[j >> 1] |= r << (16 * (j & 1));
        }
    }
    __syncthreads();   // (A) tile histogram complete
    // ---- (2) block scan of the tile histogram with ONE barrier; the tile's run in every partition is reserved
    //          with one global atomic per non-empty partition whose result is only consumed after the staging
    //          phase (round 1 waited for it here: 10 % of the warp samples) ----
    const uint32_t first = tid * per;
    uint32_t       s = 0;
    for (uint32_t k = 0; k < per; ++k)
        if (first + k < nbins) s += cnt[first + k];
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    if (OPT && tid == 0) *s_over = 0u;
    __syncthreads();   // (B) warp totals visible (every warp scans them redundantly: no second barrier)
    uint32_t wbase = 0, staged = 0;
    {
        const uint32_t w = lane < (uint32_t)NW ? warp_sums[lane] : 0u;
        uint32_t       wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        wbase = __shfl_sync(kFullMask, wi - w, (int)wid);
        staged = __shfl_sync(kFullMask, wi, NW - 1);   // tuples of this tile (< count when rows were filtered out)
    }
    // global position of stage slot i of partition b is gdelta[b] + i; a run that leaves its OPT region goes to
    // the overflow array at ovdelta[b] + i
    auto finalize = [&](uint32_t b, uint32_t cb, uint32_t oldv, uint32_t runv) {
        gdelta[b] = oldv - runv;
        if constexpr (OPT) {
            const uint32_t lim = (b + 1u) * opt.opt_cap;
            if (oldv + cb > lim) {   // part of this run does not fit the region any more
                const uint32_t from = max(oldv, lim);
                const uint32_t ovd  = atomicAdd(opt.ov_cursor, oldv + cb - from);
                ovdelta[b]          = ovd - (from - (oldv - runv));
                *s_over             = 1u;
            }
        }
    };
    const uint32_t run0 = wbase + incl - s;
    uint32_t       run  = run0;
    uint32_t       old[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {   // the first PER bins of the thread: reservation in flight across the staging
        old[k] = 0;
        if ((uint32_t)k < per && first + k < nbins) {
            const uint32_t cb = cnt[first + k];
            loc[first + k]    = run;
            if (cb) old[k] = atomicAdd(&cursor[first + k], cb);
            run += cb;
        }
    }
    for (uint32_t k = PER; k < per; ++k) {   // more than PER * NT partitions: the rest synchronously
        const uint32_t b = first + k;
        if (b < nbins) {
            const uint32_t cb = cnt[b];
            loc[b]            = run;
            if (cb) finalize(b, cb, atomicAdd(&cursor[b], cb), run);
            run += cb;
            cnt[b] = 0;
        }
    }
    // the carried column of the tile (its lines were pulled into L2 one tile ahead): every load is issued here, ahead
    // of the barrier, so that ONE L2 latency is exposed per tile — ncu showed 17 % of the warp samples waiting on these
    // loads when each was issued right where its value was staged (profiles/r2_summary.txt)
    // (not in the PRED instances: with the validity mask and the predicate batches live as well the 64 registers of
    // the batch spilled — 712 bytes of spill stores per thread and tile in the exchange plan's probe scatter — so a
    // filtered scatter loads the carried value of each surviving row where it stages it)
    constexpr bool kHoistCarry = CARRY && FULL && !PRED;
    [[maybe_unused]] uint64_t craw[kHoistCarry ? U : 1];
    if constexpr (kHoistCarry) {
#pragma unroll
        for (int j = 0; j < U; j += 2) {   // registers j, j+1 hold two consecutive rows: one 128-bit load for both
            const ulonglong2 v = ld_stream_u64x2(opt.carry_col + base + ((uint32_t)((j >> 1) * NT) + tid) * 2u);
            craw[j]            = v.x;
            craw[j + 1]        = v.y;
        }
    }
    __syncthreads();   // (C) local offsets visible
    // ---- (3) tuples into shared memory in partition order ----
    // (carried values that were not hoisted — filtered and ragged tiles — are loaded eight rows at a time, then staged:
    // one exposed latency per eight rows instead of one per row)
    constexpr int SB = 8;
    static_assert(U % SB == 0, "staging batches");
#pragma unroll
    for (int j0 = 0; j0 < U; j0 += SB) {
        [[maybe_unused]] uint32_t cval[SB];
        if constexpr (CARRY && !kHoistCarry) {
#pragma unroll
            for (int u = 0; u < SB; ++u) {
                const int      j  = j0 + u;
                const uint32_t li = FULL ? ((uint32_t)((j >> 1) * NT) + tid) * 2u + (uint32_t)(j & 1)
                                         : tile_local_index<NT, U>(j, vec);
                cval[u] = row_on(j, li) ? (uint32_t)ld_stream_u64(opt.carry_col + base + li) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < SB; ++u) {
            const int      j  = j0 + u;
            const uint32_t li = FULL ? ((uint32_t)((j >> 1) * NT) + tid) * 2u + (uint32_t)(j & 1)
                                     : tile_local_index<NT, U>(j, vec);
            if (row_on(j, li)) {
                TupT t;
                t.key = keys[j];
                if constexpr (CARRY) {
                    if constexpr (kHoistCarry) t.rid = narrow_key<uint32_t>(craw[j]);
                    else t.rid = cval[u];
                } else if constexpr (TUPIN) {
                    t.rid = (uint32_t)(raw[j] >> 32);
                } else {
                    t.rid = (uint32_t)base + li;
                }
                if constexpr (sizeof(KeyT) == 8) t.pad = 0;
                const uint32_t slot = loc[(uint32_t)keys[j] & mask] + ((rank2[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                B200_DCHECK(slot < (uint32_t)(NT * U));
                stage[slot] = t;
            }
        }
    }
    // ---- (4) the reservations have arrived by now ----
    run = run0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if ((uint32_t)k < per && first + k < nbins) {
            const uint32_t cb = cnt[first + k];
            if (cb) finalize(first + k, cb, old[k], run);
            run += cb;
            cnt[first + k] = 0;   // ready for the next tile
        }
    }
    __syncthreads();   // (D) stage, gdelta (and the overflow flag) visible
    // keys and ranks are dead: put the next tile's loads in flight before the copy-out so that their latency hides
    // behind the stores (they are narrowed at the top of the next tile, not here)
    if (has_next) {
        load_tile_raw<NT, U>(src, nbase, ncount, nvec, raw);
        if constexpr (CARRY) prefetch_carry_tile<NT, U>(opt.carry_col, nbase, ncount);
        if constexpr (PRED) {
            for (int cc = 0; cc < opt.pred.ncols; ++cc) prefetch_carry_tile<NT, U>(opt.pred.col[cc], nbase, ncount);
        }
    }
    // ---- (5) copy-out: a warp stores consecutive addresses inside each run ----
    bool over = false;
    if constexpr (OPT) over = *s_over != 0u;
    if (!over) {
        auto put = [&](uint32_t i) {
            const TupT     t   = stage[i];
            const uint32_t pos = gdelta[(uint32_t)t.key & mask] + i;
            if constexpr (OPT) B200_DCHECK(pos < ((((uint32_t)t.key & mask) + 1u) * opt.opt_cap));
            B200_DCHECK(opt.out_rows == 0 || pos < opt.out_rows);
            out[pos] = t;
        };
        if constexpr (FULL && !PRED) {
#pragma unroll
            for (int k = 0; k < U; ++k) put((uint32_t)(k * NT) + tid);
        } else {
            for (uint32_t i = tid; i < staged; i += NT) put(i);
        }
    } else {
        if constexpr (OPT) {
            for (uint32_t i = tid; i < staged; i += NT) {
                const TupT     t   = stage[i];
                const uint32_t b   = (uint32_t)t.key & mask;
                const uint32_t pos = gdelta[b] + i;
                if (pos >= (b + 1u) * opt.opt_cap) {
                    B200_DCHECK((uint32_t)(ovdelta[b] + i) < src.n);   // the overflow array holds src.n tuples (32-bit wrap-around arithmetic)
                    static_cast<TupT *>(opt.ov_out)[ovdelta[b] + i] = t;
                } else {
                    B200_DCHECK(pos >= b * opt.opt_cap);
                    out[pos] = t;
                }
            }
        }
    }
    __syncthreads();   // (E) stage and bin arrays free for the next tile
}

template <int NT, int U, int MINB, typename KeyT, bool OPT, bool CARRY = false, bool PRED = false, bool TUPIN = false>
__global__ void __launch_bounds__(NT, MINB)
radix_scatter_kernel(KeySrc src, uint32_t radix_bits, uint32_t *__restrict__ cursor,
                     typename TupOf<KeyT>::type *__restrict__ out, const OptArgs opt) {
    using TupT = typename TupOf<KeyT>::type;
    constexpr uint32_t TILE = NT * U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TupT     *stage  = reinterpret_cast<TupT *>(smem_raw);
    uint32_t *cnt    = reinterpret_cast<uint32_t *>(stage + TILE);
    const uint32_t nbins = 1u << radix_bits;
    const uint32_t mask  = nbins - 1u;
    uint32_t *loc     = cnt + nbins;
    uint32_t *gdelta  = loc + nbins;
    uint32_t *ovdelta = gdelta + nbins;   // OPT only (the launch reserves 4 bin arrays then)
    [[maybe_unused]] uint32_t *hotbits = gdelta + (OPT ? 2u : 1u) * nbins;   // PRED with a hot-key table: kHotBitmapWords more
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    __shared__ uint32_t s_over;

    // row ids are 32-bit: tile bases fit 32 bits as well
    const uint64_t n      = opt.n_dev ? (uint64_t)*opt.n_dev : src.n;   // (a count only the device knows)
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    bool           vec_ok = src.ids == nullptr && ((reinterpret_cast<uintptr_t>(src.col) & 15) == 0) &&
                  (!CARRY || (reinterpret_cast<uintptr_t>(opt.carry_col) & 15) == 0);
    if constexpr (PRED) {
        for (int cc = 0; cc < opt.pred.ncols; ++cc) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(opt.pred.col[cc]) & 15) == 0;
    }
    const uint32_t per    = (nbins + NT - 1) / NT;

    for (uint32_t b = threadIdx.x; b < nbins; b += NT) cnt[b] = 0;
    if constexpr (PRED) {
        if (opt.pred.hot_keys != nullptr) {   // (the launch reserved the bitmap)
            for (uint32_t w = threadIdx.x; w < kHotBitmapWords; w += NT) hotbits[w] = 0u;
            __syncthreads();
            for (uint32_t sl = threadIdx.x; sl < kHotSlots; sl += NT) {
                const uint32_t k = __ldg(opt.pred.hot_keys + sl);
                if (k != kHotEmpty) atomicOr(&hotbits[hot_bit(k) >> 5], 1u << (hot_bit(k) & 31u));
            }
        }
    }
    uint64_t raw[U];
    uint64_t tile = blockIdx.x;
    if (tile < ntiles) {
        const uint64_t base  = tile * TILE;
        const uint32_t count = (uint32_t)min((uint64_t)TILE, n - base);
        load_tile_raw<NT, U>(src, base, count, vec_ok && count == TILE, raw);
        if constexpr (CARRY) prefetch_carry_tile<NT, U>(opt.carry_col, base, count);
        if constexpr (PRED) {
            for (int cc = 0; cc < opt.pred.ncols; ++cc) prefetch_carry_tile<NT, U>(opt.pred.col[cc], base, count);
        }
    }
    __syncthreads();

    for (; tile < ntiles; tile += gridDim.x) {
        const uint64_t base     = tile * TILE;
        const uint32_t count    = (uint32_t)min((uint64_t)TILE, n - base);
        const bool     vec      = vec_ok && count == TILE;
        const uint64_t ntile    = tile + gridDim.x;
        const bool     has_next = ntile < ntiles;
        const uint64_t nbase    = ntile * TILE;
        const uint32_t ncount   = has_next ? (uint32_t)min((uint64_t)TILE, n - nbase) : 0u;
        const bool     nvec     = vec_ok && ncount == TILE;
        if constexpr (PRED) {
            if (vec)
                scatter_tile<NT, U, KeyT, true, OPT, CARRY, true>(src, raw, base, count, vec, nbase, ncount, nvec, has_next,
                                                                  nbins, mask, per, stage, cnt, loc, gdelta, warp_sums,
                                                                  cursor, out, opt, ovdelta, &s_over, hotbits);
            else
                scatter_tile<NT, U, KeyT, false, OPT, CARRY, true>(src, raw, base, count, vec, nbase, ncount, nvec, has_next,
                                                                   nbins, mask, per, stage, cnt, loc, gdelta, warp_sums,
                                                                   cursor, out, opt, ovdelta, &s_over, hotbits);
        } else if (vec) {
            scatter_tile<NT, U, KeyT, true, OPT, CARRY, false, TUPIN>(src, raw, base, count, vec, nbase, ncount, nvec, has_next,
                                                        nbins, mask, per, stage, cnt, loc, gdelta, warp_sums, cursor,
                                                        out, opt, ovdelta, &s_over);
        } else {
            scatter_tile<NT, U, KeyT, false, OPT, CARRY, false, TUPIN>(src, raw, base, count, vec, nbase, ncount, nvec, has_next,
                                                         nbins, mask, per, stage, cnt, loc, gdelta, warp_sums, cursor,
                                                         out, opt, ovdelta, &s_over);
        }
    }
}

// Two-pass partition: histogram over the low `cbits` of the partition index from the fine histogram
// (coarse[c] = sum of fine[p] over p with p & (2^cbits - 1) == c).  One CTA.
static __global__ void __launch_bounds__(1024)
coarse_hist_kernel(const uint32_t *__restrict__ fine, uint32_t nparts, uint32_t cbits, uint32_t *__restrict__ coarse) {
    const uint32_t nc = 1u << cbits;
    for (uint32_t c = threadIdx.x; c < nc; c += 1024) {
        uint32_t sum = 0;
        for (uint32_t p = c; p < nparts; p += nc) sum += fine[p];
        coarse[c] = sum;
    }
}

// Build-side scatter that also carries up to NPAY projected columns into
// partition order (early materialisation): the SUM of a build-side projection
// then reads pay_out[k][position in the partition buffer] — a dense, L2-resident
// window — instead of one random DRAM gather per match (measured on B200:
// random 8-byte gathers run at ~40 G/s whatever their L2 fetch size).  The build
// side is the small relation, so this kernel is the plain variant of
// radix_scatter_kernel: no prefetch, no predicate-free instance.
constexpr int kMaxPeers = 8;
struct PayArgs {
    const uint64_t *col[2];
    const uint32_t *ids[2];
    uint64_t       *out[2];
    // multi-GPU broadcast-scatter: when ndst > 0 every tuple (and payload) is
    // stored into ndst destination buffers with the same layout — this GPU's
    // own and its peers' (CUDA IPC mappings, written over NVLink) — instead of
    // `out`; row ids are rid_base + position so they stay global
    int       ndst;
    int       carry32;   // store (uint32)col[0][row] in the tuple's row-id slot instead of the row id
    uint32_t  rid_base;
    void     *tup_dst[kMaxPeers];
    uint64_t *pay_dst[2][kMaxPeers];
};
// TUPIN: the input "column" is an array of packed 32-bit-key tuples {key32, rid32}
// (the overflow of an OPT scatter); the row id comes from the tuple, not the position.
// CARRY: the tuple's row-id slot carries (uint32)pay.col[0][row] — one build-side SUM column whose values fit 32
// bits needs no separate payload array at all (and half the bytes on the wire in the multi-GPU broadcast).
template <int NT, int U, typename KeyT, int NPAY, bool TUPIN, bool CARRY = false>
__global__ void __launch_bounds__(NT)
radix_scatter_pay_kernel(KeySrc src, uint32_t radix_bits, uint32_t *__restrict__ cursor,
                         typename TupOf<KeyT>::type *__restrict__ out, PayArgs pay) {
    using TupT = typename TupOf<KeyT>::type;
    constexpr uint32_t TILE = NT * U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TupT     *stage  = reinterpret_cast<TupT *>(smem_raw);
    uint64_t *pstage = reinterpret_cast<uint64_t *>(stage + TILE);   // [NPAY][TILE]
    uint32_t *cnt    = reinterpret_cast<uint32_t *>(pstage + (size_t)NPAY * TILE);
    const uint32_t nbins = 1u << radix_bits;
    const uint32_t mask  = nbins - 1u;
    uint32_t *loc    = cnt + nbins;
    uint32_t *gdelta = loc + nbins;
    __shared__ uint32_t warp_sums[NT / 32 + 1];

    const uint64_t n      = src.n;
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    const uint32_t per    = (nbins + NT - 1) / NT;
    for (uint32_t b = threadIdx.x; b < nbins; b += NT) cnt[b] = 0;

    // keys (and payload values) of a tile live in registers; the next tile's are loaded while the
    // current one is copied out
    KeyT     keys[U];
    uint32_t in_rid[TUPIN ? U : 1];
    uint64_t pvals[NPAY > 0 ? NPAY : 1][U];
    uint32_t cvals[CARRY ? U : 1];
    auto load_tile = [&](uint64_t tile) {
        const uint64_t base  = tile * TILE;
        const uint32_t count = (uint32_t)min((uint64_t)TILE, n - base);
        if constexpr (TUPIN) {
            uint64_t raw[U];
            load_tile_keys<NT, U, uint64_t>(src, base, count, false, raw);
#pragma unroll
            for (int j = 0; j < U; ++j) {
                keys[j]   = (KeyT)(uint32_t)raw[j];
                in_rid[j] = (uint32_t)(raw[j] >> 32);
            }
        } else {
            load_tile_keys<NT, U, KeyT>(src, base, count, false, keys);
        }
#pragma unroll
        for (int k = 0; k < NPAY; ++k) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t li  = (uint32_t)(j * NT) + threadIdx.x;
                const uint32_t rid = (uint32_t)base + li;
                pvals[k][j]        = li < count ? ld_stream_u64(pay.col[k] + (pay.ids[k] ? pay.ids[k][rid] : rid)) : 0ull;
            }
        }
        if constexpr (CARRY) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint32_t li  = (uint32_t)(j * NT) + threadIdx.x;
                const uint32_t rid = (uint32_t)base + li;
                cvals[j] = li < count ? (uint32_t)ld_stream_u64(pay.col[0] + (pay.ids[0] ? pay.ids[0][rid] : rid)) : 0u;
            }
        }
    };
    uint64_t tile = blockIdx.x;
    if (tile < ntiles) load_tile(tile);
    __syncthreads();
    for (; tile < ntiles; tile += gridDim.x) {
        const uint64_t base  = tile * TILE;
        const uint32_t count = (uint32_t)min((uint64_t)TILE, n - base);
        uint16_t rank[U];
#pragma unroll
        for (int j = 0; j < U; ++j)
            if ((uint32_t)(j * NT) + threadIdx.x < count)
                rank[j] = (uint16_t)atomicAdd(&cnt[(uint32_t)keys[j] & mask], 1u);
        __syncthreads();
        {
            const uint32_t first = threadIdx.x * per;
            uint32_t       s     = 0;
            for (uint32_t k = 0; k < per; ++k)
                if (first + k < nbins) s += cnt[first + k];
            uint32_t run = block_exclusive_scan<NT>(s, warp_sums);
            for (uint32_t k = 0; k < per; ++k) {
                const uint32_t b = first + k;
                if (b < nbins) {
                    const uint32_t c = cnt[b];
                    loc[b]           = run;
                    if (c) gdelta[b] = atomicAdd(&cursor[b], c) - run;
                    run += c;
                    cnt[b] = 0;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = (uint32_t)(j * NT) + threadIdx.x;
            if (li < count) {
                const uint32_t rid = (uint32_t)base + li;
                const uint32_t pos = loc[(uint32_t)keys[j] & mask] + rank[j];
                TupT           t;
                t.key = keys[j];
                if constexpr (TUPIN) t.rid = in_rid[j];
                else if constexpr (CARRY) t.rid = cvals[j];
                else t.rid = pay.rid_base + rid;
                if constexpr (sizeof(KeyT) == 8) t.pad = 0;
                stage[pos] = t;
#pragma unroll
                for (int k = 0; k < NPAY; ++k) pstage[(size_t)k * TILE + pos] = pvals[k][j];
            }
        }
        __syncthreads();
        if (tile + gridDim.x < ntiles) load_tile(tile + gridDim.x);
        for (uint32_t i = threadIdx.x; i < count; i += NT) {
            const TupT     t = stage[i];
            const uint32_t o = gdelta[(uint32_t)t.key & mask] + i;
            if (pay.ndst == 0) {
                out[o] = t;
#pragma unroll
                for (int k = 0; k < NPAY; ++k) pay.out[k][o] = pstage[(size_t)k * TILE + i];
            } else {
                for (int d = 0; d < pay.ndst; ++d) {
                    static_cast<TupT *>(pay.tup_dst[d])[o] = t;
#pragma unroll
                    for (int k = 0; k < NPAY; ++k) pay.pay_dst[k][d][o] = pstage[(size_t)k * TILE + i];
                }
            }
        }
        __syncthreads();
    }
}

// Multi-GPU broadcast of a locally partitioned build shard: CTA p
// copies this GPU's segment of partition p (tuples and up to two payload
// columns) to its place inside the GLOBAL partition layout of every
// destination buffer — this GPU's own and the peers' CUDA-IPC mappings.  The
// stores of a warp are 256 contiguous bytes, so they travel over NVLink as
// full packets (scattering 8-byte tuples straight into peer memory measured
// 134 GB/s).
struct SegCopyArgs {
    const uint64_t *src_tup;       // 8-byte tuples {key32, rid32}
    const uint64_t *src_pay[2];
    const uint32_t *src_off;       // [nparts + 1] local partition offsets
    const uint32_t *dst_start;     // [nparts] start of this GPU's segment in the global layout
    int             ndst, npay;
    int             first;         // destinations are visited first+1, first+2, ... (mod ndst): every rank starts
                                   // at a different peer, so no destination's NVLink ingress is hit by all at once
    uint64_t       *dst_tup[kMaxPeers];
    uint64_t       *dst_pay[2][kMaxPeers];
};
static __global__ void __launch_bounds__(256) segment_broadcast_kernel(const SegCopyArgs s) {
    const uint32_t p     = blockIdx.x;
    const uint32_t first = s.src_off[p];
    const uint32_t count = s.src_off[p + 1] - first;
    const uint32_t dst   = s.dst_start[p];
    constexpr int  UN    = 4;
    // every element is loaded once and stored to all destinations; 4 elements per thread in flight
    for (int a = 0; a <= s.npay; ++a) {
        const uint64_t *src = a == 0 ? s.src_tup : s.src_pay[a - 1];
        for (uint32_t i0 = threadIdx.x; i0 < count; i0 += 256 * UN) {
            uint64_t v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const uint32_t i = i0 + (uint32_t)u * 256;
                v[u]             = i < count ? ld_stream_u64(src + first + i) : 0ull;
            }
            for (int j = 1; j <= s.ndst; ++j) {
                const int d    = (s.first + j) % s.ndst;
                uint64_t *out  = (a == 0 ? s.dst_tup[d] : s.dst_pay[a - 1][d]) + dst;
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const uint32_t i = i0 + (uint32_t)u * 256;
                    if (i < count) out[i] = v[u];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K6 + K7 (+K9): per-partition build and probe in shared memory.
//
// A persistent grid pulls work items from an atomic counter.  An item is
// (partition p, build chunk, probe slice): the CTA builds a bucket-chain table
// over <= cap build tuples in shared memory (rhjoin.c:219-273 keeps 64-bit
// bucket[]/chain[] arrays in DRAM and hashes with `% prime`; here heads/next
// are 16-bit, the hash is multiplicative on the key bits above the radix) and
// streams the probe slice through it (rhjoin.c:141-217), comparing full keys,
// so duplicates on both sides yield the cross product.  Only keys and chain
// links live in shared memory; the build row id of a match is re-read from the
// partition buffer (still L2-resident).
//
// Matches are not handled inside the chain walk: the lane pushes
// (table index, probe row id) into its warp's shared-memory queue, and the
// warp drains the queue 32 entries at a time with every lane active, so the
// gathers of a drain are all in flight together and the probe loop never
// waits on them.
//   MODE_COUNT: count matches per item (sizes the pair output exactly)
//   MODE_WRITE: materialise (build row id, probe row id) pairs
//   MODE_SUM  : fold the pairs straight into the SUM checksums
//               (inter_res.c:320-339) — the final join of a query never
//               materialises its output.
// DIRECT = no partition pass: one "partition" read lazily through KeySrc;
// used when the build side already fits a shared-memory table.
// ---------------------------------------------------------------------------
enum JoinMode { MODE_COUNT = 0, MODE_WRITE = 1, MODE_SUM = 2 };
constexpr int kWarpQueue = 64;   // entries per warp; drained when >= 32

struct JoinArgs {
    KeySrc          src_b, src_p;   // DIRECT
    const void     *tup_b, *tup_p;  // partitioned
    const uint32_t *off_b, *off_p, *cnt_p, *item_start;   // probe side of partition p: [off_p[p], off_p[p] + cnt_p[p])
    uint32_t        nparts, radix_bits, cap, slice, slots_log2;
    uint32_t        n_items_direct, sc_direct;
    uint32_t       *work_counter;
    unsigned long long *total;        // matches (COUNT, SUM)
    unsigned long long *item_count;   // per item: written by COUNT, read by WRITE
    unsigned long long *out_cursor;   // WRITE
    uint32_t       *out_b, *out_p;    // WRITE
    // segmented build side (multi-GPU, rank-major layout): the build tuples of partition p are nseg runs, run r
    // holding seg_cnt[r * nparts + p] tuples from physical position seg_off[r * nparts + p]; off_b[] then
    // describes the VIRTUAL concatenation of the runs (tag_join_kernel<SEG = true>)
    int             nseg;
    const uint32_t *seg_off, *seg_cnt;
    // multi-GPU broadcast in flight (SEG only, wait_flags != nullptr): run r of a partition may be read once
    // wait_flags[c * kMaxPeers + r] >= *wait_epoch, c = the chunk of region r (chunk_rows rows each, seg_rows rows
    // per region) that holds the run's last tuple — the flags are written by rank r's copy engine behind each
    // chunk of its broadcast, so the join of the first partitions overlaps the rest of the transfer
    const uint32_t *wait_flags, *wait_epoch;
    uint32_t        chunk_rows, seg_rows;
    uint32_t       *wait_error;       // set to 1 when a flag did not arrive within the spin budget
    // DIRECT only: predicates folded into the probe-side load (a filtered base relation probing a small build side);
    // valid_p counts the probe rows that passed (an empty filter makes the whole query NULL, query.c:360-369)
    PredSet             pred_p;
    unsigned long long *valid_p;
    int             nproj;            // SUM
    int             need_brid;        // SUM: some build-side projection is gathered through the row id
    ProjDesc        proj[kMaxProj];
    unsigned long long *sums;
};

template <typename KeyT>
__device__ __forceinline__ uint32_t table_hash(KeyT key, uint32_t radix_bits, uint32_t slots_log2) {
    if constexpr (sizeof(KeyT) == 8) {
        const uint64_t x = (uint64_t)key >> radix_bits;
        return ((uint32_t)(x ^ (x >> 32)) * 0x9E3779B1u) >> (32 - slots_log2);
    } else {
        return (((uint32_t)key >> radix_bits) * 0x9E3779B1u) >> (32 - slots_log2);
    }
}

template <typename KeyT> struct ProbeTup {
    KeyT     key;
    uint32_t rid;
};

// Shared-memory bucket-chain table of hash_join_kernel (64-bit keys): a key array, a 16-bit link array
// and 16-bit chain heads — rhjoin.c:219-273's bucket[]/chain[] in shared memory.
template <typename KeyT> struct TableView;
template <> struct TableView<uint64_t> {
    uint64_t *keys;
    uint16_t *next;
    uint16_t *heads;
    uint32_t  nslots;
    __device__ __forceinline__ TableView(unsigned char *smem, uint32_t cap, uint32_t slots_log2) {
        keys   = reinterpret_cast<uint64_t *>(smem);
        next   = reinterpret_cast<uint16_t *>(keys + cap);
        heads  = next + cap;
        nslots = 1u << slots_log2;
    }
    __device__ __forceinline__ void clear(uint32_t tid, uint32_t nt) {
        for (uint32_t s = tid; s < nslots; s += nt) heads[s] = (uint16_t)kEmpty16;
    }
    __device__ __forceinline__ void insert(uint32_t i, uint64_t key, uint32_t h) {
        unsigned short *slot = reinterpret_cast<unsigned short *>(&heads[h]);
        unsigned short  old  = *slot, assumed;
        do {
            assumed = old;
            old     = atomicCAS(slot, assumed, (unsigned short)i);
        } while (old != assumed);
        put(i, key, old);
    }
    static __host__ __device__ uint32_t slots_log2_for(uint32_t cap) {
        uint32_t l = 4;
        while ((1u << l) < cap) ++l;
        return l;
    }
    __device__ __forceinline__ void put(uint32_t i, uint64_t key, uint32_t nx) {
        keys[i] = key;
        next[i] = (uint16_t)nx;
    }
    __device__ __forceinline__ void get(uint32_t i, uint64_t &key, uint32_t &nx) const {
        key = keys[i];
        nx  = next[i];
    }
    static __host__ __device__ size_t bytes(uint32_t cap, uint32_t slots_log2) {
        return (size_t)cap * 10 + ((size_t)2 << slots_log2);
    }
};

// NP = number of fused SUM projections the kernel is compiled for (>= nproj)
template <int NT, int U, typename KeyT, bool DIRECT, int MODE, int NP>
__global__ void __launch_bounds__(NT)
hash_join_kernel(const JoinArgs a) {
    using TupT = typename TupOf<KeyT>::type;
    constexpr int NW = NT / 32;
    constexpr int NPA = NP > 0 ? NP : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TableView<KeyT> tab(smem_raw, a.cap, a.slots_log2);
    const uint32_t hash_log2 = a.slots_log2;

    __shared__ uint32_t           s_item[6];   // valid, b_start, b_count, p_start, p_count, w
    __shared__ uint32_t           s_cursor;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_cnt;
    __shared__ uint32_t           wq_cnt[NW];
    __shared__ uint2              wq[MODE == MODE_COUNT ? 1 : NW][MODE == MODE_COUNT ? 1 : kWarpQueue];

    const uint32_t tid  = threadIdx.x;
    const int      lane = tid & 31;
    const int      wid  = tid >> 5;
    const TupT *tup_b = static_cast<const TupT *>(a.tup_b);
    const TupT *tup_p = static_cast<const TupT *>(a.tup_p);

    unsigned long long my_matches = 0;
    uint32_t           my_valid   = 0;   // DIRECT with predicates: probe rows that passed
    unsigned long long my_sum[NPA];
    unsigned long long pend[NPA];   // values loaded by the previous drain
#pragma unroll
    for (int k = 0; k < NPA; ++k) my_sum[k] = pend[k] = 0;
    if (lane == 0) wq_cnt[wid] = 0;

    uint32_t b_start = 0;

    // drain `take` (<= 32) queue entries [have - take, have) of this warp, one per lane
    auto drain = [&](uint32_t have, uint32_t take) {
        if constexpr (MODE != MODE_COUNT) {
            const bool     mine = (uint32_t)lane < take;
            B200_DCHECK(take <= have && have <= (uint32_t)kWarpQueue);
            const uint2    e    = wq[wid][mine ? have - take + lane : 0];
            uint32_t       brid = 0;
            if (mine) {
                if constexpr (DIRECT) brid = b_start + e.x;
                else brid = tup_b[b_start + e.x].rid;
            }
            if constexpr (MODE == MODE_SUM) {
#pragma unroll
                for (int k = 0; k < NPA; ++k) {
                    my_sum[k] += pend[k];
                    pend[k] = 0;
                    if (k < a.nproj && mine) {
                        const uint32_t r  = a.proj[k].side == 0 ? brid : e.y;
                        const uint32_t rr = a.proj[k].ids ? __ldg(a.proj[k].ids + r) : r;
                        pend[k]           = __ldg(a.proj[k].col + rr);
                    }
                }
            } else {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&s_cursor, take);
                pos = __shfl_sync(kFullMask, pos, 0);
                if (mine) {
                    a.out_b[s_base + pos + lane] = brid;
                    a.out_p[s_base + pos + lane] = e.y;
                }
            }
        }
    };

    // one match of table position `idx` with probe row `prid`
    auto on_match = [&](uint32_t idx, uint32_t prid) {
        if constexpr (MODE != MODE_COUNT) {
            const uint32_t slot = atomicAdd(&wq_cnt[wid], 1u);
            if (slot < (uint32_t)kWarpQueue) {
                wq[wid][slot] = make_uint2(idx, prid);
            } else {
                // queue full (many matches per probe): handle the match in place
                uint32_t brid;
                if constexpr (DIRECT) brid = b_start + idx;
                else brid = tup_b[b_start + idx].rid;
                if constexpr (MODE == MODE_SUM) {
#pragma unroll
                    for (int kk = 0; kk < NPA; ++kk) {
                        if (kk < a.nproj) {
                            const uint32_t r  = a.proj[kk].side == 0 ? brid : prid;
                            const uint32_t rr = a.proj[kk].ids ? __ldg(a.proj[kk].ids + r) : r;
                            my_sum[kk] += __ldg(a.proj[kk].col + rr);
                        }
                    }
                } else {
                    const uint32_t pos = atomicAdd(&s_cursor, 1u);
                    a.out_b[s_base + pos] = brid;
                    a.out_p[s_base + pos] = prid;
                }
            }
        }
    };

    // one probe tuple through the table; returns the number of matches pushed or counted
    auto probe_one = [&](KeyT key, uint32_t prid) -> uint32_t {
        uint32_t nh = 0;
        {
            uint32_t idx = tab.heads[table_hash<KeyT>(key, a.radix_bits, hash_log2)];
            while (idx != kEmpty16) {
                KeyT     k;
                uint32_t nx;
                tab.get(idx, k, nx);
                if (k == key) {
                    ++nh;
                    on_match(idx, prid);
                }
                idx = nx;
            }
        }
        return nh;
    };

    for (;;) {
        // ---- fetch one work item (warp 0) ---------------------------------
        if (tid < 32) {
            uint32_t w = 0;
            if (lane == 0) w = atomicAdd(a.work_counter, 1u);
            w = __shfl_sync(kFullMask, w, 0);
            uint32_t valid = 0, bs = 0, bc = 0, ps = 0, pc = 0;
            if constexpr (DIRECT) {
                if (w < a.n_items_direct) {
                    valid               = 1;
                    const uint32_t rch  = w / a.sc_direct;
                    const uint32_t ssl  = w % a.sc_direct;
                    bs                  = rch * a.cap;
                    bc                  = min(a.cap, a.src_b.n - bs);
                    ps                  = ssl * a.slice;
                    pc                  = min(a.slice, a.src_p.n - ps);
                }
            } else {
                const uint32_t n_items = a.item_start[a.nparts];
                if (w < n_items) {
                    valid       = 1;
                    // 32-ary search for the partition p with
                    // item_start[p] <= w < item_start[p+1]
                    uint32_t lo = 0, hi = a.nparts;
                    while (hi - lo > 1) {
                        const uint32_t step = (hi - lo + 31) / 32;
                        const uint32_t idx  = lo + (uint32_t)lane * step;
                        const bool     le   = idx < hi && a.item_start[idx] <= w;
                        const uint32_t c    = __popc(__ballot_sync(kFullMask, le));   // >= 1
                        const uint32_t nlo  = lo + (c - 1) * step;
                        hi                  = min(hi, nlo + step);
                        lo                  = nlo;
                    }
                    const uint32_t p   = lo;
                    const uint32_t k   = w - a.item_start[p];
                    const uint32_t b0  = a.off_b[p], b1 = a.off_b[p + 1];
                    const uint32_t p0  = a.off_p[p], p1 = p0 + a.cnt_p[p];
                    const uint32_t sc  = (p1 - p0 + a.slice - 1) / a.slice;
                    const uint32_t rch = k / sc, ssl = k % sc;
                    bs                 = b0 + rch * a.cap;
                    bc                 = min(a.cap, b1 - bs);
                    ps                 = p0 + ssl * a.slice;
                    pc                 = min(a.slice, p1 - ps);
                }
            }
            if (lane == 0) {
                s_item[0] = valid;
                s_item[1] = bs;
                s_item[2] = bc;
                s_item[3] = ps;
                s_item[4] = pc;
                s_item[5] = w;
                s_cnt     = 0ull;
                s_cursor  = 0u;
            }
            if constexpr (MODE == MODE_WRITE) {
                if (lane == 0 && valid) s_base = atomicAdd(a.out_cursor, a.item_count[w]);
            }
        }
        __syncthreads();
        if (s_item[0] == 0u) break;
        b_start                = s_item[1];
        const uint32_t b_count = s_item[2];
        const uint32_t p_start = s_item[3], p_count = s_item[4];
        const uint32_t item_w  = s_item[5];

        auto load_probe = [&](uint32_t li, ProbeTup<KeyT> &t) {
            if (li < p_count) {
                if constexpr (DIRECT) {
                    t.rid = p_start + li;
                    t.key = (KeyT)(a.src_p.ids ? __ldg(a.src_p.col + ld_stream_u32(a.src_p.ids + t.rid))
                                               : ld_stream_u64(a.src_p.col + t.rid));
                    if (a.pred_p.npred) {
                        if (preds_hold_row(a.pred_p, t.rid)) ++my_valid;
                        else t.rid = 0xFFFFFFFFu;   // filtered out (never a row id: at most 2^32 - 1 rows)
                    }
                } else if constexpr (sizeof(KeyT) == 8) {
                    const ulonglong2 v = ld_stream_u64x2(reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                    t.key = v.x;
                    t.rid = (uint32_t)v.y;
                } else {
                    const uint64_t v = ld_stream_u64(reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                    t.key = (uint32_t)v;
                    t.rid = (uint32_t)(v >> 32);
                }
            } else {
                t.key = 0;
                t.rid = 0;
            }
        };
        // first probe batch in flight while the table is built
        ProbeTup<KeyT> cur[U], nxt[U];
#pragma unroll
        for (int j = 0; j < U; ++j) load_probe((uint32_t)(j * NT) + tid, cur[j]);

        // ---- build (K6) ---------------------------------------------------
        tab.clear(tid, NT);
        __syncthreads();
        for (uint32_t i = tid; i < b_count; i += NT) {
            KeyT key;
            if constexpr (DIRECT) {
                const uint32_t rid = b_start + i;
                key = (KeyT)(a.src_b.ids ? a.src_b.col[a.src_b.ids[rid]] : a.src_b.col[rid]);
            } else {
                key = tup_b[b_start + i].key;
            }
            tab.insert(i, key, table_hash<KeyT>(key, a.radix_bits, hash_log2));
        }
        __syncthreads();

        // ---- probe (K7) ---------------------------------------------------
        unsigned long long item_matches = 0;
        uint32_t           queued       = 0;   // warp-uniform mirror of wq_cnt[wid]
        for (uint32_t off = 0; off < p_count; off += NT * U) {
#pragma unroll
            for (int j = 0; j < U; ++j) load_probe(off + (uint32_t)(NT * U + j * NT) + tid, nxt[j]);
            const bool full_round = off + NT * U <= p_count;
#pragma unroll
            for (int j = 0; j < U; ++j) {
                uint32_t nh = 0;
                if ((full_round || off + (uint32_t)(j * NT) + tid < p_count) && (!DIRECT || cur[j].rid != 0xFFFFFFFFu))
                    nh = probe_one(cur[j].key, cur[j].rid);
                if constexpr (MODE == MODE_COUNT) {
                    item_matches += nh;
                } else {
                    my_matches += nh;
                    queued += __reduce_add_sync(kFullMask, nh);
                    if (queued >= 32u) {
                        __syncwarp();
                        uint32_t have = min(queued, (uint32_t)kWarpQueue);
                        while (have >= 32u) {
                            drain(have, 32u);
                            have -= 32u;
                        }
                        __syncwarp();
                        if (lane == 0) wq_cnt[wid] = have;
                        __syncwarp();
                        queued = have;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) cur[j] = nxt[j];
        }
        if constexpr (MODE != MODE_COUNT) {
            // leftovers of this item (queue entries are relative to this item's build chunk)
            __syncwarp();
            if (queued) drain(queued, queued);
            __syncwarp();
            if (lane == 0) wq_cnt[wid] = 0;
        }
        if constexpr (MODE == MODE_COUNT) {
            const unsigned long long ws = warp_sum_u64(item_matches);
            if (lane == 0 && ws) atomicAdd(&s_cnt, ws);
            __syncthreads();
            if (tid == 0) {
                a.item_count[item_w] = s_cnt;
                if (s_cnt) atomicAdd(a.total, s_cnt);
            }
        }
        __syncthreads();   // table and s_item are reused by the next item
    }

    if constexpr (DIRECT) {
        if (a.pred_p.npred && a.valid_p) {   // (every mode loads each probe row exactly once per build chunk)
            const unsigned long long wv = warp_sum_u64(my_valid);
            if (lane == 0 && wv) atomicAdd(a.valid_p, wv);
        }
    }
    if constexpr (MODE == MODE_SUM) {
        // K9: warp-shuffle reduction, one atomicAdd(u64) per warp and projection
        const unsigned long long wm = warp_sum_u64(my_matches);
        if (lane == 0 && wm) atomicAdd(a.total, wm);
#pragma unroll
        for (int k = 0; k < NPA; ++k) {
            if (k < a.nproj) {
                const unsigned long long ws = warp_sum_u64(my_sum[k] + pend[k]);
                if (lane == 0 && wm) atomicAdd(a.sums + k, ws);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K6 + K7 (+K9) for 32-bit keys: tag table, one shared-memory access per probe.
//
// hash_join_kernel above walks data-dependent chains with full keys in shared
// memory; ncu showed the probe bound by the SM's L1/shared-memory pipe (random
// LDS cost ~3.5 wavefronts per warp whatever their width, and every chain step
// is another one) and by instruction issue (~100 warp instructions per probe).
// This kernel makes a probe ONE 32-bit LDS:
//   * y = mix(key >> radix_bits) is a bijection on the 32 - radix_bits bits
//     that differ inside a partition; slot = low L bits of y, tag = the rest
//     (<= 15 bits).  Slot and tag together identify the key exactly, so a tag
//     match IS a key match: no key array in shared memory, no verification;
//   * a slot word is [tag | has_next | position of the build tuple in its
//     chunk]; build tuples that collide in a slot are chained through a second
//     word array of the same format (tagnext[pos] = word of the next element);
//   * the probe never branches per lane: lanes whose slot word has a matching
//     tag, or a chain behind it, push an 8-byte entry into the warp's queue at
//     a position computed from ballots; the queue is drained 32 entries at a
//     time with every lane active (chain walks re-push the matches they find;
//     payload gathers of a drain are all in flight together and are consumed
//     by the next drain).
// Used for partitioned joins of 32-bit keys with radix_bits + slots_log2 >= 17; a table holds up
// to 32766 build tuples (2 CTAs/SM of 512 threads for small tables, 1 CTA of 1024 for large ones).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

// flag written by a peer GPU (copy engine or remote store) into this GPU's memory
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= epoch; bounded (about 2 s) so that a lost peer cannot hang the GPU: returns false on time-out
__device__ __forceinline__ bool spin_until_epoch(const uint32_t *flag, uint32_t epoch) {
    if (ld_acquire_sys_u32(flag) >= epoch) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys_u32(flag) < epoch) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000ll) return false;
    }
    return true;
}

// work item of the join kernels: (build chunk, probe slice)
struct JoinItem {
    uint32_t valid, b_start, b_count, p_start, p_count, w, part;
};

template <bool DIRECT>
__device__ __forceinline__ JoinItem fetch_join_item(const JoinArgs &a, int lane) {
    JoinItem it{0, 0, 0, 0, 0, 0, 0};
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(a.work_counter, 1u);
    w    = __shfl_sync(kFullMask, w, 0);
    it.w = w;
    if constexpr (DIRECT) {
        if (w < a.n_items_direct) {
            it.valid           = 1;
            const uint32_t rch = w / a.sc_direct;
            const uint32_t ssl = w % a.sc_direct;
            it.b_start         = rch * a.cap;
            it.b_count         = min(a.cap, a.src_b.n - it.b_start);
            it.p_start         = ssl * a.slice;
            it.p_count         = min(a.slice, a.src_p.n - it.p_start);
        }
    } else {
        const uint32_t n_items = a.item_start[a.nparts];
        if (w < n_items) {
            it.valid    = 1;
            uint32_t lo = 0, hi = a.nparts;   // 32-ary search: item_start[p] <= w < item_start[p+1]
            while (hi - lo > 1) {
                const uint32_t step = (hi - lo + 31) / 32;
                const uint32_t idx  = lo + (uint32_t)lane * step;
                const bool     le   = idx < hi && a.item_start[idx] <= w;
                const uint32_t c    = __popc(__ballot_sync(kFullMask, le));   // >= 1
                const uint32_t nlo  = lo + (c - 1) * step;
                hi                  = min(hi, nlo + step);
                lo                  = nlo;
            }
            const uint32_t p   = lo;
            it.part            = p;
            const uint32_t k   = w - a.item_start[p];
            const uint32_t b0  = a.off_b[p], b1 = a.off_b[p + 1];
            const uint32_t p0  = a.off_p[p], p1 = p0 + a.cnt_p[p];
            const uint32_t sc  = (p1 - p0 + a.slice - 1) / a.slice;
            const uint32_t rch = k / sc, ssl = k % sc;
            it.b_start         = b0 + rch * a.cap;
            it.b_count         = min(a.cap, b1 - it.b_start);
            it.p_start         = p0 + ssl * a.slice;
            it.p_count         = min(a.slice, p1 - it.p_start);
        }
    }
    return it;
}

constexpr uint32_t kTagShift  = 16;            // word = [tag 16 | has_next 1 | position 15]
constexpr uint32_t kNextBit   = 0x8000u;
constexpr uint32_t kIdxMask   = 0x7FFFu;
constexpr uint32_t kEmptyWord = 0xFFFF7FFFu;   // tag all ones (never a real tag: tags have <= 15 bits), no chain
constexpr int      kTagQueue  = 64;            // entries per warp (<= 31 left + 32 pushed per probe)
// queue entry word: [31] match at position | [30:16] tag of the probe key | [15] walk the chain behind position | [14:0] position
constexpr uint32_t kQMatch    = 0x80000000u;

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// x = key >> radix_bits.  tag = x >> L, slot = (low L bits of x) ^ hash(tag): given slot and tag the low bits of
// x are slot ^ hash(tag), so (slot, tag) <-> x is a bijection and equal (slot, tag) means equal keys.
__device__ __forceinline__ void slot_and_tag(uint32_t x, uint32_t L, uint32_t smask, uint32_t &slot, uint32_t &tag) {
    tag  = x >> L;
    slot = (x ^ (tag * 0x9E3779B1u)) & smask;
}
__device__ __forceinline__ uint64_t ld_gather_u64(const uint64_t *p) {
    uint64_t v;   // random 8-byte gather: fetch 64 B from DRAM, not the default 128 B
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// K64: genuinely 64-bit keys (16-byte tuples).  The table is the same one-word-per-slot tag table, but slot and tag
// come from a hash of the key bits above the radix, so a tag match is only a candidate: the queue entry carries
// the probe key and the drain compares it with the build tuple's key — which it reads anyway, in the same 16-byte
// load as the row id / carried value.  The probe stays one shared-memory load; rhjoin.c:154-216 compares full keys
// on every chain step.
template <int NT, int MINB, int G, int MODE, int NP, bool SEG = false, bool K64 = false>
__global__ void __launch_bounds__(NT, MINB)
tag_join_kernel(const JoinArgs a) {
    using KeyT = typename std::conditional<K64, uint64_t, uint32_t>::type;
    using TupT = typename TupOf<KeyT>::type;
    static_assert(!(K64 && SEG), "the segmented (multi-GPU) build side carries 32-bit keys");
    constexpr int      NW  = NT / 32;
    constexpr int      NPA = NP > 0 ? NP : 1;
    constexpr int      QN  = kTagQueue;
    constexpr uint32_t QE  = K64 ? 16u : 8u;   // bytes per queue entry: {word, probe row id[, probe key]}
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: slots [4 << L bytes] | tagnext [4 * cap bytes] | queues [NW * QN * QE bytes]
    const uint32_t L       = a.slots_log2;
    const uint32_t smask   = (1u << L) - 1u;
    const uint32_t s_slots = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t s_next  = s_slots + (4u << L);
    const uint32_t s_queue = s_next + 4u * a.cap;
    uint32_t      *slots   = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t      *tagnext = slots + (1u << L);

    __shared__ uint32_t           s_item[6];
    __shared__ uint32_t           s_cursor;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_cnt;
    __shared__ uint32_t           s_seg_vend[kMaxPeers], s_seg_delta[kMaxPeers];   // SEG: see bphys()

    const uint32_t tid  = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t wid  = tid >> 5;
    const uint32_t lt   = (1u << lane) - 1u;
    const uint32_t my_q = s_queue + wid * (uint32_t)QN * QE;
    // virtual build position (off_b space) -> physical index in tup_b / part_vals
    auto bphys = [&](uint32_t v) -> uint32_t {
        if constexpr (SEG) {
            int r = 0;
            while (r + 1 < a.nseg && v >= s_seg_vend[r]) ++r;
            return v + s_seg_delta[r];
        } else {
            return v;
        }
    };
    const TupT *tup_b = static_cast<const TupT *>(a.tup_b);
    const TupT *tup_p = static_cast<const TupT *>(a.tup_p);
    // slot and tag of a key: exact (a bijection on the bits above the radix) for 32-bit keys, a hash for 64-bit keys
    auto slot_tag = [&](KeyT key, uint32_t &slot, uint32_t &tag) {
        if constexpr (K64) {
            const uint64_t x = (uint64_t)key >> a.radix_bits;
            const uint32_t h = ((uint32_t)x ^ (uint32_t)(x >> 32)) * 0x9E3779B1u;
            slot             = (h ^ (h >> 15)) & smask;
            tag              = h >> 17;   // 15 bits: never the all-ones tag of an empty slot
        } else {
            slot_and_tag((uint32_t)key >> a.radix_bits, L, smask, slot, tag);
        }
    };
    // build tuple at physical position p: row id (or carried value) and, for 64-bit keys, the key to verify against
    auto build_rid = [&](uint32_t p, KeyT &bkey) -> uint32_t {
        if constexpr (K64) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(tup_b + p);
            bkey               = v.x;
            return (uint32_t)v.y;
        } else {
            bkey = 0;
            return SEG ? __ldcg(&tup_b[p].rid) : tup_b[p].rid;
        }
    };

    unsigned long long my_matches = 0;
    unsigned long long my_sum[NPA], pend[NPA];
#pragma unroll
    for (int k = 0; k < NPA; ++k) my_sum[k] = pend[k] = 0;
    uint32_t b_start = 0;
    [[maybe_unused]] uint32_t b_cnt = 0;   // build tuples of the current item (checked build)
    uint32_t queued  = 0;   // warp-uniform number of entries in this warp's queue

    // complete up to 32 match entries, one per lane (is_match: this lane holds one): verification of a 64-bit
    // candidate, then the SUM gathers / the pair / the count.  Payload loads are left in flight (pend[]) and consumed
    // by the next call, so a run of calls keeps a warp's worth of gathers outstanding.
    auto complete = [&](bool is_match, uint32_t ex, uint32_t ey, [[maybe_unused]] KeyT ekey) {
        B200_DCHECK(!is_match || (ex & kIdxMask) < b_cnt);   // a position inside the build chunk
        [[maybe_unused]] uint32_t brid64 = 0;
        if constexpr (K64) {
            if (is_match) {   // verify the candidate: one 16-byte load gives the key and the row id / carried value
                KeyT bkey;
                brid64   = build_rid(bphys(b_start + (ex & kIdxMask)), bkey);
                is_match = bkey == ekey;
            }
        }
        if constexpr (MODE == MODE_SUM) {
            const uint32_t bpos = bphys(b_start + (ex & kIdxMask));
            uint32_t       brid = 0;
            if constexpr (K64) brid = brid64;
            else if (is_match && a.need_brid) brid = SEG ? __ldcg(&tup_b[bpos].rid) : tup_b[bpos].rid;
#pragma unroll
            for (int k = 0; k < NPA; ++k) {
                my_sum[k] += pend[k];
                pend[k] = 0;
                if (k < a.nproj && is_match) {
                    if (a.proj[k].part_vals == B200_PROJ_IN_RID) {
                        pend[k] = a.proj[k].side == 0 ? brid : ey;   // the row-id slot carries the value
                    } else if (a.proj[k].part_vals) {
                        pend[k] = a.proj[k].part_vals[bpos];   // dense window of this partition, L2-resident
                    } else {
                        const uint32_t r  = a.proj[k].side == 0 ? brid : ey;
                        const uint32_t rr = a.proj[k].ids ? __ldg(a.proj[k].ids + r) : r;
                        pend[k]           = ld_gather_u64(a.proj[k].col + rr);
                    }
                }
            }
            my_matches += is_match ? 1u : 0u;
        } else if constexpr (MODE == MODE_WRITE) {
            const uint32_t bal = __ballot_sync(kFullMask, is_match);
            uint32_t       pos = 0;
            if (lane == 0 && bal) pos = atomicAdd(&s_cursor, (uint32_t)__popc(bal));
            pos = __shfl_sync(kFullMask, pos, 0);
            if (is_match) {
                const uint32_t o  = pos + __popc(bal & lt);
                a.out_b[s_base + o] = K64 ? brid64 : tup_b[bphys(b_start + (ex & kIdxMask))].rid;
                a.out_p[s_base + o] = ey;
            }
        } else {
            my_matches += is_match ? 1u : 0u;
        }
    };
    // entry `slot` of this warp's queue
    auto q_load = [&](uint32_t slot, uint2 &e, KeyT &ekey) {
        const uint32_t qa = my_q + slot * QE;
        e                 = lds_v2(qa);
        ekey              = 0;
        if constexpr (K64) {
            const uint2 kk = lds_v2(qa + 8u);
            ekey           = (uint64_t)kk.x | ((uint64_t)kk.y << 32);
        }
    };

    // pop `take` (<= 32) entries, one per lane: a match entry is completed, a chain entry is walked and the matches
    // it finds are pushed back as match entries.  When the queue cannot take another step's worth of them, the walk
    // pauses and the newest entries — all of them matches this walk pushed — are completed, 32 at a time with every
    // lane's loads in flight together.  (Completing a match right where the walk found it made every chain step wait
    // for a dependent global load: config 5 x100 has keys with thousands of duplicates, and one of its queries took
    // 0.97 s in this kernel instead of 0.03 s.)
    auto drain = [&](uint32_t take) {
        B200_DCHECK(take <= 32u && take <= queued && queued <= (uint32_t)QN);
        queued -= take;
        const bool mine = lane < take;
        uint2      e;
        KeyT       ekey;
        q_load(queued + (mine ? lane : 0u), e, ekey);
        const uint32_t t = (e.x >> 16) & 0x7FFFu;
        complete(mine && (e.x & kQMatch) != 0u, e.x, e.y, ekey);
        // ---- chain entries ----
        const uint32_t base    = queued;   // entries below were queued by the probe loop and may be chain entries
        bool           walking = mine && (e.x & kNextBit) != 0u;
        uint32_t       pos     = e.x & kIdxMask;
        while (__any_sync(kFullMask, walking)) {
            if (queued + 32u > (uint32_t)QN) {   // no room for a step's worth of hits: complete the newest matches
                __syncwarp();
                const uint32_t n = min(32u, queued - base);
                B200_DCHECK(n >= 1u);
                queued -= n;
                uint2 e2;
                KeyT  k2;
                q_load(queued + (lane < n ? lane : 0u), e2, k2);
                complete(lane < n, e2.x, e2.y, k2);
                __syncwarp();
                continue;
            }
            B200_DCHECK(!walking || pos < a.cap);
            uint32_t w = kEmptyWord;
            if (walking) w = lds_u32(s_next + pos * 4u);
            const bool     hit = walking && (w >> kTagShift) == t;
            const uint32_t bal = __ballot_sync(kFullMask, hit);
            if (hit) {
                const uint32_t slot = queued + __popc(bal & lt);
                B200_DCHECK(slot < (uint32_t)QN);
                sts_v2(my_q + slot * QE, kQMatch | (t << 16) | (w & kIdxMask), e.y);
                if constexpr (K64) sts_v2(my_q + slot * QE + 8u, (uint32_t)ekey, (uint32_t)((uint64_t)ekey >> 32));
            }
            queued += (uint32_t)__popc(bal);
            walking = walking && (w & kNextBit) != 0u;
            pos     = w & kIdxMask;
        }
        __syncwarp();
    };

    for (;;) {
        if (tid < 32) {
            const JoinItem it = fetch_join_item<false>(a, (int)lane);
            if constexpr (SEG) {
                if (it.valid) {
                    // lane r owns run r of this partition: inclusive scan of the run lengths over the lanes
                    const uint32_t cnt = (int)lane < a.nseg ? a.seg_cnt[lane * a.nparts + it.part] : 0u;
                    uint32_t       inc = cnt;
#pragma unroll
                    for (int d = 1; d < kMaxPeers; d <<= 1) {
                        const uint32_t up = __shfl_up_sync(kFullMask, inc, d);
                        if ((int)lane >= d) inc += up;
                    }
                    if ((int)lane < a.nseg) {
                        const uint32_t vstart = a.off_b[it.part] + inc - cnt;
                        const uint32_t phys   = a.seg_off[lane * a.nparts + it.part];
                        s_seg_vend[lane]      = a.off_b[it.part] + inc;
                        s_seg_delta[lane]     = phys - vstart;
                        if (a.wait_flags != nullptr && cnt != 0u) {
                            const uint32_t chunk = (phys + cnt - 1u - lane * a.seg_rows) / a.chunk_rows;
                            if (!spin_until_epoch(a.wait_flags + chunk * kMaxPeers + lane, *a.wait_epoch))
                                *a.wait_error = 1u;
                        }
                    }
                }
            }
            if (lane == 0) {
                s_item[0] = it.valid;
                s_item[1] = it.b_start;
                s_item[2] = it.b_count;
                s_item[3] = it.p_start;
                s_item[4] = it.p_count;
                s_item[5] = it.w;
                s_cnt     = 0ull;
                s_cursor  = 0u;
                if constexpr (MODE == MODE_WRITE) {
                    if (it.valid) s_base = atomicAdd(a.out_cursor, a.item_count[it.w]);
                }
            }
        }
        __syncthreads();
        if (s_item[0] == 0u) break;
        b_start                = s_item[1];
        const uint32_t b_count = s_item[2];
        const uint32_t p_start = s_item[3], p_count = s_item[4];
        const uint32_t item_w  = s_item[5];
        b_cnt                  = b_count;

        // (a padding lane of the last round is recognised by its index, never by a row-id sentinel: the row-id slot
        // may carry an arbitrary 32-bit SUM value)
        auto load_probe = [&](uint32_t li, KeyT &key, uint32_t &rid) {
            key = 0;
            rid = 0;
            if (li < p_count) {
                if constexpr (K64) {
                    const ulonglong2 v = ld_stream_u64x2(reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                    key                = v.x;
                    rid                = (uint32_t)v.y;
                } else {
                    const uint64_t v = ld_stream_u64(reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                    key              = (uint32_t)v;
                    rid              = (uint32_t)(v >> 32);
                }
            }
        };
        // first group in flight while the table is built
        KeyT     ckey[G], nkey[G];
        uint32_t crid[G], nrid[G];
#pragma unroll
        for (int j = 0; j < G; ++j) load_probe((uint32_t)(j * NT) + tid, ckey[j], crid[j]);

        // ---- build (K6) ---------------------------------------------------
        for (uint32_t s = tid; s <= smask; s += NT) slots[s] = kEmptyWord;
        __syncthreads();
        // eight build keys per thread are loaded together before their (atomic, hence ordered) inserts:
        // one exposed global-memory latency per eight inserts instead of one per insert
        constexpr int KB = 8;
        for (uint32_t i0 = tid; i0 < b_count; i0 += NT * KB) {
            KeyT bk[KB];
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const uint32_t i = i0 + (uint32_t)u * NT;
                // (SEG: the tuples were written by peer GPUs while this kernel may already be running — read them
                // at L2, the point of coherence, not through the non-coherent path)
                if constexpr (K64) {
                    bk[u] = i < b_count ? ld_stream_u64(&tup_b[b_start + i].key) : 0ull;
                } else {
                    bk[u] = i < b_count ? (SEG ? __ldcg(&tup_b[bphys(b_start + i)].key)
                                               : ld_stream_u32(&tup_b[bphys(b_start + i)].key))
                                        : 0u;
                }
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const uint32_t i = i0 + (uint32_t)u * NT;
                if (i < b_count) {
                    uint32_t h, t;
                    slot_tag(bk[u], h, t);
                    B200_DCHECK(h <= smask && i < a.cap && i <= kIdxMask && t <= 0x7FFFu);
                    uint32_t old = slots[h], assumed;
                    do {
                        assumed           = old;
                        const uint32_t nw = (t << kTagShift) | (assumed != kEmptyWord ? kNextBit : 0u) | i;
                        old               = atomicCAS(&slots[h], assumed, nw);
                    } while (old != assumed);
                    tagnext[i] = old;
                }
            }
        }
        __syncthreads();

        // ---- probe (K7) ---------------------------------------------------
        // groups of G probes, ping-pong between two register sets: while one group is probed the next one's
        // tuples are in flight
        auto probe_group = [&](const KeyT (&key)[G], const uint32_t (&rid)[G], uint32_t first) {
            uint32_t w[G], t[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                uint32_t h;
                slot_tag(key[j], h, t[j]);
                w[j] = lds_u32(s_slots + h * 4u);
            }
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const bool     ok   = first + (uint32_t)(j * NT) + tid < p_count;
                const bool     is_m = (w[j] >> kTagShift) == t[j];
                const bool     need = ok && (is_m || (w[j] & kNextBit) != 0u);
                const uint32_t act  = __ballot_sync(kFullMask, need);
                if (act) {
                    // queued <= 31 here and a probe adds <= 32 entries: no overflow (QN = 64)
                    if (need) {
                        B200_DCHECK(queued + (uint32_t)__popc(act & lt) < (uint32_t)QN);
                        const uint32_t qs = my_q + (queued + __popc(act & lt)) * QE;
                        sts_v2(qs, (is_m ? kQMatch : 0u) | (t[j] << 16) | (w[j] & 0xFFFFu), rid[j]);
                        if constexpr (K64) sts_v2(qs + 8u, (uint32_t)key[j], (uint32_t)((uint64_t)key[j] >> 32));
                    }
                    queued += __popc(act);
                    __syncwarp();
                    while (queued >= 32u) drain(32u);
                }
            }
        };
        for (uint32_t off = 0; off < p_count; off += 2 * NT * G) {
#pragma unroll
            for (int j = 0; j < G; ++j) load_probe(off + (uint32_t)(NT * G + j * NT) + tid, nkey[j], nrid[j]);
            probe_group(ckey, crid, off);
#pragma unroll
            for (int j = 0; j < G; ++j) load_probe(off + (uint32_t)(2 * NT * G + j * NT) + tid, ckey[j], crid[j]);
            if (off + NT * G < p_count) probe_group(nkey, nrid, off + (uint32_t)(NT * G));
        }
        // leftovers of this item (queue entries are relative to this item's build chunk)
        __syncwarp();
        while (queued) drain(min(queued, 32u));
        if constexpr (MODE == MODE_COUNT) {
            const unsigned long long ws = warp_sum_u64(my_matches);
            my_matches                  = 0;
            if (lane == 0 && ws) atomicAdd(&s_cnt, ws);
            __syncthreads();
            if (tid == 0) {
                a.item_count[item_w] = s_cnt;
                if (s_cnt) atomicAdd(a.total, s_cnt);
            }
        }
        __syncthreads();   // table and s_item are reused by the next item
    }

    if constexpr (MODE == MODE_SUM) {
        const unsigned long long wm = warp_sum_u64(my_matches);
        if (lane == 0 && wm) atomicAdd(a.total, wm);
#pragma unroll
        for (int k = 0; k < NPA; ++k) {
            if (k < a.nproj) {
                const unsigned long long ws = warp_sum_u64(my_sum[k] + pend[k]);
                if (lane == 0 && wm) atomicAdd(a.sums + k, ws);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Ballot compaction shared by K1 and K11.  Element (j, lane) of warp `wid` in a
// tile sits at tile_first + wid*32*U + j*32 + lane, so loads are coalesced and
// the emitted indices stay ascending inside a tile.  Each tile reserves its
// output run with ONE global atomicAdd.
// ---------------------------------------------------------------------------
template <int NT, int U>
__device__ __forceinline__ void emit_compacted(const bool (&keep)[U], uint64_t warp_first,
                                               uint32_t *__restrict__ out,
                                               unsigned long long *__restrict__ out_count,
                                               uint32_t *warp_sums, uint32_t *s_base) {
    const int lane = threadIdx.x & 31;
    unsigned  bal[U];
    uint32_t  wc = 0;
#pragma unroll
    for (int j = 0; j < U; ++j) {
        bal[j] = __ballot_sync(kFullMask, keep[j]);
        wc += __popc(bal[j]);
    }
    const uint32_t excl  = block_exclusive_scan<NT>(lane == 0 ? wc : 0u, warp_sums);
    const uint32_t total = warp_sums[NT / 32];
    if (threadIdx.x == 0) *s_base = total ? (uint32_t)atomicAdd(out_count, (unsigned long long)total) : 0u;
    __syncthreads();
    uint32_t       wbase = *s_base + __shfl_sync(kFullMask, excl, 0);
    const unsigned lt    = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < U; ++j) {
        if (keep[j]) out[wbase + __popc(bal[j] & lt)] = (uint32_t)(warp_first + (uint32_t)(j * 32 + lane));
        wbase += __popc(bal[j]);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// K1: column scan with one predicate (filter.c:92-190).  The constant is the
// reference's 32-bit int converted to uint64 by the usual C rules
// (filter.c:118 compares `uint64_t > int`).  Emits row ids (ids == nullptr)
// or positions in the intermediate (ids != nullptr).
// cmp: 0 '<', 1 '>', 2 '='.
// ---------------------------------------------------------------------------
template <int NT, int U>
__global__ void __launch_bounds__(NT)
scan_filter_kernel(KeySrc src, int cmp, uint64_t constant, uint32_t *__restrict__ out,
                   unsigned long long *__restrict__ out_count) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    __shared__ uint32_t s_base;
    constexpr uint32_t  TILE   = NT * U;
    const uint64_t      n      = src.n;
    const uint64_t      ntiles = (n + TILE - 1) / TILE;
    const int           lane   = threadIdx.x & 31;
    const int           wid    = threadIdx.x >> 5;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t warp_first = tile * TILE + (uint64_t)wid * (32 * U);
        uint64_t       v[U];
        bool           keep[U];
        if (src.ids) {
            uint32_t id[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint64_t i = warp_first + (uint32_t)(j * 32 + lane);
                id[j]            = i < n ? ld_stream_u32(src.ids + i) : 0u;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint64_t i = warp_first + (uint32_t)(j * 32 + lane);
                v[j]             = i < n ? __ldg(src.col + id[j]) : 0ull;
            }
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const uint64_t i = warp_first + (uint32_t)(j * 32 + lane);
                v[j]             = i < n ? ld_stream_u64(src.col + i) : 0ull;
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t i = warp_first + (uint32_t)(j * 32 + lane);
            const bool     k = cmp == 0 ? (v[j] < constant) : cmp == 1 ? (v[j] > constant) : (v[j] == constant);
            keep[j]          = k && i < n;
        }
        emit_compacted<NT, U>(keep, warp_first, out, out_count, warp_sums, &s_base);
    }
}

// K11 (inter_res.c:376-385) and SelfJoin (inter_res.c:234-263): positions p
// with colA[ta ? ta[p] : p] == colB[tb ? tb[p] : p].
template <int NT, int U>
__global__ void __launch_bounds__(NT)
inter_equal_kernel(const uint64_t *__restrict__ col_a, const uint32_t *__restrict__ ta,
                   const uint64_t *__restrict__ col_b, const uint32_t *__restrict__ tb, uint32_t n,
                   uint32_t *__restrict__ out, unsigned long long *__restrict__ out_count) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    __shared__ uint32_t s_base;
    constexpr uint32_t  TILE   = NT * U;
    const uint64_t      ntiles = ((uint64_t)n + TILE - 1) / TILE;
    const int           lane   = threadIdx.x & 31;
    const int           wid    = threadIdx.x >> 5;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t warp_first = tile * TILE + (uint64_t)wid * (32 * U);
        bool           keep[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t i = warp_first + (uint32_t)(j * 32 + lane);
            bool           k = false;
            if (i < n) {
                const uint32_t ra = ta ? ld_stream_u32(ta + i) : (uint32_t)i;
                const uint32_t rb = tb ? ld_stream_u32(tb + i) : (uint32_t)i;
                k                 = __ldg(col_a + ra) == __ldg(col_b + rb);
            }
            keep[j] = k;
        }
        emit_compacted<NT, U>(keep, warp_first, out, out_count, warp_sums, &s_base);
    }
}

// ---------------------------------------------------------------------------
// K8: gather every active row-id column of an intermediate through one list
// of positions (inter_res.c:79-98, 119-137; filter.c:60-76; inter_res.c:304-313).
// One pass over the positions serves up to kMaxGather columns.
// ---------------------------------------------------------------------------
struct GatherArgs {
    const uint32_t *pos;
    uint32_t        m;
    int             ncols;
    const uint32_t *in[kMaxGather];
    uint32_t       *out[kMaxGather];
};

static __global__ void __launch_bounds__(256) gather_columns_kernel(const GatherArgs g) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.m; i += stride) {
        const uint32_t p = ld_stream_u32(g.pos + i);
#pragma unroll
        for (int c = 0; c < kMaxGather; ++c)
            if (c < g.ncols) g.out[c][i] = __ldg(g.in[c] + p);
    }
}

// CartesianInterResults (inter_res.c:405-418): row index = i * n2 + j.
static __global__ void __launch_bounds__(256)
cartesian_kernel(const uint32_t *__restrict__ in, uint32_t n1, uint32_t n2, int from_first,
                 uint32_t *__restrict__ out) {
    const uint64_t total  = (uint64_t)n1 * n2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride)
        out[k] = from_first ? in[k / n2] : in[k % n2];
}

// ---------------------------------------------------------------------------
// K9: SUM checksum (inter_res.c:332-333) for up to kMaxProj projections in one
// pass: sums[k] += col_k[ids_k ? ids_k[j] : j]; warp-shuffle reduction and one
// atomicAdd(u64) per warp.
// ---------------------------------------------------------------------------
struct ChecksumArgs {
    uint32_t            m;
    int                 nproj;
    const uint64_t     *col[kMaxProj];
    const uint32_t     *ids[kMaxProj];
    unsigned long long *sums;
};

static __global__ void __launch_bounds__(256) checksum_kernel(const ChecksumArgs c) {
    unsigned long long acc[kMaxProj];
#pragma unroll
    for (int k = 0; k < kMaxProj; ++k) acc[k] = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < c.m; j += stride) {
#pragma unroll
        for (int k = 0; k < kMaxProj; ++k) {
            if (k < c.nproj) {
                const uint32_t r = c.ids[k] ? ld_stream_u32(c.ids[k] + j) : (uint32_t)j;
                acc[k] += __ldg(c.col[k] + r);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxProj; ++k) {
        if (k < c.nproj) {
            const unsigned long long ws = warp_sum_u64(acc[k]);
            if ((threadIdx.x & 31) == 0 && ws) atomicAdd(c.sums + k, ws);
        }
    }
}

// Column statistics at registration (relation_map.c:53-83: min, max and the number of distinct values, the inputs of
// stats.c's selectivity formulas).  min / max: one pass.  distinct: the reference marks a `unsigned short` array of
// min(u - l + 1, 50 000 000) entries — entry v - l when the range is below 50 000 000, else entry (v - l) % 5 000 000
// (sic: five million, relation_map.c:71) — and counts the marked entries; here the array is a bitmap in L2
// (at most 6.25 MB), marked with atomicOr and counted with popc.
static __global__ void __launch_bounds__(256)
column_minmax_kernel(const uint64_t *__restrict__ col, uint64_t n, unsigned long long *__restrict__ out_min,
                     unsigned long long *__restrict__ out_max) {
    unsigned long long lo = ~0ull, hi = 0;
    const uint64_t     stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long v = ld_stream_u64(col + i);
        lo                         = v < lo ? v : lo;
        hi                         = v > hi ? v : hi;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(kFullMask, lo, d), b = __shfl_xor_sync(kFullMask, hi, d);
        lo                         = a < lo ? a : lo;
        hi                         = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out_min, lo);
        atomicMax(out_max, hi);
    }
}
static __global__ void __launch_bounds__(256)
column_mark_kernel(const uint64_t *__restrict__ col, uint64_t n, uint64_t lo, uint64_t modulus, uint32_t *__restrict__ bits) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t e = ld_stream_u64(col + i) - lo;
        if (modulus) e %= modulus;
        atomicOr(&bits[e >> 5], 1u << (e & 31));
    }
}
static __global__ void __launch_bounds__(256)
bitmap_count_kernel(const uint32_t *__restrict__ bits, uint64_t nwords, unsigned long long *__restrict__ out) {
    unsigned long long c      = 0;
    const uint64_t     stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += stride) c += __popc(bits[i]);
    c = warp_sum_u64(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// Column maximum at registration (selects the 32-bit-key kernels).
static __global__ void __launch_bounds__(256)
column_max_kernel(const uint64_t *__restrict__ col, uint64_t n, unsigned long long *__restrict__ out) {
    unsigned long long m      = 0;
    const uint64_t     stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long v = ld_stream_u64(col + i);
        m                          = v > m ? v : m;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFullMask, m, d);
        m                          = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// multi-GPU build side: from the all-gathered per-rank histograms hist_all[world][nparts] compute the
// global histogram and this rank's scatter start inside the global partition layout (one CTA)
template <int NT>
__global__ void __launch_bounds__(NT)
build_cursors_kernel(const uint32_t *__restrict__ hist_all, uint32_t world, uint32_t rank, uint32_t nparts,
                     uint32_t *__restrict__ total, uint32_t *__restrict__ my_start) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    const uint32_t per   = (nparts + NT - 1) / NT;
    const uint32_t first = threadIdx.x * per;
    uint32_t       s     = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts)
            for (uint32_t r = 0; r < world; ++r) s += hist_all[r * nparts + b];
    }
    uint32_t run = block_exclusive_scan<NT>(s, warp_sums);
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            uint32_t tot = 0, before = 0;
            for (uint32_t r = 0; r < world; ++r) {
                const uint32_t c = hist_all[r * nparts + b];
                if (r < rank) before += c;
                tot += c;
            }
            total[b]    = tot;
            my_start[b] = run + before;
            run += tot;
        }
    }
}

// rank-major build layout: region r of the build buffer holds rank r's shard in partition order.  From the
// all-gathered histograms: seg_off[r][p] = r * seg_rows + (exclusive scan of hist_all[r] over p) and the global
// histogram total[p].  One CTA; rank r is scanned by warp... simply one rank after the other.
template <int NT>
__global__ void __launch_bounds__(NT)
segment_offsets_kernel(const uint32_t *__restrict__ hist_all, uint32_t world, uint32_t nparts, uint32_t seg_rows,
                       uint32_t *__restrict__ seg_off, uint32_t *__restrict__ total, uint32_t seg_head) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    const uint32_t per   = (nparts + NT - 1) / NT;
    const uint32_t first = threadIdx.x * per;
    for (uint32_t r = 0; r < world; ++r) {
        uint32_t s = 0;
        for (uint32_t k = 0; k < per; ++k)
            if (first + k < nparts) s += hist_all[r * nparts + first + k];
        uint32_t run = block_exclusive_scan<NT>(s, warp_sums);
        for (uint32_t k = 0; k < per; ++k) {
            const uint32_t b = first + k;
            if (b < nparts) {
                seg_off[r * nparts + b] = r * seg_rows + seg_head + run;   // (seg_head: tuple slots a region keeps for its header)
                run += hist_all[r * nparts + b];
            }
        }
        __syncthreads();
    }
    for (uint32_t b = threadIdx.x; b < nparts; b += NT) {
        uint32_t t = 0;
        for (uint32_t r = 0; r < world; ++r) t += hist_all[r * nparts + b];
        total[b] = t;
    }
}

// ---------------------------------------------------------------------------
// Radix-sharded exchange ("all-to-all" row; K10).  Rank g owns the
// partitions p with owner(p) = (p * world) >> radix_bits — a contiguous range,
// i.e. the TOP bits of the partition id — and receives, from every rank, that
// rank's segment of each owned partition.  The receive buffer of an owner is
// partition-major, source-rank-minor, so each owned partition is contiguous
// and the local join runs on it with the masked histogram `own_total`.
//
// exchange_cursors_kernel (one CTA), from the all-gathered histograms
// hist_all[world][nparts]:
//   src_off[p]   exclusive scan of hist_all[rank]  (+ src_off[nparts] = n local)
//   dst_start[p] where this rank's segment of p starts in owner(p)'s buffer
//   own_total[p] global size of p if this rank owns it, else 0
//   need[0]      rows this rank receives (capacity check), need[1] = 1 if that
//                exceeds `cap`: the copy kernel then drops what does not fit,
//                own_total is all zero (the join reads nothing) and the
//                caller must fail the step
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t partition_owner(uint32_t p, uint32_t world, uint32_t radix_bits) {
    return (uint32_t)(((uint64_t)p * world) >> radix_bits);
}
template <int NT>
__global__ void __launch_bounds__(NT)
exchange_cursors_kernel(const uint32_t *__restrict__ hist_all, uint32_t world, uint32_t rank, uint32_t radix_bits,
                        uint32_t cap, uint32_t *__restrict__ src_off, uint32_t *__restrict__ dst_start,
                        uint32_t *__restrict__ own_total, uint32_t *__restrict__ need) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    __shared__ uint32_t owner_base[kMaxPeers + 1];   // global prefix at the first partition of every owner
    const uint32_t nparts = 1u << radix_bits;
    const uint32_t per    = (nparts + NT - 1) / NT;
    const uint32_t first  = threadIdx.x * per;
    // pass 1: local offsets of this rank's staging buffer
    uint32_t s = 0;
    for (uint32_t k = 0; k < per; ++k)
        if (first + k < nparts) s += hist_all[rank * nparts + first + k];
    uint32_t run = block_exclusive_scan<NT>(s, warp_sums);
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            src_off[b] = run;
            run += hist_all[rank * nparts + b];
            if (b == nparts - 1) src_off[nparts] = run;
        }
    }
    // pass 2: global prefix of the partition totals; an owner's buffer starts at its first partition
    s = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts)
            for (uint32_t r = 0; r < world; ++r) s += hist_all[r * nparts + b];
    }
    const uint32_t start = block_exclusive_scan<NT>(s, warp_sums);
    const uint32_t grand = warp_sums[NT / 32];
    run = start;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            if (b == 0 || partition_owner(b, world, radix_bits) != partition_owner(b - 1, world, radix_bits))
                owner_base[partition_owner(b, world, radix_bits)] = run;
            for (uint32_t r = 0; r < world; ++r) run += hist_all[r * nparts + b];
        }
    }
    if (threadIdx.x == 0) owner_base[world] = grand;
    __syncthreads();
    // (world <= nparts, so every owner has at least one partition and has set its base)
    const uint32_t mine = owner_base[rank + 1] - owner_base[rank];
    const bool     over = mine > cap;   // the join must not read past the receive buffer: it gets nothing
    run = start;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            uint32_t tot = 0, before = 0;
            for (uint32_t r = 0; r < world; ++r) {
                const uint32_t c = hist_all[r * nparts + b];
                if (r < rank) before += c;
                tot += c;
            }
            const uint32_t o = partition_owner(b, world, radix_bits);
            dst_start[b]     = run - owner_base[o] + before;
            own_total[b]     = (o == rank && !over) ? tot : 0u;
            run += tot;
        }
    }
    if (threadIdx.x == 0) {
        need[0] = mine;
        need[1] = over ? 1u : 0u;
    }
}

// The exchange itself: the staging buffer holds this rank's shard in partition order (src_off); element e of
// partition p goes to owner(p)'s receive buffer at dst_start[p] + (e - src_off[p]).  Flat decomposition (a
// thread block takes 1024 consecutive staged elements whatever partition they belong to), so a skewed
// partition is copied by as many CTAs as it has kilo-elements.  Consecutive lanes copy consecutive elements:
// a warp's stores are 256 contiguous bytes in the peer's memory (full NVLink packets) except where a
// partition ends.  REWRITE: the row-id slot of a tuple is replaced by its position in the receive buffer, so
// that payload columns copied alongside (same index) can be read through the "row id" after the exchange.
struct ExchangeArgs {
    const uint64_t *src_tup;
    const uint64_t *src_pay[2];
    const uint32_t *src_off;     // [nparts + 1]
    const uint32_t *dst_start;   // [nparts]
    uint32_t        n, radix_bits, world, cap;
    int             npay, rewrite_rid;
    uint64_t       *dst_tup[kMaxPeers];
    uint64_t       *dst_pay[2][kMaxPeers];
};
static __global__ void __launch_bounds__(256) segment_exchange_kernel(const ExchangeArgs x) {
    constexpr int  UN     = 4;
    const uint32_t nparts = 1u << x.radix_bits;
    const uint32_t base   = blockIdx.x * (256u * UN);
    // partition of this thread's first element: last p with src_off[p] <= e
    uint32_t e = base + threadIdx.x;
    uint32_t lo = 0, hi = nparts;   // invariant: src_off[lo] <= e < src_off[hi] (src_off[nparts] = n > e)
    if (e < x.n) {
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(x.src_off + mid) <= e) lo = mid; else hi = mid;
        }
    }
    uint32_t p = lo;
#pragma unroll
    for (int u = 0; u < UN; ++u, e += 256u) {
        if (e >= x.n) break;
        while (__ldg(x.src_off + p + 1) <= e) ++p;   // empty partitions are skipped as well
        const uint32_t d   = partition_owner(p, x.world, x.radix_bits);
        const uint32_t pos = __ldg(x.dst_start + p) + (e - __ldg(x.src_off + p));
        if (pos >= x.cap) continue;                  // receive buffer too small: flagged by exchange_cursors_kernel
        uint64_t t = ld_stream_u64(x.src_tup + e);
        if (x.rewrite_rid) t = (t & 0xFFFFFFFFull) | ((uint64_t)pos << 32);
        x.dst_tup[d][pos] = t;
        for (int k = 0; k < x.npay; ++k) x.dst_pay[k][d][pos] = ld_stream_u64(x.src_pay[k] + e);
    }
}

// cursors of the histogram-free probe-side scatter: partition p starts at p * opt_cap
static __global__ void init_opt_cursors_kernel(uint32_t *__restrict__ cursor, uint32_t nparts, uint32_t opt_cap) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nparts) cursor[b] = b * opt_cap;
}

// Synthetic columns of BASELINE.json's configs, generated in HBM
// (include/b200_synth.h is the single definition shared with the CPU side).
static __global__ void __launch_bounds__(256)
synth_column_kernel(uint64_t *__restrict__ out, uint64_t first, uint64_t n, int kind, uint64_t k, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = b200_synth_value(kind, first + i, k, seed);
}

// Widening copies for the read-back entry points (tests only).
static __global__ void widen_u32_kernel(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
static __global__ void narrow_u64_kernel(const uint64_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (uint32_t)in[i];
}
template <typename TupT>
__global__ void unpack_tuples_kernel(const TupT *__restrict__ in, uint64_t n, uint64_t *__restrict__ keys,
                                     uint64_t *__restrict__ rids) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        keys[i] = in[i].key;
        rids[i] = in[i].rid;
    }
}

}  // namespace b200
