// kernels.cuh — hand-written sm_100a kernels of the join hot path.
//
// Reference loops replaced (file:line in /root/reference):
//   K1  filter.c:115-170            -> scan_filter_kernel
//   K2  inter_res.c:200-204,223-227 -> fused away: KeySrc is a lazy (column,
//                                      row-id list) view, never an AoS copy
//   K3  preprocess.c:189-192        -> radix_hist_kernel
//   K4  preprocess.c:83-102         -> partition_plan_kernel
//   K5  preprocess.c:262-296,350-359-> radix_scatter_kernel
//   K6  rhjoin.c:227-248,270-271    -> hash_join_kernel, build phase
//   K7  rhjoin.c:154-216            -> hash_join_kernel, probe phase
//   K8  inter_res.c:79-98,119-137;
//       filter.c:60-76; inter_res.c:304-313 -> gather_columns_kernel
//   K9  inter_res.c:332-333         -> checksum_kernel / hash_join_kernel<SUM>
//   K11 inter_res.c:376-385         -> inter_equal_kernel
//
// Everything is HBM-bound integer work: no tensor cores.  Row ids and
// positions are 32-bit on the device; keys are 32-bit when the column maximum
// allows it (8-byte partition tuples) and 64-bit otherwise (16-byte tuples).
#pragma once

#include "types.cuh"
#include "../../include/b200_synth.h"

#include <cuda_runtime.h>

#include <cstdint>

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

template <int NT, int U, typename KeyT>
__device__ __forceinline__ void load_tile_keys(const KeySrc &src, uint64_t base, uint32_t count,
                                               bool vec, KeyT (&keys)[U]) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < U; j += 2) {
            const uint32_t li = ((uint32_t)((j >> 1) * NT + (int)threadIdx.x)) * 2u;
            ulonglong2     v  = ld_stream_u64x2(src.col + base + li);
            keys[j]           = (KeyT)v.x;
            keys[j + 1]       = (KeyT)v.y;
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
template <int NT, int U, typename KeyT>
__global__ void __launch_bounds__(NT)
radix_hist_kernel(KeySrc src, uint32_t radix_bits, uint32_t *__restrict__ ghist) {
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
            if (tile_local_index<NT, U>(j, vec) < count)
                atomicAdd(&sh_hist[(uint32_t)keys[j] & mask], 1u);
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
template <int NT>
__global__ void __launch_bounds__(NT)
partition_plan_kernel(const uint32_t *__restrict__ hist_b, const uint32_t *__restrict__ hist_p,
                      uint32_t nparts, uint32_t cap, uint32_t slice, uint32_t *__restrict__ off_b,
                      uint32_t *__restrict__ off_p, uint32_t *__restrict__ cur_b,
                      uint32_t *__restrict__ cur_p, uint32_t *__restrict__ item_start) {
    __shared__ uint32_t warp_sums[NT / 32 + 1];
    const uint32_t      per   = (nparts + NT - 1) / NT;
    const uint32_t      first = threadIdx.x * per;
    uint32_t            sb = 0, sp = 0, si = 0;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t b = first + k;
        if (b < nparts) {
            const uint32_t cb = hist_b[b], cp = hist_p[b];
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
            const uint32_t cb = hist_b[b], cp = hist_p[b];
            off_b[b] = eb;
            cur_b[b] = eb;
            off_p[b] = ep;
            cur_p[b] = ep;
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
// Per tile of NT*U keys: (1) coalesced key loads, (2) shared-memory atomics
// give every tuple its rank inside its partition, (3) a block scan turns the
// tile histogram into local offsets and reserves the tile's run in every
// partition with ONE global atomicAdd per non-empty partition, (4) tuples are
// written to shared memory in partition order, (5) copied out so that a warp
// stores consecutive addresses inside each run.  All CTAs advance the same
// 2^bits cursors, so the write frontier is a few hundred KB and partially
// written sectors are completed in L2 before they reach HBM.
// Order inside a partition is not the reference's (stable) order; only the
// multiset matters downstream.
// ---------------------------------------------------------------------------
template <int NT, int U, typename KeyT>
__global__ void __launch_bounds__(NT, 2)
radix_scatter_kernel(KeySrc src, uint32_t radix_bits, uint32_t *__restrict__ cursor,
                     typename TupOf<KeyT>::type *__restrict__ out) {
    using TupT = typename TupOf<KeyT>::type;
    constexpr uint32_t TILE = NT * U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TupT     *stage  = reinterpret_cast<TupT *>(smem_raw);
    uint32_t *cnt    = reinterpret_cast<uint32_t *>(stage + TILE);
    const uint32_t nbins = 1u << radix_bits;
    const uint32_t mask  = nbins - 1u;
    uint32_t *loc    = cnt + nbins;
    uint32_t *gdelta = loc + nbins;
    __shared__ uint32_t warp_sums[NT / 32 + 1];

    const uint64_t n      = src.n;
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    const bool     vec_ok = src.ids == nullptr && ((reinterpret_cast<uintptr_t>(src.col) & 15) == 0);
    const uint32_t per    = (nbins + NT - 1) / NT;

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t base  = tile * TILE;
        const uint32_t count = (uint32_t)min((uint64_t)TILE, n - base);
        const bool     vec   = vec_ok && count == TILE;
        for (uint32_t b = threadIdx.x; b < nbins; b += NT) cnt[b] = 0;
        KeyT keys[U];
        load_tile_keys<NT, U, KeyT>(src, base, count, vec, keys);
        __syncthreads();
        uint16_t rank[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (tile_local_index<NT, U>(j, vec) < count)
                rank[j] = (uint16_t)atomicAdd(&cnt[(uint32_t)keys[j] & mask], 1u);
        }
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
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t li = tile_local_index<NT, U>(j, vec);
            if (li < count) {
                TupT t;
                t.key = keys[j];
                t.rid = (uint32_t)(base + li);
                if constexpr (sizeof(KeyT) == 8) t.pad = 0;
                stage[loc[(uint32_t)keys[j] & mask] + rank[j]] = t;
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < count; i += NT) {
            const TupT t = stage[i];
            out[gdelta[(uint32_t)t.key & mask] + i] = t;
        }
        __syncthreads();
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
// so duplicates on both sides yield the cross product.
//   MODE_COUNT: count matches per item (sizes the pair output exactly)
//   MODE_WRITE: materialise (build row id, probe row id) pairs
//   MODE_SUM  : fold the pairs straight into the SUM checksums
//               (inter_res.c:320-339) — the final join of a query never
//               materialises its output.
// DIRECT = no partition pass: one "partition" read lazily through KeySrc;
// used when the build side already fits a shared-memory table.
// ---------------------------------------------------------------------------
enum JoinMode { MODE_COUNT = 0, MODE_WRITE = 1, MODE_SUM = 2 };

struct JoinArgs {
    KeySrc          src_b, src_p;   // DIRECT
    const void     *tup_b, *tup_p;  // partitioned
    const uint32_t *off_b, *off_p, *item_start;
    uint32_t        nparts, radix_bits, cap, slice, slots_log2;
    uint32_t        n_items_direct, sc_direct;
    uint32_t       *work_counter;
    unsigned long long *total;        // matches (COUNT, SUM)
    unsigned long long *item_count;   // per item: written by COUNT, read by WRITE
    unsigned long long *out_cursor;   // WRITE
    uint32_t       *out_b, *out_p;    // WRITE
    int             nproj;            // SUM
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

template <int NT, int U, typename KeyT, bool DIRECT, int MODE>
__global__ void __launch_bounds__(NT)
hash_join_kernel(const JoinArgs a) {
    using TupT = typename TupOf<KeyT>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyT     *t_keys  = reinterpret_cast<KeyT *>(smem_raw);
    uint32_t *t_rids  = reinterpret_cast<uint32_t *>(t_keys + a.cap);
    uint16_t *t_next  = reinterpret_cast<uint16_t *>(t_rids + a.cap);
    uint16_t *t_heads = t_next + a.cap;
    const uint32_t nslots = 1u << a.slots_log2;

    __shared__ uint32_t           s_item[5];   // valid, b_start, b_count, p_start, p_count
    __shared__ uint32_t           s_cursor;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_cnt;

    const int lane = threadIdx.x & 31;
    const TupT *tup_b = static_cast<const TupT *>(a.tup_b);
    const TupT *tup_p = static_cast<const TupT *>(a.tup_p);

    unsigned long long my_matches = 0;
    unsigned long long my_sum[kMaxProj];
#pragma unroll
    for (int k = 0; k < kMaxProj; ++k) my_sum[k] = 0;

    for (;;) {
        // ---- fetch one work item (warp 0) ---------------------------------
        if (threadIdx.x < 32) {
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
                    const uint32_t p0  = a.off_p[p], p1 = a.off_p[p + 1];
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
                s_cnt     = 0ull;
                s_cursor  = 0u;
            }
            if constexpr (MODE == MODE_WRITE) {
                if (lane == 0 && valid) s_base = atomicAdd(a.out_cursor, a.item_count[w]);
            }
            if constexpr (MODE == MODE_COUNT) {
                if (lane == 0) s_item[0] = valid | (w << 1);   // keep w for the per-item count
            }
        }
        __syncthreads();
        const uint32_t item_word = s_item[0];
        if ((item_word & 1u) == 0u) break;
        const uint32_t b_start = s_item[1], b_count = s_item[2];
        const uint32_t p_start = s_item[3], p_count = s_item[4];

        // ---- build (K6) ---------------------------------------------------
        for (uint32_t s = threadIdx.x; s < nslots; s += NT) t_heads[s] = (uint16_t)kEmpty16;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < b_count; i += NT) {
            KeyT     key;
            uint32_t rid;
            if constexpr (DIRECT) {
                rid = b_start + i;
                key = (KeyT)(a.src_b.ids ? a.src_b.col[a.src_b.ids[rid]] : a.src_b.col[rid]);
            } else {
                const TupT t = tup_b[b_start + i];
                key          = t.key;
                rid          = t.rid;
            }
            t_keys[i] = key;
            t_rids[i] = rid;
            const uint32_t  h    = table_hash<KeyT>(key, a.radix_bits, a.slots_log2);
            unsigned short *slot = reinterpret_cast<unsigned short *>(&t_heads[h]);
            unsigned short  old  = *slot, assumed;
            do {
                assumed = old;
                old     = atomicCAS(slot, assumed, (unsigned short)i);
            } while (old != assumed);
            t_next[i] = old;
        }
        __syncthreads();

        // ---- probe (K7) ---------------------------------------------------
        unsigned long long item_matches = 0;
        for (uint32_t off = 0; off < p_count; off += NT * U) {
            KeyT     pkey[U];
            uint32_t prid[U];
            bool     pval[U];
            if constexpr (DIRECT) {
                if (a.src_p.ids) {
                    uint32_t id[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        const uint32_t li = off + j * NT + threadIdx.x;
                        pval[j]           = li < p_count;
                        prid[j]           = p_start + li;
                        id[j]             = pval[j] ? ld_stream_u32(a.src_p.ids + prid[j]) : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        pkey[j] = pval[j] ? (KeyT)__ldg(a.src_p.col + id[j]) : (KeyT)0;
                } else {
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        const uint32_t li = off + j * NT + threadIdx.x;
                        pval[j]           = li < p_count;
                        prid[j]           = p_start + li;
                        pkey[j] = pval[j] ? (KeyT)ld_stream_u64(a.src_p.col + prid[j]) : (KeyT)0;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const uint32_t li = off + j * NT + threadIdx.x;
                    pval[j]           = li < p_count;
                    if (pval[j]) {
                        if constexpr (sizeof(KeyT) == 8) {
                            const ulonglong2 v = ld_stream_u64x2(
                                reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                            pkey[j] = v.x;
                            prid[j] = (uint32_t)v.y;
                        } else {
                            const uint64_t v =
                                ld_stream_u64(reinterpret_cast<const uint64_t *>(tup_p + p_start + li));
                            pkey[j] = (uint32_t)v;
                            prid[j] = (uint32_t)(v >> 32);
                        }
                    } else {
                        pkey[j] = 0;
                        prid[j] = 0;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                uint32_t idx = kEmpty16;
                if (pval[j]) idx = t_heads[table_hash<KeyT>(pkey[j], a.radix_bits, a.slots_log2)];
                // warp-synchronous chain walk: every lane stays in the loop
                // until the longest chain of the warp is exhausted
                while (__any_sync(kFullMask, idx != kEmpty16)) {
                    const bool live = idx != kEmpty16;
                    const bool hit  = live && t_keys[live ? idx : 0] == pkey[j];
                    if constexpr (MODE == MODE_COUNT) {
                        item_matches += hit ? 1ull : 0ull;
                    } else if constexpr (MODE == MODE_SUM) {
                        if (hit) {
                            ++my_matches;
                            const uint32_t brid = t_rids[idx];
#pragma unroll
                            for (int k = 0; k < kMaxProj; ++k) {
                                if (k < a.nproj) {
                                    const uint32_t r  = a.proj[k].side == 0 ? brid : prid[j];
                                    const uint32_t rr = a.proj[k].ids ? __ldg(a.proj[k].ids + r) : r;
                                    my_sum[k] += __ldg(a.proj[k].col + rr);
                                }
                            }
                        }
                    } else {
                        const unsigned hits = __ballot_sync(kFullMask, hit);
                        if (hits) {
                            const int leader = __ffs(hits) - 1;
                            uint32_t  basepos = 0;
                            if (lane == leader) basepos = atomicAdd(&s_cursor, (uint32_t)__popc(hits));
                            basepos = __shfl_sync(kFullMask, basepos, leader);
                            if (hit) {
                                const unsigned long long pos =
                                    s_base + basepos + __popc(hits & ((1u << lane) - 1u));
                                a.out_b[pos] = t_rids[idx];
                                a.out_p[pos] = prid[j];
                            }
                        }
                    }
                    if (live) idx = t_next[idx];
                }
            }
        }
        if constexpr (MODE == MODE_COUNT) {
            const unsigned long long ws = warp_sum_u64(item_matches);
            if (lane == 0 && ws) atomicAdd(&s_cnt, ws);
            __syncthreads();
            if (threadIdx.x == 0) {
                a.item_count[item_word >> 1] = s_cnt;
                if (s_cnt) atomicAdd(a.total, s_cnt);
            }
        }
        __syncthreads();   // table and s_item are reused by the next item
    }

    if constexpr (MODE == MODE_SUM) {
        // K9: warp-shuffle reduction, one atomicAdd(u64) per warp and projection
        const unsigned long long wm = warp_sum_u64(my_matches);
        if (lane == 0 && wm) atomicAdd(a.total, wm);
#pragma unroll
        for (int k = 0; k < kMaxProj; ++k) {
            if (k < a.nproj) {
                const unsigned long long ws = warp_sum_u64(my_sum[k]);
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

__global__ void __launch_bounds__(256) gather_columns_kernel(const GatherArgs g) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.m; i += stride) {
        const uint32_t p = ld_stream_u32(g.pos + i);
#pragma unroll
        for (int c = 0; c < kMaxGather; ++c)
            if (c < g.ncols) g.out[c][i] = __ldg(g.in[c] + p);
    }
}

// CartesianInterResults (inter_res.c:405-418): row index = i * n2 + j.
__global__ void __launch_bounds__(256)
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

__global__ void __launch_bounds__(256) checksum_kernel(const ChecksumArgs c) {
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

// Column maximum at registration (selects the 32-bit-key kernels).
__global__ void __launch_bounds__(256)
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

// Synthetic columns of BASELINE.json's configs, generated in HBM
// (include/b200_synth.h is the single definition shared with the CPU side).
__global__ void __launch_bounds__(256)
synth_column_kernel(uint64_t *__restrict__ out, uint64_t first, uint64_t n, int kind, uint64_t k, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = b200_synth_value(kind, first + i, k, seed);
}

// Widening copies for the read-back entry points (tests only).
__global__ void widen_u32_kernel(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
__global__ void narrow_u64_kernel(const uint64_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out) {
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
