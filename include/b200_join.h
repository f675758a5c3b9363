/*
 * b200_join.h — C-ABI of the B200 join library (libb200join.so).
 *
 * Part 1 re-declares, with the reference's exact signatures, the 16 operator
 * symbols that query.c:ExecuteQuery imports (`nm -u query.o`).
 * A build that replaces the reference's rhjoin.o preprocess.o results.o
 * filter.o inter_res.o by this library and relinks the untouched handler.o
 * query.o best_tree.o stats.o scheduler.o relation_map.o relation_list.o is a
 * link-time drop-in (oracle/Makefile target `_ref/radixhash_b200_dropin`).
 *
 * Part 2 is what the shim adds: device lifecycle, relation registration
 * (the hook after relation_map.c:InitRelationMap, handler.c:52), result
 * read-back for tests, and the kernel-level entry points the parity tests and
 * bench.py call.  Plain pointers and sizes only; no torch types.
 *
 * Error convention.  Part-1 operators follow the reference: a NULL `result*`
 * means "empty => the whole query prints NULL" (query.c:360, 439); internal
 * inconsistencies and CUDA failures print to stderr and exit(2) like
 * rhjoin.c:285, filter.c:185, query.c:424 — there is no CPU fallback.
 * Part-2 functions return 0 on success, non-zero on failure with the message
 * available from b200_last_error().
 */
#ifndef B200_JOIN_H
#define B200_JOIN_H

#include "b200_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Part 1 — the reference operator API (same names, arguments, meaning)
 * ---------------------------------------------------------------------- */

/* inter_res.h:11 / inter_res.c:26 — fresh, empty intermediate with one slot
 * per query binding. */
int InitInterResults(inter_res **head, int num_of_relations);

/* inter_res.h:17 / inter_res.c:175 */
void FreeInterResults(inter_res *var);

/* filter.h:9 / filter.c:92-190 — scan `binding.column ⋄ (uint64)(int)value`
 * over the base column (emits base row ids) or through the intermediate's
 * row-id list (emits positions).  NULL when nothing qualifies. */
result *Filter(inter_res *head, filter_pred *filter_p, relation_map *map,
               int *query_relations);

/* filter.h:15 / filter.c:11-89 — install a row-id list for a fresh binding,
 * or compact every active column of the node that holds the binding. */
int InsertSingleRowIdsToInterResult(inter_res **head, int relation_num,
                                    result *res);

/* inter_res.h:33 / inter_res.c:208-231 (+ScanInterResults 182-206) — the key
 * vector a join consumes: col[i] for base rows, or col[T[b][p]] for every
 * position p of the intermediate.  Returned lazily (no AoS materialisation). */
relation *GetRelation(int given_rel, int column, inter_res *inter,
                      relation_map *map, int *query_relations);

/* rhjoin.h:9 / rhjoin.c:13-111 — all pairs (x,y) with keyR[x]==keyS[y].
 * NULL only when an input is empty (rhjoin.c:15-16); an empty join returns a
 * non-NULL result with current_load 0 (rhjoin.c:356-359). */
result *RadixHashJoin(relation *relR, relation *relS, scheduler *sched);

/* inter_res.h:26 / inter_res.c:34-152 */
int InsertJoinToInterResults(inter_res *head, int rel1, int rel2, result *res);

/* inter_res.h:61 / inter_res.c:352-361 */
int AreActiveInInter(inter_res *inter, int rel1, int rel2);

/* inter_res.h:64 / inter_res.c:363-389 — second predicate between two
 * bindings of the same node: filter on the intermediate. */
int JoinInterNode(inter_res **inter, relation_map *rel_map, int relation1,
                  int column1, int relation2, int column2, int *relations);

/* inter_res.h:48 / inter_res.c:265-318 */
void MergeInterNodes(inter_res **inter);

/* inter_res.h:68 / inter_res.c:391-428 */
void CartesianInterResults(inter_res **inter);

/* inter_res.h:55 / inter_res.c:320-339 — prints one line of SUM checksums. */
void CalculateQueryResults(inter_res *inter, relation_map *map,
                           batch_listnode *query);

/* inter_res.h:58 / inter_res.c:341-350 */
void PrintNullResults(batch_listnode *query);

/* inter_res.h:44 / inter_res.c:234-263 — same-binding column equality
 * (`0.1=0.2`).  The reference version is buggy; this one
 * implements the contract: row ids / positions with col1 == col2. */
result *SelfJoin(int given_rel, int column1, int column2, inter_res **inter,
                 relation_map *map, int *query_relations);

/* results.h:19 / results.c:144-153 */
void FreeResult(result *head);

/* preprocess.h:18 / preprocess.c:213-218 */
void FreeRelation(relation *rel);

/* ------------------------------------------------------------------------
 * Part 2 — what the shim adds
 * ---------------------------------------------------------------------- */

/* Select the CUDA device, create the memory pool.  Idempotent.  Called
 * implicitly (device 0, or $B200_DEVICE) by the first operator if omitted, so
 * the reference's unmodified handler.o still works. */
int  b200_init(int device);
void b200_shutdown(void);
/* The calling thread works on `device` from now on (b200_device_malloc, copies,
 * operators): one host thread per GPU is how a single process drives several
 * (host/b200_engine -g N).  -1 returns to the process default. */
int  b200_set_thread_device(int device);
const char *b200_last_error(void);
/* 1 when the library was built from the CUDA sources (always, for this
 * library; the CPU oracle exports the same symbol returning 0). */
int  b200_is_cuda(void);

/* Upload every column of `count` relations once (the contest's untimed
 * preparation phase).  Device copies are keyed by the host column pointer.
 * Columns that were never registered are uploaded on first use. */
int  b200_register_relations(const relation_map *map, int count);
/* relation_map.c:53-83 on the GPU: fill map[r].col_stats[j] = {l = min, u = max,
 * f = rows, d = distinct values} from the device copies (columns not yet
 * registered are uploaded first).  d is the reference's count: a marker array
 * of min(u - l + 1, 50 000 000) entries indexed by v - l, or by
 * (v - l) % 5 000 000 once the range reaches the cap.  These are the inputs of
 * stats.c / best_tree.c's join ordering; a host without the reference's
 * relation_map.o (host/b200_engine.c) gets them from here. */
int  b200_compute_column_stats(relation_map *map, int count);
/* Make an already device-resident column known under a host-side key pointer
 * (bench.py: data generated in HBM).  `max_value` may be UINT64_MAX when
 * unknown; it only selects the 32-bit-key kernels when < 2^32. */
int  b200_register_device_column(const uint64_t *host_key,
                                 const uint64_t *device_ptr, uint64_t n,
                                 uint64_t max_value);
/* (Re-)upload one column from a host buffer, synchronously on the calling
 * thread's stream; used by the end-to-end bench arm. */
int  b200_upload_column(const uint64_t *host_col, uint64_t n);
void b200_unregister_all(void);
/* Drop the device copies of these relations' columns.  Device copies are keyed
 * by the HOST column pointer, so a caller that frees or rewrites a registered
 * host column (the reference never does: its columns are a read-only mmap for
 * the life of the process, relation_map.c:24-31) must call this first. */
int  b200_unregister_relations(const relation_map *map, int count);

/* Device memory for callers that keep columns resident in HBM and pass DEVICE
 * pointers (location 1 of b200_join_sum, b200_register_device_column). */
void *b200_device_malloc(uint64_t bytes);
void  b200_device_free(void *device_ptr);
/* Preparation-phase hook (next to b200_register_relations, handler.c:52): reserve `bytes` of device memory in the
 * library's pool now, so that no query pays the driver's physical allocations (5-100 ms each on a cold process).
 * Returns the bytes reserved. */
uint64_t b200_reserve_device_memory(uint64_t bytes);
int   b200_copy_to_device(void *device_dst, const void *host_src, uint64_t bytes);
int   b200_copy_to_host(void *host_dst, const void *device_src, uint64_t bytes);

/* The calling thread's CUDA stream (cudaStream_t as void*), so a caller can
 * record events on it; b200_set_stream adopts a caller-owned stream. */
void *b200_get_stream(void);
int   b200_set_stream(void *cuda_stream);
int   b200_synchronize(void);

/* CalculateQueryResults without the printf: sums[i] for projection i and the
 * number of rows of the final intermediate. */
int  b200_calculate_sums(inter_res *inter, relation_map *map,
                         batch_listnode *query, uint64_t *sums,
                         uint64_t *num_rows);

/* Lazy last join (default on; B200_LAZY_JOIN=0 or b200_set_lazy_join(0) turns
 * it off; returns the previous setting): RadixHashJoin returns a deferred
 * result, InsertJoinToInterResults parks it on the intermediate, any operator
 * that looks at the intermediate next materialises it exactly as the eager
 * path would, and CalculateQueryResults / b200_calculate_sums run a join that
 * is still parked fused with the SUMs, so the pairs of a query's last join
 * are never written (query.c:408-461 is unchanged: it only reads
 * inter_res::next).  result::current_load of a deferred result is 0 until one
 * of the read-back calls below has materialised it. */
int  b200_set_lazy_join(int on);

/* Filter fusion (B200_FUSE_FILTERS / b200_set_fuse_filters: 0 off, 1 where it
 * pays = default, 2 wherever possible; returns the previous setting).  Mode 1
 * estimates the surviving rows from the column statistics the way stats.c does
 * and fuses when at least 2^18 rows and 1/64 of the relation survive; a more
 * selective filter is scanned the eager way when the binding is first used,
 * which leaves the join a small relation.  Filter on a base relation of at
 * least 2^18 rows that is in no intermediate yet returns a deferred result,
 * InsertSingleRowIdsToInterResult parks its predicate (up to 4 predicates on 3
 * columns per binding), GetRelation hands the join a key vector that carries
 * them, and the partition kernels evaluate them in their load stage: no scan
 * per predicate, no row-id list, no host round trip, no compaction gather
 * (query.c:337-399 unchanged).  An operator that looks at the binding before
 * a join does scans the parked predicates the eager way.  A fused filter that
 * lets nothing through still yields the reference's NULL line:
 * CalculateQueryResults prints it; after b200_calculate_sums ask
 * b200_last_result_null(). */
int  b200_set_fuse_filters(int on);
int  b200_last_result_null(void);

/* Read-back for tests: copy a result (row ids, or pairs as r[],s[]) and one
 * intermediate column to host as uint64. */
int  b200_result_kind(const result *res);               /* 1 row ids, 2 pairs */
int  b200_result_rowids_to_host(const result *res, uint64_t *out);
int  b200_result_pairs_to_host(const result *res, uint64_t *out_r,
                               uint64_t *out_s);
int  b200_inter_column_to_host(const inter_res *node, int binding,
                               uint64_t *out);

/* Kernel-level entry points, host buffers in and out (parity tests).
 * K1 scan_filter (filter.c:115-170): ids==NULL scans col[0..n) and emits row
 * ids; otherwise scans col[ids[p]] for p in [0,n_ids) and emits positions. */
int  b200_scan_filter(const uint64_t *col, uint64_t n, const uint64_t *ids,
                      uint64_t n_ids, char cmp, int value, uint64_t *out,
                      uint64_t *out_n);
/* K3-K5 radix partition (preprocess.c:13-178) on `radix_bits` low bits:
 * out_hist[b], out_psum[b] (-1 for empty buckets, preprocess.c:91-96) and the
 * partition-contiguous copy (keys and original positions). */
int  b200_radix_partition(const uint64_t *keys, uint64_t n, int radix_bits,
                          uint64_t *out_keys, uint64_t *out_row_ids,
                          uint64_t *out_hist, int64_t *out_psum);
/* K6-K7 partitioned hash join on two key vectors; writes up to `cap` pairs,
 * always returns the true pair count in *out_m. */
int  b200_hash_join_pairs(const uint64_t *keys_r, uint64_t n_r,
                          const uint64_t *keys_s, uint64_t n_s,
                          uint64_t *out_r, uint64_t *out_s, uint64_t cap,
                          uint64_t *out_m);
/* K9 checksum (inter_res.c:332-333): sum of col[ids[j]] mod 2^64. */
int  b200_gather_sum(const uint64_t *col, uint64_t n, const uint64_t *ids,
                     uint64_t m, uint64_t *out_sum);

/* Fill a DEVICE buffer with rows [first, first+n) of a synthetic column of
 * BASELINE.json's configs (kinds and parameters: include/b200_synth.h). */
int  b200_synth_column(uint64_t *device_out, uint64_t first, uint64_t n,
                       int kind, uint64_t k, uint64_t seed);

/* Tuning knobs for tests/bench (0 = library default): radix bits of the
 * partition pass and forcing the 64-bit-key kernels. */
int  b200_set_tuning(int radix_bits, int force_key64);

/* The fused bench/serving entry: `0 1|0.c=1.c|…` style two-relation equi-join
 * straight into SUM checksums — join(keys_r, keys_s) then, per projection i,
 * sum over all matching pairs of proj[i][row id of side proj_side[i]]
 * (0 = R, 1 = S).  No pair materialisation (the final join of a query is
 * folded into inter_res.c:320-339's SUM).
 *   location 0: every pointer is a HOST buffer; the call uploads them on its
 *               stream, runs the join and copies sums back (end-to-end arm).
 *   location 1: every pointer is a DEVICE buffer already resident in HBM.
 * `max_key` (or UINT64_MAX) bounds both key vectors. */
int  b200_join_sum(const uint64_t *keys_r, uint64_t n_r,
                   const uint64_t *keys_s, uint64_t n_s, uint64_t max_key,
                   int n_proj, const uint64_t *const *proj,
                   const int *proj_side, int location, uint64_t *out_sums,
                   uint64_t *out_matches);

/* ---- staged join on caller-owned DEVICE buffers (multi-GPU plans) ----------
 * One process per GPU drives the phases of rhjoin.c:13-111 separately so that
 * the exchange between GPUs can sit between them (sigmod-2018_b200/sharding.py):
 *   hist      counts of key & (2^bits - 1)              (preprocess.c:181-195)
 *   scatter   into partition order at caller-computed cursors; the build shard
 *             is partitioned locally and each partition segment is then copied
 *             to d_dst_start[p] inside up to 8 destination buffers — this
 *             GPU's and its peers' IPC-mapped ones (broadcast over NVLink with
 *             256-byte stores) — together with up to two payload columns
 *   join_sum  per-partition build + probe + SUM on this GPU's buffers
 * 32-bit keys only (all keys < 2^32); row ids are rid_base + position.
 * b200_ipc_* wrap cudaIpcGetMemHandle / OpenMemHandle for buffers obtained from
 * b200_device_malloc. */
int   b200_ipc_export(const void *device_ptr, unsigned char *out_handle64);
void *b200_ipc_import(const unsigned char *handle64);
int   b200_ipc_close(void *imported_ptr);
int   b200_radix_bits_for(uint64_t n_build);
int   b200_stage_hist(const uint64_t *d_keys, uint64_t n, int radix_bits,
                      uint32_t *d_hist);
int   b200_stage_scatter_build(const uint64_t *d_keys, uint64_t n,
                               uint32_t rid_base, int radix_bits,
                               const uint32_t *d_hist_local,
                               const uint32_t *d_dst_start, int ndst,
                               void *const *tup_dst, int npay,
                               const uint64_t *const *pay_cols,
                               uint64_t *const *pay_dst, int phase);
/* npay == 1 with pay_dst == NULL: the one payload column holds 32-bit values
 * and travels in the row-id slot of the build tuples (no payload buffers; pass
 * (const uint64_t *)1 as that projection's proj_part_vals to the join).
 * phase 0 = partition + broadcast; 1 = only the local partition pass (staged
 * inside the library); 2 = only the broadcast of what phase 1 staged, so that a
 * caller can put other work (the probe-side scatter, on another stream)
 * between the two. */
int   b200_stage_scatter_probe(const uint64_t *d_keys, uint64_t n,
                               int radix_bits, uint32_t *d_cursor,
                               void *d_tup_out);
/* Histogram-free probe side: 2^bits regions of b200_opt_region_cap() tuples in
 * d_tup_out; what does not fit goes to d_ov (n tuples) / *d_ovcnt and is
 * partitioned exactly inside b200_stage_join_sum. */
uint32_t b200_opt_region_cap(uint64_t n_probe, int radix_bits);
int   b200_stage_scatter_probe_opt(const uint64_t *d_keys, uint64_t n,
                                   int radix_bits, uint32_t opt_cap,
                                   uint32_t *d_cursor, void *d_tup_out,
                                   void *d_ov, uint32_t *d_ovcnt);
/* The same with a probe-side SUM column whose values fit 32 bits carried in
 * the row-id slot of the probe tuples (streamed by the scatter instead of
 * gathered per match by the join): pass that projection to the join with
 * proj_part_vals = (uint64_t*)1. */
int   b200_stage_scatter_probe_opt_carry(const uint64_t *d_keys, uint64_t n,
                                         int radix_bits, uint32_t opt_cap,
                                         uint32_t *d_cursor, void *d_tup_out,
                                         void *d_ov, uint32_t *d_ovcnt,
                                         const uint64_t *d_carry_col);
/* d_hist_p is the probe histogram (opt_cap == 0) or the cursor array that
 * b200_stage_scatter_probe_opt left behind (opt_cap > 0, with d_ov/d_ovcnt). */
int   b200_stage_join_sum(const void *d_tup_b, const uint32_t *d_hist_b,
                          const void *d_tup_p, const uint32_t *d_hist_p,
                          int radix_bits, int n_proj,
                          const uint64_t *const *proj_cols,
                          const int *proj_side,
                          const uint64_t *const *proj_part_vals,
                          uint32_t opt_cap, const void *d_ov,
                          const uint32_t *d_ovcnt,
                          uint64_t *out_sums, uint64_t *out_matches);

/* From the all-gathered per-rank build histograms hist_all[world][2^bits]:
 * the global histogram and this rank's start inside every partition. */
int   b200_stage_build_cursors(const uint32_t *d_hist_all, int world, int rank,
                               int radix_bits, uint32_t *d_total,
                               uint32_t *d_my_start);
/* b200_stage_join_sum without the host read-back: d_result (DEVICE, n_proj + 2
 * u64) receives {matches, sums..., overflow count} on the stream, so the
 * caller can all-reduce it in place.  A non-zero overflow count means the
 * overflow of the histogram-free scatter still has to be joined: call
 * b200_stage_join_sum (synchronous) instead for that step. */
int   b200_stage_join_sum_async(const void *d_tup_b, const uint32_t *d_hist_b,
                                const void *d_tup_p, const uint32_t *d_hist_p,
                                int radix_bits, int n_proj,
                                const uint64_t *const *proj_cols,
                                const int *proj_side,
                                const uint64_t *const *proj_part_vals,
                                uint32_t opt_cap, const void *d_ov,
                                const uint32_t *d_ovcnt, uint64_t *d_result);

/* Rank-major build layout (the broadcast runs on the copy engines): every rank
 * partitions its build shard straight into region `rank` of its own build
 * buffer (b200_stage_scatter_build_local), copies that region verbatim into the
 * same region of every peer's buffer (b200_copy_device_async on IPC-mapped
 * pointers: one large copy per peer), and the join reads partition p as nseg
 * runs: d_hist_all[nseg][2^bits] are the all-gathered histograms, region r
 * starts at r * seg_rows.  d_result != NULL: asynchronous, {matches, sums...,
 * overflow count} written to that DEVICE buffer; else synchronous into
 * out_sums / out_matches (overflow pass included). */
int   b200_stage_scatter_build_local(const uint64_t *d_keys, uint64_t n,
                                     uint32_t rid_base, int radix_bits,
                                     const uint32_t *d_hist_local,
                                     void *d_tup_out, int npay,
                                     const uint64_t *const *pay_cols,
                                     uint64_t *const *pay_out);
int   b200_copy_device_async(void *dst, const void *src, uint64_t bytes);
int   b200_stage_join_sum_seg(const void *d_tup_b, const uint32_t *d_hist_all,
                              int nseg, uint32_t seg_rows, const void *d_tup_p,
                              const uint32_t *d_hist_p, int radix_bits,
                              int n_proj, const uint64_t *const *proj_cols,
                              const int *proj_side,
                              const uint64_t *const *proj_part_vals,
                              uint32_t opt_cap, const void *d_ov,
                              const uint32_t *d_ovcnt, uint64_t *d_result,
                              uint64_t *out_sums, uint64_t *out_matches);

/* Radix-sharded exchange (the all-to-all plan of SURVEY 8e; absent in the
 * reference, which is single-process): rank g owns the partitions p with
 * (p * world) >> radix_bits == g and receives every rank's segment of them;
 * an owner's receive buffer is partition-major, source-rank-minor.
 * b200_stage_exchange_cursors: from the all-gathered histograms
 * d_hist_all[world][2^bits], d_src_off[2^bits + 1] = offsets of this rank's
 * locally partitioned shard (b200_stage_scatter_build_local), d_dst_start[p] =
 * where its segment of p starts in the owner's buffer, d_own_total[p] = global
 * size of p if this rank owns it else 0 (the histogram the local join runs
 * on; all zero when the capacity is exceeded), d_need[0] = rows this rank
 * receives, d_need[1] = 1 if that exceeds cap (the caller must fail the step).
 * b200_stage_exchange_segments: copies every staged tuple (and up to two
 * payload columns, pay_dst[k * world + d]) to tup_dst[owner] with stores of
 * 256 contiguous bytes per warp (peer buffers are CUDA-IPC mappings: NVLink);
 * positions >= cap are dropped.  rewrite_rid: the row-id slot of a tuple
 * becomes its position in the receive buffer, so that payload columns copied
 * alongside are addressed by it.  A projection whose 32-bit value travels in
 * the row-id slot is passed to the join with proj_part_vals = (uint64_t*)1 on
 * either side. */
int   b200_stage_exchange_cursors(const uint32_t *d_hist_all, int world,
                                  int rank, int radix_bits, uint32_t cap,
                                  uint32_t *d_src_off, uint32_t *d_dst_start,
                                  uint32_t *d_own_total, uint32_t *d_need);
int   b200_stage_exchange_segments(const void *d_src_tup, int npay,
                                   const uint64_t *const *src_pay, uint64_t n,
                                   int radix_bits, int world,
                                   const uint32_t *d_src_off,
                                   const uint32_t *d_dst_start, uint32_t cap,
                                   int rewrite_rid, void *const *tup_dst,
                                   uint64_t *const *pay_dst);

/* ---- multi-GPU plans behind the C ABI: no torch, no NCCL ---------------------
 * The join of rhjoin.c:13-111 shards by radix bucket (rhjoin.c:42-57: one
 * JoinJob per bucket pair, nothing shared between buckets).  One rank drives
 * one GPU: a process per GPU (peers' memory through CUDA IPC: b200_multi_export
 * / b200_multi_connect_ipc) or a host thread per GPU inside one process
 * (b200_multi_connect_ptr, b200_join_sum_multi).  Ranks synchronise through
 * epoch flags in each other's memory; a step enqueues device work only and is
 * captured in a CUDA graph when there are several ranks ($B200_MULTI_GRAPH=0/1).
 *   B200_PLAN_BROADCAST  small build side (config 2): every rank partitions its
 *       build shard once into its region of a rank-major build buffer
 *       (histogram in the region's head), its copy engines push the region to
 *       every peer with one copy and one flag each, the probe shard is
 *       partitioned locally meanwhile and never moves, the join waits per
 *       partition for the regions holding its runs.  ($B200_BCAST=pull: every
 *       rank fetches its peers' regions with a TMA kernel instead; slower.)
 *   B200_PLAN_EXCHANGE   radix-sharded all-to-all (config 4): both shards are
 *       partitioned locally (the probe shard in chunks), the exchange kernel
 *       stores every partition into its owner's receive buffer over NVLink (the
 *       exchange of chunk c under the partition pass of chunk c + 1); owners are
 *       contiguous partition ranges cut on the global histogram so that skewed
 *       keys do not overload one GPU; every owner joins what it received.  Hot
 *       keys (a sample of the probe keys decides) are not exchanged at all: every
 *       rank learns the count and SUM of the build rows carrying them, and probe
 *       rows with a hot key are joined where they are, during the histogram pass.
 * Query shape: join(build.key = probe.key) with SUM(build column) and / or
 * SUM(probe column); keys and SUM values below 2^32 (they travel in 8-byte
 * tuples).  Every rank must create its plan with the same totals / maxima. */
#define B200_PLAN_BROADCAST 0
#define B200_PLAN_EXCHANGE  1
typedef struct b200_multi b200_multi;
typedef struct {
    int      plan, rank, world, device;
    uint64_t n_build_total, n_probe_total;         /* rows over all ranks        */
    uint64_t n_build_local, n_probe_local;         /* this rank's position shard */
    uint64_t n_build_local_max, n_probe_local_max; /* largest shard of any rank  */
    int      has_build_sum, has_probe_sum;
    int      radix_bits;                           /* 0 = automatic              */
    int      chunks;                               /* 0 = default (4 copy chunks / 8 probe chunks) */
    uint64_t recv_rows_build, recv_rows_probe;     /* exchange: receive capacity per rank, 0 = mean + 1/8 */
    int      hot_keys;                             /* exchange: 0 = hot keys are joined where they are (default), -1 = off */
} b200_multi_config;
b200_multi *b200_multi_create(const b200_multi_config *cfg);
void  b200_multi_destroy(b200_multi *plan);
int   b200_multi_export(b200_multi *plan, unsigned char *out_handle64);
void *b200_multi_shared_ptr(b200_multi *plan);
int   b200_multi_connect_ipc(b200_multi *plan, int peer, const unsigned char *handle64);
int   b200_multi_connect_ptr(b200_multi *plan, int peer, void *peer_shared, int peer_device);
int   b200_multi_radix_bits(b200_multi *plan);
/* One step on the calling thread's stream, DEVICE pointers to this rank's
 * shards.  phases: 0 = the whole step; otherwise a bit mask of the step's
 * phases (broadcast: 1 partition + broadcast, 2 join + publish, 4 reduce;
 * exchange: 1 build histogram + hot-key candidates, 2 hot-key table + build-side
 * aggregates, 4 probe histograms, 8 partition + exchange, 16 join + publish,
 * 32 reduce)
 * so that a test can drive several ranks on ONE GPU phase by phase — kernels
 * that wait on one another must never share a GPU. */
int   b200_multi_enqueue(b200_multi *plan, const uint64_t *d_build_keys, const uint64_t *d_build_sum,
                         const uint64_t *d_probe_keys, const uint64_t *d_probe_sum, int phases);
/* Synchronise; out_sums = {SUM(build column)?, SUM(probe column)?} over ALL
 * ranks (the same on every rank).  Non-zero: a peer timed out, or the exchange
 * receive buffers were too small for this input. */
int   b200_multi_finish(b200_multi *plan, uint64_t *out_sums, uint64_t *out_matches);
int   b200_multi_received(b200_multi *plan, uint64_t *out_rows_build_probe);
/* One process, one host thread per GPU (devices 0 .. n_gpus-1), DEVICE-resident
 * position shards [g]; `steps` > 0 additionally times that many steps and
 * returns the slowest rank's milliseconds per step. */
int   b200_join_sum_multi(int n_gpus, int plan_kind, const uint64_t *const *d_build_keys,
                          const uint64_t *const *d_build_sum, const uint64_t *n_build,
                          const uint64_t *const *d_probe_keys, const uint64_t *const *d_probe_sum,
                          const uint64_t *n_probe, int steps, uint64_t *out_sums, uint64_t *out_matches,
                          double *out_ms);

/* Per-kernel device times of the calling thread's last RadixHashJoin /
 * b200_join_sum, measured with CUDA events on its stream when profiling is
 * enabled with b200_set_profiling(1).  Names: "hist_b", "hist_p", "scan",
 * "scatter_b", "scatter_p", "scatter_pc" (the probe scatter when it also carries a
 * SUM column), "join", "exchange" (b = build side, p = probe side).  Returns milliseconds, or a negative value if the
 * kernel did not run. */
int    b200_set_profiling(int on);
/* ... summed over every scope of that name since profiling was last enabled (out_scopes: how many) */
double b200_sum_kernel_ms(const char *name, int *out_scopes);
double b200_last_kernel_ms(const char *name);
/* Number of kernels this library launched since the counter was reset. */
uint64_t b200_kernel_launches(int reset);

#ifdef __cplusplus
}
#endif
#endif /* B200_JOIN_H */
