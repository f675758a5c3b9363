// engine.cuh — host-side orchestration of the join pipeline on one GPU:
// column registry (device copies keyed by the host column pointer), lazy key
// vectors, radix partition + shared-memory hash join in its three output
// modes, compaction scans, gathers and checksums.  Everything runs on the
// calling thread's stream (common.cuh).
#pragma once

#include "common.cuh"
#include "types.cuh"

#include <vector>

namespace b200 {

struct DevColumn {
    const uint64_t *d       = nullptr;
    uint64_t        n       = 0;
    uint64_t        max_val = UINT64_MAX;
    // relation_map.c:53-83's other statistics, computed on the GPU when the column is uploaded (0 / 0: unknown,
    // e.g. a caller-registered device column): smallest value and number of distinct values (the reference's count)
    uint64_t        min_val  = 0;
    uint64_t        distinct = 0;
};
// min, max and the reference's distinct count of a DEVICE column (relation_map.c:53-83), on the calling thread's stream
void device_column_stats(const uint64_t *d_col, uint64_t n, uint64_t *out_min, uint64_t *out_max, uint64_t *out_distinct);

// Device copy of a host column (relation_map.columns[j]); uploads on a miss.
DevColumn lookup_column(const uint64_t *host_col, uint64_t n);
void      register_host_column(const uint64_t *host_col, uint64_t n, bool replace);
void      register_device_column(const uint64_t *host_key, const uint64_t *dev, uint64_t n,
                                 uint64_t max_val);
void      unregister_all_columns();
uint64_t  known_column_max(const uint64_t *dev_ptr);
void      unregister_column(const uint64_t *host_col);

// Lazy key vector with ownership of the row-id list it reads through.
struct KeyVec {
    KeySrc    src{nullptr, nullptr, 0};
    DevBufPtr ids_owner;
    uint64_t  max_val = UINT64_MAX;
    // filter predicates of this (base) relation evaluated inside the partition kernels' load stage instead of by a
    // scan + row-id list: rows that fail do not take part; row ids stay base row ids (ids == nullptr)
    PredSet   preds;
};

struct Tuning {
    int      radix_bits  = 0;   // 0 = automatic
    int      force_key64 = 0;
    uint32_t cap32       = 17408;    // build tuples per shared-memory table, 32-bit keys
    uint32_t cap64       = 12288;    // ... 64-bit keys
    uint32_t slice       = 1u << 18; // probe tuples per work item
    int      carry32     = 1;        // a single 32-bit build-side SUM column travels in the tuple's row-id slot
    int      carry_probe = 1;        // ... and a single 32-bit probe-side SUM column in the probe tuples', when the
                                     // expected matches make streaming it cheaper than gathering it (run_join)
    int      opt_partition = 1;      // histogram-free probe-side scatter for the fused join -> SUM
    int      early_mat   = 1;        // carry build-side SUM projections through the scatter
    int      scatter_cfg = 1;        // see engine.cu PartCfg
    int      max_bits    = 12;
    int      tag64       = 1;        // partitioned 64-bit-key joins: 1 = tag table with verified candidates (histogram-free
                                     // probe side, carried SUM values), 0 = chained table with the keys in shared memory
    int      debug       = 0;
};
Tuning &tuning();

enum class JoinOut { Pairs, Sum };

struct JoinResult {
    // rows of R / S that passed their fused predicates (UINT64_MAX: that side had none).  Zero means the filter
    // was empty: the reference prints NULL for the whole query then (query.c:360-369)
    uint64_t  valid_r = UINT64_MAX, valid_s = UINT64_MAX;
    uint64_t  m = 0;              // number of matching pairs
    DevBufPtr r_ids, s_ids;       // Pairs: 32-bit row ids, m each
    uint64_t  sums[kMaxProj] = {0};
};

// proj[k].side is relative to the call: 0 = R side, 1 = S side.
JoinResult run_join(const KeyVec &R, const KeyVec &S, JoinOut mode, int nproj, const ProjDesc *proj);

// Radix partition only (tests): tuples in partition order, hist and offsets.
struct PartitionOut {
    DevBufPtr tuples;    // Tup32 or Tup64
    DevBufPtr hist;      // u32[nparts]
    bool      key64 = false;
};
PartitionOut run_partition(const KeyVec &src, int radix_bits);

// K1 / K11 compaction scans: returns a u32 list and its length.
struct IdList {
    DevBufPtr ids;
    uint64_t  n = 0;
};
IdList run_filter(const KeySrc &src, char cmp, int value);
IdList run_filter_u64(const KeySrc &src, int cmp_code, uint64_t constant);   // cmp_code: 0 '<', 1 '>', 2 '='
IdList run_inter_equal(const uint64_t *col_a, const uint32_t *ta, const uint64_t *col_b,
                       const uint32_t *tb, uint64_t n);

// K8: out[c] = in[c][pos[i]] for every listed column.
std::vector<DevBufPtr> run_gather(const uint32_t *pos, uint64_t m, const std::vector<const uint32_t *> &in);
// K9
void run_checksum(uint64_t m, int nproj, const uint64_t *const *cols, const uint32_t *const *ids,
                  uint64_t *out_sums);

// CartesianInterResults (inter_res.c:405-418): out[i*n2+j] = in[i] or in[j].
void run_cartesian(const uint32_t *in, uint64_t n1, uint64_t n2, bool from_first, uint32_t *out);
void run_synth_column(uint64_t *d_out, uint64_t first, uint64_t n, int kind, uint64_t k, uint64_t seed);
// width conversions between host-facing uint64 ids and device uint32 ids
void widen_ids(const uint32_t *d_in, uint64_t n, uint64_t *d_out);
void narrow_ids(const uint64_t *d_in, uint64_t n, uint32_t *d_out);
void unpack_partition(const PartitionOut &p, uint64_t n, uint64_t *d_keys, uint64_t *d_rids);

// staged join on caller-owned device buffers (multi-GPU plans), 32-bit keys
void       stage_hist(const uint64_t *d_keys, uint64_t n, int bits, uint32_t *d_hist);
void       stage_scatter_build(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int bits,
                               const uint32_t *d_hist_local, const uint32_t *d_dst_start, int ndst,
                               void *const *tup_dst, int npay, const uint64_t *const *pay_cols,
                               uint64_t *const *pay_dst, int phase = 0);
void       stage_scatter_probe(const uint64_t *d_keys, uint64_t n, int bits, uint32_t *d_cursor, void *d_tup_out);
uint32_t   opt_region_cap(uint64_t n_probe, int bits);
int        auto_radix_bits(uint64_t n_build, bool key64, bool chained64 = false);
void       stage_scatter_probe_opt(const uint64_t *d_keys, uint64_t n, int bits, uint32_t opt_cap, uint32_t *d_cursor,
                                   void *d_tup_out, void *d_ov, uint32_t *d_ovcnt,
                                   const uint64_t *carry_col = nullptr);
// caller-owned control memory for the stages of a step that must not allocate (CUDA-graph capture; multi.cu)
struct StageScratch {
    void  *ptr   = nullptr;
    size_t bytes = 0;
};
size_t stage_scratch_bytes(int bits, int nseg);
// the segmented join may start while the peers' broadcast is still in flight (kernels.cuh JoinArgs::wait_flags)
struct JoinWait {
    const uint32_t *flags, *epoch;
    uint32_t        chunk_rows;
    uint32_t       *error;
};
JoinResult stage_join_sum(const void *d_tup_b, const uint32_t *d_hist_b, const void *d_tup_p,
                          const uint32_t *d_hist_p, int bits, int nproj, const ProjDesc *proj, uint32_t opt_cap,
                          const void *d_ov, const uint32_t *d_ovcnt, unsigned long long *d_result = nullptr,
                          int nseg = 0, uint32_t seg_rows = 0, const StageScratch *scr = nullptr,
                          const JoinWait *wait = nullptr, uint32_t seg_head = 0);
// d_off_out (optional): receives the local partition offsets [2^bits + 1]
void       stage_scatter_build_local(const uint64_t *d_keys, uint64_t n, uint32_t rid_base, int bits,
                                     const uint32_t *d_hist_local, void *d_tup_out, int npay,
                                     const uint64_t *const *pay_cols, uint64_t *const *pay_out,
                                     const StageScratch *scr = nullptr, uint32_t *d_off_out = nullptr,
                                     const PredSet *skip = nullptr);
// the same result through two partition passes (coarse into d_tmp, then fine): for 2^11 partitions and more
void       stage_scatter_two_pass(const uint64_t *d_keys, uint64_t n, int bits, const uint32_t *d_hist_local, void *d_tup_out,
                                  const uint64_t *carry_col, void *d_tmp, const StageScratch *scr, uint32_t *d_off_out,
                                  const PredSet *skip = nullptr);
void       stage_build_cursors(const uint32_t *d_hist_all, int world, int rank, int bits, uint32_t *d_total,
                               uint32_t *d_my_start);
void       stage_exchange_cursors(const uint32_t *d_hist_all, int world, int rank, int bits, uint32_t cap,
                                  uint32_t *d_src_off, uint32_t *d_dst_start, uint32_t *d_own_total, uint32_t *d_need);
void       stage_exchange_segments(const void *d_src_tup, int npay, const uint64_t *const *src_pay, uint64_t n,
                                   int bits, int world, const uint32_t *d_src_off, const uint32_t *d_dst_start,
                                   uint32_t cap, int rewrite_rid, void *const *tup_dst, uint64_t *const *pay_dst);

// runtime control (engine.cu)
void               request_device(int device);   // before the first use
int                device_index();
void               set_profiling(bool on);
const std::string &last_error_string();

// small helpers
uint64_t read_counter(const unsigned long long *d_ptr);   // D2H + sync on ctx stream
int      grid_for(uint64_t work_items, int per_block, int max_blocks_per_sm);
void     set_reserved_sms(int n);   // SMs grid_for leaves free on the calling thread (0 = none)

}  // namespace b200
