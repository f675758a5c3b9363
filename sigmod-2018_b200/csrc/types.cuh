// types.cuh — plain data types shared between the kernels (kernels.cuh, only
// included by engine.cu) and the host-side operator code.
#pragma once

#include <cstdint>

namespace b200 {

constexpr int kMaxGather = 8;
constexpr int      kMaxProj  = 8;

// Lazy key vector = the GetRelation contract (inter_res.c:182-231):
// key(i) = col[i] when ids == nullptr, else col[ids[i]]; the row id handed to
// the join is i itself (base row, or position in the intermediate).
struct KeySrc {
    const uint64_t *col;
    const uint32_t *ids;
    uint32_t        n;
};

// Filter predicates folded into the load stage of the partition kernels (query.c:337-399 runs the
// filters of a binding before its first join, filter.c:92-190 scans once per predicate): row r of a base relation
// takes part in the join iff every predicate col[p.col][r] (cmp) p.k holds — cmp 0 '<', 1 '>', 2 '=' on uint64,
// the constant being the reference's 32-bit int converted by the usual C rules (filter.c:118).
constexpr int kMaxPred     = 4;
constexpr int kMaxPredCols = 3;
struct PredSet {
    int             npred = 0, ncols = 0;
    const uint64_t *col[kMaxPredCols] = {nullptr, nullptr, nullptr};
    struct {
        int      col, cmp;
        uint64_t k;
    } p[kMaxPred] = {};
    // multi-GPU exchange plan: rows whose key sits in this direct-mapped table of hot keys (multi_kernels.cuh) were
    // joined during the histogram pass and are skipped; *hot_n == 0 disables the check
    const uint32_t *hot_keys = nullptr, *hot_n = nullptr;
};

struct alignas(8) Tup32 {
    uint32_t key;
    uint32_t rid;
};
struct alignas(16) Tup64 {
    uint64_t key;
    uint32_t rid;
    uint32_t pad;
};
template <typename KeyT> struct TupOf;
template <> struct TupOf<uint32_t> { using type = Tup32; };
template <> struct TupOf<uint64_t> { using type = Tup64; };

// One SUM projection of the fused final join: value = col[ids ? ids[r] : r]
// where r is the build-side (side 0) or probe-side (side 1) row id of a match.
struct ProjDesc {
    const uint64_t *col;
    const uint32_t *ids;
    int             side;
    // build-side projections only: the projected values already permuted into
    // partition order by the scatter (value of build tuple i of the partition
    // buffer), or nullptr when the value is gathered through the row id
    const uint64_t *part_vals;
    // kInRid: the (32-bit) value travels in the row-id slot of the build tuple itself
};
#define B200_PROJ_IN_RID (reinterpret_cast<const uint64_t *>(1))

}  // namespace b200
