// operators.cu — Part 1 of include/b200_join.h: the reference's operator API
// (the 16 symbols query.c:ExecuteQuery imports) implemented on device-resident
// data.  Each function states the reference lines whose contract it keeps
// (SURVEY Appendix B restates those contracts without the reference's linked
// lists).  `relation`, `result` and `inter_data` are opaque to the caller, so
// they carry device buffers behind the reference's public prefix.
#include "../../include/b200_join.h"
#include "engine.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <vector>

using namespace b200;

namespace {

struct B200Relation : relation {
    KeyVec kv;
};

enum ResultKind { kRowIds = 1, kPairs = 2, kDeferredJoin = 3, kDeferredFilter = 4 };
struct B200Result : result {
    int       kind = 0;
    DevBufPtr a, b;   // row ids / (R ids, S ids)
    uint64_t  n = 0;
    KeyVec    kr, ks;   // kDeferredJoin: the two inputs of a join that has not run yet
    bool      fused_r = false, fused_s = false;   // join inputs whose filter predicates ran inside the join's load stage
    bool      null_query = false;                 // ... and one of those filters turned out empty
    // kDeferredFilter: one predicate on a base column that has not been scanned (see PendingFilter)
    const uint64_t *f_col = nullptr;
    int             f_cmp = 0;
    uint64_t        f_k   = 0;
    double          f_sel = 1.0;   // estimated selectivity of this predicate
};

// Lazy last join ("fold the last join into the SUM"; query.c:408-461 runs the joins one after
// the other and only then CalculateQueryResults).  Unless B200_LAZY_JOIN=0, RadixHashJoin returns a DEFERRED
// result and InsertJoinToInterResults parks it on the node it belongs to.  Whatever operator looks at the
// intermediate next materialises it first (resolve_all: the pairs are produced and inserted exactly as the
// eager path does), so every join but the last behaves as before; when the next operator is
// CalculateQueryResults the join runs fused with the SUMs (run_join(..., JoinOut::Sum)) and its pairs are never
// written.  The caller cannot tell the difference: it only reads inter_res::next (query.c:453, 462), and
// AreActiveInInter sees non-NULL placeholders for the parked bindings.
struct PendingJoin {
    bool   active = false;
    KeyVec kr, ks;
    int    rel1 = -1, rel2 = -1;
    int    active_side = -1;   // -1: both bindings are new to the node; 0 / 1: rel1 / rel2 was already in it
};

// Filter fusion (query.c:337-399 runs every filter of a binding before its first join, each one a full
// scan, filter.c:92-190).  A filter on a LARGE base relation that is in no intermediate node yet is not scanned:
// Filter returns a deferred result, InsertSingleRowIdsToInterResult parks its predicate on a node that holds nothing
// but that binding (which reads as active from then on), further filters on the binding join the parked set, and
// GetRelation hands the join a key vector that carries the predicates: the partition kernels evaluate them in
// their load stage — no row-id list, no host round trip, no compaction gather — and the join reports base row
// ids for the binding, exactly what the eager path's row-id list would have translated its positions into.
// Any other operator that looks at the node first scans the parked predicates the eager way (resolve_filter).  An
// empty filter still makes the whole query NULL (query.c:360-369): the kernels count the rows that passed.
struct PendingFilter {
    bool     active  = false;
    int      binding = -1;
    uint64_t rows    = 0;      // rows of the base relation
    double   est_sel = 1.0;    // estimated fraction of rows that pass (stats.c's formulas on the column statistics)
    PredSet  preds;
};

struct B200InterData : inter_data {
    std::vector<DevBufPtr> bufs;
    PendingJoin            pending;
    PendingFilter          pfilter;
    bool                   null_query = false;   // a fused filter was empty: the query's answer is NULL
};

uint64_t *const kParked  = reinterpret_cast<uint64_t *>(8);    // table[] entry of a binding whose join is parked
uint64_t *const kParkedF = reinterpret_cast<uint64_t *>(16);   // ... whose filter predicates are parked
constexpr uint64_t kFuseMinRows = 1ull << 18;                  // smaller relations are filtered the eager way

thread_local bool t_null_result = false;   // see b200_last_result_null
// 0 = off, 1 = where it pays (default), 2 = wherever possible (tests); -1: B200_FUSE_FILTERS from the environment
std::atomic<int> g_fuse_filters{-1};
int fuse_filters_mode() {
    int v = g_fuse_filters.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("B200_FUSE_FILTERS");
        v             = e ? std::max(0, std::min(atoi(e), 2)) : 1;
        g_fuse_filters.store(v, std::memory_order_relaxed);
    }
    return v;
}
bool fuse_filters_enabled() { return fuse_filters_mode() != 0; }

// stats.c:88-259's selectivity estimates from min / max / distinct (uniform values): the fraction of rows that pass
double estimate_selectivity(const DevColumn &c, int cmp, uint64_t k) {
    const double lo = (double)c.min_val, hi = (double)c.max_val, kk = (double)k;
    if (c.max_val == UINT64_MAX || hi < lo) return 0.5;
    const double range = hi - lo + 1.0;
    double       sel;
    if (cmp == 2) sel = (kk < lo || kk > hi) ? 0.0 : 1.0 / (c.distinct ? (double)c.distinct : range);
    else if (cmp == 0) sel = (kk - lo) / range;          // v < k
    else sel = (hi - kk) / range;                        // v > k
    return std::min(1.0, std::max(0.0, sel));
}

// Fusing pays when enough rows survive: the fused kernels stream the key column and every predicate column over ALL
// rows of the relation and keep it out of the small-build-side plans, while the eager scans cost one pass per
// predicate but leave the join a relation of the filtered size.
bool worth_fusing(const PendingFilter &pf) {
    if (fuse_filters_mode() == 2) return true;
    const double est = pf.est_sel * (double)pf.rows;
    return est >= (double)kFuseMinRows && est * 64.0 >= (double)pf.rows;
}

std::atomic<int> g_lazy_join{-1};   // -1: B200_LAZY_JOIN from the environment (default on)
bool lazy_join_enabled() {
    int v = g_lazy_join.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("B200_LAZY_JOIN");
        v             = e ? (atoi(e) != 0) : 1;
        g_lazy_join.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

B200InterData *idata(const inter_res *node) { return static_cast<B200InterData *>(node->data); }

B200InterData *new_inter_data(int num_rel, uint64_t num_tuples) {
    auto *d       = new B200InterData();
    d->num_tuples = num_tuples;
    d->table      = static_cast<uint64_t **>(calloc((size_t)num_rel, sizeof(uint64_t *)));
    d->bufs.resize((size_t)num_rel);
    return d;
}
void free_inter_data(B200InterData *d) {
    free(d->table);
    delete d;
}
void set_column(B200InterData *d, int b, DevBufPtr buf) {
    d->table[b] = buf ? reinterpret_cast<uint64_t *>(buf->ptr) : nullptr;
    d->bufs[(size_t)b] = std::move(buf);
}
const uint32_t *column_ids(const B200InterData *d, int b) {
    return reinterpret_cast<const uint32_t *>(d->table[b]);
}

inter_res *new_node(int num_rel) {
    auto *n             = static_cast<inter_res *>(malloc(sizeof(inter_res)));
    n->next             = nullptr;
    n->num_of_relations = num_rel;
    n->data             = new_inter_data(num_rel, 0);
    return n;
}

B200Result *new_result(int kind, uint64_t n, DevBufPtr a, DevBufPtr b) {
    auto *r         = new B200Result();
    r->buff         = nullptr;
    r->next         = nullptr;
    r->current_load = n;
    r->kind         = kind;
    r->n            = n;
    r->a            = std::move(a);
    r->b            = std::move(b);
    return r;
}

// node that holds binding b, or NULL (filter.c:98-104, inter_res.c:184-189)
inter_res *find_node(inter_res *head, int b) {
    for (; head; head = head->next)
        if (head->data->table[b] != nullptr) return head;
    return nullptr;
}

// compact every active column of `node` through a list of positions
// (filter.c:42-81)
void compact_node(inter_res *node, const uint32_t *pos, uint64_t m) {
    B200InterData                 *old = idata(node);
    std::vector<const uint32_t *>  in;
    std::vector<int>               which;
    for (int j = 0; j < node->num_of_relations; ++j)
        if (old->table[j]) {
            in.push_back(column_ids(old, j));
            which.push_back(j);
        }
    std::vector<DevBufPtr> out = run_gather(pos, m, in);
    B200InterData         *nd  = new_inter_data(node->num_of_relations, m);
    for (size_t k = 0; k < which.size(); ++k) set_column(nd, which[k], out[k]);
    free_inter_data(old);
    node->data = nd;
}

DevColumn binding_column(relation_map *map, int *query_relations, int binding, int column) {
    relation_map &rm = map[query_relations[binding]];
    return lookup_column(rm.columns[column], rm.num_tuples);
}

// pairs (position in R's key vector, position in S's) into the node: inter_res.c:39-62 when both bindings are
// new to it, 64-141 when one of them is active (every active column is gathered through its positions)
void insert_pairs(inter_res *node, int rel1, int rel2, int active_side, uint64_t m, DevBufPtr r_ids, DevBufPtr s_ids) {
    if (active_side < 0) {
        B200InterData *d = idata(node);
        d->num_tuples    = m;
        set_column(d, rel1, std::move(r_ids));
        set_column(d, rel2, std::move(s_ids));
        return;
    }
    const bool      has1    = active_side == 0;
    DevBufPtr       pos     = has1 ? r_ids : s_ids;
    DevBufPtr       fresh   = has1 ? s_ids : r_ids;
    const int       new_rel = has1 ? rel2 : rel1;
    compact_node(node, pos->as<uint32_t>(), m);
    set_column(idata(node), new_rel, fresh);
}

// run the parked join of `node`, if any, and insert its pairs
void resolve(inter_res *node) {
    B200InterData *d = idata(node);
    if (!d->pending.active) return;
    PendingJoin pj    = std::move(d->pending);
    d->pending        = PendingJoin{};
    // the placeholders go first: compact_node gathers every non-NULL column
    if (pj.active_side != 0) d->table[pj.rel1] = nullptr;
    if (pj.active_side != 1) d->table[pj.rel2] = nullptr;
    if (pj.active_side < 0) d->num_tuples = 0;
    JoinResult j = run_join(pj.kr, pj.ks, JoinOut::Pairs, 0, nullptr);
    if (j.valid_r == 0 || j.valid_s == 0) d->null_query = true;   // a fused filter was empty
    insert_pairs(node, pj.rel1, pj.rel2, pj.active_side, j.m, j.r_ids, j.s_ids);
}
void resolve_all(inter_res *head) {
    for (; head; head = head->next) resolve(head);
}

// scan the parked predicates of `node` the eager way (filter.c:92-190 once per predicate, filter.c:11-89 in between)
void resolve_filter(inter_res *node) {
    B200InterData *d = idata(node);
    if (!d->pfilter.active) return;
    PendingFilter pf = d->pfilter;
    d->pfilter       = PendingFilter{};
    const int b      = pf.binding;
    d->table[b]      = nullptr;
    d->num_tuples    = 0;
    DevBufPtr ids;
    uint64_t  n = pf.rows;
    for (int i = 0; i < pf.preds.npred && n > 0; ++i) {
        KeySrc src{pf.preds.col[pf.preds.p[i].col], ids ? ids->as<uint32_t>() : nullptr, (uint32_t)n};
        IdList l = run_filter_u64(src, pf.preds.p[i].cmp, pf.preds.p[i].k);
        if (ids && l.n) {   // positions in the previous list -> base row ids
            std::vector<DevBufPtr> out = run_gather(l.ids->as<uint32_t>(), l.n, {ids->as<uint32_t>()});
            ids                        = out[0];
        } else {
            ids = l.ids;
        }
        n = l.n;
    }
    if (n == 0) {
        d->null_query = true;
        ids           = dev_alloc(16);
    }
    d->num_tuples = n;
    set_column(d, b, ids);
}
void resolve_filters(inter_res *head) {
    for (; head; head = head->next) resolve_filter(head);
}
// the parked filter of `binding` was consumed by a join: its node is an empty node again
void clear_parked_filter(inter_res *head, int binding) {
    for (; head; head = head->next) {
        B200InterData *d = idata(head);
        if (d->pfilter.active && d->pfilter.binding == binding) {
            d->pfilter        = PendingFilter{};
            d->table[binding] = nullptr;
            d->num_tuples     = 0;
        }
    }
}
bool any_null_query(inter_res *head) {
    for (; head; head = head->next)
        if (idata(head)->null_query) return true;
    return false;
}

// a deferred join result that is looked at directly (tests) turns into its pairs
void materialise(B200Result *r) {
    if (r && r->kind == kDeferredFilter) {   // looked at directly (tests): scan now
        IdList l        = run_filter_u64(KeySrc{r->f_col, nullptr, (uint32_t)r->n}, r->f_cmp, r->f_k);
        r->kind         = kRowIds;
        r->a            = l.ids;
        r->n            = l.n;
        r->current_load = l.n;
        return;
    }
    if (!r || r->kind != kDeferredJoin) return;
    JoinResult j    = run_join(r->kr, r->ks, JoinOut::Pairs, 0, nullptr);
    if (j.valid_r == 0 || j.valid_s == 0) r->null_query = true;
    r->kind         = kPairs;
    r->n            = j.m;
    r->current_load = j.m;
    r->a            = j.r_ids;
    r->b            = j.s_ids;
    r->kr = r->ks = KeyVec{};
}

}  // namespace

extern "C" {

// inter_res.c:26-32
int InitInterResults(inter_res **head, int num_of_rel) {
    ensure_init();
    *head = new_node(num_of_rel);
    return 0;
}

// inter_res.c:175-180
void FreeInterResults(inter_res *var) {
    while (var) {
        inter_res *next = var->next;
        free_inter_data(idata(var));
        free(var);
        var = next;
    }
}

// filter.c:92-190
result *Filter(inter_res *head, filter_pred *filter_p, relation_map *map, int *query_relations) {
    resolve_all(head);
    const int  rel  = filter_p->relation;
    DevColumn  col  = binding_column(map, query_relations, rel, filter_p->column);
    inter_res *node = find_node(head, rel);
    int        code;
    switch (filter_p->comperator) {
        case '<': code = 0; break;
        case '>': code = 1; break;
        case '=': code = 2; break;
        default:
            fprintf(stderr, "Wrong comperator in filter function\n");   // filter.c:184-186
            exit(2);
    }
    if (fuse_filters_enabled()) {
        bool defer = false;
        if (node && idata(node)->pfilter.active) {
            // another filter on a binding whose predicates are parked: it joins them while there is room
            const PredSet &ps = idata(node)->pfilter.preds;
            bool known = false;
            for (int c = 0; c < ps.ncols; ++c) known = known || ps.col[c] == col.d;
            defer = ps.npred < kMaxPred && (known || ps.ncols < kMaxPredCols);
            if (!defer) resolve_filter(node);
        } else if (!node && col.n >= kFuseMinRows) {
            defer = true;
        }
        if (defer) {
            B200Result *r = new_result(kDeferredFilter, 0, nullptr, nullptr);
            r->f_col      = col.d;
            r->f_cmp      = code;
            r->f_k        = (uint64_t)(int64_t)filter_p->value;   // `uint64_t (cmp) int`: sign-extended (filter.c:118)
            r->f_sel      = estimate_selectivity(col, code, r->f_k);
            r->n          = col.n;
            r->kr.max_val = col.max_val;
            return r;
        }
    }
    KeySrc     src;
    src.col = col.d;
    if (node) {   // filter.c:124-133: scan through the row ids, emit positions
        src.ids = column_ids(idata(node), rel);
        src.n   = (uint32_t)node->data->num_tuples;
    } else {      // filter.c:113-122: scan the base column, emit row ids
        src.ids = nullptr;
        src.n   = (uint32_t)col.n;
    }
    IdList l = run_filter(src, filter_p->comperator, filter_p->value);
    if (l.n == 0) return nullptr;   // caller prints NULLs (query.c:360-369)
    return new_result(kRowIds, l.n, l.ids, nullptr);
}

// filter.c:11-89.  Deviation (documented in DESIGN.md): the reference advances
// the caller's head pointer while searching (filter.c:82) and dereferences
// NULL when the binding is in no node and the first node is occupied
// (SURVEY §8 quirk 2); here the list head is left alone and a new node is
// appended in that case.
int InsertSingleRowIdsToInterResult(inter_res **head, int relation_num, result *res) {
    B200Result *r = static_cast<B200Result *>(res);
    B200_REQUIRE(r && (r->kind == kRowIds || r->kind == kDeferredFilter),
                 "InsertSingleRowIdsToInterResult needs a row-id result");
    resolve_all(*head);
    if (r->kind == kDeferredFilter) {
        // park the predicate: on the node that already holds this binding's parked set, else on the first empty
        // node (filter.c:19-40 installs a fresh binding there), else on a new node
        inter_res *node = *head, *last = nullptr, *target = nullptr;
        for (; node; last = node, node = node->next) {
            B200InterData *d = idata(node);
            if (d->pfilter.active && d->pfilter.binding == relation_num) { target = node; break; }
        }
        if (!target)
            for (node = *head; node; node = node->next)
                if (node->data->num_tuples == 0 && !idata(node)->pending.active) { target = node; break; }
        if (!target) {
            last->next = new_node(last->num_of_relations);
            target     = last->next;
        }
        B200InterData *d  = idata(target);
        PendingFilter &pf = d->pfilter;
        if (!pf.active) {
            pf                     = PendingFilter{};
            pf.active              = true;
            pf.binding             = relation_num;
            pf.rows                = r->n;
            d->table[relation_num] = kParkedF;
            d->num_tuples          = 1;   // "not an empty node"; the real count exists once the predicates ran
        }
        int c = 0;
        while (c < pf.preds.ncols && pf.preds.col[c] != r->f_col) ++c;
        B200_REQUIRE(c < kMaxPredCols && pf.preds.npred < kMaxPred, "parked filter set is full");
        if (c == pf.preds.ncols) pf.preds.col[pf.preds.ncols++] = r->f_col;
        pf.preds.p[pf.preds.npred].col = c;
        pf.preds.p[pf.preds.npred].cmp = r->f_cmp;
        pf.preds.p[pf.preds.npred].k   = r->f_k;
        ++pf.preds.npred;
        pf.est_sel *= r->f_sel;
        return 1;
    }
    inter_res *node = *head, *last = nullptr;
    for (; node; last = node, node = node->next) {
        if (node->data->num_tuples == 0) {   // filter.c:19-40: first instance of the node
            B200InterData *d = idata(node);
            d->num_tuples    = r->n;
            set_column(d, relation_num, r->a);
            return 1;
        }
        if (node->data->table[relation_num] != nullptr) {   // filter.c:42-81
            compact_node(node, r->a->as<uint32_t>(), r->n);
            return 1;
        }
    }
    last->next = new_node(last->num_of_relations);
    return InsertSingleRowIdsToInterResult(&last->next, relation_num, res);
}

// inter_res.c:208-231 (+182-206).  K2 is fused away: the key vector stays a
// (column, row-id list) view and is read by the partition / join kernels.
relation *GetRelation(int given_rel, int column, inter_res *inter, relation_map *map, int *query_relations) {
    DevColumn  col  = binding_column(map, query_relations, given_rel, column);
    resolve_all(inter);
    inter_res *node = inter ? find_node(inter, given_rel) : nullptr;
    auto      *rel  = new B200Relation();
    rel->tuples     = nullptr;
    rel->kv.src.col = col.d;
    rel->kv.max_val = col.max_val;
    if (node && idata(node)->pfilter.active && !worth_fusing(idata(node)->pfilter)) resolve_filter(node);
    if (node && idata(node)->pfilter.active) {
        // the binding's filters are parked: the join evaluates them in its load stage and reports base row ids
        rel->kv.src.ids = nullptr;
        rel->kv.src.n   = (uint32_t)col.n;
        rel->kv.preds   = idata(node)->pfilter.preds;
        rel->num_tuples = col.n;
        return rel;
    }
    if (node) {
        B200InterData *d   = idata(node);
        rel->kv.src.ids    = column_ids(d, given_rel);
        rel->kv.ids_owner  = d->bufs[(size_t)given_rel];
        rel->kv.src.n      = (uint32_t)d->num_tuples;
        rel->num_tuples    = d->num_tuples;
    } else {
        rel->kv.src.ids = nullptr;
        rel->kv.src.n   = (uint32_t)col.n;
        rel->num_tuples = col.n;
    }
    return rel;
}

// rhjoin.c:13-111
result *RadixHashJoin(relation *relR, relation *relS, scheduler *sched) {
    (void)sched;   // kernels run on the calling thread's stream
    if (relR->num_tuples == 0 || relS->num_tuples == 0) return nullptr;   // rhjoin.c:15-16
    B200Relation *r = static_cast<B200Relation *>(relR);
    B200Relation *s = static_cast<B200Relation *>(relS);
    if (lazy_join_enabled()) {
        // deferred: the key vectors keep their row-id lists alive (ids_owner) after the caller frees relR / relS
        B200Result *d = new_result(kDeferredJoin, 0, nullptr, nullptr);
        d->kr         = r->kv;
        d->ks         = s->kv;
        d->fused_r    = r->kv.preds.npred > 0;
        d->fused_s    = s->kv.preds.npred > 0;
        return d;
    }
    JoinResult    j = run_join(r->kv, s->kv, JoinOut::Pairs, 0, nullptr);
    // a fused filter that let nothing through is the reference's empty Filter: the whole query is NULL (query.c:360-369)
    if (j.valid_r == 0 || j.valid_s == 0) return nullptr;
    // an empty join is a non-NULL result with zero pairs (rhjoin.c:356-359)
    B200Result *res = new_result(kPairs, j.m, j.r_ids, j.s_ids);
    res->fused_r    = r->kv.preds.npred > 0;
    res->fused_s    = s->kv.preds.npred > 0;
    return res;
}

// inter_res.c:34-152
int InsertJoinToInterResults(inter_res *head, int rel1, int rel2, result *res) {
    B200Result *r = static_cast<B200Result *>(res);
    B200_REQUIRE(r && (r->kind == kPairs || r->kind == kDeferredJoin), "InsertJoinToInterResults needs a join result");
    resolve_all(head);
    // a side whose parked filter ran (or will run) inside the join comes back as base row ids: its node is empty again
    if (r->fused_r) clear_parked_filter(head, rel1);
    if (r->fused_s) clear_parked_filter(head, rel2);
    if (r->null_query) idata(head)->null_query = true;
    inter_res *node = head, *last = nullptr;
    for (; node; last = node, node = node->next) {
        B200InterData *d    = idata(node);
        const bool     none = d->num_tuples == 0;   // inter_res.c:39-62
        const bool     has1 = d->table[rel1] != nullptr, has2 = d->table[rel2] != nullptr;
        if (!none && has1 == has2) continue;
        // inter_res.c:64-102 (rel1 active) / 104-141 (rel2 active)
        const int active_side = none ? -1 : has1 ? 0 : 1;
        if (r->kind == kDeferredJoin) {
            // park the join on this node; its bindings read as active from now on
            d->pending.active      = true;
            d->pending.kr          = r->kr;
            d->pending.ks          = r->ks;
            d->pending.rel1        = rel1;
            d->pending.rel2        = rel2;
            d->pending.active_side = active_side;
            if (active_side != 0) d->table[rel1] = kParked;
            if (active_side != 1) d->table[rel2] = kParked;
            if (none) d->num_tuples = 1;   // "not an empty node"; the real count exists once the join has run
            return 1;
        }
        insert_pairs(node, rel1, rel2, active_side, r->n, r->a, r->b);
        return 1;
    }
    // inter_res.c:147-151: neither side lives in a node yet
    last->next = new_node(last->num_of_relations);
    InsertJoinToInterResults(last->next, rel1, rel2, res);
    return 0;
}

// inter_res.c:352-361
int AreActiveInInter(inter_res *inter, int rel1, int rel2) {
    for (; inter; inter = inter->next)
        if (inter->data->table[rel1] != nullptr && inter->data->table[rel2] != nullptr) return 1;
    return 0;
}

// inter_res.c:363-389
int JoinInterNode(inter_res **inter, relation_map *rel_map, int rel1, int col1, int rel2, int col2,
                  int *relations) {
    resolve_all(*inter);
    resolve_filters(*inter);
    inter_res *node = *inter;
    for (; node; node = node->next)
        if (node->data->table[rel1] != nullptr && node->data->table[rel2] != nullptr) break;
    if (!node) return 0;
    DevColumn      ca = binding_column(rel_map, relations, rel1, col1);
    DevColumn      cb = binding_column(rel_map, relations, rel2, col2);
    B200InterData *d  = idata(node);
    IdList l = run_inter_equal(ca.d, column_ids(d, rel1), cb.d, column_ids(d, rel2), d->num_tuples);
    // the reference hands an empty list to InsertSingleRowIds... which then
    // sees num_results == 0; the node simply becomes empty
    compact_node(node, l.ids->as<uint32_t>(), l.n);
    return 1;
}

// inter_res.c:265-318: a later node that shares an active binding with an
// earlier one is folded into it through that binding's positions.
void MergeInterNodes(inter_res **inter) {
    resolve_all(*inter);   // (a node that only parks a filter shares no materialised binding with any other: it is skipped)
    for (inter_res *head = *inter; head; head = head->next) {
        bool merged = true;
        while (merged) {
            merged = false;
            for (inter_res *prev = head; prev->next && !merged; prev = prev->next) {
                inter_res *cand = prev->next;
                for (int i = 0; i < head->num_of_relations && !merged; ++i) {
                    if (head->data->table[i] == nullptr || cand->data->table[i] == nullptr) continue;
                    // inter_res.c:287-318: head.table[j][k] = cand.table[j][head.table[i][k]]
                    B200InterData                *hd = idata(head);
                    B200InterData                *cd = idata(cand);
                    std::vector<const uint32_t *> in;
                    std::vector<int>              which;
                    for (int j = 0; j < head->num_of_relations; ++j)
                        if (cd->table[j]) {
                            in.push_back(column_ids(cd, j));
                            which.push_back(j);
                        }
                    DevBufPtr              pos_owner = hd->bufs[(size_t)i];
                    std::vector<DevBufPtr> out = run_gather(pos_owner->as<uint32_t>(), hd->num_tuples, in);
                    for (size_t k = 0; k < which.size(); ++k) set_column(hd, which[k], out[k]);
                    prev->next = cand->next;
                    free_inter_data(cd);
                    free(cand);
                    merged = true;
                }
            }
        }
    }
}

// inter_res.c:391-428: cross product of the remaining nodes (tail first).
void CartesianInterResults(inter_res **inter) {
    inter_res *cur = *inter;
    if (cur->next == nullptr) return;
    resolve_all(cur);
    resolve_filters(cur);
    CartesianInterResults(&cur->next);
    inter_res     *nxt = cur->next;
    B200InterData *a = idata(cur), *b = idata(nxt);
    if (a->num_tuples * b->num_tuples == 0) return;   // inter_res.c:398
    const uint64_t total = a->num_tuples * b->num_tuples;
    B200_REQUIRE(total <= kMaxRows, "cartesian product exceeds 2^32-1 rows");
    B200InterData *nd = new_inter_data(cur->num_of_relations, total);
    for (int z = 0; z < cur->num_of_relations; ++z) {
        const bool in_a = a->table[z] != nullptr, in_b = b->table[z] != nullptr;
        if (!in_a && !in_b) continue;
        DevBufPtr out = dev_alloc(total * sizeof(uint32_t));
        run_cartesian(in_a ? column_ids(a, z) : column_ids(b, z), a->num_tuples, b->num_tuples, in_a,
                      out->as<uint32_t>());
        set_column(nd, z, out);
    }
    free_inter_data(a);
    cur->data = nd;
    free_inter_data(b);
    free(nxt);
    cur->next = nullptr;
}

// inter_res.c:320-339 without the printf
int b200_calculate_sums(inter_res *inter, relation_map *map, batch_listnode *query, uint64_t *sums,
                        uint64_t *num_rows) {
    const int                      nv = query->views->num_of_elements;
    std::vector<const uint64_t *>  cols((size_t)nv);
    std::vector<const uint32_t *>  ids((size_t)nv);
    resolve_filters(inter);   // a filter still parked here was never consumed by a join: scan it now
    B200InterData                 *d = idata(inter);
    t_null_result                    = false;
    if (d->pending.active) {
        // the last join is still parked on the node: run it fused with the SUMs.  A projection on a binding that
        // was in the node before the join reads through that binding's row ids at the position the join reports
        // for the node's side; one on a binding the join brings in reads the base row of the other side.
        const PendingJoin &pj = d->pending;
        ProjDesc           pd[kMaxProj];
        bool               fusable = nv <= kMaxProj;
        for (int i = 0; i < nv && fusable; ++i) {
            const int index  = query->views->data[i][0] - '0';
            const int column = query->views->data[i][2] - '0';
            if (index < 0 || index >= inter->num_of_relations || d->table[index] == nullptr) {
                fusable = false;
                break;
            }
            DevColumn col = binding_column(map, query->relations, index, column);
            int       side;
            const uint32_t *through = nullptr;
            if (pj.active_side < 0) {
                if (index != pj.rel1 && index != pj.rel2) { fusable = false; break; }
                side = index == pj.rel1 ? 0 : 1;
                if (pj.rel1 == pj.rel2) fusable = false;
            } else {
                const int new_rel = pj.active_side == 0 ? pj.rel2 : pj.rel1;
                if (index == new_rel) {
                    side = 1 - pj.active_side;
                } else {
                    side    = pj.active_side;
                    through = column_ids(d, index);
                }
            }
            pd[i] = ProjDesc{col.d, through, side, nullptr};
        }
        if (fusable) {
            JoinResult j = run_join(pj.kr, pj.ks, JoinOut::Sum, nv, pd);
            for (int i = 0; i < nv; ++i) sums[i] = j.sums[i];
            if (num_rows) *num_rows = j.m;
            t_null_result = any_null_query(inter) || j.valid_r == 0 || j.valid_s == 0;
            return 0;
        }
        resolve(inter);
        d = idata(inter);
    }
    for (int i = 0; i < nv; ++i) {
        // single-digit binding and column, as in inter_res.c:325-327
        const int index  = query->views->data[i][0] - '0';
        const int column = query->views->data[i][2] - '0';
        DevColumn col    = binding_column(map, query->relations, index, column);
        cols[(size_t)i]  = col.d;
        ids[(size_t)i]   = column_ids(d, index);
        B200_REQUIRE(ids[(size_t)i] != nullptr || d->num_tuples == 0,
                     "projection on a binding that is not in the intermediate result");
    }
    run_checksum(d->num_tuples, nv, cols.data(), ids.data(), sums);
    if (num_rows) *num_rows = d->num_tuples;
    t_null_result = any_null_query(inter);
    return 0;
}

// 1 when the calling thread's last b200_calculate_sums found that a filter fused into a join had let nothing
// through: the reference prints NULL for every projection then (its Filter returns NULL, query.c:360-369)
int b200_last_result_null(void) { return t_null_result ? 1 : 0; }

int b200_set_fuse_filters(int mode) {
    const int before = fuse_filters_mode();
    g_fuse_filters.store(std::max(0, std::min(mode, 2)), std::memory_order_relaxed);
    return before;
}

// inter_res.c:320-339
void CalculateQueryResults(inter_res *inter, relation_map *map, batch_listnode *query) {
    const int             nv = query->views->num_of_elements;
    std::vector<uint64_t> sums((size_t)nv);
    b200_calculate_sums(inter, map, query, sums.data(), nullptr);
    if (t_null_result) {
        PrintNullResults(query);
        return;
    }
    for (int i = 0; i < nv; ++i) {
        printf("%lu", (unsigned long)sums[(size_t)i]);
        if (i != nv - 1) printf(" ");
    }
    printf("\n");
}

// inter_res.c:341-350
void PrintNullResults(batch_listnode *query) {
    for (int i = 0; i < query->views->num_of_elements; ++i) {
        printf("NULL");
        if (i != query->views->num_of_elements - 1) printf(" ");
    }
    printf("\n");
}

// inter_res.c:234-263 (contract, not the reference's buggy indexing)
result *SelfJoin(int given_rel, int column1, int column2, inter_res **inter, relation_map *map,
                 int *query_relations) {
    DevColumn  c1   = binding_column(map, query_relations, given_rel, column1);
    DevColumn  c2   = binding_column(map, query_relations, given_rel, column2);
    resolve_all(*inter);
    resolve_filters(*inter);
    inter_res *node = find_node(*inter, given_rel);
    IdList     l;
    if (node) {
        const uint32_t *t = column_ids(idata(node), given_rel);
        l                 = run_inter_equal(c1.d, t, c2.d, t, node->data->num_tuples);
    } else {
        l = run_inter_equal(c1.d, nullptr, c2.d, nullptr, c1.n);
    }
    if (l.n == 0) return nullptr;
    return new_result(kRowIds, l.n, l.ids, nullptr);
}

// results.c:144-153
void FreeResult(result *head) {
    if (head) delete static_cast<B200Result *>(head);
}

// preprocess.c:213-218
void FreeRelation(relation *rel) {
    if (rel) delete static_cast<B200Relation *>(rel);
}

// ---- read-back helpers for tests (Part 2) ---------------------------------
int b200_set_lazy_join(int on) {
    const int before = lazy_join_enabled() ? 1 : 0;
    g_lazy_join.store(on ? 1 : 0, std::memory_order_relaxed);
    return before;
}

int b200_result_kind(const result *res) {
    materialise(static_cast<B200Result *>(const_cast<result *>(res)));   // a deferred join shows up as its pairs
    return res ? static_cast<const B200Result *>(res)->kind : 0;
}

static int ids_to_host(const uint32_t *d, uint64_t n, uint64_t *out) {
    if (n == 0) return 0;
    Context  &c   = ctx();
    DevBufPtr tmp = dev_alloc(n * sizeof(uint64_t));
    widen_ids(d, n, tmp->as<uint64_t>());
    B200_CUDA(cudaMemcpyAsync(out, tmp->ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
    B200_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200_result_rowids_to_host(const result *res, uint64_t *out) {
    materialise(static_cast<B200Result *>(const_cast<result *>(res)));   // a deferred filter is scanned now
    const B200Result *r = static_cast<const B200Result *>(res);
    if (!r || r->kind != kRowIds) return 1;
    return ids_to_host(r->a->as<uint32_t>(), r->n, out);
}

int b200_result_pairs_to_host(const result *res, uint64_t *out_r, uint64_t *out_s) {
    materialise(static_cast<B200Result *>(const_cast<result *>(res)));
    const B200Result *r = static_cast<const B200Result *>(res);
    if (!r || r->kind != kPairs) return 1;
    ids_to_host(r->a->as<uint32_t>(), r->n, out_r);
    return ids_to_host(r->b->as<uint32_t>(), r->n, out_s);
}

int b200_inter_column_to_host(const inter_res *node, int binding, uint64_t *out) {
    resolve(const_cast<inter_res *>(node));
    resolve_filter(const_cast<inter_res *>(node));
    const B200InterData *d = idata(node);
    if (!d->table[binding]) return 1;
    return ids_to_host(column_ids(d, binding), d->num_tuples, out);
}

}  // extern "C"
