"""ctypes binding of libb200join.so (include/b200_join.h, include/b200_abi.h).

Part 1 binds the reference's operator API with the reference's own struct
layouts (reference structs.h, cited in include/b200_abi.h); `execute_query`
drives those operators in the order the reference's only caller does
(reference query.c:325-467).  Part 2 binds the shim's own entry points
(lifecycle, registration, kernel-level calls, the fused join->SUM).

No torch types appear here; buffers are numpy arrays (host) or raw device
addresses (ints) for the `location=1` entry points.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

import numpy as np

__all__ = [
    "LIB_PATH", "load_library", "lib", "declared_symbols",
    "RelationMapArray", "DeviceRelationMap", "parse_query", "execute_query", "execute_batch", "QueryResult",
    "scan_filter", "radix_partition", "hash_join_pairs", "gather_sum", "join_sum",
    "join_sum_device", "synth_column_device", "DeviceColumn", "kernel_launches", "last_kernel_ms",
    "SYNTH_PERM", "SYNTH_PAYLOAD", "SYNTH_ZIPF", "SYNTH_UNIFORM", "SYNTH_IOTA",
    "SEED_R", "SEED_S", "CMultiConfig", "PLAN_BROADCAST", "PLAN_EXCHANGE",
]

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libb200join.so"
INCLUDE_DIR = _HERE.parent / "include"

SYNTH_PERM, SYNTH_PAYLOAD, SYNTH_ZIPF, SYNTH_UNIFORM, SYNTH_IOTA = range(5)
SEED_R = 0x51670D180001
SEED_S = 0x51670D180002

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


# --------------------------------------------------------------------------
# ABI structs (include/b200_abi.h <- reference structs.h)
# --------------------------------------------------------------------------
class CResult(C.Structure):
    pass


CResult._fields_ = [("buff", C.c_char_p), ("next", C.POINTER(CResult)), ("current_load", C.c_uint64)]


class CRelation(C.Structure):
    _fields_ = [("tuples", C.c_void_p), ("num_tuples", C.c_uint64)]


class CInterData(C.Structure):
    _fields_ = [("num_tuples", C.c_uint64), ("table", C.POINTER(C.c_void_p))]


class CInterRes(C.Structure):
    pass


CInterRes._fields_ = [("data", C.POINTER(CInterData)), ("num_of_relations", C.c_int),
                      ("next", C.POINTER(CInterRes))]


class CColumnStats(C.Structure):
    _fields_ = [("l", C.c_uint64), ("u", C.c_uint64), ("f", C.c_double), ("d", C.c_double)]


class CRelationMap(C.Structure):
    _fields_ = [("num_tuples", C.c_uint64), ("num_columns", C.c_uint64),
                ("columns", C.POINTER(u64p)), ("col_stats", C.POINTER(CColumnStats))]


class CFilterPred(C.Structure):
    _fields_ = [("relation", C.c_int), ("column", C.c_int), ("value", C.c_int), ("comperator", C.c_char)]


class CQueryStringArray(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_char_p)), ("num_of_elements", C.c_int)]


class CBatchListnode(C.Structure):
    pass


CBatchListnode._fields_ = [("num_of_relations", C.c_int), ("relations", C.POINTER(C.c_int)),
                           ("predicate_list", C.c_void_p), ("views", C.POINTER(CQueryStringArray)),
                           ("next", C.POINTER(CBatchListnode))]


class CMultiConfig(C.Structure):
    """b200_multi_config (include/b200_join.h)"""
    _fields_ = [("plan", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("device", C.c_int),
                ("n_build_total", C.c_uint64), ("n_probe_total", C.c_uint64),
                ("n_build_local", C.c_uint64), ("n_probe_local", C.c_uint64),
                ("n_build_local_max", C.c_uint64), ("n_probe_local_max", C.c_uint64),
                ("has_build_sum", C.c_int), ("has_probe_sum", C.c_int), ("radix_bits", C.c_int), ("chunks", C.c_int),
                ("recv_rows_build", C.c_uint64), ("recv_rows_probe", C.c_uint64), ("hot_keys", C.c_int)]


PLAN_BROADCAST, PLAN_EXCHANGE = 0, 1


# --------------------------------------------------------------------------
# library loading
# --------------------------------------------------------------------------
_lib = None


def declared_symbols() -> list[str]:
    """Every function include/b200_join.h declares."""
    text = (INCLUDE_DIR / "b200_join.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)))


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libb200join.so; there is no fallback when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # B200_LIB selects another build of the same library (the checked build, lib/libb200join_checked.so)
    p = Path(path) if path else Path(os.environ.get("B200_LIB") or LIB_PATH)
    if not p.exists():
        raise ImportError(
            f"{p} is missing: build it with `make -C sigmod-2018_b200/csrc` "
            "(or __graft_entry__.build()); this package has no CPU fallback")
    lib = C.CDLL(str(p))
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def lib() -> C.CDLL:
    return load_library()


def _declare(L: C.CDLL) -> None:
    P = C.POINTER
    sig = {
        # Part 1 — reference operator API
        "InitInterResults": (C.c_int, [P(P(CInterRes)), C.c_int]),
        "FreeInterResults": (None, [P(CInterRes)]),
        "Filter": (P(CResult), [P(CInterRes), P(CFilterPred), P(CRelationMap), P(C.c_int)]),
        "InsertSingleRowIdsToInterResult": (C.c_int, [P(P(CInterRes)), C.c_int, P(CResult)]),
        "GetRelation": (P(CRelation), [C.c_int, C.c_int, P(CInterRes), P(CRelationMap), P(C.c_int)]),
        "RadixHashJoin": (P(CResult), [P(CRelation), P(CRelation), C.c_void_p]),
        "InsertJoinToInterResults": (C.c_int, [P(CInterRes), C.c_int, C.c_int, P(CResult)]),
        "AreActiveInInter": (C.c_int, [P(CInterRes), C.c_int, C.c_int]),
        "JoinInterNode": (C.c_int, [P(P(CInterRes)), P(CRelationMap), C.c_int, C.c_int, C.c_int, C.c_int,
                                    P(C.c_int)]),
        "MergeInterNodes": (None, [P(P(CInterRes))]),
        "CartesianInterResults": (None, [P(P(CInterRes))]),
        "CalculateQueryResults": (None, [P(CInterRes), P(CRelationMap), P(CBatchListnode)]),
        "PrintNullResults": (None, [P(CBatchListnode)]),
        "SelfJoin": (P(CResult), [C.c_int, C.c_int, C.c_int, P(P(CInterRes)), P(CRelationMap), P(C.c_int)]),
        "FreeResult": (None, [P(CResult)]),
        "FreeRelation": (None, [P(CRelation)]),
        # Part 2 — the shim
        "b200_init": (C.c_int, [C.c_int]),
        "b200_shutdown": (None, []),
        "b200_set_thread_device": (C.c_int, [C.c_int]),
        "b200_last_error": (C.c_char_p, []),
        "b200_is_cuda": (C.c_int, []),
        "b200_register_relations": (C.c_int, [P(CRelationMap), C.c_int]),
        "b200_compute_column_stats": (C.c_int, [P(CRelationMap), C.c_int]),
        "b200_register_device_column": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
        "b200_upload_column": (C.c_int, [C.c_void_p, C.c_uint64]),
        "b200_unregister_all": (None, []),
        "b200_unregister_relations": (C.c_int, [P(CRelationMap), C.c_int]),
        "b200_device_malloc": (C.c_void_p, [C.c_uint64]),
        "b200_device_free": (None, [C.c_void_p]),
        "b200_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
        "b200_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
        "b200_get_stream": (C.c_void_p, []),
        "b200_set_stream": (C.c_int, [C.c_void_p]),
        "b200_synchronize": (C.c_int, []),
        "b200_calculate_sums": (C.c_int, [P(CInterRes), P(CRelationMap), P(CBatchListnode), u64p, u64p]),
        "b200_set_lazy_join": (C.c_int, [C.c_int]),
        "b200_set_fuse_filters": (C.c_int, [C.c_int]),
        "b200_last_result_null": (C.c_int, []),
        "b200_result_kind": (C.c_int, [P(CResult)]),
        "b200_result_rowids_to_host": (C.c_int, [P(CResult), u64p]),
        "b200_result_pairs_to_host": (C.c_int, [P(CResult), u64p, u64p]),
        "b200_inter_column_to_host": (C.c_int, [P(CInterRes), C.c_int, u64p]),
        "b200_scan_filter": (C.c_int, [u64p, C.c_uint64, u64p, C.c_uint64, C.c_char, C.c_int, u64p, u64p]),
        "b200_radix_partition": (C.c_int, [u64p, C.c_uint64, C.c_int, u64p, u64p, u64p, i64p]),
        "b200_hash_join_pairs": (C.c_int, [u64p, C.c_uint64, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p]),
        "b200_gather_sum": (C.c_int, [u64p, C.c_uint64, u64p, C.c_uint64, u64p]),
        "b200_synth_column": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64]),
        "b200_set_tuning": (C.c_int, [C.c_int, C.c_int]),
        "b200_join_sum": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int,
                                    P(C.c_void_p), P(C.c_int), C.c_int, u64p, u64p]),
        "b200_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p]),
        "b200_ipc_import": (C.c_void_p, [C.c_char_p]),
        "b200_ipc_close": (C.c_int, [C.c_void_p]),
        "b200_radix_bits_for": (C.c_int, [C.c_uint64]),
        "b200_stage_hist": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
        "b200_stage_scatter_build": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_int, P(C.c_void_p), C.c_int, P(C.c_void_p), P(C.c_void_p),
                                               C.c_int]),
        "b200_stage_scatter_probe": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
        "b200_opt_region_cap": (C.c_uint32, [C.c_uint64, C.c_int]),
        "b200_stage_scatter_build_local": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p,
                                                     C.c_void_p, C.c_int, P(C.c_void_p), P(C.c_void_p)]),
        "b200_copy_device_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
        "b200_stage_join_sum_seg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, P(C.c_void_p), P(C.c_int), P(C.c_void_p), C.c_uint32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, u64p, u64p]),
        "b200_stage_build_cursors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
        "b200_stage_exchange_cursors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]),
        "b200_stage_exchange_segments": (C.c_int, [C.c_void_p, C.c_int, P(C.c_void_p), C.c_uint64, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, P(C.c_void_p),
                                                   P(C.c_void_p)]),
        "b200_stage_join_sum_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                P(C.c_void_p), P(C.c_int), P(C.c_void_p), C.c_uint32, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
        "b200_stage_scatter_probe_opt": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p]),
        "b200_stage_scatter_probe_opt_carry": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "b200_stage_join_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          P(C.c_void_p), P(C.c_int), P(C.c_void_p), C.c_uint32, C.c_void_p,
                                          C.c_void_p, u64p, u64p]),
        "b200_multi_create": (C.c_void_p, [P(CMultiConfig)]),
        "b200_multi_destroy": (None, [C.c_void_p]),
        "b200_multi_export": (C.c_int, [C.c_void_p, C.c_char_p]),
        "b200_multi_shared_ptr": (C.c_void_p, [C.c_void_p]),
        "b200_multi_connect_ipc": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
        "b200_multi_connect_ptr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
        "b200_multi_radix_bits": (C.c_int, [C.c_void_p]),
        "b200_multi_enqueue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
        "b200_multi_finish": (C.c_int, [C.c_void_p, u64p, u64p]),
        "b200_multi_received": (C.c_int, [C.c_void_p, u64p]),
        "b200_join_sum_multi": (C.c_int, [C.c_int, C.c_int, P(C.c_void_p), P(C.c_void_p), u64p, P(C.c_void_p),
                                          P(C.c_void_p), u64p, C.c_int, u64p, u64p, P(C.c_double)]),
        "b200_set_profiling": (C.c_int, [C.c_int]),
        "b200_reserve_device_memory": (C.c_uint64, [C.c_uint64]),
        "b200_last_kernel_ms": (C.c_double, [C.c_char_p]),
        "b200_sum_kernel_ms": (C.c_double, [C.c_char_p, C.POINTER(C.c_int)]),
        "b200_kernel_launches": (C.c_uint64, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libb200join: " + (lib().b200_last_error() or b"error").decode())


# --------------------------------------------------------------------------
# relation_map (reference structs.h:133-139) over numpy columns
# --------------------------------------------------------------------------
class RelationMapArray:
    """`relation_map rel_map[n]` as the reference's handler.c:51-52 builds it,
    with columns pointing at numpy arrays instead of an mmap."""

    def __init__(self, relations: list[list[np.ndarray]]):
        self.columns = [[_u64(c) for c in rel] for rel in relations]
        n = len(self.columns)
        self.array = (CRelationMap * n)()
        self._keep = []
        for r, cols in enumerate(self.columns):
            rows = len(cols[0]) if cols else 0
            ptrs = (u64p * len(cols))(*[_ptr(c) for c in cols])
            stats = (CColumnStats * len(cols))()
            for j, c in enumerate(cols):
                stats[j].l = int(c.min()) if rows else 0
                stats[j].u = int(c.max()) if rows else 0
                stats[j].f = float(rows)
                stats[j].d = float(len(np.unique(c))) if rows <= (1 << 22) else float(rows)
            self._keep += [ptrs, stats]
            self.array[r].num_tuples = rows
            self.array[r].num_columns = len(cols)
            self.array[r].columns = ptrs
            self.array[r].col_stats = stats
        # device copies are keyed by host pointer: drop entries a previous, freed
        # array may have left under the same addresses
        self.unregister()

    def __len__(self):
        return len(self.columns)

    def register(self):
        _check(lib().b200_register_relations(self.array, len(self)))

    def device_stats(self):
        """relation_map.c:53-83 computed on the GPU: per relation and column (l, u, f, d)."""
        _check(lib().b200_compute_column_stats(self.array, len(self)))
        return [[(int(self.array[r].col_stats[j].l), int(self.array[r].col_stats[j].u),
                  float(self.array[r].col_stats[j].f), float(self.array[r].col_stats[j].d))
                 for j in range(len(cols))] for r, cols in enumerate(self.columns)]

    def unregister(self):
        if _lib is not None:
            _lib.b200_unregister_relations(self.array, len(self))

    def __del__(self):
        try:
            self.unregister()
        except Exception:
            pass


class DeviceRelationMap:
    """`relation_map rel_map[n]` over columns that are already resident in HBM (bench.py: data generated on the
    device).  relations[r] = [(device_ptr, rows, max_value), ...].  The library keys device copies by the column
    pointer the caller's relation_map holds (b200_register_device_column); here that key is the device address
    itself, so nothing is uploaded."""

    def __init__(self, relations):
        n = len(relations)
        self.array = (CRelationMap * n)()
        self._keep = []
        L = lib()
        for r, cols in enumerate(relations):
            rows = cols[0][1] if cols else 0
            ptrs = (u64p * len(cols))(*[C.cast(C.c_void_p(p), u64p) for p, _, _ in cols])
            stats = (CColumnStats * len(cols))()
            for j, (p, nrows, mx) in enumerate(cols):
                stats[j].l, stats[j].u, stats[j].f, stats[j].d = 0, int(mx), float(nrows), float(nrows)
                _check(L.b200_register_device_column(p, p, nrows, int(mx)))
            self._keep += [ptrs, stats]
            self.array[r].num_tuples = rows
            self.array[r].num_columns = len(cols)
            self.array[r].columns = ptrs
            self.array[r].col_stats = stats

    def __len__(self):
        return len(self.array)

    def register(self):
        pass

    def unregister(self):
        if _lib is not None:
            _lib.b200_unregister_relations(self.array, len(self))


# --------------------------------------------------------------------------
# query text -> predicates (reference query.c:44-249 semantics)
# --------------------------------------------------------------------------
class ParsedQuery:
    def __init__(self, relations, filters, joins, views):
        self.relations, self.filters, self.joins, self.views = relations, filters, joins, views


def parse_query(text: str) -> ParsedQuery:
    """`r0 r1|preds|b.c b.c`.  A predicate whose right side has no '.' is a
    filter (query.c:120-130); filters are pushed to the list head, i.e. run in
    reverse textual order (query.c:150-157); joins keep textual order.
    A constant on the left (`3<0.1`) is not mirrored and is
    rejected here."""
    rel_s, pred_s, view_s = text.strip().split("|")
    relations = [int(t) for t in rel_s.split()]
    filters, joins = [], []
    for p in pred_s.split("&"):
        m = re.fullmatch(r"(\d+)\.(\d+)([<>=])(\d+)(?:\.(\d+))?", p.strip())
        if not m:
            raise ValueError(f"unsupported predicate {p!r}")
        b1, c1, op, rhs, c2 = m.groups()
        if c2 is None:
            filters.insert(0, (int(b1), int(c1), op, int(rhs)))
        else:
            if op != "=":
                raise ValueError("joins are equi-joins")
            joins.append((int(b1), int(c1), int(rhs), int(c2)))
    views = [(int(v[0]), int(v[2])) for v in view_s.split()]   # single digits, inter_res.c:325-327
    return ParsedQuery(relations, filters, joins, views)


class QueryResult:
    def __init__(self, sums, rows):
        self.sums, self.rows = sums, rows

    def line(self) -> str:
        if self.sums is None:
            return " ".join(["NULL"] * self.rows)   # rows = number of views here
        return " ".join(str(s) for s in self.sums)


def _make_batch_node(q: ParsedQuery):
    rels = (C.c_int * len(q.relations))(*q.relations)
    strs = [f"{b}.{c}".encode() for b, c in q.views]
    data = (C.c_char_p * len(strs))(*strs)
    views = CQueryStringArray(data, len(strs))
    node = CBatchListnode()
    node.num_of_relations = len(q.relations)
    node.relations = rels
    node.predicate_list = None
    node.views = C.pointer(views)
    node.next = None
    node._keep = (rels, strs, data, views)
    return node


def execute_query(text: str, rel_map: RelationMapArray) -> QueryResult:
    """The reference's ExecuteQuery (query.c:325-467) over the operator API:
    filters first (NULL result => the whole query is NULL, query.c:360-369),
    then the joins — AreActiveInInter ? JoinInterNode : GetRelation x2 ->
    RadixHashJoin -> InsertJoinToInterResults -> MergeInterNodes — then
    CartesianInterResults and the SUM projection.  Joins run in textual order
    (the reference reorders them with JoinEnum, which changes cost, not
    results)."""
    L = lib()
    q = parse_query(text)
    node = _make_batch_node(q)
    rels = node.relations
    inter = C.POINTER(CInterRes)()
    L.InitInterResults(C.byref(inter), len(q.relations))
    try:
        for b, c, op, k in q.filters:
            fp = CFilterPred(b, c, k, op.encode())
            res = L.Filter(inter, C.byref(fp), rel_map.array, rels)
            if not res:
                return QueryResult(None, len(q.views))
            L.InsertSingleRowIdsToInterResult(C.byref(inter), b, res)
            L.FreeResult(res)
        for b1, c1, b2, c2 in q.joins:
            if b1 == b2:
                res = L.SelfJoin(b1, c1, c2, C.byref(inter), rel_map.array, rels)
                if not res:
                    return QueryResult(None, len(q.views))
                L.InsertSingleRowIdsToInterResult(C.byref(inter), b1, res)
                L.FreeResult(res)
                continue
            if L.AreActiveInInter(inter, b1, b2):
                L.JoinInterNode(C.byref(inter), rel_map.array, b1, c1, b2, c2, rels)
                continue
            r1 = L.GetRelation(b1, c1, inter, rel_map.array, rels)
            r2 = L.GetRelation(b2, c2, inter, rel_map.array, rels)
            res = L.RadixHashJoin(r1, r2, None)
            L.FreeRelation(r1)
            L.FreeRelation(r2)
            if not res:
                return QueryResult(None, len(q.views))
            L.InsertJoinToInterResults(inter, b1, b2, res)
            L.FreeResult(res)
            if inter.contents.next:
                L.MergeInterNodes(C.byref(inter))
        if inter.contents.next:
            L.CartesianInterResults(C.byref(inter))
        sums = (C.c_uint64 * len(q.views))()
        rows = C.c_uint64(0)
        _check(L.b200_calculate_sums(inter, rel_map.array, C.byref(node), sums, C.byref(rows)))
        if L.b200_last_result_null():      # a filter fused into a join let nothing through (query.c:360-369)
            return QueryResult(None, len(q.views))
        return QueryResult([int(s) for s in sums], int(rows.value))
    finally:
        L.FreeInterResults(inter)


def execute_batch(queries: list[str], rel_map: RelationMapArray, workers: int = 4) -> list[QueryResult]:
    """A batch of queries on `workers` host threads (the reference's scheduler.c as a
    stream scheduler).  The reference runs a batch sequentially on one thread (handler.c:78-89) and its
    job queue cannot overlap two queries (one global barrier counter, structs.h:223).  Here a job is a
    whole query: every worker thread owns a CUDA stream inside the library (thread-local context), ctypes
    releases the GIL during the calls, and results come back in submission order so the output stays the
    reference's."""
    from concurrent.futures import ThreadPoolExecutor
    rel_map.register()          # uploads happen once, before the workers start
    if workers <= 1:
        return [execute_query(q, rel_map) for q in queries]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(lambda q: execute_query(q, rel_map), queries))


# --------------------------------------------------------------------------
# kernel-level entry points (host buffers in and out)
# --------------------------------------------------------------------------
def scan_filter(col, cmp: str, value: int, ids=None) -> np.ndarray:
    col = _u64(col)
    idv = _u64(ids) if ids is not None else None
    n_out = len(idv) if idv is not None else len(col)
    out = np.empty(max(n_out, 1), dtype=np.uint64)
    cnt = C.c_uint64(0)
    _check(lib().b200_scan_filter(_ptr(col), len(col), _ptr(idv) if idv is not None else None,
                                  len(idv) if idv is not None else 0, cmp.encode(), int(value), _ptr(out),
                                  C.byref(cnt)))
    return out[: cnt.value].copy()


def radix_partition(keys, radix_bits: int):
    keys = _u64(keys)
    n = len(keys)
    ok, orid = np.empty(max(n, 1), np.uint64), np.empty(max(n, 1), np.uint64)
    hist = np.empty(1 << radix_bits, np.uint64)
    psum = np.empty(1 << radix_bits, np.int64)
    _check(lib().b200_radix_partition(_ptr(keys), n, radix_bits, _ptr(ok), _ptr(orid), _ptr(hist),
                                      psum.ctypes.data_as(i64p)))
    return ok[:n], orid[:n], hist, psum


def hash_join_pairs(keys_r, keys_s, cap: int | None = None):
    kr, ks = _u64(keys_r), _u64(keys_s)
    m = C.c_uint64(0)
    if cap is None:
        dummy = np.empty(1, np.uint64)
        _check(lib().b200_hash_join_pairs(_ptr(kr), len(kr), _ptr(ks), len(ks), _ptr(dummy), _ptr(dummy), 0,
                                          C.byref(m)))
        cap = int(m.value)
    out_r, out_s = np.empty(max(cap, 1), np.uint64), np.empty(max(cap, 1), np.uint64)
    _check(lib().b200_hash_join_pairs(_ptr(kr), len(kr), _ptr(ks), len(ks), _ptr(out_r), _ptr(out_s), cap,
                                      C.byref(m)))
    k = min(cap, int(m.value))
    return out_r[:k], out_s[:k], int(m.value)


def gather_sum(col, ids) -> int:
    col, ids = _u64(col), _u64(ids)
    out = C.c_uint64(0)
    _check(lib().b200_gather_sum(_ptr(col), len(col), _ptr(ids), len(ids), C.byref(out)))
    return int(out.value)


def join_sum(keys_r, keys_s, proj, proj_side, max_key: int | None = None):
    """Fused join -> SUM with HOST buffers (the end-to-end entry point)."""
    kr, ks = _u64(keys_r), _u64(keys_s)
    pj = [_u64(p) for p in proj]
    if max_key is None:
        max_key = int(max(kr.max() if len(kr) else 0, ks.max() if len(ks) else 0))
    ptrs = (C.c_void_p * max(len(pj), 1))(*[p.ctypes.data for p in pj])
    sides = (C.c_int * max(len(pj), 1))(*proj_side)
    sums = (C.c_uint64 * max(len(pj), 1))()
    m = C.c_uint64(0)
    _check(lib().b200_join_sum(kr.ctypes.data, len(kr), ks.ctypes.data, len(ks), max_key, len(pj), ptrs, sides, 0,
                               sums, C.byref(m)))
    return [int(s) for s in sums[: len(pj)]], int(m.value)


def join_sum_device(keys_r: int, n_r: int, keys_s: int, n_s: int, proj: list[int], proj_side: list[int],
                    max_key: int):
    """Fused join -> SUM on DEVICE addresses already resident in HBM."""
    ptrs = (C.c_void_p * max(len(proj), 1))(*proj)
    sides = (C.c_int * max(len(proj), 1))(*proj_side)
    sums = (C.c_uint64 * max(len(proj), 1))()
    m = C.c_uint64(0)
    _check(lib().b200_join_sum(keys_r, n_r, keys_s, n_s, max_key, len(proj), ptrs, sides, 1, sums, C.byref(m)))
    return [int(s) for s in sums[: len(proj)]], int(m.value)


class DeviceColumn:
    """n uint64 values in HBM owned through the C-ABI (no torch)."""

    def __init__(self, n: int):
        self.n = n
        self.ptr = lib().b200_device_malloc(8 * n)
        if not self.ptr:
            raise MemoryError("b200_device_malloc failed")

    def to_host(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.empty(max(count, 1), np.uint64)
        _check(lib().b200_copy_to_host(out.ctypes.data, self.ptr + 8 * first, 8 * count))
        return out[:count]

    def free(self):
        if self.ptr:
            lib().b200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synth_column_device(device_ptr: int, first: int, n: int, kind: int, k: int, seed: int) -> None:
    _check(lib().b200_synth_column(device_ptr, first, n, kind, k, seed & 0xFFFFFFFFFFFFFFFF))


def kernel_launches(reset: bool = False) -> int:
    return int(lib().b200_kernel_launches(1 if reset else 0))


def last_kernel_ms(name: str) -> float:
    return float(lib().b200_last_kernel_ms(name.encode()))
