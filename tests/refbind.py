"""ctypes binding of the UNMODIFIED reference, compiled by oracle/Makefile into
oracle/_ref/libref_ops.so (+ the ref_driver / radixhash binaries).  Used to pin
the oracle restatement and to generate tests/golden/.  Struct layouts are the
reference's structs.h.  Test infrastructure only; never read /root/reference at
run time (only the prebuilt files under oracle/_ref/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / "oracle" / "_ref"
LIB = REF_DIR / "libref_ops.so"
N_LSB = 4   # structs.h:11, compiled in
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


class Tuple(C.Structure):
    _fields_ = [("value", C.c_uint64), ("row_id", C.c_uint64)]


class Relation(C.Structure):
    _fields_ = [("tuples", C.POINTER(Tuple)), ("num_tuples", C.c_uint64)]


class Result(C.Structure):
    pass


Result._fields_ = [("buff", C.c_void_p), ("next", C.POINTER(Result)), ("current_load", C.c_uint64)]


class Reordered(C.Structure):
    _fields_ = [("hist_size", C.c_int), ("psum", i64p), ("hist", u64p), ("rel_array", C.POINTER(Relation))]


class BcIndex(C.Structure):
    _fields_ = [("index_size", C.c_int), ("start", C.c_int), ("end", C.c_int), ("bucket", i64p), ("chain", i64p)]


_lib = None
_sched = {}


def available() -> bool:
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(LIB))
        L.HashFunction1.restype = C.c_uint64
        L.HashFunction1.argtypes = [C.c_uint64, C.c_uint64]
        L.HashFunction2.restype = C.c_uint64
        L.HashFunction2.argtypes = [C.c_uint64, C.c_uint64]
        L.FindNextPrime.restype = C.c_uint64
        L.FindNextPrime.argtypes = [C.c_uint64]
        L.SchedulerInit.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.RadixHashJoin.restype = C.POINTER(Result)
        L.RadixHashJoin.argtypes = [C.POINTER(Relation), C.POINTER(Relation), C.c_void_p]
        L.ReorderArray.restype = None
        L.ReorderArray.argtypes = [C.POINTER(Relation), C.POINTER(Relation), C.POINTER(C.POINTER(Reordered)),
                                   C.POINTER(C.POINTER(Reordered)), C.c_void_p]
        L.FreeReorderRelation.argtypes = [C.POINTER(Reordered)]
        L.InitIndex.argtypes = [C.POINTER(C.POINTER(BcIndex)), C.c_int, C.c_int]
        L.CreateIndex.argtypes = [C.POINTER(Reordered), C.POINTER(C.POINTER(BcIndex)), C.c_int]
        L.DeleteIndex.argtypes = [C.POINTER(C.POINTER(BcIndex))]
        L.FreeResult.argtypes = [C.POINTER(Result)]
        _lib = L
    return _lib


def scheduler(threads: int = 4):
    if threads not in _sched:
        s = C.c_void_p()
        lib().SchedulerInit(C.byref(s), threads)
        _sched[threads] = s
    return _sched[threads]


def _relation(keys: np.ndarray, rids=None):
    """GetRelation's output (inter_res.c:223-227): AoS {value, row_id = i}."""
    n = len(keys)
    aos = np.empty((max(n, 1), 2), np.uint64)
    aos[:n, 0] = keys
    aos[:n, 1] = np.arange(n, dtype=np.uint64) if rids is None else rids
    rel = Relation(aos.ctypes.data_as(C.POINTER(Tuple)), n)
    rel._keep = aos
    return rel


def find_next_prime(n: int) -> int:
    return int(lib().FindNextPrime(n))


def hash1(num: int, n: int) -> int:
    return int(lib().HashFunction1(num, n))


def radix_hash_join(keys_r, keys_s, threads: int = 4):
    """The reference's RadixHashJoin (rhjoin.c:13-111): pairs in list order, or
    None when it returns NULL."""
    L = lib()
    r, s = _relation(np.asarray(keys_r, np.uint64)), _relation(np.asarray(keys_s, np.uint64))
    res = L.RadixHashJoin(C.byref(r), C.byref(s), scheduler(threads))
    if not res:
        return None
    chunks = []
    node = res
    while node:
        n = int(node.contents.current_load)
        if n:
            buf = (C.c_uint64 * (2 * n)).from_address(node.contents.buff)
            chunks.append(np.frombuffer(buf, np.uint64).reshape(n, 2).copy())
        node = node.contents.next
    L.FreeResult(res)
    if not chunks:
        return np.empty(0, np.uint64), np.empty(0, np.uint64)
    allp = np.concatenate(chunks)
    return allp[:, 0].copy(), allp[:, 1].copy()


def reorder(keys_r, keys_s, threads: int = 4):
    """ReorderArray (preprocess.c:13-178) on N_LSB bits for both relations:
    [(keys, rids, hist, psum)] x 2, or None when it yields NULL."""
    L = lib()
    r, s = _relation(np.asarray(keys_r, np.uint64)), _relation(np.asarray(keys_s, np.uint64))
    nr, ns = C.POINTER(Reordered)(), C.POINTER(Reordered)()
    L.ReorderArray(C.byref(r), C.byref(s), C.byref(nr), C.byref(ns), scheduler(threads))
    if not nr or not ns:
        return None
    out = []
    for x in (nr, ns):
        rr = x.contents
        n = int(rr.rel_array.contents.num_tuples)
        aos = np.frombuffer((C.c_uint64 * (2 * n)).from_address(C.addressof(rr.rel_array.contents.tuples.contents)),
                            np.uint64).reshape(n, 2).copy()
        hist = np.array([rr.hist[i] for i in range(rr.hist_size)], np.uint64)
        psum = np.array([rr.psum[i] for i in range(rr.hist_size)], np.int64)
        out.append((aos[:, 0].copy(), aos[:, 1].copy(), hist, psum))
        L.FreeReorderRelation(x)
    return out


def create_index(keys):
    """InitIndex + CreateIndex (rhjoin.c:253-273, 219-250) on ONE bucket made
    of `keys` (all keys must share their low N_LSB bits)."""
    L = lib()
    keys = np.asarray(keys, np.uint64)
    n = len(keys)
    rel = _relation(keys)
    nb = 1 << N_LSB
    b = int(keys[0]) & (nb - 1)
    hist = (C.c_uint64 * nb)()
    psum = (C.c_int64 * nb)(*([-1] * nb))
    hist[b], psum[b] = n, 0
    rr = Reordered(nb, psum, hist, C.pointer(rel))
    ind = C.POINTER(BcIndex)()
    L.InitIndex(C.byref(ind), n, 0)
    L.CreateIndex(C.byref(rr), C.byref(ind), b)
    size = int(ind.contents.index_size)
    bucket = np.array([ind.contents.bucket[i] for i in range(size)], np.int64)
    chain = np.array([ind.contents.chain[i] for i in range(n)], np.int64)
    L.DeleteIndex(C.byref(ind))
    return size, bucket, chain


def write_relation_file(path, columns):
    """The contest's binary format (relation_map.c:39-50)."""
    cols = [np.ascontiguousarray(c, np.uint64) for c in columns]
    with open(path, "wb") as f:
        np.array([len(cols[0]), len(cols)], np.uint64).tofile(f)
        for c in cols:
            c.tofile(f)


def run_driver(binary: str, relations, queries, threads: int = 4, timeout: int = 600):
    """Run oracle/_ref/<binary> (ref_driver = the reference's ExecuteQuery on
    its own operators; b200_driver = the same query.o over libb200join.so) on
    in-memory relations written to temporary files; returns the output lines."""
    exe = REF_DIR / binary
    with tempfile.TemporaryDirectory() as tmp:
        specs = []
        for i, cols in enumerate(relations):
            p = os.path.join(tmp, f"r{i}")
            write_relation_file(p, cols)
            specs.append("file:" + p)
        out = subprocess.run([str(exe), "-t", str(threads), *specs, "--", *queries], capture_output=True,
                             text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"{binary} exited {out.returncode}: {out.stderr[-2000:]}")
    return out.stdout.splitlines()


class RelationListNode(C.Structure):
    pass


RelationListNode._fields_ = [("filename", C.c_char_p), ("fd", C.c_int), ("next", C.POINTER(RelationListNode))]


class ColumnStats(C.Structure):
    _fields_ = [("l", C.c_uint64), ("u", C.c_uint64), ("f", C.c_double), ("d", C.c_double)]


class RelationMap(C.Structure):
    _fields_ = [("num_tuples", C.c_uint64), ("num_columns", C.c_uint64), ("columns", C.POINTER(C.POINTER(C.c_uint64))),
                ("col_stats", C.POINTER(ColumnStats))]


def init_relation_map(paths):
    """The reference's loader (relation_map.c:13-88) on relation files: per relation and column (l, u, f, d)."""
    L = lib()
    L.InitRelationMap.argtypes = [C.POINTER(RelationListNode), C.POINTER(RelationMap)]
    nodes = [RelationListNode(str(p).encode(), -1, None) for p in paths]
    for a, b in zip(nodes, nodes[1:]):
        a.next = C.pointer(b)
    maps = (RelationMap * len(paths))()
    L.InitRelationMap(C.byref(nodes[0]), maps)
    out = []
    for m in maps:
        out.append([(int(m.col_stats[j].l), int(m.col_stats[j].u), float(m.col_stats[j].f), float(m.col_stats[j].d))
                    for j in range(m.num_columns)])
    return out
