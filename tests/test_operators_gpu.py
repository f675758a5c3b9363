"""The reference's operator API on the GPU library, driven the way the
reference's only caller (query.c:ExecuteQuery) drives it, against the golden
lines of the reference, the oracle executor and the shipped `small` workload."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_cases import GOLDEN, col, load_small, query_relations, small_queries

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("i", range(len(GOLDEN["queries"])))
def test_query_golden_lines(gpu, i):
    case = GOLDEN["queries"][i]
    rm = gpu.RelationMapArray(query_relations(case))
    assert gpu.execute_query(case["query"], rm).line() == case["line"]


EXTRA = [
    "0 1 2|0.0=1.0&0.1=2.1&1.2=2.2|2.0 1.1",                 # triangle: JoinInterNode
    "0 1 2|0.0=1.0&0.1<30&1.1>20&2.2=5&1.2=2.1|0.2 1.0 2.0",  # filters on three bindings
    "0 1 2 0|0.0=1.0&1.1=2.1&2.2=3.2&3.0<20|3.1 0.1",
    "0 1|0.1=0.2&0.0=1.0|1.1 0.0",                            # same-binding self join
    "0 1 2|0.0=1.0&0.1=1.1&1.2=2.2&0.2=2.0|0.0 1.0 2.0",
    "0 1|0.0=1.0|0.0 0.0 1.1 1.1",                            # duplicated projections
]


@pytest.mark.parametrize("q", EXTRA)
def test_queries_against_oracle_executor(gpu, orc, q):
    rels = [[col(n, 48, 4000 + 10 * r + c) for c in range(3)] for r, n in enumerate([700, 1100, 400])]
    rm = gpu.RelationMapArray(rels)
    assert gpu.execute_query(q, rm).line() == orc.execute_query(q, rels)


def test_operator_level_walkthrough(gpu, orc):
    """Filter -> InsertSingleRowIds -> GetRelation -> RadixHashJoin ->
    InsertJoinToInterResults, reading every intermediate back (Appendix B)."""
    import ctypes as C
    h, L = gpu.host, gpu.lib()
    rels = [[col(3000, 200, 50), col(3000, 50, 51)], [col(5000, 200, 52), col(5000, 1 << 20, 53)]]
    rm = gpu.RelationMapArray(rels)
    rm.register()
    binds = (C.c_int * 2)(0, 1)
    inter = C.POINTER(h.CInterRes)()
    L.InitInterResults(C.byref(inter), 2)
    fp = h.CFilterPred(0, 1, 25, b"<")
    res = L.Filter(inter, C.byref(fp), rm.array, binds)
    want_ids = orc.filter_scan(rels[0][1], "<", 25)
    assert res and res.contents.current_load == len(want_ids) and L.b200_result_kind(res) == 1
    got = np.empty(len(want_ids), np.uint64)
    L.b200_result_rowids_to_host(res, got.ctypes.data_as(h.u64p))
    assert np.array_equal(np.sort(got), want_ids)
    L.InsertSingleRowIdsToInterResult(C.byref(inter), 0, res)
    L.FreeResult(res)
    assert inter.contents.data.contents.num_tuples == len(want_ids)
    r0 = L.GetRelation(0, 0, inter, rm.array, binds)
    r1 = L.GetRelation(1, 0, inter, rm.array, binds)
    assert r0.contents.num_tuples == len(want_ids) and r1.contents.num_tuples == 5000
    jr = L.RadixHashJoin(r0, r1, None)
    L.FreeRelation(r0)
    L.FreeRelation(r1)
    t0 = np.empty(len(want_ids), np.uint64)
    L.b200_inter_column_to_host(inter, 0, t0.ctypes.data_as(h.u64p))
    o_r, o_s = orc.radix_hash_join(rels[0][0][t0.astype(np.int64)], rels[1][0], 4)
    # (with the lazy last join the result is deferred until it is looked at: ask for its kind first)
    assert jr and L.b200_result_kind(jr) == 2 and jr.contents.current_load == len(o_r)
    L.InsertJoinToInterResults(inter, 0, 1, jr)
    L.FreeResult(jr)
    m = len(o_r)
    c0, c1 = np.empty(m, np.uint64), np.empty(m, np.uint64)
    L.b200_inter_column_to_host(inter, 0, c0.ctypes.data_as(h.u64p))
    L.b200_inter_column_to_host(inter, 1, c1.ctypes.data_as(h.u64p))
    assert np.array_equal(rels[0][0][c0.astype(np.int64)], rels[1][0][c1.astype(np.int64)])
    want = np.stack([t0[o_r.astype(np.int64)], o_s])
    assert np.array_equal(np.sort(c0 * np.uint64(1 << 20) + c1), np.sort(want[0] * np.uint64(1 << 20) + want[1]))
    assert L.AreActiveInInter(inter, 0, 1) == 1
    L.FreeInterResults(inter)


@pytest.mark.parametrize("lazy", [0, 1])
def test_lazy_last_join_equals_eager(gpu, orc, lazy):
    """Every query shape of this file, with the last join parked and fused into the SUMs (lazy) and with every
    join materialised (eager): the same lines as the oracle's executor; plus the parked join being forced by a
    read-back, by a second join on the same node, and surviving up to the SUMs."""
    import ctypes as C
    h, L = gpu.host, gpu.lib()
    before = L.b200_set_lazy_join(lazy)
    try:
        rels = [[col(n, 48, 4000 + 10 * r + c) for c in range(3)] for r, n in enumerate([700, 1100, 400])]
        rm = gpu.RelationMapArray(rels)
        for q in EXTRA:
            assert gpu.execute_query(q, rm).line() == orc.execute_query(q, rels), q
        small = load_small()
        if small is not None:
            rms = gpu.RelationMapArray(small)
            rms.register()
            queries, golden = small_queries()
            for q, want in zip(queries, golden):
                assert gpu.execute_query(q, rms).line() == want, q
        # operator level: a parked join read back through the intermediate equals the eager pairs
        rels2 = [[col(3000, 200, 50), col(3000, 50, 51)], [col(5000, 200, 52), col(5000, 1 << 20, 53)]]
        rm2 = gpu.RelationMapArray(rels2)
        rm2.register()
        binds = (C.c_int * 2)(0, 1)
        inter = C.POINTER(h.CInterRes)()
        L.InitInterResults(C.byref(inter), 2)
        r0 = L.GetRelation(0, 0, inter, rm2.array, binds)
        r1 = L.GetRelation(1, 0, inter, rm2.array, binds)
        jr = L.RadixHashJoin(r0, r1, None)
        L.FreeRelation(r0)
        L.FreeRelation(r1)
        L.InsertJoinToInterResults(inter, 0, 1, jr)
        L.FreeResult(jr)
        assert L.AreActiveInInter(inter, 0, 1) == 1
        o_r, o_s = orc.radix_hash_join(rels2[0][0], rels2[1][0], 4)
        want = [orc.checksum(rels2[0][1], o_r), orc.checksum(rels2[1][1], o_s)]
        views = [b"0.1", b"1.1"]
        arr = (C.c_char_p * 2)(*views)
        qsa = h.CQueryStringArray(C.cast(arr, C.POINTER(C.c_char_p)), 2)
        node = h.CBatchListnode(2, binds, None, C.pointer(qsa), None)
        sums = (C.c_uint64 * 2)()
        rows = C.c_uint64(0)
        assert L.b200_calculate_sums(inter, rm2.array, C.byref(node), sums, C.byref(rows)) == 0
        assert [int(sums[0]), int(sums[1])] == want and rows.value == len(o_r)
        c0 = np.empty(len(o_r), np.uint64)
        assert L.b200_inter_column_to_host(inter, 0, c0.ctypes.data_as(h.u64p)) == 0      # forces the parked join
        assert np.array_equal(np.sort(c0), np.sort(o_r))
        assert inter.contents.data.contents.num_tuples == len(o_r)
        assert L.b200_calculate_sums(inter, rm2.array, C.byref(node), sums, C.byref(rows)) == 0
        assert [int(sums[0]), int(sums[1])] == want and rows.value == len(o_r)
        L.FreeInterResults(inter)
    finally:
        L.b200_set_lazy_join(before)


def test_small_workload_execute_query(gpu):
    """BASELINE config 1 at operator level: all 50 small.work queries equal the
    reference's golden small.result."""
    rels = load_small()
    if rels is None:
        pytest.skip("small workload data not present (oracle/_ref/small)")
    rm = gpu.RelationMapArray(rels)
    rm.register()
    queries, golden = small_queries()
    for q, want in zip(queries, golden):
        assert gpu.execute_query(q, rm).line() == want, q


@pytest.mark.parametrize("workers", [2, 8])
def test_small_workload_concurrent_queries_on_streams(gpu, workers):
    """BASELINE config 5's mechanism at the shipped scale: the 50 queries of small.work run concurrently,
    one CUDA stream per worker thread, and still produce small.result line by line."""
    rels = load_small()
    if rels is None:
        pytest.skip("small workload data not present (oracle/_ref/small)")
    rm = gpu.RelationMapArray(rels)
    queries, golden = small_queries()
    got = gpu.execute_batch(queries * 2, rm, workers=workers)
    assert [r.line() for r in got] == golden * 2


@pytest.mark.parametrize("workers", [1, 4])
def test_small_workload_standalone_c_host(workers):
    """BASELINE config 1 through host/b200_engine.c — our own C host (protocol of handler.c, loader of
    relation_map.c, parser of query.c, a scheduler whose jobs are whole queries on per-thread CUDA streams) —
    fed the harness protocol on stdin; output must equal the reference's small.result."""
    exe = ROOT / "host" / "b200_engine"
    small = ROOT / "oracle" / "_ref" / "small"
    if not exe.exists() or not (small / "r0").exists():
        pytest.skip("host/b200_engine or the small workload data not present")
    stdin = (small / "small.init").read_text() + "Done\n" + (small / "small.work").read_text()
    out = subprocess.run([str(exe), "-w", str(workers)], input=stdin, capture_output=True, text=True, cwd=small,
                         timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout == (small / "small.result").read_text()


def test_small_workload_dropin_binary():
    """BASELINE config 1 through the link-time drop-in: the reference's own
    handler.o/query.o/best_tree.o/stats.o/relation_map.o over libb200join.so,
    fed the harness protocol on stdin; output must equal small.result."""
    exe = ROOT / "oracle" / "_ref" / "radixhash_b200"
    small = ROOT / "oracle" / "_ref" / "small"
    if not exe.exists() or not (small / "r0").exists():
        pytest.skip("drop-in binary / small workload not built (make -C oracle ref dropin)")
    stdin = (small / "small.init").read_text() + "Done\n" + (small / "small.work").read_text()
    out = subprocess.run([str(exe)], input=stdin, capture_output=True, text=True, cwd=small, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout == (small / "small.result").read_text()


def test_small_workload_through_the_contest_harness():
    exe = ROOT / "oracle" / "_ref" / "radixhash_b200"
    harness = ROOT / "oracle" / "_ref" / "harness"
    small = ROOT / "oracle" / "_ref" / "small"
    if not exe.exists() or not harness.exists() or not (small / "r0").exists():
        pytest.skip("harness / drop-in binary not built")
    out = subprocess.run(["timeout", "120", str(harness), "small.init", "small.work", "small.result", str(exe)],
                         capture_output=True, text=True, cwd=small, timeout=180)
    assert out.returncode == 0, (out.stdout + out.stderr)[-2000:]
    assert int(out.stdout.strip().split()[-1]) >= 0     # elapsed ms


# ---- BASELINE config 3 (scaled down): 4-way chain join, range/equality filters on the fact
# relation, 3-column SUM projection — the reference itself (oracle/_ref/ref_driver), the link-time
# drop-in (the reference's query.o/best_tree.o over libb200join.so) and the operator API must print
# the same line.  All filters sit on binding 0, constants < 2^31 (quirk 5).
CONFIG3_QUERY = "0 1 2 3|0.1=1.0&1.1=2.0&2.1=3.0&0.3>2499&0.3<7500&0.4=1|0.2 1.2 3.1"
CONFIG3_SPECS = [
    "synth:{f}:iota,uni{d1}@11,pay@12,uni10000@13,uni4@14",       # F: pk, fk->D1, measure, filter cols
    "synth:{d1}:iota,uni{d2}@21,pay@22",                           # D1: pk, fk->D2, payload
    "synth:{d2}:iota,uni{d3}@31,pay@32",                           # D2
    "synth:{d3}:iota,pay@41",                                      # D3
]


def _config3_relations(orc, f, d1, d2, d3):
    iota = lambda n: np.arange(n, dtype=np.uint64)
    uni = lambda n, mod, seed: orc.synth_column(n, 3, mod, seed)
    pay = lambda n, seed: orc.synth_column(n, 1, 0, seed)
    return [[iota(f), uni(f, d1, 11), pay(f, 12), uni(f, 10000, 13), uni(f, 4, 14)],
            [iota(d1), uni(d1, d2, 21), pay(d1, 22)],
            [iota(d2), uni(d2, d3, 31), pay(d2, 32)],
            [iota(d3), pay(d3, 41)]]


@pytest.mark.parametrize("f,d1,d2,d3", [(200_000, 1 << 14, 1 << 10, 1 << 6), (3_000_000, 1 << 18, 1 << 14, 1 << 10)])
def test_config3_chain_join_matches_the_reference(gpu, orc, f, d1, d2, d3):
    ref_driver = ROOT / "oracle" / "_ref" / "ref_driver"
    if not ref_driver.exists():
        pytest.skip("oracle/_ref/ref_driver not built")
    specs = [s.format(f=f, d1=d1, d2=d2, d3=d3) for s in CONFIG3_SPECS]
    want = subprocess.run([str(ref_driver), "-t", "4", *specs, "--", CONFIG3_QUERY], capture_output=True, text=True,
                          timeout=600)
    assert want.returncode == 0, want.stderr[-1000:]
    want_line = want.stdout.splitlines()[0]
    assert "NULL" not in want_line
    # operator API (textual join order)
    rels = _config3_relations(orc, f, d1, d2, d3)
    rm = gpu.RelationMapArray(rels)
    assert gpu.execute_query(CONFIG3_QUERY, rm).line() == want_line
    # the reference's own ExecuteQuery + JoinEnum over the CUDA operators
    drop = ROOT / "oracle" / "_ref" / "b200_driver"
    if drop.exists():
        got = subprocess.run([str(drop), "-t", "1", *specs, "--", CONFIG3_QUERY], capture_output=True, text=True,
                             timeout=600)
        assert got.returncode == 0, got.stderr[-1000:]
        assert got.stdout.splitlines()[0] == want_line


# ---- differential fuzz: random conjunctive queries over random relations, GPU operators vs the oracle ----
def _random_query(g, nrel, ncols):
    nb = int(g.integers(2, 5))                                  # bindings
    rels = [int(g.integers(0, nrel)) for _ in range(nb)]
    preds, joined = [], {0}
    order = list(range(1, nb))
    g.shuffle(order)
    for b in order:                                              # a spanning tree keeps the query connected
        a = int(g.choice(sorted(joined)))
        preds.append(f"{a}.{int(g.integers(0, ncols))}={b}.{int(g.integers(0, ncols))}")
        joined.add(b)
    for _ in range(int(g.integers(0, 3))):                       # extra edges: cycles / duplicate pairs / self joins
        a, b = int(g.integers(0, nb)), int(g.integers(0, nb))
        preds.append(f"{a}.{int(g.integers(0, ncols))}={b}.{int(g.integers(0, ncols))}")
    for _ in range(int(g.integers(0, 3))):                       # filters, possibly on several bindings
        op = "<>="[int(g.integers(0, 3))]
        preds.append(f"{int(g.integers(0, nb))}.{int(g.integers(0, ncols))}{op}{int(g.integers(0, 400))}")
    g.shuffle(preds)
    views = " ".join(f"{int(g.integers(0, nb))}.{int(g.integers(0, ncols))}" for _ in range(int(g.integers(1, 4))))
    return " ".join(map(str, rels)) + "|" + "&".join(preds) + "|" + views


@pytest.mark.parametrize("seed", range(6))
def test_random_queries_against_the_oracle(gpu, orc, seed):
    g = np.random.default_rng(1000 + seed)
    sizes = [int(x) for x in g.integers(1, 300, 4)]           # fan-out per join <= 6: results stay small
    rels = [[col(n, int(g.integers(50, 400)), 7000 + 100 * seed + 10 * r + c) for c in range(3)]
            for r, n in enumerate(sizes)]
    rm = gpu.RelationMapArray(rels)
    for _ in range(25):
        q = _random_query(g, len(rels), 3)
        assert gpu.execute_query(q, rm).line() == orc.execute_query(q, rels), q
