"""Filter fusion: filters on large base relations are not scanned one by one (filter.c:92-190) but
evaluated inside the load stage of the join's partition kernels.  Every query runs fused and the eager way, with the
last join lazy and eager, against the oracle executor (query.c:325-467 restated): probe side, build side, an
unpartitioned join, through an intermediate, several bindings, more predicates than the fused set holds, and an
empty filter (the reference's NULL line)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N0, N1, N2 = 1 << 19, (1 << 20) + 333, 5000


@pytest.fixture(scope="module")
def rels(orc):
    uni = lambda n, mod, seed: orc.synth_column(n, 3, mod, seed)
    r0 = [uni(N0, 3000, 1), uni(N0, 1000, 2), uni(N0, 16, 3), uni(N0, 1 << 20, 4)]
    r1 = [uni(N1, 3000, 5), uni(N1, 1000, 6), uni(N1, 16, 7)]
    r2 = [np.arange(N2, dtype=np.uint64), uni(N2, 1000, 8), uni(N2, 3000, 9)]
    return [r0, r1, r2]


QUERIES = [
    "0 1|0.0=1.0&0.1<5|0.2 1.1 0.3",                          # filtered side is the smaller one: build side
    "0 1|0.0=1.0&1.1>100&1.1<130&1.2=7|0.2 1.1",              # three predicates on two columns, probe side
    "0 2|0.0=1.2&0.1<50|0.3 1.1",                             # large filtered relation probes an unpartitioned build side
    "2 0 1|0.2=1.0&1.1<3&1.0=2.0&2.1=999|0.1 1.2 2.2",        # both large bindings filtered, chain through an intermediate
    "2 0 1|0.2=1.0&1.0=2.0&2.1<2|0.1 2.2",                    # the filtered binding joins an intermediate
    "0 1|0.0=1.0&0.1<40&0.1>10&0.2=3&0.2<9&0.3>1000&0.3<900000|0.3 1.2",   # more predicates than the fused set holds
    "0 1|0.0=1.0&0.1>2000000000|0.2 1.1",                     # empty filter: NULL
    "0 1|0.0=1.0&1.1=5000|0.2",                               # empty filter on the probe side
    "0 1 2|0.0=1.0&0.1<2&0.2=2.0|1.1 2.1",                    # filtered binding joined twice
    "0 0|0.0=1.0&0.1<1&1.1>998|0.3 1.3",                      # two bindings of the same relation, both filtered
]


@pytest.mark.parametrize("q", QUERIES)
def test_fused_filters_match_the_oracle(gpu, orc, rels, q):
    L = gpu.lib()
    want = orc.execute_query(q, rels)
    rm = gpu.RelationMapArray(rels)
    rm.register()
    try:
        for fuse in (2, 1, 0):        # wherever possible / where the estimate says it pays / never
            for lazy in (1, 0):
                L.b200_set_fuse_filters(fuse)
                before = L.b200_set_lazy_join(lazy)
                try:
                    assert gpu.execute_query(q, rm).line() == want, (q, fuse, lazy)
                finally:
                    L.b200_set_lazy_join(before)
    finally:
        L.b200_set_fuse_filters(1)
        rm.unregister()


def test_fusion_is_chosen_by_estimated_selectivity(gpu, orc, rels):
    """Mode 1 (default): a filter that keeps a quarter of a large relation is fused — no scan, no row-id list, no
    host round trip of its own, so fewer kernels than the eager path; one that keeps 0.2 % is scanned the eager way
    (the join then takes the tiny unpartitioned plan) and launches exactly what mode 0 launches."""
    L = gpu.lib()
    rm = gpu.RelationMapArray(rels)
    rm.register()

    def launches(q, mode):
        L.b200_set_fuse_filters(mode)
        gpu.execute_query(q, rm)
        gpu.kernel_launches(reset=True)
        line = gpu.execute_query(q, rm).line()
        assert line == orc.execute_query(q, rels), (q, mode)
        return gpu.kernel_launches()

    try:
        wide, narrow = "0 1|0.0=1.0&1.1<500&1.1>0|0.2 1.1", "0 1|0.0=1.0&1.1>100&1.1<130&1.2=7|0.2 1.1"
        assert launches(wide, 1) == launches(wide, 2) < launches(wide, 0)
        assert launches(narrow, 1) == launches(narrow, 0)
    finally:
        L.b200_set_fuse_filters(1)
        rm.unregister()


def test_deferred_filter_result_reads_back_as_row_ids(gpu, orc, rels):
    """A deferred Filter result that is looked at directly is scanned on the spot."""
    import ctypes as C
    h, L = gpu.host, gpu.lib()
    rm = gpu.RelationMapArray(rels)
    rm.register()
    try:
        binds = (C.c_int * 2)(0, 1)
        inter = C.POINTER(h.CInterRes)()
        L.InitInterResults(C.byref(inter), 2)
        fp = h.CFilterPred(0, 1, 25, b"<")
        res = L.Filter(inter, C.byref(fp), rm.array, binds)
        want = orc.filter_scan(rels[0][1], "<", 25)
        assert res and L.b200_result_kind(res) in (1, 4)
        got = np.empty(len(want), np.uint64)
        assert L.b200_result_rowids_to_host(res, got.ctypes.data_as(h.u64p)) == 0
        assert res.contents.current_load == len(want) and np.array_equal(np.sort(got), want)
        L.FreeResult(res)
        L.FreeInterResults(inter)
    finally:
        rm.unregister()
