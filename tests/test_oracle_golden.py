"""The oracle restatement against the committed golden vectors generated from
the unmodified reference (tests/golden/make_golden.py).  Runs anywhere gcc is."""
import numpy as np
import pytest

from golden_cases import GOLDEN, join_inputs, query_relations, sha


def test_next_prime_golden(orc):
    for n, p in GOLDEN["next_prime"].items():
        assert orc.next_prime(int(n)) == p


@pytest.mark.parametrize("i", range(len(GOLDEN["joins"])))
def test_join_golden(orc, i):
    case = GOLDEN["joins"][i]
    kr, ks, pay_r, pay_s = join_inputs(case)
    r, s = orc.radix_hash_join(kr, ks, GOLDEN["n_lsb"])
    assert len(r) == case["m"]
    assert sha(r, s) == case["pairs_sha256"]            # exact order
    assert orc.checksum(pay_r, r) == case["sum_r"]
    assert orc.checksum(pay_s, s) == case["sum_s"]
    sums, m = orc.join_sum(kr, ks, [pay_r, pay_s], [0, 1], GOLDEN["n_lsb"])
    assert m == case["m"] and sums == [case["sum_r"], case["sum_s"]]


@pytest.mark.parametrize("i", range(len(GOLDEN["reorders"])))
def test_reorder_golden(orc, i):
    case = GOLDEN["reorders"][i]
    kr = join_inputs({"nr": case["n"], "ns": 1, "domain": case["domain"], "seed": case["seed"]})[0]
    ok, orid, hist, psum = orc.reorder(kr, GOLDEN["n_lsb"])
    assert [int(x) for x in hist] == case["hist"]
    assert [int(x) for x in psum] == case["psum"]
    assert sha(ok, orid) == case["tuples_sha256"]


@pytest.mark.parametrize("i", range(len(GOLDEN["queries"])))
def test_query_golden(orc, i):
    case = GOLDEN["queries"][i]
    assert orc.execute_query(case["query"], query_relations(case)) == case["line"]


def test_results_do_not_depend_on_radix_bits(orc):
    """the reference prints identical results for N_LSB 4/6/10."""
    case = GOLDEN["joins"][0]
    kr, ks, pay_r, pay_s = join_inputs(case)
    for bits in (1, 6, 10):
        sums, m = orc.join_sum(kr, ks, [pay_r, pay_s], [0, 1], bits)
        assert m == case["m"] and sums == [case["sum_r"], case["sum_s"]]


def test_filter_semantics(orc):
    col = np.array([5, 1, 7, 3, 7, 0], np.uint64)
    assert list(orc.filter_scan(col, ">", 3)) == [0, 2, 4]
    assert list(orc.filter_scan(col, "<", 3)) == [1, 5]
    assert list(orc.filter_scan(col, "=", 7)) == [2, 4]
    ids = np.array([4, 4, 0, 5], np.uint64)
    assert list(orc.filter_scan(col, ">", 3, ids)) == [0, 1, 2]     # positions, not row ids
    assert len(orc.filter_scan(col, ">", 100)) == 0


def test_synth_generator_properties(orc):
    for k in (1, 5, 16, 20):
        p = orc.synth_column(1 << k, 0, k, 0x51670D180001)
        assert len(np.unique(p)) == 1 << k and int(p.max()) == (1 << k) - 1   # bijection on k bits
    z = orc.synth_column(200000, 2, 16, 7)
    assert int(z.max()) < (1 << 16)
    _, counts = np.unique(z, return_counts=True)
    assert counts.max() > 0.04 * len(z)       # hottest key carries ~1/(k+1) of the mass


def test_column_stats_golden(orc):
    """relation_map.c:53-83 (min, max, the reference's distinct count incl. its modulo branch) restated in the oracle."""
    from golden_cases import GOLDEN, col
    assert GOLDEN["stats"]
    for case in GOLDEN["stats"]:
        for j, ((d, off), want) in enumerate(zip(case["domains"], case["stats"])):
            c = col(case["n"], d, case["seed"] * 100 + j) + np.uint64(off)
            l, u, dd = orc.column_stats(c)
            assert (l, u, float(case["n"]), float(dd)) == tuple(want), (case, j)
