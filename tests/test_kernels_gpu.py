"""Parity of the CUDA kernels (through the C-ABI) with the oracle restatement
and the committed golden vectors of the reference.  Bit-exact: everything is
u64 integer work.  Pair/partition ORDER is not compared (SURVEY §8 quirk 7:
only multisets reach the checksums)."""
import numpy as np
import pytest

from golden_cases import GOLDEN, col, join_inputs

pytestmark = pytest.mark.gpu


def sorted_pairs(r, s):
    order = np.lexsort((s, r))
    return np.stack([np.asarray(r)[order], np.asarray(s)[order]])


# ---- K1 scan_filter (filter.c:92-190) --------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 2047, 2048, 2049, 100003, 1 << 20])
@pytest.mark.parametrize("cmp", ["<", ">", "="])
def test_scan_filter_base_column(gpu, orc, n, cmp):
    c = col(n, 100, 1000 + n)
    got = gpu.scan_filter(c, cmp, 37)
    assert np.array_equal(np.sort(got), orc.filter_scan(c, cmp, 37))   # ascending row ids in the oracle


@pytest.mark.parametrize("n_ids", [0, 1, 1000, 70001])
def test_scan_filter_through_row_ids_emits_positions(gpu, orc, n_ids):
    c = col(5000, 1000, 3)
    ids = col(n_ids, 5000, 4)
    for cmp in "<>=":
        got = gpu.scan_filter(c, cmp, 500, ids)
        assert np.array_equal(np.sort(got), orc.filter_scan(c, cmp, 500, ids))


def test_scan_filter_constant_is_a_c_int(gpu, orc):
    c = np.array([0, 5, 2**31 - 1, 2**31, 2**63, 2**64 - 1], np.uint64)
    for cmp, k in [(">", 2**31 - 1), ("<", 1), ("=", 0), (">", 0)]:
        assert np.array_equal(np.sort(gpu.scan_filter(c, cmp, k)), orc.filter_scan(c, cmp, k))


# ---- K3-K5 radix partition (preprocess.c:13-178) ---------------------------
@pytest.mark.parametrize("n,bits,domain", [(0, 4, 10), (1, 4, 10), (1000, 4, 1 << 20), (8192, 6, 1 << 30),
                                           (8193, 4, 16), (300007, 10, 1 << 28), (300007, 12, 1 << 40),
                                           (1 << 20, 11, 1 << 24), (50000, 1, 7)])
def test_radix_partition_matches_oracle(gpu, orc, n, bits, domain):
    keys = col(n, domain, 77 + n)
    gk, gr, gh, gp = gpu.radix_partition(keys, bits)
    ok, orid, oh, op = orc.reorder(keys, bits)
    assert np.array_equal(gh, oh)          # histogram (HistJob + merge)
    assert np.array_equal(gp, op)          # psum incl. -1 for empty buckets
    assert np.array_equal(keys[gr.astype(np.int64)], gk)   # row ids travel with their keys
    start = 0
    for b in range(1 << bits):             # same multiset in every partition
        h = int(oh[b])
        if h:
            assert np.array_equal(np.sort(gr[start:start + h]), np.sort(orid[start:start + h])), b
        start += h


def test_radix_partition_golden_histograms(gpu):
    for case in GOLDEN["reorders"]:
        kr = join_inputs({"nr": case["n"], "ns": 1, "domain": case["domain"], "seed": case["seed"]})[0]
        _, _, gh, gp = gpu.radix_partition(kr, GOLDEN["n_lsb"])
        assert [int(x) for x in gh] == case["hist"] and [int(x) for x in gp] == case["psum"]


# ---- K6-K7 build + probe (rhjoin.c:13-111) ---------------------------------
@pytest.mark.parametrize("i", range(len(GOLDEN["joins"])))
def test_join_pairs_golden(gpu, orc, i):
    case = GOLDEN["joins"][i]
    kr, ks, pay_r, pay_s = join_inputs(case)
    r, s, m = gpu.hash_join_pairs(kr, ks)
    assert m == case["m"] == len(r)
    assert orc.checksum(pay_r, r) == case["sum_r"] and orc.checksum(pay_s, s) == case["sum_s"]
    o = orc.radix_hash_join(kr, ks, GOLDEN["n_lsb"])
    assert np.array_equal(sorted_pairs(r, s), sorted_pairs(*o))


@pytest.mark.parametrize("nr,ns,domain,bits", [
    (100000, 300000, 1 << 16, 0),      # automatic
    (100000, 300000, 1 << 16, 3),      # forced partition count, build chunks > table capacity
    (20000, 20000, 3, 0),              # massive duplicates on both sides: 3 keys
    (5000, 5000, 1, 0),                # one key: full cross product, 25M pairs
    (70000, 10, 1 << 50, 0),           # 64-bit keys, probe side tiny (sides swap)
    (1 << 18, 1 << 20, 1 << 18, 12),
])
def test_join_pairs_shapes(gpu, orc, nr, ns, domain, bits):
    kr, ks = col(nr, domain, nr + 1), col(ns, domain, ns + 2)
    gpu.lib().b200_set_tuning(bits, 0)
    try:
        r, s, m = gpu.hash_join_pairs(kr, ks)
    finally:
        gpu.lib().b200_set_tuning(0, 0)
    o_r, o_s = orc.radix_hash_join(kr, ks, 4)
    assert m == len(o_r)
    assert np.array_equal(kr[r.astype(np.int64)], ks[s.astype(np.int64)])
    if m <= 2_000_000:
        assert np.array_equal(sorted_pairs(r, s), sorted_pairs(o_r, o_s))
    else:   # order-free digest of the multiset
        mix = lambda a, b: int(((a * np.uint64(0x9E3779B97F4A7C15)) ^ (b + np.uint64(0x1234567))).sum(dtype=np.uint64))
        assert mix(r, s) == mix(o_r, o_s)


def test_join_64bit_and_32bit_key_kernels_agree(gpu, orc):
    kr, ks = col(50000, 1 << 20, 5), col(200000, 1 << 20, 6)
    want = sorted_pairs(*orc.radix_hash_join(kr, ks, 4))
    for force64 in (0, 1):
        gpu.lib().b200_set_tuning(0, force64)
        try:
            r, s, _ = gpu.hash_join_pairs(kr, ks)
        finally:
            gpu.lib().b200_set_tuning(0, 0)
        assert np.array_equal(sorted_pairs(r, s), want)


def test_join_empty_inputs(gpu):
    e, k = np.empty(0, np.uint64), np.arange(10, dtype=np.uint64)
    assert gpu.hash_join_pairs(e, k)[2] == 0 and gpu.hash_join_pairs(k, e)[2] == 0
    assert gpu.hash_join_pairs(k, k + np.uint64(100))[2] == 0


# ---- K9 checksum (inter_res.c:320-339) -------------------------------------
@pytest.mark.parametrize("m", [0, 1, 255, 256, 100001])
def test_gather_sum(gpu, orc, m):
    c = col(4096, 1 << 63, 9) * np.uint64(3)          # forces wrap-around mod 2^64
    ids = col(m, 4096, 10)
    assert gpu.gather_sum(c, ids) == orc.checksum(c, ids)


# ---- fused join -> SUM (the bench path) -------------------------------------
@pytest.mark.parametrize("i", range(len(GOLDEN["joins"])))
def test_join_sum_golden(gpu, i):
    case = GOLDEN["joins"][i]
    kr, ks, pay_r, pay_s = join_inputs(case)
    sums, m = gpu.join_sum(kr, ks, [pay_r, pay_s, pay_r], [0, 1, 0])
    assert m == case["m"] and sums == [case["sum_r"], case["sum_s"], case["sum_r"]]
    sums, m = gpu.join_sum(ks, kr, [pay_r, pay_s], [1, 0])      # sides swapped
    assert m == case["m"] and sums == [case["sum_r"], case["sum_s"]]


def test_join_sum_device_carries_registered_32bit_payload(gpu, orc):
    """A build-side SUM column registered with a maximum below 2^32 travels in the row-id slot of the build
    tuples; a wider one (or an unregistered pointer) takes the payload-array path.  Same checksums."""
    kr_bits, ks_bits = 18, 21
    nr, ns = 1 << kr_bits, 1 << ks_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 0, ks_bits, gpu.SEED_S)
    ps = orc.synth_column(ns, 1, 0, 8)
    for wide in (False, True):
        pr = orc.synth_column(nr, 1, 0, 7) * np.uint64((1 << 40) + 1 if wide else 1)
        want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
        cols = [gpu.DeviceColumn(len(a)) for a in (kr, ks, pr, ps)]
        for c, a in zip(cols, (kr, ks, pr, ps)):
            gpu.lib().b200_copy_to_device(c.ptr, a.ctypes.data, 8 * len(a))
        gpu.lib().b200_register_device_column(cols[2].ptr, cols[2].ptr, nr, int(pr.max()))
        got, m = gpu.join_sum_device(cols[0].ptr, nr, cols[1].ptr, ns, [cols[2].ptr, cols[3].ptr], [0, 1], ns - 1)
        assert m == wm and got == want
        got, m = gpu.join_sum_device(cols[1].ptr, ns, cols[0].ptr, nr, [cols[3].ptr, cols[2].ptr], [0, 1], ns - 1)
        assert m == wm and got == [want[1], want[0]]
        gpu.lib().b200_unregister_all()
        for c in cols:
            c.free()


@pytest.mark.parametrize("kr_bits,ks_bits,zipf,wide", [(18, 21, False, False), (18, 21, False, True),
                                                       (16, 21, False, False), (16, 21, True, False),
                                                       (17, 20, True, False)])
def test_join_sum_device_carries_registered_probe_payload(gpu, orc, kr_bits, ks_bits, zipf, wide):
    """A probe-side SUM column registered with a maximum below 2^32 is streamed into the row-id slot of the probe
    tuples by the histogram-free scatter when matches are not rare (build rows / key domain >= 1/24: 1/8 and, with
    Zipf keys over the build domain, 1 — there the overflow pass carries the values too); a wider column, an
    unregistered one or a sparser join (1/32) gathers per match.  Same checksums, both argument orders."""
    nr, ns = 1 << kr_bits, 1 << ks_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 2, kr_bits, 21) if zipf else orc.synth_column(ns, 0, ks_bits, gpu.SEED_S)
    max_key = nr - 1 if zipf else ns - 1
    pr = orc.synth_column(nr, 1, 0, 7)
    ps = orc.synth_column(ns, 1, 0, 8) * np.uint64((1 << 40) + 1 if wide else 1)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    cols = [gpu.DeviceColumn(len(a)) for a in (kr, ks, pr, ps)]
    try:
        for c, a in zip(cols, (kr, ks, pr, ps)):
            gpu.lib().b200_copy_to_device(c.ptr, a.ctypes.data, 8 * len(a))
        for register in (False, True):
            if register:
                gpu.lib().b200_register_device_column(cols[2].ptr, cols[2].ptr, nr, int(pr.max()))
                gpu.lib().b200_register_device_column(cols[3].ptr, cols[3].ptr, ns, int(ps.max()))
            got, m = gpu.join_sum_device(cols[0].ptr, nr, cols[1].ptr, ns, [cols[2].ptr, cols[3].ptr], [0, 1], max_key)
            assert m == wm and got == want
            got, m = gpu.join_sum_device(cols[1].ptr, ns, cols[0].ptr, nr, [cols[3].ptr, cols[2].ptr, cols[3].ptr],
                                         [0, 1, 0], max_key)          # two probe-side projections: no carrying
            assert m == wm and got == [want[1], want[0], want[1]]
            got, m = gpu.join_sum_device(cols[1].ptr, ns, cols[0].ptr, nr, [cols[3].ptr], [0], max_key)
            assert m == wm and got == [want[1]]
    finally:
        gpu.lib().b200_unregister_all()
        for c in cols:
            c.free()


def test_carried_payload_may_hold_all_ones(gpu, orc):
    """A carried SUM value of exactly 2^32 - 1 is data, not a padding marker: the join recognises padding lanes by
    their index (round-1 advisor finding: rows carrying 0xFFFFFFFF were dropped)."""
    kr_bits, ks_bits = 17, 20
    nr, ns = 1 << kr_bits, 1 << ks_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 2, kr_bits, 31)                  # every probe row matches
    pr = orc.synth_column(nr, 1, 0, 7)
    ps = orc.synth_column(ns, 1, 0, 8)
    pr[::3] = np.uint64(0xFFFFFFFF)
    ps[::2] = np.uint64(0xFFFFFFFF)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    cols = [gpu.DeviceColumn(len(a)) for a in (kr, ks, pr, ps)]
    try:
        for c, a in zip(cols, (kr, ks, pr, ps)):
            gpu.lib().b200_copy_to_device(c.ptr, a.ctypes.data, 8 * len(a))
        gpu.lib().b200_register_device_column(cols[2].ptr, cols[2].ptr, nr, 0xFFFFFFFF)
        gpu.lib().b200_register_device_column(cols[3].ptr, cols[3].ptr, ns, 0xFFFFFFFF)
        got, m = gpu.join_sum_device(cols[0].ptr, nr, cols[1].ptr, ns, [cols[2].ptr, cols[3].ptr], [0, 1], nr - 1)
        assert m == wm == ns and got == want
    finally:
        gpu.lib().b200_unregister_all()
        for c in cols:
            c.free()


@pytest.mark.parametrize("kr_bits,ks_bits", [(12, 16), (16, 20), (20, 22)])
def test_join_sum_config2_shape_scaled_down(gpu, orc, kr_bits, ks_bits):
    """BASELINE config 2 at reduced size: unique permutation keys, probe
    selectivity 2^(kr-ks), checked against the oracle's full pipeline."""
    nr, ns = 1 << kr_bits, 1 << ks_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 0, ks_bits, gpu.SEED_S)
    pr = orc.synth_column(nr, 1, 0, gpu.SEED_R + 1)
    ps = orc.synth_column(ns, 1, 0, gpu.SEED_S + 1)
    want, m = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    got, gm = gpu.join_sum(kr, ks, [pr, ps], [0, 1])
    assert m == gm == nr and got == want


# ---- staged join (the phases the multi-GPU plan drives), single GPU ----------
@pytest.mark.parametrize("kr_bits,ks_bits,zipf,carry,rank_major,carry_probe", [
    (15, 18, False, False, False, False), (18, 21, False, False, False, False), (16, 21, True, False, False, False),
    (18, 21, False, True, False, False), (18, 21, False, True, True, False), (16, 21, True, False, True, False),
    (18, 21, False, True, True, True), (16, 21, True, False, False, True), (15, 18, False, True, False, True)])
def test_staged_join_matches_fused_join(gpu, orc, kr_bits, ks_bits, zipf, carry, rank_major, carry_probe):
    """sharding.BroadcastScatterJoin with world = 1: hist -> cursors -> scatter (build side with an
    early-materialised payload, through the multi-destination path) -> join_sum, against the oracle."""
    torch = pytest.importorskip("torch")
    nr, ns = 1 << kr_bits, 1 << ks_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)
    # zipf: the histogram-free probe scatter overflows its regions -> exact second pass
    ks = orc.synth_column(ns, 2, kr_bits, 77) if zipf else orc.synth_column(ns, 0, ks_bits, gpu.SEED_S)
    pr = orc.synth_column(nr, 1, 0, gpu.SEED_R + 1)
    ps = orc.synth_column(ns, 1, 0, gpu.SEED_S + 1)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    dev = torch.device("cuda:0")
    gpu.lib().b200_set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        t = {n: torch.from_numpy(a.view(np.int64).copy()).to(dev) for n, a in
             [("kr", kr), ("ks", ks), ("pr", pr), ("ps", ps)]}
        plan = gpu.sharding.BroadcastScatterJoin(gpu, torch, None, 0, 1, nr, nr, ns, 1, dev, carry32=carry,
                                                 rank_major=rank_major, carry_probe=carry_probe)
        for _ in range(2):     # buffers are reused across steps
            got, m = plan.step(t["kr"].data_ptr(), [t["pr"].data_ptr()], t["ks"].data_ptr(),
                               [t["pr"].data_ptr(), t["ps"].data_ptr()], [0, 1])
            assert m == wm and got == want
        plan.close()
    finally:
        torch.cuda.synchronize()   # the library keeps using torch's (default) stream afterwards


# ---- BASELINE config 4 shape (scaled down): Zipf(theta = 1) probe keys over a unique build side ----
@pytest.mark.parametrize("kr_bits,ns", [(12, 200_000), (17, 3_000_000)])
def test_join_sum_zipf_probe_side(gpu, orc, kr_bits, ns):
    nr = 1 << kr_bits
    kr = orc.synth_column(nr, 0, kr_bits, gpu.SEED_R)            # every key of [0, 2^k) once
    ks = orc.synth_column(ns, 2, kr_bits, 99)                     # hottest key ~ 1/(k+1) of all probes
    pr = orc.synth_column(nr, 1, 0, 5)
    ps = orc.synth_column(ns, 1, 0, 6)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    assert wm == ns                                               # every probe matches exactly once
    got, m = gpu.join_sum(kr, ks, [pr, ps], [0, 1])
    assert m == wm and got == want
    got, m = gpu.join_sum(ks, kr, [ps, pr], [0, 1])               # skewed side first: it becomes the probe side anyway
    assert m == wm and got == [want[1], want[0]]


def test_join_skewed_build_side_duplicates(gpu, orc):
    """Zipf keys on BOTH sides: long duplicate chains in the tables and many matches per probe."""
    kr = orc.synth_column(60_000, 2, 10, 1)
    ks = orc.synth_column(80_000, 2, 10, 2)
    pr, ps = orc.synth_column(60_000, 1, 0, 3), orc.synth_column(80_000, 1, 0, 4)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    got, m = gpu.join_sum(kr, ks, [pr, ps], [0, 1])
    assert m == wm and got == want
    r, s, mp = gpu.hash_join_pairs(kr[:20000], ks[:20000])
    o_r, o_s = orc.radix_hash_join(kr[:20000], ks[:20000], 4)
    assert mp == len(o_r) and orc.checksum(pr, r) == orc.checksum(pr, o_r) and orc.checksum(ps, s) == orc.checksum(ps, o_s)


# ---- differential fuzz of the fused join -> SUM over sizes, key widths, radix bits and table capacities ----
@pytest.mark.parametrize("seed", range(10))
def test_join_sum_fuzz(gpu, orc, seed):
    g = np.random.default_rng(500 + seed)
    nr = int(g.integers(1, 400_000))
    ns = int(g.integers(1, 3_000_000))
    wide = bool(g.integers(0, 4) == 0)                                   # one in four: keys above 2^32
    domain = int(g.integers(max(2, nr // 4), 4 * nr + 4))                # <= ~4 matches per probe on average
    kr = col(nr, domain, 9000 + seed)
    ks = col(ns, domain, 9100 + seed)
    if wide:
        kr, ks = kr * np.uint64(0x100000001), ks * np.uint64(0x100000001)
    pr, ps = col(nr, 1 << 40, 9200 + seed), col(ns, 1 << 62, 9300 + seed)
    want, wm = orc.join_sum(kr, ks, [pr, ps, ps], [0, 1, 1], 4)
    bits = int(g.choice([0, 0, 3, 6, 9, 12]))                            # 0 = automatic
    gpu.lib().b200_set_tuning(bits, 0)
    try:
        got, m = gpu.join_sum(kr, ks, [pr, ps, ps], [0, 1, 1])
    finally:
        gpu.lib().b200_set_tuning(0, 0)
    assert m == wm and got == want, (nr, ns, domain, wide, bits)


def test_join_pairs_zipf_probe_side_overflows_the_regions(gpu, orc):
    """Pair-materialising join with 2^21 Zipf probe keys: the histogram-free scatter overflows and the
    probe side is partitioned again exactly; the pairs must still be the oracle's multiset."""
    kr_bits, ns = 16, 1 << 21
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 2, kr_bits, 31)
    r, s, m = gpu.hash_join_pairs(kr, ks)
    o_r, o_s = orc.radix_hash_join(kr, ks, 4)
    assert m == ns == len(o_r)
    assert np.array_equal(kr[r.astype(np.int64)], ks[s.astype(np.int64)])
    assert np.array_equal(np.sort(s), np.arange(ns, dtype=np.uint64))          # every probe row exactly once
    pr = orc.synth_column(1 << kr_bits, 1, 0, 5)
    assert orc.checksum(pr, r) == orc.checksum(pr, o_r)


# ---- rank-major (segmented) build side: several "ranks" emulated on one GPU ----------------------------------
@pytest.mark.parametrize("world,carry", [(2, False), (3, False), (4, True)])
def test_segmented_build_side_join(gpu, orc, world, carry):
    """Every emulated rank partitions its build shard into its own region of ONE build buffer (what the
    copy-engine broadcast produces on every GPU); the join then reads each partition as `world` runs."""
    import ctypes as C
    L = gpu.lib()
    kr_bits, ks_bits = 17, 20
    nr, ns = (1 << kr_bits) - 5, 1 << ks_bits
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)[:nr]
    ks = orc.synth_column(ns, 0, ks_bits, gpu.SEED_S) % np.uint64(1 << kr_bits)     # ~8 probes per build key
    pr = orc.synth_column(nr, 1, 0, 3)
    ps = orc.synth_column(ns, 1, 0, 4)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    bits = int(L.b200_radix_bits_for(nr))
    P = 1 << bits
    seg_rows = (nr + world - 1) // world
    d = {k: gpu.DeviceColumn(len(a)) for k, a in (("kr", kr), ("ks", ks), ("pr", pr), ("ps", ps))}
    for k, a in (("kr", kr), ("ks", ks), ("pr", pr), ("ps", ps)):
        L.b200_copy_to_device(d[k].ptr, a.ctypes.data, 8 * len(a))
    tup_b, pay_b, tup_p = gpu.DeviceColumn(world * seg_rows), gpu.DeviceColumn(world * seg_rows), gpu.DeviceColumn(ns)
    hist_all = gpu.DeviceColumn(world * P)           # u32[world][P] inside a u64 buffer
    hist_p, cur_p = gpu.DeviceColumn(P), gpu.DeviceColumn(P)
    for r in range(world):
        first = r * seg_rows
        cnt = min(seg_rows, nr - first)
        h_r = hist_all.ptr + 4 * r * P
        assert L.b200_stage_hist(d["kr"].ptr + 8 * first, cnt, bits, h_r) == 0
        cols = (C.c_void_p * 1)(d["pr"].ptr + 8 * first)
        outs = None if carry else (C.c_void_p * 1)(pay_b.ptr + 8 * first)
        assert L.b200_stage_scatter_build_local(d["kr"].ptr + 8 * first, cnt, first, bits, h_r, tup_b.ptr + 8 * first, 1,
                                                cols, outs) == 0
    assert L.b200_stage_hist(d["ks"].ptr, ns, bits, hist_p.ptr) == 0
    h = np.empty(P, np.uint32)
    L.b200_copy_to_host(h.ctypes.data, hist_p.ptr, 4 * P)
    cur = (np.cumsum(h, dtype=np.uint64) - h).astype(np.uint32)
    L.b200_copy_to_device(cur_p.ptr, cur.ctypes.data, 4 * P)
    assert L.b200_stage_scatter_probe(d["ks"].ptr, ns, bits, cur_p.ptr, tup_p.ptr) == 0
    pc = (C.c_void_p * 2)(d["pr"].ptr, d["ps"].ptr)
    sides = (C.c_int * 2)(0, 1)
    part = (C.c_void_p * 2)(1 if carry else pay_b.ptr, None)
    sums = (C.c_uint64 * 2)()
    m = C.c_uint64(0)
    assert L.b200_stage_join_sum_seg(tup_b.ptr, hist_all.ptr, world, seg_rows, tup_p.ptr, hist_p.ptr, bits, 2, pc, sides,
                                     part, 0, None, None, None, sums, C.byref(m)) == 0
    assert int(m.value) == wm and [int(sums[0]), int(sums[1])] == want


# ---- radix-sharded exchange (all-to-all plan): several "ranks" emulated on one GPU ----------------------------
def _to_dev(gpu, a):
    d = gpu.DeviceColumn(max(len(a), 1))
    if len(a):
        gpu.lib().b200_copy_to_device(d.ptr, a.ctypes.data, a.nbytes)
    return d


def _from_dev(gpu, ptr, n, dtype):
    out = np.empty(n, dtype)
    gpu.lib().b200_copy_to_host(out.ctypes.data, ptr, out.nbytes)
    return out


@pytest.mark.parametrize("world,bits", [(1, 3), (2, 5), (3, 6), (5, 10), (8, 12)])
def test_exchange_cursors_match_host_layout(gpu, world, bits):
    """exchange_cursors_kernel against sharding.exchange_layout on random histograms with empty partitions."""
    L = gpu.lib()
    P = 1 << bits
    g = np.random.default_rng(world * 100 + bits)
    hist = g.integers(0, 5000, size=(world, P)).astype(np.uint32)
    hist[:, g.integers(0, P, size=max(P // 8, 1))] = 0
    hist[g.integers(0, world), :] //= 3
    d_hist = _to_dev(gpu, hist.reshape(-1))
    src_off, dst_start, own, need = (gpu.DeviceColumn(P + 1) for _ in range(4))
    for rank in range(world):
        w_src, w_dst, w_own, w_need = gpu.sharding.exchange_layout(hist, rank, bits)
        for cap in (w_need, max(w_need - 1, 0)):
            assert L.b200_stage_exchange_cursors(d_hist.ptr, world, rank, bits, cap, src_off.ptr, dst_start.ptr,
                                                 own.ptr, need.ptr) == 0
            assert np.array_equal(_from_dev(gpu, src_off.ptr, P + 1, np.uint32), w_src.astype(np.uint32))
            assert np.array_equal(_from_dev(gpu, dst_start.ptr, P, np.uint32), w_dst.astype(np.uint32))
            over = w_need > cap                      # flagged: the join is given nothing to read
            assert np.array_equal(_from_dev(gpu, own.ptr, P, np.uint32), w_own.astype(np.uint32) * (0 if over else 1))
            assert _from_dev(gpu, need.ptr, 2, np.uint32).tolist() == [w_need, 1 if w_need > cap else 0]


@pytest.mark.parametrize("world,carry,npay,zipf,bits", [
    (1, True, 1, False, 0), (2, True, 1, False, 0), (3, False, 1, False, 0), (4, False, 2, True, 0),
    (8, True, 1, True, 0), (8, False, 2, False, 7), (2, False, 0, False, 0)])
def test_exchange_plan_emulated_ranks(gpu, orc, world, carry, npay, zipf, bits):
    """The whole data path of sharding.ShardedExchangeJoin with `world` ranks emulated on one GPU (the peers'
    receive buffers are ordinary local buffers): per rank hist -> cursors -> local partition pass -> exchange
    kernel; per owner the join of its partitions; the sums over owners must equal the oracle's single join."""
    import ctypes as C
    L = gpu.lib()
    kr_bits, ns = 17, (1 << 20) + 77
    nr = (1 << kr_bits) - 9
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)[:nr]
    ks = (orc.synth_column(ns, 2, kr_bits, 41) if zipf
          else orc.synth_column(ns, 0, 20, gpu.SEED_S) % np.uint64(1 << kr_bits))
    wide = np.uint64(1) if carry else np.uint64(0x10000000001)          # payload values beyond 32 bits unless carried
    pays_r = [orc.synth_column(nr, 1, 0, 3 + k) * wide for k in range(npay)]
    pays_s = [orc.synth_column(ns, 1, 0, 13 + k) * wide for k in range(npay)]
    want, wm = orc.join_sum(kr, ks, pays_r + pays_s, [0] * npay + [1] * npay, 4)
    bits = bits or int(L.b200_radix_bits_for(nr))
    P = 1 << bits
    sh = gpu.sharding
    d_kr, d_ks = _to_dev(gpu, kr), _to_dev(gpu, ks)
    d_pr, d_ps = [_to_dev(gpu, p) for p in pays_r], [_to_dev(gpu, p) for p in pays_s]
    spans = [[sh.shard_bounds(n, r, world) for r in range(world)] for n in (nr, ns)]
    hist = np.zeros((2, world, P), np.uint32)
    d_hist = gpu.DeviceColumn(2 * world * P)
    for s, d_keys in ((0, d_kr), (1, d_ks)):
        for r, (first, cnt) in enumerate(spans[s]):
            assert L.b200_stage_hist(d_keys.ptr + 8 * first, cnt, bits, d_hist.ptr + 4 * (s * world + r) * P) == 0
    hist = _from_dev(gpu, d_hist.ptr, 2 * world * P, np.uint32).reshape(2, world, P)
    cap = [max(sh.exchange_layout(hist[s], r, bits)[3] for r in range(world)) for s in (0, 1)]
    npay_x = 0 if carry else npay                                        # payload arrays that travel separately
    recv = [[gpu.DeviceColumn(max(cap[s], 1)) for _ in range(world)] for s in (0, 1)]
    recv_pay = [[[gpu.DeviceColumn(max(cap[s], 1)) for _ in range(world)] for _ in range(npay_x)] for s in (0, 1)]
    own = [[gpu.DeviceColumn(P) for _ in range(world)] for s in (0, 1)]
    src_off, dst_start, need = gpu.DeviceColumn(P + 1), gpu.DeviceColumn(P), gpu.DeviceColumn(2)
    arr = lambda ptrs: (C.c_void_p * max(len(ptrs), 1))(*ptrs)
    for s, d_keys, d_pays, n_tot in ((0, d_kr, d_pr, nr), (1, d_ks, d_ps, ns)):
        stage = gpu.DeviceColumn(max(c for _, c in spans[s]))
        stage_pay = [gpu.DeviceColumn(max(c for _, c in spans[s])) for _ in range(npay_x)]
        for r, (first, cnt) in enumerate(spans[s]):
            h_all = d_hist.ptr + 4 * s * world * P
            assert L.b200_stage_exchange_cursors(h_all, world, r, bits, cap[s], src_off.ptr, dst_start.ptr,
                                                 own[s][r].ptr, need.ptr) == 0
            assert _from_dev(gpu, need.ptr, 2, np.uint32)[1] == 0
            cols = arr([p.ptr + 8 * first for p in d_pays])
            outs = None if carry else arr([p.ptr for p in stage_pay])
            assert L.b200_stage_scatter_build_local(d_keys.ptr + 8 * first, cnt, 0, bits, h_all + 4 * r * P, stage.ptr,
                                                    npay, cols, outs) == 0
            pay_dst = arr([recv_pay[s][k][d].ptr for k in range(npay_x) for d in range(world)])
            assert L.b200_stage_exchange_segments(stage.ptr, npay_x, arr([p.ptr for p in stage_pay]), cnt, bits, world,
                                                  src_off.ptr, dst_start.ptr, cap[s], 0 if (carry or s == 0) else 1,
                                                  arr([recv[s][d].ptr for d in range(world)]), pay_dst) == 0
        L.b200_synchronize()
    sums_all, m_all = [0] * (2 * npay), 0
    for r in range(world):
        cols, sides, part = [], [], []
        for k in range(npay):
            cols.append(d_pr[k].ptr); sides.append(0)
            part.append(1 if carry else recv_pay[0][k][r].ptr)
        for k in range(npay):
            cols.append(d_ps[k].ptr if carry else recv_pay[1][k][r].ptr); sides.append(1)
            part.append(1 if carry else None)
        k2 = 2 * npay
        sums = (C.c_uint64 * max(k2, 1))()
        m = C.c_uint64(0)
        assert L.b200_stage_join_sum(recv[0][r].ptr, own[0][r].ptr, recv[1][r].ptr, own[1][r].ptr, bits, k2, arr(cols),
                                     (C.c_int * max(k2, 1))(*sides), arr(part), 0, None, None, sums, C.byref(m)) == 0
        m_all += int(m.value)
        sums_all = [(a + int(b)) % (1 << 64) for a, b in zip(sums_all, sums[:k2])]
    assert m_all == wm and sums_all == want


def test_exchange_overflow_is_flagged_not_written(gpu, orc):
    """A receive buffer that is too small: the cursors kernel raises the flag and the exchange kernel drops the
    rows beyond the capacity instead of writing past the buffer."""
    import ctypes as C
    L = gpu.lib()
    n, bits, world = 50_000, 6, 2
    P = 1 << bits
    keys = orc.synth_column(n, 0, 20, 5)
    d_keys = _to_dev(gpu, keys)
    d_hist = gpu.DeviceColumn(world * P)
    L.b200_copy_to_device(d_hist.ptr, np.zeros(world * P, np.uint32).ctypes.data, 4 * world * P)
    assert L.b200_stage_hist(d_keys.ptr, n, bits, d_hist.ptr) == 0           # rank 0 holds everything, rank 1 nothing
    hist = _from_dev(gpu, d_hist.ptr, world * P, np.uint32).reshape(world, P)
    need0 = gpu.sharding.exchange_layout(hist, 0, bits)[3]
    cap = need0 - 100
    guard = 4096
    recv = [gpu.DeviceColumn(cap + guard) for _ in range(world)]
    sentinel = np.full(cap + guard, 0xABCDABCDABCDABCD, np.uint64)
    for r in recv:
        L.b200_copy_to_device(r.ptr, sentinel.ctypes.data, sentinel.nbytes)
    src_off, dst_start, own, need = (gpu.DeviceColumn(P + 1) for _ in range(4))
    assert L.b200_stage_exchange_cursors(d_hist.ptr, world, 0, bits, cap, src_off.ptr, dst_start.ptr, own.ptr,
                                         need.ptr) == 0
    assert _from_dev(gpu, need.ptr, 2, np.uint32).tolist() == [need0, 1]
    stage = gpu.DeviceColumn(n)
    assert L.b200_stage_scatter_build_local(d_keys.ptr, n, 0, bits, d_hist.ptr, stage.ptr, 0, None, None) == 0
    tup_dst = (C.c_void_p * world)(*[r.ptr for r in recv])
    assert L.b200_stage_exchange_segments(stage.ptr, 0, None, n, bits, world, src_off.ptr, dst_start.ptr, cap, 0,
                                          tup_dst, None) == 0
    L.b200_synchronize()
    for r in recv:
        tail = _from_dev(gpu, r.ptr + 8 * cap, guard, np.uint64)
        assert np.all(tail == np.uint64(0xABCDABCDABCDABCD))


# ---- column statistics at registration (relation_map.c:53-83) ----
def test_column_stats_match_the_reference_loader(gpu, orc):
    """min / max / the reference's distinct count on the GPU against the golden vectors of the reference's own
    InitRelationMap (direct marker array, the modulo branch beyond a range of 50 000 000, a constant column) and
    against the oracle restatement on the shipped small relations."""
    from golden_cases import GOLDEN, col, load_small
    for case in GOLDEN["stats"]:
        cols = [col(case["n"], d, case["seed"] * 100 + j) + np.uint64(off) for j, (d, off) in enumerate(case["domains"])]
        rm = gpu.RelationMapArray([cols])
        assert rm.device_stats()[0] == [tuple(x) for x in case["stats"]], case
        rm.unregister()
    small = load_small()
    if small is not None:
        rm = gpu.RelationMapArray(small)
        got = rm.device_stats()
        for r, rel in enumerate(small):
            for j, c in enumerate(rel):
                l, u, d = orc.column_stats(c)
                assert got[r][j] == (l, u, float(len(c)), float(d)), (r, j)
        rm.unregister()


# ---- genuinely 64-bit keys: 16-byte tuples, histogram-free probe side, carried SUM values, verified tag table ----
@pytest.mark.parametrize("kr_bits,ks_bits,zipf", [(16, 21, False), (18, 21, False), (16, 21, True), (12, 16, False)])
def test_join_sum_64bit_keys(gpu, orc, kr_bits, ks_bits, zipf):
    """Keys beyond 2^32 (the low bits carry the radix, the high bits differ too): fused join -> SUM with registered
    32-bit SUM columns (carried in the tuples), with wide ones (gathered) and through the pair-materialising path.
    Zipf probe keys overflow the histogram-free regions: the probe side is then partitioned again, exactly."""
    nr, ns = 1 << kr_bits, 1 << ks_bits
    wide_key = lambda k: k | ((k * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFF00000000000))
    kr = wide_key(orc.synth_column(nr, 0, kr_bits, gpu.SEED_R))
    ks = wide_key(orc.synth_column(ns, 2, kr_bits, 21) if zipf else orc.synth_column(ns, 0, ks_bits, gpu.SEED_S))
    assert int(kr.max()) > (1 << 44)
    pr, ps = orc.synth_column(nr, 1, 0, 7), orc.synth_column(ns, 1, 0, 8)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    cols = [gpu.DeviceColumn(len(a)) for a in (kr, ks, pr, ps)]
    try:
        for c, a in zip(cols, (kr, ks, pr, ps)):
            gpu.lib().b200_copy_to_device(c.ptr, a.ctypes.data, 8 * len(a))
        max_key = int(max(kr.max(), ks.max()))
        for register in (False, True):
            if register:
                gpu.lib().b200_register_device_column(cols[2].ptr, cols[2].ptr, nr, int(pr.max()))
                gpu.lib().b200_register_device_column(cols[3].ptr, cols[3].ptr, ns, int(ps.max()))
            got, m = gpu.join_sum_device(cols[0].ptr, nr, cols[1].ptr, ns, [cols[2].ptr, cols[3].ptr], [0, 1], max_key)
            assert m == wm and got == want, register
            got, m = gpu.join_sum_device(cols[1].ptr, ns, cols[0].ptr, nr, [cols[3].ptr, cols[2].ptr, cols[3].ptr],
                                         [0, 1, 0], max_key)
            assert m == wm and got == [want[1], want[0], want[1]], register
    finally:
        gpu.lib().b200_unregister_all()
        for c in cols:
            c.free()
    if wm <= 3_000_000:
        r, s_, m = gpu.hash_join_pairs(kr, ks)
        assert m == wm and orc.checksum(pr, r) == want[0] and orc.checksum(ps, s_) == want[1]
        assert np.array_equal(kr[r.astype(np.int64)], ks[s_.astype(np.int64)])


def test_join_64bit_keys_that_share_their_low_32_bits(gpu, orc):
    """Keys that differ only above bit 32 land in the same partition, often the same slot with the same tag: the
    drain's key comparison is what keeps them apart."""
    n = 1 << 17
    low = orc.synth_column(n, 3, 1 << 12, 3)
    kr = low | (orc.synth_column(n, 3, 8, 4) << np.uint64(40))
    ks = low[::-1].copy() | (orc.synth_column(n, 3, 8, 5) << np.uint64(40))
    o_r, o_s = orc.radix_hash_join(kr, ks, 4)
    r, s_, m = gpu.hash_join_pairs(kr, ks)
    assert m == len(o_r) and np.array_equal(sorted_pairs(r, s_), sorted_pairs(o_r, o_s))
    p = orc.synth_column(n, 1, 0, 9)
    assert gpu.join_sum(kr, ks, [p, p], [0, 1]) == orc.join_sum(kr, ks, [p, p], [0, 1], 4)


# ---- repeatability: the same join many times over (what a race in the tile pipeline, the shared-memory CAS chains
#      or the warp queues would disturb; compute-sanitizer's racecheck is not available on the GPU pool) ----
@pytest.mark.parametrize("wide", [False, True])
def test_join_sum_is_repeatable(gpu, orc, wide):
    nr, ns = 1 << 18, (1 << 22) + 12345
    kr = orc.synth_column(nr, 0, 18, gpu.SEED_R)
    ks = orc.synth_column(ns, 3, 1 << 19, gpu.SEED_S)      # every build key is probed ~8 times, half the probes miss
    if wide:
        spread = lambda k: k | ((k * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFF00000000000))
        kr, ks = spread(kr), spread(ks)
    pr, ps = orc.synth_column(nr, 1, 0, 7), orc.synth_column(ns, 1, 0, 8)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    cols = [gpu.DeviceColumn(len(a)) for a in (kr, ks, pr, ps)]
    try:
        for c, a in zip(cols, (kr, ks, pr, ps)):
            gpu.lib().b200_copy_to_device(c.ptr, a.ctypes.data, 8 * len(a))
        max_key = int(max(kr.max(), ks.max()))
        for register in (False, True):      # gathered SUM columns, then carried in the tuples
            if register:
                gpu.lib().b200_register_device_column(cols[2].ptr, cols[2].ptr, nr, int(pr.max()))
                gpu.lib().b200_register_device_column(cols[3].ptr, cols[3].ptr, ns, int(ps.max()))
            for rep in range(20):
                got, m = gpu.join_sum_device(cols[0].ptr, nr, cols[1].ptr, ns, [cols[2].ptr, cols[3].ptr], [0, 1], max_key)
                assert (got, m) == (want, wm), (register, rep)
    finally:
        gpu.lib().b200_unregister_all()
        for c in cols:
            c.free()
