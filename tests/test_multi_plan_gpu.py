"""The C-ABI multi-GPU plans (csrc/multi.cu, b200_multi_*) on ONE GPU: several ranks live in this process, their
"peer" memory is plain device memory, and the step is driven phase by phase over all ranks (kernels that wait on one
another must never share a GPU, so phase i of every rank completes before phase i + 1 of any rank starts).  Same
kernels, flags and layouts as on N GPUs; tests/multi_gpu_worker.py runs them over real CUDA IPC + NVLink.
Everything is compared bit-exactly with the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(gpu, orc, plan, world, kr, ks, pr, ps, has_b=True, has_p=True, reps=2, **kw):
    sh = gpu.sharding
    L = gpu.lib()
    nr, ns = len(kr), len(ks)
    proj = ([pr] if has_b else []) + ([ps] if has_p else [])
    sides = ([0] if has_b else []) + ([1] if has_p else [])
    want, wm = orc.join_sum(kr, ks, proj, sides, 4)
    bounds_r = [sh.shard_bounds(nr, r, world) for r in range(world)]
    bounds_s = [sh.shard_bounds(ns, r, world) for r in range(world)]
    counts = [(bounds_r[r][1], bounds_s[r][1]) for r in range(world)]
    cols, plans = [], []
    try:
        for r in range(world):
            (fr, cr), (fs, cs) = bounds_r[r], bounds_s[r]
            dev = []
            for a in (kr[fr:fr + cr], pr[fr:fr + cr], ks[fs:fs + cs], ps[fs:fs + cs]):
                d = gpu.DeviceColumn(max(len(a), 1))
                if len(a):
                    L.b200_copy_to_device(d.ptr, np.ascontiguousarray(a).ctypes.data, 8 * len(a))
                dev.append(d)
            cols.append(dev)
            plans.append(sh.MultiJoin(gpu, None, r, world, 0, plan, cr, cs, has_b, has_p,
                                      peers_in_process={"counts": counts}, **kw))
        for p in plans:
            p.connect_in_process(plans, [0] * world)
        phases = (1, 2, 4) if plan == gpu.PLAN_BROADCAST else (1, 2, 4, 8, 16, 32)
        for _ in range(reps):          # buffers, flags and epochs are reused across steps
            for ph in phases:
                for r in range(world):
                    d = cols[r]
                    plans[r].enqueue(d[0].ptr, d[1].ptr if has_b else None, d[2].ptr, d[3].ptr if has_p else None, ph)
                L.b200_synchronize()
            for r in range(world):
                got, m = plans[r].finish()
                assert (got, m) == (want, wm), (r, got, m, want, wm)
        return plans, [p.received() for p in plans] if plan == gpu.PLAN_EXCHANGE else None
    finally:
        recv = None
        for p in plans:
            p.close()
        for dev in cols:
            for d in dev:
                d.free()


@pytest.mark.parametrize("world,kr_bits,ns,chunks", [(1, 16, 1 << 21, 0), (2, 18, (1 << 22) + 77, 4), (3, 17, 3 << 20, 3),
                                                    (2, 15, 70_001, 2), (4, 16, (1 << 22) + 5, 8)])
def test_broadcast_plan_emulated_ranks(gpu, orc, world, kr_bits, ns, chunks):
    nr = (1 << kr_bits) - 3                                    # ragged build shards
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)[:nr]
    ks = orc.synth_column(ns, 0, kr_bits + 3, gpu.SEED_S)      # 1/8 of the probe rows match
    pr, ps = orc.synth_column(nr, 1, 0, 3), orc.synth_column(ns, 1, 0, 4)
    _run(gpu, orc, gpu.PLAN_BROADCAST, world, kr, ks, pr, ps, chunks=chunks)


@pytest.mark.parametrize("world,chunks", [(2, 0), (3, 2), (4, 8)])
def test_broadcast_plan_pull_variant(gpu, orc, world, chunks, monkeypatch):
    """B200_BCAST=pull: every rank fetches the peers' regions with a kernel (loads over NVLink) that raises the chunk
    flags the join waits on, instead of pushing its own region with the copy engines."""
    monkeypatch.setenv("B200_BCAST", "pull")
    monkeypatch.setenv("B200_BCAST_SMS", "4")
    kr = orc.synth_column(1 << 17, 0, 17, gpu.SEED_R)[: (1 << 17) - 5]
    ks = orc.synth_column((1 << 22) + 3, 0, 20, gpu.SEED_S)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(len(ks), 1, 0, 4)
    _run(gpu, orc, gpu.PLAN_BROADCAST, world, kr, ks, pr, ps, chunks=chunks)


def test_broadcast_plan_skewed_probe_keys_take_the_overflow_pass(gpu, orc):
    """Zipf probe keys overflow the histogram-free regions: finish() redoes the join with the exact overflow pass
    (world = 1: with several ranks on one GPU the second result exchange could not be driven phase by phase)."""
    kr_bits, ns = 16, 1 << 21
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 2, kr_bits, 77)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(ns, 1, 0, 4)
    _run(gpu, orc, gpu.PLAN_BROADCAST, 1, kr, ks, pr, ps)


@pytest.mark.parametrize("has_b,has_p", [(True, False), (False, True), (False, False)])
def test_broadcast_plan_projection_subsets(gpu, orc, has_b, has_p):
    kr = orc.synth_column(1 << 16, 0, 16, gpu.SEED_R)
    ks = orc.synth_column(1 << 21, 0, 19, gpu.SEED_S)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(len(ks), 1, 0, 4)
    _run(gpu, orc, gpu.PLAN_BROADCAST, 2, kr, ks, pr, ps, has_b, has_p, reps=1)


@pytest.mark.parametrize("hot", [True, False])
@pytest.mark.parametrize("world,kr_bits,ns,zipf,chunks,bits", [(1, 16, 1 << 20, True, 0, 0), (2, 17, (1 << 21) + 5, True, 4, 0),
                                                              (3, 16, (1 << 20) + 77, False, 2, 0), (4, 15, 9, True, 0, 0),
                                                              (2, 17, 1 << 20, False, 8, 8), (8, 16, 1 << 21, True, 2, 0),
                                                              (2, 17, (1 << 22) + 70, True, 2, 12), (1, 16, (1 << 21) + 6, False, 2, 11)])
def test_exchange_plan_emulated_ranks(gpu, orc, world, kr_bits, ns, zipf, chunks, bits, hot, monkeypatch):
    """hot: keys a large share of a sample of the probe rows carries are joined where they are, during the histogram
    pass (Zipf inputs have them, the uniform ones do not); off: everything is exchanged.  The last two cases have
    2^11 / 2^12 partitions and chunks of >= 2^20 rows: their probe chunks take the two-pass partition."""
    if bits >= 11:
        monkeypatch.setenv("B200_TWO_PASS_BITS", "11")      # (off by default)
    nr = (1 << kr_bits) - 9
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)[:nr]
    ks = orc.synth_column(ns, 2, kr_bits, 41) if zipf else orc.synth_column(ns, 0, kr_bits + 2, gpu.SEED_S)
    pr, ps = orc.synth_column(nr, 1, 0, 3), orc.synth_column(ns, 1, 0, 13)
    _run(gpu, orc, gpu.PLAN_EXCHANGE, world, kr, ks, pr, ps, chunks=chunks, radix_bits=bits,
         recv_rows_build=nr, recv_rows_probe=ns, hot_keys=hot)


def test_exchange_plan_hot_keys_with_duplicate_build_keys(gpu, orc):
    """A hot key that several build rows carry: every probe row with it matches all of them (count and SUM of the
    build rows are replicated, the probe value is multiplied by the count)."""
    ns = 1 << 21
    kr = orc.synth_column(1 << 16, 3, 1 << 14, 5)                 # 2^16 build rows over 2^14 keys: ~4 rows per key
    ks = orc.synth_column(ns, 2, 14, 41)                          # Zipf over the same 2^14 keys (perm14)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(ns, 1, 0, 13)
    for has_b, has_p in ((True, True), (False, True), (True, False)):
        _run(gpu, orc, gpu.PLAN_EXCHANGE, 2, kr, ks, pr, ps, has_b, has_p, reps=1, recv_rows_build=len(kr), recv_rows_probe=ns)


def test_exchange_plan_balances_owners_under_zipf(gpu, orc):
    """Zipf(1.0) probe keys: the hottest key alone is 1/k of the probe rows.  Ownership cuts placed on the global
    histogram keep the most loaded owner within 10 % of the mean (equal-width ranges would not)."""
    world, kr_bits, ns = 4, 16, 1 << 22
    kr = orc.synth_column(1 << kr_bits, 0, kr_bits, gpu.SEED_R)
    ks = orc.synth_column(ns, 2, kr_bits, 99)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(ns, 1, 0, 13)
    _, recv = _run(gpu, orc, gpu.PLAN_EXCHANGE, world, kr, ks, pr, ps, reps=1, recv_rows_build=len(kr), recv_rows_probe=ns,
                   radix_bits=10)       # 1024 partitions (a 2^16-row build side alone would get 16: too coarse to cut)
    rows = [b + p for b, p in recv]
    # (probe rows with a hot key are joined where they are and never exchanged: fewer than ns arrive)
    assert sum(p for _, p in recv) < ns and sum(b for b, _ in recv) == len(kr)
    assert max(rows) <= 1.10 * (sum(rows) / world), rows


def test_exchange_plan_reports_too_small_receive_buffers(gpu, orc):
    kr = orc.synth_column(1 << 15, 0, 15, gpu.SEED_R)
    ks = orc.synth_column(1 << 19, 2, 15, 5)
    pr, ps = orc.synth_column(len(kr), 1, 0, 3), orc.synth_column(len(ks), 1, 0, 13)
    with pytest.raises(RuntimeError, match="too small"):
        _run(gpu, orc, gpu.PLAN_EXCHANGE, 2, kr, ks, pr, ps, reps=1, recv_rows_build=len(kr), recv_rows_probe=len(ks) // 8)
