"""The N > 1 host logic (sigmod-2018_b200/sharding.py) on CPU: two gloo ranks,
position shards from the shared generator, all-gather of the build side, the
oracle as the local join, u64 all-reduce of the checksums; the result must
equal the single-process oracle."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kr_bits, ks_bits, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_package
    import orc
    b200 = load_package()
    sh = b200.sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nr, ns = (1 << kr_bits) - 3, (1 << ks_bits) - 5     # ragged shards
    r_first, r_cnt = sh.shard_bounds(nr, rank, world)
    s_first, s_cnt = sh.shard_bounds(ns, rank, world)
    r0 = orc.synth_column(r_cnt, 0, kr_bits, b200.SEED_R, first=r_first)
    r1 = orc.synth_column(r_cnt, 1, 0, b200.SEED_R + 1, first=r_first) * np.uint64(0x1000000000001)  # wraps
    s0 = orc.synth_column(s_cnt, 0, ks_bits, b200.SEED_S, first=s_first)
    s1 = orc.synth_column(s_cnt, 1, 0, b200.SEED_S + 1, first=s_first) * np.uint64(0x1000000000003)
    as_t = lambda a: torch.from_numpy(a.view(np.int64).copy())
    r0_all = sh.allgather_column(as_t(r0), nr, dist).numpy().view(np.uint64)
    r1_all = sh.allgather_column(as_t(r1), nr, dist).numpy().view(np.uint64)
    sums, m = orc.join_sum(r0_all, s0, [r1_all, s1], [0, 1], 4)
    sums, m = sh.allreduce_checksums(sums, m, dist)
    if rank == 0:
        q.put((sums, m))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_broadcast_plan_matches_single_process(orc, b200, world):
    import torch  # noqa: F401  (imported before the fork so the children do not pay for it again)
    import torch.multiprocessing as mp
    kr_bits, ks_bits = 10, 14
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kr_bits, ks_bits, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nr, ns = (1 << kr_bits) - 3, (1 << ks_bits) - 5
    r0 = orc.synth_column(nr, 0, kr_bits, b200.SEED_R)
    r1 = orc.synth_column(nr, 1, 0, b200.SEED_R + 1) * np.uint64(0x1000000000001)
    s0 = orc.synth_column(ns, 0, ks_bits, b200.SEED_S)
    s1 = orc.synth_column(ns, 1, 0, b200.SEED_S + 1) * np.uint64(0x1000000000003)
    want = orc.join_sum(r0, s0, [r1, s1], [0, 1], 4)
    assert got == (want[0], want[1])


def test_shard_bounds_cover_everything(b200):
    sh = b200.sharding
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            spans = [sh.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_u64_sums_survive_int64_allreduce(b200):
    sh = b200.sharding
    vals = [0, 1, 2**63 - 1, 2**63, 2**64 - 1]
    assert sh.i64_to_u64(sh.u64_to_i64(vals)) == vals
    a, b = 2**64 - 5, 17
    s = (np.int64(sh.u64_to_i64([a])[0]) + np.int64(sh.u64_to_i64([b])[0]))
    assert sh.i64_to_u64([s])[0] == (a + b) % 2**64


# ---- radix-sharded exchange plan (sharding.exchange_layout / partition_owner) under gloo ----------------------
def _exchange_worker(rank, world, port, kr_bits, ns, bits, q):
    """Every rank partitions its shards, sends each partition segment to the owner at the position
    exchange_layout assigns, the owner joins what it received with the oracle; checksums are all-reduced."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_package
    import orc
    b200 = load_package()
    sh = b200.sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = 1 << bits
    nr = (1 << kr_bits) - 3
    received = []
    for n, kind, k, seed in ((nr, 0, kr_bits, b200.SEED_R), (ns, 2, kr_bits, 77)):      # unique build, Zipf probe
        first, cnt = sh.shard_bounds(n, rank, world)
        keys = orc.synth_column(cnt, kind, k, seed, first=first)
        pay = orc.synth_column(cnt, 1, 0, seed + 1, first=first) * np.uint64(0x1000000000001)
        part = (keys & np.uint64(P - 1)).astype(np.int64)
        hist = np.bincount(part, minlength=P).astype(np.int64)
        hist_all = [torch.zeros(P, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(hist_all, torch.from_numpy(hist))
        hist_all = np.stack([h.numpy() for h in hist_all])
        src_off, dst_start, own_total, need = sh.exchange_layout(hist_all, rank, bits)
        order = np.argsort(part, kind="stable")                   # the local partition pass
        assert np.array_equal(src_off[:-1], np.cumsum(hist) - hist)
        p_sorted = part[order]
        pos = dst_start[p_sorted].astype(np.int64) + (np.arange(cnt) - src_off[p_sorted].astype(np.int64))
        owner = sh.partition_owner(p_sorted, world, bits).astype(np.int64)
        assert np.all(np.diff(owner) >= 0)                        # owners are contiguous in partition order
        rows = np.stack([pos.astype(np.uint64), keys[order], pay[order]], axis=1)        # [cnt, 3] u64
        send_counts = np.bincount(owner, minlength=world)
        recv_counts = [int(hist_all[s][sh.partition_owner(np.arange(P), world, bits) == rank].sum())
                       for s in range(world)]
        assert sum(recv_counts) == need
        out = torch.zeros((need, 3), dtype=torch.int64)
        dist.all_to_all_single(out, torch.from_numpy(rows.view(np.int64).copy()),
                               output_split_sizes=recv_counts, input_split_sizes=send_counts.tolist())
        got = out.numpy().view(np.uint64)
        buf_k, buf_v = np.zeros(need, np.uint64), np.zeros(need, np.uint64)
        assert np.array_equal(np.sort(got[:, 0]), np.arange(need, dtype=np.uint64))      # every slot exactly once
        buf_k[got[:, 0].astype(np.int64)] = got[:, 1]
        buf_v[got[:, 0].astype(np.int64)] = got[:, 2]
        # the receive buffer is partition-major: partition ids are non-decreasing and match own_total
        pid = (buf_k & np.uint64(P - 1)).astype(np.int64)
        assert np.all(np.diff(pid) >= 0)
        assert np.array_equal(np.bincount(pid, minlength=P).astype(np.uint64), own_total)
        received.append((buf_k, buf_v))
    (bk, bv), (pk, pv) = received
    sums, m = orc.join_sum(bk, pk, [bv, pv], [0, 1], 4) if len(bk) and len(pk) else ([0, 0], 0)
    sums, m = sh.allreduce_checksums(sums, m, dist)
    if rank == 0:
        q.put((sums, m))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bits", [(2, 6), (3, 5)])
def test_exchange_plan_matches_single_process(orc, b200, world, bits):
    import torch  # noqa: F401
    import torch.multiprocessing as mp
    kr_bits, ns = 10, 20_011
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, kr_bits, ns, bits, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nr = (1 << kr_bits) - 3
    r0 = orc.synth_column(nr, 0, kr_bits, b200.SEED_R)
    r1 = orc.synth_column(nr, 1, 0, b200.SEED_R + 1) * np.uint64(0x1000000000001)
    s0 = orc.synth_column(ns, 2, kr_bits, 77)
    s1 = orc.synth_column(ns, 1, 0, 78) * np.uint64(0x1000000000001)
    want = orc.join_sum(r0, s0, [r1, s1], [0, 1], 4)
    assert got == (want[0], want[1])


def test_partition_owner_ranges(b200):
    sh = b200.sharding
    for world in (1, 2, 3, 5, 8):
        for bits in (3, 6, 12):
            owner = sh.partition_owner(np.arange(1 << bits), world, bits)
            assert owner[0] == 0 and owner[-1] == world - 1 and np.all(np.diff(owner.astype(np.int64)) >= 0)
            counts = np.bincount(owner.astype(np.int64), minlength=world)
            assert counts.min() >= (1 << bits) // world and counts.max() <= -(-(1 << bits) // world)
