"""Multi-GPU parity worker (one process per GPU, launched by tests/test_multi_gpu.py through torchrun):
sharding.ShardedExchangeJoin and sharding.BroadcastScatterJoin over real CUDA-IPC peer buffers and NCCL, against
the CPU oracle on the same synthetic relations.  Prints "MULTI-GPU PARITY OK <cases>" on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_package
    import orc

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b200 = load_package()
    L = b200.lib()
    assert L.b200_init(local) == 0
    L.b200_set_stream(torch.cuda.current_stream().cuda_stream)
    sh = b200.sharding

    def shard(a, n):
        first, cnt = sh.shard_bounds(n, rank, world)
        return torch.from_numpy(a[first:first + cnt].view(np.int64).copy()).to(dev)

    cases = 0
    # ---- exchange plan: (build rows, probe rows, carried?, payload columns per side, Zipf probe keys?, radix bits)
    for nr, ns, carry, npay, zipf, bits in [((1 << 17) - 9, (1 << 20) + 77, True, 1, False, None),
                                            (1 << 16, (1 << 21) + 5, False, 2, True, None),
                                            ((1 << 17) - 1, 1 << 20, False, 1, False, 8),
                                            (1 << 15, 3 * world + 1, True, 1, True, None)]:
        k = int(np.ceil(np.log2(nr + 1)))
        kr = orc.synth_column(1 << k, 0, k, b200.SEED_R)[:nr]
        ks = orc.synth_column(ns, 2, k, 41) if zipf else orc.synth_column(ns, 0, 22, b200.SEED_S) % np.uint64(1 << k)
        wide = np.uint64(1) if carry else np.uint64(0x10000000001)
        pr = [orc.synth_column(nr, 1, 0, 3 + j) * wide for j in range(npay)]
        ps = [orc.synth_column(ns, 1, 0, 13 + j) * wide for j in range(npay)]
        want, wm = orc.join_sum(kr, ks, pr + ps, [0] * npay + [1] * npay, 4)
        d_kr, d_ks = shard(kr, nr), shard(ks, ns)
        d_pr, d_ps = [shard(p, nr) for p in pr], [shard(p, ns) for p in ps]
        plan = sh.ShardedExchangeJoin(b200, torch, dist, rank, world, nr, d_kr.numel(), d_ks.numel(), npay, npay, dev,
                                      size_from=(d_kr.data_ptr(), d_ks.data_ptr()), carry_build=carry,
                                      carry_probe=carry, bits=bits)
        for _ in range(2):
            got, m = plan.step(d_kr.data_ptr(), [p.data_ptr() for p in d_pr], d_ks.data_ptr(),
                               [p.data_ptr() for p in d_ps])
            assert m == wm and got == want, ("exchange", nr, ns, carry, npay, zipf, got, m, want, wm)
        plan.close()
        cases += 1
    # a receive buffer that is too small must fail the step on every rank, not corrupt memory
    nr, ns = 1 << 16, 1 << 20
    kr, ks = orc.synth_column(nr, 0, 16, b200.SEED_R), orc.synth_column(ns, 2, 16, 5)
    d_kr, d_ks = shard(kr, nr), shard(ks, ns)
    plan = sh.ShardedExchangeJoin(b200, torch, dist, rank, world, nr, d_kr.numel(), d_ks.numel(), 0, 0, dev,
                                  recv_rows_build=nr, recv_rows_probe=ns // world // 2)
    try:
        plan.step(d_kr.data_ptr(), [], d_ks.data_ptr(), [])
        raise SystemExit("a too-small receive buffer was not reported")
    except RuntimeError as e:
        assert "too small" in str(e)
    plan.close()
    cases += 1
    # ---- broadcast plans (equal shards) ----
    nr, ns = 1 << 18, 1 << 22
    kr, ks = orc.synth_column(nr, 0, 18, b200.SEED_R), orc.synth_column(ns, 0, 22, b200.SEED_S)
    pr, ps = orc.synth_column(nr, 1, 0, 3), orc.synth_column(ns, 1, 0, 4)
    want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
    d_kr, d_ks, d_pr, d_ps = shard(kr, nr), shard(ks, ns), shard(pr, nr), shard(ps, ns)
    for carry, rank_major, carry_probe in ((False, False, False), (True, False, False), (True, True, False),
                                           (True, True, True), (False, False, True)):
        plan = sh.BroadcastScatterJoin(b200, torch, dist, rank, world, nr, d_kr.numel(), d_ks.numel(), 1, dev,
                                       carry32=carry, rank_major=rank_major, carry_probe=carry_probe)
        for _ in range(2):
            got, m = plan.step(d_kr.data_ptr(), [d_pr.data_ptr()], d_ks.data_ptr(),
                               [d_pr.data_ptr(), d_ps.data_ptr()], [0, 1])
            assert m == wm and got == want, ("broadcast", carry, rank_major, carry_probe, got, m, want, wm)
        plan.close()
        cases += 1
    # ---- the C-ABI plans (csrc/multi.cu): no NCCL in the step, flags in peer memory over CUDA IPC ----
    for plan_kind, nr, ns, zipf in [(b200.PLAN_BROADCAST, (1 << 18) - 5, (1 << 22) + 9, False),
                                    (b200.PLAN_BROADCAST, 1 << 16, 1 << 21, True),      # overflow pass + 2nd result round
                                    (b200.PLAN_EXCHANGE, (1 << 17) - 9, (1 << 21) + 77, True),
                                    (b200.PLAN_EXCHANGE, 1 << 16, 3 * world + 1, False)]:
        k = int(np.ceil(np.log2(nr + 1)))
        kr = orc.synth_column(1 << k, 0, k, b200.SEED_R)[:nr]
        ks = orc.synth_column(ns, 2, k, 41) if zipf else orc.synth_column(ns, 0, k + 2, b200.SEED_S)
        pr, ps = orc.synth_column(nr, 1, 0, 3), orc.synth_column(ns, 1, 0, 13)
        want, wm = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
        d_kr, d_ks, d_pr, d_ps = shard(kr, nr), shard(ks, ns), shard(pr, nr), shard(ps, ns)
        for graph, bcast in ((0, "push"), (1, "push"), (0, "pull"), (1, "pull")):
            if bcast == "pull" and plan_kind != b200.PLAN_BROADCAST:
                continue
            os.environ["B200_MULTI_GRAPH"] = str(graph)
            os.environ["B200_BCAST"] = bcast
            plan = sh.MultiJoin(b200, dist, rank, world, local, plan_kind, d_kr.numel(), d_ks.numel(),
                                recv_rows_build=nr, recv_rows_probe=ns)
            for _ in range(3):
                got, m = plan.step(d_kr.data_ptr(), d_pr.data_ptr(), d_ks.data_ptr(), d_ps.data_ptr())
                assert m == wm and got == want, ("multi", plan_kind, graph, nr, ns, zipf, got, m, want, wm)
            torch.cuda.synchronize()
            dist.barrier()
            plan.close()
        cases += 1
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"MULTI-GPU PARITY OK {cases}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
