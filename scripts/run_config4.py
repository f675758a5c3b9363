#!/usr/bin/env python
"""BASELINE config 4: Zipf(theta = 1) probe side x unique build side, radix-sharded across the GPUs of one box
with an all-to-all over NVLink (sharding.ShardedExchangeJoin).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/run_config4.py [--build-bits 27] [--probe-rows 2000000000] [--steps 10] [--warmup 3]

Full size is --build-bits 27 --probe-rows 2000000000 on 8 GPUs; on fewer GPUs scale both by N/8 to keep the
per-GPU load (e.g. N = 2: --build-bits 25 --probe-rows 500000000).  Rows start position-sharded; keys and
payloads come from the generator shared with the CPU oracle (include/b200_synth.h).  The expected checksums are
computed independently of the join kernels with torch (a dense lookup table over the build keys, which are a
permutation of [0, 2^k)): every probe row matches exactly once.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))


def _ck(L, rc):
    """Raise when a C-ABI call failed (never `assert call(...) == 0`: python -O strips asserts and with them the call)."""
    if rc != 0:
        raise RuntimeError("libb200join: " + (L.b200_last_error() or b"error").decode())


def main():
    # ONE JSON line on stdout: libraries (NCCL prints its version there) get stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-bits", type=int, default=27)
    ap.add_argument("--probe-rows", type=int, default=2_000_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--uniform", action="store_true", help="uniform probe keys instead of Zipf (control)")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    from conftest import load_package

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b200 = load_package()
    L = b200.lib()
    _ck(L, L.b200_init(local))
    stream = torch.cuda.current_stream()
    L.b200_set_stream(stream.cuda_stream)
    sh = b200.sharding

    k = args.build_bits
    nr, ns = 1 << k, args.probe_rows
    r_first, nr_loc = sh.shard_bounds(nr, rank, world)
    s_first, ns_loc = sh.shard_bounds(ns, rank, world)

    def synth(n, first, kind, kk, seed):
        t = torch.empty(n, dtype=torch.int64, device=dev)
        b200.synth_column_device(t.data_ptr(), first, n, kind, kk, seed)
        return t

    r0 = synth(nr_loc, r_first, b200.SYNTH_PERM, k, b200.SEED_R)
    r1 = synth(nr_loc, r_first, b200.SYNTH_PAYLOAD, 0, b200.SEED_R + 1)
    if args.uniform:
        s0 = synth(ns_loc, s_first, b200.SYNTH_UNIFORM, nr, b200.SEED_S)
    else:
        s0 = synth(ns_loc, s_first, b200.SYNTH_ZIPF, k, b200.SEED_S)
    s1 = synth(ns_loc, s_first, b200.SYNTH_PAYLOAD, 0, b200.SEED_S + 1)

    # expected checksums without the join kernels: lut[key] = R.c1 of the row holding that key
    if world > 1:
        r0_all = sh.allgather_column(r0, nr, dist)
        r1_all = sh.allgather_column(r1, nr, dist)
    else:
        r0_all, r1_all = r0, r1
    lut = torch.empty(nr, dtype=torch.int64, device=dev)
    lut[r0_all] = r1_all
    want = torch.zeros(3, dtype=torch.int64, device=dev)
    chunk = 1 << 26
    for a in range(0, ns_loc, chunk):
        want[0] += lut[s0[a:a + chunk]].sum()
    want[1] = s1.sum()
    want[2] = ns_loc
    hottest = int(torch.bincount(s0[: min(ns_loc, 1 << 24)] & 0xFFFF, minlength=1 << 16).max().item())
    del lut, r0_all, r1_all
    if world > 1:
        dist.all_reduce(want)
    want = [int(x) for x in want.cpu().numpy().view(np.uint64)]

    plan = sh.ShardedExchangeJoin(b200, torch, dist if world > 1 else None, rank, world, nr, nr_loc, ns_loc, 1, 1, dev,
                                  size_from=(r0.data_ptr(), s0.data_ptr()), carry_build=True, carry_probe=True)
    torch.cuda.synchronize()

    def step():
        return plan.step(r0.data_ptr(), [r1.data_ptr()], s0.data_ptr(), [s1.data_ptr()])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sums, m = step()
    if [sums[0], sums[1], m] != want:
        raise SystemExit(f"checksum mismatch: got {sums} m={m}, want {want}")
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        sums, m = step()
    ev1.record(stream)
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if [sums[0], sums[1], m] != want:
        raise SystemExit(f"checksum mismatch: got {sums} m={m}, want {want}")

    L.b200_set_profiling(1)
    per_kernel = {}
    for _ in range(3):
        step()
        torch.cuda.synchronize()
        for name in ("scatter_b", "exchange", "join"):
            v = b200.last_kernel_ms(name)
            if v >= 0:
                per_kernel.setdefault(name, []).append(round(v, 4))
    L.b200_set_profiling(0)
    need = plan.need.cpu().tolist()
    recv = torch.tensor([need[0][0], need[1][0]], dtype=torch.int64, device=dev)
    recv_max = recv.clone()
    if world > 1:
        dist.all_reduce(recv_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        wire = 8 * (nr_loc + ns_loc) * (world - 1) / max(world, 1)      # expected bytes leaving this rank per step
        print(json.dumps({
            "workload": f"config4: Zipf(1.0) probe {ns} rows x unique build 2^{k}, radix-sharded x{world}"
                        if not args.uniform else f"control: uniform probe {ns} rows x unique build 2^{k}, x{world}",
            "n_gpus": world, "ms_per_step": ms, "probe_tuples_per_s": ns / (ms * 1e-3), "steps": args.steps,
            "radix_bits": plan.bits, "checksums": sums, "matches": m, "checksum_ok": True,
            "rows_received_max_rank": [int(x) for x in recv_max.cpu().tolist()],
            "rows_received_mean": [nr / world, ns / world],
            "nvlink_out_bytes_per_rank": wire, "nvlink_out_gbs_per_rank": wire / (ms * 1e-3) / 1e9,
            "hottest_16bit_key_share_sample": hottest / min(ns_loc, 1 << 24),
            "last_kernel_ms_rank0": per_kernel,
        }))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
