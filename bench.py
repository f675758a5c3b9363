#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload: synthetic 2-relation uint64 equi-join,
|R| = 2^24 build x |S| = 2^28 probe, unique permutation keys, query
`0 1|0.0=1.0|0.1 1.1` (SUM of one payload column per side).  One "step" is one
pass of the whole hot path over that input: radix histogram + scatter of both
sides, per-partition shared-memory build + probe, SUM projection.

  value   probe tuples/s, inputs resident in HBM (CUDA events on the stream
          the kernels are launched on, max over ranks)
  e2e     same metric through the public C-ABI call with HOST (pinned)
          buffers: H2D copies of all four columns and the D2H of the sums are
          inside the timed region
  roofline  dominant kernel (per-kernel CUDA-event time measured live) against
          MEASURED_PEAKS.json; canonical algorithmic bytes and
          the bytes of the narrower encoding actually used are both given
  cpu_baseline  the reference's own CPU join (oracle/_ref/ref_driver, built
          from the unmodified reference) on a bounded sample, host cores stated

N > 1 (torchrun, one rank per GPU): R and S start position-sharded; every rank
scatters its build shard into ALL ranks' partition buffers with P2P stores over
NVLink (the broadcast of the small side fused into the scatter kernel, SURVEY
§8e), partitions its probe shard locally, joins, and the checksums are
all-reduced (sharding.BroadcastScatterJoin).  Strong scaling.

`--impl reference` times the reference's CPU implementation (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
KR_BITS, KS_BITS = 24, 28
QUERY = "0 1|0.0=1.0|0.1 1.1"
METRIC = "join_probe_throughput"
UNIT = "probe tuples/s"
# canonical algorithmic bytes, SURVEY §8d: 40 B per input row + 16 B per match (k = 2 projections)
CANON_PER_INPUT, CANON_PER_MATCH = 40, 16


def _ck(L, rc):
    """Raise when a C-ABI call failed (never `assert call(...) == 0`: python -O strips asserts and with them the call)."""
    if rc != 0:
        raise RuntimeError("libb200join: " + (L.b200_last_error() or b"error").decode())


def load_package():
    name = "sigmod2018_b200"
    pkg_dir = ROOT / "sigmod-2018_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref/ref_driver = unmodified reference objects)
# --------------------------------------------------------------------------
def cpu_reference_run(sample_kr: int, sample_ks: int, reps: int, warm: int, threads: int):
    """Runs the reference's ExecuteQuery on a config-2-shaped sample; returns
    (probe tuples/s from the median timed repetition, description dict)."""
    exe = ROOT / "oracle" / "_ref" / "ref_driver"
    nr, ns = 1 << sample_kr, 1 << sample_ks
    sample = (f"config-2 shape scaled to |R|=2^{sample_kr} x |S|=2^{sample_ks} (same generator, same query), "
              f"{reps} timed repetitions after {warm}")
    if exe.exists():
        cmd = [str(exe), "-t", str(threads), "-r", str(reps + warm),
               f"synth:{nr}:perm{sample_kr}@0x51670D180001,pay@0x51670D180002",
               f"synth:{ns}:perm{sample_ks}@0x51670D180002,pay@0x51670D180003", "--", QUERY]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
        if out.returncode != 0:
            raise RuntimeError("ref_driver failed: " + out.stderr[-500:])
        secs = json.loads(out.stderr.strip().splitlines()[-1])["seconds"][warm:]
        lines = out.stdout.splitlines()
        kind = "reference"
    else:
        # the oracle port (single-threaded restatement), only when the reference was not compiled
        sys.path.insert(0, str(ROOT / "tests"))
        import orc
        kr = orc.synth_column(nr, 0, sample_kr, 0x51670D180001)
        ks = orc.synth_column(ns, 0, sample_ks, 0x51670D180002)
        pr = orc.synth_column(nr, 1, 0, 0x51670D180002)
        ps = orc.synth_column(ns, 1, 0, 0x51670D180003)
        secs, lines = [], []
        for i in range(reps + warm):
            t0 = time.perf_counter()
            sums, _ = orc.join_sum(kr, ks, [pr, ps], [0, 1], 4)
            if i >= warm:
                secs.append(time.perf_counter() - t0)
            lines.append(" ".join(map(str, sums)))
        kind, threads = "port", 1
    t = statistics.median(secs)
    return ns / t, {"kind": kind, "cores": threads, "sample": sample, "seconds_per_step": t,
                    "checksum_line": lines[-1] if lines else None}


def reference_threads() -> int:
    """Scheduler pool size for the reference (scheduler.c:9, run-time argument): min(cores, 16).
    The join fans out over 2^N_LSB = 16 buckets (rhjoin.c:42-57) and every PartitionJob re-scans its input
    once per bucket it spans (preprocess.c:262-296), so more than 16 threads only add work.  With 16 threads
    the reference intermittently loses about one bucket of tuples
    (tests/test_oracle_vs_reference.py::test_reference_loses_pairs_at_16_threads): cpu_reference_checked()
    verifies the checksum line it printed and falls back to 8 threads when it is wrong."""
    return max(1, min(os.cpu_count() or 1, 16))


def expected_checksum_line(sample_kr: int, sample_ks: int) -> str:
    """The result line of the config-2-shaped sample, from the generator alone (every R key matches once)."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import orc
    nr, ns = 1 << sample_kr, 1 << sample_ks
    pr = orc.synth_column(nr, 1, 0, 0x51670D180002)
    ks = orc.synth_column(ns, 0, sample_ks, 0x51670D180002)
    ps = orc.synth_column(ns, 1, 0, 0x51670D180003)
    return f"{int(pr.sum(dtype=np.uint64))} {int(ps[ks < nr].sum(dtype=np.uint64))}"


def cpu_reference_checked(sample_kr: int, sample_ks: int, reps: int, warm: int):
    """cpu_reference_run with the widest thread count whose output is right."""
    want = expected_checksum_line(sample_kr, sample_ks)
    threads = reference_threads()
    while True:
        value, info = cpu_reference_run(sample_kr, sample_ks, reps, warm, threads)
        info["checksum_ok"] = info.get("checksum_line") == want
        if info["checksum_ok"] or threads <= 8 or info["kind"] != "reference":
            return value, info
        threads = 8


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    skr, sks = 21, 25   # 1/8 of config 2: each step is seconds of CPU work
    value, info = cpu_reference_checked(skr, sks, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "config2: 2-relation uint64 equi-join |R|=2^24 x |S|=2^28, SUM projection",
                   "query": QUERY, "measured_on": info["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "checksum_ok": info["checksum_ok"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    b200 = load_package()
    L = b200.lib()
    _ck(L, L.b200_init(local))
    stream = torch.cuda.current_stream()
    L.b200_set_stream(stream.cuda_stream)        # kernels run on torch's current stream: torch events see them

    nr, ns = 1 << KR_BITS, 1 << KS_BITS
    max_key = ns - 1
    # position shards (N = 1: the whole relations)
    nr_loc, ns_loc = nr // world, ns // world
    r_first, s_first = rank * nr_loc, rank * ns_loc

    def synth(n, first, kind, k, seed):
        t = torch.empty(n, dtype=torch.int64, device=dev)
        b200.synth_column_device(t.data_ptr(), first, n, kind, k, seed)
        return t

    r0 = synth(nr_loc, r_first, b200.SYNTH_PERM, KR_BITS, b200.SEED_R)
    r1 = synth(nr_loc, r_first, b200.SYNTH_PAYLOAD, 0, b200.SEED_R + 1)
    s0 = synth(ns_loc, s_first, b200.SYNTH_PERM, KS_BITS, b200.SEED_S)
    s1 = synth(ns_loc, s_first, b200.SYNTH_PAYLOAD, 0, b200.SEED_S + 1)
    # the build-side SUM column's maximum is a column statistic (relation_map.c:53-61 keeps min/max per column); the
    # library uses it to carry 32-bit values inside the build tuples
    r1_max = int(r1.max().item())
    if world > 1:
        t_max = torch.tensor([r1_max], dtype=torch.int64, device=dev)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        r1_max = int(t_max.item())
    L.b200_register_device_column(r1.data_ptr(), r1.data_ptr(), nr_loc, r1_max)
    # ... and the probe-side SUM column's: with it the library may stream that column into the probe tuples instead
    # of gathering it per match
    s1_max = int(s1.max().item())
    L.b200_register_device_column(s1.data_ptr(), s1.data_ptr(), ns_loc, s1_max)
    plan = None
    # B200_PLAN: "copy" (default) / "scatter" = the two broadcast plans (small build side: config 2's shape);
    # "exchange" = radix-sharded all-to-all of both sides (config 4's plan), here for comparison
    plan_kind = os.environ.get("B200_PLAN", "copy")
    s1_max_all = s1_max
    if world > 1:
        t_max = torch.tensor([s1_max], dtype=torch.int64, device=dev)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        s1_max_all = int(t_max.item())
    if world > 1 and plan_kind == "exchange":
        plan = b200.sharding.ShardedExchangeJoin(b200, torch, dist, rank, world, nr, nr_loc, ns_loc, 1, 1, dev,
                                                 size_from=(r0.data_ptr(), s0.data_ptr()),
                                                 carry_build=r1_max < (1 << 32),
                                                 carry_probe=s1_max_all < (1 << 32))
    elif world > 1:
        plan = b200.sharding.BroadcastScatterJoin(b200, torch, dist, rank, world, nr, nr_loc, ns_loc, 1, dev,
                                                  carry32=r1_max < (1 << 32), rank_major=plan_kind == "copy",
                                                  carry_probe=s1_max_all < (1 << 32))   # matches per probe row: 1/16
    torch.cuda.synchronize()

    def step(kr=None, pr=None, ks=None, ps=None):
        """One pass of the hot path; returns ([sum R.c1, sum S.c1], matches): this rank's at N = 1, the
        all-reduced result at N > 1 (sharding.BroadcastScatterJoin)."""
        kr, pr = (r0 if kr is None else kr), (r1 if pr is None else pr)
        ks, ps = (s0 if ks is None else ks), (s1 if ps is None else ps)
        if world > 1 and plan_kind == "exchange":
            return plan.step(kr.data_ptr(), [pr.data_ptr()], ks.data_ptr(), [ps.data_ptr()])
        if world > 1:
            return plan.step(kr.data_ptr(), [pr.data_ptr()], ks.data_ptr(), [pr.data_ptr(), ps.data_ptr()], [0, 1])
        return b200.join_sum_device(kr.data_ptr(), nr, ks.data_ptr(), ns_loc, [pr.data_ptr(), ps.data_ptr()], [0, 1],
                                    max_key)

    def reduce_sums(sums, m):
        if world > 1:
            return sums, m          # plan.step already all-reduced them
        return [int(x) for x in sums], int(m)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # expected checksums from size-independent properties (no oracle in the product path)
    want_r = int(r1.sum().item())
    want_s_t = s1[s0 < nr].sum().reshape(1)
    want = torch.stack([torch.tensor(want_r, device=dev), want_s_t[0]])
    if world > 1:
        dist.all_reduce(want)
    want = [int(x) for x in want.cpu().numpy().view(np.uint64)]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs a moment to start: it samples warm-up and timed steps
    for _ in range(args.warmup):
        sums, m = step()
    barrier()
    b200.kernel_launches(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        sums, m = step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = b200.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    sums, m = reduce_sums(sums, m)
    if m != nr or sums != want:
        raise SystemExit(f"checksum mismatch: got {sums} m={m}, want {want} m={nr}")

    # ---- per-kernel device times (profiling events on, outside the timed region) ----
    L.b200_set_profiling(1)
    per_kernel = {}
    for _ in range(9):
        step()
        torch.cuda.synchronize()
        for name in ("hist", "hist_b", "hist_p", "scan", "scatter_b", "broadcast", "exchange", "scatter_p",
                     "scatter_pc", "join"):
            v = b200.last_kernel_ms(name)
            if v >= 0:
                per_kernel.setdefault(name, []).append(v)
    L.b200_set_profiling(0)
    per_kernel_range = {k: [min(v), max(v)] for k, v in per_kernel.items()}
    per_kernel = {k: statistics.median(v) for k, v in per_kernel.items()}

    # ---- end to end through the C-ABI with HOST buffers (N = 1 shard per rank) ----
    e2e = None
    if not args.no_e2e:
        host = {}
        for name, t in (("r0", r0), ("r1", r1), ("s0", s0), ("s1", s1)):
            h = torch.empty(t.shape, dtype=torch.int64, pin_memory=True)
            h.copy_(t)
            host[name] = h
        torch.cuda.synchronize()
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            if world > 1:
                # every rank uploads its shards, then the same multi-GPU plan as step()
                d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                return step(d["r0"], d["r1"], d["s0"], d["s1"])
            ptrs = (C.c_void_p * 2)(host["r1"].data_ptr(), host["s1"].data_ptr())
            sides = (C.c_int * 2)(0, 1)
            out = (C.c_uint64 * 2)()
            mm = C.c_uint64(0)
            rc = L.b200_join_sum(host["r0"].data_ptr(), nr, host["s0"].data_ptr(), ns_loc, max_key, 2, ptrs, sides, 0,
                                 out, C.byref(mm))
            _ck(L, rc)
            return [int(out[0]), int(out[1])], int(mm.value)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            es, em = e2e_step()
        barrier()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
        es, em = reduce_sums(es, em)
        if em != nr or es != want:
            raise SystemExit(f"e2e checksum mismatch: {es} {em}")
        e2e = {"value": ns / t_e2e, "unit": UNIT, "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": 8 * 2 * (nr + ns), "d2h_bytes_per_step": 8 * 3 * world,
               "host_memory": "pinned"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peaks()
    t_s = ms_step * 1e-3
    canon_bytes = CANON_PER_INPUT * (nr + ns) + CANON_PER_MATCH * nr
    # dominant kernel + its algorithmic bytes per launch (DESIGN.md "Kernels and rooflines")
    # canonical widths: key 8 B, partition tuple 16 B; encoded: the 8-B packed tuple used for keys < 2^32
    n_p, n_b = ns_loc, nr
    kernel_bytes = {
        "hist_p": (8 * n_p, 8 * n_p),
        "hist_b": (8 * n_b, 8 * n_b),
        "scatter_p": ((8 + 16) * n_p, (8 + 8) * n_p),
        "scatter_b": ((8 + 16) * n_b, (8 + 8) * n_b),
        "join": (16 * (n_p + n_b) + 16 * nr // world, 8 * (n_p + n_b) + 16 * nr // world),
    }
    if "scatter_pc" in per_kernel:
        # the probe scatter streamed S.c1 into the tuples (early materialisation): it reads 8 B per row more, and the
        # canonical bytes of that projection (8 B per match) are served by it instead of by the join
        m_loc = nr // world
        kernel_bytes["scatter_pc"] = ((8 + 16) * n_p + 8 * m_loc, (8 + 8 + 8) * n_p)
        kernel_bytes["join"] = (16 * (n_p + n_b) + 8 * m_loc, 8 * (n_p + n_b))
    traffic = {}
    tpath = ROOT / "profiles" / "r1_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text())
    roofline = None
    if per_kernel:
        dom = max((k for k in per_kernel if k in kernel_bytes), key=lambda k: per_kernel[k])
        canon, enc = kernel_bytes[dom]
        dur = per_kernel[dom] * 1e-3
        roofline = {"bound": "hbm", "kernel": dom, "achieved": canon / dur / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": canon / dur / 1e9 / peak,
                    "traffic": traffic.get(dom) if world == 1 else None, "traffic_source": traffic.get("source"),
                    "achieved_encoded": enc / dur / 1e9, "frac_encoded": enc / dur / 1e9 / peak,
                    "launch_ms": per_kernel[dom], "peak_source": peak_kind + " copy bandwidth (MEASURED_PEAKS.json)",
                    "bytes_per_launch_canonical": canon, "bytes_per_launch_encoded": enc,
                    "per_kernel_ms": per_kernel, "per_kernel_ms_min_max": per_kernel_range}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = cpu_reference_checked(21, 25, 2, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                        "sample": info["sample"], "seconds_per_step": info["seconds_per_step"],
                        "checksum_ok": info["checksum_ok"]}

    line = {
        "metric": METRIC, "value": ns / t_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "config2: 2-relation uint64 equi-join |R|=2^24 x |S|=2^28, unique permutation keys, "
                               "SUM projection", "query": QUERY, "rows_build": nr, "rows_probe": ns,
                   "matches": nr, "seeds": [hex(b200.SEED_R), hex(b200.SEED_S)],
                   "l2": "inputs (4.6 GB) larger than L2; no flush",
                   "parallelism": "single GPU" if world == 1 else
                   f"R and S position-sharded x{world}; both sides radix-partitioned locally, every partition stored "
                   "into its owner rank's buffers over NVLink (CUDA IPC; all-to-all), owners join, u64 all-reduce "
                   "of sums" if plan_kind == "exchange" else
                   f"R and S position-sharded x{world}; build shard scattered into every rank's partition buffers "
                   "by P2P stores over NVLink (CUDA IPC), probe shard partitioned locally, u64 all-reduce of sums"},
        "hbm": {"canonical_bytes_per_step": canon_bytes, "achieved_gbs": canon_bytes / t_s / 1e9 / world,
                "frac_of_peak": canon_bytes / t_s / 1e9 / world / peak, "peak_gbs": peak, "peak_source": peak_kind},
        "checksums": sums, "matches": m,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version there) get stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
