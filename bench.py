#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]

Default (what the driver runs) is config 2: synthetic 2-relation uint64 equi-join, |R| = 2^24 build x
|S| = 2^28 probe, unique permutation keys, query `0 1|0.0=1.0|0.1 1.1`.  One "step" is one pass of the whole hot
path over that input: radix partition of both sides, per-partition shared-memory build + probe, SUM projection.

  value     probe tuples/s, inputs resident in HBM (CUDA events on the stream the kernels are launched on, max
            over ranks)
  e2e       same metric through the public C-ABI call with HOST (pinned) buffers: the H2D copies of all four
            columns and the D2H of the sums are inside the timed region
  roofline  dominant kernel (per-kernel CUDA-event time measured live) against MEASURED_PEAKS.json; canonical
            algorithmic bytes and the bytes of the narrower encoding actually moved are both given
  cpu_baseline  the reference's own CPU join (oracle/_ref/ref_driver = the unmodified reference objects) on the
            SAME config, host cores stated

N > 1 (torchrun, one rank per GPU): the C-ABI multi-GPU plan (csrc/multi.cu, sharding.MultiJoin) — no NCCL in the
step.  `--impl reference` times the reference's CPU implementation of the same config (rank 0 only).
`--config 3|4|5` run the other BASELINE configs (one JSON line each, same keys; see run_config3/4/5).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import re
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
SEED_R, SEED_S = 0x51670D180001, 0x51670D180002
REF_BUDGET_S = 420.0          # wall-clock budget of the reference arm's timed repetitions


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


def config2_dict(world: int) -> dict:
    """The `config` object of the JSON line: the same in both arms (the driver compares them)."""
    nr, ns = 1 << KR_BITS, 1 << KS_BITS
    return {"workload": "config2: 2-relation uint64 equi-join |R|=2^24 x |S|=2^28, unique permutation keys, "
                        "SUM projection", "query": QUERY, "rows_build": nr, "rows_probe": ns, "matches": nr,
            "seeds": [hex(SEED_R), hex(SEED_S)], "l2": "inputs (4.6 GB) larger than L2; no flush",
            "parallelism": parallelism_text(world)}


def parallelism_text(world: int) -> str:
    if world == 1:
        return "single GPU"
    exchange = os.environ.get("B200_PLAN", "broadcast") == "exchange"
    return (f"R and S position-sharded x{world}; " +
            ("both sides radix-partitioned locally, every partition stored into its owner's buffers over NVLink "
             "(exchange kernel, CUDA IPC), owners join" if exchange else
             "each build shard partitioned once and pushed to every GPU by the copy engines (one region copy and one "
             "flag in peer memory per peer), probe shard partitioned locally meanwhile, the join waits for the regions it "
             "reads") +
            "; no NCCL in the step, results summed from peer-written slots, the step replayed from a CUDA graph")


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
def ref_driver_run(specs: list[str], query: str, reps: int, warm: int, threads: int, budget_s: float):
    """Runs the reference's ExecuteQuery (query.c:325-467) on synthetic relations; returns (seconds per timed
    repetition, last result line, load seconds).  Stops early when the time budget is spent (never before
    warm + 1 repetitions)."""
    exe = ROOT / "oracle" / "_ref" / "ref_driver"
    cmd = [str(exe), "-t", str(threads), "-r", str(reps + warm), "-w", str(warm), "-T", str(budget_s)] + specs + ["--", query]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1700)
    if out.returncode != 0:
        raise RuntimeError("ref_driver failed: " + out.stderr[-500:])
    info = json.loads(out.stderr.strip().splitlines()[-1])
    secs = info["seconds"][warm:]
    lines = out.stdout.splitlines()
    return secs, (lines[-1] if lines else None), info.get("load_s")


def config2_specs(kr: int, ks: int) -> list[str]:
    return [f"synth:{1 << kr}:perm{kr}@{hex(SEED_R)},pay@{hex(SEED_R + 1)}",
            f"synth:{1 << ks}:perm{ks}@{hex(SEED_S)},pay@{hex(SEED_S + 1)}"]


def expected_config2_line(kr: int, ks: int) -> str:
    """The result line of a config-2-shaped join from the generator alone (every R key matches exactly once)."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import orc
    nr, ns = 1 << kr, 1 << ks
    pr = orc.synth_column(nr, 1, 0, SEED_R + 1)
    total_r = int(pr.sum(dtype=np.uint64))
    total_s, chunk = 0, 1 << 24
    for a in range(0, ns, chunk):          # bounded memory at full size
        k = orc.synth_column(min(chunk, ns - a), 0, ks, SEED_S, first=a)
        p = orc.synth_column(min(chunk, ns - a), 1, 0, SEED_S + 1, first=a)
        total_s += int(p[k < nr].sum(dtype=np.uint64))
    return f"{total_r} {total_s % (1 << 64)}"


def reference_threads() -> int:
    """Scheduler pool size for the reference (scheduler.c:9, run-time argument): min(cores, 16).  The join fans
    out over 2^N_LSB = 16 buckets (rhjoin.c:42-57).  With 16 threads the reference intermittently loses about one
    bucket of tuples (tests/test_oracle_vs_reference.py::test_reference_loses_pairs_at_16_threads), so every line
    it prints is checked and the run falls back to 8 threads when one is wrong."""
    return max(1, min(os.cpu_count() or 1, 16))


def cpu_reference_config2(reps: int, warm: int, budget_s: float, kr: int = KR_BITS, ks: int = KS_BITS):
    """The reference's CPU join on config 2 at FULL size (same generator, same query): (probe tuples/s, info)."""
    exe = ROOT / "oracle" / "_ref" / "ref_driver"
    ns = 1 << ks
    if not exe.exists():
        # the oracle port (single-threaded restatement), only where the reference was not compiled
        sys.path.insert(0, str(ROOT / "tests"))
        import orc
        skr, sks = min(kr, 20), min(ks, 24)
        a = [orc.synth_column(1 << skr, 0, skr, SEED_R), orc.synth_column(1 << sks, 0, sks, SEED_S),
             orc.synth_column(1 << skr, 1, 0, SEED_R + 1), orc.synth_column(1 << sks, 1, 0, SEED_S + 1)]
        secs = []
        for i in range(warm + max(1, min(reps, 3))):
            t0 = time.perf_counter()
            orc.join_sum(a[0], a[1], [a[2], a[3]], [0, 1], 4)
            if i >= warm:
                secs.append(time.perf_counter() - t0)
        t = statistics.median(secs)
        return (1 << sks) / t, {"kind": "port", "cores": 1, "seconds_per_step": t, "steps_timed": len(secs),
                                "sample": f"oracle port on |R|=2^{skr} x |S|=2^{sks} (reference not compiled here)",
                                "checksum_ok": True, "same_config": False}
    want = expected_config2_line(kr, ks)
    threads = reference_threads()
    while True:
        secs, line, load_s = ref_driver_run(config2_specs(kr, ks), QUERY, reps, warm, threads, budget_s)
        ok = line == want
        if ok or threads <= 8:
            break
        threads = 8
    t = statistics.median(secs)
    capped = len(secs) < reps
    sample = (f"config 2 at full size |R|=2^{kr} x |S|=2^{ks} (same generator, same query), {len(secs)} timed "
              f"repetitions after {warm}" + (f" ({reps} asked; stopped at the {budget_s:.0f} s budget)" if capped else ""))
    return ns / t, {"kind": "reference", "cores": threads, "seconds_per_step": t, "steps_timed": len(secs),
                    "sample": sample, "checksum_ok": ok, "same_config": (kr, ks) == (KR_BITS, KS_BITS),
                    "load_s": load_s, "capped": capped}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != 2:
        raise SystemExit("--impl reference is the config-2 arm; the other configs time the reference inside their own line")
    value, info = cpu_reference_config2(args.steps, args.warmup, REF_BUDGET_S, KR_BITS - args.scale_bits,
                                        KS_BITS - args.scale_bits)
    cfg = config2_dict(args.gpus)
    if args.scale_bits:
        cfg["workload"] += f" -- SCALED DOWN by 2^{args.scale_bits} (--scale-bits: a quick check, not the benchmark)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": info["steps_timed"],
        "ms_per_step": info["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "checksum_ok": info["checksum_ok"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------
# shared set-up of the B200 arms
# --------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # rendezvous and set-up traffic only (IPC handles, expected checksums, max-over-ranks of the timing):
            # the step itself never calls a collective
            dist.init_process_group("nccl", device_id=self.dev)
        self.b200 = load_package()
        self.L = self.b200.lib()
        _ck(self.L, self.L.b200_init(self.local))
        self.stream = torch.cuda.current_stream()
        self.L.b200_set_stream(self.stream.cuda_stream)   # kernels run on torch's current stream: torch events see them

    def synth(self, n, first, kind, k, seed):
        t = self.torch.empty(max(n, 1), dtype=self.torch.int64, device=self.dev)[:n]
        if n:
            self.b200.synth_column_device(t.data_ptr(), first, n, kind, k, seed)
        return t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def traffic_record():
    """ncu-measured DRAM bytes per launch of the dominant kernels: a CITATION of a committed capture (it says which
    commit and profile it belongs to), not something measured in this run."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = ROOT / "profiles" / name
        if p.exists():
            d = json.loads(p.read_text())
            d["file"] = "profiles/" + name
            return d
    return {}


# --------------------------------------------------------------------------
# config 2 (the driver's bench)
# --------------------------------------------------------------------------
def run_config2(args):
    import numpy as np
    env = Env(args)
    torch, dist, b200, L, dev = env.torch, env.dist, env.b200, env.L, env.dev
    world, rank = env.world, env.rank
    kr_bits, ks_bits = KR_BITS - args.scale_bits, KS_BITS - args.scale_bits   # --scale-bits: quick checks only
    nr, ns = 1 << kr_bits, 1 << ks_bits
    max_key = ns - 1
    nr_loc, ns_loc = nr // world, ns // world
    r_first, s_first = rank * nr_loc, rank * ns_loc
    r0 = env.synth(nr_loc, r_first, b200.SYNTH_PERM, kr_bits, SEED_R)
    r1 = env.synth(nr_loc, r_first, b200.SYNTH_PAYLOAD, 0, SEED_R + 1)
    s0 = env.synth(ns_loc, s_first, b200.SYNTH_PERM, ks_bits, SEED_S)
    s1 = env.synth(ns_loc, s_first, b200.SYNTH_PAYLOAD, 0, SEED_S + 1)
    # column maxima are column statistics (relation_map.c:53-61 keeps min/max per column): with them the library
    # carries 32-bit SUM values inside the partition tuples
    r1_max, s1_max = int(r1.max().item()), int(s1.max().item())
    L.b200_register_device_column(r1.data_ptr(), r1.data_ptr(), nr_loc, r1_max)
    L.b200_register_device_column(s1.data_ptr(), s1.data_ptr(), ns_loc, s1_max)
    plan = None
    plan_kind = os.environ.get("B200_PLAN", "broadcast")
    if world > 1:
        if max(env.max_over_ranks(r1_max), env.max_over_ranks(s1_max)) >= (1 << 32):
            raise SystemExit("the C-ABI multi-GPU plans carry SUM values below 2^32")
        plan = b200.sharding.MultiJoin(b200, dist, rank, world, env.local,
                                       b200.PLAN_EXCHANGE if plan_kind == "exchange" else b200.PLAN_BROADCAST,
                                       nr_loc, ns_loc, True, True)
    torch.cuda.synchronize()

    def step(kr=None, pr=None, ks=None, ps=None):
        """One pass of the hot path; returns ([sum R.c1, sum S.c1], matches) over ALL ranks."""
        kr, pr = (r0 if kr is None else kr), (r1 if pr is None else pr)
        ks, ps = (s0 if ks is None else ks), (s1 if ps is None else ps)
        if world > 1:
            return plan.step(kr.data_ptr(), pr.data_ptr(), ks.data_ptr(), ps.data_ptr())
        return b200.join_sum_device(kr.data_ptr(), nr, ks.data_ptr(), ns_loc, [pr.data_ptr(), ps.data_ptr()], [0, 1],
                                    max_key)

    # expected checksums from size-independent properties (no oracle in the product path)
    want = torch.stack([r1.sum(), s1[s0 < nr].sum()])
    if world > 1:
        dist.all_reduce(want)
    want = [int(x) for x in want.cpu().numpy().view(np.uint64)]

    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs a moment to start: it samples warm-up and timed steps
    for _ in range(args.warmup):
        sums, m = step()
    env.barrier()
    b200.kernel_launches(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev0.record(env.stream)
    for _ in range(args.steps):
        sums, m = step()
    ev1.record(env.stream)
    env.barrier()
    ms_step = env.max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = b200.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    if m != nr or [int(x) for x in sums] != want:
        raise SystemExit(f"checksum mismatch: got {sums} m={m}, want {want} m={nr}")

    # ---- per-kernel device times (profiling events on, outside the timed region) ----
    L.b200_set_profiling(1)
    per_kernel = {}
    for _ in range(9):
        step()
        torch.cuda.synchronize()
        for name in ("hist", "hist_b", "hist_p", "scan", "scatter_b", "exchange", "scatter_p", "scatter_pc", "join"):
            v = b200.last_kernel_ms(name)
            if v >= 0:
                per_kernel.setdefault(name, []).append(v)
    L.b200_set_profiling(0)
    per_kernel_range = {k: [min(v), max(v)] for k, v in per_kernel.items()}
    per_kernel = {k: statistics.median(v) for k, v in per_kernel.items()}

    # ---- end to end through the C-ABI with HOST buffers ----
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
            _ck(L, L.b200_join_sum(host["r0"].data_ptr(), nr, host["s0"].data_ptr(), ns_loc, max_key, 2, ptrs, sides, 0,
                                   out, C.byref(mm)))
            return [int(out[0]), int(out[1])], int(mm.value)

        e2e_step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            es, em = e2e_step()
        env.barrier()
        t_e2e = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        if em != nr or [int(x) for x in es] != want:
            raise SystemExit(f"e2e checksum mismatch: {es} {em}")
        e2e = {"value": ns / t_e2e, "unit": UNIT, "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
               "h2d_bytes_per_step": 8 * 2 * (nr + ns), "d2h_bytes_per_step": 8 * 3 * world,
               "host_memory": "pinned"}

    if plan is not None:
        plan.close()
    if rank != 0:
        env.close()
        return

    peak, peak_kind = measured_peaks()
    t_s = ms_step * 1e-3
    canon_bytes = CANON_PER_INPUT * (nr + ns) + CANON_PER_MATCH * nr
    # dominant kernel + its algorithmic bytes per launch: canonical widths (key 8 B,
    # partition tuple 16 B) and the encoding actually moved (8-B packed tuples for keys < 2^32)
    n_p, n_b = ns_loc, nr
    kernel_bytes = {
        "hist_p": (8 * n_p, 8 * n_p),
        "hist_b": (8 * n_b // world, 8 * n_b // world),
        "scatter_p": ((8 + 16) * n_p, (8 + 8) * n_p),
        "scatter_b": ((8 + 16) * n_b // world, (8 + 8) * n_b // world),
        "join": (16 * (n_p + n_b) + 16 * nr // world, 8 * (n_p + n_b) + 16 * nr // world),
    }
    if "scatter_pc" in per_kernel:
        # the probe scatter streamed S.c1 into the tuples (early materialisation): it reads 8 B per row more, and the
        # canonical bytes of that projection (8 B per match) are served by it instead of by the join
        m_loc = nr // world
        kernel_bytes["scatter_pc"] = ((8 + 16) * n_p + 8 * m_loc, (8 + 8 + 8) * n_p)
        kernel_bytes["join"] = (16 * (n_p + n_b) + 8 * m_loc, 8 * (n_p + n_b))
    traffic = traffic_record()
    roofline = None
    if per_kernel:
        dom = max((k for k in per_kernel if k in kernel_bytes), key=lambda k: per_kernel[k])
        canon, enc = kernel_bytes[dom]
        dur = per_kernel[dom] * 1e-3
        roofline = {"bound": "hbm", "kernel": dom, "achieved": canon / dur / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": canon / dur / 1e9 / peak,
                    "traffic": traffic.get(dom) if world == 1 and not args.scale_bits else None,
                    "traffic_source": traffic.get("source"), "traffic_commit": traffic.get("commit"),
                    "traffic_file": traffic.get("file"),
                    "achieved_encoded": enc / dur / 1e9, "frac_encoded": enc / dur / 1e9 / peak,
                    "launch_ms": per_kernel[dom], "peak_source": peak_kind + " copy bandwidth (MEASURED_PEAKS.json)",
                    "bytes_per_launch_canonical": canon, "bytes_per_launch_encoded": enc,
                    "per_kernel_ms": per_kernel, "per_kernel_ms_min_max": per_kernel_range}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, info = cpu_reference_config2(3, 1, 90.0, kr_bits, ks_bits)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                        "sample": info["sample"], "seconds_per_step": info["seconds_per_step"],
                        "checksum_ok": info["checksum_ok"], "same_config": info["same_config"]}

    cfg = config2_dict(world)
    if args.scale_bits:
        cfg["workload"] += f" -- SCALED DOWN by 2^{args.scale_bits} (--scale-bits: a quick check, not the benchmark)"
    line = {
        "metric": METRIC, "value": ns / t_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
        "hbm": {"canonical_bytes_per_step": canon_bytes, "achieved_gbs": canon_bytes / t_s / 1e9 / world,
                "frac_of_peak": canon_bytes / t_s / 1e9 / world / peak, "peak_gbs": peak, "peak_source": peak_kind},
        "checksums": [int(x) for x in sums], "matches": m,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line))
    env.close()


# --------------------------------------------------------------------------
# config 3: 4-way chain join with range / equality filters over a 200 M-row fact relation
# --------------------------------------------------------------------------
CONFIG3_QUERY = "0 1 2 3|0.1=1.0&1.1=2.0&2.1=3.0&0.3>2499&0.3<7500&0.4=1|0.2 1.2 3.1"


def config3_shape(scale_bits: int):
    f = 200_000_000 >> scale_bits
    d1, d2, d3 = (1 << 24) >> scale_bits, (1 << 20) >> min(scale_bits, 8), (1 << 16) >> min(scale_bits, 4)
    # (rows, [(kind, k, seed), ...]) — the columns the reference driver (oracle/_ref/ref_driver) is handed as well
    return [(f, [("iota", 0, 0), ("uni", d1, 11), ("pay", 0, 12), ("uni", 10000, 13), ("uni", 4, 14)]),
            (d1, [("iota", 0, 0), ("uni", d2, 21), ("pay", 0, 22)]),
            (d2, [("iota", 0, 0), ("uni", d3, 31), ("pay", 0, 32)]),
            (d3, [("iota", 0, 0), ("pay", 0, 41)])]


def config3_specs(shape) -> list[str]:
    def col(kind, k, seed):
        return "iota" if kind == "iota" else (f"pay@{seed}" if kind == "pay" else f"uni{k}@{seed}")
    return [f"synth:{rows}:" + ",".join(col(*c) for c in cols) for rows, cols in shape]


def run_config3(args):
    env = Env(args)
    if env.world != 1:
        raise SystemExit("config 3 is a single-GPU config")
    torch, b200, L = env.torch, env.b200, env.L
    shape = config3_shape(args.scale_bits)
    kinds = {"iota": b200.SYNTH_IOTA, "uni": b200.SYNTH_UNIFORM, "pay": b200.SYNTH_PAYLOAD}
    rels = [[env.synth(rows, 0, kinds[k], kk, seed) for k, kk, seed in cols] for rows, cols in shape]
    F, D1, D2, D3 = rels
    torch.cuda.synchronize()
    # expected sums without the join kernels: every dimension key is its row number, so a join is an index
    mask = (F[3] > 2499) & (F[3] < 7500) & (F[4] == 1)
    f1 = F[1][mask]
    d2 = D1[1][f1]
    d3 = D2[1][d2]
    want = [int(F[2][mask].sum().item()), int(D1[2][f1].sum().item()), int(D3[1][d3].sum().item())]
    n_hit = int(mask.sum().item())
    del mask, f1, d2, d3
    rel_map = b200.DeviceRelationMap([[(c.data_ptr(), c.numel(), int(c.max().item()) if c.numel() else 0) for c in rel]
                                      for rel in rels])

    def query():
        return b200.execute_query(CONFIG3_QUERY, rel_map)

    sampler = ClockSampler(env.local)
    sampler.start()
    for _ in range(args.warmup):
        res = query()
    torch.cuda.synchronize()
    b200.kernel_launches(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(env.stream)
    for _ in range(args.steps):
        res = query()
    ev1.record(env.stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = b200.kernel_launches()
    clocks = sampler.stop()
    ok = res.sums == want
    if not ok:
        raise SystemExit(f"config 3 checksum mismatch: got {res.sums}, want {want}")
    L.b200_set_profiling(1)
    per_kernel = {}
    for _ in range(5):
        query()
        torch.cuda.synchronize()
        for name in ("filter", "filter_fused", "gather", "hist_b", "hist_p", "scatter_b", "scatter_p", "scatter_pc",
                     "join", "join_write", "checksum"):
            v = b200.last_kernel_ms(name)
            if v >= 0:
                per_kernel.setdefault(name, []).append(v)
    L.b200_set_profiling(0)
    per_kernel = {k: statistics.median(v) for k, v in per_kernel.items()}
    # canonical bytes for the textual left-deep order with the measured cardinalities
    f_rows, d1_rows, d2_rows, d3_rows = (s[0] for s in shape)
    m = n_hit
    canon = (2 * 8 * f_rows + 8 * m          # two distinct filter columns + the surviving row-id list
             + (48 * m + 40 * d1_rows) + 8 * m * 3      # J1 (fact side through the row-id list), out a = 1
             + (48 * m + 40 * d2_rows) + 8 * m * 5      # J2, out a = 2
             + (48 * m + 40 * d3_rows)                  # J3 (final: not materialised)
             + 8 * 3 * m + 8 * 3 * m)                   # projection: gathers + ids
    peak, peak_kind = measured_peaks()
    cpu_baseline = None
    if not args.no_cpu_baseline and (ROOT / "oracle" / "_ref" / "ref_driver").exists():
        secs, line, _ = ref_driver_run(config3_specs(shape), CONFIG3_QUERY, 2, 0, 8, 60.0)
        t = statistics.median(secs)
        cpu_baseline = {"value": f_rows / t, "unit": "fact rows/s", "cores": 8, "kind": "reference",
                        "sample": f"the same query on the same relations through the reference's ExecuteQuery, {len(secs)} "
                                  "repetitions (8 threads: at 16 the reference loses tuples)",
                        "seconds_per_step": t, "checksum_ok": line == " ".join(map(str, want)), "line": line}
    dom = max(per_kernel, key=per_kernel.get) if per_kernel else None
    print(json.dumps({
        "metric": "chain_join_fact_rows_per_s", "value": f_rows / (ms * 1e-3), "unit": "fact rows/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "config3: 4-way chain join with range/equality filters and a 3-column SUM projection over a "
                               f"{f_rows}-row fact relation", "query": CONFIG3_QUERY, "rows": [s[0] for s in shape],
                   "fact_rows_after_filters": m, "l2": "inputs (8.6 GB) larger than L2; no flush",
                   "host": "host.py execute_query over the reference's operator API (ExecuteQuery's order)"},
        "hbm": {"canonical_bytes_per_step": canon, "achieved_gbs": canon / (ms * 1e-3) / 1e9,
                "frac_of_peak": canon / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak, "peak_source": peak_kind},
        "roofline": {"bound": "hbm", "kernel": dom, "per_kernel_ms": per_kernel, "achieved": canon / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": canon / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "note": "whole query: canonical bytes of all operators / device time of the query"},
        "checksums": res.sums, "checksum_ok": ok, "rows": res.rows, "cpu_baseline": cpu_baseline,
        "gpu_launches": launches, "clocks": clocks,
    }))
    env.close()


# --------------------------------------------------------------------------
# config 4: Zipf(1.0) probe x unique build, radix-sharded across the GPUs with an all-to-all over NVLink
# --------------------------------------------------------------------------
def run_config4(args):
    import numpy as np
    env = Env(args)
    torch, dist, b200, L, dev = env.torch, env.dist, env.b200, env.L, env.dev
    world, rank = env.world, env.rank
    sh = b200.sharding
    # full size on 8 GPUs; fewer GPUs keep the per-GPU load (rows scale with N / 8)
    k = 27 - args.scale_bits - {8: 0, 4: 1, 2: 2, 1: 3}[world]
    nr = 1 << k
    ns = (2_000_000_000 >> args.scale_bits) * world // 8
    r_first, nr_loc = sh.shard_bounds(nr, rank, world)
    s_first, ns_loc = sh.shard_bounds(ns, rank, world)
    r0 = env.synth(nr_loc, r_first, b200.SYNTH_PERM, k, SEED_R)
    r1 = env.synth(nr_loc, r_first, b200.SYNTH_PAYLOAD, 0, SEED_R + 1)
    s0 = env.synth(ns_loc, s_first, b200.SYNTH_UNIFORM if args.uniform else b200.SYNTH_ZIPF, nr if args.uniform else k,
                   SEED_S)
    s1 = env.synth(ns_loc, s_first, b200.SYNTH_PAYLOAD, 0, SEED_S + 1)
    # expected checksums without the join kernels: lut[key] = R.c1 of the row holding that key
    r0_all = sh.allgather_column(r0, nr, dist) if world > 1 else r0
    r1_all = sh.allgather_column(r1, nr, dist) if world > 1 else r1
    lut = torch.empty(nr, dtype=torch.int64, device=dev)
    lut[r0_all] = r1_all
    want = torch.zeros(3, dtype=torch.int64, device=dev)
    for a in range(0, ns_loc, 1 << 26):
        want[0] += lut[s0[a:a + (1 << 26)]].sum()
    want[1] = s1.sum()
    want[2] = ns_loc
    hot = int(torch.bincount(s0[: min(ns_loc, 1 << 24)] & 0xFFFF, minlength=1 << 16).max().item())
    del lut, r0_all, r1_all
    if world > 1:
        dist.all_reduce(want)
    want = [int(x) for x in want.cpu().numpy().view(np.uint64)]
    plan = sh.MultiJoin(b200, dist if world > 1 else None, rank, world, env.local, b200.PLAN_EXCHANGE, nr_loc, ns_loc,
                        True, True, chunks=args.chunks, radix_bits=args.radix_bits)
    torch.cuda.synchronize()

    def step():
        return plan.step(r0.data_ptr(), r1.data_ptr(), s0.data_ptr(), s1.data_ptr())

    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        sums, m = step()
    if [sums[0], sums[1], m] != want:
        raise SystemExit(f"config 4 checksum mismatch: got {sums} m={m}, want {want}")
    env.barrier()
    b200.kernel_launches(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(env.stream)
    for _ in range(args.steps):
        sums, m = step()
    ev1.record(env.stream)
    env.barrier()
    ms = env.max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = b200.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    ok = [sums[0], sums[1], m] == want
    recv = plan.received()
    recv_all = [None] * world
    if world > 1:
        dist.all_gather_object(recv_all, recv)
    else:
        recv_all = [recv]
    L.b200_set_profiling(1)
    per_kernel = {}
    for _ in range(3):
        step()
        torch.cuda.synchronize()
        for name in ("hot_sample", "hot_table", "hot_build", "hist", "layout", "scatter_b", "exchange", "join"):
            v = b200.last_kernel_ms(name)
            if v >= 0:
                per_kernel.setdefault(name, []).append(round(v, 4))
    L.b200_set_profiling(0)
    bits = plan.bits
    plan.close()
    if rank == 0:
        rows = [b + p for b, p in recv_all]
        # mean bytes leaving a rank per step: the rows that were exchanged (probe rows with a hot key are joined where
        # they are and never travel), of which (world - 1) / world land on another GPU
        wire = 8 * sum(rows) / world * (world - 1) / max(world, 1)
        canon = 40 * (nr + ns) + 16 * ns
        peak, peak_kind = measured_peaks()
        nvlink_peak = 770.0
        print(json.dumps({
            "metric": METRIC, "value": ns / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": (f"config4: Zipf(1.0) probe {ns} rows x unique build 2^{k}, radix-sharded x{world} with an "
                                    "all-to-all over NVLink" if not args.uniform else
                                    f"control: uniform probe {ns} rows x unique build 2^{k}, x{world}"),
                       "query": QUERY, "full_size": world == 8 and args.scale_bits == 0, "radix_bits": bits,
                       "parallelism": "exchange plan (csrc/multi.cu): local partition passes, exchange kernel stores into the "
                                      "owners' receive buffers, ownership cuts balanced on the global histogram"},
            "checksums": sums, "matches": m, "checksum_ok": ok,
            "rows_received_per_rank": recv_all, "rows_received_max_over_mean": max(rows) / (sum(rows) / world),
            "probe_rows_joined_in_place": ns - sum(p for _, p in recv_all),
            "hbm": {"canonical_bytes_per_step": canon, "achieved_gbs": canon / (ms * 1e-3) / 1e9 / world,
                    "frac_of_peak": canon / (ms * 1e-3) / 1e9 / world / peak, "peak_gbs": peak, "peak_source": peak_kind},
            "nvlink": {"out_bytes_per_rank": wire, "out_gbs_per_rank_over_step": wire / (ms * 1e-3) / 1e9,
                       "exchange_kernel_ms_rank0": per_kernel.get("exchange"),
                       "peak_gbs_per_direction": nvlink_peak, "peak_source": "measured peer copy (B200_PROFILING.md)",
                       "frac_of_peak_over_step": wire / (ms * 1e-3) / 1e9 / nvlink_peak},
            "roofline": {"bound": "nvlink" if world > 1 else "hbm", "achieved": wire / (ms * 1e-3) / 1e9, "peak": nvlink_peak,
                         "unit": "GB/s", "frac": wire / (ms * 1e-3) / 1e9 / nvlink_peak, "traffic": None,
                         "per_kernel_ms_rank0": per_kernel},
            "hottest_16bit_key_share_sample": hot / max(min(ns_loc, 1 << 24), 1),
            "cpu_baseline": None, "gpu_launches": launches, "clocks": clocks,
        }))
    env.close()


# --------------------------------------------------------------------------
# config 5: the small.work batch on the small schema scaled xF, concurrent queries on GPU streams
# --------------------------------------------------------------------------
def run_config5(args):
    factor = args.factor
    if int(os.environ.get("RANK", "0")) != 0:
        return
    work = Path(args.workdir or f"/tmp/scaled_small_{factor}")
    if not (work / "scaled.work").exists():
        subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scaled_small.py"), str(factor), str(work)], check=True,
                       stdout=sys.stderr)
    stdin = "\n".join((work / "scaled.init").read_text().split()) + "\nDone\n" + (work / "scaled.work").read_text()
    engine = ROOT / "host" / "b200_engine"
    runs, lines_by_w = {}, {}
    for w in args.workers:
        t0 = time.perf_counter()
        out = subprocess.run([str(engine), "-w", str(w)], input=stdin, capture_output=True, text=True, cwd=work,
                             env=dict(os.environ, B200_TIMING="1"), timeout=3000)
        wall = time.perf_counter() - t0
        if out.returncode != 0:
            raise SystemExit(f"b200_engine -w {w} failed (rc {out.returncode}): {out.stderr[-800:]}")
        batches = [float(x) for x in re.findall(r"workers: ([0-9.]+) s", out.stderr)]
        nq = len(out.stdout.splitlines())
        runs[str(w)] = {"batch_seconds": batches, "query_seconds": sum(batches), "queries": nq,
                        "queries_per_s": nq / sum(batches) if batches else None, "process_wall_s": wall,
                        "startup": (out.stderr.splitlines() or [""])[0]}
        lines_by_w[w] = out.stdout.splitlines()
    base = lines_by_w[args.workers[0]]
    same = all(lines_by_w[w] == base for w in args.workers)
    parity = None
    ref_bin = ROOT / "oracle" / "_ref" / "radixhash"
    if args.check_reference and ref_bin.exists():
        t0 = time.perf_counter()
        ref = subprocess.run([str(ref_bin)], input=stdin, capture_output=True, text=True, cwd=work, timeout=3000)
        ref_lines = ref.stdout.splitlines()
        diff = [i for i, (a, b) in enumerate(zip(ref_lines, base)) if a != b]
        parity = {"reference_wall_s": time.perf_counter() - t0, "lines": len(ref_lines),
                  "identical": len(ref_lines) - len(diff), "differing_queries": [i + 1 for i in diff],
                  "reference_rc": ref.returncode}
    best = min(runs.values(), key=lambda r: r["query_seconds"])
    print(json.dumps({
        "metric": "batch_queries_per_s", "value": best["queries_per_s"], "unit": "queries/s", "n_gpus": 1, "steps": 1,
        "warmup": 0, "ms_per_step": best["query_seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"config5: small.work (50 queries, 5 batches) on the 14-relation small schema scaled x{factor} "
                               "(rows and key domains), concurrent queries on per-worker CUDA streams (host/b200_engine)",
                   "factor": factor, "workers": args.workers},
        "runs_by_workers": runs, "output_identical_across_worker_counts": same, "parity_vs_reference": parity,
        "lines_head": base[:3], "gpu_launches": None,
    }))


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version there) get stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scale-bits", type=int, default=0, help="shrink the workload by 2^n (quick checks only)")
    ap.add_argument("--uniform", action="store_true", help="config 4 control: uniform probe keys")
    ap.add_argument("--chunks", type=int, default=0, help="config 4: probe chunks of the exchange (0 = default)")
    ap.add_argument("--radix-bits", type=int, default=0, help="config 4: radix bits of the exchange plan (0 = automatic)")
    ap.add_argument("--factor", type=int, default=1000, help="config 5: scale factor of the small schema")
    ap.add_argument("--workers", type=lambda s: [int(x) for x in s.split(",")], default=[1, 4, 8])
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--check-reference", action="store_true", help="config 5: also run the reference binary and diff")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == 2:
        run_config2(args)
    elif args.config == 3:
        run_config3(args)
    elif args.config == 4:
        run_config4(args)
    else:
        run_config5(args)


if __name__ == "__main__":
    main()
