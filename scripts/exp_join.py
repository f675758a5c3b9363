#!/usr/bin/env python
"""Kernel-time experiments on the config-2 join (GPU box): per-kernel CUDA-event
times for different projection sets.  usage: python scripts/exp_join.py [kr_bits ks_bits]"""
import importlib.util
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("sigmod2018_b200", ROOT / "sigmod-2018_b200" / "__init__.py",
                                              submodule_search_locations=[str(ROOT / "sigmod-2018_b200")])
b200 = importlib.util.module_from_spec(spec)
sys.modules["sigmod2018_b200"] = b200
spec.loader.exec_module(b200)

kr_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ks_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 28
L = b200.lib()
L.b200_init(0)
nr, ns = 1 << kr_bits, 1 << ks_bits
r0, r1, s0, s1 = (b200.DeviceColumn(n) for n in (nr, nr, ns, ns))
b200.synth_column_device(r0.ptr, 0, nr, b200.SYNTH_PERM, kr_bits, b200.SEED_R)
b200.synth_column_device(r1.ptr, 0, nr, b200.SYNTH_PAYLOAD, 0, b200.SEED_R + 1)
b200.synth_column_device(s0.ptr, 0, ns, b200.SYNTH_PERM, ks_bits, b200.SEED_S)
b200.synth_column_device(s1.ptr, 0, ns, b200.SYNTH_PAYLOAD, 0, b200.SEED_S + 1)
L.b200_set_profiling(1)
variants = {"no projection": ([], []), "R payload only": ([r1.ptr], [0]), "S payload only": ([s1.ptr], [1]),
            "both": ([r1.ptr, s1.ptr], [0, 1])}
for name, (proj, sides) in variants.items():
    times = {}
    for _ in range(6):
        sums, m = b200.join_sum_device(r0.ptr, nr, s0.ptr, ns, proj, sides, ns - 1)
        for k in ("hist_b", "hist_p", "scan", "scatter_b", "scatter_p", "join"):
            times.setdefault(k, []).append(b200.last_kernel_ms(k))
    print(f"{name:16s} m={m} " + " ".join(f"{k}={statistics.median(v[1:]):.3f}" for k, v in times.items()))
