#!/usr/bin/env python
"""Operator-API path at config-2 scale: GetRelation x2 -> RadixHashJoin (pairs materialised) ->
InsertJoinToInterResults -> CalculateQueryResults, on device-resident columns (wall clock per query and
per-kernel CUDA-event times).  usage: python scripts/exp_pairs.py [kr_bits ks_bits]"""
import ctypes as C, importlib.util, statistics, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("sigmod2018_b200", ROOT / "sigmod-2018_b200" / "__init__.py",
                                              submodule_search_locations=[str(ROOT / "sigmod-2018_b200")])
b200 = importlib.util.module_from_spec(spec); sys.modules["sigmod2018_b200"] = b200; spec.loader.exec_module(b200)
h = b200.host
kr_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ks_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 28
L = b200.lib(); L.b200_init(0)
nr, ns = 1 << kr_bits, 1 << ks_bits
cols = {k: b200.DeviceColumn(n) for k, n in (("r0", nr), ("r1", nr), ("s0", ns), ("s1", ns))}
b200.synth_column_device(cols["r0"].ptr, 0, nr, b200.SYNTH_PERM, kr_bits, b200.SEED_R)
b200.synth_column_device(cols["r1"].ptr, 0, nr, b200.SYNTH_PAYLOAD, 0, b200.SEED_R + 1)
b200.synth_column_device(cols["s0"].ptr, 0, ns, b200.SYNTH_PERM, ks_bits, b200.SEED_S)
b200.synth_column_device(cols["s1"].ptr, 0, ns, b200.SYNTH_PAYLOAD, 0, b200.SEED_S + 1)
# a relation_map whose host column pointers are only registry keys (the columns live in HBM)
rm = (h.CRelationMap * 2)()
keep = []
for r, (names, n, kmax) in enumerate(((("r0", "r1"), nr, nr - 1), (("s0", "s1"), ns, ns - 1))):
    ptrs = (h.u64p * 2)()
    for j, name in enumerate(names):
        fake = C.cast(C.c_void_p(0x1000 + 0x100 * r + 8 * j), h.u64p)
        ptrs[j] = fake
        L.b200_register_device_column(0x1000 + 0x100 * r + 8 * j, cols[name].ptr, n, kmax if j == 0 else (1 << 24) - 1)
    keep.append(ptrs)
    rm[r].num_tuples, rm[r].num_columns, rm[r].columns = n, 2, ptrs

class FakeMap:
    array = rm
    def register(self): pass
L.b200_set_profiling(1)
times = []
for rep in range(5):
    t0 = time.perf_counter()
    res = b200.execute_query("0 1|0.0=1.0|0.1 1.1", FakeMap())
    times.append(time.perf_counter() - t0)
    kt = {k: b200.last_kernel_ms(k) for k in ("hist_b", "hist_p", "scan", "scatter_b", "scatter_p", "join", "join_write")}
print("result", res.line(), "rows", res.rows)
print("wall ms per query:", [round(t * 1e3, 2) for t in times])
print("last query kernels (ms):", {k: round(v, 3) for k, v in kt.items()})
