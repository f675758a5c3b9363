#!/usr/bin/env python
"""Queries/s of the small.work batch with S concurrent worker threads (one CUDA stream each) on the GPU
library, next to the reference binary's time for the same batch.  usage: python scripts/exp_batch.py"""
import importlib.util, sys, time, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
spec = importlib.util.spec_from_file_location("sigmod2018_b200", ROOT / "sigmod-2018_b200" / "__init__.py",
                                              submodule_search_locations=[str(ROOT / "sigmod-2018_b200")])
b200 = importlib.util.module_from_spec(spec); sys.modules["sigmod2018_b200"] = b200; spec.loader.exec_module(b200)
from golden_cases import load_small, small_queries
rels = load_small()
queries, golden = small_queries()
b200.lib().b200_init(0)
rm = b200.RelationMapArray(rels)
b200.execute_batch(queries, rm, 1)          # warm-up: uploads, module load, memory pool
for workers in (1, 2, 4, 8, 16):
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        got = b200.execute_batch(queries, rm, workers)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert [r.line() for r in got] == golden
    print(f"workers={workers:2d}: {best*1e3:7.1f} ms for {len(queries)} queries = {len(queries)/best:7.1f} queries/s")
small = ROOT / "oracle" / "_ref" / "small"
stdin = (small / "small.init").read_text() + "Done\n" + (small / "small.work").read_text()
t0 = time.perf_counter()
subprocess.run([str(ROOT / "oracle" / "_ref" / "radixhash")], input=stdin, capture_output=True, text=True, cwd=small)
print(f"reference radixhash, whole process (load + 50 queries): {(time.perf_counter()-t0)*1e3:.1f} ms")
