#!/bin/bash
# 1 x B200: the checked-build test, then config 5 at x100 (diffed against the reference binary) and x1000.
set -u
T=${1:-r2y}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_checked_build_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "checked or repeatable" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -25 gpurun_out/${T}_pytest.log | cut -c1-300
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "query ms", round(d["ms_per_step"], 2), "q/s", round(d["value"], 1), "same output for every worker count:",
          d["output_identical_across_worker_counts"], "vs reference:", d["parity_vs_reference"])
    for w, r in d["runs_by_workers"].items():
        print("   workers", w, "batches", r["batch_seconds"], "wall", round(r["process_wall_s"], 2), r["startup"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 500 python bench.py --config 5 --factor 100 --check-reference > gpurun_out/${T}_config5_x100.json 2> gpurun_out/${T}_config5_x100.err; show gpurun_out/${T}_config5_x100.json; tail -3 gpurun_out/${T}_config5_x100.err | cut -c1-300
timeout 600 python bench.py --config 5 --factor 1000 > gpurun_out/${T}_config5.json 2> gpurun_out/${T}_config5.err; show gpurun_out/${T}_config5.json; tail -3 gpurun_out/${T}_config5.err | cut -c1-300
