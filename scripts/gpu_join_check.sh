#!/bin/bash
# 1 x B200 after a change to the join kernel: every GPU test, config 5 per query (x100, default against the chained
# 64-bit table), and the config-2 line for both key widths.
set -u
T=${1:-r2w}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -4 gpurun_out/${T}_pytest.log
python scripts/exp_config5.py 100 "" B200_TAG64=0 > gpurun_out/${T}_exp_config5.log 2>&1; grep -A6 "^===" gpurun_out/${T}_exp_config5.log | cut -c1-220
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 200 python bench.py --steps 50 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; show gpurun_out/${T}_bench.json
B200_FORCE_KEY64=1 timeout 200 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_key64.json 2> gpurun_out/${T}_bench_key64.err; show gpurun_out/${T}_bench_key64.json
timeout 300 python bench.py --config 5 --factor 100 > gpurun_out/${T}_config5_x100.json 2> gpurun_out/${T}_config5_x100.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_config5_x100.json')); print({w: r['batch_seconds'] for w, r in d['runs_by_workers'].items()})"
