#!/bin/bash
# One-call GPU check for a 1 x B200 box (about 4 minutes):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/gpu_round_check.sh'
# GPU tests, smoke, the default bench line, and the bench under the tuning knobs worth re-checking after a
# kernel change (scatter tile configurations, carried probe column on/off).  Everything lands in gpurun_out/.
set -u
T=${1:-r2}
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -4 gpurun_out/${T}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["per_kernel_ms"].items()},
      "hbm frac", round(d["hbm"]["frac_of_peak"], 4), d.get("clocks"))
PY
}
timeout 300 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err && show gpurun_out/${T}_bench.json
for cfg in 0 2 3; do
    B200_SCATTER_CFG=$cfg timeout 120 python bench.py --steps 50 --no-e2e --no-cpu-baseline \
        > gpurun_out/${T}_bench_cfg$cfg.json 2> gpurun_out/${T}_bench_cfg$cfg.err && show gpurun_out/${T}_bench_cfg$cfg.json
done
B200_CARRY_PROBE=0 timeout 120 python bench.py --steps 50 --no-e2e --no-cpu-baseline \
    > gpurun_out/${T}_bench_nocarry.json 2> gpurun_out/${T}_bench_nocarry.err && show gpurun_out/${T}_bench_nocarry.json
