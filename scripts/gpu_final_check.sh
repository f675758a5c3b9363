#!/bin/bash
# Round-end check on a 1 x B200 box (about 10 minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_final_check.sh r2z'
# every GPU test (incl. the checked build), smoke, the default bench line (e2e + cpu_baseline at full size), config 3 at
# full size with the reference timed beside it, config 5 at x100 (diffed against the reference binary) and x1000.
set -u
T=${1:-r2z}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -4 gpurun_out/${T}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get("roofline") or {}
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "value", f'{d["value"]:.4g}', d["unit"],
          {k: round(v, 4) for k, v in (r.get("per_kernel_ms") or {}).items()}, "frac", r.get("frac"),
          "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
          "ok", d.get("checksum_ok", d.get("output_identical_across_worker_counts")), d.get("parity_vs_reference"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 400 python bench.py --steps 50 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; show gpurun_out/${T}_bench.json
timeout 300 python bench.py --config 3 --steps 20 > gpurun_out/${T}_config3.json 2> gpurun_out/${T}_config3.err; show gpurun_out/${T}_config3.json
B200_FUSE_FILTERS=0 timeout 200 python bench.py --config 3 --steps 20 --no-cpu-baseline > gpurun_out/${T}_config3_eager.json 2> gpurun_out/${T}_config3_eager.err; show gpurun_out/${T}_config3_eager.json
timeout 400 python bench.py --config 5 --factor 100 --check-reference > gpurun_out/${T}_config5_x100.json 2> gpurun_out/${T}_config5_x100.err; show gpurun_out/${T}_config5_x100.json
timeout 500 python bench.py --config 5 --factor 1000 > gpurun_out/${T}_config5.json 2> gpurun_out/${T}_config5.err; show gpurun_out/${T}_config5.json
tail -3 gpurun_out/${T}_config5.err
timeout 300 python bench.py --config 4 --gpus 1 --steps 10 > gpurun_out/${T}_config4_n1.json 2> gpurun_out/${T}_config4_n1.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_config4_n1.json')); print('config4 n1', d['ms_per_step'], d['roofline']['per_kernel_ms_rank0'], d['checksum_ok'])"
