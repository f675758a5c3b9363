#!/bin/bash
# 1 x B200: multi-plan, checked-build and fused-filter tests, config 5 x100 per query with the host-side gap logger
# (three runs), config 4 at the per-GPU load of the full config on one GPU
set -u
T=${1:-r2s}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_multi_plan_gpu.py tests/test_checked_build_gpu.py tests/test_filter_fusion_gpu.py -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -4 gpurun_out/${T}_pytest.log
python scripts/exp_config5.py 100 B200_SLOWLOG=4 B200_SLOWLOG=4 B200_SLOWLOG=4 > gpurun_out/${T}_exp_config5.log 2>&1; grep -E "^===|slow:" gpurun_out/${T}_exp_config5.log | cut -c1-330 | head -60
timeout 300 python bench.py --config 4 --gpus 1 --steps 10 > gpurun_out/${T}_config4_n1.json 2> gpurun_out/${T}_config4_n1.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_config4_n1.json')); print('config4 n1', d['ms_per_step'], d['roofline']['per_kernel_ms_rank0'], d['checksum_ok'])"
for f in 100 1000; do
  timeout 500 python bench.py --config 5 --factor $f > gpurun_out/${T}_config5_x$f.json 2> gpurun_out/${T}_config5_x$f.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_config5_x$f.json')); print($f, d['ms_per_step'], {w: (r['batch_seconds'], r['startup'][-60:]) for w, r in d['runs_by_workers'].items()})"
done
