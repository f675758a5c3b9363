#!/bin/bash
set -u
T=${1:-r2v}
mkdir -p gpurun_out
python scripts/exp_config5.py 100 "" "" B200_TAG64=0 "" > gpurun_out/${T}_exp_config5.log 2>&1; grep -A8 "^===" gpurun_out/${T}_exp_config5.log | cut -c1-260
timeout 300 python bench.py --config 3 --steps 20 --no-cpu-baseline > gpurun_out/${T}_config3.json 2> gpurun_out/${T}_config3.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_config3.json')); print('config3', d['ms_per_step'], d['roofline']['per_kernel_ms'], d['checksum_ok'])"
