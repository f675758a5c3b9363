#!/bin/bash
# 1 x B200: every GPU test, then config 4 (one GPU, the full config's per-GPU load) with 2^12 partitions, probe chunks
# partitioned in one pass (B200_TWO_PASS_BITS=99) and in two
set -u
T=${1:-r2n}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -4 gpurun_out/${T}_pytest.log
for v in 99 11; do
  B200_TWO_PASS_BITS=$v timeout 300 python bench.py --config 4 --gpus 1 --steps 10 --radix-bits 12 > gpurun_out/${T}_config4_n1_tp$v.json 2> gpurun_out/${T}_config4_n1_tp$v.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_config4_n1_tp$v.json')); print('two-pass from $v bits:', d['ms_per_step'], d['roofline']['per_kernel_ms_rank0'], d['checksum_ok'])"
  tail -2 gpurun_out/${T}_config4_n1_tp$v.err | cut -c1-200
done
