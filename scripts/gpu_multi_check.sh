#!/bin/bash
# Multi-GPU check on one box with N GPUs:
#   /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_multi_check.sh N tag'
# parity over real CUDA IPC + NVLink (tests/test_multi_gpu.py), the config-2 bench line at N (with and without a CUDA
# graph of the step), config 4 at the per-GPU load of the full config.  Everything lands in gpurun_out/.
set -u
N=${1:-2}; T=${2:-r2m}; STEPS4=${3:-10}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_multi.log)
tail -5 gpurun_out/${T}_pytest_multi.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get("roofline") or {}
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), r.get("per_kernel_ms") or r.get("per_kernel_ms_rank0"),
          "e2e", (d.get("e2e") or {}).get("ms_per_step"), d.get("rows_received_max_over_mean"), d.get("checksum_ok"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for g in 0 1; do
  B200_MULTI_GRAPH=$g timeout 300 $TR --master-port $((29500 + g)) bench.py --gpus $N --steps 50 --warmup 5 --no-e2e \
      > gpurun_out/${T}_bench_n${N}_graph$g.json 2> gpurun_out/${T}_bench_n${N}_graph$g.err
  show gpurun_out/${T}_bench_n${N}_graph$g.json; tail -3 gpurun_out/${T}_bench_n${N}_graph$g.err
done
timeout 600 $TR --master-port 29510 bench.py --config 4 --gpus $N --steps $STEPS4 --warmup 3 \
    > gpurun_out/${T}_config4_n${N}.json 2> gpurun_out/${T}_config4_n${N}.err
show gpurun_out/${T}_config4_n${N}.json; tail -3 gpurun_out/${T}_config4_n${N}.err
