#!/bin/bash
# The 8-GPU box in one call: config-2 bench lines at N = 8, 4, 2 (plain launches and the step captured in a CUDA
# graph) and config 4 at FULL size on 8 GPUs.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale_check.sh tag'
set -u
T=${1:-r2s}
mkdir -p gpurun_out
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
port=29600
for N in 8 4; do
  for g in 1 0; do
    port=$((port + 1))
    B200_MULTI_GRAPH=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus $N --steps 50 --warmup 5 --no-e2e \
        > gpurun_out/${T}_bench_n${N}_graph$g.json 2> gpurun_out/${T}_bench_n${N}_graph$g.err
    show gpurun_out/${T}_bench_n${N}_graph$g.json
  done
done
timeout 60 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
show gpurun_out/${T}_bench_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --config 4 --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_config4_n8.json 2> gpurun_out/${T}_config4_n8.err
show gpurun_out/${T}_config4_n8.json; tail -3 gpurun_out/${T}_config4_n8.err
