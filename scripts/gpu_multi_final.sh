#!/bin/bash
# N x B200 (N = 4): the 2-GPU tests over real peers, config 2 at 2 and N GPUs, config 4 at N GPUs (per-GPU load of the
# full config).   gpurun --gpus 4 -- 'bash scripts/gpu_multi_final.sh 4 r2q'
set -u
N=${1:-4}; T=${2:-r2q}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/${T}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_multi.log)
tail -3 gpurun_out/${T}_pytest_multi.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), r.get("per_kernel_ms") or r.get("per_kernel_ms_rank0"), d.get("checksum_ok"),
          d.get("rows_received_max_over_mean"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for n in 2 $N; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 50 --warmup 5 --no-e2e \
      > gpurun_out/${T}_bench_n$n.json 2> gpurun_out/${T}_bench_n$n.err
  show gpurun_out/${T}_bench_n$n.json; tail -2 gpurun_out/${T}_bench_n$n.err | cut -c1-200
done
timeout 400 $TR --nproc-per-node $N --master-port 29650 bench.py --config 4 --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${T}_config4_n$N.json 2> gpurun_out/${T}_config4_n$N.err
show gpurun_out/${T}_config4_n$N.json; tail -2 gpurun_out/${T}_config4_n$N.err | cut -c1-200
