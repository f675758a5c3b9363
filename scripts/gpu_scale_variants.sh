#!/bin/bash
# 8-GPU box: push (copy engines) against pull (fetch kernel) broadcast at N = 8 and 4, step captured in a CUDA graph.
set -u
T=${1:-r2k}
mkdir -p gpurun_out
port=29900
one() {  # N name env...
  N=$1; name=$2; shift 2
  port=$((port + 1))
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 50 --warmup 5 --no-e2e > gpurun_out/${T}_n${N}_$name.json 2> gpurun_out/${T}_n${N}_$name.err
  python - gpurun_out/${T}_n${N}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
one 8 ce1g B200_MULTI_GRAPH=1
one 8 pull12g B200_BCAST=pull B200_BCAST_SMS=12 B200_MULTI_GRAPH=1
one 8 pull16g B200_BCAST=pull B200_BCAST_SMS=16 B200_MULTI_GRAPH=1
one 8 pull8g B200_BCAST=pull B200_BCAST_SMS=8 B200_MULTI_GRAPH=1
one 4 pull12g B200_BCAST=pull B200_BCAST_SMS=12 B200_MULTI_GRAPH=1
