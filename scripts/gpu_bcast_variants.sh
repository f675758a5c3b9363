#!/bin/bash
# Broadcast-plan variants of the config-2 bench on N GPUs: SM broadcast kernel with R reserved SMs, copy engines with
# K chunks.   gpurun --gpus N -- 'bash scripts/gpu_bcast_variants.sh N tag'
set -u
N=${1:-2}; T=${2:-r2v}
mkdir -p gpurun_out
port=29700
one() {  # name, env...
  name=$1; shift
  port=$((port + 1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 40 --warmup 5 --no-e2e > gpurun_out/${T}_n${N}_$name.json 2> gpurun_out/${T}_n${N}_$name.err
  python - gpurun_out/${T}_n${N}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
one sm16 B200_BCAST_SMS=16
one sm8 B200_BCAST_SMS=8
one sm4 B200_BCAST_SMS=4
one sm2 B200_BCAST_SMS=2
one sm8g B200_BCAST_SMS=8 B200_MULTI_GRAPH=1
one ce1 B200_BCAST_CE=1 B200_BCAST_CHUNKS=1
one ce2 B200_BCAST_CE=1 B200_BCAST_CHUNKS=2
one ce1g B200_BCAST_CE=1 B200_BCAST_CHUNKS=1 B200_MULTI_GRAPH=1
