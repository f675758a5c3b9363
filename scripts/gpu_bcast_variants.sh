#!/bin/bash
# Broadcast-plan variants of the config-2 bench on N GPUs: copy engines pushing (K chunks), or a fetch kernel pulling
# the peers' regions on R reserved SMs.   gpurun --gpus N -- 'bash scripts/gpu_bcast_variants.sh N tag'
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
one ce1 B200_BCAST_CHUNKS=1
one ce1g B200_BCAST_CHUNKS=1 B200_MULTI_GRAPH=1
one pull12 B200_BCAST=pull B200_BCAST_SMS=12
one pull12g B200_BCAST=pull B200_BCAST_SMS=12 B200_MULTI_GRAPH=1
one pull6g B200_BCAST=pull B200_BCAST_SMS=6 B200_MULTI_GRAPH=1
one pull20g B200_BCAST=pull B200_BCAST_SMS=20 B200_MULTI_GRAPH=1
one pull12k2g B200_BCAST=pull B200_BCAST_SMS=12 B200_BCAST_CHUNKS=2 B200_MULTI_GRAPH=1
