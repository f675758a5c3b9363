#!/bin/bash
# ncu evidence for the config-2 step on one B200 (B200_PROFILING.md recipe): the plain command first, then the launch
# list (cold-cache, serialised: compare SHARES), then --set full of the scatter and join kernels.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_profile.sh r2'
set -u
T=${1:-r2}
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${T}_plain.log 2> gpurun_out/${T}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD \
    > gpurun_out/${T}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'radix_scatter_kernel|tag_join_kernel|radix_scatter_pay|radix_hist' \
    -s 8 -c 8 -o gpurun_out/prof_${T} -f $CMD > gpurun_out/${T}_ncu_f.log 2>&1
ls -la gpurun_out/prof_${T}.ncu-rep gpurun_out/${T}_launches.csv
tail -3 gpurun_out/${T}_ncu_f.log
