#!/bin/bash
# compute-sanitizer over a small subset of the GPU tests (ONE tool per gpurun call, B200_PROFILING.md):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh memcheck'
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh racecheck'
# The subset covers the hand-rolled pieces: warp queues with inline LDS/STS and __syncwarp in the join kernels,
# shared-memory CAS chains, the tile pipeline of the scatter (join pairs shapes, Zipf overflow, 64-bit keys, carried
# payloads), the multi-GPU plans with emulated ranks (flags, hot keys, the TMA fetch kernel) and fused filters.
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='test_join_pairs_shapes and 100000 or test_join_sum_zipf_probe_side and 200000 or test_join_sum_64bit_keys and 12-16 or test_carried_payload_may_hold_all_ones or test_join_sum_config2_shape_scaled_down and 12-16 or test_scan_filter_base_column and 100003 or test_radix_partition_matches_oracle and 300007 or test_column_stats or test_broadcast_plan_emulated_ranks and 70001 or test_broadcast_plan_pull_variant and 2-0 or test_exchange_plan_emulated_ranks and 4-15-9 or test_exchange_plan_hot_keys_with_duplicate_build_keys or test_fused_filters_match_the_oracle and 999'
timeout 1000 compute-sanitizer --tool $TOOL --error-exitcode 97 --log-file gpurun_out/sanitizer_${TOOL}.log \
    python -m pytest tests/test_kernels_gpu.py tests/test_multi_plan_gpu.py tests/test_filter_fusion_gpu.py -q -x -k "$SEL" > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
echo "sanitizer $TOOL rc=$?" | tee -a gpurun_out/sanitizer_${TOOL}_pytest.log
tail -5 gpurun_out/sanitizer_${TOOL}_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Race" gpurun_out/sanitizer_${TOOL}.log | sort | uniq -c | head -20
