"""The checked build (make -C sigmod-2018_b200/csrc checked: the same sources with -DB200_CHECKED, i.e. device-side
bounds checks on stage buffers, tag tables, warp queues and output positions, common.cuh B200_DCHECK) over the subset
of the GPU tests that exercises the hand-rolled shared-memory structures.  compute-sanitizer is closed on the GPU pool
this repository is developed on (profiles/r2_sanitizer_unavailable.txt); a failed device assertion aborts the child
process (cudaErrorAssert is fatal in the library), so exit code 0 means no check fired and every result matched the
oracle."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
CHECKED = ROOT / "sigmod-2018_b200" / "lib" / "libb200join_checked.so"

SUBSET = ("test_join_pairs_shapes and 100000 or test_join_sum_zipf_probe_side and 200000 or test_join_sum_64bit_keys and 12-16 "
          "or test_carried_payload_may_hold_all_ones or test_join_sum_config2_shape_scaled_down and 12-16 "
          "or test_radix_partition_matches_oracle and 300007 or test_broadcast_plan_emulated_ranks and 70001 "
          "or test_broadcast_plan_pull_variant and 2-0 or test_exchange_plan_emulated_ranks and 4-15-9 "
          "or test_exchange_plan_hot_keys_with_duplicate_build_keys or test_fused_filters_match_the_oracle and 999 "
          "or test_small_workload_execute_query")


def test_subset_passes_on_the_checked_build():
    assert CHECKED.exists(), f"{CHECKED} is missing: make -C sigmod-2018_b200/csrc checked (or __graft_entry__.build())"
    env = dict(os.environ, B200_LIB=str(CHECKED))
    p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-s", "-p", "no:cacheprovider",
                        "tests/test_kernels_gpu.py", "tests/test_multi_plan_gpu.py", "tests/test_filter_fusion_gpu.py",
                        "tests/test_operators_gpu.py", "-k", SUBSET],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and "failed" not in p.stdout, tail
