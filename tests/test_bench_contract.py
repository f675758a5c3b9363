"""bench.py's output contract on the arm that runs without a GPU: `--impl reference` prints ONE JSON line
with the keys the driver reads (the GPU arm prints the same keys plus roofline/clocks; it is exercised on
the GPU box by the driver itself)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    if not (ROOT / "oracle" / "_ref" / "ref_driver").exists():
        pytest.skip("oracle/_ref/ref_driver not built (the port would take minutes at this sample size)")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--scale-bits", "4"],      # 1/16 of config 2: the full size takes a minute and 14 GB here
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "join_probe_throughput" and d["unit"] == "probe tuples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["checksum_ok"] is True          # the reference's printed sums were verified
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
