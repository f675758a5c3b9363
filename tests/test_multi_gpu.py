"""N > 1 on real GPUs: runs only where the box has at least two (the driver's 1-GPU round-end tier skips it;
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` is how it is exercised)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
    except (OSError, subprocess.TimeoutExpired):
        return 0
    return sum(1 for line in out.splitlines() if line.startswith("GPU "))


@pytest.mark.gpu
def test_multi_gpu_plans_match_the_oracle():
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI-GPU PARITY OK" in r.stdout
