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


@pytest.mark.gpu
def test_c_host_answers_a_config2_query_on_two_gpus(tmp_path):
    """host/b200_engine -g 2: relations sharded over the GPUs when they are loaded, the join through the library's
    multi-GPU plan from ONE process (a host thread per GPU, peer access, no NCCL) — the C side of SURVEY §8e — and
    the other queries of the batch on GPU 0 as before.  Lines equal the oracle executor's."""
    if _gpu_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import orc
    import refbind
    nr, ns = 1 << 18, (1 << 22) + 6
    r = [orc.synth_column(nr, 0, 18, 11), orc.synth_column(nr, 1, 0, 12)]
    s = [orc.synth_column(ns, 0, 21, 13), orc.synth_column(ns, 1, 0, 14), orc.synth_column(ns, 3, 100, 15)]
    refbind.write_relation_file(tmp_path / "r0", r)
    refbind.write_relation_file(tmp_path / "r1", s)
    queries = ["0 1|0.0=1.0|0.1 1.1", "1 0|0.0=1.0|1.1", "0 1|0.0=1.0&1.2<50|0.1 1.1", "0 1|0.0=1.0|1.1 0.1 1.2"]
    stdin = "r0\nr1\nDone\n" + "\n".join(queries) + "\nF\n"
    out = subprocess.run([str(ROOT / "host" / "b200_engine"), "-g", "2"], input=stdin, capture_output=True, text=True,
                         cwd=tmp_path, timeout=300, env=dict(__import__("os").environ, B200_TIMING="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.splitlines() == [orc.execute_query(q, [r, s]) for q in queries]
    assert out.stderr.count("multi-GPU join on 2 GPUs") == 2          # the two config-2-shaped queries
