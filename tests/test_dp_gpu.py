"""Real multi-rank data parallelism on GPUs (NCCL): spawns tests/dp_worker.py under torchrun when at least two
GPUs are visible (skipped on a one-GPU box; the host-side logic is covered on CPU with gloo in
tests/test_host_cpu.py).  A 2-GPU run of this file is committed under profiles/ (r2_dp_2gpu.log)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_n_rank_gradients_equal_oracle_per_shard_average():
    n = 2 if torch.cuda.device_count() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-6000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, "data-parallel worker failed"
    assert "all data-parallel parity cases passed" in r.stdout
