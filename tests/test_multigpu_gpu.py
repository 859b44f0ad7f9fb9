"""N-GPU run (one process per GPU, peer-memory exchange over NVLink, then NCCL collectives) against the 1-GPU run of
the same job: the final generator / discriminator states must be BIT-IDENTICAL (the only cross-GPU arithmetic is the
sum of the workers' feedbacks, performed in a fixed order).  Needs >= 2 GPUs; skipped on a one-GPU box (the protocol
itself is covered on CPU by tests/test_exchange_gloo.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("graph", [False, True])
def test_two_gpus_bit_identical_to_one(exchange, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MDGAN_EXCHANGE=exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(REPO / "tools" / "multigpu_check.py"), "--workers", "4"]
    if graph:
        cmd.append("--graph")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "BIT-IDENTICAL" in out.stdout, out.stdout[-2000:]
