"""N-GPU run (one process per GPU, peer-memory exchange over NVLink, then NCCL collectives) against the 1-GPU run of
the same job: the final generator / discriminator states must be BIT-IDENTICAL (the only cross-GPU arithmetic is the
sum of the workers' feedbacks, performed in a fixed order).  Runs on min(8, device_count) GPUs with as many workers
(rounded up to even) and discriminator swaps; skipped on a one-GPU box -- there the same check is part of every
`bench.py --gpus N` run (`multi_gpu_bit_identical` in its JSON line) and the protocol itself is covered on CPU by
tests/test_exchange_gloo.py."""
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
def test_n_gpus_bit_identical_to_one(exchange, graph):
    nproc = min(8, torch.cuda.device_count())
    if nproc < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, MDGAN_EXCHANGE=exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(REPO / "tools" / "multigpu_check.py"), "--workers",
           str(max(4, nproc + (nproc & 1)))]
    if graph:
        cmd.append("--graph")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "BIT-IDENTICAL" in out.stdout, out.stdout[-2000:]
