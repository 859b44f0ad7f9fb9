"""The reference arm of the benchmark (CPU only): the unmodified reference installed under oracle/_ref runs as N+1 gloo
processes through oracle/ref_harness/time_reference.py, and `bench.py --impl reference` prints the contract's JSON
line.  Skipped when neither oracle/_ref nor /root/reference is present."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle.ref_harness import install_ref, time_reference  # noqa: E402

needs_ref = pytest.mark.skipif(time_reference.reference_src() is None, reason="no reference sources on this host")


@needs_ref
def test_installed_reference_is_unmodified():
    if not (install_ref.DEST / "MANIFEST.json").exists():
        pytest.skip("oracle/_ref not installed (build() installs it when /root/reference is present)")
    assert install_ref.verify()
    mf = json.loads((install_ref.DEST / "MANIFEST.json").read_text())["files"]
    assert {"bootstrap.py", "actors/server.py", "actors/worker.py", "datasets/CIFAR10.py", "datasets/CelebA.py"} <= set(mf)
    assert all(rec["sha256_source"] == rec["sha256_copy"] for rec in mf.values())


@needs_ref
def test_reference_arm_json_line():
    """K = 2 workers (3 processes), CIFAR-10 shape, b = 4, a swap every iteration: the reference's own loop."""
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2", "--dataset",
                          "CIFAR10", "--batch", "4", "--steps", "3", "--warmup", "1", "--swap-interval", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "reference" and "3 processes" in d["cpu_baseline"]["sample"]
    assert d["value"] > 0 and d["unit"] == "worker-it/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["swap_interval"] == 1 and "swap every 1" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    import os

    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""
