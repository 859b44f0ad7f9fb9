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


def test_recorded_bench_line_meets_the_contract():
    """The JSON line `bench.py` printed on a B200 for the final tree (profiles/r02_bench_celeba_k1_final.json, copied from
    gpurun_out/) carries every key of the measurement contract, and its derived fields are consistent."""
    path = REPO / "profiles" / "r02_bench_celeba_k1_final.json"
    d = json.loads(path.read_text().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "CelebA-shape" in d["config"]["workload"] and "model" not in d["config"]
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["traffic"] is not None
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["unit"] == d["unit"] and "UNMODIFIED reference" in c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] <= d["value"] * 1.02 and e["value"] != d["value"], "e2e is measured, not copied from the device leg"
    assert e["value"] / c["value"] > 50, "end-to-end against the reference on the same box's host cores"
    k = d["clocks"]
    assert k["sm_mhz"] > 0.9 * k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"] > 0
    assert {"MNIST_DCGAN_b64", "CIFAR10_b64", "MNIST_MLP_b64"} <= set(d["shapes"]) and "error" not in d["shapes"]["MNIST_MLP_b64"]
