#!/usr/bin/env python
"""Time the UNMODIFIED reference (oracle/_ref/src, or /root/reference/src in the build container) on the host CPU
cores: `bootstrap.py --backend gloo --device cpu` as N+1 OS processes (one server + N workers, mp.spawn), synthetic
data, `OMP_NUM_THREADS = floor(cores / (N+1))` -- SURVEY.md section 8(d), BASELINE.md section 2.  The per-iteration
time is the server's own `end.epoch_calculation - start.epoch_calculation` from the CSV the reference writes.

Measurement infrastructure only (used by bench.py --impl reference / cpu_baseline and by the CPU tests).
"""
import csv
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_INSTALLED = HERE.parent / "_ref" / "src"


def reference_src():
    """Where the unmodified reference sources are: $MDGAN_REFERENCE_SRC, the installed copy, or /root/reference."""
    env = os.environ.get("MDGAN_REFERENCE_SRC")
    for cand in ([Path(env)] if env else []) + [REF_INSTALLED, Path("/root/reference/src")]:
        if (cand / "bootstrap.py").exists():
            return cand
    return None


def _free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def time_distributed(dataset: str, n_workers: int, batch: int, steps: int, warmup: int, swap_interval: int = 10 ** 9,
                     threads: int = 0, seed: int = 3, timeout_s: float = 800.0) -> dict:
    """Runs warmup + steps generator iterations of the reference and returns
    {"ms_per_step", "steps", "warmup", "cores", "threads_per_process", "processes", "phases_ms": {...}, "wall_s"}."""
    src = reference_src()
    if src is None:
        raise FileNotFoundError("no reference sources (oracle/_ref not installed and /root/reference absent)")
    cores = os.cpu_count() or 1
    procs = n_workers + 1
    threads = threads or max(1, cores // procs)
    epochs = warmup + steps
    out = Path(tempfile.mkdtemp(prefix="mdgan_ref_"))
    (out / "logs").mkdir()
    env = dict(os.environ)
    env["PYTHONPATH"] = f"{HERE / 'stubs'}:{src}:" + env.get("PYTHONPATH", "")
    env["MDGAN_SYNTH_M"] = str(n_workers * 16 * batch)
    env["OMP_NUM_THREADS"] = env["MKL_NUM_THREADS"] = str(threads)
    env["CUDA_VISIBLE_DEVICES"] = ""     # the reference's CPU path: its evaluation device must not grab a GPU
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID",
              "GROUP_RANK", "LOCAL_WORLD_SIZE", "ROLE_RANK", "ROLE_WORLD_SIZE"):
        env.pop(k, None)                 # bootstrap.py does its own rendezvous (mp.spawn)
    # an even world size (odd K, e.g. the single-GPU configuration K = 1) is refused by the reference's CLI guard only
    # (bootstrap.py:163-164): launch_any_workers.py repeats the launcher's tail without it
    entry = src / "bootstrap.py" if procs % 2 == 1 else HERE / "launch_any_workers.py"
    cmd = [sys.executable, str(entry), "--backend", "gloo", "--world_size", str(procs),
           "--dataset", dataset, "--ranks", f"0..{n_workers}", "--epochs", str(epochs), "--local_epochs", "1",
           "--swap_interval", str(swap_interval), "--device", "cpu", "--batch_size", str(batch), "--iid", "1",
           "--seed", str(seed), "--master_addr", "127.0.0.1", "--master_port", str(_free_port()),
           "--log_interval", str(10 ** 9), "--generator_lr", "0.0002", "--discriminator_lr", "0.0002",
           "--beta_1", "0.5", "--beta_2", "0.999"]
    t0 = time.time()
    try:
        r = subprocess.run(cmd, cwd=out, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           timeout=timeout_s)
        if r.returncode != 0:
            raise RuntimeError("reference run failed:\n" + r.stdout[-3000:])
        rows = list(csv.DictReader(open(out / "logs" / f"mdgan.{n_workers}.{dataset}.server.logs.csv")))
    finally:
        wall = time.time() - t0
        shutil.rmtree(out, ignore_errors=True)
    if len(rows) != epochs:
        raise RuntimeError(f"reference wrote {len(rows)} rows, expected {epochs}")
    timed = rows[warmup:]
    dur = [float(x["end.epoch_calculation"]) - float(x["start.epoch_calculation"]) for x in timed]

    def phase(name):
        vals = [float(x[f"end.{name}"]) - float(x[f"start.{name}"]) for x in timed
                if x.get(f"end.{name}") not in (None, "", "None") and x.get(f"start.{name}") not in (None, "", "None")]
        return 1e3 * sum(vals) / len(vals) if vals else None

    return {"ms_per_step": 1e3 * sum(dur) / len(dur), "steps": len(dur), "warmup": warmup, "cores": cores,
            "threads_per_process": threads, "processes": procs, "wall_s": wall, "source": str(src),
            "launcher": entry.name,
            "phases_ms": {n: phase(n) for n in ("generate_data", "send_data", "recv_data", "agg_gradients",
                                                "calc_gradients")}}


if __name__ == "__main__":
    import json

    a = sys.argv[1:]
    print(json.dumps(time_distributed(a[0] if a else "CIFAR10", int(a[1]) if len(a) > 1 else 1,
                                      int(a[2]) if len(a) > 2 else 8, int(a[3]) if len(a) > 3 else 3, 1)))
