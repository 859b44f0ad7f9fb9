#!/usr/bin/env python
"""Run the UNMODIFIED reference (/root/reference/src) on CPU in this container.

Test infrastructure only (see oracle/README.md).  Used to (1) pin the oracle
restatement (oracle/mdgan_oracle.py) against the reference's real N+1-process
gloo run, and (2) generate the fixtures under tests/golden/.  /root/reference is
not present on the GPU box, so nothing that runs there may import this file.

    python run_reference.py distributed --dataset CIFAR10 --workers 2 --batch_size 8 \
        --epochs 3 --swap_interval 100 --out /tmp/ref_run
    python run_reference.py standalone --dataset CIFAR10 --batch_size 8 --epochs 3 --out /tmp/ref_sa

The reference's scripts write `logs/`, `weights/`, `saved_images/` relative to
the cwd, so each run happens in a scratch cwd (`--out`).
"""
import argparse
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("MDGAN_REFERENCE_SRC", "/root/reference/src"))


def _env(m: int, threads: int) -> dict:
    env = dict(os.environ)
    env["PYTHONPATH"] = f"{HERE / 'stubs'}:{REF_SRC}:" + env.get("PYTHONPATH", "")
    env["MDGAN_SYNTH_M"] = str(m)
    env["OMP_NUM_THREADS"] = str(threads)
    env["MKL_NUM_THREADS"] = str(threads)
    return env


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["distributed", "standalone"])
    ap.add_argument("--dataset", default="CIFAR10")
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--batch_size", type=int, default=8)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--swap_interval", type=int, default=10**6)
    ap.add_argument("--local_epochs", type=int, default=1)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--iid", type=int, default=1)
    ap.add_argument("--beta_1", type=float, default=0.5)
    ap.add_argument("--beta_2", type=float, default=0.999)
    ap.add_argument("--lr", type=float, default=0.0002)
    ap.add_argument("--samples", type=int, default=0, help="dataset size M (default workers*16*batch)")
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--master_port", default="12377")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    if not REF_SRC.is_dir():
        print(f"reference sources not found at {REF_SRC}", file=sys.stderr)
        return 2
    out = Path(a.out)
    (out / "logs").mkdir(parents=True, exist_ok=True)  # standalone_gan.py opens logs/ before mkdir
    m = a.samples or max(a.workers, 1) * 16 * a.batch_size
    env = _env(m, a.threads)
    if a.mode == "distributed":
        ws = a.workers + 1
        cmd = [sys.executable, str(REF_SRC / "bootstrap.py"), "--backend", "gloo",
               "--world_size", str(ws), "--dataset", a.dataset, "--ranks", f"0..{a.workers}",
               "--epochs", str(a.epochs), "--local_epochs", str(a.local_epochs),
               "--swap_interval", str(a.swap_interval), "--device", "cpu",
               "--batch_size", str(a.batch_size), "--iid", str(a.iid), "--seed", str(a.seed),
               "--master_addr", "127.0.0.1", "--master_port", a.master_port,
               "--log_interval", str(10**9), "--generator_lr", str(a.lr),
               "--discriminator_lr", str(a.lr), "--beta_1", str(a.beta_1), "--beta_2", str(a.beta_2)]
    else:
        cmd = [sys.executable, str(REF_SRC / "standalone_gan.py"), "--dataset", a.dataset,
               "--epochs", str(a.epochs), "--local_epochs", str(a.local_epochs),
               "--batch_size", str(a.batch_size), "--log_interval", str(10**9),
               "--generator_lr", str(a.lr), "--discriminator_lr", str(a.lr), "--device", "cpu",
               "--seed", str(a.seed), "--beta_1", str(a.beta_1), "--beta_2", str(a.beta_2)]
    r = subprocess.run(cmd, cwd=out, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    (out / "run.log").write_text(r.stdout)
    if r.returncode != 0:
        print(r.stdout[-4000:], file=sys.stderr)
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
