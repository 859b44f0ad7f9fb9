#!/usr/bin/env python
"""Generate tests/golden/*.pt from the UNMODIFIED reference, run in the build container (CPU, gloo).

    python tests/golden/make_golden.py [case ...]        # default: every case

For each case of oracle/ref_harness/pin_oracle.py::CASES this runs the reference's own bootstrap.py /
standalone_gan.py (N+1 processes, synthetic data), checks that the oracle restatement reproduces it, and stores a
compact fixture of the REFERENCE's outputs:
  * per-iteration mean_d_loss of every worker and the swap_with log (from the reference's CSV logs);
  * for every tensor of the final generator / discriminator state_dicts: shape, dtype, sum, sum of |x|, and every
    STRIDE-th element (the full checkpoints are 3-14 MB each; the sample keeps a fixture under ~100 kB).
The fixtures are consumed by tests/test_oracle_golden.py, which re-runs the oracle with THIS repo's plugin model
definitions (distributed-gan_b200/datasets) -- pinning both the oracle and the plugin ports.  /root/reference is
needed only here, never at test time.
"""
import csv
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO / "oracle" / "ref_harness"))
import pin_oracle  # noqa: E402

STRIDE = 211


def summarise(sd):
    out = {}
    for k, v in sd.items():
        f = v.detach().reshape(-1)
        out[k] = {"shape": tuple(v.shape), "dtype": str(v.dtype), "sum": f.double().sum().item(),
                  "abssum": f.double().abs().sum().item(), "sample": f[::STRIDE].clone()}
    return out


def main():
    only = sys.argv[1:]
    for case in pin_oracle.CASES:
        pos, extra = pin_oracle.case_args(case)
        name, mode, dataset, workers, batch, epochs, swap_interval, seed = pos
        if only and name not in only:
            continue
        r = pin_oracle.run_case(*pos, **extra)
        assert r["G_maxdiff"] == 0.0 and r["D_maxdiff"] == 0.0 and r["loss_maxdiff"] < 1e-12 and r["swaps_bit_exact"], r
        out = Path(r["out"])
        fx = {"case": dict(name=name, mode=mode, dataset=dataset, workers=workers, batch=batch, epochs=epochs,
                           swap_interval=swap_interval, seed=seed, beta_1=0.5, samples=max(workers, 1) * 16 * batch, **extra),
              "stride": STRIDE, "torch": torch.__version__}
        if mode == "distributed":
            fx["G"] = summarise(torch.load(out / "weights" / "generator_final.pt"))
            fx["D"], fx["mean_d_loss"], fx["swap_with"] = [], [], []
            for n in range(workers):
                fx["D"].append(summarise(torch.load(out / "weights" / f"worker_{n + 1}" / "discriminator.pth")))
                rows = list(csv.DictReader(open(out / "logs" / f"mdgan.{workers}.{dataset}.worker.{n + 1}.logs.csv")))
                fx["mean_d_loss"].append([float(row["mean_d_loss"]) for row in rows])
                fx["swap_with"].append([int(row["swap_with"]) if row["swap_with"] != "" else None for row in rows])
        else:
            fx["G"] = summarise(torch.load(out / "weights" / f"netG_epoch_{epochs - 1}.pth"))
            fx["D"] = [summarise(torch.load(out / "weights" / f"netD_epoch_{epochs - 1}.pth"))]
            rows = list(csv.DictReader(open(out / "logs" / f"{dataset}.standalone.logs.csv")))
            fx["mean_d_loss"] = [[float(row["mean_d_loss"]) for row in rows]]
            fx["mean_g_loss"] = [[float(row["mean_g_loss"]) for row in rows]]
        torch.save(fx, HERE / f"{name}.pt")
        size = (HERE / f"{name}.pt").stat().st_size
        print(f"{name}: reference == oracle bit-exact; fixture {size / 1024:.0f} kB", flush=True)


if __name__ == "__main__":
    main()
