#!/usr/bin/env python
"""Pin oracle/mdgan_oracle.py against the unmodified reference, run here on CPU.

For each case: (1) run /root/reference/src/bootstrap.py (or standalone_gan.py) through
run_reference.py (N+1 gloo processes, synthetic data, 1 thread per process); (2) run
the oracle restatement in this process with the reference's OWN model classes
(imported from /root/reference/src/datasets/<NAME>.py) and the same flags; (3) compare
final state_dicts, per-iteration mean_d_loss and the swap log.  Build-container only
(/root/reference does not exist on the GPU box).  Test infrastructure only.

    python oracle/ref_harness/pin_oracle.py            # all cases, prints a table
"""
import csv
import importlib
import os
import sys
import tempfile
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF_SRC = Path(os.environ.get("MDGAN_REFERENCE_SRC", "/root/reference/src"))
sys.path.insert(0, str(HERE / "stubs"))
sys.path.insert(0, str(REF_SRC))
sys.path.insert(0, str(REPO))

import run_reference  # noqa: E402
import synthetic_tv  # noqa: E402
from oracle.mdgan_oracle import OracleMDGAN, OracleStandalone  # noqa: E402

CASES = [
    # name, mode, dataset, workers, batch, epochs, swap_interval, seed
    ("cifar_n2", "distributed", "CIFAR10", 2, 8, 4, 10**6, 3),
    ("cifar_n4_swap", "distributed", "CIFAR10", 4, 4, 5, 2, 3),
    ("celeba_n2", "distributed", "CelebA", 2, 4, 3, 10**6, 3),
    ("mnist_n2", "distributed", "MNIST", 2, 8, 3, 10**6, 3),
    ("cifar_standalone", "standalone", "CIFAR10", 0, 8, 3, 0, 1),
    # the MLP plugin (always-on dropout drawn from each actor's global RNG stream) with swaps, and standalone (ONE stream
    # shared by the loader seeds, the noise and the dropout draws)
    ("mnist_n4_swap", "distributed", "MNIST", 4, 4, 4, 2, 3),
    ("mnist_standalone", "standalone", "MNIST", 0, 8, 4, 0, 1),
    # optional 9th field: extra launcher flags -- two local epochs per iteration, the non-iid (contiguous) split, a swap
    ("cifar_n2_le2_noniid", "distributed", "CIFAR10", 2, 4, 3, 2, 3, {"local_epochs": 2, "iid": 0}),
]


def case_args(case):
    """(the 8 positional fields, the extras dict) of a CASES entry."""
    return tuple(case[:8]), dict(case[8]) if len(case) > 8 else {}


def _maxdiff(sd_a, sd_b):
    worst = 0.0
    for k in sd_a:
        a, b = sd_a[k].double(), sd_b[k].double()
        worst = max(worst, (a - b).abs().max().item())
    return worst


def run_case(name, mode, dataset, workers, batch, epochs, swap_interval, seed, keep=None, local_epochs=1, iid=1):
    torch.set_num_threads(1)
    out = Path(keep) if keep else Path(tempfile.mkdtemp(prefix=f"ref_{name}_"))
    m = max(workers, 1) * 16 * batch
    argv = [mode, "--dataset", dataset, "--workers", str(workers), "--batch_size", str(batch),
            "--epochs", str(epochs), "--seed", str(seed), "--out", str(out), "--threads", "1",
            "--beta_1", "0.5"]
    if mode == "distributed":
        argv += ["--swap_interval", str(swap_interval), "--local_epochs", str(local_epochs), "--iid", str(iid)]
    sys.argv = ["run_reference.py"] + argv
    rc = run_reference.main()
    assert rc == 0, f"reference run failed for {name}"

    os.environ["MDGAN_SYNTH_M"] = str(m)
    ds = getattr(synthetic_tv, dataset)()
    mod = importlib.import_module(f"datasets.{dataset}")
    res = {"name": name, "out": str(out)}
    if mode == "distributed":
        o = OracleMDGAN(mod.Generator, mod.Discriminator, ds, workers, batch, mod.Z_DIM, mod.SHAPE,
                        seed=seed, beta_1=0.5, swap_interval=swap_interval, local_epochs=local_epochs, iid=bool(iid))
        d_losses, pairs_log = [], []
        for e in range(epochs):
            r = o.step(e, record=False)
            d_losses.append(r["mean_d_loss"])
            pairs_log.append(r["pairs"])
        ref_g = torch.load(out / "weights" / "generator_final.pt")
        res["G_maxdiff"] = _maxdiff(ref_g, o.G.state_dict())
        worst_d, worst_loss, swaps_ok = 0.0, 0.0, True
        for n in range(workers):
            ref_d = torch.load(out / "weights" / f"worker_{n+1}" / "discriminator.pth")
            worst_d = max(worst_d, _maxdiff(ref_d, o.D[n].state_dict()))
            rows = list(csv.DictReader(open(out / "logs" / f"mdgan.{workers}.{dataset}.worker.{n+1}.logs.csv")))
            for e, row in enumerate(rows):
                worst_loss = max(worst_loss, abs(float(row["mean_d_loss"]) - d_losses[e][n]))
                ref_sw = row["swap_with"]
                if pairs_log[e] is None:
                    swaps_ok &= ref_sw == ""
                else:
                    partner = {a: c for a, c in pairs_log[e].tolist()}
                    partner.update({c: a for a, c in pairs_log[e].tolist()})
                    swaps_ok &= ref_sw != "" and int(ref_sw) == partner[n + 1]
        res.update(D_maxdiff=worst_d, loss_maxdiff=worst_loss, swaps_bit_exact=swaps_ok)
        res["oracle"] = o
    else:
        o = OracleStandalone(mod.Generator, mod.Discriminator, ds, batch, mod.Z_DIM, seed=seed, beta_1=0.5)
        losses = [o.step() for _ in range(epochs)]
        ref_g = torch.load(out / "weights" / f"netG_epoch_{epochs-1}.pth")
        ref_d = torch.load(out / "weights" / f"netD_epoch_{epochs-1}.pth")
        res["G_maxdiff"] = _maxdiff(ref_g, o.G.state_dict())
        res["D_maxdiff"] = _maxdiff(ref_d, o.D.state_dict())
        rows = list(csv.DictReader(open(out / "logs" / f"{dataset}.standalone.logs.csv")))
        res["loss_maxdiff"] = max(abs(float(r["mean_d_loss"]) - l["mean_d_loss"]) for r, l in zip(rows, losses))
        res["swaps_bit_exact"] = True
        res["oracle"] = o
    return res


if __name__ == "__main__":
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    for case in CASES:
        if only and case[0] not in only:
            continue
        pos, extra = case_args(case)
        r = run_case(*pos, **extra)
        print(f"{r['name']:18s} G_maxdiff={r['G_maxdiff']:.3e} D_maxdiff={r['D_maxdiff']:.3e} "
              f"loss_maxdiff={r['loss_maxdiff']:.3e} swaps_bit_exact={r['swaps_bit_exact']}", flush=True)
