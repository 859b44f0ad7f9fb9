#!/usr/bin/env python
"""Repeatability stress: every kernel of the step is deterministic by construction (fixed reduction orders), so the
same inputs must give the same BITS on every repetition.  Runs generator forward + backward and a discriminator
training forward/backward REPS times from identical state and reports the first buffer that ever differs -- a race in a
pipelined kernel shows up here long before it shows up in a tolerance test.

    python tools/stress_repeat.py [dataset] [n] [reps]
"""
import copy, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import init_model, plugin
from mdgan_b200.nets import DiscNet, GenNet


def snapshot(net, extra):
    bufs = {}
    for name in ("z", "a", "da", "dz"):
        for i, t in enumerate(getattr(net, name)):
            if t is not None:
                bufs[f"{name}[{i}]"] = t.clone()
    bufs["grad"] = net.state.grad.clone()
    for k, v in extra.items():
        bufs[k] = v.clone()
    return bufs


def run(name="CIFAR10", n=128, reps=200):
    dev = torch.device("cuda:0")
    mod = plugin(name)
    g = torch.Generator().manual_seed(5)
    z = torch.randn((n, mod.Z_DIM), generator=g).to(dev)
    s = (torch.randn((n, *mod.SHAPE), generator=g) * 0.01).to(dev)
    real = (torch.rand((n // 2, *mod.SHAPE), generator=g) * 2 - 1).to(dev)
    fake = torch.tanh(torch.randn((n // 2, *mod.SHAPE), generator=g)).to(dev)
    gen = GenNet(init_model(mod.Generator, 9), mod.Z_DIM, mod.SHAPE, n, dev, lr=2e-4, beta_1=0.5, beta_2=0.999)
    disc = DiscNet(init_model(mod.Discriminator, 5), mod.SHAPE, n // 2, dev, lr=2e-4, beta_1=0.5, beta_2=0.999)
    first_g = first_d = None
    bad = 0
    for rep in range(reps):
        X = gen.forward(z)
        gen.backward(s, 1.0 / 64)
        snap_g = snapshot(gen, {"X": X})
        disc.img[: n // 2].copy_(real); disc.img[n // 2: n].copy_(fake)
        disc.forward(disc.img, 2, disc.labels_train)
        disc.backward(disc.img, 2, train=True)
        snap_d = snapshot(disc, {"loss": disc.loss})
        torch.cuda.synchronize()
        if first_g is None:
            first_g, first_d = snap_g, snap_d
            continue
        for label, a, b in (("G", first_g, snap_g), ("D", first_d, snap_d)):
            for k in a:
                if not torch.equal(a[k], b[k]):
                    diff = (a[k] - b[k]).abs()
                    print(f"rep {rep}: {label}.{k} differs: {int((diff > 0).sum())} elements, max {diff.max().item():.3e}", flush=True)
                    bad += 1
                    break
    print(f"stress_repeat {name} n={n} reps={reps}: {'REPEATABLE' if bad == 0 else f'{bad} MISMATCHING REPETITIONS'}", flush=True)
    return bad


if __name__ == "__main__":
    a = sys.argv[1:]
    sys.exit(1 if run(a[0] if a else "CIFAR10", int(a[1]) if len(a) > 1 else 128, int(a[2]) if len(a) > 2 else 200) else 0)
