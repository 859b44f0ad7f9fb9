#!/usr/bin/env python
"""Print the per-tensor parity errors of DiscNet / GenNet against the torch-CPU oracle functions."""
import copy, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import init_model, plugin, relerr
from mdgan_b200.nets import DiscNet, GenNet
from oracle.mdgan_oracle import d_feedback, d_train_step
dev = torch.device("cuda:0")
import itertools
for (name, b), prec in itertools.product([("CIFAR10", 8), ("CIFAR10", 64), ("CelebA", 16), ("MNIST_DCGAN", 32)], [1, 0]):
    print(f"=== precision {'tf32x3' if prec else 'tf32'} ===")
    mod = plugin(name)
    D = init_model(mod.Discriminator, 5)
    g = torch.Generator().manual_seed(77)
    real = torch.rand((b, *mod.SHAPE), generator=g) * 2 - 1
    x_d = torch.tanh(torch.randn((b, *mod.SHAPE), generator=g)); x_g = torch.tanh(torch.randn((b, *mod.SHAPE), generator=g))
    net = DiscNet(D, mod.SHAPE, b, dev, 2e-4, 0.5, 0.999, precision=prec)
    ref = copy.deepcopy(D); opt = torch.optim.Adam(ref.parameters(), lr=2e-4, betas=(0.5, 0.999))
    rl = d_train_step(ref, opt, real, x_d); rg = {n: p.grad.clone() for n, p in ref.named_parameters()}
    rlg, rfb = d_feedback(ref, x_g)
    loss = net.train_step(real.to(dev), x_d.to(dev)).item()
    print(f"{name} b={b} D loss {loss:.6f} ref {rl.item():.6f}")
    for n, gr in rg.items():
        print(f"   grad {n:24s} relerr {relerr(net.state.g[n], gr):.3e}  |ref|max {gr.abs().max():.3e}")
    lg, fb = net.feedback_step(x_g.to(dev)), net.feedback
    print(f"   loss_gen {lg.item():.6f} ref {rlg.item():.6f}  feedback relerr {relerr(fb, rfb):.3e}")
    n = 2 * b
    Gm = init_model(mod.Generator, 9)
    z = torch.randn((n, mod.Z_DIM, 1, 1), generator=g); s = torch.randn((n, *mod.SHAPE), generator=g) * 0.01
    gnet = GenNet(Gm, mod.Z_DIM, mod.SHAPE, n, dev, 2e-4, 0.5, 0.999, precision=prec)
    refg = copy.deepcopy(Gm); X = refg(z)
    grads = torch.autograd.grad(X, list(refg.parameters()), grad_outputs=s / 64)
    Xg = gnet.forward(z.to(dev).view(n, mod.Z_DIM)); gnet.backward(s.to(dev), 1 / 64)
    print(f"   G X relerr {relerr(Xg, X):.3e}")
    for (pn, _), gr in zip(refg.named_parameters(), grads):
        print(f"   G grad {pn:22s} relerr {relerr(gnet.state.g[pn], gr):.3e}")

# ---- calibration: how far is torch's OWN cuDNN path (TF32 on / off) from torch CPU fp32 on the same problem?
print("\n== calibration: torch CUDA (cuDNN) vs torch CPU fp32, CIFAR10 b=64 ==")
mod = plugin("CIFAR10"); b = 64
D = init_model(mod.Discriminator, 5)
g = torch.Generator().manual_seed(77)
real = torch.rand((b, *mod.SHAPE), generator=g) * 2 - 1
x_d = torch.tanh(torch.randn((b, *mod.SHAPE), generator=g)); x_g = torch.tanh(torch.randn((b, *mod.SHAPE), generator=g))
ref = copy.deepcopy(D); opt = torch.optim.Adam(ref.parameters(), lr=2e-4, betas=(0.5, 0.999))
d_train_step(ref, opt, real, x_d); rg = {n: p.grad.clone() for n, p in ref.named_parameters()}
_, rfb = d_feedback(ref, x_g)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    Dg = copy.deepcopy(D).to(dev)
    optg = torch.optim.Adam(Dg.parameters(), lr=2e-4, betas=(0.5, 0.999))
    crit = torch.nn.BCELoss()
    Dg.zero_grad()
    l = crit(Dg(real.to(dev)), torch.ones(b, device=dev)) + crit(Dg(x_d.to(dev)), torch.zeros(b, device=dev))
    l.backward(); gg = {n: p.grad.clone() for n, p in Dg.named_parameters()}; optg.step()
    xg = x_g.to(dev).requires_grad_(True)
    crit(Dg(xg), torch.ones(b, device=dev)).backward()
    print(f" allow_tf32={tf32}: feedback relerr {relerr(xg.grad, rfb):.3e}; " + " ".join(f"{n}:{relerr(gg[n], rg[n]):.2e}" for n in rg))
