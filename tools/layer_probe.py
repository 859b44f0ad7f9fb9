#!/usr/bin/env python
"""Layer-by-layer parity of GenNet.backward / DiscNet.backward intermediates against torch autograd (fp64 CPU)."""
import copy, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import init_model, plugin, relerr, nchw
from mdgan_b200.nets import DiscNet, GenNet

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "CIFAR10"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mod = plugin(name)
g = torch.Generator().manual_seed(78)

# ------------------------------------------------------------------ generator
Gm = init_model(mod.Generator, 9)
z = torch.randn((n, mod.Z_DIM, 1, 1), generator=g)
s = torch.randn((n, *mod.SHAPE), generator=g) * 0.01
net = GenNet(Gm, mod.Z_DIM, mod.SHAPE, n, dev, 2e-4, 0.5, 0.999, precision=prec)
ref = copy.deepcopy(Gm).double()
acts = []
x = z.double()
for m in ref.main:
    x = m(x)
    x.retain_grad()
    acts.append((type(m).__name__, x))
x.backward(s.double() / 64)
net.forward(z.to(dev).view(n, mod.Z_DIM))
net.backward(s.to(dev), 1 / 64)
torch.cuda.synchronize()
# map: layer l -> conv output index 3l, bn out 3l+1, relu out 3l+2
for l in range(len(net.L) - 1):
    zc, zb, za = acts[3 * l][1], acts[3 * l + 1][1], acts[3 * l + 2][1]
    print(f"G layer {l}: z {relerr(nchw(net.z[l]), zc):.2e} a {relerr(nchw(net.a[l]), za):.2e} "
          f"da {relerr(nchw(net.da[l]), za.grad):.2e} dz {relerr(nchw(net.dz[l]), zc.grad):.2e}")
    e = (nchw(net.da[l]).double().cpu() - za.grad).abs()
    per_img = e.amax(dim=(1, 2, 3)) / za.grad.abs().max()
    bad = (per_img > 1e-4).nonzero().flatten().tolist()
    print("    da bad images:", bad[:20], "count", len(bad))
    e = (nchw(net.dz[l]).double().cpu() - zc.grad).abs()
    per_img = e.amax(dim=(1, 2, 3)) / zc.grad.abs().max()
    bad = (per_img > 1e-4).nonzero().flatten().tolist()
    per_ch = e.amax(dim=(0, 2, 3)) / zc.grad.abs().max()
    badc = (per_ch > 1e-4).nonzero().flatten().tolist()
    print("    dz bad images:", bad[:20], "count", len(bad), " bad channels", badc[:20], "count", len(badc))
for (pn, p) in ref.named_parameters():
    print(f"   G grad {pn:22s} relerr {relerr(net.state.g[pn], p.grad):.3e}")

# ------------------------------------------------------------------ discriminator (train step on real || x_d)
b = n // 2
D = init_model(mod.Discriminator, 5)
g = torch.Generator().manual_seed(77)
real = torch.rand((b, *mod.SHAPE), generator=g) * 2 - 1
x_d = torch.tanh(torch.randn((b, *mod.SHAPE), generator=g))
dnet = DiscNet(D, mod.SHAPE, b, dev, 2e-4, 0.5, 0.999, precision=prec)
for dt in (torch.float64, torch.float32):
    ref = copy.deepcopy(D).to(dt)
    crit = torch.nn.BCELoss()
    acts = {}
    outs = []
    for xin, lab in ((real, 1.0), (x_d, 0.0)):
        x = xin.to(dt)
        per = []
        for m in ref.main:
            x = m(x)
            if not isinstance(m, (torch.nn.LeakyReLU,)) or True:
                x.retain_grad()
            per.append((type(m).__name__, x))
        outs.append(per)
        loss = crit(x.view(-1), torch.full((b,), lab, dtype=dt))
        loss.backward()
    if dt == torch.float64:
        dnet.img[:b].copy_(real.to(dev)); dnet.img[b:2 * b].copy_(x_d.to(dev))
        dnet.forward(dnet.img, 2, dnet.labels_train)
        dnet.backward(dnet.img, 2, train=True)
        torch.cuda.synchronize()
    names = [nm for nm, _ in outs[0]]
    print(f"--- D reference dtype {dt}: modules {names}")
    # find conv outputs (pre-BN) and activation outputs per layer
    conv_idx = [i for i, nm in enumerate(names) if nm == "Conv2d"]
    for l in range(len(dnet.L) - 1):
        ci = conv_idx[l]
        ai = conv_idx[l + 1] - 1
        zc = torch.cat([outs[0][ci][1], outs[1][ci][1]]); zcg = torch.cat([outs[0][ci][1].grad, outs[1][ci][1].grad])
        za = torch.cat([outs[0][ai][1], outs[1][ai][1]]); zag = torch.cat([outs[0][ai][1].grad, outs[1][ai][1].grad])
        msg = f"D layer {l}: a {relerr(nchw(dnet.a[l][:n]), za):.2e} da {relerr(nchw(dnet.da[l][:n]), zag):.2e} dz {relerr(nchw(dnet.dz[l][:n]), zcg):.2e}"
        if dnet.z[l] is not None:
            msg += f" z {relerr(nchw(dnet.z[l][:n]), zc):.2e}"
        print(msg)
    for (pn, p) in ref.named_parameters():
        print(f"   D grad {pn:22s} relerr {relerr(dnet.state.g[pn], p.grad):.3e}")
    if dt == torch.float32:
        l = 1
        ci = conv_idx[l]
        zcg = torch.cat([outs[0][ci][1].grad, outs[1][ci][1].grad]).double()
        ours = nchw(dnet.dz[l][:n]).double().cpu()
        e = (ours - zcg).abs() / zcg.abs().max()
        idx = (e > 1e-4).nonzero()
        print("elements of dz[1] off by > 1e-4 of max:", idx.shape[0], "of", e.numel())
        ybn = torch.cat([outs[0][ci + 2][1], outs[1][ci + 2][1]]).double()  # post-LeakyReLU (inplace: also the BN output)
        for t in idx[:12].tolist():
            print("   at", t, "err", float(e[tuple(t)]), "y(fp32 ref)", float(ybn[tuple(t)]), "y ours", float(nchw(dnet.a[l][:n]).cpu()[tuple(t)]))
