#!/usr/bin/env python
"""Hardware probe: run every kernel family once against torch (CPU fp64 truth) and print the error table.
Used while bringing the sm_100a kernels up (descriptor encodings cannot be checked without a B200)."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "distributed-gan_b200"))
from mdgan_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def report(name, err, tol=3e-3):
    print(f"{name:44s} relerr={err:.3e}  {'OK' if err < tol else 'FAIL'}", flush=True)
    return err < tol


def probe_down(n, C, N, H, force_bn=0):
    x = torch.randn(n, C, H, H)
    W = torch.randn(N, C, 4, 4) * 0.05
    ref = F.conv2d(x.double(), W.double(), stride=2, padding=1)
    wp = ops.pack_down(W.to(dev))
    out = torch.empty(n, H // 2, H // 2, N, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), wp, ops.MODE_DOWN, N, out, (n, H // 2, H // 2), (H, H), force_bn=force_bn)
    torch.cuda.synchronize()
    return report(f"DOWN n={n} C={C} N={N} H={H} bn={force_bn}", relerr(nchw(out), ref))


def probe_up(n, C, N, H, nchw_out=False, force_bn=0):
    x = torch.randn(n, C, H, H)
    W = torch.randn(C, N, 4, 4) * 0.05
    ref = F.conv_transpose2d(x.double(), W.double(), stride=2, padding=1)
    wq = ops.pack_up(W.to(dev))
    if nchw_out:
        out = torch.empty(n, N, 2 * H, 2 * H, device=dev)
    else:
        out = torch.empty(n, 2 * H, 2 * H, N, device=dev)
    ops.conv_gemm(nhwc(x).to(dev), wq, ops.MODE_UP, N, out, (n, H, H), (H, H), out_nchw=nchw_out, force_bn=force_bn)
    torch.cuda.synchronize()
    got = out if nchw_out else nchw(out)
    return report(f"UP n={n} C={C} N={N} H={H} nchw={nchw_out} bn={force_bn}", relerr(got, ref))


def probe_dense(n, C, N, k):
    z = torch.randn(n, C, 1, 1)
    W = torch.randn(C, N, k, k) * 0.05
    ref = F.conv_transpose2d(z.double(), W.double())
    wp = ops.pack_dense(W.to(dev))
    Cp = wp.shape[1]
    zp = torch.zeros(n, Cp, device=dev)
    ops.pad_rows(z.view(n, C).to(dev), zp, round_tf32=True)
    out = torch.empty(n, k, k, N, device=dev)
    ops.conv_gemm(zp, wp, ops.MODE_DENSE, k * k * N, out, (n, 1, 1), (1, 1))
    torch.cuda.synchronize()
    return report(f"DENSE n={n} C={C} N={N} k={k}", relerr(nchw(out), ref))


def probe_wgrad(n, C1, C2, Hl, lbo=0, sbo=0):
    """conv wgrad: lo = dOut [n,Hl,Hl,C1], hi = input [n,2Hl,2Hl,C2] -> dW [C1,C2,4,4]"""
    x = torch.randn(n, C2, 2 * Hl, 2 * Hl)
    dout = torch.randn(n, C1, Hl, Hl)
    xr = x.double().requires_grad_(False)
    W = torch.zeros(C1, C2, 4, 4, dtype=torch.double, requires_grad=True)
    F.conv2d(xr, W, stride=2, padding=1).backward(dout.double())
    ref = W.grad
    splits = ops.wgrad_splits(n, Hl, Hl, C1, C2, 0)
    partial = torch.empty(splits * 16 * C1 * C2, device=dev)
    grad = torch.empty(C1, C2, 4, 4, device=dev)
    ops.wgrad_gemm(nhwc(dout).to(dev), nhwc(x).to(dev), partial, (n, Hl, Hl), 0, splits)
    ops.wgrad_unpack(partial, grad, 0, splits, C1, C1, C2)
    torch.cuda.synchronize()
    return report(f"WGRAD n={n} C1={C1} C2={C2} Hl={Hl} splits={splits} lbo={lbo} sbo={sbo}", relerr(grad, ref))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", _lib.load().mdgan_abi_version(), "dev-check", _lib.load().mdgan_check_device())
    t0 = time.time()
    ok = True
    ok &= probe_down(8, 64, 128, 16)
    ok &= probe_down(8, 64, 128, 16, force_bn=128)
    ok &= probe_down(8, 64, 128, 16, force_bn=32)
    ok &= probe_down(4, 128, 256, 8)
    ok &= probe_down(3, 256, 512, 8)          # ragged M (3*16=48 rows)
    ok &= probe_up(8, 512, 256, 4)
    ok &= probe_up(8, 256, 128, 8)
    ok &= probe_up(4, 128, 64, 16)
    ok &= probe_up(4, 64, 3, 16, nchw_out=True)
    ok &= probe_up(4, 128, 3, 16, nchw_out=True)
    ok &= probe_dense(16, 100, 512, 4)
    ok &= probe_dense(130, 100, 256, 7)
    w_ok = probe_wgrad(8, 128, 64, 8)
    if not w_ok:
        for lbo, sbo in [(512, 4096), (4096, 1024), (1024, 4096), (4096, 4096), (4096, 128), (128, 4096), (4096, 256)]:
            if probe_wgrad(8, 128, 64, 8, lbo, sbo):
                print(f"  -> WGRAD encoding that works: lbo={lbo} sbo={sbo}")
    ok &= probe_wgrad(8, 256, 128, 4)
    ok &= probe_wgrad(16, 128, 64, 16)
    ok &= probe_wgrad(5, 512, 256, 4)
    print("ALL OK" if ok and w_ok else "SOME FAILED", f"({time.time()-t0:.1f}s)")
