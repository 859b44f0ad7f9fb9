#!/usr/bin/env python
"""MMA-only rate of conv_gemm (gathers skipped, dbg=2) for several forced tile widths: is SS-mode bound by the A read?"""
import os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200"))
from mdgan_b200 import _lib, ops
dev = torch.device("cuda:0")
n, C, N, H = 512, 128, 256, 16   # M = 32768 rows -> 256 row tiles
x = torch.randn(n, H, H, C, device=dev)
W = torch.randn(N, C, 4, 4, device=dev) * 0.05
out = torch.empty(n, H // 2, H // 2, N, device=dev)
for dbg in (2, 0):
    pass  # debug probe flags were removed after the round-1 analysis
    for prec in (0, 1):
        wp = ops.pack_down(W, precision=prec)
        for bn in (32, 64, 128):
            if prec == 1 and bn > 64:
                continue
            fn = lambda: ops.conv_gemm(x, wp, ops.MODE_DOWN, N, out, (n, H // 2, H // 2), (H, H), precision=prec, force_bn=bn)
            fn(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10): fn()
            e.record(); torch.cuda.synchronize()
            us = s.elapsed_time(e) * 100
            ctas = 256 * (N // bn)
            waves = -(-ctas // 148)
            ksteps = 16 * C // 32
            mmas = ksteps * 4 * (3 if prec else 1)
            clk = us * 1e-6 * 1.9e9 / waves / mmas
            print(f"dbg {dbg} prec {prec} bn {bn}: {us:8.1f} us, {ctas} CTAs ({waves} waves), ~{clk:6.1f} clk per MMA (128x{bn}x8) at 1.9 GHz")
