#!/usr/bin/env python
"""Time the tensor-core kernels on the layer shapes of the benchmark workloads (CUDA events, L2 flushed between reps).

    python tools/conv_bench.py [precision 0|1] [dbg flags]
"""
import os, sys, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200"))
from mdgan_b200 import _lib, ops
dev = torch.device("cuda:0")
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pass  # debug probe flags were removed after the round-1 analysis
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

print(f"precision {'tf32x3' if prec else 'tf32'} dbg {dbg}")
# (label, mode, n, C, N, H_src)
layers = [("celebaD c2 down", 0, 128, 64, 128, 32), ("celebaD c3 down", 0, 128, 128, 256, 16), ("celebaD c4 down", 0, 128, 256, 512, 8),
          ("celebaG L2 up", 1, 128, 512, 256, 4), ("celebaG L3 up", 1, 128, 256, 128, 8), ("celebaG L4 up", 1, 128, 128, 64, 16),
          ("celebaG L5 up(nchw,3)", 1, 128, 64, 3, 32), ("mnistD c2 down", 0, 128, 64, 128, 14), ("mnistG L2 up", 1, 128, 256, 128, 7),
          ("b1024 celebaD c3", 0, 2048, 128, 256, 16),
          ("mnistG dgrad L2 down", 0, 128, 128, 256, 14), ("mnistD dgrad c2 up", 1, 128, 128, 64, 7),
          ("mnistD c2 fb down", 0, 64, 64, 128, 14), ("cifarD c3 down", 0, 128, 128, 256, 8), ("cifarG L3 up", 1, 128, 256, 128, 8)]
only = os.environ.get('CONV_BENCH_ONLY')
if only:
    layers = [l for l in layers if only in l[0]]
if os.environ.get('WGRAD_BENCH_ONLY'):
    layers = []
for label, mode, n, C, N, H in layers:
    x = torch.randn(n, H, H, C, device=dev)
    if mode == 0:
        W = torch.randn(N, C, 4, 4, device=dev) * 0.05
        wp = ops.pack_down(W, precision=prec)
        Hg = H // 2
        out = torch.empty(n, Hg, Hg, N, device=dev)
        fn = lambda: ops.conv_gemm(x, wp, ops.MODE_DOWN, N, out, (n, Hg, Hg), (H, H), precision=prec)
        flops = 2.0 * n * Hg * Hg * 16 * C * N
    else:
        W = torch.randn(C, N, 4, 4, device=dev) * 0.05
        wp = ops.pack_up(W, precision=prec)
        nchw = N <= 3
        out = torch.empty((n, N, 2 * H, 2 * H) if nchw else (n, 2 * H, 2 * H, N), device=dev)
        fn = lambda: ops.conv_gemm(x, wp, ops.MODE_UP, N, out, (n, H, H), (H, H), out_nchw=nchw, precision=prec)
        flops = 2.0 * n * H * H * 16 * C * N
    us = timeit(fn)
    print(f"  {label:24s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (algorithmic)")
# wgrad
wonly = os.environ.get('WGRAD_BENCH_ONLY')
for label, n, C1, C2, Hl in [] if only else [w for w in [("celebaD c3 wgrad", 128, 256, 128, 8), ("celebaD c4 wgrad", 128, 512, 256, 4), ("celebaG L3 wgrad", 128, 256, 128, 8),
                             ("celebaD c2 wgrad", 128, 128, 64, 16), ("b1024 celebaD c3 wgrad", 2048, 256, 128, 8)] if not wonly or wonly in w[0]]:
    lo = torch.randn(n, Hl, Hl, C1, device=dev); hi = torch.randn(n, 2 * Hl, 2 * Hl, C2, device=dev)
    splits = ops.wgrad_splits(n, Hl, Hl, C1, C2, ops.MODE_DOWN)
    partial = torch.empty(splits * 16 * C1 * C2, device=dev)
    fn = lambda: ops.wgrad_gemm(lo, hi, partial, (n, Hl, Hl), ops.MODE_DOWN, splits, precision=prec)
    us = timeit(fn)
    flops = 2.0 * n * Hl * Hl * 16 * C1 * C2
    print(f"  {label:24s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (algorithmic)  splits {splits}")
