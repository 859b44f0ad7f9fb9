#!/usr/bin/env python
"""Measured parity numbers of the whole loop against the oracle, per configuration and mode (tests/parity.py), as one
JSON line each -- the source of the tolerances in tests/parity.py (TOL = at most 10x the worst value printed here).

    python tools/parity_report.py [quick]
"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import run_engine_vs_oracle

CASES = [  # name, N, b, epochs, swap_interval
    ("CIFAR10", 1, 16, 3, 10**6), ("CIFAR10", 2, 8, 4, 2), ("CIFAR10", 4, 8, 4, 1), ("MNIST_DCGAN", 2, 16, 3, 10**6),
    ("CelebA", 2, 8, 3, 1), ("CelebA", 8, 8, 2, 1), ("CIFAR10", 2, 10, 3, 2),
    ("MNIST_DCGAN", 1, 64, 2, 10**6), ("MNIST_DCGAN", 2, 64, 2, 10**6), ("CIFAR10", 4, 64, 3, 1), ("CelebA", 2, 64, 2, 1),
    ("CelebA", 8, 64, 2, 1),
]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    CASES = CASES[:2]
for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("trajectory", "unpatched", "free")):
    for c in CASES:
        t0 = time.time()
        r = run_engine_vs_oracle(*c, mode=mode)
        r.update(case=list(c), seconds=round(time.time() - t0, 1))
        print(json.dumps(r), flush=True)
