#!/usr/bin/env python
"""Print the whole-loop parity report (engine vs oracle) for a few configurations."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import run_engine_vs_oracle
cfgs = [("CIFAR10", 1, 16, 3, 10**6), ("CIFAR10", 2, 8, 4, 2), ("CIFAR10", 4, 8, 4, 1), ("MNIST_DCGAN", 2, 16, 3, 10**6),
        ("CelebA", 2, 8, 3, 1), ("CelebA", 8, 8, 2, 1)]
import os
sel = os.environ.get('PROBE_ONLY')
if sel:
    cfgs = [c for c in cfgs if f'{c[0]}-{c[1]}' == sel]
for c in cfgs:
    print(c, run_engine_vs_oracle(*c, mode='trajectory'), flush=True)
    print(c, run_engine_vs_oracle(*c, mode='free'), flush=True)
