#!/usr/bin/env python
"""Launcher for the UNMODIFIED reference actors with an ODD number of workers (e.g. K = 1, the single-GPU benchmark
configuration).  The reference's own launcher refuses an even world size (/root/reference/src/bootstrap.py:163-164,
because its discriminator swap pairs workers up); the training loop itself (actors/server.py, actors/worker.py,
bootstrap.init_process / run) has no such requirement as long as no swap is due.  This script repeats the tail of
bootstrap.py's `__main__` block (lines 166-187) verbatim in behaviour -- plugin import, partitioner, mp.spawn of
bootstrap.init_process with bootstrap.run -- and skips only that guard.  Same command line as bootstrap.py.

Measurement infrastructure only (oracle/ref_harness/time_reference.py uses it when K is odd).
"""
import importlib

import torch.multiprocessing as mp

import bootstrap  # the reference module: parses the command line at import, defines init_process / run

if __name__ == "__main__":
    args = bootstrap.args
    if ".." in args.ranks:
        a, b = args.ranks.split("..")
        ranks = list(range(int(a), int(b) + 1))
    else:
        ranks = [int(r) for r in args.ranks.split(",")]
    mod = importlib.import_module(f"datasets.{args.dataset}")
    partioner = mod.Partitioner(args.world_size, 0)
    partioner.load_data()
    mp.spawn(bootstrap.init_process,
             args=(args, ranks, partioner, mod.SHAPE, mod.Z_DIM, mod.Generator, mod.Discriminator, bootstrap.run),
             nprocs=len(ranks), join=True)
