#!/usr/bin/env python
"""N-GPU run (one process per GPU) against the 1-GPU run of the same job: final generator / discriminator states and
the swap permutations must be bit-identical (mdgan_b200/selfcheck.py).

    torchrun --nproc-per-node N tools/multigpu_check.py [--workers K] [--graph] [--dataset CIFAR10] [--swap 2]
"""
import argparse, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200"))
from mdgan_b200.selfcheck import multi_gpu_bit_identity

ap = argparse.ArgumentParser()
ap.add_argument("--workers", type=int, default=0, help="default: the smallest even number >= world size")
ap.add_argument("--dataset", default="CIFAR10")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--epochs", type=int, default=5)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--swap", type=int, default=2)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
r = multi_gpu_bit_identity(rank, world, dev, a.dataset, a.workers or None, a.batch, a.epochs, a.swap, a.graph)
if rank == 0:
    print(f"multigpu_check world={world} workers={r['workers']} graph={a.graph} exchange={r['exchange']} "
          f"swaps={r['swaps']}: {'BIT-IDENTICAL' if r['ok'] else 'MISMATCH ' + json.dumps(r['mismatches'])}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if r["ok"] else 1)
