#!/usr/bin/env python
"""N-GPU run (one process per GPU, NCCL) against the 1-GPU run of the same job: final generator / discriminator
states must be bit-identical (the only cross-GPU arithmetic is the feedback reduce, a sum of disjoint slots).

    torchrun --nproc-per-node 2 tools/multigpu_check.py [--workers 4] [--graph]
"""
import argparse, os, sys
import torch, torch.distributed as dist
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-gan_b200"))
import importlib
from datasets.DataPartitioner import SyntheticImages
from mdgan_b200 import routing
from mdgan_b200.engine import EngineConfig, MDGANEngine
from mdgan_b200.node import _DeviceBatches
import bootstrap

ap = argparse.ArgumentParser()
ap.add_argument("--workers", type=int, default=4)
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
mod = importlib.import_module(f"datasets.{a.dataset}")
N, b = a.workers, a.batch
dataset = SyntheticImages(mod.SHAPE, N * 4 * b)
shards = routing.split_dataset(len(dataset), N, True)


def run(proc, n_procs, graph):
    local = routing.workers_of_process(proc, n_procs, N)
    discs = {}
    for n in local:
        bootstrap._seed_actor(3 + n + 1)
        d = mod.Discriminator(); d.apply(bootstrap._weights_init); discs[n] = d
    gen = None
    bootstrap._seed_actor(3)
    if proc == 0:
        gen = mod.Generator(); gen.apply(bootstrap._weights_init)
    cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=a.swap, z_source="host")
    src = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], b), dev, mod.SHAPE) for n in local}
    eng = MDGANEngine(cfg, proc, n_procs, dev, gen, discs, src)
    for e in range(a.epochs):
        if graph and e == 2:
            eng.capture()
        eng.iteration(e)
    torch.cuda.synchronize()
    eng.sync_modules()
    eng.close()
    return gen, discs, eng


gen_n, discs_n, eng_n = run(rank, world, a.graph)
# gather the discriminator states on rank 0
states = {n: {k: v.cpu() for k, v in d.state_dict().items()} for n, d in discs_n.items()}
gathered = [None] * world
dist.all_gather_object(gathered, states)
dist.barrier()
if rank == 0:
    merged = {}
    for g in gathered:
        merged.update(g)
    from mdgan_b200.exchange import Exchange
    gen_1, discs_1, eng_1 = run(0, 1, False)
    gen_2, discs_2, eng_2 = run(0, 1, False)
    self_ok = all(torch.equal(v, gen_2.state_dict()[k]) for k, v in gen_1.state_dict().items())
    for n in range(N):
        self_ok &= all(torch.equal(v, discs_2[n].state_dict()[k]) for k, v in discs_1[n].state_dict().items())
    print("1-GPU run repeated twice in one process:", "bit-identical" if self_ok else "DIFFERENT", flush=True)
    ok = True
    for k, v in gen_1.state_dict().items():
        same = torch.equal(v.cpu(), gen_n.state_dict()[k].cpu())
        ok &= same
        if not same:
            print("G mismatch", k, flush=True) if False else print("G mismatch", k, (v.cpu() - gen_n.state_dict()[k].cpu()).abs().max().item())
    for n in range(N):
        for k, v in discs_1[n].state_dict().items():
            same = torch.equal(v.cpu(), merged[n][k])
            ok &= same
            if not same:
                print(f"D{n + 1} mismatch", k, (v.cpu().float() - merged[n][k].float()).abs().max().item())
    print(f"multigpu_check world={world} workers={N} graph={a.graph}: {'BIT-IDENTICAL' if ok else 'MISMATCH'} "
          f"(d_loss {eng_1.mean_d_loss()})", flush=True)
dist.barrier()
dist.destroy_process_group()
