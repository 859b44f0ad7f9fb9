"""Multi-GPU self-check: the N-process run (one GPU each, peer-memory or NCCL exchange, discriminator swaps included)
must leave BIT-IDENTICAL generator / discriminator states to the one-process run of the same job on one GPU.

Why this can hold bit for bit: every worker's D step and feedback are computed by the same kernels on the same inputs
wherever the worker lives (/root/reference/src/actors/worker.py:193-233 is per-worker work), the generated batch is
copied, not recomputed, and the only cross-GPU arithmetic -- the sum of the workers' feedbacks per generated batch
(/root/reference/src/actors/server.py:266-302) -- is taken in ascending worker order in both layouts.

Used by bench.py (`multi_gpu_bit_identical` in the JSON line), tools/multigpu_check.py and tests/test_multigpu_gpu.py.
Collective: every rank of the default process group must call `multi_gpu_bit_identity` with the same arguments.
"""
from __future__ import annotations

import importlib
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import routing
from .engine import EngineConfig, MDGANEngine


def _run(mod, dataset, shards, proc: int, n_procs: int, dev: torch.device, n_workers: int, batch: int, epochs: int,
         swap: int, graph: bool, seed: int):
    import bootstrap
    from .node import _DeviceBatches

    local = routing.workers_of_process(proc, n_procs, n_workers)
    discs = {}
    for n in local:
        bootstrap._seed_actor(seed + n + 1)
        d = mod.Discriminator()
        d.apply(bootstrap._weights_init)
        discs[n] = d
    gen = None
    bootstrap._seed_actor(seed)   # the server's stream: noise and swap permutations are drawn from it on process 0
    if proc == 0:
        gen = mod.Generator()
        gen.apply(bootstrap._weights_init)
    cfg = EngineConfig(n_workers=n_workers, batch_size=batch, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE),
                       swap_interval=swap, z_source="host")
    src = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], batch), dev, mod.SHAPE) for n in local}
    eng = MDGANEngine(cfg, proc, n_procs, dev, gen, discs, src)
    pairs_log = []
    for e in range(epochs):
        if graph and e == 2:
            eng.capture()
        eng.iteration(e)
        pairs_log.append(None if eng.last_pairs is None else eng.last_pairs.clone())
    torch.cuda.synchronize(dev)
    eng.sync_modules()
    losses = eng.mean_d_loss()
    mode = getattr(eng.exchange, "mode", "nccl") if n_procs > 1 else "none"
    eng.close()
    return gen, discs, losses, pairs_log, mode


def multi_gpu_bit_identity(rank: int, world: int, dev: torch.device, dataset_name: str = "CIFAR10",
                           n_workers: Optional[int] = None, batch: int = 16, epochs: int = 5, swap: int = 2,
                           graph: bool = True, seed: int = 3) -> Dict[str, object]:
    """Returns (on every rank) {"ok": bool, "world": N, "workers": K, "swaps": count, "exchange": mode, "mismatches":
    [names]}.  K defaults to the smallest even number >= world (swaps need an even worker count)."""
    from datasets.DataPartitioner import SyntheticImages

    mod = importlib.import_module(f"datasets.{dataset_name}")
    K = n_workers or (world + (world & 1) if world > 1 else 2)
    dataset = SyntheticImages(mod.SHAPE, K * 4 * batch)
    shards = routing.split_dataset(len(dataset), K, True)
    gen_n, discs_n, _, pairs_n, mode = _run(mod, dataset, shards, rank, world, dev, K, batch, epochs, swap, graph, seed)
    states = {n: {k: v.detach().cpu() for k, v in d.state_dict().items()} for n, d in discs_n.items()}
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, states)
    else:
        gathered = [states]
    result = [None]
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        gen_1, discs_1, _, pairs_1, _ = _run(mod, dataset, shards, 0, 1, dev, K, batch, epochs, swap, False, seed)
        bad = []
        for k, v in gen_1.state_dict().items():
            if not torch.equal(v.cpu(), gen_n.state_dict()[k].cpu()):
                bad.append(f"G.{k}")
        for n in range(K):
            for k, v in discs_1[n].state_dict().items():
                if not torch.equal(v.cpu(), merged[n][k]):
                    bad.append(f"D{n + 1}.{k}")
        for e, (a, c) in enumerate(zip(pairs_1, pairs_n)):
            if (a is None) != (c is None) or (a is not None and not torch.equal(a, c)):
                bad.append(f"pairs@{e}")
        result[0] = {"ok": not bad, "world": world, "workers": K, "dataset": dataset_name, "batch": batch,
                     "iterations": epochs, "swaps": sum(p is not None for p in pairs_n), "cuda_graph": graph,
                     "exchange": mode, "mismatches": bad[:8]}
    if world > 1:
        dist.broadcast_object_list(result, src=0)
    return result[0]
