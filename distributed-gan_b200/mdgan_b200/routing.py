"""Host-side control logic of the MD-GAN step: everything that must be BIT-EXACT with the reference.

None of this touches the GPU.  Each function restates one piece of the reference's actor code with the same torch
CPU calls in the same order, so the random streams (shard permutation, real-batch order, swap pairs, noise in
parity mode) are identical to a reference run with the same `--seed`:

    k, routing      /root/reference/src/actors/server.py:116-120,238-239
    shards          /root/reference/src/actors/server.py:46-64,151-154
    real batches    /root/reference/src/actors/worker.py:78-89,162-167
    swap schedule   /root/reference/src/actors/server.py:315-324 ; worker.py:239-240
    actor -> GPU    SURVEY.md H3: world_size = N+1 actors (server rank 0, workers 1..N) on N GPU processes;
                    the server shares GPU process 0 with worker 1.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import os

import torch
import torch.utils.data


def num_workers(world_size: int) -> int:
    """server.py:116 -- N = world_size - 1 (world_size counts the server)."""
    if world_size < 2:
        raise ValueError("world_size must be >= 2 (one server + at least one worker)")
    return world_size - 1


def num_generated_batches(n_workers: int) -> int:
    """server.py:120 -- k = max(floor(ln N), 2)."""
    return max(math.floor(math.log(n_workers)), 2)


def route(n: int, k: int) -> Tuple[int, int]:
    """server.py:238-239 -- 0-based worker n trains on (X_g, X_d) = (K[n % k], K[(n + 1) % k])."""
    return n % k, (n + 1) % k


def feedback_slot(n: int, k: int) -> int:
    """Slot of the [k, b, C, H, W] grad-output buffer that worker n's feedback is summed into (server.py:272)."""
    return n % k


def split_dataset(dataset_size: int, n_workers: int, iid: bool) -> Tuple[torch.Tensor, ...]:
    """server.py:46-64 with the private seed-0 generator of server.py:151-153."""
    if iid:
        g = torch.Generator()
        g.manual_seed(0)
        idx = torch.randperm(dataset_size, generator=g)
    else:
        idx = torch.arange(dataset_size)
    return torch.chunk(idx, n_workers)


def swap_due(epoch: int, swap_interval: int, n_workers: int) -> bool:
    """server.py:315-317 and worker.py:239-240."""
    return n_workers > 1 and epoch % swap_interval == 0 and epoch > 0


def draw_swap_pairs(n_workers: int) -> torch.Tensor:
    """server.py:321-324 -- from the server's GLOBAL torch RNG; [N/2, 2] int32 of 1-based worker ranks."""
    if n_workers % 2 != 0:
        raise ValueError("discriminator swap needs an even number of workers (bootstrap.py:163-164)")
    return torch.randperm(n_workers, dtype=torch.int).view(-1, 2) + 1


def partners_from_pairs(pairs: torch.Tensor) -> Dict[int, int]:
    """{worker rank -> partner rank} (what each worker receives at worker.py:243-246)."""
    out: Dict[int, int] = {}
    for a, c in pairs.tolist():
        out[a] = c
        out[c] = a
    return out


class RealBatchStream:
    """worker.py:78-89,162-167 -- DataLoader(Subset(dataset, shard), b, shuffle=True, generator seed 0); the
    iterator is re-created on exhaustion.  Yields CPU fp32 [b, C, H, W] batches."""

    def __init__(self, dataset, shard: torch.Tensor, batch_size: int, num_workers: Optional[int] = None):
        """num_workers (default: MDGAN_LOADER_WORKERS, 0 like the reference): loader processes that decode /
        transform samples ahead of the training loop.  The sampler, and therefore the order and content of the
        batches, does not depend on it (the DataLoader returns batches in sampler order)."""
        g = torch.Generator()
        g.manual_seed(0)
        self.batch_size = batch_size
        if num_workers is None:
            num_workers = int(os.environ.get("MDGAN_LOADER_WORKERS", "0"))
        kw = dict(num_workers=num_workers, prefetch_factor=4) if num_workers > 0 else {}
        self.loader = torch.utils.data.DataLoader(torch.utils.data.Subset(dataset, shard), batch_size=batch_size,
                                                  shuffle=True, generator=g, **kw)
        self.it = iter(self.loader)

    def next(self) -> torch.Tensor:
        try:
            batch = next(self.it)[0]
        except StopIteration:
            self.it = iter(self.loader)
            batch = next(self.it)[0]
        if batch.shape[0] != self.batch_size:
            # the reference would crash here (BCELoss against fixed-size labels, worker.py:114-115,199)
            raise ValueError(f"ragged real batch of {batch.shape[0]} (batch_size {self.batch_size}): the shard size "
                             "must be a multiple of the batch size")
        return batch


# ---------------------------------------------------------------------------------------------- actor placement
def workers_of_process(proc: int, n_procs: int, n_workers: int) -> List[int]:
    """0-based worker indices hosted by GPU process `proc` (contiguous blocks; one worker per GPU when
    n_procs == n_workers, all of them when n_procs == 1)."""
    if n_procs < 1 or n_procs > n_workers:
        raise ValueError(f"need 1 <= n_procs ({n_procs}) <= n_workers ({n_workers})")
    base, extra = divmod(n_workers, n_procs)
    start = proc * base + min(proc, extra)
    return list(range(start, start + base + (1 if proc < extra else 0)))


def process_of_worker(n: int, n_procs: int, n_workers: int) -> int:
    for p in range(n_procs):
        if n in workers_of_process(p, n_procs, n_workers):
            return p
    raise ValueError(n)


def parse_ranks(spec: str) -> List[int]:
    """bootstrap.py:150-159 -- "a..b" | "a,b,c" | "n"."""
    if ".." in spec:
        a, b = spec.split("..")
        return list(range(int(a), int(b) + 1))
    if "," in spec:
        return [int(r) for r in spec.split(",")]
    if spec.isdigit():
        return [int(spec)]
    raise ValueError("Invalid rank format")
