"""Server (generator) actor -- same entry point as /root/reference/src/actors/server.py:67-87.

`start(...)` keeps the reference's keyword signature.  What changed underneath: the reference runs
N+1 OS processes (server rank 0 + N workers) over gloo; here there is one process per GPU over NCCL, and the GPU
process that runs the server also runs the worker(s) placed on GPU 0, because NCCL forbids two ranks per GPU.  The
co-resident workers are handed in through the extra keyword `colocated_workers` (bootstrap.py does this); without
it `start` refuses to run rather than leaving GPU 0 without a discriminator.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Tuple

import os

import torch
import torch.utils.data

from mdgan_b200 import routing
from mdgan_b200.engine import EngineConfig
from mdgan_b200.node import run_node


def _split_dataset(dataset_size: int, world_size: int, iid: bool = False):
    """Reference helper server.py:46-64 (`world_size` here is the number of workers, as at its call site :153)."""
    return routing.split_dataset(dataset_size, world_size, iid)


def start(
    backend: str,
    rank: int,
    generator_lr: float,
    world_size: int,
    batch_size: int,
    epochs: int,
    log_interval: int,
    generator: torch.nn.Module,
    dataset: torch.utils.data.Dataset,
    z_dim: int,
    log_folder: Path,
    image_shape: Tuple[int, int, int],
    dataset_name: str,
    device: torch.device = torch.device("cpu"),
    n_samples: int = 5,
    iid: bool = True,
    swap_interval: int = 1,
    beta_1: float = 0.5,
    beta_2: float = 0.999,
    *,
    colocated_workers: Optional[Dict[int, torch.nn.Module]] = None,
    discriminator_lr: Optional[float] = None,
    local_epochs: int = 1,
    n_procs: Optional[int] = None,
    z_source: str = "host",
):
    if rank != 0:
        raise ValueError("the server is rank 0 (bootstrap.py:100)")
    N = routing.num_workers(world_size)
    n_procs = n_procs or N
    hosted = routing.workers_of_process(0, n_procs, N)
    if not colocated_workers or sorted(colocated_workers) != [n + 1 for n in hosted]:
        raise RuntimeError(
            f"server.start: GPU process 0 also hosts worker rank(s) {[n + 1 for n in hosted]}; pass their "
            "discriminator modules as colocated_workers={rank: module} (bootstrap.py does). The reference's "
            "separate server process does not exist on the one-process-per-GPU NCCL layout.")
    cfg = EngineConfig(n_workers=N, batch_size=batch_size, z_dim=z_dim, image_shape=tuple(image_shape),
                       generator_lr=generator_lr,
                       discriminator_lr=generator_lr if discriminator_lr is None else discriminator_lr,
                       beta_1=beta_1, beta_2=beta_2, swap_interval=swap_interval, local_epochs=local_epochs,
                       z_source=z_source, prefetch_host=os.environ.get("MDGAN_PREFETCH", "1") == "1")
    return run_node(backend=backend, proc=0, n_procs=n_procs, world_size=world_size, device=torch.device(device),
                    cfg=cfg, generator=generator, discriminators={r - 1: m for r, m in colocated_workers.items()},
                    dataset=dataset, epochs=epochs, log_interval=log_interval, log_folder=Path(log_folder),
                    dataset_name=dataset_name, iid=iid, n_samples=n_samples)
