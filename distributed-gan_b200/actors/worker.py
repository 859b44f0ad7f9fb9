"""Worker (discriminator) actor -- same entry point as /root/reference/src/actors/worker.py:20-41.

One GPU process runs one worker (or several, when there are fewer GPUs than workers: pass the others through
`colocated_workers`).  Worker rank 1 always lives in the server's process (actors/server.py), so calling
`start(rank=1, ...)` directly is an error.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional, Tuple

import os

import torch

from datasets.DataPartitioner import DataPartitioner
from mdgan_b200 import routing
from mdgan_b200.engine import EngineConfig
from mdgan_b200.node import run_node


def start(
    backend: str,
    rank: int,
    world_size: int,
    data_partitioner: DataPartitioner,
    discriminator_lr: float,
    generator_lr: float,
    epochs: int,
    swap_interval: int,
    local_epochs: int,
    log_interval: int,
    discriminator: torch.nn.Module,
    generator: torch.nn.Module,
    batch_size: int,
    image_shape: Tuple[int, int, int],
    log_folder: Path,
    dataset_name: str,
    device: torch.device = torch.device("cpu"),
    z_dim: int = 100,
    beta_1: float = 0.0,
    beta_2: float = 0.999,
    *,
    colocated_workers: Optional[Dict[int, torch.nn.Module]] = None,
    n_procs: Optional[int] = None,
    iid: bool = True,
) -> None:
    N = routing.num_workers(world_size)
    n_procs = n_procs or N
    proc = routing.process_of_worker(rank - 1, n_procs, N)
    if proc == 0:
        raise RuntimeError(f"worker rank {rank} shares GPU process 0 with the server: start it through "
                           "actors.server.start(colocated_workers=...)")
    hosted = routing.workers_of_process(proc, n_procs, N)
    discs = {rank - 1: discriminator}
    for r, m in (colocated_workers or {}).items():
        discs[r - 1] = m
    if sorted(discs) != hosted:
        raise RuntimeError(f"GPU process {proc} hosts worker ranks {[n + 1 for n in hosted]}, got {sorted(r + 1 for r in discs)}")
    cfg = EngineConfig(n_workers=N, batch_size=batch_size, z_dim=z_dim, image_shape=tuple(image_shape),
                       generator_lr=generator_lr, discriminator_lr=discriminator_lr, beta_1=beta_1, beta_2=beta_2,
                       swap_interval=swap_interval, local_epochs=local_epochs,
                       prefetch_host=os.environ.get("MDGAN_PREFETCH", "1") == "1")
    run_node(backend=backend, proc=proc, n_procs=n_procs, world_size=world_size, device=torch.device(device), cfg=cfg,
             generator=None, discriminators=discs, dataset=data_partitioner.train_dataset, epochs=epochs,
             log_interval=log_interval, log_folder=Path(log_folder), dataset_name=dataset_name, iid=iid)
