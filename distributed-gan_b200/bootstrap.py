"""Launcher of the distributed MD-GAN run -- CLI-compatible with /root/reference/src/bootstrap.py:30-51.

Every reference flag is accepted with the same name and default.  Differences forced by the B200 layout
(SURVEY.md H3), all additive:
  * one OS process per GPU (NCCL), not one per actor: `--ranks 0..N` still names the N+1 actors, which are placed
    on `--gpus` processes (default min(N, visible GPUs)); the server shares process 0 with worker 1;
  * `--device` must be cuda (the reference default "cpu" is refused: there is no CPU fallback);
  * `--world_size 2` (N = 1, the "K=1" baseline) is allowed; otherwise world_size must be odd as in the reference
    (bootstrap.py:163-164) because the discriminator swap pairs workers;
  * `--z_source host|device`, `--precision tf32x3|tf32`, `--synthetic M` are new optional flags.
"""
import argparse
import importlib
import logging
import os
import random
import sys
from pathlib import Path
from typing import Dict, List

import numpy as np
import torch
import torch.multiprocessing as mp
import torch.nn as nn

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from mdgan_b200 import routing  # noqa: E402


def _weights_init(m: nn.Module) -> None:
    """DCGAN init by class name (reference bootstrap.py:17-27)."""
    name = type(m).__name__
    if "Conv" in name:
        m.weight.data.normal_(0.0, 0.02)
    elif "BatchNorm" in name:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--backend", type=str, default="nccl")
    p.add_argument("--world_size", type=int, default=2)
    p.add_argument("--dataset", type=str, default="cifar")
    p.add_argument("--ranks", type=str, default="0,1,2")
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--swap_interval", type=int, default=1)
    p.add_argument("--local_epochs", type=int, default=10)
    p.add_argument("--model", type=str, default="cifar")
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--log_interval", type=int, default=50)
    p.add_argument("--generator_lr", type=float, default=0.001)
    p.add_argument("--discriminator_lr", type=float, default=0.004)
    p.add_argument("--device", type=str, default="cpu")
    p.add_argument("--master_addr", type=str, default="localhost")
    p.add_argument("--master_port", type=str, default="1234")
    p.add_argument("--network_interface", type=str, required=False)
    p.add_argument("--iid", type=int, default=1)
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--beta_1", type=float, default=0.0)
    p.add_argument("--beta_2", type=float, default=0.999)
    # additions
    p.add_argument("--gpus", type=int, default=0, help="GPU processes (default: min(workers, visible GPUs))")
    p.add_argument("--z_source", type=str, default="host", choices=["host", "device"])
    p.add_argument("--precision", type=str, default=None, choices=["tf32x3", "tf32"])
    p.add_argument("--synthetic", type=int, default=0, help="train on M synthetic images (sets MDGAN_SYNTH_M)")
    return p


def _seed_actor(seed: int) -> None:
    """Per-actor seeding of the reference (bootstrap.py:138-145): every actor owns a stream seeded --seed + rank."""
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def init_process(proc: int, args, n_procs: int, partitioner, image_shape, z_dim, generator_cls, discriminator_cls) -> None:
    from actors import server, worker

    N = routing.num_workers(args.world_size)
    os.environ["RANK"] = str(proc)
    partitioner.rank = proc
    log_folder = Path("logs")
    log_folder.mkdir(parents=True, exist_ok=True)
    hosted = routing.workers_of_process(proc, n_procs, N)
    discs: Dict[int, nn.Module] = {}
    for n in hosted:  # each worker actor builds its model on its own seed (bootstrap.py:75-76)
        _seed_actor(args.seed + n + 1)
        d = discriminator_cls().to(dtype=torch.float32)
        d.apply(_weights_init)
        # the worker actor's global RNG stream goes on from here in the reference (its dropout draws, MNIST.py:89-94);
        # several actors share this process, so the stream's state travels with the model (mlp_nets.MlpDiscNet)
        d._mdgan_rng_state = torch.get_rng_state()
        discs[n + 1] = d
    device = torch.device(args.device)
    if proc == 0:
        _seed_actor(args.seed + 0)  # the server's stream stays live: noise and swap pairs keep drawing from it
        g = generator_cls().to(dtype=torch.float32)
        g.apply(_weights_init)
        server.start(backend=args.backend, rank=0, world_size=args.world_size, batch_size=args.batch_size,
                     epochs=args.epochs, generator=g, dataset=partitioner.train_dataset, device=device,
                     image_shape=image_shape, generator_lr=args.generator_lr, z_dim=z_dim,
                     log_interval=args.log_interval, log_folder=log_folder, iid=args.iid == 1,
                     dataset_name=args.dataset, swap_interval=args.swap_interval, beta_1=args.beta_1,
                     beta_2=args.beta_2, colocated_workers=discs, discriminator_lr=args.discriminator_lr,
                     local_epochs=args.local_epochs, n_procs=n_procs, z_source=args.z_source)
    else:
        first = hosted[0] + 1
        others = {r: m for r, m in discs.items() if r != first}
        worker.start(backend=args.backend, rank=first, world_size=args.world_size, batch_size=args.batch_size,
                     swap_interval=args.swap_interval, data_partitioner=partitioner, epochs=args.epochs,
                     discriminator=discs[first], device=device, local_epochs=args.local_epochs,
                     image_shape=image_shape, log_interval=args.log_interval, generator=None,
                     discriminator_lr=args.discriminator_lr, generator_lr=args.generator_lr, z_dim=z_dim,
                     log_folder=log_folder, dataset_name=args.dataset, beta_1=args.beta_1, beta_2=args.beta_2,
                     colocated_workers=others, n_procs=n_procs, iid=args.iid == 1)


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(message)s")
    os.environ["MASTER_ADDR"] = args.master_addr
    os.environ["MASTER_PORT"] = args.master_port
    if args.network_interface:
        os.environ["NCCL_SOCKET_IFNAME"] = args.network_interface
    if args.precision:
        os.environ["MDGAN_PRECISION"] = args.precision
    if args.synthetic:
        os.environ["MDGAN_SYNTH_M"] = str(args.synthetic)

    ranks: List[int] = routing.parse_ranks(args.ranks)
    N = routing.num_workers(args.world_size)
    if args.world_size % 2 == 0 and N != 1:
        raise ValueError("World size must be odd")
    if sorted(ranks) != list(range(args.world_size)):
        raise ValueError(f"--ranks must name all {args.world_size} actors 0..{N} of this single-node job "
                         "(multi-node placement is out of scope for the B200 build)")
    if not args.device.startswith("cuda"):
        raise RuntimeError(f"--device {args.device}: the B200 build runs the training step on CUDA only; "
                           "pass --device cuda (there is no CPU fallback)")
    n_gpus = torch.cuda.device_count()
    if n_gpus == 0:
        raise RuntimeError("no CUDA device visible")
    n_procs = args.gpus or min(N, n_gpus)
    os.environ["WORLD_SIZE"] = str(n_procs)

    dataset_module = importlib.import_module(f"datasets.{args.dataset}")
    partitioner = dataset_module.Partitioner(args.world_size, 0)
    partitioner.load_data()
    spawn_args = (args, n_procs, partitioner, dataset_module.SHAPE, dataset_module.Z_DIM, dataset_module.Generator,
                  dataset_module.Discriminator)
    if n_procs == 1:
        init_process(0, *spawn_args)
    else:
        mp.spawn(init_process, args=spawn_args, nprocs=n_procs, join=True)


if __name__ == "__main__":
    main()
