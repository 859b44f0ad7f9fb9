"""Single-process classic GAN loop on the B200 kernels -- CLI-compatible with
/root/reference/src/standalone_gan.py:58-72 (the reference's comparison baseline, BASELINE.json config 1).

Per iteration (standalone_gan.py:180-227): real batch, fake = G(z); D step on BCE(D(real),1) + BCE(D(fake),0); then
the G step BCE(D(fake),1) back-propagated through the just-updated D into G.  With the engine's primitives that is
DiscNet.train_step(real, fake) -> DiscNet.feedback_step(fake) -> GenNet.backward(feedback) -> Adam.
The host RNG calls happen in the reference's order (model construction, DataLoader iterator, per-step randn), so a
run with the same --seed sees the same data order and noise.
"""
import argparse
import csv
import importlib
import random
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

COLUMNS = [
    "epoch", "start.epoch", "end.epoch", "start.epoch_calculation", "start.discriminator_train",
    "end.discriminator_train", "start.generator_train", "start.generate_data", "end.generate_data",
    "end.generator_train", "end.epoch_calculation", "start.calc_gradients", "end.calc_gradients", "absolut_step",
    "mean_d_loss", "mean_g_loss", "start.train", "end.train", "start.fid", "end.fid", "start.is", "end.is", "fid", "is",
]


def _weights_init(m: nn.Module) -> None:
    name = type(m).__name__
    if "Conv" in name:
        m.weight.data.normal_(0.0, 0.02)
    elif "BatchNorm" in name:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="cifar")
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--local_epochs", type=int, default=10)
    p.add_argument("--model", type=str, default="cifar")
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--log_interval", type=int, default=10)
    p.add_argument("--n_samples_fid", type=int, default=10)
    p.add_argument("--generator_lr", type=float, default=0.0002)
    p.add_argument("--discriminator_lr", type=float, default=0.0002)
    p.add_argument("--device", type=str, default="cpu")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--beta_1", type=float, default=0.0)
    p.add_argument("--beta_2", type=float, default=0.999)
    return p


class Standalone:
    def __init__(self, dataset_module, dataset, batch_size: int, device: torch.device, seed: int, generator_lr: float,
                 discriminator_lr: float, beta_1: float, beta_2: float, factory=None):
        from mdgan_b200.engine import CudaNetFactory, EngineConfig

        np.random.seed(seed)
        random.seed(seed)
        torch.manual_seed(seed)  # standalone_gan.py:74-80
        self.G = dataset_module.Generator()
        self.D = dataset_module.Discriminator()
        self.G.apply(_weights_init)
        self.D.apply(_weights_init)
        self.b, self.z_dim, self.device = batch_size, dataset_module.Z_DIM, device
        self.loader = DataLoader(dataset, batch_size=batch_size, shuffle=True)
        self.it = iter(self.loader)
        cfg = EngineConfig(n_workers=1, batch_size=batch_size, z_dim=self.z_dim, image_shape=tuple(dataset_module.SHAPE),
                           generator_lr=generator_lr, discriminator_lr=discriminator_lr, beta_1=beta_1, beta_2=beta_2)
        fac = factory or CudaNetFactory(device)
        self.gen = fac.generator(self.G, cfg, batch_size)
        self.disc = fac.discriminator(self.D, cfg)
        if getattr(self.disc, "stage_host", None) is not None:
            # MLP plugin (datasets/MNIST.py): one process, ONE global RNG -- the dropout draws of the three discriminator
            # forwards interleave with the noise draws on the same stream (standalone_gan.py:190-214)
            self.disc.rng = torch.default_generator
        self.real_dev = torch.empty((batch_size, *dataset_module.SHAPE), device=device)
        self.z_dev = torch.empty((batch_size, self.z_dim), device=device)

    def step(self):
        try:
            real = next(self.it)[0]
        except StopIteration:
            self.it = iter(self.loader)
            real = next(self.it)[0]
        self.real_dev.copy_(real, non_blocking=True)
        self.z_dev.copy_(torch.randn(self.b, self.z_dim, 1, 1).view(self.b, self.z_dim), non_blocking=True)
        if getattr(self.disc, "stage_host", None) is not None:
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()   # the previous step's mask upload has read the staging buffer
            self.disc.stage_host()
            self.disc.upload_host()
        fake = self.gen.forward(self.z_dev)
        d_loss = self.disc.train_step(self.real_dev, fake)
        g_loss = self.disc.feedback_step(fake)
        self.gen.backward(self.disc.feedback, 1.0)
        self.gen.adam()
        return d_loss, g_loss


def main(argv=None) -> None:
    args = build_parser().parse_args(argv)
    if not args.device.startswith("cuda"):
        raise RuntimeError(f"--device {args.device}: the B200 build runs on CUDA only; pass --device cuda")
    if args.local_epochs != 1:
        # the reference itself only works with 1 (the graph of `fake` is freed after the first errG.backward())
        raise ValueError("standalone_gan supports --local_epochs 1 only")
    device = torch.device(args.device)
    dataset_module = importlib.import_module(f"datasets.{args.dataset}")
    partitioner = dataset_module.Partitioner(0, 0)
    partitioner.load_data()
    run = Standalone(dataset_module, partitioner.train_dataset, args.batch_size, device, args.seed, args.generator_lr,
                     args.discriminator_lr, args.beta_1, args.beta_2)
    logs = Path("logs")
    logs.mkdir(parents=True, exist_ok=True)
    with open(logs / f"{args.dataset}.standalone.logs.csv", "a", encoding="utf-8") as f:
        w = csv.DictWriter(f, fieldnames=COLUMNS)
        w.writeheader()
        for epoch in range(args.epochs):
            row = {c: None for c in COLUMNS}
            t0 = time.time()
            d_loss, g_loss = run.step()
            row.update({"epoch": epoch, "start.epoch": t0, "start.epoch_calculation": t0, "start.train": t0,
                        "absolut_step": epoch * args.local_epochs, "mean_d_loss": d_loss.item(),
                        "mean_g_loss": g_loss.item()})
            row["end.epoch_calculation"] = row["end.epoch"] = time.time()
            print(f"Epoch {epoch}, Step 0, Loss D {row['mean_d_loss']}, Loss G {row['mean_g_loss']}")
            w.writerow(row)
    run.gen.state.store_to(run.G)
    run.disc.state.store_to(run.D)
    weights = Path("weights")
    weights.mkdir(parents=True, exist_ok=True)
    torch.save(run.G.state_dict(), weights / f"netG_epoch_{args.epochs - 1}.pth")
    torch.save(run.D.state_dict(), weights / f"netD_epoch_{args.epochs - 1}.pth")


if __name__ == "__main__":
    main()
