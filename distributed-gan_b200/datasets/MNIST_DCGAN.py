"""MNIST-shape (1x28x28) DCGAN plugin.

BASELINE.json config 1/2 ask for a "DCGAN on synthetic MNIST-shape 1x28x28", which the reference does not ship
(its datasets/MNIST.py:74-120 is an MLP with always-on dropout).  This plugin follows the same
contract (Partitioner / Generator / Discriminator / SHAPE / Z_DIM); its oracle is this same nn.Module run by
stock PyTorch inside the oracle loops.  7 -> 14 -> 28 generator, mirrored discriminator.
"""
from typing import Tuple

import torch
from torch import nn

from datasets.DataPartitioner import TorchvisionPartitioner

SHAPE: Tuple[int, int, int] = (1, 28, 28)
NDF: int = 64
NGF: int = 64
Z_DIM: int = 100


def _load(path: str, train: bool):
    from torchvision import transforms
    from torchvision.datasets import MNIST

    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=(0.5,), std=(0.5,))])
    return MNIST(root=path, train=train, download=False, transform=tf)


class Partitioner(TorchvisionPartitioner):
    def __init__(self, world_size: int, rank: int, path: str = "data/mnist"):
        super().__init__(world_size, rank, path, SHAPE, _load)


class Discriminator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.main = nn.Sequential(
            nn.Conv2d(1, NDF, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True),                 # 28 -> 14
            nn.Conv2d(NDF, NDF * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(NDF * 2), nn.LeakyReLU(0.2, inplace=True),  # 7
            nn.Conv2d(NDF * 2, 1, 7, 1, 0, bias=False), nn.Sigmoid(),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.main(x).view(-1, 1).squeeze(1)


class Generator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.main = nn.Sequential(
            nn.ConvTranspose2d(Z_DIM, NGF * 4, 7, 1, 0, bias=False), nn.BatchNorm2d(NGF * 4), nn.ReLU(True),        # 7
            nn.ConvTranspose2d(NGF * 4, NGF * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(NGF * 2), nn.ReLU(True),      # 14
            nn.ConvTranspose2d(NGF * 2, 1, 4, 2, 1, bias=False), nn.Tanh(),                                       # 28
        )

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        return self.main(z)
