"""CIFAR-10 plugin: 3x32x32 DCGAN (same architecture and state_dict keys `main.<i>.*` as
/root/reference/src/datasets/CIFAR10.py:76-140, so checkpoints are interchangeable)."""
from typing import Tuple

import torch
from torch import nn

from datasets.DataPartitioner import TorchvisionPartitioner

SHAPE: Tuple[int, int, int] = (3, 32, 32)
NDF: int = 64
NGF: int = 64
Z_DIM: int = 100


def _load(path: str, train: bool):
    from torchvision import transforms
    from torchvision.datasets import CIFAR10

    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    return CIFAR10(root=path, train=train, download=False, transform=tf)


class Partitioner(TorchvisionPartitioner):
    def __init__(self, world_size: int, rank: int, path: str = "data/cifar10"):
        super().__init__(world_size, rank, path, SHAPE, _load)


class Discriminator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        widths = [SHAPE[0], NDF, NDF * 2, NDF * 4]
        seq = []
        for i in range(3):  # 32 -> 16 -> 8 -> 4
            seq.append(nn.Conv2d(widths[i], widths[i + 1], 4, 2, 1, bias=False))
            if i > 0:
                seq.append(nn.BatchNorm2d(widths[i + 1]))
            seq.append(nn.LeakyReLU(0.2, inplace=True))
        seq += [nn.Conv2d(widths[-1], 1, 4, 1, 0, bias=False), nn.Sigmoid()]
        self.main = nn.Sequential(*seq)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.main(x).view(-1, 1).squeeze(1)


class Generator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        widths = [NGF * 8, NGF * 4, NGF * 2]
        seq = [nn.ConvTranspose2d(Z_DIM, widths[0], 4, 1, 0, bias=False), nn.BatchNorm2d(widths[0]), nn.ReLU(True)]
        for i in range(2):  # 4 -> 8 -> 16
            seq += [nn.ConvTranspose2d(widths[i], widths[i + 1], 4, 2, 1, bias=False), nn.BatchNorm2d(widths[i + 1]),
                    nn.ReLU(True)]
        seq += [nn.ConvTranspose2d(widths[-1], SHAPE[0], 4, 2, 1, bias=False), nn.Tanh()]  # 16 -> 32
        self.main = nn.Sequential(*seq)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        return self.main(z)
