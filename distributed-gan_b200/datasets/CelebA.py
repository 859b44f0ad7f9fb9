"""CelebA plugin: 3x64x64 DCGAN with the reference's attribute names (cv1..cv5 / bn2..bn4, tconv1..tconv5 /
bn1..bn4), its conv biases on cv2/cv3 and the slope-0.01 first LeakyReLU
(/root/reference/src/datasets/CelebA.py:75-142)."""
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from datasets.DataPartitioner import TorchvisionPartitioner

SHAPE: Tuple[int, int, int] = (3, 64, 64)
NDF: int = 64
NGF: int = 64
Z_DIM: int = 100


def _load(path: str, train: bool):
    from torchvision import transforms
    from torchvision.datasets import CelebA

    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3),
                             transforms.Resize((64, 64))])
    return CelebA(root=path, split="train" if train else "test", download=False, transform=tf)


class Partitioner(TorchvisionPartitioner):
    def __init__(self, world_size: int, rank: int, path: str = "data/celeba"):
        super().__init__(world_size, rank, path, SHAPE, _load)


class Discriminator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        c = [SHAPE[0], NDF, NDF * 2, NDF * 4, NDF * 8]
        self.cv1 = nn.Conv2d(c[0], c[1], 4, 2, 1, bias=False)   # 64 -> 32
        self.cv2 = nn.Conv2d(c[1], c[2], 4, 2, 1)               # 32 -> 16 (bias kept, as in the reference)
        self.bn2 = nn.BatchNorm2d(c[2])
        self.cv3 = nn.Conv2d(c[2], c[3], 4, 2, 1)               # 16 -> 8
        self.bn3 = nn.BatchNorm2d(c[3])
        self.cv4 = nn.Conv2d(c[3], c[4], 4, 2, 1, bias=False)   # 8 -> 4
        self.bn4 = nn.BatchNorm2d(c[4])
        self.cv5 = nn.Conv2d(c[4], 1, 4, 1, 0, bias=False)      # 4 -> 1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = F.leaky_relu(self.cv1(x))  # default slope 0.01
        for cv, bn in ((self.cv2, self.bn2), (self.cv3, self.bn3), (self.cv4, self.bn4)):
            h = F.leaky_relu(bn(cv(h)), 0.2)
        return torch.sigmoid(self.cv5(h)).view(-1, 1).squeeze(1)


class Generator(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        c = [NGF * 8, NGF * 4, NGF * 2, NGF]
        self.tconv1 = nn.ConvTranspose2d(Z_DIM, c[0], 4, 1, 0, bias=False)  # 1 -> 4
        self.bn1 = nn.BatchNorm2d(c[0])
        self.tconv2 = nn.ConvTranspose2d(c[0], c[1], 4, 2, 1, bias=False)   # 4 -> 8
        self.bn2 = nn.BatchNorm2d(c[1])
        self.tconv3 = nn.ConvTranspose2d(c[1], c[2], 4, 2, 1, bias=False)   # 8 -> 16
        self.bn3 = nn.BatchNorm2d(c[2])
        self.tconv4 = nn.ConvTranspose2d(c[2], c[3], 4, 2, 1, bias=False)   # 16 -> 32
        self.bn4 = nn.BatchNorm2d(c[3])
        self.tconv5 = nn.ConvTranspose2d(c[3], SHAPE[0], 4, 2, 1, bias=False)  # 32 -> 64

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        h = z
        for tc, bn in ((self.tconv1, self.bn1), (self.tconv2, self.bn2), (self.tconv3, self.bn3), (self.tconv4, self.bn4)):
            h = F.relu(bn(tc(h)))
        return torch.tanh(self.tconv5(h))
