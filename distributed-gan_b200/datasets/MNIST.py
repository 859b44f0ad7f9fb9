"""MNIST plugin with the reference's MLP models (/root/reference/src/datasets/MNIST.py:74-120: four Linear
layers, LeakyReLU(0.2), always-active functional dropout 0.3).

The plugin loads and its models run under stock PyTorch (the oracle / CPU baseline use them), but the B200 engine
covers the DCGAN conv family only and refuses this model loudly (`UnsupportedModelError`, no fallback): the MLP is
not a dense-conv hot path and its dropout masks come from each worker's global RNG stream, which has no
bit-compatible device equivalent.  Use `--dataset MNIST_DCGAN` for MNIST-shape runs on the GPU.
"""
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from datasets.DataPartitioner import TorchvisionPartitioner
from datasets.MNIST_DCGAN import _load

SHAPE: Tuple[int, int, int] = (1, 28, 28)
NDF: int = 64
NGF: int = 64
Z_DIM: int = 100
_PIXELS = SHAPE[0] * SHAPE[1] * SHAPE[2]


class Partitioner(TorchvisionPartitioner):
    def __init__(self, world_size: int, rank: int, path: str = "data/mnist"):
        super().__init__(world_size, rank, path, SHAPE, _load)


class Discriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(_PIXELS, 1024)
        self.fc2 = nn.Linear(1024, 512)
        self.fc3 = nn.Linear(512, 256)
        self.fc4 = nn.Linear(256, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = x.view(x.shape[0], -1)
        for fc in (self.fc1, self.fc2, self.fc3):
            h = F.dropout(F.leaky_relu(fc(h), 0.2), 0.3)  # training=True by default: always active
        return torch.sigmoid(self.fc4(h)).flatten()


class Generator(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(Z_DIM, 256)
        self.fc2 = nn.Linear(256, 512)
        self.fc3 = nn.Linear(512, 1024)
        self.fc4 = nn.Linear(1024, _PIXELS)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        h = z.view(z.shape[0], -1)
        for fc in (self.fc1, self.fc2, self.fc3):
            h = F.leaky_relu(fc(h), 0.2)
        return torch.tanh(self.fc4(h)).view(-1, *SHAPE)
