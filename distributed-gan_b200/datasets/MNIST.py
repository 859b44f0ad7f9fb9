"""MNIST plugin with the reference's MLP models (/root/reference/src/datasets/MNIST.py:74-120: four Linear
layers per net, LeakyReLU(0.2), always-active functional dropout 0.3 in the discriminator).

The B200 engine runs this family on its own kernels (mdgan_b200/mlp_nets.py, csrc/mlp.cu: fp32 SGEMM with the
bias / activation / dropout / gate tail fused).  The dropout masks are drawn on the host from each worker's torch RNG
stream in the reference's call order and uploaded with the iteration's inputs, so they are the reference's masks bit
for bit (tests/test_mlp_host.py, tests/test_mlp_gpu.py).  `--dataset MNIST_DCGAN` is the MNIST-shape DCGAN that
BASELINE.json's configs 1/2 name.
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
