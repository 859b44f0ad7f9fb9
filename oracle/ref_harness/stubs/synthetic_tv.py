"""Picklable in-memory synthetic replacement for torchvision's MNIST / CIFAR10 /
CelebA dataset classes (no network in the build container).  Data recipe from
BASELINE.md section 2: U(-1,1) fp32 images from a private generator seeded 1234,
label 0, M = $MDGAN_SYNTH_M samples.  Test infrastructure only.
"""
import os

import torch
import torch.utils.data

_SHAPES = {"MNIST": (1, 28, 28), "CIFAR10": (3, 32, 32), "CelebA": (3, 64, 64)}


class _Synthetic(torch.utils.data.Dataset):
    KIND = "CIFAR10"

    def __init__(self, root=None, train=True, download=False, transform=None, split="train", **kw):
        m = int(os.environ.get("MDGAN_SYNTH_M", "256"))
        g = torch.Generator().manual_seed(1234)
        self.data = torch.rand((m, *_SHAPES[self.KIND]), generator=g) * 2 - 1

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        return self.data[int(i)], 0


class MNIST(_Synthetic):
    KIND = "MNIST"


class CIFAR10(_Synthetic):
    KIND = "CIFAR10"


class CelebA(_Synthetic):
    KIND = "CelebA"
