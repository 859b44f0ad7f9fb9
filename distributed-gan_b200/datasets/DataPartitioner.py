"""Partitioner contract of the dataset plugins (mirror of /root/reference/src/datasets/DataPartitioner.py:6-59).

A plugin module `datasets.<NAME>` exposes `Partitioner(world_size, rank)`, zero-argument `Generator` and
`Discriminator` classes, `SHAPE` and `Z_DIM` (consumed by bootstrap.py exactly like the reference's
bootstrap.py:167-180).  `TorchvisionPartitioner` implements the contract once for all three datasets; when the
torchvision files are not on disk (this project has no network access) or MDGAN_SYNTH_M is set it serves the
synthetic U(-1,1) images described in BASELINE.md section 2.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch
import torch.utils.data


class DataPartitioner:
    def __init__(self, world_size: int, rank: int):
        raise NotImplementedError

    def get_subset_from_indices(self, indices: List[int], train: bool = True) -> torch.utils.data.Subset:
        raise NotImplementedError

    def load_data(self) -> None:
        raise NotImplementedError

    def shuffle(self) -> None:
        raise NotImplementedError

    def get_train_partition(self, partition_id: int) -> Tuple[torch.utils.data.Subset, int, int]:
        raise NotImplementedError

    def get_test_partition(self, partition_id: int) -> Tuple[torch.utils.data.Subset, int, int]:
        raise NotImplementedError

    @property
    def train_dataset(self) -> torch.utils.data.Dataset:
        raise NotImplementedError

    @property
    def test_dataset(self) -> torch.utils.data.Dataset:
        raise NotImplementedError


def _get_partition(world_size: int, partition_id: int, dataset) -> Tuple[torch.utils.data.Subset, int, int]:
    """Contiguous slice `partition_id` of `world_size` (reference DataPartitioner.py:62-76, which hard-codes a
    1000-sample universe at :66; kept for interface parity, the actors never call it)."""
    size = 1000
    length = size // world_size
    start = partition_id * length
    end = start + length
    if partition_id == world_size - 1 and size - end > 0:
        end = size
    return torch.utils.data.Subset(dataset, range(start, end)), start, end


class SyntheticImages(torch.utils.data.Dataset):
    """M images ~ U(-1,1), fp32, from a private generator seeded 1234; label 0."""

    def __init__(self, shape: Tuple[int, int, int], m: int):
        g = torch.Generator().manual_seed(1234)
        self.data = torch.rand((m, *shape), generator=g) * 2 - 1

    def __len__(self) -> int:
        return self.data.shape[0]

    def __getitem__(self, i):
        return self.data[int(i)], 0


class TorchvisionPartitioner(DataPartitioner):
    def __init__(self, world_size: int, rank: int, path: str, shape: Tuple[int, int, int],
                 loader: Optional[Callable[[str, bool], torch.utils.data.Dataset]] = None):
        self.world_size, self.rank, self.path, self.shape = world_size, rank, path, shape
        self._loader = loader
        self._train = None
        self._test = None

    def load_data(self) -> None:
        m = os.environ.get("MDGAN_SYNTH_M")
        if m is None and self._loader is not None:
            try:
                self._train = self._loader(self.path, True)
                self._test = self._loader(self.path, False)
                return
            except Exception as e:  # no files on disk and no network
                raise RuntimeError(
                    f"dataset files not found under {self.path} ({e}); set MDGAN_SYNTH_M=<samples> to train on "
                    "synthetic images") from e
        m = int(m or 1024)
        self._train = SyntheticImages(self.shape, m)
        self._test = SyntheticImages(self.shape, max(m // 8, 1))

    def get_subset_from_indices(self, indices, train: bool = True) -> torch.utils.data.Subset:
        return torch.utils.data.Subset(self._train if train else self._test, indices)

    def shuffle(self) -> None:
        self._train = torch.utils.data.Subset(self._train, torch.randperm(len(self._train)))
        self._test = torch.utils.data.Subset(self._test, torch.randperm(len(self._test)))

    def get_train_partition(self, partition_id: int):
        return _get_partition(self.world_size, partition_id, self._train)

    def get_test_partition(self, partition_id: int):
        return _get_partition(self.world_size, partition_id, self._test)

    @property
    def train_dataset(self):
        return self._train

    @property
    def test_dataset(self):
        return self._test
