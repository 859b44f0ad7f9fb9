"""Real-data ingest on CPU: the dataset plugins read torchvision-format files from disk
through the REAL torchvision dataset class and transforms (MNIST idx files written by tests/mnist_files.py -- there is
no network to fetch the originals), and the shard / batch order on top of them is the reference's, bit for bit
(/root/reference/src/actors/server.py:46-64,152-167, worker.py:70-89,162-167, datasets/MNIST.py:39-47)."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

from mnist_files import write_mnist_idx
from util import plugin

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle.ref_harness import time_reference  # noqa: E402


@pytest.mark.parametrize("name", ["MNIST_DCGAN", "MNIST"])
def test_mnist_files_through_torchvision(tmp_path, monkeypatch, name):
    import torchvision

    from mdgan_b200 import routing

    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)
    root = write_mnist_idx(tmp_path / "data" / "mnist", 192, 40)
    p = plugin(name).Partitioner(3, 0, path=str(root))
    p.load_data()
    assert isinstance(p.train_dataset, torchvision.datasets.MNIST) and isinstance(p.test_dataset, torchvision.datasets.MNIST)
    assert len(p.train_dataset) == 192 and len(p.test_dataset) == 40
    x, y = p.train_dataset[5]
    raw = p.train_dataset.data[5].float() / 255.0                      # ToTensor, then Normalize(0.5, 0.5)
    assert x.dtype == torch.float32 and tuple(x.shape) == (1, 28, 28) and torch.equal(x[0], (raw - 0.5) / 0.5)
    assert x.min() >= -1 and x.max() <= 1 and 0 <= int(y) <= 9
    # the worker's loader (worker.py:78-89,162-167) on the server's shard (server.py:46-64,152-154), iid and non-iid
    for iid in (True, False):
        shards = routing.split_dataset(len(p.train_dataset), 2, iid)
        if not iid:
            assert shards[0].tolist() == list(range(96)) and shards[1].tolist() == list(range(96, 192))
        shard = shards[1]
        g = torch.Generator()
        g.manual_seed(0)
        ref = torch.utils.data.DataLoader(p.get_subset_from_indices(shard), batch_size=16, shuffle=True, generator=g)
        expect = [b[0] for b in ref] + [b[0] for b in ref][:2]         # one epoch (6 batches) + into the next
        for nw in (0, 2):
            stream = routing.RealBatchStream(p.train_dataset, shard, 16, num_workers=nw)
            for e in expect:
                assert torch.equal(stream.next(), e)
            del stream


def test_missing_files_fail_loudly(tmp_path, monkeypatch):
    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)
    p = plugin("MNIST_DCGAN").Partitioner(3, 0, path=str(tmp_path / "nothing_here"))
    with pytest.raises(RuntimeError, match="MDGAN_SYNTH_M"):
        p.load_data()


@pytest.mark.skipif(time_reference.reference_src() is None, reason="no reference sources on this host")
def test_mnist_partitioner_equals_the_references(tmp_path, monkeypatch):
    """Same files, the reference's own Partitioner class (datasets/MNIST.py:19-47, loaded from the unmodified copy)
    against this repo's: every train and test sample identical."""
    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)
    root = write_mnist_idx(tmp_path / "data" / "mnist", 64, 16)
    src = Path(time_reference.reference_src()) / "datasets" / "MNIST.py"
    plugin("MNIST")                                                    # `datasets.DataPartitioner` resolves to the package
    spec = importlib.util.spec_from_file_location("reference_datasets_MNIST", src)
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    ref = ref_mod.Partitioner(3, 0, path=str(root))
    ref.load_data()                                                    # download=True finds the files and fetches nothing
    for name in ("MNIST", "MNIST_DCGAN"):
        ours = plugin(name).Partitioner(3, 0, path=str(root))
        ours.load_data()
        for a, b in ((ours.train_dataset, ref.train_dataset), (ours.test_dataset, ref.test_dataset)):
            assert len(a) == len(b)
            for i in range(len(a)):
                assert torch.equal(a[i][0], b[i][0]) and a[i][1] == b[i][1]
        sub_a, sub_b = ours.get_subset_from_indices([3, 1, 2]), ref.get_subset_from_indices([3, 1, 2])
        assert all(torch.equal(sub_a[i][0], sub_b[i][0]) for i in range(3))
