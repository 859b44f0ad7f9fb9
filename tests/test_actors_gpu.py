"""The reference-facing entry points on a GPU: bootstrap.py (server + workers through actors.server.start) and
standalone_gan.py, run with the reference's flags on synthetic data, checked against the oracle and for the
reference's output files (CSV schemas, weight files with the reference's state_dict keys)."""
import csv
import os

import pytest
import torch

from parity import FREE_TOL, l2err
from util import plugin

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gpus", [1, 2])
def test_bootstrap_matches_oracle(tmp_path, monkeypatch, gpus):
    """gpus = 1: server + both workers in one process; gpus = 2 (needs two GPUs): one process per GPU spawned by
    bootstrap.main exactly like the reference's mp.spawn, peer-memory exchange, phase graphs, CSV spans from events."""
    import bootstrap

    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200.node import SERVER_COLUMNS, WORKER_COLUMNS
    from oracle.mdgan_oracle import OracleMDGAN

    N, b, epochs, m = 2, 8, 5, 2 * 4 * 8
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("MDGAN_PRECISION", "tf32x3")
    bootstrap.main(["--backend", "nccl", "--world_size", str(N + 1), "--ranks", f"0..{N}", "--dataset", "CIFAR10",
                    "--epochs", str(epochs), "--local_epochs", "1", "--swap_interval", "2", "--device", "cuda",
                    "--batch_size", str(b), "--iid", "1", "--seed", "3", "--beta_1", "0.5", "--generator_lr", "0.0002",
                    "--discriminator_lr", "0.0002", "--log_interval", "1000", "--gpus", str(gpus), "--synthetic", str(m),
                    "--master_addr", "127.0.0.1", "--master_port", "29533"])
    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)
    mod = plugin("CIFAR10")
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, m), N, b, mod.Z_DIM, mod.SHAPE,
                         seed=3, beta_1=0.5, swap_interval=2)
    ref = [oracle.step(e, record=False) for e in range(epochs)]

    g = torch.load(tmp_path / "weights" / "generator_final.pt")
    assert list(g.keys()) == list(oracle.G.state_dict().keys())
    flat = lambda sd: torch.cat([v.reshape(-1).double() for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k])
    assert l2err(flat(g), flat(oracle.G.state_dict())) <= FREE_TOL["weights_l2"]
    for n in range(N):
        d = torch.load(tmp_path / "weights" / f"worker_{n + 1}" / "discriminator.pth")
        rd = oracle.D[n].state_dict()
        assert list(d.keys()) == list(rd.keys())
        assert l2err(flat(d), flat(rd)) <= FREE_TOL["weights_l2"]
        assert all(int(d[k]) == int(rd[k]) == 3 * epochs for k in d if k.endswith("num_batches_tracked"))
        rows = list(csv.DictReader(open(tmp_path / "logs" / f"mdgan.{N}.CIFAR10.worker.{n + 1}.logs.csv")))
        assert list(rows[0].keys()) == WORKER_COLUMNS and len(rows) == epochs
        for e, row in enumerate(rows):
            assert abs(float(row["mean_d_loss"]) - ref[e]["mean_d_loss"][n]) <= FREE_TOL["loss"] * abs(ref[e]["mean_d_loss"][n])
            partner = None
            if ref[e]["pairs"] is not None:
                pairs = {a: c for a, c in ref[e]["pairs"].tolist()}
                pairs.update({c: a for a, c in ref[e]["pairs"].tolist()})
                partner = pairs[n + 1]
            assert (row["swap_with"] == "" and partner is None) or int(row["swap_with"]) == partner
        assert float(rows[0]["size.model"]) == 2.5332183837890625
    srows = list(csv.DictReader(open(tmp_path / "logs" / f"mdgan.{N}.CIFAR10.server.logs.csv")))
    assert list(srows[0].keys()) == SERVER_COLUMNS and len(srows) == epochs
    assert [r["swap"] for r in srows] == ["False", "False", "True", "False", "True"]
    for r in srows:  # device-event spans: ordered phases of positive length
        t = [float(r[c]) for c in ("start.epoch_calculation", "end.generate_data", "end.recv_data", "end.agg_gradients",
                                   "end.epoch_calculation")]
        assert all(b_ >= a_ for a_, b_ in zip(t, t[1:])) and t[-1] > t[0]
    assert (tmp_path / "saved_images" / "real_images.png").exists()
    assert (tmp_path / "saved_images" / f"generated_epoch_{epochs - 1}.png").exists()
    # the asynchronous snapshot of the last epoch (side-stream copy + writer thread) holds exactly the final state
    g_last = torch.load(tmp_path / "weights" / f"generator_{epochs - 1}.pt")
    assert list(g_last.keys()) == list(g.keys())
    assert all(torch.equal(g_last[k], g[k]) and g_last[k].dtype == g[k].dtype and g_last[k].shape == g[k].shape for k in g)


def test_standalone_matches_oracle(tmp_path, monkeypatch):
    import standalone_gan
    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleStandalone

    b, epochs, m = 16, 3, 64
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("MDGAN_SYNTH_M", str(m))
    standalone_gan.main(["--dataset", "CIFAR10", "--epochs", str(epochs), "--local_epochs", "1", "--batch_size", str(b),
                         "--device", "cuda", "--seed", "1", "--beta_1", "0.5", "--log_interval", "1000"])
    mod = plugin("CIFAR10")
    oracle = OracleStandalone(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, m), b, mod.Z_DIM, seed=1,
                              beta_1=0.5)
    ref = [oracle.step() for _ in range(epochs)]
    rows = list(csv.DictReader(open(tmp_path / "logs" / "CIFAR10.standalone.logs.csv")))
    assert len(rows) == epochs
    for row, r in zip(rows, ref):
        assert abs(float(row["mean_d_loss"]) - r["mean_d_loss"]) <= FREE_TOL["loss"] * abs(r["mean_d_loss"])
        assert abs(float(row["mean_g_loss"]) - r["mean_g_loss"]) <= FREE_TOL["loss"] * abs(r["mean_g_loss"])
    g = torch.load(tmp_path / "weights" / f"netG_epoch_{epochs - 1}.pth")
    d = torch.load(tmp_path / "weights" / f"netD_epoch_{epochs - 1}.pth")
    flat = lambda sd: torch.cat([v.reshape(-1).double() for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k])
    assert list(g.keys()) == list(oracle.G.state_dict().keys()) and list(d.keys()) == list(oracle.D.state_dict().keys())
    assert l2err(flat(g), flat(oracle.G.state_dict())) <= FREE_TOL["weights_l2"]
    assert l2err(flat(d), flat(oracle.D.state_dict())) <= FREE_TOL["weights_l2"]


def test_bootstrap_on_mnist_files(tmp_path, monkeypatch):
    """Real-data ingest end to end: bootstrap.py trains from MNIST-format files on disk through the
    real torchvision.datasets.MNIST class (tests/mnist_files.py writes them; no --synthetic), streamed host batches,
    early upload, phase graphs, a swap -- against the oracle on the same dataset object.  The shard is 3 batches long, so
    the 4 iterations also cross the loader's epoch boundary (worker.py:162-167)."""
    import bootstrap
    import torchvision

    from mnist_files import write_mnist_idx
    from oracle.mdgan_oracle import OracleMDGAN

    N, b, epochs = 2, 8, 4
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)
    monkeypatch.setenv("MDGAN_PRECISION", "tf32x3")
    write_mnist_idx(tmp_path / "data" / "mnist", N * 3 * b, 16)
    bootstrap.main(["--backend", "nccl", "--world_size", str(N + 1), "--ranks", f"0..{N}", "--dataset", "MNIST_DCGAN",
                    "--epochs", str(epochs), "--local_epochs", "1", "--swap_interval", "2", "--device", "cuda",
                    "--batch_size", str(b), "--iid", "1", "--seed", "3", "--beta_1", "0.5", "--generator_lr", "0.0002",
                    "--discriminator_lr", "0.0002", "--log_interval", "1000", "--gpus", "1",
                    "--master_addr", "127.0.0.1", "--master_port", "29534"])
    mod = plugin("MNIST_DCGAN")
    part = mod.Partitioner(N + 1, 0)
    part.load_data()
    assert isinstance(part.train_dataset, torchvision.datasets.MNIST) and len(part.train_dataset) == N * 3 * b
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, part.train_dataset, N, b, mod.Z_DIM, mod.SHAPE, seed=3,
                         beta_1=0.5, swap_interval=2)
    ref = [oracle.step(e, record=False) for e in range(epochs)]
    flat = lambda sd: torch.cat([v.reshape(-1).double() for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k])
    g = torch.load(tmp_path / "weights" / "generator_final.pt")
    assert l2err(flat(g), flat(oracle.G.state_dict())) <= FREE_TOL["weights_l2"]
    for n in range(N):
        d = torch.load(tmp_path / "weights" / f"worker_{n + 1}" / "discriminator.pth")
        assert l2err(flat(d), flat(oracle.D[n].state_dict())) <= FREE_TOL["weights_l2"]
        rows = list(csv.DictReader(open(tmp_path / "logs" / f"mdgan.{N}.MNIST_DCGAN.worker.{n + 1}.logs.csv")))
        assert len(rows) == epochs
        for e, row in enumerate(rows):   # a wrong or stale real batch moves the loss by O(1), far outside this bound
            assert abs(float(row["mean_d_loss"]) - ref[e]["mean_d_loss"][n]) <= FREE_TOL["loss"] * abs(ref[e]["mean_d_loss"][n])
