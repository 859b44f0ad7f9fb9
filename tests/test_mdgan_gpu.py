"""Whole-loop parity on one GPU: N logical workers + server in one process (SURVEY.md build step 6) against the
oracle, over several iterations including discriminator swaps.  Criteria and tolerances: tests/parity.py
(along-trajectory parity with TOL, free-running drift with FREE_TOL)."""
import pytest
import torch

from parity import run_engine_vs_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n_workers,b,epochs,swap", [
    ("CIFAR10", 1, 16, 3, 10**6),        # K = 1 baseline (no swap)
    ("CIFAR10", 2, 8, 4, 2),             # swap at epoch 2
    ("CIFAR10", 4, 8, 4, 1),             # BASELINE config 3 shape: N = 4, swap every epoch
    ("MNIST_DCGAN", 2, 16, 3, 10**6),    # BASELINE config 2 shape
    ("CelebA", 2, 8, 3, 1),
    ("CelebA", 8, 8, 2, 1),              # BASELINE config 4 shape: N = 8 (k = 2)
    ("CIFAR10", 2, 10, 3, 2),            # the reference's default batch size (shared-args.sh: batch_size=10), ragged tiles
    # the shipped / benchmarked shapes at the benchmarked batch size (BASELINE.json configs 2-4)
    ("MNIST_DCGAN", 1, 64, 2, 10**6),
    ("MNIST_DCGAN", 2, 64, 2, 10**6),
    ("CIFAR10", 4, 64, 3, 1),            # config 3: K = 4, swap every epoch
    ("CelebA", 2, 64, 2, 1),
    ("CelebA", 8, 64, 2, 1),             # config 4: K = 8
])
def test_engine_matches_oracle(name, n_workers, b, epochs, swap):
    r = run_engine_vs_oracle(name, n_workers, b, epochs, swap, mode="trajectory")
    assert r["pairs_bit_exact"], "swap permutation must be bit-exact with the reference's RNG stream"
    assert r["num_batches_tracked_exact"]
    assert r["ok"], r


@pytest.mark.parametrize("name,n_workers,b,epochs,swap", [
    ("CIFAR10", 2, 8, 4, 2),
    ("CelebA", 2, 4, 3, 1),
    ("MNIST_DCGAN", 4, 8, 3, 1),
])
def test_engine_free_running_drift(name, n_workers, b, epochs, swap):
    """The engine carries its own weights, Adam moments and BatchNorm buffers across iterations (no re-sync)."""
    r = run_engine_vs_oracle(name, n_workers, b, epochs, swap, mode="free")
    assert r["pairs_bit_exact"] and r["num_batches_tracked_exact"]
    assert r["ok"], r


@pytest.mark.parametrize("name,n_workers,b,epochs,swap", [
    ("CIFAR10", 2, 8, 3, 2),
    ("MNIST_DCGAN", 1, 64, 2, 10**6),    # the MNIST-shape benchmark workload
    ("CIFAR10", 4, 64, 2, 1),
    ("CelebA", 2, 64, 2, 1),
])
def test_engine_unpatched_iteration(name, n_workers, b, epochs, swap):
    """Every iteration starts from the reference's state and then runs COMPLETELY un-patched in the engine -- its own
    post-Adam discriminator weights feed the feedback pass, its own feedback feeds the generator backward.  Generated
    batch, losses and running statistics are held to the tight bounds, the quantities behind the sign-like first Adam
    steps to parity.UNPATCHED_TOL."""
    r = run_engine_vs_oracle(name, n_workers, b, epochs, swap, mode="unpatched")
    assert r["pairs_bit_exact"] and r["num_batches_tracked_exact"]
    assert r["ok"], r


def test_engine_local_epochs_two():
    r = run_engine_vs_oracle("CIFAR10", 2, 8, 2, 10**6, local_epochs=2)
    assert r["ok"], r


def test_engine_refuses_cpu_and_mlp():
    from mdgan_b200.engine import CudaNetFactory
    from mdgan_b200.plan import UnsupportedModelError, extract_plan
    from util import plugin

    with pytest.raises(RuntimeError):
        CudaNetFactory(torch.device("cpu"))
    with pytest.raises(UnsupportedModelError):
        extract_plan(plugin("MNIST").Discriminator(), "discriminator", (1, 28, 28))
