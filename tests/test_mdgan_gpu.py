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


@pytest.mark.parametrize("graph", [False, True])
def test_early_upload_is_bit_identical(monkeypatch, graph):
    """The early upload (MDGAN_PREFETCH_H2D: next iteration's noise + real batches copied to shadow device buffers on
    a copy stream while the current iteration runs, engine.upload_ahead) must not change a single bit: same host RNG
    order, same batches, same step -- against the run that uploads on the compute stream at the start of every
    iteration.  Streamed host batches, two workers in one process, a swap (across which nothing is prefetched)."""
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import _DeviceBatches
    from parity import build_actor_modules
    from util import plugin

    mod = plugin("CIFAR10")
    N, b, epochs = 2, 16, 7
    dev = torch.device("cuda", 0)
    data = SyntheticImages(mod.SHAPE, N * 4 * b)
    final = {}
    for ahead in ("1", "0"):
        monkeypatch.setenv("MDGAN_PREFETCH_H2D", ahead)
        g, discs = build_actor_modules(mod, N, 3)
        cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=3,
                           z_source="host", prefetch_host=True)
        shards = routing.split_dataset(len(data), N, True)
        src = {n: _DeviceBatches(routing.RealBatchStream(data, shards[n], b), dev, tuple(mod.SHAPE)) for n in range(N)}
        eng = MDGANEngine(cfg, 0, 1, dev, g, discs, src)
        assert eng._h2d_ahead == (ahead == "1")
        used = 0
        for e in range(epochs):
            if graph and e == 2:
                eng.capture()
            eng.stage_inputs()
            used += int(eng._ahead)
            eng.device_iteration()
            eng.prefetch_next(e, last=(e == epochs - 1))
            eng.maybe_swap(e)
            eng.mean_d_loss()   # the per-iteration read-back of the actors (synchronises)
        expect = sum(1 for e in range(epochs - 1) if not routing.swap_due(e, 3, N))   # never across a swap draw
        assert used == (expect if ahead == "1" else 0) and expect == 5
        torch.cuda.synchronize()
        final[ahead] = [eng.gen.state.state_f32.clone(), eng.gen.state.m.clone()] + \
                       [eng.disc[n].state.state_f32.clone() for n in range(N)] + [eng.d_loss.clone(), eng.g_loss.clone()]
        eng.close()
    for a, c in zip(final["1"], final["0"]):
        assert torch.equal(a, c)
