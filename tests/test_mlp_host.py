"""Host logic of the MLP path (datasets/MNIST.py:74-120 on the engine) WITHOUT a GPU: the kernels are replaced by their
plain-torch statements (tests/mlp_ref_ops.py), everything else is the product code -- plan extraction, mlp_nets
(layer wiring, gradient routing, the dropout masks drawn from each worker's RNG stream in the reference's order),
engine.MDGANEngine (staging, routing, feedback slots, swap).  Checked against the oracle and against the golden
fixture of the UNMODIFIED reference's own MNIST run (tests/golden/mnist_n2.pt).  The CUDA kernels themselves are
checked against the same plain-torch statements in tests/test_mlp_gpu.py."""
import os
import sys
from pathlib import Path

import pytest
import torch

HERE = Path(__file__).resolve().parent
for _p in (str(HERE.parent), str(HERE.parent / "distributed-gan_b200"), str(HERE)):   # spawned processes import this file
    if _p not in sys.path:
        sys.path.insert(0, _p)

import mlp_ref_ops  # noqa: E402
from parity import build_actor_modules, l2err
from util import plugin

GOLDEN = Path(__file__).resolve().parent / "golden"


class _CpuMlpFactory:
    def __init__(self, device):
        self.device = device

    def generator(self, module, cfg, n_samples):
        from mdgan_b200.mlp_nets import MlpGenNet

        return MlpGenNet(module, cfg.z_dim, cfg.image_shape, n_samples, self.device, cfg.generator_lr, cfg.beta_1, cfg.beta_2)

    def discriminator(self, module, cfg):
        from mdgan_b200.mlp_nets import MlpDiscNet

        return MlpDiscNet(module, cfg.image_shape, cfg.batch_size, self.device, cfg.discriminator_lr, cfg.beta_1,
                          cfg.beta_2, local_epochs=cfg.local_epochs)


class _HostBatches:
    def __init__(self, stream):
        self.stream, self.cur = stream, None

    def stage(self):
        self.cur = self.stream.next().clone()

    def __call__(self):
        return self.cur


def _engine(mod, dataset, N, b, swap, local_epochs=1, prefetch=False, seed=3):
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine

    g, discs = build_actor_modules(mod, N, seed)
    cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=swap,
                       local_epochs=local_epochs, z_source="host", prefetch_host=prefetch)
    shards = routing.split_dataset(len(dataset), N, True)
    src = {n: _HostBatches(routing.RealBatchStream(dataset, shards[n], b)) for n in range(N)}
    dev = torch.device("cpu")
    return MDGANEngine(cfg, 0, 1, dev, g, discs, src, factory=_CpuMlpFactory(dev))


@pytest.mark.parametrize("N,b,epochs,swap,local_epochs,prefetch", [
    (2, 8, 4, 2, 1, False),
    (2, 8, 4, 2, 1, True),      # masks of iteration e+1 drawn while e "runs": the worker streams must not care
    (4, 4, 3, 1, 2, False),     # two local epochs: 12 + 3 mask draws per worker and iteration
    (1, 16, 2, 10 ** 6, 1, False),
])
def test_mlp_engine_matches_oracle(monkeypatch, N, b, epochs, swap, local_epochs, prefetch):
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops
    from oracle.mdgan_oracle import OracleMDGAN

    mlp_ref_ops.patch(monkeypatch, ops)
    torch.set_num_threads(1)
    mod = plugin("MNIST")
    data = SyntheticImages(mod.SHAPE, N * 4 * b)
    eng = _engine(mod, data, N, b, swap, local_epochs, prefetch)
    assert all(type(d).__name__ == "MlpDiscNet" for d in eng.disc.values()) and type(eng.gen).__name__ == "MlpGenNet"
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, data, N, b, mod.Z_DIM, mod.SHAPE, seed=3, beta_1=0.5,
                         swap_interval=swap, local_epochs=local_epochs)
    for e in range(epochs):
        eng.stage_inputs()
        eng.device_iteration()
        eng.prefetch_next(e, last=(e == epochs - 1))
        pairs = eng.maybe_swap(e)
        ref = oracle.step(e, record=True)
        assert (pairs is None) == (ref["pairs"] is None) and (pairs is None or torch.equal(pairs, ref["pairs"]))
        assert l2err(eng.X, ref["X"]) < 1e-5, e
        for n in range(N):
            assert abs(eng.mean_d_loss()[n] - ref["mean_d_loss"][n]) <= 1e-5 * abs(ref["mean_d_loss"][n]), (e, n)
            assert abs(float(eng.g_loss[n]) - float(ref["loss_gen"][n])) <= 1e-5 * abs(float(ref["loss_gen"][n])), (e, n)
    eng.sync_modules()
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    assert l2err(flat(eng.gen_module.state_dict()), flat(oracle.G.state_dict())) < 1e-4
    for n in range(N):
        assert list(eng.disc_modules[n].state_dict().keys()) == list(oracle.D[n].state_dict().keys())
        assert l2err(flat(eng.disc_modules[n].state_dict()), flat(oracle.D[n].state_dict())) < 1e-4


@pytest.mark.parametrize("name", ["mnist_n2", "mnist_n4_swap"])
def test_mlp_engine_matches_the_references_own_run(monkeypatch, name):
    """tests/golden/mnist_n2.pt, mnist_n4_swap.pt: per-iteration mean_d_loss, the swap log and the final state_dicts of
    the UNMODIFIED reference (bootstrap.py, N+1 gloo processes, datasets/MNIST.py with its always-on dropout)."""
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops

    mlp_ref_ops.patch(monkeypatch, ops)
    torch.set_num_threads(1)
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    c = fx["case"]
    assert c["dataset"] == "MNIST" and c["mode"] == "distributed"
    mod = plugin("MNIST")
    data = SyntheticImages(mod.SHAPE, c["samples"])
    eng = _engine(mod, data, c["workers"], c["batch"], c["swap_interval"], seed=c["seed"])
    swaps = 0
    for e in range(c["epochs"]):
        eng.iteration(e, last=(e == c["epochs"] - 1))
        for n in range(c["workers"]):
            ref = fx["mean_d_loss"][n][e]
            assert abs(eng.mean_d_loss()[n] - ref) <= 1e-5 * abs(ref), (e, n)
            assert eng.swap_partner(n) == fx["swap_with"][n][e], "swap permutation must be the reference's"
            swaps += fx["swap_with"][n][e] is not None
    assert (swaps > 0) == (name == "mnist_n4_swap")
    eng.sync_modules()

    def check(sd, fxs):
        assert list(sd.keys()) == list(fxs.keys())
        for k, v in sd.items():
            f = fxs[k]
            sample = v.detach().reshape(-1)[::fx["stride"]]
            # weights after Adam: a gradient element at rounding level moves its weight by a fraction of lr whichever
            # implementation computes it (sign-like first steps); a wrong gradient moves it by ~lr = 2e-4 per step
            assert (sample - f["sample"]).abs().max().item() <= 2e-5, k

    check(eng.gen_module.state_dict(), fx["G"])
    for n in range(c["workers"]):
        check(eng.disc_modules[n].state_dict(), fx["D"][n])


def test_mlp_standalone_matches_the_references_own_run(monkeypatch):
    """tests/golden/mnist_standalone.pt: the UNMODIFIED reference's standalone_gan.py on its MLP plugin (BASELINE config
    1's entry point and model) -- both losses of every step and the final nets."""
    import standalone_gan
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops

    mlp_ref_ops.patch(monkeypatch, ops)
    torch.set_num_threads(1)
    fx = torch.load(GOLDEN / "mnist_standalone.pt", weights_only=False)
    c = fx["case"]
    mod = plugin("MNIST")
    data = SyntheticImages(mod.SHAPE, c["samples"])
    dev = torch.device("cpu")
    run = standalone_gan.Standalone(mod, data, c["batch"], dev, c["seed"], 2e-4, 2e-4, c["beta_1"], 0.999,
                                    factory=_CpuMlpFactory(dev))
    for e in range(c["epochs"]):
        d_loss, g_loss = (float(x) for x in run.step())
        assert abs(d_loss - fx["mean_d_loss"][0][e]) <= 1e-5 * abs(fx["mean_d_loss"][0][e]), e
        assert abs(g_loss - fx["mean_g_loss"][0][e]) <= 1e-5 * abs(fx["mean_g_loss"][0][e]), e
    run.gen.state.store_to(run.G)
    run.disc.state.store_to(run.D)
    for sd, fxs in ((run.G.state_dict(), fx["G"]), (run.D.state_dict(), fx["D"][0])):
        assert list(sd.keys()) == list(fxs.keys())
        for k, v in sd.items():
            sample = v.detach().reshape(-1)[::fx["stride"]]
            assert (sample - fxs[k]["sample"]).abs().max().item() <= 2e-5, k   # a tenth of lr (see above)


def test_mlp_standalone_matches_oracle(monkeypatch):
    """standalone_gan.py on the MLP plugin (BASELINE config 1's own model): ONE global RNG stream carries the loader
    seeds, the noise and the nine dropout draws of every step, in the reference's order (standalone_gan.py:180-227)."""
    import standalone_gan
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops
    from oracle.mdgan_oracle import OracleStandalone

    mlp_ref_ops.patch(monkeypatch, ops)
    torch.set_num_threads(1)
    mod = plugin("MNIST")
    b, steps = 8, 5
    data = SyntheticImages(mod.SHAPE, 3 * b)           # the loader wraps around inside the run
    dev = torch.device("cpu")
    run = standalone_gan.Standalone(mod, data, b, dev, 1, 2e-4, 2e-4, 0.5, 0.999, factory=_CpuMlpFactory(dev))
    got = [tuple(float(x) for x in run.step()) for _ in range(steps)]
    oracle = OracleStandalone(mod.Generator, mod.Discriminator, data, b, mod.Z_DIM, seed=1, beta_1=0.5)
    for (d_loss, g_loss) in got:
        ref = oracle.step()
        assert abs(d_loss - ref["mean_d_loss"]) <= 1e-5 * abs(ref["mean_d_loss"])
        assert abs(g_loss - ref["mean_g_loss"]) <= 1e-5 * abs(ref["mean_g_loss"])
    run.gen.state.store_to(run.G)
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    assert l2err(flat(run.G.state_dict()), flat(oracle.G.state_dict())) < 1e-4


def _two_process_worker(proc, n_procs, port, out_dir, N, b, epochs, swap):
    """One GPU-process stand-in of a 2-process gloo job, built the way bootstrap.init_process builds it: every hosted
    worker actor on its own seed with its RNG stream handed over on the module, the server last."""
    import torch.distributed as dist

    import bootstrap
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops, routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=proc, world_size=n_procs)
    try:
        for name in mlp_ref_ops.ALL:
            setattr(ops, name, getattr(mlp_ref_ops, name))
        torch.set_num_threads(1)
        mod = plugin("MNIST")
        data = SyntheticImages(mod.SHAPE, N * 4 * b)
        local = routing.workers_of_process(proc, n_procs, N)
        discs = {}
        for n in local:
            bootstrap._seed_actor(3 + n + 1)
            d = mod.Discriminator()
            d.apply(bootstrap._weights_init)
            d._mdgan_rng_state = torch.get_rng_state()
            discs[n] = d
        gen = None
        if proc == 0:
            bootstrap._seed_actor(3)
            gen = mod.Generator()
            gen.apply(bootstrap._weights_init)
        cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=swap,
                           z_source="host", prefetch_host=True)
        shards = routing.split_dataset(len(data), N, True)
        src = {n: _HostBatches(routing.RealBatchStream(data, shards[n], b)) for n in local}
        dev = torch.device("cpu")
        eng = MDGANEngine(cfg, proc, n_procs, dev, gen, discs, src, factory=_CpuMlpFactory(dev))
        losses = []
        for e in range(epochs):
            eng.iteration(e, last=(e == epochs - 1))
            losses.append(eng.mean_d_loss())
        eng.sync_modules()
        torch.save({"G": None if gen is None else gen.state_dict(), "D": {n: discs[n].state_dict() for n in local},
                    "losses": losses, "local": local}, Path(out_dir) / f"mlp_{proc}.pt")
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_mlp_two_process_gloo_run_matches_oracle(tmp_path):
    """N > 1 processes (world_size-2 gloo, CPU): workers 1-2 with the server on process 0, workers 3-4 on process 1, each
    worker's dropout stream continued from ITS actor seed wherever it is hosted, a swap across the process boundary."""
    import torch.multiprocessing as mp

    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleMDGAN

    N, b, epochs, swap = 4, 4, 4, 2
    port = 29100 + os.getpid() % 1500
    mp.spawn(_two_process_worker, args=(2, port, str(tmp_path), N, b, epochs, swap), nprocs=2, join=True)
    res = [torch.load(tmp_path / f"mlp_{p}.pt", weights_only=False) for p in range(2)]
    torch.set_num_threads(1)
    mod = plugin("MNIST")
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, N * 4 * b), N, b, mod.Z_DIM, mod.SHAPE,
                         seed=3, beta_1=0.5, swap_interval=swap)
    ref = [oracle.step(e, record=False) for e in range(epochs)]
    assert any(r["pairs"] is not None for r in ref)
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    assert l2err(flat(res[0]["G"]), flat(oracle.G.state_dict())) < 1e-4
    assert res[0]["local"] == [0, 1] and res[1]["local"] == [2, 3]
    for r in res:
        for i, n in enumerate(r["local"]):
            assert l2err(flat(r["D"][n]), flat(oracle.D[n].state_dict())) < 1e-4, n
            for e in range(epochs):
                assert abs(r["losses"][e][i] - ref[e]["mean_d_loss"][n]) <= 1e-5 * abs(ref[e]["mean_d_loss"][n]), (e, n)


def test_mlp_plan_extraction_and_refusals():
    """plan.extract_mlp_plan on the reference's MLP models, and the loud refusals (no fallback path)."""
    import torch.nn as nn
    import torch.nn.functional as F

    from mdgan_b200.plan import UnsupportedModelError, extract_mlp_plan, extract_plan, is_mlp

    mod = plugin("MNIST")
    torch.manual_seed(0)
    d, g = mod.Discriminator(), mod.Generator()
    before = torch.get_rng_state()
    pd = extract_mlp_plan(d, "discriminator", mod.SHAPE)
    pg = extract_mlp_plan(g, "generator", (mod.Z_DIM, 1, 1))
    assert torch.equal(before, torch.get_rng_state()), "the dry run (which draws dropout masks) must not consume the RNG"
    assert [(l.n_in, l.n_out, l.act, l.drop_p) for l in pd.layers] == [
        (784, 1024, "lrelu", 0.3), (1024, 512, "lrelu", 0.3), (512, 256, "lrelu", 0.3), (256, 1, "sigmoid", 0.0)]
    assert [(l.n_in, l.n_out, l.act, l.drop_p) for l in pg.layers] == [
        (100, 256, "lrelu", 0.0), (256, 512, "lrelu", 0.0), (512, 1024, "lrelu", 0.0), (1024, 784, "tanh", 0.0)]
    assert [l.weight for l in pd.layers] == ["fc1.weight", "fc2.weight", "fc3.weight", "fc4.weight"]
    assert all(abs(l.slope - 0.2) < 1e-9 for l in pd.layers[:-1]) and pg.out_shape == (1, 28, 28) and pd.out_shape == ()
    assert is_mlp(d) and is_mlp(g) and not is_mlp(plugin("CIFAR10").Discriminator())
    with pytest.raises(UnsupportedModelError):
        extract_plan(d, "discriminator", mod.SHAPE)              # the conv-family extractor refuses Linear layers

    class ReluMlp(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(784, 32), nn.Linear(32, 1)

        def forward(self, x):
            return torch.sigmoid(self.b(F.relu(self.a(x.view(x.shape[0], -1))))).flatten()

    class NoSigmoid(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(784, 32), nn.Linear(32, 1)

        def forward(self, x):
            return self.b(F.leaky_relu(self.a(x.view(x.shape[0], -1)), 0.2)).flatten()

    class DropoutInGenerator(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = nn.Linear(100, 32), nn.Linear(32, 784)

        def forward(self, z):
            h = F.dropout(F.leaky_relu(self.a(z.view(z.shape[0], -1)), 0.2), 0.5)
            return torch.tanh(self.b(h)).view(-1, 1, 28, 28)

    for bad, role, shape in ((ReluMlp(), "discriminator", mod.SHAPE), (NoSigmoid(), "discriminator", mod.SHAPE),
                             (DropoutInGenerator(), "generator", (100, 1, 1))):
        with pytest.raises(UnsupportedModelError):
            extract_mlp_plan(bad, role, shape)


def _sharded_worker(proc, n_procs, port, out_dir, N, b, epochs, swap):
    import torch.distributed as dist

    import bootstrap
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops, routing
    from mdgan_b200.engine import EngineConfig
    from mdgan_b200.sharded import ShardedGeneratorEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=proc, world_size=n_procs)
    try:
        for name in mlp_ref_ops.ALL:
            setattr(ops, name, getattr(mlp_ref_ops, name))
        torch.set_num_threads(1)
        mod = plugin("MNIST")
        data = SyntheticImages(mod.SHAPE, N * 4 * b)
        local = routing.workers_of_process(proc, n_procs, N)
        discs = {}
        for n in local:
            bootstrap._seed_actor(3 + n + 1)
            d = mod.Discriminator()
            d.apply(bootstrap._weights_init)
            d._mdgan_rng_state = torch.get_rng_state()
            discs[n] = d
        bootstrap._seed_actor(3)            # the server's seed on EVERY process: identical generator replicas;
        gen = mod.Generator()               # only process 0 goes on drawing from this stream (noise, swap pairs)
        gen.apply(bootstrap._weights_init)
        cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=swap,
                           z_source="host", prefetch_host=True)
        shards = routing.split_dataset(len(data), N, True)
        src = {n: _HostBatches(routing.RealBatchStream(data, shards[n], b)) for n in local}
        dev = torch.device("cpu")
        eng = ShardedGeneratorEngine(cfg, proc, n_procs, dev, gen, discs, src, factory=_CpuMlpFactory(dev))
        assert eng.gen.n == 2 * b // n_procs
        losses = []
        for e in range(epochs):
            eng.iteration(e, last=(e == epochs - 1))
            losses.append(eng.mean_d_loss())
        eng.sync_modules()
        torch.save({"G": gen.state_dict(), "D": {n: discs[n].state_dict() for n in local}, "losses": losses, "local": local,
                    "X": eng.X.clone()}, Path(out_dir) / f"sharded_{proc}.pt")
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sharded_generator_two_processes(tmp_path):
    """SURVEY n2 (protocol half, BatchNorm-free generators): every process runs G forward / backward on its k*b / P rows
    of the noise batch -- noise broadcast, row blocks all-gathered, feedback all-reduced, gradients all-reduced, the same
    Adam step everywhere -- and the job still reproduces the reference (oracle) to 1e-5, swaps included; the replicas
    stay bit-identical to each other."""
    import torch.multiprocessing as mp

    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200.sharded import ShardedGeneratorEngine
    from oracle.mdgan_oracle import OracleMDGAN

    N, b, epochs, swap = 4, 4, 4, 2
    port = 29300 + os.getpid() % 1500
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path), N, b, epochs, swap), nprocs=2, join=True)
    res = [torch.load(tmp_path / f"sharded_{p}.pt", weights_only=False) for p in range(2)]
    torch.set_num_threads(1)
    mod = plugin("MNIST")
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, N * 4 * b), N, b, mod.Z_DIM, mod.SHAPE,
                         seed=3, beta_1=0.5, swap_interval=swap)
    ref = [oracle.step(e, record=True) for e in range(epochs)]
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    assert all(torch.equal(res[0]["G"][k], res[1]["G"][k]) for k in res[0]["G"]), "generator replicas must stay identical"
    assert torch.equal(res[0]["X"], res[1]["X"]) and l2err(res[0]["X"], ref[-1]["X"]) < 1e-5
    assert l2err(flat(res[0]["G"]), flat(oracle.G.state_dict())) < 1e-4
    for r in res:
        for i, n in enumerate(r["local"]):
            assert l2err(flat(r["D"][n]), flat(oracle.D[n].state_dict())) < 1e-4, n
            for e in range(epochs):
                assert abs(r["losses"][e][i] - ref[e]["mean_d_loss"][n]) <= 1e-5 * abs(ref[e]["mean_d_loss"][n]), (e, n)
    # generators with BatchNorm couple the rows of the batch: refused until the cross-process statistics exist
    from mdgan_b200.engine import EngineConfig
    cif = plugin("CIFAR10")
    cfg = EngineConfig(n_workers=2, batch_size=4, z_dim=cif.Z_DIM, image_shape=tuple(cif.SHAPE))
    with pytest.raises(NotImplementedError):
        ShardedGeneratorEngine(cfg, 0, 1, torch.device("cpu"), cif.Generator(), {}, {}, factory=_CpuMlpFactory(torch.device("cpu")))
