"""The MLP path on the GPU (reference plugin datasets/MNIST.py:74-120): csrc/mlp.cu kernels against their plain-torch
statements in fp64 (tests/mlp_ref_ops.py -- the same functions the CPU host-logic test runs the engine on), then the
engine / bootstrap.py / standalone_gan.py with the real kernels against the oracle and the reference's golden run.
Dropout masks are the reference's (host draws from the worker RNG streams), so losses are comparable sample for
sample: iteration 0 to 1e-4, later iterations to 5e-3 (Adam's sign-like first steps amplify rounding)."""
import csv
from pathlib import Path

import pytest
import torch

import mlp_ref_ops as R
from parity import build_actor_modules, l2err
from util import plugin, relerr

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
GOLDEN = Path(__file__).resolve().parent / "golden"


def _r(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


@pytest.mark.parametrize("M,K,N", [(128, 100, 256), (70, 33, 65), (16, 1024, 784), (1, 7, 1), (128, 784, 1024)])
def test_linear_forward_kernel(M, K, N):
    from mdgan_b200 import ops

    x, W, b = _r(M, K, seed=1), _r(N, K, seed=2) * 0.05, _r(N, seed=3)
    mask = (torch.rand(M, N, generator=torch.Generator().manual_seed(4)) < 0.7).to(torch.uint8).to(DEV)
    for act, use_mask in ((ops.ACT_LRELU, True), (ops.ACT_LRELU, False), (ops.ACT_TANH, False), (ops.ACT_NONE, False)):
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.linear_forward(x, W, b, out, act=act, slope=0.2, mask=mask if use_mask else None, mask_scale=1.4285715)
        ref = torch.empty((M, N), device=DEV, dtype=torch.float64)
        R.linear_forward(x, W, b, ref, act=act, slope=0.2, mask=mask if use_mask else None, mask_scale=1.4285715)
        assert relerr(out, ref) < 2e-5, (act, use_mask)
        if use_mask:
            assert torch.equal(out == 0, (mask == 0) | (ref == 0).to(out.device))
    out = torch.empty((M, N), device=DEV)
    ops.linear_forward(x, W, None, out)                       # no bias
    assert relerr(out, x.double() @ W.double().t()) < 2e-5


@pytest.mark.parametrize("M,K,N", [(128, 784, 1024), (70, 33, 65), (64, 256, 1), (16, 100, 256)])
def test_linear_backward_kernels(M, K, N):
    """dgrad (with the previous layer's mask + gate, and accumulate), wgrad in the PyTorch layout, bias gradient."""
    from mdgan_b200 import ops

    dy, W, x = _r(M, N, seed=5), _r(N, K, seed=6) * 0.05, _r(M, K, seed=7)
    h_prev = _r(M, K, seed=8)
    mask = (torch.rand(M, K, generator=torch.Generator().manual_seed(9)) < 0.7).to(torch.uint8).to(DEV)
    h_prev = h_prev * mask            # dropped elements of the gated layer's output are exactly zero
    for use_mask, acc in ((True, False), (False, False), (False, True)):
        base = _r(M, K, seed=10)
        out = base.clone()
        ref = base.double().clone()
        kw = dict(gate=h_prev, gate_slope=0.2, mask=mask if use_mask else None, mask_scale=1.4285715, accumulate=acc)
        ops.linear_dgrad(dy, W, out, **kw)
        R.linear_dgrad(dy, W, ref, **kw)
        assert relerr(out, ref) < 2e-5, (use_mask, acc)
    out, ref = torch.empty((M, K), device=DEV), torch.empty((M, K), device=DEV, dtype=torch.float64)
    ops.linear_dgrad(dy, W, out)                              # plain: the feedback dLoss/dx
    R.linear_dgrad(dy, W, ref)
    assert relerr(out, ref) < 2e-5
    dW, dWr = torch.empty((N, K), device=DEV), torch.empty((N, K), device=DEV, dtype=torch.float64)
    ops.linear_wgrad(dy, x, dW)
    R.linear_wgrad(dy, x, dWr)
    assert relerr(dW, dWr) < 2e-5
    db = torch.empty(N, device=DEV)
    ops.col_sum(dy, db)
    assert relerr(db, dy.double().sum(0)) < 2e-5


@pytest.mark.parametrize("G,b,L", [(2, 64, 256), (1, 8, 256), (2, 10, 37)])
def test_linear_head_kernels(G, b, L):
    """Linear(L,1) + sigmoid + BCELoss(mean) and its backward against autograd in fp64 (through the plain-torch
    statement, and directly)."""
    from mdgan_b200 import ops

    n = G * b
    a, w, bias = _r(n, L, seed=11), _r(L, seed=12) * 0.2, _r(1, seed=13)
    mask = (torch.rand(n, L, generator=torch.Generator().manual_seed(14)) < 0.7).to(torch.uint8).to(DEV)
    a = a * mask
    label = torch.tensor([1.0, 0.0][:G], device=DEV)
    f = dict(device=DEV, dtype=torch.float32)
    prob, terms, dlogit, loss = torch.zeros(n, **f), torch.zeros(n, **f), torch.zeros(n, **f), torch.zeros(G + 1, **f)
    counter = torch.zeros(1, device=DEV, dtype=torch.int32)
    for _ in range(2):   # twice: the block counter must be left at zero
        ops.linear_head_forward(a, w, bias, label, prob, terms, dlogit, loss, counter, G, b)
    a64 = a.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    p = torch.sigmoid(a64 @ w64 + b64)
    per = [torch.nn.functional.binary_cross_entropy(p[g * b:(g + 1) * b], torch.full((b,), float(label[g]), device=DEV, dtype=torch.float64))
           for g in range(G)]
    total = sum(per)
    total.backward()
    assert relerr(prob, p) < 1e-5 and relerr(loss[:G], torch.stack(per)) < 1e-5 and abs(loss[G].item() - total.item()) < 1e-5 * abs(total.item())
    assert int(counter.item()) == 0
    da, dw, dbias = torch.empty((n, L), **f), torch.empty(L, **f), torch.empty(1, **f)
    ops.linear_head_backward(a, w, dlogit, da, dw, dbias, mask=mask, mask_scale=1.4285715, gate_slope=0.2)
    assert relerr(dw, w64.grad) < 2e-5 and relerr(dbias, b64.grad) < 2e-5
    gated = torch.where(mask.bool(), a64.grad * 1.4285715, torch.zeros_like(a64.grad))
    gated = torch.where(a64 > 0, gated, gated * 0.2)
    assert relerr(da, gated) < 2e-5
    ops.linear_head_backward(a, w, dlogit, da, None, None)     # feedback pass: no parameter gradients, no mask
    assert relerr(da, torch.where(a64 > 0, a64.grad, a64.grad)) < 2e-5


def _run_engine(N, b, epochs, swap, local_epochs, graph, data=None, seed=3):
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import _DeviceBatches

    mod = plugin("MNIST")
    data = data if data is not None else SyntheticImages(mod.SHAPE, N * 4 * b)
    g, discs = build_actor_modules(mod, N, seed)
    cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=swap,
                       local_epochs=local_epochs, z_source="host", prefetch_host=True)
    shards = routing.split_dataset(len(data), N, True)
    src = {n: _DeviceBatches(routing.RealBatchStream(data, shards[n], b), DEV, tuple(mod.SHAPE)) for n in range(N)}
    eng = MDGANEngine(cfg, 0, 1, DEV, g, discs, src)
    assert type(eng.gen).__name__ == "MlpGenNet" and not eng._h2d_ahead
    return mod, data, eng


@pytest.mark.parametrize("N,b,epochs,swap,local_epochs,graph", [
    (2, 8, 4, 2, 1, False),
    (2, 64, 4, 2, 1, True),      # BASELINE config 2's worker count and batch, captured graph
    (4, 16, 3, 1, 2, True),      # two local epochs, swap every iteration
    (1, 64, 3, 10 ** 6, 1, True),
])
def test_mlp_engine_matches_oracle_on_gpu(N, b, epochs, swap, local_epochs, graph):
    from oracle.mdgan_oracle import OracleMDGAN

    torch.set_num_threads(4)
    mod, data, eng = _run_engine(N, b, epochs, swap, local_epochs, graph)
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, data, N, b, mod.Z_DIM, mod.SHAPE, seed=3, beta_1=0.5,
                         swap_interval=swap, local_epochs=local_epochs)
    for e in range(epochs):
        if graph and e == 2:
            eng.capture()
        eng.stage_inputs()
        eng.device_iteration()
        eng.prefetch_next(e, last=(e == epochs - 1))
        pairs = eng.maybe_swap(e)
        ref = oracle.step(e, record=True)
        tol = 1e-4 if e == 0 else 5e-3
        assert (pairs is None) == (ref["pairs"] is None) and (pairs is None or torch.equal(pairs.cpu(), ref["pairs"]))
        assert l2err(eng.X, ref["X"]) < tol, (e, l2err(eng.X, ref["X"]))
        losses = eng.mean_d_loss()
        for n in range(N):
            assert abs(losses[n] - ref["mean_d_loss"][n]) <= tol * abs(ref["mean_d_loss"][n]), (e, n)
            assert abs(float(eng.g_loss[n]) - float(ref["loss_gen"][n])) <= tol * abs(float(ref["loss_gen"][n])), (e, n)
    eng.sync_modules()
    eng.close()
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    assert l2err(flat(eng.gen_module.state_dict()), flat(oracle.G.state_dict())) < 5e-3
    for n in range(N):
        assert l2err(flat(eng.disc_modules[n].state_dict()), flat(oracle.D[n].state_dict())) < 5e-3


def test_mlp_device_mask_source():
    """z_source = "device" (the throughput mode): noise AND dropout masks come from the CUDA generator inside the captured
    step -- no host draw, no upload.  Not comparable sample for sample; checked for the keep rate, fresh masks on every
    replay, and a finite, moving loss."""
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import DeviceResidentBatches

    mod = plugin("MNIST")
    N, b = 2, 64
    data = SyntheticImages(mod.SHAPE, N * 4 * b)
    g, discs = build_actor_modules(mod, N, 3)
    cfg = EngineConfig(n_workers=N, batch_size=b, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE), swap_interval=10 ** 9,
                       z_source="device")
    shards = routing.split_dataset(len(data), N, True)
    src = {n: DeviceResidentBatches(routing.RealBatchStream(data, shards[n], b), DEV, tuple(mod.SHAPE), 4) for n in range(N)}
    eng = MDGANEngine(cfg, 0, 1, DEV, g, discs, src)
    net = eng.disc[0]
    assert net.mask_source == "device" and net.stage_host is None
    seen, losses = [], []
    for e in range(6):
        if e == 2:
            eng.capture()
        eng.iteration(e)
        seen.append(net.mask_train[0][0].clone())
        losses.append(eng.mean_d_loss()[0])
    eng.close()
    for m in seen:
        assert abs(m.float().mean().item() - 0.7) < 0.01 and int(m.max()) == 1
    assert all(not torch.equal(seen[i], seen[i + 1]) for i in range(5)), "every replay must draw fresh masks"
    assert all(l == l and 0.1 < l < 10 for l in losses) and len(set(losses)) == 6


def test_mlp_engine_matches_the_references_own_run_on_gpu():
    """tests/golden/mnist_n2.pt: the UNMODIFIED reference's MNIST run (per-iteration mean_d_loss, final generator)."""
    from datasets.DataPartitioner import SyntheticImages

    fx = torch.load(GOLDEN / "mnist_n2.pt", weights_only=False)
    c = fx["case"]
    mod = plugin("MNIST")
    data = SyntheticImages(mod.SHAPE, c["samples"])
    _, _, eng = _run_engine(c["workers"], c["batch"], c["epochs"], c["swap_interval"], 1, False, data=data, seed=c["seed"])
    for e in range(c["epochs"]):
        eng.iteration(e, last=(e == c["epochs"] - 1))
        tol = 1e-4 if e == 0 else 5e-3
        losses = eng.mean_d_loss()
        for n in range(c["workers"]):
            ref = fx["mean_d_loss"][n][e]
            assert abs(losses[n] - ref) <= tol * abs(ref), (e, n, losses[n], ref)
    eng.sync_modules()
    eng.close()
    sd = eng.gen_module.state_dict()
    assert list(sd.keys()) == list(fx["G"].keys())
    for k, v in sd.items():
        f = fx["G"][k]
        sample = v.detach().reshape(-1)[::211].cpu()
        assert (sample - f["sample"]).abs().max().item() <= 2.5 * 2e-4 * c["epochs"], k   # <= a few Adam steps of lr


def test_bootstrap_and_standalone_on_the_mlp_plugin(tmp_path, monkeypatch):
    """The reference's entry points with --dataset MNIST (its MLP models) on the GPU against the oracle."""
    import bootstrap
    import standalone_gan
    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleMDGAN, OracleStandalone

    N, b, epochs, m = 2, 8, 4, 2 * 4 * 8
    mod = plugin("MNIST")
    monkeypatch.chdir(tmp_path)
    monkeypatch.delenv("MDGAN_SYNTH_M", raising=False)   # bootstrap --synthetic sets it: restored at teardown
    bootstrap.main(["--backend", "nccl", "--world_size", str(N + 1), "--ranks", f"0..{N}", "--dataset", "MNIST",
                    "--epochs", str(epochs), "--local_epochs", "1", "--swap_interval", "2", "--device", "cuda",
                    "--batch_size", str(b), "--iid", "1", "--seed", "3", "--beta_1", "0.5", "--generator_lr", "0.0002",
                    "--discriminator_lr", "0.0002", "--log_interval", "1000", "--gpus", "1", "--synthetic", str(m),
                    "--master_addr", "127.0.0.1", "--master_port", "29535"])
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, m), N, b, mod.Z_DIM, mod.SHAPE,
                         seed=3, beta_1=0.5, swap_interval=2)
    ref = [oracle.step(e, record=False) for e in range(epochs)]
    flat = lambda sd: torch.cat([v.reshape(-1).double() for v in sd.values()])
    g = torch.load(tmp_path / "weights" / "generator_final.pt")
    assert list(g.keys()) == list(oracle.G.state_dict().keys())
    assert l2err(flat(g), flat(oracle.G.state_dict())) < 5e-3
    for n in range(N):
        d = torch.load(tmp_path / "weights" / f"worker_{n + 1}" / "discriminator.pth")
        assert l2err(flat(d), flat(oracle.D[n].state_dict())) < 5e-3
        rows = list(csv.DictReader(open(tmp_path / "logs" / f"mdgan.{N}.MNIST.worker.{n + 1}.logs.csv")))
        for e, row in enumerate(rows):
            assert abs(float(row["mean_d_loss"]) - ref[e]["mean_d_loss"][n]) <= 5e-3 * abs(ref[e]["mean_d_loss"][n]), (e, n)
    # BASELINE config 1's entry point on its own model
    monkeypatch.setenv("MDGAN_SYNTH_M", "24")
    steps = 4
    standalone_gan.main(["--dataset", "MNIST", "--epochs", str(steps), "--local_epochs", "1", "--batch_size", "8",
                         "--device", "cuda", "--seed", "1", "--beta_1", "0.5", "--log_interval", "1000"])
    so = OracleStandalone(mod.Generator, mod.Discriminator, SyntheticImages(mod.SHAPE, 24), 8, mod.Z_DIM, seed=1, beta_1=0.5)
    rows = list(csv.DictReader(open(tmp_path / "logs" / "MNIST.standalone.logs.csv")))
    assert len(rows) == steps
    for row in rows:
        r = so.step()
        assert abs(float(row["mean_d_loss"]) - r["mean_d_loss"]) <= 5e-3 * abs(r["mean_d_loss"])
        assert abs(float(row["mean_g_loss"]) - r["mean_g_loss"]) <= 5e-3 * abs(r["mean_g_loss"])
