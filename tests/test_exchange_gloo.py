"""The N>1 protocol of the engine (broadcast of the generated batches, feedback slot sum + reduce, swap permutation
send/recv, actor placement) on world_size-2 gloo/CPU, with deterministic stand-in nets plugged into the engine's
net factory.  The 2-process run must reproduce the 1-process run of the same job bit-for-bit (the only cross-process
arithmetic is a 2-operand sum), and the swap must move states exactly as the reference's pair table says."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = Path(__file__).resolve().parent
for p in (str(HERE.parent), str(HERE.parent / "distributed-gan_b200"), str(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

SHAPE, Z, B, N, EPOCHS = (1, 4, 4), 6, 4, 4, 5


class _State:
    def __init__(self, module):
        self.module = module
        self.state_f32 = torch.cat([p.detach().reshape(-1).clone() for p in module.parameters()])
        self.state_i64 = torch.zeros(1, dtype=torch.int64)

    def store_to(self, module):
        off = 0
        for p in module.parameters():
            p.data.copy_(self.state_f32[off: off + p.numel()].view_as(p))
            off += p.numel()


class FakeGen:
    def __init__(self, module, cfg, n):
        self.state, self.n, self.lr = _State(module), n, cfg.generator_lr
        self.grad = torch.zeros_like(self.state.state_f32)

    def forward(self, z):
        self.z = z.clone()
        w = self.state.state_f32.view(16, Z)
        self.X = torch.tanh(z @ w.t()).view(self.n, *SHAPE)
        return self.X

    def backward(self, S, scale):
        d = (S * (1 - self.X * self.X) * scale).view(self.n, 16)
        self.grad = (d.t() @ self.z).reshape(-1)

    def adam(self):
        self.state.state_f32 -= self.lr * 100 * self.grad


class FakeDisc:
    def __init__(self, module, cfg):
        self.state = _State(module)
        self.repacks = 0

    def repack(self):
        self.repacks += 1

    def train_step(self, real, x_d):
        w = self.state.state_f32
        w += 0.01 * (real.mean(0).reshape(-1) - x_d.mean(0).reshape(-1))
        self.state.state_i64 += 1
        return (real.reshape(B, -1) @ w).mean() - (x_d.reshape(B, -1) @ w).mean()

    def feedback_step(self, x_g, out, accumulate):
        f = (x_g.reshape(B, -1) * self.state.state_f32).view(B, *SHAPE) / B
        out.add_(f) if accumulate else out.copy_(f)
        return f.sum()


class FakeFactory:
    def generator(self, module, cfg, n):
        return FakeGen(module, cfg, n)

    def discriminator(self, module, cfg):
        return FakeDisc(module, cfg)


def _run(proc, n_procs, port, out_dir):
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine

    if n_procs > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=proc, world_size=n_procs)
    try:
        dataset = SyntheticImages(SHAPE, N * 4 * B)
        shards = routing.split_dataset(len(dataset), N, True)
        local = routing.workers_of_process(proc, n_procs, N)
        discs = {}
        for n in local:
            torch.manual_seed(100 + n)
            discs[n] = nn.Linear(16, 1, bias=False)
        gen = None
        if proc == 0:
            torch.manual_seed(3)
            gen = nn.Linear(Z, 16, bias=False)
        streams = {n: routing.RealBatchStream(dataset, shards[n], B) for n in local}
        sources = {n: streams[n].next for n in local}
        cfg = EngineConfig(n_workers=N, batch_size=B, z_dim=Z, image_shape=SHAPE, swap_interval=2, z_source="host")
        eng = MDGANEngine(cfg, proc, n_procs, torch.device("cpu"), gen, discs, sources, factory=FakeFactory())
        log = []
        for e in range(EPOCHS):
            eng.iteration(e)
            log.append({"pairs": None if eng.last_pairs is None else eng.last_pairs.clone(),
                        "partners": {n: eng.swap_partner(n) for n in local}})
        eng.sync_modules()
        res = {"G": None if gen is None else gen.weight.detach().clone(), "X": eng.X.clone(), "S": eng.S.clone(),
               "D": {n: (eng.disc[n].state.state_f32.clone(), eng.disc[n].state.state_i64.clone(), eng.disc[n].repacks)
                     for n in local},
               "d_loss": {n: eng.d_loss[i, 0].item() for i, n in enumerate(local)}, "log": log}
        torch.save(res, Path(out_dir) / f"res_{n_procs}_{proc}.pt")
    finally:
        if n_procs > 1:
            dist.barrier()
            dist.destroy_process_group()


def test_two_process_gloo_run_matches_single_process(tmp_path):
    _run(0, 1, 0, tmp_path)
    port = 29000 + os.getpid() % 2000
    mp.spawn(_run, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    one = torch.load(tmp_path / "res_1_0.pt", weights_only=False)
    two = [torch.load(tmp_path / f"res_2_{p}.pt", weights_only=False) for p in range(2)]
    assert torch.equal(one["G"], two[0]["G"]), "generator after 5 iterations"
    assert torch.equal(one["X"], two[0]["X"]) and torch.equal(one["X"], two[1]["X"]), "every process sees the same X"
    assert torch.equal(one["S"], two[0]["S"]), "group-summed feedback reduced to process 0"
    merged = {**two[0]["D"], **two[1]["D"]}
    assert sorted(merged) == [0, 1, 2, 3]
    for n in range(N):
        assert torch.equal(one["D"][n][0], merged[n][0]) and torch.equal(one["D"][n][1], merged[n][1]), n
        assert merged[n][2] == one["D"][n][2] == 2, "packed weights rebuilt after each of the 2 swaps (epochs 2, 4)"
    losses = {**two[0]["d_loss"], **two[1]["d_loss"]}
    assert all(losses[n] == one["d_loss"][n] for n in range(N))
    for e in range(EPOCHS):
        p1 = one["log"][e]["pairs"]
        for r in two:
            p2 = r["log"][e]["pairs"]
            assert (p1 is None) == (p2 is None) and (p1 is None or torch.equal(p1, p2)), "pair table bit-exact"
        assert (p1 is not None) == (e in (2, 4))
    # the swap really exchanged states: after epoch 4 worker a holds what its partner trained
    partners = {**two[0]["log"][4]["partners"], **two[1]["log"][4]["partners"]}
    assert all(partners[partners[n] - 1] - 1 == n for n in range(N))


def _single(prefetch: bool):
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine

    dataset = SyntheticImages(SHAPE, N * 4 * B)
    shards = routing.split_dataset(len(dataset), N, True)
    discs = {}
    for n in range(N):
        torch.manual_seed(100 + n)
        discs[n] = nn.Linear(16, 1, bias=False)
    torch.manual_seed(3)
    gen = nn.Linear(Z, 16, bias=False)
    streams = {n: routing.RealBatchStream(dataset, shards[n], B) for n in range(N)}
    cfg = EngineConfig(n_workers=N, batch_size=B, z_dim=Z, image_shape=SHAPE, swap_interval=2, z_source="host",
                       prefetch_host=prefetch)
    eng = MDGANEngine(cfg, 0, 1, torch.device("cpu"), gen, discs, {n: streams[n].next for n in range(N)},
                      factory=FakeFactory())
    pairs, zs = [], []
    for e in range(EPOCHS):
        eng.iteration(e, last=(e == EPOCHS - 1))
        pairs.append(None if eng.last_pairs is None else eng.last_pairs.clone())
        zs.append(eng.gen.z.clone())
    eng.sync_modules()
    return gen.weight.detach().clone(), eng.X.clone(), pairs, zs, torch.get_rng_state()


def test_host_prefetch_keeps_the_reference_rng_order():
    """prefetch_host stages iteration e+1's noise while iteration e runs, except across a swap draw (the reference
    draws the permutation before the next noise batch, server.py:321 vs :219) and after the last iteration: the noise
    batches, swap pairs, results and the final state of the global RNG must not change."""
    a, b = _single(False), _single(True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert all((p is None and q is None) or torch.equal(p, q) for p, q in zip(a[2], b[2]))
    assert all(torch.equal(p, q) for p, q in zip(a[3], b[3]))
    assert torch.equal(a[4], b[4]), "no extra draw from the global RNG"


def test_make_exchange_on_cpu_is_the_collective_exchange():
    from mdgan_b200.exchange import Exchange, make_exchange

    ex = make_exchange(0, 1, 2, torch.device("cpu"), 2, 4, SHAPE)
    assert type(ex) is Exchange and ex.mode == "nccl"


def test_exchange_refuses_uninitialised_group():
    from mdgan_b200.exchange import Exchange

    with pytest.raises(RuntimeError):
        Exchange(0, 2, 2)
