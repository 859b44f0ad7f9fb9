"""CPU oracle for the MD-GAN data-parallel training step.  TEST INFRASTRUCTURE ONLY.

This file restates, in one process and on the CPU (torch fp32, stock ATen ops),
the algorithm that the reference runs as N+1 gloo processes:

    server  : /root/reference/src/actors/server.py:213-333
    worker  : /root/reference/src/actors/worker.py:157-284
    init    : /root/reference/src/bootstrap.py:17-27,126-147
    baseline: /root/reference/src/standalone_gan.py:180-227

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it -- as the checker or the timed CPU
baseline, never as a product path.  The product (distributed-gan_b200/) never
imports anything from `oracle/` and fails loudly without its CUDA library.

Arithmetic note: the reference has no arithmetic of its own; every number it
produces comes from torch (pinned `torch==2.2.2`, /root/reference/requirements.txt:91;
this image: torch 2.11).  The oracle therefore calls the same torch ops in the
same order.  PARITY PIN: `oracle/ref_harness/pin_oracle.py` runs the unmodified
reference (N+1 processes, gloo, CPU) in the build container and checks that this
restatement reproduces its final generator/discriminator `state_dict`s, its
per-iteration `mean_d_loss` and its `swap_with` log; the pinned outputs are the
fixtures in `tests/golden/` (the reference itself ships no tests or golden
vectors: SURVEY.md section 4).

Every actor of the reference is an OS process with its own global torch RNG
seeded `--seed + rank` (bootstrap.py:138-141).  Here the actors share a process,
so each one's global-RNG state is saved/restored around its actions
(`_as_rank`).
"""
from __future__ import annotations

import contextlib
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.utils.data


# --------------------------------------------------------------------------- host logic
def weights_init(m: nn.Module) -> None:
    """bootstrap.py:17-27 / standalone_gan.py:19-29 -- DCGAN init by class name."""
    name = type(m).__name__
    if "Conv" in name:
        m.weight.data.normal_(0.0, 0.02)
    elif "BatchNorm" in name:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def num_generated_batches(n_workers: int) -> int:
    """server.py:120 -- k = max(floor(ln N), 2)."""
    return max(math.floor(math.log(n_workers)), 2)


def route(n: int, k: int) -> Tuple[int, int]:
    """server.py:238-239 -- worker n (0-based) gets X_g = K[n % k], X_d = K[(n+1) % k]."""
    return n % k, (n + 1) % k


def split_dataset(dataset_size: int, n_workers: int, iid: bool) -> Tuple[torch.Tensor, ...]:
    """server.py:46-64 with the generator of server.py:151-153 (private, seed 0)."""
    g = torch.Generator()
    g.manual_seed(0)
    if iid:
        idx = torch.randperm(dataset_size, generator=g)
    else:
        idx = torch.arange(dataset_size)
    return torch.chunk(idx, n_workers)


def draw_swap_pairs(n_workers: int) -> torch.Tensor:
    """server.py:321-324 -- drawn from the *server's global* RNG; returns ranks (1-based)."""
    return torch.randperm(n_workers, dtype=torch.int).view(-1, 2) + 1


def swap_due(epoch: int, swap_interval: int, n_workers: int) -> bool:
    """server.py:315-317, worker.py:239-240."""
    return n_workers > 1 and epoch % swap_interval == 0 and epoch > 0


class _RealBatches:
    """worker.py:78-89,162-167 -- DataLoader(Subset, b, shuffle, generator seed 0), restart on exhaustion."""

    def __init__(self, dataset, indices: torch.Tensor, batch_size: int):
        g = torch.Generator()
        g.manual_seed(0)
        self.loader = torch.utils.data.DataLoader(
            torch.utils.data.Subset(dataset, indices), batch_size=batch_size, shuffle=True, generator=g
        )
        self.it = iter(self.loader)

    def next(self) -> torch.Tensor:
        try:
            return next(self.it)[0]
        except StopIteration:
            self.it = iter(self.loader)
            return next(self.it)[0]


# --------------------------------------------------------------------------- actors
def d_train_step(D: nn.Module, opt: torch.optim.Optimizer, real: torch.Tensor, x_d: torch.Tensor) -> torch.Tensor:
    """worker.py:197-206 -- one local epoch of discriminator training; returns d_loss."""
    b = real.shape[0]
    crit = nn.BCELoss()
    ones, zeros = torch.ones(b, dtype=real.dtype), torch.zeros(b, dtype=real.dtype)
    D.zero_grad()
    d_loss_real = crit(D(real), ones)
    d_loss_fake = crit(D(x_d), zeros)
    d_loss = d_loss_real + d_loss_fake
    d_loss.backward()
    opt.step()
    return d_loss.detach()


def d_feedback(D: nn.Module, x_g: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """worker.py:220-233 -- F_n = d BCE(D(X_g), 1) / d X_g (D in train mode); returns (loss_gen, F_n)."""
    x = x_g.detach().clone().requires_grad_(True)
    loss_gen = nn.BCELoss()(D(x), torch.ones(x.shape[0], dtype=x.dtype))
    loss_gen.backward()
    return loss_gen.detach(), x.grad.detach()


def g_aggregate(G: nn.Module, K: Sequence[torch.Tensor], feedbacks: torch.Tensor, batch_size: int) -> List[torch.Tensor]:
    """server.py:266-302 -- N retain_graph VJPs through G, summed, times 1/(b*N)."""
    n_workers = feedbacks.shape[0]
    k = len(K)
    params = list(G.parameters())
    grads_sum = [torch.zeros_like(p) for p in params]
    for n in range(n_workers):
        grads = torch.autograd.grad(
            outputs=K[n % k], inputs=params, grad_outputs=feedbacks[n], retain_graph=True, allow_unused=True
        )
        for j, g in enumerate(grads):
            if g is not None:
                grads_sum[j] += g
    inv = 1.0 / (batch_size * n_workers)
    return [g * inv for g in grads_sum]


class OracleMDGAN:
    """N logical workers + 1 server, single process, reference order of operations."""

    def __init__(
        self,
        generator_cls: Callable[[], nn.Module],
        discriminator_cls: Callable[[], nn.Module],
        dataset,
        n_workers: int,
        batch_size: int,
        z_dim: int,
        image_shape: Tuple[int, int, int],
        seed: int = 3,
        generator_lr: float = 2e-4,
        discriminator_lr: float = 2e-4,
        beta_1: float = 0.5,
        beta_2: float = 0.999,
        swap_interval: int = 10**9,
        local_epochs: int = 1,
        iid: bool = True,
        dtype: torch.dtype = torch.float32,
    ):
        # dtype=float64 builds the "exact arithmetic" twin used by the tests to calibrate how far two correct
        # implementations of this (chaotic: ReLU gates + Adam's sign-like first steps) loop drift apart from
        # rounding alone; models are initialised in fp32 from the same RNG stream, then widened.
        self.dtype = dtype
        self.N, self.b, self.z_dim, self.shape = n_workers, batch_size, z_dim, tuple(image_shape)
        self.k = num_generated_batches(n_workers)
        self.swap_interval, self.local_epochs = swap_interval, local_epochs
        self._rng: Dict[int, torch.Tensor] = {}
        # bootstrap.py:138-145 then :102-103 (server) / :75-76 (workers)
        with self._as_rank(0, seed=seed + 0):
            self.G = generator_cls().to(dtype=torch.float32)
            self.G.apply(weights_init)
            self.G.train()  # server.py:171
        self.G.to(dtype)
        self.opt_g = torch.optim.Adam(self.G.parameters(), lr=generator_lr, betas=(beta_1, beta_2))
        self.D: List[nn.Module] = []
        self.opt_d: List[torch.optim.Optimizer] = []
        for r in range(1, n_workers + 1):
            with self._as_rank(r, seed=seed + r):
                d = discriminator_cls().to(dtype=torch.float32)
                d.apply(weights_init)
                d.train()  # worker.py:193
            self.D.append(d.to(dtype))
            self.opt_d.append(torch.optim.Adam(d.parameters(), lr=discriminator_lr, betas=(beta_1, beta_2)))
        self.shards = split_dataset(len(dataset), n_workers, iid)
        self.real = [_RealBatches(dataset, self.shards[n], batch_size) for n in range(n_workers)]

    @contextlib.contextmanager
    def _as_rank(self, rank: int, seed: Optional[int] = None):
        outer = torch.get_rng_state()
        if seed is not None:
            np.random.seed(seed)
            torch.manual_seed(seed)
        else:
            torch.set_rng_state(self._rng[rank])
        try:
            yield
        finally:
            self._rng[rank] = torch.get_rng_state()
            torch.set_rng_state(outer)

    # ---- one generator iteration (a reference "epoch")
    def step(self, epoch: int, record: bool = True, z: Optional[torch.Tensor] = None,
             replay_reals: Optional[Sequence[torch.Tensor]] = None, pairs: Optional[torch.Tensor] = None,
             record_gates: bool = False) -> Dict[str, object]:
        """z / replay_reals / pairs: replay another run's random draws instead of consuming this oracle's RNG streams and
        data loaders (used by the tests to run the fp64 twin on exactly the reference's inputs).
        record_gates: also return, per worker, the sign pattern (output > 0, NCHW bool) of every LeakyReLU of the
        feedback pass (`fb_gates`) -- the tests use it to tell a rounding-tied gate from a numerical error."""
        N, k, b = self.N, self.k, self.b
        out: Dict[str, object] = {}
        if z is None:
            with self._as_rank(0):
                z = torch.randn((k * b, self.z_dim, 1, 1))  # server.py:219
        z = z.to(self.dtype)
        X = self.G(z)  # server.py:220 (train-mode BN over all k*b samples)
        K = torch.chunk(X, k)  # server.py:223
        feedbacks = torch.zeros((N, b, *self.shape), dtype=self.dtype)
        d_losses, g_losses, reals, d_mid = [], [], [], []
        fb_gates: List[List[torch.Tensor]] = []
        for n in range(N):
            ig, id_ = route(n, k)
            x_g, x_d = K[ig].detach(), K[id_].detach()
            with self._as_rank(n + 1):
                real = (self.real[n].next() if replay_reals is None else replay_reals[n]).to(self.dtype)  # worker.py:162-167
                losses = torch.zeros(self.local_epochs)
                for l in range(self.local_epochs):  # worker.py:193-213
                    losses[l] = d_train_step(self.D[n], self.opt_d[n], real, x_d)
                if record:  # discriminator state after its Adam step(s), before the feedback pass (test hook)
                    d_mid.append({kk: v.detach().clone() for kk, v in self.D[n].state_dict().items()})
                if record_gates:
                    gates_n: List[torch.Tensor] = []
                    with _record_leaky_relu(gates_n):
                        loss_gen, F_n = d_feedback(self.D[n], x_g)  # worker.py:220-233
                    fb_gates.append(gates_n)
                else:
                    loss_gen, F_n = d_feedback(self.D[n], x_g)  # worker.py:220-233
            feedbacks[n] = F_n
            d_losses.append(losses.mean().item())
            g_losses.append(loss_gen.item())
            if record:
                reals.append(real)
        delta_w = g_aggregate(self.G, K, feedbacks, b)  # server.py:266-302
        self.opt_g.zero_grad()
        for p, g in zip(self.G.parameters(), delta_w):  # server.py:308-312
            p.grad = g.detach()
        self.opt_g.step()
        replay_pairs, pairs = pairs, None
        if swap_due(epoch, self.swap_interval, N):  # server.py:315-333, worker.py:239-284
            if replay_pairs is not None:
                pairs = replay_pairs
            else:
                with self._as_rank(0):
                    pairs = draw_swap_pairs(N)
            for a, c in pairs.tolist():
                sa = {kk: v.detach().clone() for kk, v in self.D[a - 1].state_dict().items()}
                sc = {kk: v.detach().clone() for kk, v in self.D[c - 1].state_dict().items()}
                self.D[a - 1].load_state_dict(sc)
                self.D[c - 1].load_state_dict(sa)
        out.update(mean_d_loss=d_losses, loss_gen=g_losses, pairs=pairs)
        if record:
            out.update(z=z, X=X.detach(), feedbacks=feedbacks, delta_w=[g.detach() for g in delta_w], real=reals, d_mid=d_mid)
        if record_gates:
            out["fb_gates"] = fb_gates
        return out


@contextlib.contextmanager
def _record_leaky_relu(store: List[torch.Tensor]):
    """While active, every torch.nn.functional.leaky_relu call (nn.LeakyReLU modules included) appends the sign pattern
    of its output to `store`.  Test hook only: the arithmetic is untouched."""
    import torch.nn.functional as F

    orig = F.leaky_relu

    def wrapped(input, negative_slope=0.01, inplace=False):
        out = orig(input, negative_slope, inplace)
        store.append(out.detach() > 0)
        return out

    F.leaky_relu = wrapped
    try:
        yield
    finally:
        F.leaky_relu = orig


class OracleStandalone:
    """standalone_gan.py:84-227 -- classic single-process GAN loop (local_epochs = 1)."""

    def __init__(self, generator_cls, discriminator_cls, dataset, batch_size, z_dim, seed=1,
                 generator_lr=2e-4, discriminator_lr=2e-4, beta_1=0.0, beta_2=0.999):
        np.random.seed(seed)
        torch.manual_seed(seed)  # standalone_gan.py:74-80
        self.G = generator_cls()
        self.D = discriminator_cls()
        self.G.apply(weights_init)
        self.D.apply(weights_init)
        self.b, self.z_dim = batch_size, z_dim
        self.loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=True)  # :126-131
        self.it = iter(self.loader)
        self.opt_d = torch.optim.Adam(self.D.parameters(), lr=discriminator_lr, betas=(beta_1, beta_2))
        self.opt_g = torch.optim.Adam(self.G.parameters(), lr=generator_lr, betas=(beta_1, beta_2))

    def step(self) -> Dict[str, float]:
        crit = nn.BCELoss()
        try:
            real = next(self.it)[0]
        except StopIteration:
            self.it = iter(self.loader)
            real = next(self.it)[0]
        ones, zeros = torch.ones(self.b), torch.zeros(self.b)
        fake = self.G(torch.randn(self.b, self.z_dim, 1, 1))  # :190-191
        self.D.zero_grad()  # :201-212
        err_real = crit(self.D(real), ones)
        err_real.backward()
        err_fake = crit(self.D(fake.detach()), zeros)
        err_fake.backward()
        self.opt_d.step()
        self.G.zero_grad()  # :217-222
        err_g = crit(self.D(fake), ones)
        err_g.backward()
        self.opt_g.step()
        return {"mean_d_loss": (err_real + err_fake).item(), "mean_g_loss": err_g.item()}
