"""Batch-sharded generator for generators WITHOUT BatchNorm -- the protocol half.

With the generator on process 0 only (engine.MDGANEngine, the reference's placement: server.py:219-223,266-312) the
other processes wait while G runs forward, backward and Adam.  Here every process holds a replica of the generator and
works on kb / P rows of the k*b noise batch (P processes):

    generate          z [k*b, z_dim] is drawn on process 0 exactly as before (host RNG order of server.py:219) and
                      broadcast; process p runs G forward on rows [p*m, (p+1)*m), m = k*b / P; the row blocks are
                      all-gathered into every process' X (replaces the broadcast of X, C4).
    train_workers     unchanged (routing, D steps, feedback into the slots of S), but the slot sums are ALL-reduced: every
                      process needs its rows of the group-summed feedback (replaces the reduce to process 0, C3).
    update_generator  process p runs G backward on S[p*m:(p+1)*m] with the reference's 1/(b*N) scale; the flat gradient
                      buffers are all-reduced (sum over row blocks = the gradient of the whole batch, by linearity of the
                      vector-Jacobian product in the rows); every process applies the same Adam step, so the replicas stay
                      identical without ever broadcasting weights.

Scope, stated plainly: a generator layer with train-mode BatchNorm couples the rows of the batch (statistics over all
k*b samples, server.py:219-220), so the DCGAN generators need cross-process statistics first -- the existing kernels allow
it without change (all-gather the per-CTA partial slices of `bn_partial` and run `bn_finalize` / `bn_bwd_finalize` over
`phases * P` phase blocks), but that variant is NOT built, and this class refuses such nets.  The MLP family
(mlp_nets.MlpGenNet, the reference's MNIST plugin) has no BatchNorm and runs as is.  Validated on two gloo processes
against the oracle (tests/test_mlp_host.py::test_sharded_generator_two_processes); it has not been run or timed on GPUs,
is not reachable from bootstrap.py, and the default engine does not depend on this module.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import routing
from .engine import EngineConfig, MDGANEngine
from .exchange import Exchange


class _AllReduceExchange(Exchange):
    """C3 as an all-reduce (every process consumes its rows of S); C4 is replaced by the all-gather in `generate`."""

    def broadcast_fakes(self, X: torch.Tensor) -> None:  # the row blocks are all-gathered by the engine
        return

    def reduce_feedback(self, S: torch.Tensor) -> None:
        if self.n_procs > 1:
            dist.all_reduce(S, op=dist.ReduceOp.SUM)


class ShardedGeneratorEngine(MDGANEngine):
    def __init__(self, cfg: EngineConfig, proc: int, n_procs: int, device: torch.device, generator: nn.Module,
                 discriminators: Dict[int, nn.Module], real_sources: Dict[int, Callable[[], torch.Tensor]], factory=None):
        """generator: on EVERY process, built like the server actor builds it (same seed => same initial weights; the
        flat state of process 0 is broadcast once anyway).  Only process 0 draws noise and swap pairs."""
        if generator is None:
            raise ValueError("the sharded generator needs a generator replica on every process")
        k = routing.num_generated_batches(cfg.n_workers)
        if (k * cfg.batch_size) % n_procs != 0:
            raise ValueError(f"k*b = {k * cfg.batch_size} rows do not split over {n_procs} processes")
        if any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in generator.modules()):
            raise NotImplementedError("generator with BatchNorm: its statistics span all k*b rows (server.py:219-220); the "
                                      "cross-process statistics variant is not built (mdgan_b200/sharded.py docstring)")
        super().__init__(cfg, proc, n_procs, device, generator if proc == 0 else None, discriminators, real_sources,
                         factory=factory, exchange=_AllReduceExchange(proc, n_procs, cfg.n_workers))
        self.rows = k * cfg.batch_size // n_procs
        self.r0 = proc * self.rows
        self.gen_module = generator
        # the base class built a full-batch net on process 0; every process runs a [rows]-sample replica instead
        self.gen = self.factory.generator(generator, cfg, self.rows)
        if n_procs > 1:
            dist.broadcast(self.gen.state.state_f32, src=0)
            self.gen.repack()
        self.X = torch.zeros((k * cfg.batch_size, *cfg.image_shape), device=device, dtype=torch.float32)

    def generate(self, staged: bool = False) -> None:
        if not staged:
            self.stage_inputs()
            self.upload_inputs()
        if self.proc == 0 and self.cfg.z_source != "host":
            self.z.normal_()
        if self.n_procs > 1:
            dist.broadcast(self.z, src=0)
        Xr = self.gen.forward(self.z[self.r0: self.r0 + self.rows])
        if self.n_procs > 1:
            dist.all_gather(list(self.X.chunk(self.n_procs)), Xr.contiguous())
        else:
            self.X.copy_(Xr)

    def update_generator(self) -> None:
        self.gen.backward(self.S[self.r0: self.r0 + self.rows], 1.0 / (self.b * self.N))
        if self.n_procs > 1:
            dist.all_reduce(self.gen.state.grad, op=dist.ReduceOp.SUM)
        self.gen.adam()

    def sync_modules(self) -> None:
        self.gen.state.store_to(self.gen_module)   # the replicas are identical: every process can write its own copy
        for n in self.local:
            self.disc[n].state.store_to(self.disc_modules[n])
