"""One GPU process of the MD-GAN training loop: the generator (process 0 only) plus the discriminator workers this
process hosts.  Mirrors, iteration for iteration, what the reference runs as N+1 OS processes:

    server loop  /root/reference/src/actors/server.py:213-333
    worker loop  /root/reference/src/actors/worker.py:157-284

Order of operations inside `iteration(epoch)` (reference order kept; the G/D synchrony of the reference is strict):
  1. process 0: z ~ N(0,1) [k*b, z_dim] (host torch RNG in parity mode, server.py:219), X = G(z)
  2. broadcast X                                    (C4, exchange.broadcast_fakes)
  3. every hosted worker n: `local_epochs` x D.train_step(real_n, X[(n+1)%k]); feedback on X[n%k] accumulated
     into slot n%k of S [k*b, C, H, W]
  4. reduce S to process 0                          (C3, exchange.reduce_feedback)
  5. process 0: G.backward(S, 1/(b*N)) -- ONE backward on the group-summed feedback instead of the reference's N
     retain_graph VJPs (server.py:271-297; equal by linearity) -- then Adam (server.py:308-312)
  6. if due: process 0 draws the swap pairs on its host RNG (server.py:321), broadcast, pairwise state exchange
     (worker.py:252-282); Adam moments stay with the rank.

The compute objects come from a factory so the host-side protocol (routing, reduce slots, swap, RNG order) can be
exercised on gloo/CPU tensors by the tests with stand-in nets; the default factory builds the CUDA nets and there is
no CPU product path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import routing
from .exchange import Exchange, make_exchange


@dataclass
class EngineConfig:
    n_workers: int
    batch_size: int
    z_dim: int
    image_shape: Tuple[int, int, int]
    generator_lr: float = 2e-4
    discriminator_lr: float = 2e-4
    beta_1: float = 0.5
    beta_2: float = 0.999
    swap_interval: int = 1
    local_epochs: int = 1
    z_source: str = "host"      # "host": torch CPU randn in the reference's RNG order (parity); "device": CUDA Philox
    prefetch_host: bool = False  # stage the NEXT iteration's host inputs (noise draw, loader batch) while the GPU runs
                                 # the current one; the host RNG order of the reference is kept (see prefetch_next)


class CudaNetFactory:
    """Builds the sm_100a nets; raises if CUDA or the kernel library is unavailable (no fallback)."""

    def __init__(self, device: torch.device):
        if device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the MD-GAN B200 engine needs a CUDA device (no CPU fallback); got " + str(device))
        from . import _lib

        lib = _lib.load()
        with torch.cuda.device(device):
            _lib.check(lib.mdgan_check_device(), "libmdgan_b200.so is sm_100a-only; device check")
        self.device = device

    def generator(self, module: nn.Module, cfg: EngineConfig, n_samples: int):
        from .nets import GenNet
        from .plan import is_mlp

        if is_mlp(module):   # the reference's Linear / LeakyReLU / dropout family (datasets/MNIST.py:74-120)
            from .mlp_nets import MlpGenNet

            return MlpGenNet(module, cfg.z_dim, cfg.image_shape, n_samples, self.device, cfg.generator_lr, cfg.beta_1,
                             cfg.beta_2)
        return GenNet(module, cfg.z_dim, cfg.image_shape, n_samples, self.device, cfg.generator_lr, cfg.beta_1, cfg.beta_2)

    def discriminator(self, module: nn.Module, cfg: EngineConfig):
        from .nets import DiscNet
        from .plan import is_mlp

        if is_mlp(module):
            from .mlp_nets import MlpDiscNet

            return MlpDiscNet(module, cfg.image_shape, cfg.batch_size, self.device, cfg.discriminator_lr, cfg.beta_1,
                              cfg.beta_2, local_epochs=cfg.local_epochs,
                              mask_source="host" if cfg.z_source == "host" else "device")
        return DiscNet(module, cfg.image_shape, cfg.batch_size, self.device, cfg.discriminator_lr, cfg.beta_1, cfg.beta_2)


class MDGANEngine:
    def __init__(self, cfg: EngineConfig, proc: int, n_procs: int, device: torch.device,
                 generator: Optional[nn.Module], discriminators: Dict[int, nn.Module],
                 real_sources: Dict[int, Callable[[], torch.Tensor]], factory=None, exchange: Optional[Exchange] = None):
        """discriminators / real_sources are keyed by 0-based worker index and must cover exactly the workers
        routing.workers_of_process(proc, n_procs, N) hosts; real_sources[n]() returns the next real batch on `device`."""
        self.cfg, self.proc, self.n_procs, self.device = cfg, proc, n_procs, device
        self.N = cfg.n_workers
        self.k = routing.num_generated_batches(self.N)
        self.b = cfg.batch_size
        self.local = routing.workers_of_process(proc, n_procs, self.N)
        if sorted(discriminators) != self.local or sorted(real_sources) != self.local:
            raise ValueError(f"process {proc} must host workers {self.local}, got {sorted(discriminators)}")
        if (generator is not None) != (proc == 0):
            raise ValueError("the generator lives on process 0 only")
        self.factory = factory or CudaNetFactory(device)
        self.exchange = exchange or make_exchange(proc, n_procs, self.N, device, self.k, self.b, cfg.image_shape)
        self.peer = getattr(self.exchange, "mode", "nccl") == "peer"
        kb = self.k * self.b
        self.gen = self.factory.generator(generator, cfg, kb) if proc == 0 else None
        self.disc = {n: self.factory.discriminator(discriminators[n], cfg) for n in self.local}
        self.real_sources = real_sources
        self.gen_module = generator
        self.disc_modules = discriminators
        f = dict(device=device, dtype=torch.float32)
        # peer mode: X is this process' symmetric copy (process 0 pushes into all of them) and the feedback goes to
        # process 0's slice buffer F[n] instead of the summed S
        self.X = self.exchange.X if self.peer else torch.zeros((kb, *cfg.image_shape), **f)
        self.S = torch.zeros((kb, *cfg.image_shape), **f)
        self.z = torch.zeros((kb, cfg.z_dim), **f)
        self.z_host = torch.zeros((kb, cfg.z_dim), dtype=torch.float32,
                                  pin_memory=(device.type == "cuda"))
        self.d_loss = torch.zeros((len(self.local), max(cfg.local_epochs, 1)), **f)
        self.g_loss = torch.zeros((len(self.local),), **f)
        self.last_pairs: Optional[torch.Tensor] = None
        self.graph = None
        self._uploaded = None
        self._staged = False
        # MDGAN_PREFETCH_H2D = 1 (default) | force | 0: with prefetch_host, the NEXT iteration's inputs are also copied to
        # the device while the current iteration runs (copy stream -> shadow buffers; the iteration then starts with two
        # device-to-device copies instead of waiting for PCIe).  See upload_ahead / upload_inputs.
        # "1" enables it on single-process runs, where it was validated (bit-identity test, A/B in profiles/); multi-GPU
        # jobs keep the compute-stream upload they were validated with until "force" has been exercised on a multi-GPU
        # box.  Nets with host-staged inputs of their own (the MLP family's dropout masks, MlpDiscNet.stage_host) use
        # the plain upload.
        h2d = os.environ.get("MDGAN_PREFETCH_H2D", "1")
        self._h2d_ahead = (cfg.prefetch_host and device.type == "cuda"
                           and (h2d == "force" or (h2d == "1" and n_procs == 1))
                           and not any(getattr(d, "stage_host", None) is not None for d in self.disc.values()))
        self._copy_stream = None
        self._ahead = False          # the shadow buffers hold the next iteration's inputs
        self._ahead_ready = None     # copy stream: shadow buffers written
        self._consumed = None        # compute stream: shadow buffers copied into the live input buffers
        self.z_next = torch.zeros((kb, cfg.z_dim), **f) if (self._h2d_ahead and proc == 0) else None
        self.iterations_done = 0

    # ------------------------------------------------------------------------------------------ phases
    def stage_inputs(self) -> None:
        """Host half of an iteration: everything that touches host RNG / host memory.  server.py:219 -- in parity
        mode (z_source == "host") the noise consumes process 0's global torch RNG exactly like the reference; real
        batches come from the host loaders in reference order (worker.py:162-167).  Fills the pinned staging buffers,
        after making sure the previous iteration's uploads have read them (the host may run iterations ahead of the
        GPU: nothing else in the loop synchronises).  A no-op when prefetch_next already staged this iteration."""
        if self._staged:
            self._staged = False
            return
        if self._uploaded is not None:
            self._uploaded.synchronize()
        kb = self.k * self.b
        if self.proc == 0 and self.cfg.z_source == "host":
            self.z_host.copy_(torch.randn((kb, self.cfg.z_dim, 1, 1)).view(kb, self.cfg.z_dim))
        for n in self.local:
            stage = getattr(self.real_sources[n], "stage", None)
            if stage is not None:
                stage()
            stage = getattr(self.disc[n], "stage_host", None)   # worker-side host draws (dropout masks, worker RNG)
            if stage is not None:
                stage()

    def upload_ahead(self) -> None:
        """Pinned staging -> SHADOW device buffers on the copy stream, while the compute stream still runs the current
        iteration (whose live input buffers must not change under it).  The copy stream first waits until the previous
        shadow contents were consumed; the staging buffers are free again once its copies are done (`_uploaded`)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._ahead_ready = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
            if self._uploaded is None:
                self._uploaded = torch.cuda.Event()
        cs = self._copy_stream
        cs.wait_event(self._consumed)
        with torch.cuda.stream(cs):
            if self.proc == 0 and self.cfg.z_source == "host":
                self.z_next.copy_(self.z_host, non_blocking=True)
            for n in self.local:
                self.real_sources[n].upload_ahead()
            self._uploaded.record(cs)
            self._ahead_ready.record(cs)
        self._ahead = True

    def upload_inputs(self) -> None:
        """Pinned staging -> device (async copies on the compute stream, always launched eagerly so that an event can
        mark the moment the staging buffers are free again).  When `upload_ahead` already moved this iteration's
        inputs to the device: wait for the copy stream and adopt the shadow buffers (device-to-device)."""
        if self._ahead:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ahead_ready)
            if self.proc == 0 and self.cfg.z_source == "host":
                self.z.copy_(self.z_next, non_blocking=True)
            for n in self.local:
                self.real_sources[n].adopt()
            self._consumed.record(cur)
            self._ahead = False
            return
        if self.proc == 0 and self.cfg.z_source == "host":
            self.z.copy_(self.z_host, non_blocking=True)
        for n in self.local:
            upload = getattr(self.real_sources[n], "upload", None)
            if upload is not None:
                upload()
            upload = getattr(self.disc[n], "upload_host", None)
            if upload is not None:
                upload()
        if self.device.type == "cuda":
            if self._uploaded is None:
                self._uploaded = torch.cuda.Event()
            self._uploaded.record()

    def prefetch_next(self, epoch: int, last: bool = False) -> None:
        """Host/GPU overlap: called right after `device_iteration` of iteration `epoch` was launched, stages iteration
        epoch + 1's host inputs while the GPU works.  Skipped when a swap is due at `epoch`: the reference draws the
        swap permutation (server.py:321) from process 0's global RNG BEFORE the next noise batch, and that order is
        kept; skipped on the last iteration so that no extra draw is consumed."""
        if not self.cfg.prefetch_host or last or routing.swap_due(epoch, self.cfg.swap_interval, self.N):
            return
        self._staged = False
        self.stage_inputs()
        self._staged = True
        if self._h2d_ahead and all(hasattr(self.real_sources[n], "upload_ahead") for n in self.local):
            self.upload_ahead()

    def generate(self, staged: bool = False) -> None:
        """staged=False: also runs the host staging and the uploads (one call per phase, as the actors use it)."""
        if not staged:
            self.stage_inputs()
            self.upload_inputs()
        if self.proc == 0:
            if self.cfg.z_source != "host":
                self.z.normal_()
            X = self.gen.forward(self.z)
            if self.peer:
                self.exchange.broadcast_fakes(X)  # push out of the net's buffer into every process' copy (+ flags)
                return
            if X.data_ptr() != self.X.data_ptr():
                self.X = X  # the net owns the [k*b, C, H, W] output buffer; broadcast straight out of it
        self.exchange.broadcast_fakes(self.X)

    def train_workers(self) -> None:
        k, b = self.k, self.b
        if not self.peer:
            self.S.zero_()
        for i, n in enumerate(self.local):
            ig, id_ = routing.route(n, k)
            x_g, x_d = self.X[ig * b:(ig + 1) * b], self.X[id_ * b:(id_ + 1) * b]
            real = self.real_sources[n]()
            net = self.disc[n]
            for l in range(self.cfg.local_epochs):
                self.d_loss[i, l].copy_(net.train_step(real, x_d))
            if self.peer:  # the last data-gradient kernel stores into process 0's F[n] over NVLink
                self.g_loss[i].copy_(net.feedback_step(x_g, out=self.exchange.feedback_slice(n), accumulate=False))
            else:
                slot = routing.feedback_slot(n, k)
                self.g_loss[i].copy_(net.feedback_step(x_g, out=self.S[slot * b:(slot + 1) * b], accumulate=True))
        self.exchange.reduce_feedback(self.S)

    def update_generator(self) -> None:
        if self.proc == 0:
            if self.peer:
                self.gen.backward(None, 1.0 / (self.b * self.N), slices=(self.exchange.F, self.k, self.N))
            else:
                self.gen.backward(self.S, 1.0 / (self.b * self.N))
            self.gen.adam()

    def maybe_swap(self, epoch: int) -> Optional[torch.Tensor]:
        if not routing.swap_due(epoch, self.cfg.swap_interval, self.N):
            self.last_pairs = None
            return None
        pairs = routing.draw_swap_pairs(self.N) if self.proc == 0 else None
        pairs = self.exchange.broadcast_pairs(pairs, self.device)
        states = {n: (self.disc[n].state.state_f32, self.disc[n].state.state_i64) for n in self.local}
        for n in self.exchange.swap_states(states, pairs):
            self.disc[n].repack()
        self.last_pairs = pairs
        return pairs

    def compute_iteration(self) -> None:
        """G forward, exchange, D steps + feedback, reduce, G backward, Adam: device work only, graph-capturable."""
        self.generate(staged=True)
        self.train_workers()
        self.update_generator()

    def _phases(self):
        return (lambda: self.generate(staged=True), self.train_workers, self.update_generator)

    def device_iteration(self, marks=None) -> None:
        """Device half of an iteration: uploads (eager) + the compute part (the captured graph(s) if there are any).
        marks: optional list of three CUDA events recorded on the compute stream after the generate / train-workers /
        update-generator phases (needs eager launches or `capture(split=True)`): the device-time spans of the
        reference's CSV columns (server.py:179-208)."""
        self.upload_inputs()
        if self.graph is None or isinstance(self.graph, list):
            for i, phase in enumerate(self.graph if self.graph is not None else self._phases()):
                phase.replay() if self.graph is not None else phase()
                if marks is not None:
                    marks[i].record()
        else:
            if marks is not None:
                raise RuntimeError("phase marks need eager launches or capture(split=True)")
            self.graph.replay()

    def capture(self, split: bool = False) -> None:
        """Capture the device half of the steady-state iteration into one CUDA graph (at b <= 128 the
        iteration is launch-bound) -- or, split=True, into three (generate + broadcast / D steps + feedback + reduce /
        G backward + Adam) replayed back to back so that CUDA events can mark the phase boundaries.  Call after at
        least one eager iteration (lazy kernel attributes, tensor maps and NCCL communicators must exist).  The swap
        stays outside the graph: it is host-driven and rare."""
        if self.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        torch.cuda.synchronize(self.device)
        if not split:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.compute_iteration()
            self.graph = graph
            return
        graphs = []
        for phase in self._phases():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=graphs[0].pool() if graphs else None):
                phase()
            graphs.append(g)
        self.graph = graphs

    def close(self) -> None:
        """Release the captured graph.  Must happen before the process group is destroyed: tearing down an NCCL
        communicator that a live CUDA graph still references hangs (observed on 2 x B200, NCCL 2.28.9)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        self.graph = None

    def iteration(self, epoch: int, last: bool = False) -> None:
        self.stage_inputs()
        self.device_iteration()
        self.prefetch_next(epoch, last)
        self.maybe_swap(epoch)
        self.iterations_done += 1

    # ------------------------------------------------------------------------------------------ results
    def mean_d_loss(self) -> List[float]:
        """worker.py:215 -- per hosted worker, mean over the local epochs (synchronises)."""
        if self.peer:
            self.exchange.check()
        return self.d_loss[:, : self.cfg.local_epochs].mean(dim=1).tolist()

    def swap_partner(self, n: int) -> Optional[int]:
        if self.last_pairs is None:
            return None
        return routing.partners_from_pairs(self.last_pairs)[n + 1]

    def sync_modules(self) -> None:
        """Write the device state back into the caller's nn.Modules (before torch.save / at the end)."""
        if self.gen is not None:
            self.gen.state.store_to(self.gen_module)
        for n in self.local:
            self.disc[n].state.store_to(self.disc_modules[n])
