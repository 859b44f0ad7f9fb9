"""Device-resident engines for the reference's second model family: the MLP generator / discriminator of
/root/reference/src/datasets/MNIST.py:74-120 (four nn.Linear per net, LeakyReLU(0.2), always-active F.dropout(0.3) in
the discriminator, tanh / sigmoid outputs).  Same interface as nets.GenNet / nets.DiscNet, so engine.MDGANEngine runs
them unchanged: `MlpDiscNet.train_step / feedback_step` = worker.py:193-233, `MlpGenNet.forward / backward / adam` =
server.py:219-223,266-312.

Kernels: csrc/mlp.cu (fp32 SGEMM with the bias / activation / dropout / gate tail in its epilogue, column sums for the
bias gradients, Linear(L,1)+sigmoid+BCE head), plus the shared tanh-backward, Adam and exchange kernels.

Dropout parity.  On the CPU F.dropout(x, p) is `noise = empty_like(x).bernoulli_(1 - p).div_(1 - p); x * noise`, drawn
from the worker's GLOBAL torch RNG (the reference seeds it with --seed + rank, bootstrap.py:138-141; the model's default
init has consumed part of the stream by then).  `MlpDiscNet` owns a torch.Generator that continues exactly that stream
and draws the keep masks on the host, in the reference's call order (per local epoch: real fc1, fc2, fc3, then X_d fc1,
fc2, fc3; then the feedback pass), into a pinned uint8 buffer that is uploaded with the iteration's other inputs.  The
masks are therefore the reference's masks bit for bit, and the captured CUDA graph reads them from fixed addresses.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .nets import FlatState, _cpu_copy
from .plan import MlpPlan, extract_mlp_plan


def _keep_scale(p: float) -> float:
    """The fp32 factor F.dropout multiplies kept elements with: ones.div_(1 - p) evaluated by torch itself."""
    return float(torch.ones(1, dtype=torch.float32).div_(1.0 - p).item())


class MlpDiscNet:
    def __init__(self, module: nn.Module, image_shape: Tuple[int, int, int], batch_size: int, device: torch.device,
                 lr: float, beta_1: float, beta_2: float, max_groups: int = 2, local_epochs: int = 1,
                 rng_state: Optional[torch.Tensor] = None, mask_source: str = "host"):
        """mask_source "host" (parity mode, see above) | "device": the keep masks come from the CUDA generator inside
        the step (no host draw, no upload; what the reference does when it runs on a GPU) -- follows EngineConfig.z_source.
        rng_state: state of the worker actor's global torch RNG right after its model was built (bootstrap stores it
        on the module as `_mdgan_rng_state`); default: this process' global RNG as it is now -- correct for a process
        that hosts exactly this one worker, like a reference worker process."""
        self.device, self.b, self.shape = device, batch_size, tuple(image_shape)
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.plan: MlpPlan = extract_mlp_plan(_cpu_copy(module), "discriminator", self.shape)
        self.state = FlatState(module, device)
        self.L = self.plan.layers
        self.hidden = self.L[:-1]
        self.local_epochs = max(int(local_epochs), 1)
        b, nmax = batch_size, max_groups * batch_size
        self.nmax, self.max_groups = nmax, max_groups
        f = dict(device=device, dtype=torch.float32)
        self.n_in = self.L[0].n_in
        self.h = [torch.empty((nmax, ly.n_out), **f) for ly in self.hidden]   # post-dropout activations
        self.d = [torch.empty((nmax, ly.n_out), **f) for ly in self.hidden]   # gradients w.r.t. the pre-activations
        self.prob = torch.zeros(nmax, **f)
        self.loss_terms = torch.zeros(nmax, **f)
        self.dlogit = torch.zeros(nmax, **f)
        self.loss = torch.zeros(max_groups + 1, **f)
        self.head_counter = torch.zeros(1, device=device, dtype=torch.int32)
        self.labels_train = torch.tensor([1.0, 0.0][:max_groups], **f)
        self.labels_ones = torch.ones(max_groups, **f)
        self.img = torch.empty((nmax, *self.shape), **f)          # real || X_d
        self.feedback = torch.empty((b, *self.shape), **f)
        # ---- dropout masks: [local epoch][layer] -> [2b, n_out] (rows 0..b-1 real, b..2b-1 X_d), then [layer] -> [b, n_out]
        self.scale = [_keep_scale(ly.drop_p) if ly.drop_p > 0 else 1.0 for ly in self.hidden]
        widths = [ly.n_out if ly.drop_p > 0 else 0 for ly in self.hidden]
        per_row = sum(widths)
        total = (self.local_epochs * 2 * b + b) * per_row
        self.mask_dev = torch.zeros(max(total, 1), device=device, dtype=torch.uint8)
        self.mask_host = torch.zeros(max(total, 1), dtype=torch.uint8, pin_memory=(device.type == "cuda"))
        self.mask_train: List[List[Optional[torch.Tensor]]] = []
        self._host_train: List[List[Optional[torch.Tensor]]] = []
        off = 0

        def carve(rows: int, w: int):
            nonlocal off
            if w == 0:
                return None, None
            dv = self.mask_dev[off: off + rows * w].view(rows, w)
            hv = self.mask_host[off: off + rows * w].view(rows, w)
            off += rows * w
            return dv, hv

        for _ in range(self.local_epochs):
            pairs = [carve(2 * b, w) for w in widths]
            self.mask_train.append([p[0] for p in pairs])
            self._host_train.append([p[1] for p in pairs])
        pairs = [carve(b, w) for w in widths]
        self.mask_fb = [p[0] for p in pairs]
        self._host_fb = [p[1] for p in pairs]
        self.has_dropout = per_row > 0
        if mask_source not in ("host", "device"):
            raise ValueError(f"mask_source must be 'host' or 'device', got {mask_source!r}")
        self.mask_source = mask_source
        if mask_source == "device":   # no host half at all (engine.stage_inputs / upload_inputs look these up)
            self.stage_host = self.upload_host = None
        self.rng = torch.Generator()
        state = rng_state if rng_state is not None else getattr(module, "_mdgan_rng_state", None)
        self.rng.set_state(state.clone() if state is not None else torch.get_rng_state())
        self._le = 0   # local epoch of the next train_step within the current iteration

    # ------------------------------------------------------------------ host half: the reference's dropout draws
    def stage_host(self) -> None:
        """Draw this iteration's keep masks in the reference's order (MNIST.py:86-94 under worker.py:197-198,222)."""
        if not self.has_dropout:
            return
        b = self.b
        for e in range(self.local_epochs):
            for part in (0, 1):   # D(real) first, then D(X_d)
                for l, ly in enumerate(self.hidden):
                    if ly.drop_p > 0:
                        noise = torch.empty((b, ly.n_out), dtype=torch.float32).bernoulli_(1.0 - ly.drop_p, generator=self.rng)
                        self._host_train[e][l][part * b:(part + 1) * b].copy_(noise)
        for l, ly in enumerate(self.hidden):   # the feedback pass D(X_g)
            if ly.drop_p > 0:
                noise = torch.empty((b, ly.n_out), dtype=torch.float32).bernoulli_(1.0 - ly.drop_p, generator=self.rng)
                self._host_fb[l].copy_(noise)

    def upload_host(self) -> None:
        """Host mode: stage_host() + upload_host() must run before every iteration (engine.stage_inputs / upload_inputs
        and standalone_gan.Standalone.step do); the step itself only reads the fixed device buffer."""
        if self.has_dropout:
            self.mask_dev.copy_(self.mask_host, non_blocking=True)

    def _draw_on_device(self, masks: List[Optional[torch.Tensor]]) -> None:
        """mask_source == "device": Bernoulli(1 - p) keep masks from the CUDA generator (graph-capturable)."""
        if self.mask_source == "device":
            for m, ly in zip(masks, self.hidden):
                if m is not None:
                    m.bernoulli_(1.0 - ly.drop_p)

    # ------------------------------------------------------------------ parameters
    def repack(self) -> None:
        """The SGEMM reads the PyTorch-layout weights directly: nothing to re-pack after Adam or a swap."""

    def adam(self) -> None:
        s = self.state
        ops.adam_step(s.params, s.grad, s.m, s.v, s.step, self.lr, self.beta_1, self.beta_2)

    # ------------------------------------------------------------------ forward / backward
    def forward(self, x: torch.Tensor, G: int, labels: torch.Tensor, masks: List[Optional[torch.Tensor]]) -> None:
        """x [G*b, n_in]; fills self.loss[0..G-1] (per-pass mean BCE) and self.loss[G] (their sum)."""
        n, P = G * self.b, self.state.p
        src = x
        for l, ly in enumerate(self.hidden):
            m = masks[l][:n] if masks[l] is not None else None
            ops.linear_forward(src, P[ly.weight], P[ly.bias] if ly.bias else None, self.h[l][:n], act=ops.ACT_LRELU,
                               slope=ly.slope, mask=m, mask_scale=self.scale[l])
            src = self.h[l][:n]
        head = self.L[-1]
        ops.linear_head_forward(src, P[head.weight].view(-1), P[head.bias] if head.bias else None, labels, self.prob,
                                self.loss_terms, self.dlogit, self.loss, self.head_counter, G, self.b)

    def backward(self, x: torch.Tensor, G: int, train: bool, masks: List[Optional[torch.Tensor]],
                 out: Optional[torch.Tensor] = None, accumulate: bool = False) -> None:
        """train=True: parameter gradients into state.grad.  train=False: dLoss/dx written (or accumulated) into `out`
        (default self.feedback)."""
        n, P, Gd = G * self.b, self.state.p, self.state.g
        head, last = self.L[-1], len(self.hidden) - 1
        m = masks[last][:n] if masks[last] is not None else None
        ops.linear_head_backward(self.h[last][:n], P[head.weight].view(-1), self.dlogit, self.d[last][:n],
                                 Gd[head.weight].view(-1) if train else None,
                                 Gd[head.bias] if (train and head.bias) else None,
                                 mask=m, mask_scale=self.scale[last], gate_slope=self.hidden[last].slope)
        for l in range(last, -1, -1):
            ly = self.hidden[l]
            src = self.h[l - 1][:n] if l > 0 else x
            if train:
                ops.linear_wgrad(self.d[l][:n], src, Gd[ly.weight])
                if ly.bias:
                    ops.col_sum(self.d[l][:n], Gd[ly.bias])
            if l > 0:
                pm = masks[l - 1][:n] if masks[l - 1] is not None else None
                ops.linear_dgrad(self.d[l][:n], P[ly.weight], self.d[l - 1][:n], gate=self.h[l - 1][:n],
                                 gate_slope=self.hidden[l - 1].slope, mask=pm, mask_scale=self.scale[l - 1])
            elif not train:
                dst = self.feedback if out is None else out
                ops.linear_dgrad(self.d[0][:n], P[ly.weight], dst.view(n, self.n_in), accumulate=accumulate)

    # ------------------------------------------------------------------ worker-level steps
    def train_step(self, real: torch.Tensor, x_d: torch.Tensor) -> torch.Tensor:
        """worker.py:197-206.  Returns a device view of d_loss = BCE(D(real),1) + BCE(D(X_d),0)."""
        b = self.b
        self.img[:b].copy_(real)
        self.img[b: 2 * b].copy_(x_d)
        masks = self.mask_train[min(self._le, self.local_epochs - 1)]
        self._le += 1
        x = self.img.view(self.nmax, self.n_in)[: 2 * b]
        self._draw_on_device(masks)
        self.forward(x, 2, self.labels_train, masks)
        self.backward(x, 2, True, masks)
        self.adam()
        return self.loss[2]

    def feedback_step(self, x_g: torch.Tensor, out: Optional[torch.Tensor] = None,
                      accumulate: bool = False) -> torch.Tensor:
        """worker.py:220-233.  dBCE(D(X_g),1)/dX_g goes to `out` (default self.feedback; accumulate=True adds)."""
        self._le = 0   # the feedback pass closes the iteration
        x = x_g.reshape(self.b, self.n_in)
        self._draw_on_device(self.mask_fb)
        self.forward(x, 1, self.labels_ones, self.mask_fb)
        self.backward(x, 1, False, self.mask_fb, out=out, accumulate=accumulate)
        return self.loss[0]


class MlpGenNet:
    def __init__(self, module: nn.Module, z_dim: int, image_shape: Tuple[int, int, int], n_samples: int,
                 device: torch.device, lr: float, beta_1: float, beta_2: float):
        self.device, self.n, self.z_dim, self.shape = device, n_samples, z_dim, tuple(image_shape)
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.plan: MlpPlan = extract_mlp_plan(_cpu_copy(module), "generator", (z_dim, 1, 1))
        self.state = FlatState(module, device)
        self.L = self.plan.layers
        self.hidden = self.L[:-1]
        n = n_samples
        f = dict(device=device, dtype=torch.float32)
        self.h = [torch.empty((n, ly.n_out), **f) for ly in self.hidden]
        self.d = [torch.empty((n, ly.n_out), **f) for ly in self.hidden]
        self.X = torch.empty((n, *self.shape), **f)
        self.dXt = torch.empty((n, *self.shape), **f)
        self.n_out = self.L[-1].n_out
        self._z: Optional[torch.Tensor] = None

    def repack(self) -> None:
        """Nothing to re-pack (see MlpDiscNet.repack)."""

    def adam(self) -> None:
        s = self.state
        ops.adam_step(s.params, s.grad, s.m, s.v, s.step, self.lr, self.beta_1, self.beta_2)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """z [n, z_dim] (device) -> X NCHW [n, C, H, W] (server.py:219-220; MNIST.py:110-120)."""
        n, P = self.n, self.state.p
        src = z.view(n, self.z_dim)
        self._z = src
        for l, ly in enumerate(self.hidden):
            ops.linear_forward(src, P[ly.weight], P[ly.bias] if ly.bias else None, self.h[l], act=ops.ACT_LRELU, slope=ly.slope)
            src = self.h[l]
        last = self.L[-1]
        ops.linear_forward(src, P[last.weight], P[last.bias] if last.bias else None, self.X.view(n, self.n_out),
                           act=ops.ACT_TANH)
        return self.X

    def backward(self, s: Optional[torch.Tensor], scale: float, slices=None) -> None:
        """grads = scale * J^T s for the per-sample sum s of the routed feedbacks (see nets.GenNet.backward); slices =
        (F [N, b, C, H, W], k, N): the per-worker feedbacks, summed per generated batch inside the tanh backward."""
        n, P, Gd = self.n, self.state.p, self.state.g
        if slices is not None:
            F_, k_, N_ = slices
            ops.tanh_backward_slices(F_, self.X, self.dXt, k_, N_, scale)
        else:
            ops.tanh_backward(s, self.X, self.dXt, scale)
        dy = self.dXt.view(n, self.n_out)
        for l in range(len(self.L) - 1, -1, -1):
            ly = self.L[l]
            src = self.h[l - 1] if l > 0 else self._z
            ops.linear_wgrad(dy, src, Gd[ly.weight])
            if ly.bias:
                ops.col_sum(dy, Gd[ly.bias])
            if l > 0:
                ops.linear_dgrad(dy, P[ly.weight], self.d[l - 1], gate=self.h[l - 1], gate_slope=self.hidden[l - 1].slope)
                dy = self.d[l - 1]
