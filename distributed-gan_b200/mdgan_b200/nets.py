"""Device-resident generator / discriminator engines built on the C-ABI kernels.

`DiscNet` executes what a reference worker does with its `nn.Module` discriminator
(/root/reference/src/actors/worker.py:193-233): the D training step on (real, X_d) with Adam, and the error
feedback dBCE(D(X_g),1)/dX_g.  `GenNet` executes the server's generator work
(/root/reference/src/actors/server.py:219-223,266-312): forward over the k*b noise batch, one backward on the
group-summed feedback (equal to the reference's N retain_graph VJPs by linearity), Adam.

Data layout in HBM (per net):
  state_f32  : [ parameters in module.parameters() order | BatchNorm running_mean/var in buffers() order ]  fp32,
               PyTorch tensor layouts, one contiguous allocation (so the discriminator swap is one send/recv);
  state_i64  : BatchNorm num_batches_tracked counters;
  grad, m, v : flat fp32 mirrors of the parameter prefix (Adam is a single launch over it);
  packed weights (TF32-rounded, K-major tensor-core operands), rebuilt after every Adam step;
  activations: NHWC fp32, pre-BN conv outputs `z`, post-activation `a`, and their gradients `da`, `dz`.
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .plan import ACT_LRELU, ACT_RELU, ConvLayer, NetPlan, extract_plan

_ACT_CODE = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU}
# MDGAN_BN_FUSED_STATS = 1 (default) | 0: BatchNorm statistics reduced in the producing GEMM's epilogue (no separate
# pass over the layer output) wherever a CTA's rows belong to one BatchNorm pass; 0 keeps the two-kernel bn_forward.
_FUSED_STATS = os.environ.get("MDGAN_BN_FUSED_STATS", "1") == "1"


def _stats_floats(plan, n_pad: int) -> int:
    row_tiles, _, phases = plan
    return phases * row_tiles * 2 * n_pad


def _conv_bn_act(src, wpacked, mode, c_out, z, a, grid, src_hw, bias, prec, bn_part, bnp, P, B, stats, bn_ws, bn_cnt, G, Pg,
                 act, slope, round_tf32, n_cols=None, fold=1):
    """conv_gemm -> train-mode BatchNorm -> activation.  With fused statistics: GEMM (+ column sums in its epilogue) ->
    bn_finalize -> bn_apply; otherwise GEMM -> bn_forward (statistics pass + apply)."""
    n_cols = c_out if n_cols is None else n_cols
    plan = ops.conv_stats_plan(grid, mode, G, prec, n_cols) if (_FUSED_STATS and bn_part is not None) else None
    nbt = B[bnp.num_batches_tracked] if bnp.num_batches_tracked else None
    if plan is not None and _stats_floats(plan, ops.n_pad_for(n_cols)) <= bn_part.numel():
        ops.conv_gemm(src, wpacked, mode, n_cols, z, grid, src_hw, bias=bias, precision=prec, bn_partial=bn_part)
        ops.bn_finalize(bn_part, plan, ops.n_pad_for(n_cols), fold, P[bnp.weight], P[bnp.bias], B[bnp.running_mean],
                        B[bnp.running_var], nbt, stats, G, Pg, c_out, eps=bnp.eps, momentum=bnp.momentum)
        ops.bn_apply(z, stats, a, G, Pg, c_out, act, slope, round_tf32=round_tf32)
        return
    ops.conv_gemm(src, wpacked, mode, n_cols, z, grid, src_hw, bias=bias, precision=prec)
    ops.bn_forward(z, a, P[bnp.weight], P[bnp.bias], B[bnp.running_mean], B[bnp.running_var], nbt, stats, bn_ws, bn_cnt, G,
                   Pg, c_out, act, slope, round_tf32=round_tf32, eps=bnp.eps, momentum=bnp.momentum)


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _bn_backward(fused_plan, bn_part, n_pad, da, z, stats, dz, dgamma, dbeta, sums, bn_ws, bn_cnt, G, Pg, Cc, act, slope,
                 round_tf32):
    """BatchNorm(+activation) backward.  fused_plan is not None: `da` already holds dy = da * act'(.) and bn_part the
    partial sums, both written by the data-gradient GEMM that produced it (ops.conv_gemm bnb=...)."""
    if fused_plan is not None:
        ops.bn_bwd_finalize(bn_part, fused_plan, n_pad, sums, dgamma, dbeta, G, Cc)
        ops.bn_bwd_apply_dy(da, z, stats, sums, dz, G, Pg, Cc, round_tf32=round_tf32)
    else:
        ops.bn_backward(da, z, stats, dz, dgamma, dbeta, sums, bn_ws, bn_cnt, G, Pg, Cc, act, slope, round_tf32=round_tf32)


class FlatState:
    """Flat device copy of a module's parameters + buffers (see module docstring)."""

    def __init__(self, module: nn.Module, device: torch.device):
        self.device = device
        self.param_names = [n for n, _ in module.named_parameters()]
        self.param_shapes = {n: tuple(p.shape) for n, p in module.named_parameters()}
        fbuf = [(n, b) for n, b in module.named_buffers() if b.dtype == torch.float32]
        ibuf = [(n, b) for n, b in module.named_buffers() if b.dtype == torch.int64]
        other = [n for n, b in module.named_buffers() if b.dtype not in (torch.float32, torch.int64)]
        if other or any(p.dtype != torch.float32 for p in module.parameters()):
            raise TypeError(f"only fp32 parameters and fp32/int64 buffers are supported (got {other})")
        self.fbuf_names = [n for n, _ in fbuf]
        self.ibuf_names = [n for n, _ in ibuf]
        self.key_order = list(module.state_dict().keys())
        self.n_params = sum(p.numel() for p in module.parameters())
        n_f = self.n_params + sum(b.numel() for _, b in fbuf)
        self.state_f32 = torch.zeros(n_f, device=device, dtype=torch.float32)
        self.state_i64 = torch.zeros(max(len(ibuf), 1), device=device, dtype=torch.int64)
        self.params = self.state_f32[: self.n_params]
        self.grad = torch.zeros(self.n_params, device=device, dtype=torch.float32)
        self.m = torch.zeros_like(self.grad)
        self.v = torch.zeros_like(self.grad)
        self.step = torch.zeros(2, device=device, dtype=torch.int32)   # (steps taken, adam block counter)
        self.p: Dict[str, torch.Tensor] = {}
        self.g: Dict[str, torch.Tensor] = {}
        self.b: Dict[str, torch.Tensor] = {}
        off = 0
        for n, prm in module.named_parameters():
            k = prm.numel()
            self.p[n] = self.state_f32[off: off + k].view(prm.shape)
            self.g[n] = self.grad[off: off + k].view(prm.shape)
            off += k
        for n, bf in fbuf:
            k = bf.numel()
            self.b[n] = self.state_f32[off: off + k].view(bf.shape)
            off += k
        for i, (n, _) in enumerate(ibuf):
            self.b[n] = self.state_i64[i: i + 1].view(())
        self.load_from(module)

    @torch.no_grad()
    def load_from(self, module: nn.Module) -> None:
        for n, prm in module.named_parameters():
            self.p[n].copy_(prm.detach().to(self.device, non_blocking=True))
        for n, bf in module.named_buffers():
            self.b[n].copy_(bf.detach().to(self.device, non_blocking=True))

    @torch.no_grad()
    def store_to(self, module: nn.Module) -> None:
        """Write the device state back into the caller's module (state_dict key order and dtypes preserved)."""
        for n, prm in module.named_parameters():
            prm.data.copy_(self.p[n].to(prm.device))
        for n, bf in module.named_buffers():
            bf.data.copy_(self.b[n].to(bf.device))

    @torch.no_grad()
    def load_adam(self, optimizer: torch.optim.Optimizer, module: nn.Module) -> None:
        """Adopt the moments and step count of a torch.optim.Adam that optimises `module.parameters()` (resume /
        hand-over of a run started elsewhere; the reference never saves optimizer state)."""
        self.m.zero_()
        self.v.zero_()
        step, off = 0, 0
        for prm in module.parameters():
            k = prm.numel()
            st = optimizer.state.get(prm, {})
            if st:
                self.m[off: off + k].copy_(st["exp_avg"].detach().reshape(-1).to(self.device, torch.float32))
                self.v[off: off + k].copy_(st["exp_avg_sq"].detach().reshape(-1).to(self.device, torch.float32))
                step = int(st["step"])
            off += k
        self.step[0:1].fill_(step)

    def state_dict_from(self, host_f32: torch.Tensor, host_i64: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The module's `state_dict()` (key order preserved) rebuilt from host copies of the two flat buffers -- used
        by the asynchronous checkpoint writer (node._SnapshotWriter)."""
        out: Dict[str, torch.Tensor] = {}
        off = 0
        for n in self.param_names:
            k = 1
            for d in self.param_shapes[n]:
                k *= d
            out[n] = host_f32[off: off + k].view(self.param_shapes[n]).clone()
            off += k
        for n in self.fbuf_names:
            shape = tuple(self.b[n].shape)
            k = self.b[n].numel()
            out[n] = host_f32[off: off + k].view(shape).clone()
            off += k
        for i, n in enumerate(self.ibuf_names):
            out[n] = host_i64[i].clone()
        return {n: out[n] for n in self.key_order}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Host copy in the module's own `state_dict()` key order (what torch.save of the reference writes)."""
        return {n: (self.p[n] if n in self.p else self.b[n]).detach().cpu().clone() for n in self.key_order}


def _cpu_copy(module: nn.Module) -> nn.Module:
    return copy.deepcopy(module).to("cpu")


class DiscNet:
    def __init__(self, module: nn.Module, image_shape: Tuple[int, int, int], batch_size: int, device: torch.device,
                 lr: float, beta_1: float, beta_2: float, max_groups: int = 2, precision: Optional[int] = None):
        self.device, self.b, self.shape = device, batch_size, tuple(image_shape)
        self.prec = ops.default_precision() if precision is None else precision
        self.rnd = self.prec == ops.TF32  # single-pass TF32: producers round tensor-core operands to nearest
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.plan: NetPlan = extract_plan(_cpu_copy(module), "discriminator", self.shape)
        self.state = FlatState(module, device)
        L = self.plan.layers
        self.L = L
        nmax = max_groups * batch_size
        self.nmax, self.max_groups = nmax, max_groups
        f = dict(device=device, dtype=torch.float32)
        self.z: List[Optional[torch.Tensor]] = []
        self.a: List[torch.Tensor] = []
        self.da: List[torch.Tensor] = []
        self.dz: List[torch.Tensor] = []
        self.stats: List[Optional[torch.Tensor]] = []
        self.sums: List[Optional[torch.Tensor]] = []
        self.wp: List[Optional[torch.Tensor]] = []
        self.wq: List[Optional[torch.Tensor]] = []
        self.partial: List[Optional[torch.Tensor]] = []
        ws = 0
        for l, ly in enumerate(L[:-1]):
            Ho, Cc = ly.h_out, ly.c_out
            shape = (nmax, Ho, Ho, Cc)
            self.z.append(torch.empty(shape, **f) if ly.bn else None)
            self.a.append(torch.empty(shape, **f))
            self.da.append(torch.empty(shape, **f))
            self.dz.append(torch.empty(shape, **f))
            self.stats.append(torch.zeros(max_groups * 4 * Cc, **f) if ly.bn else None)
            self.sums.append(torch.zeros(max_groups * 2 * Cc, **f) if ly.bn else None)
            if ly.bn:
                for G in range(1, max_groups + 1):
                    ws = max(ws, ops.bn_workspace_floats(G, batch_size * Ho * Ho, Cc))
            # conv data-grad operand (UP): conv weight [co][ci] read as [C=co][N=ci]; the image-side layer 0 runs on
            # CUDA cores straight from the PyTorch-layout weight (ops.thin_up)
            self.wq.append(torch.empty(ops.packed_shape(ops.MODE_UP, ly.c_in, Cc, precision=self.prec), **f)
                           if l >= 1 else None)
            if l >= 1:
                self.wp.append(torch.empty(ops.packed_shape(ops.MODE_DOWN, Cc, ly.c_in, precision=self.prec), **f))
                splits = ops.wgrad_splits(nmax, Ho, Ho, Cc, ly.c_in, ops.MODE_DOWN)
                self.partial.append(torch.empty(splits * 16 * Cc * ly.c_in, **f))
            else:
                self.wp.append(None)
                self.partial.append(torch.empty(ops.thin_wgrad_slices(nmax, Ho, Ho) * Cc * ly.c_in * 16, **f))
        self.bn_ws = torch.empty(max(ws, 1), **f)
        self.bn_cnt = ops.bn_counters(device)
        part = 0
        for l, ly in enumerate(L[:-1]):
            if ly.bn and l >= 1:
                for G in range(1, max_groups + 1):
                    plan = ops.conv_stats_plan((G * batch_size, ly.h_out, ly.h_out), ops.MODE_DOWN, G, self.prec)
                    if plan is not None:
                        part = max(part, _stats_floats(plan, ops.n_pad_for(ly.c_out)))
                    if l >= 2:  # the data-gradient GEMM of layer l reduces the BatchNorm-backward sums of layer l - 1
                        plan = ops.conv_stats_plan((G * batch_size, ly.h_out, ly.h_out), ops.MODE_UP, G, self.prec, ly.c_in)
                        if plan is not None:
                            part = max(part, _stats_floats(plan, ops.n_pad_for(ly.c_in)))
        self.bn_part = torch.empty(part, **f) if part else None
        head = L[-1]
        self.w_head = torch.empty(head.k * head.k * head.c_in, **f)
        self.prob = torch.zeros(nmax, **f)
        self.loss_terms = torch.zeros(nmax, **f)
        self.dlogit = torch.zeros(nmax, **f)
        self.loss = torch.zeros(max_groups + 1, **f)
        self.head_counter = torch.zeros(1, device=device, dtype=torch.int32)
        self.labels_train = torch.tensor([1.0, 0.0][:max_groups], **f)
        self.labels_ones = torch.ones(max_groups, **f)
        self.img = torch.empty((nmax, *self.shape), **f)          # real || X_d
        self.feedback = torch.empty((batch_size, *self.shape), **f)
        self._packs: Optional[ops.PackPlan] = None
        self.repack()

    # ------------------------------------------------------------------ parameters
    def repack(self) -> None:
        """One launch re-packs every tensor-core operand + the head weight from the (just updated) parameters."""
        if self._packs is None:
            P = self.state.p
            plan = ops.PackPlan(self.device)
            for l, ly in enumerate(self.L[:-1]):
                if l >= 1:
                    plan.add_up(P[ly.weight], self.wq[l])
                    plan.add_down(P[ly.weight], self.wp[l])
            plan.add_head(P[self.L[-1].weight], self.w_head.view(-1, self.L[-1].c_in))
            self._packs = plan.finalize()
        self._packs.run()

    def adam(self) -> None:
        s = self.state
        ops.adam_step(s.params, s.grad, s.m, s.v, s.step, self.lr, self.beta_1, self.beta_2)
        self.repack()

    # ------------------------------------------------------------------ forward / backward
    def forward(self, img: torch.Tensor, G: int, labels: torch.Tensor) -> None:
        """img NCHW [G*b, C, H, W]; fills self.loss[0..G-1] (per-pass mean BCE) and self.loss[G] (their sum)."""
        b, P, B = self.b, self.state.p, self.state.b
        n = G * b
        L = self.L
        l0 = L[0]
        ops.thin_down(img[:n], P[l0.weight], self.a[0][:n], act=ops.ACT_LRELU, slope=l0.slope, round_tf32=self.rnd)
        for l in range(1, len(L) - 1):
            ly = L[l]
            Ho = ly.h_out
            _conv_bn_act(self.a[l - 1][:n], self.wp[l], ops.MODE_DOWN, ly.c_out, self.z[l][:n], self.a[l][:n], (n, Ho, Ho),
                         (ly.h_in, ly.h_in), P[ly.bias] if ly.bias else None, self.prec, self.bn_part, ly.bn, P, B,
                         self.stats[l], self.bn_ws, self.bn_cnt, G, b * Ho * Ho, _ACT_CODE[ly.act], ly.slope,
                         self.rnd and L[l + 1].kind == "down")
        head = L[-1]
        ops.head_forward(self.a[-1][:n], self.w_head, labels, self.prob, self.loss_terms, self.dlogit, self.loss,
                         self.head_counter, G, b, head.k * head.k, head.c_in)

    def backward(self, img: torch.Tensor, G: int, train: bool, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> None:
        """train=True: parameter gradients into state.grad.  train=False: dLoss/d(img) written (or accumulated)
        into `out` (default self.feedback), NCHW."""
        b, P, Gd = self.b, self.state.p, self.state.g
        n = G * b
        L = self.L
        head = L[-1]
        ops.head_backward(self.a[-1][:n], self.w_head, self.dlogit, self.da[-1][:n],
                          Gd[head.weight] if train else None, n, head.k * head.k, head.c_in)
        fused = None   # plan of the fused BatchNorm-backward reduction of the NEXT layer down, if its producer did it
        for l in range(len(L) - 2, 0, -1):
            ly = L[l]
            Ho, bn = ly.h_out, ly.bn
            _bn_backward(fused, self.bn_part, ops.n_pad_for(ly.c_out), self.da[l][:n], self.z[l][:n], self.stats[l],
                         self.dz[l][:n], Gd[bn.weight] if train else None, Gd[bn.bias] if train else None, self.sums[l],
                         self.bn_ws, self.bn_cnt, G, b * Ho * Ho, ly.c_out, _ACT_CODE[ly.act], ly.slope, self.rnd)
            fused = None
            if train:  # weight gradient on the side stream, next to the data gradient below
                with ops.side_branch():
                    splits = ops.wgrad_splits(n, Ho, Ho, ly.c_out, ly.c_in, ops.MODE_DOWN)
                    ops.wgrad_gemm(self.dz[l][:n], self.a[l - 1][:n], self.partial[l], (n, Ho, Ho), ops.MODE_DOWN,
                                   splits, precision=self.prec)
                    ops.wgrad_unpack(self.partial[l], Gd[ly.weight], ops.MODE_DOWN, splits, ly.c_out, ly.c_out, ly.c_in)
            if l == 1:  # the LeakyReLU backward of layer 0 rides in this GEMM's epilogue: dz[0] = dgrad * act'(a[0])
                ops.conv_gemm(self.dz[l][:n], self.wq[l], ops.MODE_UP, ly.c_in, self.dz[0][:n], (n, Ho, Ho), (Ho, Ho),
                              precision=self.prec, gate=self.a[0][:n], gate_act=ops.ACT_LRELU, gate_slope=L[0].slope,
                              round_tf32=(self.rnd and not train))
            else:
                prev = L[l - 1]
                plan = ops.conv_stats_plan((n, Ho, Ho), ops.MODE_UP, G, self.prec, ly.c_in) \
                    if (_FUSED_STATS and self.bn_part is not None and not self.rnd and ly.c_in % 16 == 0) else None
                if plan is not None and _stats_floats(plan, ops.n_pad_for(ly.c_in)) <= self.bn_part.numel():
                    ops.conv_gemm(self.dz[l][:n], self.wq[l], ops.MODE_UP, ly.c_in, self.da[l - 1][:n], (n, Ho, Ho), (Ho, Ho),
                                  precision=self.prec, bn_partial=self.bn_part,
                                  bnb=(self.z[l - 1][:n], self.stats[l - 1], _ACT_CODE[prev.act], prev.slope, G))
                    fused = plan
                else:
                    ops.conv_gemm(self.dz[l][:n], self.wq[l], ops.MODE_UP, ly.c_in, self.da[l - 1][:n], (n, Ho, Ho),
                                  (Ho, Ho), precision=self.prec)
        l0 = L[0]
        if train:
            ops.thin_wgrad(self.dz[0][:n], img[:n], self.partial[0], Gd[l0.weight])
            ops.join_side()
        else:
            ops.thin_up(self.dz[0][:n], P[l0.weight], self.feedback if out is None else out, accumulate=accumulate)

    # ------------------------------------------------------------------ worker-level steps
    def train_step(self, real: torch.Tensor, x_d: torch.Tensor) -> torch.Tensor:
        """worker.py:197-206.  Returns a device view of d_loss = BCE(D(real),1) + BCE(D(X_d),0)."""
        b = self.b
        self.img[:b].copy_(real)
        self.img[b: 2 * b].copy_(x_d)
        self.forward(self.img, 2, self.labels_train)
        self.backward(self.img, 2, train=True)
        self.adam()
        return self.loss[2]

    def feedback_step(self, x_g: torch.Tensor, out: Optional[torch.Tensor] = None,
                      accumulate: bool = False) -> torch.Tensor:
        """worker.py:220-233.  dBCE(D(X_g),1)/dX_g goes to `out` (default self.feedback; accumulate=True adds, which
        is how feedbacks of workers sharing a generated batch are summed).  Returns the loss_gen device view."""
        self.forward(x_g, 1, self.labels_ones)
        self.backward(x_g, 1, train=False, out=out, accumulate=accumulate)
        return self.loss[0]


class GenNet:
    def __init__(self, module: nn.Module, z_dim: int, image_shape: Tuple[int, int, int], n_samples: int,
                 device: torch.device, lr: float, beta_1: float, beta_2: float, precision: Optional[int] = None):
        self.device, self.n, self.z_dim, self.shape = device, n_samples, z_dim, tuple(image_shape)
        self.prec = ops.default_precision() if precision is None else precision
        self.rnd = self.prec == ops.TF32
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.plan: NetPlan = extract_plan(_cpu_copy(module), "generator", (z_dim, 1, 1))
        self.state = FlatState(module, device)
        L = self.plan.layers
        self.L = L
        n = n_samples
        f = dict(device=device, dtype=torch.float32)
        self.zc = _pad(z_dim, 32)
        self.zp = torch.zeros((n, self.zc), **f)
        self.z, self.a, self.da, self.dz, self.stats, self.sums = [], [], [], [], [], []
        self.wq: List[Optional[torch.Tensor]] = []      # forward operands (UP), l >= 1
        self.wp_dg: List[Optional[torch.Tensor]] = []   # data-grad operands (DOWN), 1 <= l <= L-2
        self.partial: List[Optional[torch.Tensor]] = []
        ws = 0
        for l, ly in enumerate(L[:-1]):
            Ho, Cc = ly.h_out, ly.c_out
            shape = (n, Ho, Ho, Cc)
            self.z.append(torch.empty(shape, **f))
            self.a.append(torch.empty(shape, **f))
            self.da.append(torch.empty(shape, **f))
            self.dz.append(torch.empty(shape, **f))
            self.stats.append(torch.zeros(4 * Cc, **f))
            self.sums.append(torch.zeros(2 * Cc, **f))
            ws = max(ws, ops.bn_workspace_floats(1, n * Ho * Ho, Cc))
        self.bn_ws = torch.empty(max(ws, 1), **f)
        self.bn_cnt = ops.bn_counters(device)
        l0 = L[0]
        self.kk = l0.k * l0.k
        part = 0
        plan = ops.conv_stats_plan((n, 1, 1), ops.MODE_DENSE, 1, self.prec)
        if plan is not None:
            part = _stats_floats(plan, ops.n_pad_for(self.kk * l0.c_out))
        for ly in L[1:-1]:
            plan = ops.conv_stats_plan((n, ly.h_in, ly.h_in), ops.MODE_UP, 1, self.prec, ly.c_out)
            if plan is not None:
                part = max(part, _stats_floats(plan, ops.n_pad_for(ly.c_out)))
            plan = ops.conv_stats_plan((n, ly.h_in, ly.h_in), ops.MODE_DOWN, 1, self.prec)  # data gradient: BN bwd of l - 1
            if plan is not None:
                part = max(part, _stats_floats(plan, ops.n_pad_for(ly.c_in)))
        self.bn_part = torch.empty(part, **f) if part else None
        self.wp_dense = torch.empty(ops.packed_shape(ops.MODE_DENSE, l0.c_out, z_dim, self.kk, self.prec), **f)
        sp0 = ops.wgrad_splits(n, 1, 1, self.zc, self.kk * l0.c_out, ops.MODE_DENSE)
        self.partial.append(torch.empty(sp0 * self.zc * self.kk * l0.c_out, **f))
        self.wq.append(None)
        self.wp_dg.append(None)
        for l in range(1, len(L)):
            ly = L[l]
            self.wq.append(torch.empty(ops.packed_shape(ops.MODE_UP, ly.c_out, ly.c_in, precision=self.prec), **f)
                           if l < len(L) - 1 else None)   # the image-side last layer runs on CUDA cores (ops.thin_up)
            if l < len(L) - 1:
                self.wp_dg.append(torch.empty(ops.packed_shape(ops.MODE_DOWN, ly.c_in, ly.c_out, precision=self.prec), **f))
                sp = ops.wgrad_splits(n, ly.h_in, ly.h_in, ly.c_in, ly.c_out, ops.MODE_DOWN)
                self.partial.append(torch.empty(sp * 16 * ly.c_in * ly.c_out, **f))
            else:
                self.wp_dg.append(None)
                self.partial.append(torch.empty(ops.thin_wgrad_slices(n, ly.h_in, ly.h_in) * ly.c_in * ly.c_out * 16, **f))
        self.X = torch.empty((n, *self.shape), **f)
        self.dXt = torch.empty((n, *self.shape), **f)
        self._packs: Optional[ops.PackPlan] = None
        self.repack()

    def repack(self) -> None:
        if self._packs is None:
            P, L = self.state.p, self.L
            plan = ops.PackPlan(self.device)
            plan.add_dense(P[L[0].weight], self.wp_dense)
            for l in range(1, len(L) - 1):
                plan.add_up(P[L[l].weight], self.wq[l])
                plan.add_down(P[L[l].weight], self.wp_dg[l])
            self._packs = plan.finalize()
        self._packs.run()

    def adam(self) -> None:
        s = self.state
        ops.adam_step(s.params, s.grad, s.m, s.v, s.step, self.lr, self.beta_1, self.beta_2)
        self.repack()

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """z [n, z_dim] (device) -> X NCHW [n, C, H, W]; train-mode BatchNorm over all n samples (server.py:219-220)."""
        n, P, B, L = self.n, self.state.p, self.state.b, self.L
        ops.pad_rows(z.view(n, self.z_dim), self.zp, round_tf32=self.rnd)
        for l in range(len(L) - 1):
            ly = L[l]
            Ho = ly.h_out
            if l == 0:   # ConvTranspose2d on a 1x1 input: GEMM columns are (position, channel)
                _conv_bn_act(self.zp, self.wp_dense, ops.MODE_DENSE, ly.c_out, self.z[0], self.a[0], (n, 1, 1), (1, 1), None,
                             self.prec, self.bn_part, ly.bn, P, B, self.stats[0], self.bn_ws, self.bn_cnt, 1, n * Ho * Ho,
                             _ACT_CODE[ly.act], ly.slope, self.rnd, n_cols=self.kk * ly.c_out, fold=self.kk)
            else:
                _conv_bn_act(self.a[l - 1], self.wq[l], ops.MODE_UP, ly.c_out, self.z[l], self.a[l], (n, ly.h_in, ly.h_in),
                             (ly.h_in, ly.h_in), None, self.prec, self.bn_part, ly.bn, P, B, self.stats[l], self.bn_ws,
                             self.bn_cnt, 1, n * Ho * Ho, _ACT_CODE[ly.act], ly.slope, self.rnd)
        last = L[-1]
        ops.thin_up(self.a[-1], P[last.weight], self.X, act_tanh=True)
        return self.X

    def backward(self, s: Optional[torch.Tensor], scale: float, slices=None) -> None:
        """s NCHW [n, C, H, W]: per-sample sum of the feedbacks routed to that sample; grads = scale * J^T s.
        slices = (F [N, b, C, H, W], k, N) instead of s: the per-worker feedbacks, summed per generated batch inside
        the tanh-backward kernel (peer-memory exchange)."""
        n, P, Gd, L = self.n, self.state.p, self.state.g, self.L
        last = L[-1]
        if slices is not None:
            F_, k_, N_ = slices
            ops.tanh_backward_slices(F_, self.X, self.dXt, k_, N_, scale)
        else:
            ops.tanh_backward(s, self.X, self.dXt, scale)
        with ops.side_branch():  # weight gradients run next to the data-gradient chain (ops.side_branch)
            ops.thin_wgrad(self.a[-1], self.dXt, self.partial[-1], Gd[last.weight])
        ops.thin_down(self.dXt, P[last.weight], self.da[-1], act=ops.ACT_NONE)
        fused = None
        for l in range(len(L) - 2, -1, -1):
            ly, bn = L[l], L[l].bn
            Ho = ly.h_out
            _bn_backward(fused, self.bn_part, ops.n_pad_for(ly.c_out), self.da[l], self.z[l], self.stats[l], self.dz[l],
                         Gd[bn.weight], Gd[bn.bias], self.sums[l], self.bn_ws, self.bn_cnt, 1, n * Ho * Ho, ly.c_out,
                         _ACT_CODE[ly.act], ly.slope, self.rnd)
            fused = None
            if l >= 1:
                hi = ly.h_in
                with ops.side_branch():
                    splits = ops.wgrad_splits(n, hi, hi, ly.c_in, ly.c_out, ops.MODE_DOWN)
                    ops.wgrad_gemm(self.a[l - 1], self.dz[l], self.partial[l], (n, hi, hi), ops.MODE_DOWN, splits,
                                   precision=self.prec)
                    ops.wgrad_unpack(self.partial[l], Gd[ly.weight], ops.MODE_DOWN, splits, ly.c_in, ly.c_in, ly.c_out)
                prev = L[l - 1]
                plan = ops.conv_stats_plan((n, hi, hi), ops.MODE_DOWN, 1, self.prec) \
                    if (_FUSED_STATS and self.bn_part is not None and not self.rnd and ly.c_in % 16 == 0) else None
                if plan is not None and _stats_floats(plan, ops.n_pad_for(ly.c_in)) <= self.bn_part.numel():
                    ops.conv_gemm(self.dz[l], self.wp_dg[l], ops.MODE_DOWN, ly.c_in, self.da[l - 1], (n, hi, hi), (Ho, Ho),
                                  precision=self.prec, bn_partial=self.bn_part,
                                  bnb=(self.z[l - 1], self.stats[l - 1], _ACT_CODE[prev.act], prev.slope, 1))
                    fused = plan
                else:
                    ops.conv_gemm(self.dz[l], self.wp_dg[l], ops.MODE_DOWN, ly.c_in, self.da[l - 1], (n, hi, hi), (Ho, Ho),
                                  precision=self.prec)
            else:
                c2 = self.kk * ly.c_out
                splits = ops.wgrad_splits(n, 1, 1, self.zc, c2, ops.MODE_DENSE)
                ops.wgrad_gemm(self.zp, self.dz[0].view(n, 1, 1, c2), self.partial[0], (n, 1, 1), ops.MODE_DENSE, splits,
                               precision=self.prec)
                ops.wgrad_unpack(self.partial[0], Gd[ly.weight], ops.MODE_DENSE, splits, self.z_dim, self.zc, c2,
                                 N=ly.c_out, KK=self.kk)
        ops.join_side()
