"""Tensor-level wrappers over the kernel library (one function per exported kernel family).

Every launch goes through the torch custom-op layer `torch.ops.mdgan_b200.*` (mdgan_b200/torch_ops.py), a thin shim
over the C-ABI of include/mdgan_b200.h: tensors in, raw device pointers + sizes + the current CUDA stream out.
Activations are NHWC fp32 CUDA tensors, parameters keep their PyTorch layout.  Launches are stream-ordered, so the
calls can be captured into a CUDA graph.  No fallback: a non-CUDA tensor or a failed launch raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .torch_ops import ns as _K

import os

MODE_DOWN, MODE_UP, MODE_DENSE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
TF32, TF32X3 = 0, 1
_PRECISION_NAMES = {"tf32": TF32, "tf32x3": TF32X3}


def default_precision() -> int:
    """MDGAN_PRECISION = tf32x3 (default; ~fp32 accuracy, the parity mode) | tf32 (single-pass, cuDNN-TF32-like)."""
    name = os.environ.get("MDGAN_PRECISION", "tf32x3").lower()
    if name not in _PRECISION_NAMES:
        raise ValueError(f"MDGAN_PRECISION must be one of {sorted(_PRECISION_NAMES)}, got {name!r}")
    return _PRECISION_NAMES[name]


# Optional launch observer (bench.py / tools): called as observer(name, n_kernels, flops, bytes) and must return a
# context manager entered around the launch.  flops / bytes are the ALGORITHMIC work of the call.
_observer = None


def set_observer(observer) -> None:
    global _observer
    _observer = observer


def _run(name: str, n_kernels: int, flops: float, nbytes: float, call) -> None:
    if _observer is None:
        call()
        return
    with _observer(name, n_kernels, flops, nbytes):
        call()


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.MdganLibraryError("MD-GAN kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.MdganLibraryError("MD-GAN kernels need contiguous tensors")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


# ----------------------------------------------------------------------------- branch overlap (fork / join)
# The weight gradient of a layer and the data gradient that feeds the next layer's backward are independent, so the
# weight-gradient branch can run on a side stream (also inside a captured CUDA graph, where the fork/join become
# graph edges).  Opt-in (MDGAN_OVERLAP=1): both GEMMs are already sized to one wave of one-CTA-per-SM tiles (split-K
# in the weight gradient), so on B200 the overlapped step measured 0.607 ms against 0.592 ms on one stream (MNIST
# shape, b = 64) and 2.52 against 2.54 ms (CelebA shape).
_side_streams = {}
_overlap = os.environ.get("MDGAN_OVERLAP", "0") == "1"


class side_branch:
    """with side_branch(): launches go to the device's side stream, ordered after everything already queued on the
    current stream.  join_side() makes the current stream wait for the side stream."""

    def __enter__(self):
        self._ctx = None
        if not _overlap:
            return self
        main = torch.cuda.current_stream()
        key = main.device.index
        side = _side_streams.get(key)
        if side is None:
            side = _side_streams[key] = torch.cuda.Stream(device=main.device)
        side.wait_stream(main)
        self._ctx = torch.cuda.stream(side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


def join_side() -> None:
    if not _overlap:
        return
    main = torch.cuda.current_stream()
    side = _side_streams.get(main.device.index)
    if side is not None:
        main.wait_stream(side)


def n_pad_for(n: int) -> int:
    """Packed-weight row padding: a multiple of the narrowest tensor-core tile (16)."""
    return _pad(n, 16) if n < 32 else _pad(n, 32)


# ----------------------------------------------------------------------------- weights
def packed_shape(mode: int, N: int, Cc: int, KK: int = 16, precision: int = TF32) -> Tuple[int, int]:
    """Shape of the packed operand; tf32x3 stacks the `lo` matrix under the `hi` one (twice the rows)."""
    m = 2 if precision == TF32X3 else 1
    if mode == MODE_DOWN:
        return m * n_pad_for(N), 16 * _pad(Cc, 32)
    if mode == MODE_UP:
        return m * 4 * n_pad_for(N), 4 * _pad(Cc, 32)
    return m * KK * N, _pad(Cc, 32)


def _split_of(out: torch.Tensor, rows_hi: int) -> int:
    if out.shape[0] == rows_hi:
        return 0
    if out.shape[0] == 2 * rows_hi:
        return 1
    raise _lib.MdganLibraryError(f"packed weight buffer has {out.shape[0]} rows, expected {rows_hi} or {2 * rows_hi}")


def pack_down(W: torch.Tensor, out: Optional[torch.Tensor] = None, precision: int = TF32) -> torch.Tensor:
    """W [N, C, 4, 4] -> [N_pad, 16*C_pad] (K-major, K = (tap, c)); conv fwd / convT dgrad operand."""
    N, Cc = W.shape[0], W.shape[1]
    Np, Cp = n_pad_for(N), _pad(Cc, 32)
    if out is None:
        out = torch.empty(packed_shape(MODE_DOWN, N, Cc, precision=precision), device=W.device, dtype=torch.float32)
    _run("pack_weights", 1, 0, 4 * (W.numel() + out.numel()),
         lambda: _K.pack_weights(W, out, 0, N, Cc, Np, Cp, 16, _split_of(out, Np)))
    return out


def pack_up(W: torch.Tensor, out: Optional[torch.Tensor] = None, precision: int = TF32) -> torch.Tensor:
    """W [C, N, 4, 4] -> [4 phases * N_pad, 4*C_pad]; convT fwd / conv dgrad operand."""
    Cc, N = W.shape[0], W.shape[1]
    Np, Cp = n_pad_for(N), _pad(Cc, 32)
    if out is None:
        out = torch.empty(packed_shape(MODE_UP, N, Cc, precision=precision), device=W.device, dtype=torch.float32)
    _run("pack_weights", 1, 0, 4 * (W.numel() + out.numel()),
         lambda: _K.pack_weights(W, out, 1, N, Cc, Np, Cp, 16, _split_of(out, 4 * Np)))
    return out


def pack_dense(W: torch.Tensor, out: Optional[torch.Tensor] = None, precision: int = TF32) -> torch.Tensor:
    """W [C, N, k, k] (ConvTranspose2d on a 1x1 input) -> [k*k*N, C_pad]."""
    Cc, N, KK = W.shape[0], W.shape[1], W.shape[2] * W.shape[3]
    Cp = _pad(Cc, 32)
    if out is None:
        out = torch.empty(packed_shape(MODE_DENSE, N, Cc, KK, precision), device=W.device, dtype=torch.float32)
    _run("pack_weights", 1, 0, 4 * (W.numel() + out.numel()),
         lambda: _K.pack_weights(W, out, 2, N, Cc, KK * N, Cp, KK, _split_of(out, KK * N)))
    return out


class PackPlan:
    """All weight re-packs of one network as ONE launch (mdgan_pack_weights_multi), issued after its Adam step."""

    def __init__(self, device: torch.device):
        self.device = device
        self.jobs = []      # rows of 11 ints
        self.keep = []      # the tensors the recorded pointers refer to
        self.blocks = 0
        self.nbytes = 0.0
        self.table: Optional[torch.Tensor] = None

    def _add(self, W, out, mode, N, Cc, Np, Cp, KK, total, split):
        if self.table is not None:
            raise _lib.MdganLibraryError("PackPlan is already finalized")
        _ptr(W), _ptr(out)  # CUDA + contiguous checks
        self.jobs.append([W.data_ptr(), out.data_ptr(), mode, N, Cc, Np, Cp, KK, total, self.blocks, split])
        self.keep += [W, out]
        # modes 0 / 1 are tiled (one block per output row and 32 channels), the others element-wise
        self.blocks += (Np * (Cp // 32)) if mode <= 1 else (total + 255) // 256
        self.nbytes += 4.0 * (W.numel() + out.numel())

    def add_down(self, W, out):
        N, Cc = W.shape[0], W.shape[1]
        Np, Cp = n_pad_for(N), _pad(Cc, 32)
        self._add(W, out, 0, N, Cc, Np, Cp, 16, Np * 16 * Cp, _split_of(out, Np))

    def add_up(self, W, out):
        Cc, N = W.shape[0], W.shape[1]
        Np, Cp = n_pad_for(N), _pad(Cc, 32)
        self._add(W, out, 1, N, Cc, Np, Cp, 16, 4 * Np * 4 * Cp, _split_of(out, 4 * Np))

    def add_dense(self, W, out):
        Cc, N, KK = W.shape[0], W.shape[1], W.shape[2] * W.shape[3]
        Cp = _pad(Cc, 32)
        self._add(W, out, 2, N, Cc, KK * N, Cp, KK, KK * N * Cp, _split_of(out, KK * N))

    def add_head(self, w, wt):
        Cc, HW = w.shape[1], w.shape[2] * w.shape[3]
        self._add(w, wt, 3, HW, Cc, HW, Cc, 1, HW * Cc, 0)

    def finalize(self) -> "PackPlan":
        words = _lib.load().mdgan_pack_job_words()
        if any(len(j) != words for j in self.jobs):
            raise _lib.MdganLibraryError("pack job record size mismatch with the library")
        self.table = torch.tensor(self.jobs, dtype=torch.int64).to(self.device)
        return self

    def run(self) -> None:
        if self.table is None:
            self.finalize()
        _run("pack_weights", 1, 0, self.nbytes,
             lambda: _K.pack_weights_multi(self.table, len(self.jobs), self.blocks))


# ----------------------------------------------------------------------------- tensor-core GEMMs
def conv_gemm(src: torch.Tensor, wpacked: torch.Tensor, mode: int, N: int, out: torch.Tensor,
              grid: Tuple[int, int, int], src_hw: Tuple[int, int], bias: Optional[torch.Tensor] = None,
              out_nchw: bool = False, act_tanh: bool = False, round_tf32: bool = False, accumulate: bool = False,
              precision: int = TF32, force_bn: int = 0, gate: Optional[torch.Tensor] = None, gate_act: int = ACT_NONE,
              gate_slope: float = 0.0, bn_partial: Optional[torch.Tensor] = None, bnb=None) -> torch.Tensor:
    """grid = (n_img, Hg, Wg): the low-resolution row grid; src is NHWC [n_img, Hs, Ws, C].  With precision TF32X3
    `wpacked` must hold the hi and lo matrices (pack_*(..., precision=TF32X3)).  gate (NHWC, shaped like out): the
    result is multiplied by act'(gate) -- the (Leaky)ReLU backward fused into the data gradient that feeds it.
    bn_partial ([phases * row tiles, 2, N_pad], see conv_stats_shape): the GEMM epilogue also reduces the BatchNorm
    statistics of its output; finish with bn_finalize + bn_apply instead of bn_forward.
    bnb = (z, stats, act, slope, groups) with bn_partial (data-gradient GEMMs): `out` is the gradient of the output of
    BatchNorm(z)+act; the epilogue stores dy = da * act'(.) and reduces the BatchNorm-backward sums; finish with
    bn_bwd_finalize + bn_bwd_apply_dy instead of bn_backward."""
    bz, bst, bact, bslope, bgroups = bnb if bnb is not None else (None, None, 0, 0.0, 1)
    n_img, Hg, Wg = grid
    Hs, Ws = src_hw
    Cc = src.shape[-1]
    phases = 4 if mode == MODE_UP else 1
    N_pad = wpacked.shape[0] // phases // (2 if precision == TF32X3 else 1)
    taps = {MODE_DOWN: 16, MODE_UP: 16, MODE_DENSE: 1}[mode]  # UP: 4 phases x 4 taps per low-res position
    flops = 2.0 * n_img * Hg * Wg * taps * Cc * N
    nbytes = 4.0 * (src.numel() + out.numel() + wpacked.numel())
    _run(("conv_down", "conv_up", "conv_dense")[mode], 1, flops, nbytes,
         lambda: _K.conv_gemm(src, wpacked, out, bias, n_img, Hg, Wg, Hs, Ws, Cc,
                                             mode, N, N_pad, int(out_nchw), int(act_tanh), int(round_tf32),
                                             int(accumulate), precision, force_bn, gate, gate_act,
                                             float(gate_slope), bn_partial, bz, bst, bact, float(bslope), bgroups))
    return out


def conv_rows_per_tile(Hg: int, Wg: int, precision: int) -> int:
    """GEMM rows one CTA of conv_gemm owns (0: the fused BatchNorm statistics are unavailable in this mode)."""
    return _lib.load().mdgan_conv_rows_per_tile(Hg, Wg, precision)


def conv_stats_plan(grid: Tuple[int, int, int], mode: int, groups: int, precision: int, n_cols: int = 0):
    """(row_tiles, tiles_per_group, phases) of the fused statistics of conv_gemm over `grid` = (n_img, Hg, Wg) with
    n_cols output channels holding `groups` equally sized BatchNorm passes, or None when a CTA's rows would straddle two passes / the mode has no
    fused statistics (callers then use bn_forward)."""
    n_img, Hg, Wg = grid
    rpt = conv_rows_per_tile(Hg, Wg, precision)
    M = n_img * Hg * Wg
    if rpt <= 0 or M % groups != 0 or (M // groups) % rpt != 0:
        return None
    phases = _lib.load().mdgan_conv_stat_phases(mode, n_pad_for(n_cols) if n_cols else 0, Hg, Wg, precision)
    return M // rpt, M // groups // rpt, phases


def wgrad_splits(n_img: int, Hl: int, Wl: int, C1: int, C2: int, mode: int) -> int:
    return _lib.load().mdgan_wgrad_splits(n_img, Hl, Wl, C1, C2, mode)


def wgrad_gemm(lo: torch.Tensor, hi: torch.Tensor, partial: torch.Tensor, grid: Tuple[int, int, int], mode: int,
               splits: int, precision: int = TF32) -> torch.Tensor:
    n_img, Hl, Wl = grid
    taps = 16 if mode == MODE_DOWN else 1
    flops = 2.0 * n_img * Hl * Wl * taps * lo.shape[-1] * hi.shape[-1]
    nbytes = 4.0 * (lo.numel() + hi.numel() + taps * lo.shape[-1] * hi.shape[-1])
    _run("wgrad_gemm", 1, flops, nbytes,
         lambda: _K.wgrad_gemm(lo, hi, partial, n_img, Hl, Wl, lo.shape[-1],
                                              hi.shape[-1], mode, splits, precision))
    return partial


def wgrad_unpack(partial: torch.Tensor, grad: torch.Tensor, mode: int, splits: int, C1: int, C1p: int, C2: int,
                 N: int = 0, KK: int = 0) -> torch.Tensor:
    _run("wgrad_unpack", 1, 0, 4.0 * grad.numel() * (splits + 1),
         lambda: _K.wgrad_unpack(partial, grad, mode, splits, C1, C1p, C2, N, KK))
    return grad


def reduce_slices(partial: torch.Tensor, out: torch.Tensor, slices: int) -> torch.Tensor:
    _run("reduce_slices", 1, 0, 4.0 * out.numel() * (slices + 1),
         lambda: _K.reduce_slices(partial, out, slices, out.numel()))
    return out


# ----------------------------------------------------------------------------- thin (image-side) layers
def thin_down(img: torch.Tensor, W: torch.Tensor, out: torch.Tensor, act: int = ACT_NONE, slope: float = 0.0,
              round_tf32: bool = False) -> torch.Tensor:
    n, ci, Hi, Wi = img.shape
    _run("thin_down", 1, 2.0 * out.numel() * 16 * ci, 4.0 * (img.numel() + out.numel() + W.numel()),
         lambda: _K.thin_down(img, W, out, n, ci, Hi, Wi, W.shape[0], act, slope,
                                             int(round_tf32)))
    return out


def thin_up(src: torch.Tensor, W: torch.Tensor, out: torch.Tensor, act_tanh: bool = False,
            accumulate: bool = False) -> torch.Tensor:
    """src NHWC [n, H, W, C], W [C, N, 4, 4] (N in {1, 3}) -> out NCHW [n, N, 2H, 2W]; optional tanh / += ."""
    n, H, Wd, Cc = src.shape
    N = W.shape[1]
    _run("thin_up", 1, 2.0 * n * H * Wd * 16 * Cc * N, 4.0 * (src.numel() + out.numel() * (2 if accumulate else 1)),
         lambda: _K.thin_up(src, W, out, n, H, Wd, Cc, N, int(act_tanh),
                                           int(accumulate)))
    return out


def thin_wgrad_slices(n_img: int, Hl: int, Wl: int) -> int:
    return _lib.load().mdgan_thin_wgrad_slices(n_img, Hl, Wl)


def thin_wgrad(feat: torch.Tensor, img: torch.Tensor, partial: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    n, ci, Hi, Wi = img.shape
    Hl, Wl, C1 = Hi // 2, Wi // 2, feat.shape[-1]
    _run("thin_wgrad", 1, 2.0 * feat.numel() * 16 * ci, 4.0 * (feat.numel() + img.numel() + grad.numel()),
         lambda: _K.thin_wgrad(feat, img, partial, n, ci, Hl, Wl, C1))
    return reduce_slices(partial, grad, thin_wgrad_slices(n, Hl, Wl))


# ----------------------------------------------------------------------------- BatchNorm + activation
def bn_workspace_floats(G: int, Pg: int, Cc: int) -> int:
    return _lib.load().mdgan_bn_workspace_floats(G, Pg, Cc)


def bn_counters(device) -> torch.Tensor:
    """The 32 zero-initialised slab counters mdgan_bn_forward / mdgan_bn_backward need (one tensor per net is enough:
    launches on one stream use it one after the other and leave it zero)."""
    return torch.zeros(32, dtype=torch.int32, device=device)


def bn_forward(x, out, gamma, beta, running_mean, running_var, nbt, stats, workspace, counters, G, Pg, Cc, act, slope,
               round_tf32=False, eps=1e-5, momentum=0.1):
    # algorithmic bytes: read x for the statistics, read x + write out for the normalisation
    _run("bn_forward", 2, 0, 4.0 * 3 * G * Pg * Cc,
         lambda: _K.bn_forward(x, out, gamma, beta, running_mean,
                                              running_var, nbt, stats, workspace, counters,
                                              G, Pg, Cc, eps, momentum, act, slope, int(round_tf32)))
    return out


def bn_finalize(partial, plan, col_stride, fold, gamma, beta, running_mean, running_var, nbt, stats, G, Pg, Cc,
                eps=1e-5, momentum=0.1):
    """Statistics reduced by the producing GEMM (conv_gemm bn_partial; plan from conv_stats_plan) -> stats / running."""
    row_tiles, tpg, phases = plan
    _run("bn_forward", 1, 0, 4.0 * 2 * phases * row_tiles * col_stride,
         lambda: _K.bn_finalize(partial, phases, row_tiles, tpg, col_stride, fold, gamma, beta, running_mean, running_var,
                                nbt, stats, G, Pg, Cc, eps, momentum))


def bn_apply(x, stats, out, G, Pg, Cc, act, slope, round_tf32=False):
    _run("bn_forward", 1, 0, 4.0 * 2 * G * Pg * Cc,
         lambda: _K.bn_apply(x, stats, out, G, Pg, Cc, act, slope, int(round_tf32)))
    return out


def bn_backward(da, x, stats, dx, dgamma, dbeta, sums, workspace, counters, G, Pg, Cc, act, slope, round_tf32=False):
    # algorithmic bytes: read da + x for the sums, read da + x and write dx for the gradient
    _run("bn_backward", 2, 0, 4.0 * 5 * G * Pg * Cc,
         lambda: _K.bn_backward(da, x, stats, dx, dgamma, dbeta,
                                               sums, workspace, counters, G, Pg, Cc, act, slope,
                                               int(round_tf32)))
    return dx


def bn_bwd_finalize(partial, plan, col_stride, sums, dgamma, dbeta, G, Cc):
    row_tiles, tpg, phases = plan
    _run("bn_backward", 1, 0, 4.0 * 2 * phases * row_tiles * col_stride,
         lambda: _K.bn_bwd_finalize(partial, phases, row_tiles, tpg, col_stride, sums, dgamma, dbeta, G, Cc))


def bn_bwd_apply_dy(dy, x, stats, sums, dx, G, Pg, Cc, round_tf32=False):
    _run("bn_backward", 1, 0, 4.0 * 3 * G * Pg * Cc,
         lambda: _K.bn_bwd_apply_dy(dy, x, stats, sums, dx, G, Pg, Cc, int(round_tf32)))
    return dx


def act_backward(da, a, dz, act, slope, round_tf32=False):
    _run("act_backward", 1, 0, 4.0 * 3 * a.numel(),
         lambda: _K.act_backward(da, a, dz, a.numel(), act, slope, int(round_tf32)))
    return dz


def tanh_backward(s, x, out, scale: float):
    _run("tanh_backward", 1, 0, 4.0 * 3 * x.numel(),
         lambda: _K.tanh_backward(s, x, out, x.numel(), scale))
    return out


def tanh_backward_slices(F, x, out, k: int, N: int, scale: float):
    """F [N, b*C*H*W] feedback slices (worker order), x / out [k*b, C, H, W]: group sum per generated batch + tanh'."""
    n_per = x.numel() // k
    _run("tanh_backward", 1, 0, 4.0 * (F.numel() + 2 * x.numel()),
         lambda: _K.tanh_backward_slices(F, x, out, n_per, k, N, scale))
    return out


# ----------------------------------------------------------------------------- peer-memory exchange (NVLink)
def peer_signal(flag_addrs: torch.Tensor, n: int, epoch: torch.Tensor, advance: bool):
    _run("peer_signal", 1, 0, 4.0 * n,
         lambda: _K.peer_signal(flag_addrs, n, epoch, int(advance)))


def peer_timeout_ms() -> int:
    """MDGAN_PEER_TIMEOUT_S (default 600): wall time a flag wait may last before the kernel traps (dead peer)."""
    return max(1, int(float(os.environ.get("MDGAN_PEER_TIMEOUT_S", "600")) * 1000))


def peer_wait(flags: torch.Tensor, n: int, epoch: torch.Tensor, advance: bool, err: torch.Tensor):
    _run("peer_wait", 1, 0, 4.0 * n,
         lambda: _K.peer_wait(flags, n, epoch, int(advance), err, peer_timeout_ms()))


def peer_push(src: torch.Tensor, dst_addrs: torch.Tensor, n_dst: int):
    _run("peer_push", 1, 0, 4.0 * src.numel() * (1 + n_dst),
         lambda: _K.peer_push(src, dst_addrs, n_dst, src.numel()))


def peer_push_multicast(src: torch.Tensor, mc_dst: torch.Tensor):
    """mc_dst: a tensor view of the NVSwitch multicast address of the symmetric buffer (exchange.PeerExchange)."""
    _run("peer_push", 1, 0, 4.0 * src.numel() * 2,
         lambda: _K.peer_push_multicast(src, mc_dst, src.numel()))


# ----------------------------------------------------------------------------- head / loss / optimiser
def head_pack(w, wt):
    """w PyTorch [1, C, k, k] -> wt [k*k, C] (the NHWC order of the activations)."""
    Cc, HW = w.shape[1], w.shape[2] * w.shape[3]
    _run("head_pack", 1, 0, 8.0 * w.numel(), lambda: _K.head_pack(w, wt, HW, Cc))
    return wt


def head_forward(a, w, label, prob, loss_terms, dlogit, loss, counter, G, b, HW, Cc):
    """counter: one zero-initialised int32 device scalar (block counter of the fused loss reduction)."""
    _run("head_forward", 1, 2.0 * G * b * HW * Cc, 4.0 * (G * b * HW * Cc + HW * Cc),
         lambda: _K.head_forward(a, w, label, prob, loss_terms,
                                                dlogit, loss, counter, G, b, HW, Cc))


def head_backward(a, w, dlogit, da, dw, n_total, HW, Cc):
    _run("head_backward", 1, 4.0 * n_total * HW * Cc, 4.0 * (2 * n_total * HW * Cc + 2 * HW * Cc),
         lambda: _K.head_backward(a, w, dlogit, da, dw, n_total, HW, Cc))


def adam_step(p, g, m, v, step_count, lr, beta1, beta2, eps=1e-8):
    """step_count: int32 [2] on the device = (steps taken, zero-initialised block counter)."""
    if step_count.numel() < 2:
        raise _lib.MdganLibraryError("adam_step: step_count must hold 2 int32 (step, block counter)")
    _run("adam_step", 1, 0, 28.0 * p.numel(),
         lambda: _K.adam_step(p, g, m, v, p.numel(), step_count, lr, beta1,
                                             beta2, eps))


def pad_rows(x, out, round_tf32=False):
    rows, cin = x.shape
    _run("pad_rows", 1, 0, 4.0 * (x.numel() + out.numel()),
         lambda: _K.pad_rows(x, out, rows, cin, out.shape[1], int(round_tf32)))
    return out


def sum_slices(x, out, count: int, stride: int):
    _run("sum_slices", 1, 0, 4.0 * out.numel() * (count + 1),
         lambda: _K.sum_slices(x, out, out.numel(), count, stride))
    return out


# ----------------------------------------------------------------------------- Linear layers (the reference's MLP plugin)
def _sgemm(name, A, B, out, M, N, K, a_rs, a_cs, b_rs, b_cs, bias=None, act=ACT_NONE, slope=0.0, mask=None, mask_scale=1.0,
           gate=None, gate_slope=1.0, accumulate=False):
    if mask is not None and (mask.dtype != torch.uint8 or mask.numel() != M * N):
        raise _lib.MdganLibraryError("sgemm: the dropout mask must be uint8 [M, N]")
    if out.numel() != M * N or (gate is not None and gate.numel() != M * N) or (bias is not None and bias.numel() != N):
        raise _lib.MdganLibraryError("sgemm: operand shapes do not match M, N")
    _run(name, 1, 2.0 * M * N * K, 4.0 * (M * K + K * N + M * N),
         lambda: _K.sgemm(A, B, out, M, N, K, a_rs, a_cs, b_rs, b_cs, bias, act, slope, mask, mask_scale, gate,
                          gate_slope, int(accumulate)))
    return out


def linear_forward(x, W, bias, out, act=ACT_NONE, slope=0.0, mask=None, mask_scale=1.0):
    """out [M, N] = dropout(act(x [M, K] @ W [N, K]^T + bias)); mask: uint8 keep mask [M, N] or None."""
    M, K = x.shape
    N = W.shape[0]
    return _sgemm("linear_forward", x, W, out, M, N, K, K, 1, 1, K, bias=bias, act=act, slope=slope, mask=mask,
                  mask_scale=mask_scale)


def linear_dgrad(dy, W, out, gate=None, gate_slope=1.0, mask=None, mask_scale=1.0, accumulate=False):
    """out [M, K] (+)= (dy [M, N] @ W [N, K]) through the dropout mask and LeakyReLU gate of the layer that produced the
    [M, K] activations (gate = that layer's output)."""
    M, N = dy.shape
    K = W.shape[1]
    return _sgemm("linear_dgrad", dy, W, out, M, K, N, N, 1, K, 1, mask=mask, mask_scale=mask_scale, gate=gate,
                  gate_slope=gate_slope, accumulate=accumulate)


def linear_wgrad(dy, x, dW):
    """dW [N, K] = dy [M, N]^T @ x [M, K] (PyTorch Linear.weight layout)."""
    M, N = dy.shape
    K = x.shape[1]
    return _sgemm("linear_wgrad", dy, x, dW, N, K, M, 1, N, K, 1)


def col_sum(x, out):
    M, N = x.shape
    _run("col_sum", 1, 0, 4.0 * (x.numel() + N), lambda: _K.col_sum(x, out, M, N))
    return out


def linear_head_forward(a, w, bias, label, prob, loss_terms, dlogit, loss, counter, G, b):
    L = a.shape[1]
    _run("head_forward", 1, 2.0 * G * b * L, 4.0 * (G * b * L + L),
         lambda: _K.linear_head_forward(a, w, bias, label, prob, loss_terms, dlogit, loss, counter, G, b, L))


def linear_head_backward(a, w, dlogit, da, dw, dbias, mask=None, mask_scale=1.0, gate_slope=1.0):
    n, L = a.shape
    _run("head_backward", 1, 4.0 * n * L, 4.0 * (2 * n * L + 2 * L),
         lambda: _K.linear_head_backward(a, w, dlogit, mask, mask_scale, gate_slope, da, dw, dbias, n, L))
