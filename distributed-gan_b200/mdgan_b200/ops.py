"""Tensor-level wrappers over the C-ABI launchers (one function per exported kernel family).

Activations are NHWC fp32 CUDA tensors, parameters keep their PyTorch layout.  Every wrapper launches on
torch's current CUDA stream, so the calls can be captured into a CUDA graph.  No fallback: a non-CUDA tensor
or a failed launch raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

import os

MODE_DOWN, MODE_UP, MODE_DENSE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
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
        _lib.check(call(), name)
        return
    with _observer(name, n_kernels, flops, nbytes):
        _lib.check(call(), name)


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
         lambda: _lib.load().mdgan_pack_weights(_ptr(W), _ptr(out), 0, N, Cc, Np, Cp, 16, _split_of(out, Np), _stream()))
    return out


def pack_up(W: torch.Tensor, out: Optional[torch.Tensor] = None, precision: int = TF32) -> torch.Tensor:
    """W [C, N, 4, 4] -> [4 phases * N_pad, 4*C_pad]; convT fwd / conv dgrad operand."""
    Cc, N = W.shape[0], W.shape[1]
    Np, Cp = n_pad_for(N), _pad(Cc, 32)
    if out is None:
        out = torch.empty(packed_shape(MODE_UP, N, Cc, precision=precision), device=W.device, dtype=torch.float32)
    _run("pack_weights", 1, 0, 4 * (W.numel() + out.numel()),
         lambda: _lib.load().mdgan_pack_weights(_ptr(W), _ptr(out), 1, N, Cc, Np, Cp, 16, _split_of(out, 4 * Np),
                                                _stream()))
    return out


def pack_dense(W: torch.Tensor, out: Optional[torch.Tensor] = None, precision: int = TF32) -> torch.Tensor:
    """W [C, N, k, k] (ConvTranspose2d on a 1x1 input) -> [k*k*N, C_pad]."""
    Cc, N, KK = W.shape[0], W.shape[1], W.shape[2] * W.shape[3]
    Cp = _pad(Cc, 32)
    if out is None:
        out = torch.empty(packed_shape(MODE_DENSE, N, Cc, KK, precision), device=W.device, dtype=torch.float32)
    _run("pack_weights", 1, 0, 4 * (W.numel() + out.numel()),
         lambda: _lib.load().mdgan_pack_weights(_ptr(W), _ptr(out), 2, N, Cc, KK * N, Cp, KK, _split_of(out, KK * N),
                                                _stream()))
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
             lambda: _lib.load().mdgan_pack_weights_multi(_ptr(self.table), len(self.jobs), self.blocks, _stream()))


# ----------------------------------------------------------------------------- tensor-core GEMMs
def conv_gemm(src: torch.Tensor, wpacked: torch.Tensor, mode: int, N: int, out: torch.Tensor,
              grid: Tuple[int, int, int], src_hw: Tuple[int, int], bias: Optional[torch.Tensor] = None,
              out_nchw: bool = False, act_tanh: bool = False, round_tf32: bool = False, accumulate: bool = False,
              precision: int = TF32, force_bn: int = 0, gate: Optional[torch.Tensor] = None, gate_act: int = ACT_NONE,
              gate_slope: float = 0.0) -> torch.Tensor:
    """grid = (n_img, Hg, Wg): the low-resolution row grid; src is NHWC [n_img, Hs, Ws, C].  With precision TF32X3
    `wpacked` must hold the hi and lo matrices (pack_*(..., precision=TF32X3)).  gate (NHWC, shaped like out): the
    result is multiplied by act'(gate) -- the (Leaky)ReLU backward fused into the data gradient that feeds it."""
    n_img, Hg, Wg = grid
    Hs, Ws = src_hw
    Cc = src.shape[-1]
    phases = 4 if mode == MODE_UP else 1
    N_pad = wpacked.shape[0] // phases // (2 if precision == TF32X3 else 1)
    taps = {MODE_DOWN: 16, MODE_UP: 16, MODE_DENSE: 1}[mode]  # UP: 4 phases x 4 taps per low-res position
    flops = 2.0 * n_img * Hg * Wg * taps * Cc * N
    nbytes = 4.0 * (src.numel() + out.numel() + wpacked.numel())
    _run(("conv_down", "conv_up", "conv_dense")[mode], 1, flops, nbytes,
         lambda: _lib.load().mdgan_conv_gemm(_ptr(src), _ptr(wpacked), _ptr(out), _ptr(bias), n_img, Hg, Wg, Hs, Ws, Cc,
                                             mode, N, N_pad, int(out_nchw), int(act_tanh), int(round_tf32),
                                             int(accumulate), precision, force_bn, _ptr(gate), gate_act,
                                             float(gate_slope), _stream()))
    return out


def wgrad_splits(n_img: int, Hl: int, Wl: int, C1: int, C2: int, mode: int) -> int:
    return _lib.load().mdgan_wgrad_splits(n_img, Hl, Wl, C1, C2, mode)


def wgrad_gemm(lo: torch.Tensor, hi: torch.Tensor, partial: torch.Tensor, grid: Tuple[int, int, int], mode: int,
               splits: int, precision: int = TF32) -> torch.Tensor:
    n_img, Hl, Wl = grid
    taps = 16 if mode == MODE_DOWN else 1
    flops = 2.0 * n_img * Hl * Wl * taps * lo.shape[-1] * hi.shape[-1]
    nbytes = 4.0 * (lo.numel() + hi.numel() + taps * lo.shape[-1] * hi.shape[-1])
    _run("wgrad_gemm", 1, flops, nbytes,
         lambda: _lib.load().mdgan_wgrad_gemm(_ptr(lo), _ptr(hi), _ptr(partial), n_img, Hl, Wl, lo.shape[-1],
                                              hi.shape[-1], mode, splits, precision, _stream()))
    return partial


def wgrad_unpack(partial: torch.Tensor, grad: torch.Tensor, mode: int, splits: int, C1: int, C1p: int, C2: int,
                 N: int = 0, KK: int = 0) -> torch.Tensor:
    _run("wgrad_unpack", 1, 0, 4.0 * grad.numel() * (splits + 1),
         lambda: _lib.load().mdgan_wgrad_unpack(_ptr(partial), _ptr(grad), mode, splits, C1, C1p, C2, N, KK, _stream()))
    return grad


def reduce_slices(partial: torch.Tensor, out: torch.Tensor, slices: int) -> torch.Tensor:
    _run("reduce_slices", 1, 0, 4.0 * out.numel() * (slices + 1),
         lambda: _lib.load().mdgan_reduce_slices(_ptr(partial), _ptr(out), slices, out.numel(), _stream()))
    return out


# ----------------------------------------------------------------------------- thin (image-side) layers
def thin_down(img: torch.Tensor, W: torch.Tensor, out: torch.Tensor, act: int = ACT_NONE, slope: float = 0.0,
              round_tf32: bool = False) -> torch.Tensor:
    n, ci, Hi, Wi = img.shape
    _run("thin_down", 1, 2.0 * out.numel() * 16 * ci, 4.0 * (img.numel() + out.numel() + W.numel()),
         lambda: _lib.load().mdgan_thin_down(_ptr(img), _ptr(W), _ptr(out), n, ci, Hi, Wi, W.shape[0], act, slope,
                                             int(round_tf32), _stream()))
    return out


def thin_up(src: torch.Tensor, W: torch.Tensor, out: torch.Tensor, act_tanh: bool = False,
            accumulate: bool = False) -> torch.Tensor:
    """src NHWC [n, H, W, C], W [C, N, 4, 4] (N in {1, 3}) -> out NCHW [n, N, 2H, 2W]; optional tanh / += ."""
    n, H, Wd, Cc = src.shape
    N = W.shape[1]
    _run("thin_up", 1, 2.0 * n * H * Wd * 16 * Cc * N, 4.0 * (src.numel() + out.numel() * (2 if accumulate else 1)),
         lambda: _lib.load().mdgan_thin_up(_ptr(src), _ptr(W), _ptr(out), n, H, Wd, Cc, N, int(act_tanh),
                                           int(accumulate), _stream()))
    return out


def thin_wgrad_slices(n_img: int, Hl: int, Wl: int) -> int:
    return _lib.load().mdgan_thin_wgrad_slices(n_img, Hl, Wl)


def thin_wgrad(feat: torch.Tensor, img: torch.Tensor, partial: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    n, ci, Hi, Wi = img.shape
    Hl, Wl, C1 = Hi // 2, Wi // 2, feat.shape[-1]
    _run("thin_wgrad", 1, 2.0 * feat.numel() * 16 * ci, 4.0 * (feat.numel() + img.numel() + grad.numel()),
         lambda: _lib.load().mdgan_thin_wgrad(_ptr(feat), _ptr(img), _ptr(partial), n, ci, Hl, Wl, C1, _stream()))
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
         lambda: _lib.load().mdgan_bn_forward(_ptr(x), _ptr(out), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                              _ptr(running_var), _ptr(nbt), _ptr(stats), _ptr(workspace), _ptr(counters),
                                              G, Pg, Cc, eps, momentum, act, slope, int(round_tf32), _stream()))
    return out


def bn_backward(da, x, stats, dx, dgamma, dbeta, sums, workspace, counters, G, Pg, Cc, act, slope, round_tf32=False):
    # algorithmic bytes: read da + x for the sums, read da + x and write dx for the gradient
    _run("bn_backward", 2, 0, 4.0 * 5 * G * Pg * Cc,
         lambda: _lib.load().mdgan_bn_backward(_ptr(da), _ptr(x), _ptr(stats), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                                               _ptr(sums), _ptr(workspace), _ptr(counters), G, Pg, Cc, act, slope,
                                               int(round_tf32), _stream()))
    return dx


def act_backward(da, a, dz, act, slope, round_tf32=False):
    _run("act_backward", 1, 0, 4.0 * 3 * a.numel(),
         lambda: _lib.load().mdgan_act_backward(_ptr(da), _ptr(a), _ptr(dz), a.numel(), act, slope, int(round_tf32),
                                                _stream()))
    return dz


def tanh_backward(s, x, out, scale: float):
    _run("tanh_backward", 1, 0, 4.0 * 3 * x.numel(),
         lambda: _lib.load().mdgan_tanh_backward(_ptr(s), _ptr(x), _ptr(out), x.numel(), scale, _stream()))
    return out


def tanh_backward_slices(F, x, out, k: int, N: int, scale: float):
    """F [N, b*C*H*W] feedback slices (worker order), x / out [k*b, C, H, W]: group sum per generated batch + tanh'."""
    n_per = x.numel() // k
    _run("tanh_backward", 1, 0, 4.0 * (F.numel() + 2 * x.numel()),
         lambda: _lib.load().mdgan_tanh_backward_slices(_ptr(F), _ptr(x), _ptr(out), n_per, k, N, scale, _stream()))
    return out


# ----------------------------------------------------------------------------- peer-memory exchange (NVLink)
def peer_signal(flag_addrs: torch.Tensor, n: int, epoch: torch.Tensor, advance: bool):
    _run("peer_signal", 1, 0, 4.0 * n,
         lambda: _lib.load().mdgan_peer_signal(_ptr(flag_addrs), n, _ptr(epoch), int(advance), _stream()))


def peer_timeout_ms() -> int:
    """MDGAN_PEER_TIMEOUT_S (default 600): wall time a flag wait may last before the kernel traps (dead peer)."""
    return max(1, int(float(os.environ.get("MDGAN_PEER_TIMEOUT_S", "600")) * 1000))


def peer_wait(flags: torch.Tensor, n: int, epoch: torch.Tensor, advance: bool, err: torch.Tensor):
    _run("peer_wait", 1, 0, 4.0 * n,
         lambda: _lib.load().mdgan_peer_wait(_ptr(flags), n, _ptr(epoch), int(advance), _ptr(err),
                                             C.c_longlong(peer_timeout_ms()), _stream()))


def peer_push(src: torch.Tensor, dst_addrs: torch.Tensor, n_dst: int):
    _run("peer_push", 1, 0, 4.0 * src.numel() * (1 + n_dst),
         lambda: _lib.load().mdgan_peer_push(_ptr(src), _ptr(dst_addrs), n_dst, src.numel(), _stream()))


# ----------------------------------------------------------------------------- head / loss / optimiser
def head_pack(w, wt):
    """w PyTorch [1, C, k, k] -> wt [k*k, C] (the NHWC order of the activations)."""
    Cc, HW = w.shape[1], w.shape[2] * w.shape[3]
    _run("head_pack", 1, 0, 8.0 * w.numel(), lambda: _lib.load().mdgan_head_pack(_ptr(w), _ptr(wt), HW, Cc, _stream()))
    return wt


def head_forward(a, w, label, prob, loss_terms, dlogit, loss, counter, G, b, HW, Cc):
    """counter: one zero-initialised int32 device scalar (block counter of the fused loss reduction)."""
    _run("head_forward", 1, 2.0 * G * b * HW * Cc, 4.0 * (G * b * HW * Cc + HW * Cc),
         lambda: _lib.load().mdgan_head_forward(_ptr(a), _ptr(w), _ptr(label), _ptr(prob), _ptr(loss_terms),
                                                _ptr(dlogit), _ptr(loss), _ptr(counter), G, b, HW, Cc, _stream()))


def head_backward(a, w, dlogit, da, dw, n_total, HW, Cc):
    _run("head_backward", 1, 4.0 * n_total * HW * Cc, 4.0 * (2 * n_total * HW * Cc + 2 * HW * Cc),
         lambda: _lib.load().mdgan_head_backward(_ptr(a), _ptr(w), _ptr(dlogit), _ptr(da), _ptr(dw), n_total, HW, Cc,
                                                 _stream()))


def adam_step(p, g, m, v, step_count, lr, beta1, beta2, eps=1e-8):
    """step_count: int32 [2] on the device = (steps taken, zero-initialised block counter)."""
    if step_count.numel() < 2:
        raise _lib.MdganLibraryError("adam_step: step_count must hold 2 int32 (step, block counter)")
    _run("adam_step", 1, 0, 28.0 * p.numel(),
         lambda: _lib.load().mdgan_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(step_count), lr, beta1,
                                             beta2, eps, _stream()))


def pad_rows(x, out, round_tf32=False):
    rows, cin = x.shape
    _run("pad_rows", 1, 0, 4.0 * (x.numel() + out.numel()),
         lambda: _lib.load().mdgan_pad_rows(_ptr(x), _ptr(out), rows, cin, out.shape[1], int(round_tf32), _stream()))
    return out


def sum_slices(x, out, count: int, stride: int):
    _run("sum_slices", 1, 0, 4.0 * out.numel() * (count + 1),
         lambda: _lib.load().mdgan_sum_slices(_ptr(x), _ptr(out), out.numel(), count, stride, _stream()))
    return out
