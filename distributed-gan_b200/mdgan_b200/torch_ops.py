"""torch custom-op layer over the C-ABI: every launcher of include/mdgan_b200.h that works on device memory is
registered as `torch.ops.mdgan_b200.<name>` (CUDA dispatch key only -- a CPU tensor has no kernel to dispatch to, there
is no fallback).  The op takes tensors where the C function takes device pointers, passes scalars through, launches on
torch's current CUDA stream and raises MdganLibraryError on a non-zero return code.  Mutated arguments are annotated
(`Tensor(a!)`), so the ops are visible to the dispatcher, profiler and graph capture like any other torch op; the
product (mdgan_b200/ops.py) calls the kernels only through this layer.

Spec mini-language (one string per op): comma-separated `<kind> <name>` in the C argument order without the trailing
stream; kinds: T tensor (read), T! tensor (written), T? optional tensor (read), T?! optional tensor (written),
int, float.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Tuple

import torch

from . import _lib

NAMESPACE = "mdgan_b200"

SPECS: Dict[str, str] = {
    "conv_gemm": "T src, T wpacked, T! dst, T? bias, int n_img, int Hg, int Wg, int Hs, int Ws, int C, int mode, int N, "
                 "int N_pad, int out_nchw, int act, int round_tf32, int accumulate, int precision, int force_bn, "
                 "T? gate, int gate_act, float gate_slope, T?! bn_partial, T? bnb_z, T? bnb_stats, int bnb_act, "
                 "float bnb_slope, int bnb_groups",
    "wgrad_gemm": "T lo, T hi, T! partial, int n_img, int Hl, int Wl, int C1, int C2, int mode, int splits, int precision",
    "pack_weights": "T W, T! out, int mode, int N, int C, int N_pad, int C_pad, int KK, int split",
    "wgrad_unpack": "T partial, T! grad, int mode, int splits, int C1, int C1p, int C2, int N, int KK",
    "pack_weights_multi": "T! jobs, int n_jobs, int total_blocks",   # writes through the pointers stored in `jobs`
    "reduce_slices": "T partial, T! out, int slices, int n",
    "bn_forward": "T x, T! out, T gamma, T beta, T?! running_mean, T?! running_var, T?! nbt, T! stats, T! workspace, "
                  "T! counters, int G, int Pg, int C, float eps, float momentum, int act, float slope, int round_tf32",
    "bn_finalize": "T partial, int phases, int row_tiles, int tiles_per_group, int col_stride, int fold, T gamma, T beta, "
                   "T?! running_mean, T?! running_var, T?! nbt, T! stats, int G, int Pg, int C, float eps, float momentum",
    "bn_apply": "T x, T stats, T! out, int G, int Pg, int C, int act, float slope, int round_tf32",
    "bn_bwd_finalize": "T partial, int phases, int row_tiles, int tiles_per_group, int col_stride, T! sums, T?! dgamma, "
                       "T?! dbeta, int G, int C",
    "bn_bwd_apply_dy": "T dy, T x, T stats, T sums, T! dx, int G, int Pg, int C, int round_tf32",
    "bn_backward": "T da, T x, T stats, T! dx, T?! dgamma, T?! dbeta, T! sums, T! workspace, T! counters, int G, int Pg, "
                   "int C, int act, float slope, int round_tf32",
    "act_backward": "T da, T a, T! dz, int n, int act, float slope, int round_tf32",
    "tanh_backward": "T s, T x, T! out, int n, float scale",
    "head_pack": "T w, T! wt, int HW, int C",
    "head_forward": "T a, T wt, T label, T! prob, T! loss_terms, T! dlogit, T! loss, T! counter, int G, int b, int HW, int C",
    "head_backward": "T a, T wt, T dlogit, T! da, T?! dw, int n_total, int HW, int C",
    "adam_step": "T! p, T g, T! m, T! v, int n, T! step_count, float lr, float beta1, float beta2, float eps",
    "pad_rows": "T x, T! out, int rows, int cols_in, int cols_out, int round_tf32",
    "sum_slices": "T x, T! out, int n, int count, int stride",
    "peer_signal": "T! flag_addrs, int n, T! epoch, int advance",    # writes to the peer-mapped addresses it holds
    "peer_wait": "T flags, int n, T! epoch, int advance, T! err, int timeout_ms",
    "peer_push": "T src, T! dst_addrs, int n_dst, int n",
    "peer_push_multicast": "T src, T! mc_dst, int n",
    "tanh_backward_slices": "T F, T x, T! out, int n_per_slot, int k, int N, float scale",
    "thin_down": "T img, T W, T! out, int n_img, int CI, int Hi, int Wi, int N, int act, float slope, int round_tf32",
    "thin_up": "T src, T W, T! out, int n_img, int H, int Wd, int C, int N, int act_tanh, int accumulate",
    "thin_wgrad": "T feat, T img, T! partial, int n_img, int CI, int Hl, int Wl, int C1",
    "sgemm": "T A, T B, T! C, int M, int N, int K, int a_rs, int a_cs, int b_rs, int b_cs, T? bias, int act, float slope, "
             "T? mask, float mask_scale, T? gate, float gate_slope, int accumulate",
    "col_sum": "T x, T! out, int M, int N",
    "linear_head_forward": "T a, T w, T? bias, T label, T! prob, T! loss_terms, T! dlogit, T! loss, T! counter, int G, "
                           "int b, int L",
    "linear_head_backward": "T a, T w, T dlogit, T? mask, float mask_scale, float gate_slope, T! da, T?! dw, T?! dbias, "
                            "int n_total, int L",
}

_lib_handle = None


def _parse(spec: str) -> List[Tuple[str, str]]:
    out = []
    for part in spec.split(","):
        kind, name = part.split()
        out.append((kind, name))
    return out


def _schema(name: str, args: List[Tuple[str, str]]) -> str:
    parts, alias = [], iter("abcdefghijklmnop")
    for kind, arg in args:
        if kind.startswith("T"):
            t = "Tensor"
            if "!" in kind:
                t += f"({next(alias)}!)"
            if "?" in kind:
                t += "?"
            parts.append(f"{t} {arg}")
        else:
            parts.append(f"{kind} {arg}")
    return f"{name}({', '.join(parts)}) -> ()"


def _make_impl(name: str, args: List[Tuple[str, str]]) -> Callable:
    cname = f"mdgan_{name}"
    ctypes_args = _lib.SIGNATURES[cname][1]
    if len(ctypes_args) != len(args) + 1:
        raise _lib.MdganLibraryError(f"torch op spec of {name} does not match the C signature")
    is_tensor = [kind.startswith("T") for kind, _ in args]

    def impl(*values):
        fn = getattr(_lib.load(), cname)
        call = []
        for v, tens in zip(values, is_tensor):
            if not tens:
                call.append(v)
            elif v is None:
                call.append(None)
            else:
                if not v.is_cuda:
                    raise _lib.MdganLibraryError("MD-GAN kernels need CUDA tensors (no CPU fallback)")
                if not v.is_contiguous():
                    raise _lib.MdganLibraryError("MD-GAN kernels need contiguous tensors")
                call.append(C.c_void_p(v.data_ptr()))
        _lib.check(fn(*call, C.c_void_p(torch.cuda.current_stream().cuda_stream)), cname)

    return impl


_registered = None


def register() -> "torch.library.Library":
    """Define and implement the ops once per process; returns the library handle (kept alive by this module)."""
    global _registered
    if _registered is not None:
        return _registered
    lib = torch.library.Library(NAMESPACE, "DEF")
    for name, spec in SPECS.items():
        args = _parse(spec)
        lib.define(_schema(name, args))
        lib.impl(name, _make_impl(name, args), "CUDA")
    _registered = lib
    return lib


register()
ns = getattr(torch.ops, NAMESPACE)
