"""ctypes binding of the C-ABI shared library `libmdgan_b200.so` (declared in include/mdgan_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("MDGAN_B200_LIB", _HERE / "libmdgan_b200.so"))

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ll = C.c_longlong

# name -> (restype, argtypes); every launcher returns int (0 ok, >0 cudaError_t, <0 own code)
SIGNATURES = {
    "mdgan_abi_version": (_i, []),
    "mdgan_check_device": (_i, []),
    "mdgan_conv_gemm": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _f, _p, _p, _p, _i, _f, _i, _p]),
    "mdgan_conv_rows_per_tile": (_i, [_i, _i, _i]),
    "mdgan_conv_stat_phases": (_i, [_i, _i, _i, _i, _i]),
    "mdgan_bn_finalize": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _p]),
    "mdgan_bn_apply": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "mdgan_bn_bwd_finalize": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p]),
    "mdgan_bn_bwd_apply_dy": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "mdgan_wgrad_splits": (_i, [_i, _i, _i, _i, _i, _i]),
    "mdgan_wgrad_gemm": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mdgan_pack_weights": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mdgan_wgrad_unpack": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mdgan_pack_job_words": (_i, []),
    "mdgan_pack_weights_multi": (_i, [_p, _i, _i, _p]),
    "mdgan_reduce_slices": (_i, [_p, _p, _i, _ll, _p]),
    "mdgan_bn_workspace_floats": (_ll, [_i, _i, _i]),
    "mdgan_bn_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _f, _i, _p]),
    "mdgan_bn_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "mdgan_act_backward": (_i, [_p, _p, _p, _ll, _i, _f, _i, _p]),
    "mdgan_tanh_backward": (_i, [_p, _p, _p, _ll, _f, _p]),
    "mdgan_head_pack": (_i, [_p, _p, _i, _i, _p]),
    "mdgan_head_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "mdgan_head_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "mdgan_adam_step": (_i, [_p, _p, _p, _p, _ll, _p, _f, _f, _f, _f, _p]),
    "mdgan_pad_rows": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mdgan_sum_slices": (_i, [_p, _p, _ll, _i, _ll, _p]),
    "mdgan_peer_signal": (_i, [_p, _i, _p, _i, _p]),
    "mdgan_peer_wait": (_i, [_p, _i, _p, _i, _p, _ll, _p]),
    "mdgan_peer_push": (_i, [_p, _p, _i, _ll, _p]),
    "mdgan_peer_push_multicast": (_i, [_p, _p, _ll, _p]),
    "mdgan_tanh_backward_slices": (_i, [_p, _p, _p, _ll, _i, _i, _f, _p]),
    "mdgan_thin_down": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mdgan_thin_up": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mdgan_thin_wgrad_slices": (_i, [_i, _i, _i]),
    "mdgan_thin_wgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mdgan_sgemm": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _f, _p, _f, _p, _f, _i, _p]),
    "mdgan_col_sum": (_i, [_p, _p, _i, _i, _p]),
    "mdgan_linear_head_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "mdgan_linear_head_backward": (_i, [_p, _p, _p, _p, _f, _f, _p, _p, _p, _i, _i, _p]),
}


class MdganLibraryError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is not built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MdganLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the MD-GAN hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc > 0:
        raise MdganLibraryError(f"{what}: CUDA error {rc}")
    names = {-1: "bad argument", -2: "unsupported shape/dtype", -3: "driver entry point / tensor map failure"}
    raise MdganLibraryError(f"{what}: {names.get(rc, rc)}")
