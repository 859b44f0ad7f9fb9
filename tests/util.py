"""Shared helpers for the parity tests."""
import importlib

import torch


def relerr(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got - ref| / max |ref| in float64 (scale-normalised max error)."""
    g, r = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((g - r).abs().max() / r.abs().max().clamp_min(1e-30)).item()


def plugin(name: str):
    return importlib.import_module(f"datasets.{name}")


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def init_model(cls, seed: int):
    from oracle.mdgan_oracle import weights_init

    torch.manual_seed(seed)
    m = cls()
    m.apply(weights_init)
    m.train()
    return m


def agrees(got: torch.Tensor, ref32: torch.Tensor, ref64: torch.Tensor, tol: float) -> bool:
    """Parity criterion for gradients that flow through (Leaky)ReLU gates.

    A pre-activation that is zero to within fp32 rounding (|y| ~ 1e-7, about one element per million) lands on
    either side of the gate depending on the summation order of the producing convolution; two correct fp32
    implementations then differ by one gate, which moves the reduced gradients behind it by ~1e-2 of their
    maximum (measured: the reference's own torch-CPU fp32 run against its fp64 run, tools/layer_probe.py).  So a
    tensor passes if it is within `tol` of the reference's fp32 result, or within `tol` of the same module
    evaluated in fp64 (exact arithmetic) -- i.e. wherever the reference is stable to rounding we match it, and
    where it is not we match the exact answer."""
    return relerr(got, ref32) < tol or relerr(got, ref64) < tol
