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
