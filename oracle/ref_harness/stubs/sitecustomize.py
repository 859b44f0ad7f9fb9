"""Auto-imported by the interpreter when this directory is on PYTHONPATH.
Swaps torchvision's downloading dataset classes for the synthetic ones when
MDGAN_SYNTH_M is set (the parent *and* every mp.spawn child see it)."""
import os

if os.environ.get("MDGAN_SYNTH_M"):
    try:
        import torchvision.datasets as _tvd
        import synthetic_tv as _s

        _tvd.MNIST = _s.MNIST
        _tvd.CIFAR10 = _s.CIFAR10
        _tvd.CelebA = _s.CelebA
    except Exception as _e:  # pragma: no cover
        import sys

        print(f"[ref_harness] synthetic dataset patch failed: {_e}", file=sys.stderr)
