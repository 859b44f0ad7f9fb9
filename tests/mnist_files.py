"""MNIST-format files for the ingest tests: the four idx files torchvision.datasets.MNIST(download=False) reads from
<root>/MNIST/raw, filled with seeded synthetic digits (no network here, so the real files cannot be fetched; the
torchvision class, its PIL decode and the plugin's ToTensor + Normalize transform are the real ones)."""
import struct
from pathlib import Path

import numpy as np


def write_mnist_idx(root, n_train: int, n_test: int, seed: int = 7) -> Path:
    raw = Path(root) / "MNIST" / "raw"
    raw.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(seed)
    for prefix, n in (("train", n_train), ("t10k", n_test)):
        # blurred random strokes, uint8, so that neighbouring samples differ everywhere and values span 0..255
        img = rng.integers(0, 256, size=(n, 28, 28), dtype=np.uint8)
        lab = rng.integers(0, 10, size=(n,), dtype=np.uint8)
        with open(raw / f"{prefix}-images-idx3-ubyte", "wb") as f:
            f.write(struct.pack(">IIII", 0x00000803, n, 28, 28))
            f.write(img.tobytes())
        with open(raw / f"{prefix}-labels-idx1-ubyte", "wb") as f:
            f.write(struct.pack(">II", 0x00000801, n))
            f.write(lab.tobytes())
    return Path(root)
