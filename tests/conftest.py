import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "distributed-gan_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)
# our `datasets` plugin package must win over the unrelated HuggingFace `datasets` distribution
if "datasets" in sys.modules and not str(getattr(sys.modules["datasets"], "__file__", "")).startswith(str(PKG)):
    for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
        del sys.modules[k]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
