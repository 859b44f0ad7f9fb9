"""Stand-in for the absent `torchmetrics` package (test infrastructure only).

The reference imports FID / Inception-Score metrics at module import time
(/root/reference/src/actors/server.py:16-17, standalone_gan.py:10-11).  They
carry no hot-path arithmetic, need a network download of InceptionV3 weights,
and are only evaluated at `log_interval` boundaries, so this stub returns NaN.
"""
