"""End-to-end step time with and without the early upload (MDGAN_PREFETCH_H2D), same process, alternating legs.

The loop is bench.py's `e2e` leg: MDGANEngine.iteration with host RNG noise + streamed host batches (pinned -> device),
the captured graph, the losses read back and the stream synchronised every step, an L2 flush between steps.
    python tools/e2e_ab.py [--dataset CelebA] [--batch 64] [--steps 40] [--rounds 2]
Prints one JSON line: per leg the mean of the per-step CUDA-event intervals (ms) and the host wall time per step."""
import argparse
import importlib
import json
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402  (puts the package on sys.path)

import torch  # noqa: E402


def leg(mod, ahead: str, b: int, steps: int, dev, flush):
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import _DeviceBatches

    os.environ["MDGAN_PREFETCH_H2D"] = ahead
    shape = tuple(mod.SHAPE)
    data = SyntheticImages(shape, 16 * b)
    discs = bench.build_modules(mod, [0], bench.SEED)
    gen = bench.build_generator(mod, bench.SEED)
    cfg = EngineConfig(n_workers=1, batch_size=b, z_dim=mod.Z_DIM, image_shape=shape, generator_lr=bench.LR,
                       discriminator_lr=bench.LR, beta_1=bench.BETA_1, beta_2=bench.BETA_2, swap_interval=10 ** 9,
                       z_source="host", prefetch_host=True)
    shards = routing.split_dataset(len(data), 1, True)
    src = {0: _DeviceBatches(routing.RealBatchStream(data, shards[0], b), dev, shape)}
    eng = MDGANEngine(cfg, 0, 1, dev, gen, discs, src)
    loss_host = torch.empty((1, 2), dtype=torch.float32, pin_memory=True)
    epoch = 0
    for _ in range(3):
        eng.iteration(epoch)
        epoch += 1
    eng.capture()
    for _ in range(3):
        eng.iteration(epoch)
        epoch += 1
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        eng.iteration(epoch)
        epoch += 1
        loss_host[:, 0].copy_(eng.d_loss[:, 0], non_blocking=True)
        loss_host[:, 1].copy_(eng.g_loss, non_blocking=True)
        ev[i][1].record()
        ev[i][1].synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ms = sum(s.elapsed_time(e) for s, e in ev) / steps
    eng.close()
    return ms, wall


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="CelebA")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--rounds", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MDGAN_PRECISION", "tf32x3")
    mod = importlib.import_module(f"datasets.{a.dataset}")
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    out = {"dataset": a.dataset, "batch": a.batch, "steps": a.steps, "early_upload": [], "compute_stream_upload": []}
    for _ in range(a.rounds):
        for ahead, key in (("1", "early_upload"), ("0", "compute_stream_upload")):
            ms, wall = leg(mod, ahead, a.batch, a.steps, dev, flush)
            out[key].append({"event_ms_per_step": round(ms, 4), "wall_ms_per_step": round(wall, 4)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
