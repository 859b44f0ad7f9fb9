"""Where one iteration of the MLP workload (reference plugin datasets/MNIST.py, K = 1, b = 64) spends its device time:
the captured step between CUDA events (device-resident batches, device noise and masks), then an eager iteration with
CUDA events around every kernel-family launch (bench.OpTimer), the stream held by a spin kernel while the host queues.
    python tools/mlp_profile.py [--batch 64] [--steps 30]      -> one JSON line"""
import argparse
import importlib
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402

import torch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import ops, routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import DeviceResidentBatches

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    mod = importlib.import_module("datasets.MNIST")
    b, shape = a.batch, tuple(mod.SHAPE)
    data = SyntheticImages(shape, 16 * b)
    discs = bench.build_modules(mod, [0], bench.SEED)
    gen = bench.build_generator(mod, bench.SEED)
    cfg = EngineConfig(n_workers=1, batch_size=b, z_dim=mod.Z_DIM, image_shape=shape, generator_lr=bench.LR,
                       discriminator_lr=bench.LR, beta_1=bench.BETA_1, beta_2=bench.BETA_2, swap_interval=10 ** 9,
                       z_source="device")
    shards = routing.split_dataset(len(data), 1, True)
    src = {0: DeviceResidentBatches(routing.RealBatchStream(data, shards[0], b), dev, shape, 16)}
    eng = MDGANEngine(cfg, 0, 1, dev, gen, discs, src)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    for e in range(3):
        eng.iteration(e)
    eng.capture()
    for e in range(3):
        eng.iteration(e)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    for i in range(a.steps):
        flush.zero_()
        eng.stage_inputs()
        ev[i][0].record()
        eng.device_iteration()
        ev[i][1].record()
    torch.cuda.synchronize()
    graph_ms = sum(s.elapsed_time(e) for s, e in ev) / a.steps
    eng.graph = None
    timer = bench.OpTimer()
    reps = 4
    for rep in range(reps + 1):
        flush.zero_()
        eng.stage_inputs()
        if rep == 0:
            eng.device_iteration()
            continue
        torch.cuda._sleep(20_000_000)
        ops.set_observer(timer)
        eng.device_iteration()
        ops.set_observer(None)
        torch.cuda.synchronize()
    per = timer.summary()
    out = {"workload": f"MD-GAN MNIST reference MLP, K=1, b={b}", "graph_ms_per_step": round(graph_ms, 4),
           "eager_us_per_iter": {n: {"us": round(v["ms"] * 1e3 / reps, 2), "calls": v["calls"] // reps,
                                     "gflops": round(v["flops"] / reps / 1e9, 3),
                                     "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2)}
                                 for n, v in sorted(per.items(), key=lambda t: -t[1]["ms"])}}
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
