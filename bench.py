#!/usr/bin/env python
"""Benchmark of the MD-GAN data-parallel training step (BASELINE.json: generator iterations/s, device-timed).

    python bench.py --gpus 1 --steps K --warmup W              # our arm (sm_100a kernels), one process
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU over NCCL
    python bench.py --impl reference ...                       # the reference's CPU path (oracle port) on host cores

Workload (DESIGN.md "Measurement"): MD-GAN with K = --gpus discriminator workers, one per GPU, the generator on
rank 0 (north star "K=1/2/4/8 B200"), DCGAN on synthetic MNIST-shape 1x28x28 images, per-worker batch 64
(BASELINE.json configs[1]); --dataset / --batch select the CIFAR-10 / CelebA shapes and the batch sweep.
A step = one generator iteration: G forward over k*b noise vectors -> every worker's D step (real + X_d, Adam) and
error feedback on X_g -> feedback sum/reduce -> one G backward -> G Adam.  Per-GPU work is fixed as the number of
GPUs grows (one more worker per GPU), i.e. weak scaling; `value` is the whole-job rate
generator-iterations/s x workers (worker-iterations/s), `generator_it_s` the plain rate.

One JSON line on stdout (rank 0).  Keys beyond the base contract: roofline, cpu_baseline, e2e, clocks, gpu_launches,
per_op (device-time share of every kernel family in one instrumented iteration).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
PKG = REPO / "distributed-gan_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)
if "datasets" in sys.modules and not str(getattr(sys.modules["datasets"], "__file__", "")).startswith(str(PKG)):
    for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
        del sys.modules[k]

import torch  # noqa: E402

METRIC = "generator iters/sec x workers (device-timed MD-GAN step, one discriminator worker per GPU)"
UNIT = "worker-it/s"
SEED = 3  # run-distributed.sh:5
LR, BETA_1, BETA_2 = 2e-4, 0.5, 0.999


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--dataset", default="MNIST_DCGAN", choices=["MNIST_DCGAN", "CIFAR10", "CelebA"])
    ap.add_argument("--batch", type=int, default=64, help="per-worker batch size b")
    ap.add_argument("--workers", type=int, default=0, help="discriminator workers (default: one per GPU)")
    ap.add_argument("--precision", default="tf32x3", choices=["tf32x3", "tf32"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying the CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline budget")
    return ap.parse_args()


def workload_name(args, n_workers: int) -> str:
    shape = {"MNIST_DCGAN": "MNIST-shape 1x28x28", "CIFAR10": "CIFAR-10-shape 3x32x32", "CelebA": "CelebA-shape 3x64x64"}
    return (f"MD-GAN {shape[args.dataset]} DCGAN, K={n_workers} workers, per-worker batch {args.batch}, "
            f"k=2 generated batches, Adam lr 2e-4 betas (0.5, 0.999), swap off in the timed window")


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self) -> None:
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        with contextlib.suppress(Exception):
            self.proc.wait(timeout=5)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            with contextlib.suppress(ValueError):
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            for nm, val in zip(names, parts[3:7]):
                if val == "Active":
                    reasons.add(nm)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


class OpTimer:
    """Observer for mdgan_b200.ops.set_observer: CUDA events around every kernel-family launch of one iteration."""

    def __init__(self):
        self.records = []  # (name, n_kernels, flops, bytes, ev_start, ev_end)

    @contextlib.contextmanager
    def __call__(self, name, n_kernels, flops, nbytes):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.records.append((name, n_kernels, flops, nbytes, s, e))

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, nk, fl, by, s, e in self.records:
            a = agg.setdefault(name, {"calls": 0, "kernels": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["calls"] += 1
            a["kernels"] += nk
            a["ms"] += s.elapsed_time(e)
            a["flops"] += fl
            a["bytes"] += by
        return agg


class LaunchCounter:
    def __init__(self):
        self.kernels = 0

    @contextlib.contextmanager
    def __call__(self, name, n_kernels, flops, nbytes):
        self.kernels += n_kernels
        yield


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def build_modules(mod, workers, seed):
    """bootstrap.init_process order: every hosted worker actor on seed + rank, the server last on seed + 0."""
    import bootstrap

    discs = {}
    for n in workers:
        bootstrap._seed_actor(seed + n + 1)
        d = mod.Discriminator().to(dtype=torch.float32)
        d.apply(bootstrap._weights_init)
        discs[n] = d
    return discs


def build_generator(mod, seed):
    import bootstrap

    bootstrap._seed_actor(seed)
    g = mod.Generator().to(dtype=torch.float32)
    g.apply(bootstrap._weights_init)
    return g


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> None:
    """The reference's own algorithm on the host CPU cores: the oracle port (oracle/mdgan_oracle.py, pinned bit-exact
    to the unmodified reference run) with all the host threads torch can use.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import importlib

    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleMDGAN

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_workers = args.workers or args.gpus
    mod = importlib.import_module(f"datasets.{args.dataset}")
    dataset = SyntheticImages(mod.SHAPE, n_workers * 16 * args.batch)
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, args.batch, mod.Z_DIM, mod.SHAPE,
                         seed=SEED, generator_lr=LR, discriminator_lr=LR, beta_1=BETA_1, beta_2=BETA_2)
    t0 = time.perf_counter()
    oracle.step(0, record=False)
    first = time.perf_counter() - t0
    warm = max(1, min(args.warmup, int(20.0 / max(first, 1e-3))))
    steps = max(1, min(args.steps, int(120.0 / max(first, 1e-3))))
    for e in range(1, warm):
        oracle.step(e, record=False)
    t0 = time.perf_counter()
    for e in range(steps):
        oracle.step(warm + e, record=False)
    dt = (time.perf_counter() - t0) / steps
    value = n_workers / dt
    sample = (f"{steps} full iterations of the same workload (K={n_workers}, b={args.batch}) after {warm} warm-up, "
              f"single process, torch CPU fp32 with {cores} intra-op threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "generator_it_s": 1.0 / dt,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n_workers), "dataset": args.dataset, "batch": args.batch,
                   "workers": n_workers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_baseline(args, n_workers: int) -> dict:
    import importlib

    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleMDGAN

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mod = importlib.import_module(f"datasets.{args.dataset}")
    dataset = SyntheticImages(mod.SHAPE, n_workers * 16 * args.batch)
    rng = torch.get_rng_state()
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, args.batch, mod.Z_DIM, mod.SHAPE,
                         seed=SEED, generator_lr=LR, discriminator_lr=LR, beta_1=BETA_1, beta_2=BETA_2)
    t0 = time.perf_counter()
    oracle.step(0, record=False)
    first = time.perf_counter() - t0
    steps = max(2, min(200, int(args.cpu_seconds / max(first, 1e-3))))
    oracle.step(1, record=False)
    t0 = time.perf_counter()
    for e in range(steps):
        oracle.step(2 + e, record=False)
    dt = (time.perf_counter() - t0) / steps
    torch.set_rng_state(rng)
    return {"value": n_workers / dt, "unit": UNIT, "generator_it_s": 1.0 / dt, "cores": cores, "kind": "port",
            "sample": f"{steps} full iterations of the same workload (K={n_workers}, b={args.batch}) after 2 warm-up, "
                      f"oracle port of server.py:213-333 + worker.py:157-284, torch CPU fp32, {cores} threads"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import importlib

    import torch.distributed as dist

    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import _lib, ops, routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import DeviceResidentBatches, _DeviceBatches

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch one rank per GPU (torchrun)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.check(_lib.load().mdgan_check_device(), "device check (sm_100a)")
    os.environ["MDGAN_PRECISION"] = args.precision
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    n_workers = args.workers or args.gpus
    mod = importlib.import_module(f"datasets.{args.dataset}")
    b, shape = args.batch, tuple(mod.SHAPE)
    local = routing.workers_of_process(rank, world, n_workers)
    dataset = SyntheticImages(shape, n_workers * 16 * b)  # M = N*16*b
    shards = routing.split_dataset(len(dataset), n_workers, True)

    def make_engine(resident: bool, z_source: str):
        discs = build_modules(mod, local, SEED)
        gen = build_generator(mod, SEED) if rank == 0 else None
        cfg = EngineConfig(n_workers=n_workers, batch_size=b, z_dim=mod.Z_DIM, image_shape=shape, generator_lr=LR,
                           discriminator_lr=LR, beta_1=BETA_1, beta_2=BETA_2, swap_interval=10 ** 9, local_epochs=1,
                           z_source=z_source, prefetch_host=not resident)
        if resident:
            src = {n: DeviceResidentBatches(routing.RealBatchStream(dataset, shards[n], b), dev, shape, 16) for n in local}
        else:
            src = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], b), dev, shape) for n in local}
        return MDGANEngine(cfg, rank, world, dev, gen, discs, src)

    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)  # 512 MB > 126 MB L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------------------------------------------------------- device-resident leg (`value`)
    engine = make_engine(resident=True, z_source="device")
    for e in range(max(args.warmup, 3)):
        engine.iteration(e)
    counter = LaunchCounter()
    ops.set_observer(counter)
    engine.iteration(0)
    ops.set_observer(None)
    launches_per_step = counter.kernels
    graphed = not args.no_graph
    if graphed:
        engine.capture()
        for e in range(3):
            engine.iteration(e)
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    if rank == 0:
        sampler.start()
    for i in range(args.steps):
        flush.zero_()                 # evict L2 between timed iterations (outside the per-step event pair)
        engine.stage_inputs()         # device-to-device: next resident real batch into the step's input buffer
        ev[i][0].record()
        engine.device_iteration()     # graph replay (or eager launches with --no-graph)
        ev[i][1].record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev))
    ms_per_step = ms_total / args.steps
    gen_it_s = 1e3 / ms_per_step
    value = gen_it_s * n_workers

    # ---------------------------------------------------------------- instrumented iteration (roofline, per-op shares)
    if graphed:
        engine.graph = None
    timer = OpTimer()
    for rep in range(5):
        flush.zero_()
        engine.stage_inputs()
        if rep == 0:
            engine.device_iteration()  # eager warm-up after the graph replays
            continue
        # Hold the stream with a spin kernel (~10 ms) while the host queues the whole iteration + its events, so the
        # event intervals measure device execution back to back, not the host's per-launch latency.
        torch.cuda._sleep(20_000_000)
        ops.set_observer(timer)
        engine.device_iteration()
        ops.set_observer(None)
        torch.cuda.synchronize(dev)
    per_op = timer.summary()
    peaks = load_peaks()
    total_ms = sum(a["ms"] for a in per_op.values()) or 1.0
    top = max(per_op, key=lambda n: per_op[n]["ms"])
    a = per_op[top]
    tensor_bound = top in ("conv_down", "conv_up", "conv_dense", "wgrad_gemm")
    traffic, traffic_note = None, None
    tpath = REPO / "profiles" / "r01_traffic.json"
    if tpath.exists() and args.dataset == "MNIST_DCGAN" and b == 64:
        kname = "void wgrad_gemm_ta_kernel" if top == "wgrad_gemm" else ("void conv_gemm_ta_kernel" if tensor_bound else None)
        rec = json.loads(tpath.read_text()).get("MNIST_DCGAN_b64_final_kernels", {}).get(kname)
        if rec:
            traffic = rec["dram_bytes_per_launch_avg"]
            traffic_note = ("ncu --set full capture of this workload (profiles/r01_traffic.json), cold caches, average "
                            "over the launches of the kernel")
    if tensor_bound:
        achieved = a["flops"] / (a["ms"] * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        x3 = args.precision == "tf32x3"
        ceiling = peak / (6.0 if x3 else 2.0)
        roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "mode_ceiling": ceiling, "frac_of_mode_ceiling": achieved / ceiling,
                    "peak_note": f"{peaks['source']} dense bf16 cuBLAS (sustained) from MEASURED_PEAKS.json; kind::tf32 "
                                 "runs at half that rate and the tf32x3 parity mode issues 3 MMAs per algorithmic MAC "
                                 "(mode_ceiling = peak/6 for tf32x3, peak/2 for tf32); achieved counts ALGORITHMIC "
                                 "FLOPs (2*M*N*K of the layer), not the three tensor-core passes"}
    else:
        achieved = a["bytes"] / (a["ms"] * 1e-3) / 1e9
        peak = peaks["hbm_gbs"]
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_note": f"{peaks['source']} HBM copy bandwidth"}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    roofline["avg_launch_us"] = a["ms"] * 1e3 / max(a["calls"], 1)
    roofline["launches_timed"] = a["calls"]
    roofline["share_of_step"] = a["ms"] / total_ms
    engine.close()
    exchange_mode = getattr(engine.exchange, "mode", "nccl") if world > 1 else "none (one process)"
    shares = {n: {"share": round(v["ms"] / total_ms, 4), "us_per_iter": round(v["ms"] * 1e3 / 4, 2),
                  "calls_per_iter": v["calls"] // 4} for n, v in sorted(per_op.items(), key=lambda t: -t[1]["ms"])}
    del engine

    # ---------------------------------------------------------------- end-to-end leg (host buffers in, losses out)
    engine = make_engine(resident=False, z_source="host")
    loss_host = torch.empty((len(local), 2), dtype=torch.float32, pin_memory=True)
    for e in range(max(args.warmup, 3)):
        engine.iteration(e)
    if graphed:
        engine.capture()
        for e in range(3):
            engine.iteration(e)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for i in range(args.steps):
        flush.zero_()
        ev2[i][0].record()
        engine.iteration(i)                       # host RNG + pinned staging, H2D, the step (+ next step's host staging)
        loss_host[:, 0].copy_(engine.d_loss[:, 0], non_blocking=True)   # D2H of the step's result
        loss_host[:, 1].copy_(engine.g_loss, non_blocking=True)
        ev2[i][1].record()
        ev2[i][1].synchronize()                   # the caller reads the losses every iteration (worker.py:215)
    sync_all()
    e2e_ms = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev2)) / args.steps
    img_bytes = 4 * b * shape[0] * shape[1] * shape[2]
    h2d = len(local) * img_bytes + (4 * engine.k * b * mod.Z_DIM if rank == 0 else 0)
    e2e = {"value": n_workers * 1e3 / e2e_ms, "unit": UNIT, "generator_it_s": 1e3 / e2e_ms, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * len(local),
           "api": "MDGANEngine.iteration (the loop body of actors.server.start / actors.worker.start): host torch RNG "
                  "noise + host DataLoader batches -> pinned -> device, losses read back every iteration"}
    engine.close()
    del engine

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, n_workers)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "generator_it_s": gen_it_s, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "tf32x3 (fp32-accurate)",
        "data": "synthetic",
        "config": {"workload": workload_name(args, n_workers), "dataset": args.dataset, "batch": b, "workers": n_workers,
                   "parallelism": f"one discriminator worker per GPU x{args.gpus}, generator on rank 0",
                   "precision": args.precision, "cuda_graph": graphed, "exchange": exchange_mode,
                   "l2": "512 MB buffer written between timed iterations (outside the per-step event pairs)",
                   "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "tensor_pipe_util": {"source": "committed ncu --set full captures, profiles/r01_ncu_full_conv_wgrad.md (not measured "
                                       "by this run: ncu numbers are never bench values)",
                             "metric": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                             "conv_128_wide_tiles_pct": [53, 75], "conv_64_wide_tiles_pct": [22, 30],
                             "weight_gradient_pct": [22, 41]},
        "gpu_launches_per_step": launches_per_step, "per_op": shares,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    args = parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, torch warnings) are sent to stderr; the JSON goes to the saved descriptor.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
