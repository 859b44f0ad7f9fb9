#!/usr/bin/env python
"""Benchmark of the MD-GAN data-parallel training step (BASELINE.json: generator iterations/s, device-timed).

    python bench.py --gpus 1 --steps K --warmup W              # our arm (sm_100a kernels), one process
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU over NCCL
    python bench.py --impl reference ...                       # the UNMODIFIED reference (oracle/_ref) on the host cores:
                                                               # bootstrap.py --backend gloo --device cpu, N+1 processes

Workload (DESIGN.md "Measurement"): MD-GAN with K = --gpus discriminator workers, one per GPU, the generator on
rank 0 (north star "K=1/2/4/8 B200"), the reference's CelebA-shape 3x64x64 DCGAN, per-worker batch 64 (BASELINE.json
configs[3], the largest configuration that fits one GPU); --dataset / --batch / --swap-interval select the other
configurations.  The JSON line also carries `shapes`: the device-timed step of the MNIST-shape DCGAN (configs[1]), of
the CIFAR-10-shape DCGAN with a discriminator swap EVERY iteration inside the timed window (configs[2]) and, at one
GPU, of the reference's own MNIST plugin (the MLP with always-on dropout; device noise and masks).
A step = one generator iteration: G forward over k*b noise vectors -> every worker's D step (real + X_d, Adam) and
error feedback on X_g -> feedback sum/reduce -> one G backward -> G Adam.  Per-GPU work is fixed as the number of
GPUs grows (one more worker per GPU), i.e. weak scaling; `value` is the whole-job rate
generator-iterations/s x workers (worker-iterations/s), `generator_it_s` the plain rate.

One JSON line on stdout (rank 0).  Keys beyond the base contract: roofline, cpu_baseline, e2e (MDGANEngine.iteration
with HOST inputs -- host RNG noise, loader batches, pinned staging, H2D, losses read back every step; the next step's
inputs are staged and uploaded while the current step runs), clocks, gpu_launches, per_op (device-time share of every
kernel family in one instrumented iteration), multi_gpu_bit_identical (N > 1).
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--dataset", default="CelebA", choices=["MNIST_DCGAN", "CIFAR10", "CelebA"])
    ap.add_argument("--swap-interval", type=int, default=0,
                    help="discriminator swap every this many iterations INSIDE the timed window (0 = off)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the per-shape sub-lines (MNIST / CIFAR-10 shape)")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the multi-GPU bit-identity pre-check (N > 1)")
    ap.add_argument("--batch", type=int, default=64, help="per-worker batch size b")
    ap.add_argument("--workers", type=int, default=0, help="discriminator workers (default: one per GPU)")
    ap.add_argument("--precision", default="tf32x3", choices=["tf32x3", "tf32"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying the CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline budget")
    return ap.parse_args()


def workload_name(args, n_workers: int, dataset=None, swap=None) -> str:
    shape = {"MNIST_DCGAN": "MNIST-shape 1x28x28", "CIFAR10": "CIFAR-10-shape 3x32x32", "CelebA": "CelebA-shape 3x64x64",
             "MNIST": "MNIST 1x28x28 reference MLP (Linear / LeakyReLU / dropout 0.3)"}
    swap = args.swap_interval if swap is None else swap
    swap_txt = f"discriminator swap every {swap} iteration(s) inside the timed window" if swap > 0 and n_workers > 1 \
        else "swap off in the timed window"
    return (f"MD-GAN {shape[dataset or args.dataset]} DCGAN, K={n_workers} workers, per-worker batch {args.batch}, "
            f"k=2 generated batches, Adam lr 2e-4 betas (0.5, 0.999), {swap_txt}")


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
        d._mdgan_rng_state = torch.get_rng_state()
        discs[n] = d
    return discs


def build_generator(mod, seed):
    import bootstrap

    bootstrap._seed_actor(seed)
    g = mod.Generator().to(dtype=torch.float32)
    g.apply(bootstrap._weights_init)
    return g


# ------------------------------------------------------------------------------------------------ reference arm
REFERENCE_DATASETS = ("CIFAR10", "CelebA")   # plugins the reference ships with a DCGAN (MNIST_DCGAN is this repo's)


def _oracle_port_ms(args, n_workers: int, steps: int, warmup: int, budget_s: float):
    """The oracle restatement (single process, all host threads).  Returns (ms_per_step, steps_run, warmup_run)."""
    import importlib

    from datasets.DataPartitioner import SyntheticImages
    from oracle.mdgan_oracle import OracleMDGAN

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mod = importlib.import_module(f"datasets.{args.dataset}")
    dataset = SyntheticImages(mod.SHAPE, n_workers * 16 * args.batch)
    rng = torch.get_rng_state()
    swap = args.swap_interval if args.swap_interval > 0 else 10 ** 9
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, args.batch, mod.Z_DIM, mod.SHAPE,
                         seed=SEED, generator_lr=LR, discriminator_lr=LR, beta_1=BETA_1, beta_2=BETA_2,
                         swap_interval=swap)
    t0 = time.perf_counter()
    oracle.step(0, record=False)
    first = time.perf_counter() - t0
    fit = int(budget_s / max(first, 1e-3))
    warm = max(1, min(warmup, max(1, fit // 5)))
    steps_run = max(1, min(steps, fit - warm))
    for e in range(1, warm):
        oracle.step(e, record=False)
    t0 = time.perf_counter()
    for e in range(steps_run):
        oracle.step(warm + e, record=False)
    dt = (time.perf_counter() - t0) / steps_run
    torch.set_rng_state(rng)
    return dt * 1e3, steps_run, warm


def reference_cpu(args, n_workers: int, steps: int, warmup: int, budget_s: float) -> dict:
    """The reference's own CPU implementation of the path on this host's cores.  kind "reference": the UNMODIFIED
    reference sources (oracle/_ref, installed by oracle/ref_harness/install_ref.py) run as `bootstrap.py --backend gloo
    --device cpu`, N+1 OS processes, OMP threads = floor(cores / (N+1)).  kind "port": the oracle
    restatement in one process -- only for the MNIST-shape DCGAN, which the reference does not ship, or when
    oracle/_ref is absent."""
    cores = os.cpu_count() or 1
    note = None
    if args.dataset in REFERENCE_DATASETS:
        try:
            from oracle.ref_harness import time_reference

            if time_reference.reference_src() is not None:
                swap = args.swap_interval if args.swap_interval > 0 else 10 ** 9
                r = time_reference.time_distributed(args.dataset, n_workers, args.batch, steps, warmup, swap_interval=swap,
                                                    timeout_s=budget_s)
                return {"ms_per_step": r["ms_per_step"], "steps": r["steps"], "warmup": r["warmup"], "cores": cores,
                        "kind": "reference", "phases_ms": r["phases_ms"],
                        "sample": (f"{r['steps']} generator iterations of the same workload (K={n_workers}, b={args.batch}) "
                                   f"after {r['warmup']} warm-up, UNMODIFIED reference actors ({r['launcher']}) --backend gloo --device cpu: "
                                   f"{r['processes']} processes x {r['threads_per_process']} OMP threads on {cores} cores, server "
                                   "CSV end.epoch_calculation - start.epoch_calculation")}
            note = "oracle/_ref is not installed on this host"
        except Exception as e:  # noqa: BLE001 -- time-out or a failed spawn: say so and fall back to the port
            note = f"reference run failed ({type(e).__name__}: {str(e)[:200]})"
    else:
        note = f"the reference ships no {args.dataset} plugin"
    ms, steps_run, warm = _oracle_port_ms(args, n_workers, steps, warmup, min(budget_s, 240.0))
    return {"ms_per_step": ms, "steps": steps_run, "warmup": warm, "cores": cores, "kind": "port", "note": note,
            "sample": (f"{steps_run} generator iterations of the same workload (K={n_workers}, b={args.batch}) after {warm} "
                       f"warm-up, oracle port of server.py:213-333 + worker.py:157-284 in ONE process, torch CPU fp32, "
                       f"{cores} intra-op threads ({note})")}


def run_reference(args) -> None:
    """`--impl reference`: the reference's CPU path with every host thread it can use.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_workers = args.workers or args.gpus
    r = reference_cpu(args, n_workers, args.steps, args.warmup, budget_s=780.0)
    dt = r["ms_per_step"] * 1e-3
    value = n_workers / dt
    cpu = {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "generator_it_s": 1.0 / dt,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n_workers), "dataset": args.dataset, "batch": args.batch,
                   "workers": n_workers, "swap_interval": args.swap_interval},
        "cpu_baseline": cpu, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if r["steps"] != args.steps or r["warmup"] != args.warmup:
        line["steps_note"] = (f"ran {r['steps']} steps / {r['warmup']} warm-up instead of the requested {args.steps} / "
                              f"{args.warmup}: the CPU run is bounded to a few minutes")
    if r.get("phases_ms"):
        line["phases_ms"] = r["phases_ms"]
    emit(line)


def cpu_baseline(args, n_workers: int) -> dict:
    """cpu_baseline of our arm's line (N = 1): a bounded sample (10 iterations after 2 warm-up) of the same workload."""
    r = reference_cpu(args, n_workers, 10, 2, budget_s=max(60.0, 8 * args.cpu_seconds))
    dt = r["ms_per_step"] * 1e-3
    return {"value": n_workers / dt, "unit": UNIT, "generator_it_s": 1.0 / dt, "cores": r["cores"], "kind": r["kind"],
            "sample": r["sample"], "phases_ms": r.get("phases_ms")}


# ------------------------------------------------------------------------------------------------ our arm
def _profile_json(name: str):
    path = REPO / "profiles" / name
    return json.loads(path.read_text()) if path.exists() else None


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
    b = args.batch
    local = routing.workers_of_process(rank, world, n_workers)
    graphed = not args.no_graph
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

    # ---------------------------------------------------------------- multi-GPU correctness pre-check
    selfcheck = None
    if world > 1 and not args.no_selfcheck:
        from mdgan_b200.selfcheck import multi_gpu_bit_identity

        selfcheck = multi_gpu_bit_identity(rank, world, dev, "CIFAR10", None, 16, 5, 2, graph=True, seed=SEED)
        sync_all()

    def make_engine(mod, dataset, shards, resident: bool, z_source: str, swap: int):
        discs = build_modules(mod, local, SEED)
        gen = build_generator(mod, SEED) if rank == 0 else None
        shape = tuple(mod.SHAPE)
        cfg = EngineConfig(n_workers=n_workers, batch_size=b, z_dim=mod.Z_DIM, image_shape=shape, generator_lr=LR,
                           discriminator_lr=LR, beta_1=BETA_1, beta_2=BETA_2, swap_interval=swap if swap > 0 else 10 ** 9,
                           local_epochs=1, z_source=z_source, prefetch_host=not resident)
        if resident:
            src = {n: DeviceResidentBatches(routing.RealBatchStream(dataset, shards[n], b), dev, shape, 16) for n in local}
        else:
            src = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], b), dev, shape) for n in local}
        return MDGANEngine(cfg, rank, world, dev, gen, discs, src)

    def device_leg(ds_name: str, swap: int, steps: int, warmup: int, sample_clocks: bool):
        """`value` leg: inputs resident in HBM, the captured step (+ the host-driven swap when due) between CUDA events."""
        mod = importlib.import_module(f"datasets.{ds_name}")
        dataset = SyntheticImages(tuple(mod.SHAPE), n_workers * 16 * b)  # M = N*16*b
        shards = routing.split_dataset(len(dataset), n_workers, True)
        engine = make_engine(mod, dataset, shards, resident=True, z_source="device", swap=swap)
        epoch = 0
        for _ in range(max(warmup, 3)):
            engine.iteration(epoch)
            epoch += 1
        counter = LaunchCounter()
        ops.set_observer(counter)
        engine.stage_inputs()
        engine.device_iteration()
        ops.set_observer(None)
        if graphed:
            engine.capture()
            for _ in range(3):
                engine.iteration(epoch)
                epoch += 1
        sampler = ClockSampler(local_rank) if sample_clocks and rank == 0 else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        swaps = 0
        sync_all()
        if sampler:
            sampler.start()
        for i in range(steps):
            flush.zero_()                 # evict L2 between timed iterations (outside the per-step event pair)
            engine.stage_inputs()         # device-to-device: next resident real batch into the step's input buffer
            ev[i][0].record()
            engine.device_iteration()     # graph replay (or eager launches with --no-graph)
            if swap > 0 and engine.maybe_swap(epoch) is not None:   # server.py:315-333, worker.py:239-284
                swaps += 1
            ev[i][1].record()
            epoch += 1
        sync_all()
        clocks = sampler.stop() if sampler else None
        mine = sum(s_.elapsed_time(e_) for s_, e_ in ev) / steps
        ms = max_over_ranks(mine * steps) / steps
        per_rank = None
        if world > 1:   # every rank's own device time per step (rank 0 also runs the generator)
            try:
                t = torch.zeros(world, device=dev, dtype=torch.float64)
                t[rank] = mine
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                per_rank = [round(x, 4) for x in t.tolist()]
            except Exception:  # noqa: BLE001 -- diagnostics only
                per_rank = None
        return engine, mod, dataset, shards, {"ms_per_step": ms, "launches_per_step": counter.kernels, "swaps": swaps,
                                               "clocks": clocks, "per_rank_ms": per_rank}

    # ---------------------------------------------------------------- device-resident leg (`value`)
    engine, mod, dataset, shards, leg = device_leg(args.dataset, args.swap_interval, args.steps, args.warmup, True)
    shape = tuple(mod.SHAPE)
    ms_per_step, launches_per_step, clocks = leg["ms_per_step"], leg["launches_per_step"], leg["clocks"]
    push_mode = getattr(engine.exchange, "push_mode", None)
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
    tensor_families = ("conv_down", "conv_up", "conv_dense", "wgrad_gemm")
    tensor_bound = top in tensor_families
    traffic, traffic_note = None, None
    tj = _profile_json("r02_traffic.json")
    if tj:
        rec = tj.get(f"{args.dataset}_b{b}", {}).get(top)
        if rec:
            traffic = rec["dram_bytes_per_launch_avg"]
            traffic_note = tj.get("how")
    x3 = args.precision == "tf32x3"
    if tensor_bound:
        achieved = a["flops"] / (a["ms"] * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
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
    # all tensor-core GEMMs of the step together: algorithmic FLOPs / their device time
    t_ms = sum(per_op[n]["ms"] for n in tensor_families if n in per_op)
    t_fl = sum(per_op[n]["flops"] for n in tensor_families if n in per_op)
    if t_ms > 0:
        roofline["all_gemms"] = {"achieved": t_fl / (t_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                 "frac": t_fl / (t_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                                 "frac_of_mode_ceiling": t_fl / (t_ms * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] / (6.0 if x3 else 2.0)),
                                 "share_of_step": t_ms / total_ms}
    step_flops = sum(v["flops"] for v in per_op.values()) / 4
    # exchange kernels of the instrumented eager iteration on EVERY rank: a flag wait shows how long that rank idled
    exchange_us = None
    if world > 1:
        try:
            mine_x = {n: round(per_op[n]["ms"] * 1e3 / 4, 2) for n in ("peer_push", "peer_signal", "peer_wait") if n in per_op}
            mine_x["iteration"] = round(total_ms * 1e3 / 4, 1)
            gathered = [None] * world
            dist.all_gather_object(gathered, mine_x)
            exchange_us = gathered
        except Exception:  # noqa: BLE001 -- diagnostics only
            exchange_us = None
    engine.close()
    exchange_mode = getattr(engine.exchange, "mode", "nccl") if world > 1 else "none (one process)"
    shares = {n: {"share": round(v["ms"] / total_ms, 4), "us_per_iter": round(v["ms"] * 1e3 / 4, 2),
                  "calls_per_iter": v["calls"] // 4} for n, v in sorted(per_op.items(), key=lambda t: -t[1]["ms"])}
    del engine

    # ---------------------------------------------------------------- end-to-end leg (host buffers in, losses out)
    engine = make_engine(mod, dataset, shards, resident=False, z_source="host", swap=args.swap_interval)
    loss_host = torch.empty((len(local), 2), dtype=torch.float32, pin_memory=True)
    epoch = 0
    for _ in range(max(args.warmup, 3)):
        engine.iteration(epoch)
        epoch += 1
    if graphed:
        engine.capture()
        for _ in range(3):
            engine.iteration(epoch)
            epoch += 1
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for i in range(args.steps):
        flush.zero_()
        ev2[i][0].record()
        engine.iteration(epoch)                   # host RNG + pinned staging, H2D, the step (+ swap when due, + next step's host staging)
        epoch += 1
        loss_host[:, 0].copy_(engine.d_loss[:, 0], non_blocking=True)   # D2H of the step's result
        loss_host[:, 1].copy_(engine.g_loss, non_blocking=True)
        ev2[i][1].record()
        ev2[i][1].synchronize()                   # the caller reads the losses every iteration (worker.py:215)
    sync_all()
    e2e_ms = max_over_ranks(sum(s_.elapsed_time(e_) for s_, e_ in ev2)) / args.steps
    img_bytes = 4 * b * shape[0] * shape[1] * shape[2]
    h2d = len(local) * img_bytes + (4 * engine.k * b * mod.Z_DIM if rank == 0 else 0)
    e2e = {"value": n_workers * 1e3 / e2e_ms, "unit": UNIT, "generator_it_s": 1e3 / e2e_ms, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * len(local),
           "api": "MDGANEngine.iteration (the loop body of actors.server.start / actors.worker.start): host torch RNG "
                  "noise + host DataLoader batches -> pinned -> device, losses read back every iteration; the host staging "
                  "of step i+1 runs while step i computes" + (
                      " and so does its H2D copy (copy stream -> shadow buffers, adopted device-to-device at the start "
                      "of step i+1; MDGAN_PREFETCH_H2D=0 uploads on the compute stream)"
                      if getattr(engine, "_h2d_ahead", False) else
                      "; the H2D copy is on the compute stream (the early upload is on by default for single-process "
                      "runs only, MDGAN_PREFETCH_H2D=force)")}
    engine.close()
    del engine

    # ---------------------------------------------------------------- per-shape sub-lines (device-timed step only)
    shapes = {}
    if not args.no_shapes:
        sub_steps = min(args.steps, 20)
        for ds_name, swap in (("MNIST_DCGAN", 0), ("CIFAR10", 1)):
            if ds_name == args.dataset and swap == args.swap_interval:
                continue
            eng, _, _, _, sub = device_leg(ds_name, swap, sub_steps, 3, False)
            eng.close()
            del eng
            shapes[f"{ds_name}_b{b}" + ("_swap1" if swap and n_workers > 1 else "")] = {
                "workload": workload_name(args, n_workers, ds_name, swap), "ms_per_step": sub["ms_per_step"],
                "generator_it_s": 1e3 / sub["ms_per_step"], "value": n_workers * 1e3 / sub["ms_per_step"], "unit": UNIT,
                "steps": sub_steps, "swaps_in_timed_window": sub["swaps"], "gpu_launches_per_step": sub["launches_per_step"]}
        if world == 1:
            # the reference's own MNIST plugin (MLP with always-on dropout; the host draws its masks every step,
            # outside the per-step event pair like the other host staging; their upload is inside)
            try:
                eng, _, _, _, sub = device_leg("MNIST", 0, sub_steps, 3, False)
                eng.close()
                del eng
                shapes[f"MNIST_MLP_b{b}"] = {
                    "workload": workload_name(args, n_workers, "MNIST", 0).replace(" DCGAN,", ","),
                    "ms_per_step": sub["ms_per_step"], "generator_it_s": 1e3 / sub["ms_per_step"],
                    "value": n_workers * 1e3 / sub["ms_per_step"], "unit": UNIT, "steps": sub_steps,
                    "swaps_in_timed_window": 0, "gpu_launches_per_step": sub["launches_per_step"]}
            except Exception as e:  # noqa: BLE001 -- a sub-line must not take the headline down with it
                shapes[f"MNIST_MLP_b{b}"] = {"error": repr(e)[:300]}

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
        # `config` names the workload and is identical in both arms; how OUR arm ran it is in `setup`
        "config": {"workload": workload_name(args, n_workers), "dataset": args.dataset, "batch": b, "workers": n_workers,
                   "swap_interval": args.swap_interval},
        "setup": {"parallelism": f"one discriminator worker per GPU x{args.gpus}, generator on rank 0",
                  "precision": args.precision, "cuda_graph": graphed, "exchange": exchange_mode,
                  "swaps_in_timed_window": leg["swaps"], "push": push_mode, "per_rank_ms": leg["per_rank_ms"],
                  "exchange_us_per_rank_eager": exchange_us,
                  "l2": "512 MB buffer written between timed iterations (outside the per-step event pairs)",
                  "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step, "step_gflop_algorithmic_rank0": step_flops / 1e9,
        "step_tflops_algorithmic_rank0": step_flops / (ms_per_step * 1e-3) / 1e12, "per_op": shares, "shapes": shapes,
    }
    if selfcheck is not None:
        line["multi_gpu_bit_identical"] = bool(selfcheck["ok"])
        line["multi_gpu_check"] = selfcheck
    tp = _profile_json("r02_tensor_pipe.json")
    if tp:
        line["tensor_pipe_util"] = tp
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
