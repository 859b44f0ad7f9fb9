"""Process-level runner behind `actors.server.start` / `actors.worker.start`: rendezvous, data shards, the
training loop, and the reference's side outputs (CSV timing logs, weight files, image grids).

Outputs keep the reference's names and schemas so its notebooks keep working:
  logs/mdgan.{N}.{dataset}.server.logs.csv          columns of /root/reference/src/actors/server.py:179-208
  logs/mdgan.{N}.{dataset}.worker.{rank}.logs.csv   columns of /root/reference/src/actors/worker.py:129-152
  weights/generator_{epoch}.pt, weights/generator_final.pt            (server.py:366-367,373-374)
  weights/worker_{rank}/discriminator.pth                             (worker.py:289-292)
  saved_images/real_images.png, saved_images/generated_epoch_{e}.png  (server.py:130-149,344-352)
"""
from __future__ import annotations

import csv
import logging
import os
import queue
import threading
import time
from datetime import timedelta
from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import routing
from .engine import EngineConfig, MDGANEngine

SERVER_COLUMNS = [
    "epoch", "start.epoch", "end.epoch", "start.epoch_calculation", "end.epoch_calculation", "start.send_data",
    "end.send_data", "start.recv_data", "end.recv_data", "start.calc_gradients", "end.calc_gradients",
    "start.agg_gradients", "end.agg_gradients", "start.generate_data", "end.generate_data", "fid", "is", "start.fid",
    "end.fid", "start.is", "end.is", "size.data", "size.feedback", "start.swap", "end.swap", "swap", "size.sent",
    "size.recv",
]
WORKER_COLUMNS = [
    "epoch", "start.epoch", "end.epoch", "start.calc_gradients", "end.calc_gradients", "start.recv_data",
    "end.recv_data", "start.send", "end.send", "start.swap_recv_instruction", "end.swap_recv_instruction",
    "start.load_state_dict", "end.load_state_dict", "start.swap_recv", "end.swap_recv", "start.swap_send",
    "end.swap_send", "swap_with", "mean_d_loss", "size.model", "size.sent", "size.recv",
]
_MB = 1024 ** 2


class _DeviceBatches:
    """Host DataLoader (reference order) -> pinned staging -> device."""

    def __init__(self, stream: routing.RealBatchStream, device: torch.device, shape):
        self.stream, self.device = stream, device
        self.pinned = torch.empty((stream.batch_size, *shape), dtype=torch.float32, pin_memory=True)
        self.dev = torch.empty((stream.batch_size, *shape), dtype=torch.float32, device=device)
        self.next_dev = torch.empty_like(self.dev)   # shadow of `dev` for the early upload (upload_ahead / adopt)

    def stage(self) -> None:
        """Host half: next batch of the reference-order loader into the pinned buffer."""
        self.pinned.copy_(self.stream.next())

    def upload(self) -> None:
        """H2D from the pinned buffer (async on the current stream)."""
        self.dev.copy_(self.pinned, non_blocking=True)

    def upload_ahead(self) -> None:
        """H2D of the NEXT iteration's batch into a shadow buffer (called on the engine's copy stream while the
        current iteration still reads `dev`), MDGANEngine.upload_ahead."""
        self.next_dev.copy_(self.pinned, non_blocking=True)

    def adopt(self) -> None:
        """Shadow buffer -> the fixed buffer the (graph-captured) step reads; device-to-device on the compute stream."""
        self.dev.copy_(self.next_dev, non_blocking=True)

    def __call__(self) -> torch.Tensor:
        return self.dev


class DeviceResidentBatches:
    """Real batches already resident in HBM: the worker's whole shard is uploaded once, in the reference's batch
    order for the first epoch, and every iteration a device-to-device copy moves the next batch into the fixed
    buffer the (possibly graph-captured) step reads.  Used when the shard fits on the device (synthetic-data
    benchmarks, small datasets); `_DeviceBatches` streams from the host otherwise."""

    def __init__(self, stream: routing.RealBatchStream, device: torch.device, shape, n_batches: int):
        self.ring = torch.stack([stream.next() for _ in range(n_batches)]).to(device)
        self.dev = torch.empty((stream.batch_size, *shape), dtype=torch.float32, device=device)
        self.i = 0

    def stage(self) -> None:
        self.dev.copy_(self.ring[self.i])
        self.i = (self.i + 1) % self.ring.shape[0]

    def __call__(self) -> torch.Tensor:
        return self.dev


def _maybe_metrics():
    try:
        from torchmetrics.image.fid import FrechetInceptionDistance  # noqa: F401
        from torchmetrics.image.inception import InceptionScore  # noqa: F401

        return FrechetInceptionDistance, InceptionScore
    except Exception:
        return None, None


def _save_grid(images: torch.Tensor, path: Path, normalize: bool) -> None:
    from torchvision.transforms.functional import to_pil_image
    from torchvision.utils import make_grid

    grid = make_grid(images.cpu().float(), nrow=4, normalize=normalize, value_range=(0, 1), padding=0)
    path.parent.mkdir(parents=True, exist_ok=True)
    to_pil_image(grid).save(path)


class _SnapshotWriter:
    """The reference's periodic side outputs (image grid + generator checkpoint, server.py:336-367) without stalling
    the training loop: the generated batch and the generator's flat state are cloned on the
    compute stream (device-to-device, microseconds, so the next Adam step cannot touch the snapshot), copied to pinned
    host memory on a side stream, and a background thread writes the PNG / .pt once the copy event has fired."""

    def __init__(self, device: torch.device):
        self.device = device
        self.side = torch.cuda.Stream(device=device)
        self.q: "queue.Queue" = queue.Queue()
        self.error: Optional[BaseException] = None
        self.slots = []          # (pinned x, pinned f32, pinned i64, free-flag) ring of two
        self.thread = threading.Thread(target=self._run, name="mdgan-snapshot-writer", daemon=True)
        self.thread.start()

    def _slot(self, x, f32, i64):
        for s in self.slots:
            if s[3].is_set():
                s[3].clear()
                return s
        if len(self.slots) >= 2:       # both in flight: wait for the older one (the loop is producing faster than disk)
            s = self.slots[0]
            s[3].wait()
            s[3].clear()
            self.slots.append(self.slots.pop(0))
            return s
        s = (torch.empty(x.shape, dtype=x.dtype, pin_memory=True), torch.empty(f32.shape, dtype=f32.dtype, pin_memory=True),
             torch.empty(i64.shape, dtype=i64.dtype, pin_memory=True), threading.Event())
        self.slots.append(s)
        return s

    def submit(self, engine: MDGANEngine, image_path: Path, weights_path: Path) -> None:
        main = torch.cuda.current_stream(self.device)
        st = engine.gen.state
        x, f32, i64 = engine.X.detach().clone(), st.state_f32.clone(), st.state_i64.clone()
        hx, hf, hi, free = self._slot(x, f32, i64)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            hx.copy_(x, non_blocking=True)
            hf.copy_(f32, non_blocking=True)
            hi.copy_(i64, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        for t in (x, f32, i64):
            t.record_stream(self.side)
        self.q.put((done, hx, hf, hi, free, st, image_path, weights_path))

    def _run(self) -> None:
        while True:
            item = self.q.get()
            if item is None:
                return
            done, hx, hf, hi, free, st, image_path, weights_path = item
            try:
                done.synchronize()
                fake = hx.clone()
                sd = st.state_dict_from(hf, hi)
                free.set()
                fake = fake.repeat(1, 3, 1, 1) if fake.shape[1] < 3 else fake
                _save_grid((fake + 1) * 0.5, image_path, normalize=False)
                weights_path.parent.mkdir(parents=True, exist_ok=True)
                torch.save(sd, weights_path)
            except BaseException as e:  # noqa: BLE001 -- surfaced by close()
                self.error = e
                free.set()

    def close(self) -> None:
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise RuntimeError("snapshot writer failed") from self.error


def run_node(*, backend: str, proc: int, n_procs: int, world_size: int, device: torch.device, cfg: EngineConfig,
             generator: Optional[nn.Module], discriminators: Dict[int, nn.Module], dataset,
             epochs: int, log_interval: int, log_folder: Path, dataset_name: str, iid: bool = True, n_samples: int = 5,
             engine_hook=None) -> MDGANEngine:
    N = routing.num_workers(world_size)
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"--device {device}: the B200 MD-GAN engine runs on CUDA only (no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", proc % max(torch.cuda.device_count(), 1))
    torch.cuda.set_device(device)
    owns_pg = False
    if n_procs > 1 and not dist.is_initialized():
        if backend != "nccl":
            raise RuntimeError(f"--backend {backend}: multi-GPU runs use NCCL over NVLink (one process per GPU)")
        dist.init_process_group(backend="nccl", rank=proc, world_size=n_procs, timeout=timedelta(weeks=52),
                                device_id=device)
        owns_pg = True
        dist.barrier()
    logging.info(f"GPU process {proc}/{n_procs} on {device}: hosts workers "
                 f"{[n + 1 for n in routing.workers_of_process(proc, n_procs, N)]}" + (" and the server" if proc == 0 else ""))

    # data shards: every process derives the same seed-0 split locally (replaces the index send of server.py:157-167)
    shards = routing.split_dataset(len(dataset), N, iid)
    local = routing.workers_of_process(proc, n_procs, N)
    sources = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], cfg.batch_size), device, cfg.image_shape)
               for n in local}
    engine = MDGANEngine(cfg, proc, n_procs, device, generator, discriminators, sources)
    if engine_hook is not None:
        engine_hook(engine)

    name = f"mdgan.{N}.{dataset_name}"
    log_folder = Path(log_folder)
    log_folder.mkdir(parents=True, exist_ok=True)
    image_dir, weights_dir = Path("saved_images"), Path("weights")
    k, b = engine.k, cfg.batch_size
    img_bytes = 4 * b * cfg.image_shape[0] * cfg.image_shape[1] * cfg.image_shape[2]
    server_f = server_w = None
    if proc == 0:
        server_f = open(log_folder / f"{name}.server.logs.csv", "a", encoding="utf-8")
        server_w = csv.DictWriter(server_f, fieldnames=SERVER_COLUMNS)
        server_w.writeheader()
        g = torch.Generator()
        g.manual_seed(0)  # server.py:130-149: constant real batch for the image grid / FID
        real_eval = next(iter(torch.utils.data.DataLoader(dataset, batch_size=n_samples, shuffle=True, generator=g)))[0]
        real_eval = (real_eval.repeat(1, 3, 1, 1) if real_eval.shape[1] < 3 else real_eval)
        real_eval = (real_eval + 1) * 0.5
        _save_grid(real_eval, image_dir / "real_images.png", normalize=True)
    worker_files, worker_writers = {}, {}
    model_mb = {}
    for n in local:
        f = open(log_folder / f"{name}.worker.{n + 1}.logs.csv", "a", encoding="utf-8")
        w = csv.DictWriter(f, fieldnames=WORKER_COLUMNS)
        w.writeheader()
        worker_files[n], worker_writers[n] = f, w
        m = discriminators[n]
        model_mb[n] = (sum(p.nelement() * p.element_size() for p in m.parameters())
                       + sum(bf.nelement() * bf.element_size() for bf in m.buffers())) / _MB

    # CSV spans are DEVICE times: CUDA events on the compute stream mark the phase boundaries (start of the iteration,
    # after generate+broadcast, after the D steps+feedback+reduce, after G backward+Adam, after the swap) and are
    # converted to wall-clock seconds through one (host time, event) anchor pair; the row of an iteration is written
    # after its loss read-back, which synchronises anyway.  Steady state runs as three CUDA graphs replayed back to
    # back (captured after two eager iterations) so the marks survive graph replay; MDGAN_GRAPH=0 keeps eager launches.
    use_graph = os.environ.get("MDGAN_GRAPH", "1") == "1"
    FID, IS = _maybe_metrics()
    if FID is None and proc == 0:
        logging.warning("torchmetrics is not importable: the fid / is columns of the server CSV stay empty "
                        "(the reference requires torchmetrics, /root/reference/src/actors/server.py:16-17)")
    snapshots = _SnapshotWriter(device) if (proc == 0 and FID is None) else None
    torch.cuda.synchronize(device)
    anchor = torch.cuda.Event(enable_timing=True)
    anchor.record()
    anchor.synchronize()
    anchor_t = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def at(ev) -> float:
        return anchor_t + anchor.elapsed_time(ev) * 1e-3

    for epoch in range(epochs):
        if use_graph and epoch == 2:
            engine.capture(split=True)
        h0 = time.time()
        engine.stage_inputs()                 # host: noise draw in the reference's RNG order, loader batches -> pinned
        marks[0].record()
        engine.device_iteration(marks[1:4])   # uploads + generate / D steps + feedback / G backward + Adam
        engine.prefetch_next(epoch, last=(epoch == epochs - 1))  # next iteration's host inputs while the GPU works
        pairs = engine.maybe_swap(epoch)
        marks[4].record()
        engine.iterations_done += 1
        losses = engine.mean_d_loss()  # one small D2H per iteration, like the reference's losses.mean().item()
        marks[4].synchronize()
        t0, t1, t2, t3, t4 = (at(m) for m in marks)
        srow = {c: None for c in SERVER_COLUMNS}
        srow.update({"epoch": epoch, "start.epoch": h0, "start.epoch_calculation": t0, "swap": False,
                     "size.data": 2 * img_bytes / _MB, "size.feedback": N * img_bytes / _MB,
                     "size.sent": N * 2 * img_bytes / _MB, "size.recv": N * img_bytes / _MB,
                     "start.generate_data": t0, "end.generate_data": t1,
                     # the batch leaves inside the generate phase (peer stores / broadcast are its last kernels)
                     "start.send_data": t1, "end.send_data": t1,
                     # waiting for the workers = their D steps + feedback (+ the reduce / flag wait)
                     "start.recv_data": t1, "end.recv_data": t2,
                     # aggregation = the ONE generator backward on the group-summed feedback, fused with Adam
                     "start.agg_gradients": t2, "end.agg_gradients": t3,
                     "start.calc_gradients": t2, "end.calc_gradients": t3, "end.epoch_calculation": t4})
        wrows = {n: {c: None for c in WORKER_COLUMNS} for n in local}
        for n in local:
            wrows[n].update({"epoch": epoch, "start.epoch": h0, "size.model": model_mb[n],
                             "size.sent": img_bytes / _MB, "size.recv": 2 * img_bytes / _MB,
                             "start.recv_data": t0, "end.recv_data": t1, "start.calc_gradients": t1,
                             "end.calc_gradients": t2, "start.send": t2, "end.send": t2, "end.epoch": t4})
        if pairs is not None:
            srow.update({"swap": True, "start.swap": t3, "end.swap": t4})
            for n in local:
                wrows[n].update({"swap_with": engine.swap_partner(n), "start.swap_recv_instruction": t3,
                                 "end.swap_recv_instruction": t3, "start.swap_send": t3, "end.swap_send": t4,
                                 "start.swap_recv": t3, "end.swap_recv": t4, "start.load_state_dict": t4,
                                 "end.load_state_dict": t4})
                wrows[n]["size.sent"] += model_mb[n]
                wrows[n]["size.recv"] += model_mb[n]
        for i, n in enumerate(local):
            wrows[n]["mean_d_loss"] = losses[i]
            worker_writers[n].writerow(wrows[n])
        log_now = epoch % log_interval == 0 or epoch == epochs - 1
        if proc == 0:
            if log_now and snapshots is not None:  # server.py:336-367
                snapshots.submit(engine, image_dir / f"generated_epoch_{epoch}.png", weights_dir / f"generator_{epoch}.pt")
            elif log_now:  # with torchmetrics: FID / IS need the images now
                fake = engine.X.detach().cpu()
                fake = fake.repeat(1, 3, 1, 1) if fake.shape[1] < 3 else fake
                fake = (fake + 1) * 0.5
                _save_grid(fake, image_dir / f"generated_epoch_{epoch}.png", normalize=False)
                if FID is not None:
                    ev = fake[: min(n_samples, len(fake))]
                    srow["start.is"] = time.time()
                    inc = IS(normalize=True, splits=1)
                    inc.update(ev)
                    srow["is"] = inc.compute()[0].item()
                    srow["end.is"] = srow["start.fid"] = time.time()
                    fid = FID(normalize=True)
                    fid.update(ev, real=True)
                    fid.update(real_eval, real=False)
                    srow["fid"] = fid.compute().item()
                    srow["end.fid"] = time.time()
                engine.gen.state.store_to(generator)
                weights_dir.mkdir(parents=True, exist_ok=True)
                torch.save(generator.state_dict(), weights_dir / f"generator_{epoch}.pt")
            srow["end.epoch"] = time.time()
            server_w.writerow(srow)
        if log_now and FID is not None and n_procs > 1:
            # the synchronous evaluation above can take minutes on process 0: hold the other processes on the HOST
            # here instead of letting them replay the next iteration and spin in a flag wait on the device
            dist.barrier(group=engine.exchange.ctl_group)

    torch.cuda.synchronize(device)
    if snapshots is not None:
        snapshots.close()
    engine.sync_modules()
    engine.close()
    if proc == 0:
        weights_dir.mkdir(parents=True, exist_ok=True)
        torch.save(generator.state_dict(), weights_dir / "generator_final.pt")
        server_f.close()
    for n in local:
        d = weights_dir / f"worker_{n + 1}"
        d.mkdir(parents=True, exist_ok=True)
        torch.save(discriminators[n].state_dict(), d / "discriminator.pth")
        worker_files[n].close()
    if owns_pg:
        dist.barrier()
        dist.destroy_process_group()
    return engine
