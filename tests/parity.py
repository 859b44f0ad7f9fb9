"""Whole-loop parity harness: the CUDA engine (all N workers hosted by one GPU process) against the CPU oracle
(oracle/mdgan_oracle.py, itself pinned bit-exact to the unmodified reference) on the same seeds and inputs.
Used by tests/test_mdgan_gpu.py and by __graft_entry__.smoke()."""
from __future__ import annotations

from typing import Dict

import torch

from util import plugin, relerr

# scale-normalised max-error tolerances per tensor class, for N iterations of the default tf32x3 precision
TOL = {"loss": 2e-4, "X": 1e-3, "S": 5e-3, "weights": 5e-3, "running": 2e-3}


def build_actor_modules(mod, n_workers: int, seed: int):
    """Models built exactly like bootstrap.init_process does (per-actor seed = seed + rank, workers first, server
    last so the global torch RNG continues as the server's stream)."""
    import bootstrap

    discs = {}
    for n in range(n_workers):
        bootstrap._seed_actor(seed + n + 1)
        d = mod.Discriminator().to(dtype=torch.float32)
        d.apply(bootstrap._weights_init)
        discs[n] = d
    bootstrap._seed_actor(seed)
    g = mod.Generator().to(dtype=torch.float32)
    g.apply(bootstrap._weights_init)
    return g, discs


def run_engine_vs_oracle(name: str, n_workers: int, batch_size: int, epochs: int, swap_interval: int, seed: int = 3,
                         local_epochs: int = 1, precision=None, device: str = "cuda:0") -> Dict[str, object]:
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import _DeviceBatches
    from oracle.mdgan_oracle import OracleMDGAN

    import os

    if precision is not None:
        os.environ["MDGAN_PRECISION"] = precision
    mod = plugin(name)
    dev = torch.device(device)
    dataset = SyntheticImages(mod.SHAPE, n_workers * 4 * batch_size)
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, batch_size, mod.Z_DIM, mod.SHAPE,
                         seed=seed, beta_1=0.5, swap_interval=swap_interval, local_epochs=local_epochs)
    g, discs = build_actor_modules(mod, n_workers, seed)
    cfg = EngineConfig(n_workers=n_workers, batch_size=batch_size, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE),
                       beta_1=0.5, swap_interval=swap_interval, local_epochs=local_epochs, z_source="host")
    shards = routing.split_dataset(len(dataset), n_workers, True)
    sources = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], batch_size), dev, mod.SHAPE)
               for n in range(n_workers)}
    engine = MDGANEngine(cfg, 0, 1, dev, g, discs, sources)
    k, b = engine.k, batch_size
    worst = {"loss": 0.0, "X": 0.0, "S": 0.0, "weights": 0.0, "running": 0.0}
    pairs_ok, nbt_ok = True, True
    for e in range(epochs):
        ref = oracle.step(e, record=True)
        engine.generate()
        worst["X"] = max(worst["X"], relerr(engine.X, ref["X"]))
        engine.train_workers()
        S_ref = torch.zeros((k, b, *mod.SHAPE))
        for n in range(n_workers):
            S_ref[n % k] += ref["feedbacks"][n]
        worst["S"] = max(worst["S"], relerr(engine.S.view(k, b, *mod.SHAPE), S_ref))
        for i, l in enumerate(engine.mean_d_loss()):
            worst["loss"] = max(worst["loss"], abs(l - ref["mean_d_loss"][i]) / abs(ref["mean_d_loss"][i]))
        for i, l in enumerate(engine.g_loss.tolist()):
            worst["loss"] = max(worst["loss"], abs(l - ref["loss_gen"][i]) / abs(ref["loss_gen"][i]))
        engine.update_generator()
        pairs = engine.maybe_swap(e)
        if (pairs is None) != (ref["pairs"] is None) or (pairs is not None and not torch.equal(pairs, ref["pairs"])):
            pairs_ok = False
    engine.sync_modules()
    nets = [("G", g, oracle.G)] + [(f"D{n + 1}", discs[n], oracle.D[n]) for n in range(n_workers)]
    for label, ours, theirs in nets:
        sd, rsd = ours.state_dict(), theirs.state_dict()
        assert list(sd.keys()) == list(rsd.keys())
        for key in sd:
            if key.endswith("num_batches_tracked"):
                nbt_ok &= int(sd[key]) == int(rsd[key])
            elif "running" in key:
                worst["running"] = max(worst["running"], relerr(sd[key], rsd[key]))
            elif name == "CelebA" and key in ("cv2.bias", "cv3.bias"):
                continue  # zero-true-gradient parameters driven by rounding noise in the reference
            else:
                worst["weights"] = max(worst["weights"], relerr(sd[key], rsd[key]))
    ok = pairs_ok and nbt_ok and all(worst[c] <= TOL[c] for c in worst)
    return {"ok": ok, "pairs_bit_exact": pairs_ok, "num_batches_tracked_exact": nbt_ok, **worst}
