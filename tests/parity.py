"""Whole-loop parity harness: the CUDA engine (all N workers hosted by one GPU process) against the CPU oracle
(oracle/mdgan_oracle.py, itself pinned bit-exact to the unmodified reference) on the same seeds and inputs.
Used by tests/test_mdgan_gpu.py and by __graft_entry__.smoke().

Why two modes.  The MD-GAN loop is chaotic at the rounding level: a (Leaky)ReLU pre-activation that is zero to
within fp32 rounding flips its gate, and Adam's first steps are sign-like (update = lr * g / (|g| + 1e-8)), so the
reference's own fp32 run and the *same code in fp64* drift apart to ~4e-2 (max-norm, feedback tensors) within three
iterations (tools/drift_calibration.py; numbers in DESIGN.md).  Parity is therefore stated as:

  along-trajectory : every iteration starts from the reference's state (weights, BatchNorm buffers, Adam moments
      and step counts, taken from the fp32 oracle), runs ONE full iteration in the engine, and must reproduce
      the oracle's generated batch, per-worker losses, group-summed feedback, Adam moments and post-step weights
      within TOL -- against the fp32 oracle, or, where a rounding-tied gate makes the fp32 oracle itself deviate
      from exact arithmetic, against the fp64 twin started from the same state (`agrees`, tests/util.py).  The
      reference state is also re-imposed at the two points inside an iteration where a sign-like Adam step would
      turn one tied gate into an O(1e-1) downstream difference (after the discriminators' Adam step, before the
      generator backward), so that every phase is judged on its own arithmetic.  Swap pairs, routing and
      num_batches_tracked are bit-exact.
  free-running     : the engine carries its own state for E iterations; its drift from the fp32 oracle must stay
      within FREE_TOL (a bound of the same order as the fp32-vs-fp64 drift of the reference itself).
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List

import torch

from util import agrees, plugin, relerr

# along-trajectory tolerances (scale-normalised max error unless noted), default tf32x3 precision.  Measured on B200
# (tools/parity_report.py, gpurun_out r2: 12 configurations up to CelebA K=8 b=64): loss <= 1.2e-6, X <= 3.0e-6,
# running <= 2.0e-6, feedback of images without a tied gate <= see S; bounds are <= 10x the worst measured value.
TOL = {
    "loss": 1e-5,      # relative, per-worker mean_d_loss and loss_gen
    "X": 3e-5,         # generated batch
    "S": 1e-3,         # group-summed feedback = grad-output of the generator backward, per image (north-star bar)
    "moments": 5e-3,   # Adam exp_avg after the step (== gradient parity), rel. L2 over the flat buffer (one tied
                       # gate in the training pass moves it by ~1e-3; gate-free runs measure ~1e-6..3e-5)
    "update": 1e-1,    # rel. L2 of the applied weight update (w_after - w_before).  Adam's first steps are sign-like
                       # (lr * g / (|g| + 1e-8)): every gradient element at rounding level flips its full-lr step,
                       # so this only bounds the flipped fraction (measured 1e-4..5e-2); the gradients themselves
                       # are held to "moments", the Adam arithmetic to tests/test_kernels_gpu.py::test_adam
    "abs_w": 2.1,      # max |w_ours - w_ref| in units of lr
    "running": 2e-5,   # BatchNorm running statistics
}
# the un-patched iteration (mode "unpatched") carries the engine's own post-Adam discriminator weights into the feedback
# pass and the engine's own feedback into the generator backward: sign-like first Adam steps turn gradient elements at
# rounding level into +-lr weight differences, so its bounds are those of one chaotic step
UNPATCHED_TOL = {"loss": 1e-3, "X": 3e-5, "S_l2": 5e-2, "moments": 5e-2, "update": 3e-1, "abs_w": 2.1, "running": 1e-3}
# Feedback (judged on the reference's post-Adam discriminator weights, see run_engine_vs_oracle).  A LeakyReLU
# pre-activation that is zero to within fp32 rounding lands on either side of its gate depending on summation order; the
# gradient of THAT image then differs by O(1e-2) between two correct fp32 implementations.  The harness detects these
# events exactly: it compares the sign pattern of every LeakyReLU output of the engine's feedback pass with the fp32
# oracle's, element by element.  Images WITHOUT a sign difference must meet TOL["S"] -- all of them (S_BAD_IMAGES = 0);
# images with one are listed by index in the report (`tied_gate_images`), may not exceed TIED_IMAGES of the batch and
# must still stay within S_TIED; the whole tensor stays within S_L2 (rel. L2).
S_BAD_IMAGES = 0.0
TIED_IMAGES = 0.35
S_TIED = 1e-1
S_L2 = 1e-2
# free-running drift bounds after <= 4 iterations (rel. L2)
FREE_TOL = {"loss": 5e-2, "X": 5e-2, "weights_l2": 2e-2}


def l2err(got: torch.Tensor, ref: torch.Tensor) -> float:
    g, r = got.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    return ((g - r).norm() / r.norm().clamp_min(1e-30)).item()


def feedback_parity(S: torch.Tensor, ref32: torch.Tensor, ref64: torch.Tensor, tied: torch.Tensor):
    """S: [k, b, C, H, W]; tied: [k, b] bool, images whose feedback pass had a LeakyReLU sign difference against the
    fp32 oracle.  Returns (fraction of UNTIED images that miss TOL["S"], worst per-image error among the untied images,
    worst per-image error among the tied images, rel. L2 of the whole tensor against the closer reference)."""
    S = S.detach().double().cpu()
    per_img = []
    for ref in (ref32.double(), ref64.double()):
        scale = ref.abs().amax(dim=(2, 3, 4), keepdim=True).clamp_min(1e-30)  # per image
        per_img.append(((S - ref).abs() / scale).amax(dim=(2, 3, 4)).reshape(-1))
    e32, e = per_img[0], torch.minimum(per_img[0], per_img[1])
    tied = tied.reshape(-1)
    untied = ~tied
    # untied images: same gates as the fp32 oracle, so they are held to the fp32 oracle itself
    bad_frac = (e32[untied] > TOL["S"]).double().mean().item() if untied.any() else 0.0
    worst_untied = e32[untied].max().item() if untied.any() else 0.0
    worst_tied = e[tied].max().item() if tied.any() else 0.0
    return bad_frac, worst_untied, worst_tied, min(l2err(S, ref32), l2err(S, ref64))


def build_actor_modules(mod, n_workers: int, seed: int):
    """Models built exactly like bootstrap.init_process does (per-actor seed = seed + rank, workers first, server
    last so the global torch RNG continues as the server's stream)."""
    import bootstrap

    discs = {}
    for n in range(n_workers):
        bootstrap._seed_actor(seed + n + 1)
        d = mod.Discriminator().to(dtype=torch.float32)
        d.apply(bootstrap._weights_init)
        d._mdgan_rng_state = torch.get_rng_state()   # as bootstrap.init_process does
        discs[n] = d
    bootstrap._seed_actor(seed)
    g = mod.Generator().to(dtype=torch.float32)
    g.apply(bootstrap._weights_init)
    return g, discs


def _copy_oracle(dst, src) -> None:
    """dst <- src: module state (cast to dst's dtype) and Adam state."""
    dst.G.load_state_dict(src.G.state_dict())
    dst.opt_g.load_state_dict(copy.deepcopy(src.opt_g.state_dict()))
    for d, s, od, os_ in zip(dst.D, src.D, dst.opt_d, src.opt_d):
        d.load_state_dict(s.state_dict())
        od.load_state_dict(copy.deepcopy(os_.state_dict()))


def _load_engine(engine, oracle) -> None:
    engine.gen.state.load_from(oracle.G)
    engine.gen.state.load_adam(oracle.opt_g, oracle.G)
    engine.gen.repack()
    for n, net in engine.disc.items():
        net.state.load_from(oracle.D[n])
        net.state.load_adam(oracle.opt_d[n], oracle.D[n])
        net.repack()


def _flat(params) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1).double().cpu() for p in params])


def _adam_flat(opt, module, key) -> torch.Tensor:
    return torch.cat([opt.state[p][key].detach().reshape(-1).double().cpu() for p in module.parameters()])


def run_engine_vs_oracle(name: str, n_workers: int, batch_size: int, epochs: int, swap_interval: int, seed: int = 3,
                         local_epochs: int = 1, precision=None, device: str = "cuda:0",
                         mode: str = "trajectory") -> Dict[str, object]:
    from datasets.DataPartitioner import SyntheticImages
    from mdgan_b200 import routing
    from mdgan_b200.engine import EngineConfig, MDGANEngine
    from mdgan_b200.node import _DeviceBatches
    from oracle.mdgan_oracle import OracleMDGAN

    if precision is not None:
        os.environ["MDGAN_PRECISION"] = precision
    trace = bool(os.environ.get("MDGAN_PARITY_TRACE"))
    mod = plugin(name)
    dev = torch.device(device)
    lr = 2e-4
    dataset = SyntheticImages(mod.SHAPE, n_workers * 4 * batch_size)
    okw = dict(seed=seed, beta_1=0.5, swap_interval=swap_interval, local_epochs=local_epochs, generator_lr=lr,
               discriminator_lr=lr)
    oracle = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, batch_size, mod.Z_DIM, mod.SHAPE, **okw)
    traj = mode in ("trajectory", "unpatched")   # per-iteration comparison from the reference's state
    patched = mode == "trajectory"               # ... with the reference state re-imposed inside the iteration
    twin = OracleMDGAN(mod.Generator, mod.Discriminator, dataset, n_workers, batch_size, mod.Z_DIM, mod.SHAPE,
                       dtype=torch.float64, **okw) if traj else None
    g, discs = build_actor_modules(mod, n_workers, seed)
    cfg = EngineConfig(n_workers=n_workers, batch_size=batch_size, z_dim=mod.Z_DIM, image_shape=tuple(mod.SHAPE),
                       generator_lr=lr, discriminator_lr=lr, beta_1=0.5, swap_interval=swap_interval,
                       local_epochs=local_epochs, z_source="host")
    shards = routing.split_dataset(len(dataset), n_workers, True)
    sources = {n: _DeviceBatches(routing.RealBatchStream(dataset, shards[n], batch_size), dev, mod.SHAPE)
               for n in range(n_workers)}
    engine = MDGANEngine(cfg, 0, 1, dev, g, discs, sources)
    scratch = copy.deepcopy(discs[0])  # (constructing a new module would consume the global RNG stream)
    k, b = engine.k, batch_size
    worst = {c: 0.0 for c in (list(TOL) + ["S_bad_images", "S_l2", "S_tied", "tied_images"] if traj else FREE_TOL)}
    tol = TOL if mode != "unpatched" else {**TOL, **UNPATCHED_TOL}
    tied_log: List[str] = []
    failures: List[str] = []
    pairs_ok, nbt_ok = True, True

    def check(cls: str, what: str, ok: bool, val: float) -> None:
        worst[cls] = max(worst[cls], val)
        if not ok:
            failures.append(f"{what}={val:.3e}")

    for e in range(epochs):
        if traj and e > 0:
            _copy_oracle(twin, oracle)
            _load_engine(engine, oracle)
        w_before = {"G": _flat(oracle.G.parameters()), **{n: _flat(oracle.D[n].parameters()) for n in range(n_workers)}}
        ref = oracle.step(e, record=True, record_gates=patched)
        ref64 = twin.step(e, record=True, z=ref["z"], replay_reals=ref["real"], pairs=ref["pairs"]) if twin else None
        tied = torch.zeros((k, b), dtype=torch.bool)
        engine.generate()
        w_after_adam = {}
        if patched:
            # engine.train_workers(), with the reference's post-Adam discriminator state loaded between the training
            # step and the feedback pass: the first Adam steps are sign-like, so ONE rounding-tied gate in the
            # training pass (gradient off by ~5e-4 rel. L2) flips ~1e-4 of the lr-steps and moves the feedback of
            # that worker by ~6e-2 -- the feedback kernels are judged on the reference's weights, the Adam step on
            # the "moments"/"update"/"abs_w" metrics below.
            engine.S.zero_()
            for i, n in enumerate(engine.local):
                ig, id_ = routing.route(n, k)
                x_g, x_d = engine.X[ig * b:(ig + 1) * b], engine.X[id_ * b:(id_ + 1) * b]
                real = engine.real_sources[n]()
                net = engine.disc[n]
                for l in range(local_epochs):
                    engine.d_loss[i, l].copy_(net.train_step(real, x_d))
                w_after_adam[n] = net.state.params.detach().double().cpu().clone()
                scratch.load_state_dict(ref["d_mid"][n])
                net.state.load_from(scratch)
                net.repack()
                slot = routing.feedback_slot(n, k)
                engine.g_loss[i].copy_(net.feedback_step(x_g, out=engine.S[slot * b:(slot + 1) * b], accumulate=True))
                # LeakyReLU sign pattern of THIS feedback pass against the fp32 oracle's, element by element
                for l, g_ref in enumerate(ref["fb_gates"][n]):
                    ours = (net.a[l][:b] > 0).permute(0, 3, 1, 2).cpu()
                    diff = (ours != g_ref).flatten(1)
                    for img in diff.any(dim=1).nonzero().flatten().tolist():
                        tied[slot, img] = True
                        tied_log.append(f"iter {e} worker {n + 1} image {img} layer {l}: "
                                        f"{int(diff[img].sum())} gate(s), first at flat index {int(diff[img].nonzero()[0])}")
        else:
            engine.train_workers()
            for n in engine.local:  # the feedback pass does not touch the parameters: this is the post-Adam state
                w_after_adam[n] = engine.disc[n].state.params.detach().double().cpu().clone()
        S_ref = torch.zeros((k, b, *mod.SHAPE))
        S_ref64 = torch.zeros((k, b, *mod.SHAPE), dtype=torch.float64)
        for n in range(n_workers):
            S_ref[n % k] += ref["feedbacks"][n]
            if ref64:
                S_ref64[n % k] += ref64["feedbacks"][n]
        S = engine.S.view(k, b, *mod.SHAPE).clone()
        d_l, g_l = engine.mean_d_loss(), engine.g_loss.tolist()
        loss_err = max(max(abs(d_l[i] - ref["mean_d_loss"][i]) / abs(ref["mean_d_loss"][i]),
                           abs(g_l[i] - ref["loss_gen"][i]) / abs(ref["loss_gen"][i])) for i in range(n_workers))
        if patched:
            # the generator phase is judged on the reference's feedback (a tied gate upstream is accounted for above)
            engine.S.copy_(S_ref.view_as(engine.S).to(dev))
        engine.update_generator()
        pairs = engine.maybe_swap(e)
        if (pairs is None) != (ref["pairs"] is None) or (pairs is not None and not torch.equal(pairs, ref["pairs"])):
            pairs_ok = False
        if trace:
            print(f"  iter {e}: X {relerr(engine.X, ref['X']):.2e} S {relerr(S, S_ref):.2e} loss {loss_err:.2e}\n"
                  f"     d_loss {d_l} ref {ref['mean_d_loss']}\n     g_loss {g_l} ref {ref['loss_gen']}\n"
                  f"     per-slot S err {[relerr(S[i], S_ref[i]) for i in range(k)]}", flush=True)
        if traj:
            check("loss", f"loss@{e}", loss_err <= tol["loss"], loss_err)
            check("X", f"X@{e}", agrees(engine.X, ref["X"], ref64["X"], tol["X"]), relerr(engine.X, ref["X"]))
            bad_frac, s_err, s_tied, s_l2 = feedback_parity(S, S_ref, S_ref64, tied)
            if patched:
                check("S", f"S@{e}", s_err <= TOL["S"], s_err)
                check("S_bad_images", f"S_bad_images@{e}", bad_frac <= S_BAD_IMAGES, bad_frac)
                check("S_tied", f"S_tied@{e}", s_tied <= S_TIED, s_tied)
                check("tied_images", f"tied_images@{e}", tied.double().mean().item() <= TIED_IMAGES, tied.double().mean().item())
            check("S_l2", f"S_l2@{e}", s_l2 <= (S_L2 if patched else UNPATCHED_TOL["S_l2"]), s_l2)
            engine.sync_modules()
            # after a swap worker a holds what partner c trained: compare module-for-module (oracle swapped too)
            nets = [("G", engine.gen, g, oracle.G, oracle.opt_g, twin.G, twin.opt_g)]
            nets += [(n, engine.disc[n], discs[n], oracle.D[n], oracle.opt_d[n], twin.D[n], twin.opt_d[n])
                     for n in range(n_workers)]
            part = routing.partners_from_pairs(pairs) if pairs is not None else {}
            for label, net, ours, theirs, opt, theirs64, opt64 in nets:
                # Adam moments stay with the rank (worker.py:281 copies parameters in place), weights move
                m_ref, m_ref64 = _adam_flat(opt, theirs, "exp_avg"), _adam_flat(opt64, theirs64, "exp_avg")
                m_err = min(l2err(net.state.m, m_ref), l2err(net.state.m, m_ref64))
                check("moments", f"m[{label}]@{e}", m_err <= tol["moments"], m_err)
                if trace:
                    print(f"     moments[{label}]@{e} {m_err:.2e}", flush=True)
                if label == "G":
                    wb = w_before["G"]
                    wa_ref, wa_ref64, wa = _flat(theirs.parameters()), _flat(theirs64.parameters()), _flat(ours.parameters())
                else:  # the Adam step of worker `label`, before the reference state was loaded / any swap
                    wb = w_before[label]
                    names = [kk for kk, _ in theirs.named_parameters()]
                    wa_ref = torch.cat([ref["d_mid"][label][kk].reshape(-1).double() for kk in names])
                    wa_ref64 = torch.cat([ref64["d_mid"][label][kk].reshape(-1).double() for kk in names])
                    wa = w_after_adam[label]
                u_err = min(l2err(wa - wb, wa_ref - wb), l2err(wa - wb, wa_ref64 - wb))
                check("update", f"update[{label}]@{e}", u_err <= tol["update"], u_err)
                a_err = (wa - wa_ref).abs().max().item() / lr
                check("abs_w", f"abs_w[{label}]@{e}", a_err <= tol["abs_w"], a_err)
                sd, rsd, rsd64 = ours.state_dict(), theirs.state_dict(), theirs64.state_dict()
                assert list(sd.keys()) == list(rsd.keys())
                for key in sd:
                    if key.endswith("num_batches_tracked"):
                        nbt_ok &= int(sd[key]) == int(rsd[key])
                    elif "running_mean" in key and name == "CelebA" and label != "G":
                        # running_mean of a BatchNorm behind a biased conv tracks the chaotic bias
                        continue
                    elif "running" in key:
                        check("running", f"{label}.{key}@{e}", agrees(sd[key], rsd[key], rsd64[key], tol["running"]),
                              relerr(sd[key], rsd[key]))
        else:
            check("loss", f"loss@{e}", loss_err <= FREE_TOL["loss"], loss_err)
            check("X", f"X@{e}", l2err(engine.X, ref["X"]) <= FREE_TOL["X"], l2err(engine.X, ref["X"]))
    if not traj:
        engine.sync_modules()
        for label, ours, theirs in [("G", g, oracle.G)] + [(f"D{n + 1}", discs[n], oracle.D[n]) for n in range(n_workers)]:
            sd, rsd = ours.state_dict(), theirs.state_dict()
            assert list(sd.keys()) == list(rsd.keys())
            for key in sd:
                if key.endswith("num_batches_tracked"):
                    nbt_ok &= int(sd[key]) == int(rsd[key])
            w_err = l2err(_flat(ours.parameters()), _flat(theirs.parameters()))
            check("weights_l2", f"weights[{label}]", w_err <= FREE_TOL["weights_l2"], w_err)
    ok = pairs_ok and nbt_ok and not failures
    return {"ok": ok, "mode": mode, "pairs_bit_exact": pairs_ok, "num_batches_tracked_exact": nbt_ok,
            "failures": failures[:8], "tied_gate_images": tied_log[:40], "tied_gate_events": len(tied_log), **worst}
