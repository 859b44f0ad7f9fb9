"""The three exchanges of an MD-GAN iteration over torch.distributed (NCCL over NVLink on the GPU box; the same code
runs on gloo/CPU tensors in the host-logic tests).

Reference call sites and what replaces them:
  C4  server.py:238-246 isend of stack([K[n%k], K[(n+1)%k]]) to every worker, worker.py:181-182 recv
        -> ONE broadcast of X [k*b, C, H, W] from process 0; every process slices its chunks locally
           (for k = 2 every worker needs all of X anyway).
  C3  worker.py:232-233 send of dBCE/dX_g, server.py:234 irecv into feedbacks[n]
        -> every process sums its local workers' feedback into slot n % k of a [k*b, C, H, W] buffer
           (done by the dgrad kernel's accumulate epilogue) and ONE reduce(sum) to process 0 lands the
           group-summed grad-output directly where the generator backward reads it.
  C5  server.py:321-332 isend of the partner rank, worker.py:243-246 recv
        -> one broadcast of the int32 pair table drawn on process 0's host RNG (bit-exact with the reference).
  C6  worker.py:252-282 TensorDict irecv/send of the whole state_dict, one message per leaf
        -> one grouped isend/irecv per remote partner on the flat state (fp32 block + int64 counters) into staging
           buffers (the reference receives in place and races with its own send).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import routing


class Exchange:
    mode = "nccl"   # how C3 / C4 travel: torch.distributed collectives (NCCL on GPUs, gloo in the CPU protocol tests)

    def __init__(self, proc: int, n_procs: int, n_workers: int):
        self.proc, self.n_procs, self.n_workers = proc, n_procs, n_workers
        if n_procs > 1 and not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before building a multi-process Exchange")
        # The per-iteration collectives (C3, C4) run on the default group and may be captured into a CUDA graph; the
        # rare, host-driven swap traffic gets its own communicators so that eager operations never interleave with
        # graph-captured ones on one NCCL communicator: a host-side (gloo) group for the pair table, which is host
        # data anyway, and a second device group for the state exchange.
        self.ctl_group = self.swap_group = None
        self._swap_staging: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}   # per hosted worker, allocated once
        if n_procs > 1:
            self.ctl_group = dist.new_group(backend="gloo")
            self.swap_group = dist.new_group()
            self._connect_all_pairs()

    def _connect_all_pairs(self) -> None:
        """NCCL sets up the point-to-point channel of a pair of ranks lazily, at the first send/recv between them
        (tens of milliseconds each).  The swap partners are random (server.py:321), so without this every swap of the
        first dozens of iterations would pay for a new pair inside the training loop (measured: 240 ms per iteration at
        8 GPUs with a swap every iteration).  n-1 rounds of a shifted exchange touch every ordered pair once."""
        if dist.get_backend(self.swap_group) != "nccl" or not torch.cuda.is_available():
            return
        dev = torch.device("cuda", torch.cuda.current_device())
        tx, rx = torch.zeros(8, device=dev), torch.zeros(8, device=dev)
        for off in range(1, self.n_procs):
            to, frm = (self.proc + off) % self.n_procs, (self.proc - off) % self.n_procs
            ops = [dist.P2POp(dist.isend, tx, to, group=self.swap_group), dist.P2POp(dist.irecv, rx, frm, group=self.swap_group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        torch.cuda.synchronize(dev)

    # ---- C4
    def broadcast_fakes(self, X: torch.Tensor) -> None:
        if self.n_procs > 1:
            dist.broadcast(X, src=0)

    # ---- C3
    def reduce_feedback(self, S: torch.Tensor) -> None:
        if self.n_procs > 1:
            dist.reduce(S, dst=0, op=dist.ReduceOp.SUM)

    # ---- C5
    def broadcast_pairs(self, pairs: Optional[torch.Tensor], device: torch.device) -> torch.Tensor:
        """pairs: [N/2, 2] int32 host tensor on process 0 (None elsewhere); returns it on every process (host)."""
        if self.n_procs == 1:
            return pairs
        buf = torch.zeros((self.n_workers // 2, 2), dtype=torch.int32)
        if self.proc == 0:
            buf.copy_(pairs)
        dist.broadcast(buf, src=0, group=self.ctl_group)
        return buf

    def _staging(self, n: int, f32: torch.Tensor, i64: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Receive / scratch buffers of worker n's swap (the size of its flat state), kept across swaps: the swap is
        on the loop's critical path every `swap_interval` iterations and must not go through the allocator."""
        st = self._swap_staging.get(n)
        if st is None or st[0].shape != f32.shape or st[1].shape != i64.shape or st[0].device != f32.device:
            st = (torch.empty_like(f32), torch.empty_like(i64))
            self._swap_staging[n] = st
        return st

    # ---- C6
    def swap_states(self, local_states: Dict[int, Tuple[torch.Tensor, torch.Tensor]], pairs: torch.Tensor) -> List[int]:
        """local_states: {0-based worker index -> (state_f32, state_i64)} of the workers hosted here.
        Exchanges every hosted worker's state with its partner's, in place.  Returns the hosted workers whose state
        changed (they must rebuild their packed weights)."""
        partners = routing.partners_from_pairs(pairs)  # 1-based ranks
        changed: List[int] = []
        ops, copies = [], []
        for n in sorted(local_states):
            pn = partners[n + 1] - 1
            f32, i64 = local_states[n]
            if pn in local_states:
                if n < pn:  # both hosted here: swap through the staging pair
                    of32, oi64 = local_states[pn]
                    tf, ti = self._staging(n, f32, i64)
                    tf.copy_(f32)
                    ti.copy_(i64)
                    f32.copy_(of32)
                    i64.copy_(oi64)
                    of32.copy_(tf)
                    oi64.copy_(ti)
                changed.append(n)
                continue
            peer = routing.process_of_worker(pn, self.n_procs, self.n_workers)
            rf, ri = self._staging(n, f32, i64)
            # message order between two processes: by the SENDING worker's index on both sides
            ops.append((n, dist.P2POp(dist.isend, f32, peer, group=self.swap_group)))
            ops.append((n, dist.P2POp(dist.isend, i64, peer, group=self.swap_group)))
            ops.append((pn, dist.P2POp(dist.irecv, rf, peer, group=self.swap_group)))
            ops.append((pn, dist.P2POp(dist.irecv, ri, peer, group=self.swap_group)))
            copies.append((f32, rf, i64, ri))
            changed.append(n)
        if ops:
            sends = [op for key, op in sorted(((k, o) for k, o in ops if o.op is dist.isend), key=lambda t: t[0])]
            recvs = [op for key, op in sorted(((k, o) for k, o in ops if o.op is dist.irecv), key=lambda t: t[0])]
            for req in dist.batch_isend_irecv(sends + recvs):
                req.wait()
            for f32, rf, i64, ri in copies:
                f32.copy_(rf)
                i64.copy_(ri)
        return changed


class PeerExchange(Exchange):
    """C4 and C3 over peer-mapped memory (NVLink / NVSwitch) instead of collectives; C5 / C6 as in `Exchange`.

    Every process allocates the same three buffers in torch symmetric memory (peer-mapped at rendezvous):
      X     [k*b, C, H, W]   its copy of the generated batches; process 0 PUSHES X into all copies with one kernel of
                             plain stores to the peers' addresses (mdgan_peer_push);
      F     [N, b, C, H, W]  feedback slices, used on process 0 only: worker n's last data-gradient kernel stores
                             dBCE/dX_g STRAIGHT into process 0's F[n] over NVLink (no staging, no reduce); the generator's
                             tanh backward sums the slices of each generated batch in ascending worker order
                             (mdgan_tanh_backward_slices), which is bit-identical to the single-process accumulation;
      flags [64] int32       epoch flags: [0] "X is there" (written by process 0 into every peer), [p] "process p's
                             feedback is there" (written by p into process 0).  Release / acquire at system scope.
    Per iteration a process launches two tiny flag kernels instead of two collectives, and the whole iteration --
    exchange included -- is graph-capturable with no communicator inside the graph.
    Ordering argument (no further synchronisation is needed): process 0 overwrites the peers' X for iteration i+1 only
    after it consumed every F flag of iteration i, which each worker raises after its last read of X(i); a worker
    overwrites F[n] for iteration i+1 only after it saw X(i+1), which process 0 produces after reading F(i)."""

    mode = "peer"

    def __init__(self, proc: int, n_procs: int, n_workers: int, device: torch.device, k: int, b: int, image_shape):
        super().__init__(proc, n_procs, n_workers)
        import torch.distributed._symmetric_memory as symm

        if n_procs > 63:
            raise RuntimeError("PeerExchange: at most 63 processes (one flag word each)")
        group = dist.group.WORLD.group_name
        f = dict(dtype=torch.float32, device=device)
        self.X = symm.empty((k * b, *image_shape), **f)
        self.F = symm.empty((n_workers, b, *image_shape), **f)
        self.flags = symm.empty((64,), dtype=torch.int32, device=device)
        hx, hf, hg = (symm.rendezvous(t, group) for t in (self.X, self.F, self.flags))
        self.X.zero_()
        self.F.zero_()
        self.flags.zero_()
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.F_root = self.F if proc == 0 else hf.get_buffer(0, tuple(self.F.shape), torch.float32)
        self._handles = (hx, hf, hg)
        # NVSwitch multicast mapping of X (one multimem.st per element instead of one store per peer), when the
        # platform provides it and MDGAN_PEER_MULTICAST != 0
        import os

        self.X_mc = None
        self.push_mode = "peer stores (one per destination)"
        mc_ptr = int(getattr(hx, "multicast_ptr", 0) or 0)
        if proc == 0 and mc_ptr != 0 and os.environ.get("MDGAN_PEER_MULTICAST", "1") != "0":
            self.X_mc = _tensor_at(mc_ptr, tuple(self.X.shape), device)
            self.push_mode = "NVSwitch multicast (multimem.st)"
        if proc == 0:
            self.x_addrs = torch.tensor([int(p) for p in hx.buffer_ptrs], dtype=torch.int64).to(device)
            self.sig_addrs = torch.tensor([int(hg.buffer_ptrs[p]) for p in range(1, n_procs)] or [0],
                                          dtype=torch.int64).to(device)
            self.wait_flags, self.n_wait, self.n_sig = self.flags[1:n_procs], n_procs - 1, n_procs - 1
        else:
            self.x_addrs = None
            self.sig_addrs = torch.tensor([int(hg.buffer_ptrs[0]) + 4 * proc], dtype=torch.int64).to(device)
            self.wait_flags, self.n_wait, self.n_sig = self.flags[0:1], 1, 1
        torch.cuda.synchronize(device)
        dist.barrier()  # nobody raises a flag before every process has cleared its own

    def feedback_slice(self, n: int) -> torch.Tensor:
        """Where worker n's feedback goes: process 0's F[n] (a peer-mapped tensor on the other processes)."""
        return self.F_root[n]

    # ---- C4: process 0 pushes X into every copy and raises the peers' flag; the others wait for it
    def broadcast_fakes(self, X: torch.Tensor) -> None:
        from . import ops

        if self.proc == 0:
            if self.X_mc is not None:
                ops.peer_push_multicast(X, self.X_mc)
            else:
                ops.peer_push(X, self.x_addrs, self.n_procs)
            if self.n_sig:
                ops.peer_signal(self.sig_addrs, self.n_sig, self.epoch, advance=False)
        else:
            ops.peer_wait(self.wait_flags, self.n_wait, self.epoch, False, self.err)

    # ---- C3: the feedback is already in process 0's memory; only the flags move
    def reduce_feedback(self, S: Optional[torch.Tensor] = None) -> None:
        from . import ops

        if self.proc == 0:
            ops.peer_wait(self.wait_flags, self.n_wait, self.epoch, True, self.err)
        else:
            ops.peer_signal(self.sig_addrs, self.n_sig, self.epoch, advance=True)

    def check(self) -> None:
        """Raise if a flag wait timed out (call at a point that synchronises anyway)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("peer exchange: a flag did not arrive within the time-out (a peer process died?)")


class _RawCuda:
    """An fp32 device buffer at a raw address (not owned) for torch.as_tensor: the multicast mapping has no tensor."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _tensor_at(ptr: int, shape, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, shape), device=device)


def make_exchange(proc: int, n_procs: int, n_workers: int, device: torch.device, k: int, b: int, image_shape) -> Exchange:
    """MDGAN_EXCHANGE = peer | nccl | auto (default).  auto: peer-memory exchange when every process can set it up,
    otherwise NCCL collectives (announced on stderr; both are device paths, there is no host fallback)."""
    import os
    import sys

    want = os.environ.get("MDGAN_EXCHANGE", "auto").lower()
    if n_procs == 1 or device.type != "cuda" or want == "nccl":
        return Exchange(proc, n_procs, n_workers)
    ex, why = None, ""
    vote_group = _ctl_group()  # collective: created by every process before anything can fail
    try:
        ex = PeerExchange(proc, n_procs, n_workers, device, k, b, image_shape)
    except Exception as e:  # noqa: BLE001 -- any set-up failure means "no peer memory here"
        why = repr(e)
    ok = torch.tensor([1 if ex is not None else 0])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=vote_group)
    if int(ok.item()) == 1:
        return ex
    if want == "peer":
        raise RuntimeError(f"MDGAN_EXCHANGE=peer but the peer-memory exchange could not be set up: {why or 'on another process'}")
    print(f"[mdgan_b200] process {proc}: peer-memory exchange unavailable ({why or 'another process failed'}); "
          "using NCCL collectives", file=sys.stderr, flush=True)
    return Exchange(proc, n_procs, n_workers)


_CTL = None


def _ctl_group():
    global _CTL
    if _CTL is None:
        _CTL = dist.new_group(backend="gloo")
    return _CTL
