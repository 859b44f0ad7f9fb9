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
    def __init__(self, proc: int, n_procs: int, n_workers: int):
        self.proc, self.n_procs, self.n_workers = proc, n_procs, n_workers
        if n_procs > 1 and not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before building a multi-process Exchange")
        # The per-iteration collectives (C3, C4) run on the default group and may be captured into a CUDA graph; the
        # rare, host-driven swap traffic gets its own communicators so that eager operations never interleave with
        # graph-captured ones on one NCCL communicator: a host-side (gloo) group for the pair table, which is host
        # data anyway, and a second device group for the state exchange.
        self.ctl_group = self.swap_group = None
        if n_procs > 1:
            self.ctl_group = dist.new_group(backend="gloo")
            self.swap_group = dist.new_group()

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
                if n < pn:  # both hosted here: swap through clones
                    of32, oi64 = local_states[pn]
                    tf, ti = f32.clone(), i64.clone()
                    f32.copy_(of32)
                    i64.copy_(oi64)
                    of32.copy_(tf)
                    oi64.copy_(ti)
                changed.append(n)
                continue
            peer = routing.process_of_worker(pn, self.n_procs, self.n_workers)
            rf, ri = torch.empty_like(f32), torch.empty_like(i64)
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
