"""Minimal, race-free stand-in for `tensordict.TensorDict` (test infrastructure only).

The reference uses TensorDict purely as a container to move a discriminator
`state_dict` between two workers one leaf tensor at a time
(/root/reference/src/actors/worker.py:252-282): `TensorDict(sd, batch_size=[])
.unflatten_keys(".")`, then `.irecv(src=, return_premature=True)` / `.send(dst=)`
and finally `.flatten_keys(".")` into `load_state_dict`.

The real package wraps the *live* parameter tensors, so the reference's irecv
(posted first) overwrites storage that its own send is still reading
("race detection").  This stand-in snapshots the tensors at
construction, which yields the intended exchange (each worker receives its
partner's pre-swap state) deterministically.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


class TensorDict:
    def __init__(self, source, batch_size=None):
        self._flat = OrderedDict((k, v.detach().clone()) for k, v in source.items())

    def unflatten_keys(self, sep="."):
        return self

    def flatten_keys(self, sep="."):
        return self._flat

    def keys(self):
        return self._flat.keys()

    def items(self):
        return self._flat.items()

    def irecv(self, src, return_premature=False, init_tag=0):
        reqs = []
        for i, (_, t) in enumerate(self._flat.items()):
            reqs.append(dist.irecv(t, src=src, tag=init_tag + i + 1))
        if return_premature:
            return reqs
        for r in reqs:
            r.wait()
        return None

    def send(self, dst, init_tag=0):
        for i, (_, t) in enumerate(self._flat.items()):
            dist.send(t.contiguous(), dst=dst, tag=init_tag + i + 1)
