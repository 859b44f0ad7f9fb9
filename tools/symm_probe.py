#!/usr/bin/env python
"""Probe (N GPUs, torchrun): does torch symmetric memory work on this box -- peer pointers, multicast, P2P stores from
an ordinary kernel, inside a CUDA graph?"""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
def log(*a):
    print(f"[rank {rank}]", *a, flush=True)
try:
    t = symm.empty((1024,), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    log("rendezvous ok; multicast:", symm._SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, dev.index), "mc_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else 0,
        "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pad", hdl.signal_pad_size)
    t.fill_(float(rank))
    hdl.barrier()
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (1024,), torch.float32)
    log("peer buffer reads", pb[:2].tolist(), "ptr", hex(pb.data_ptr()))
    hdl.barrier()
    pb[512:].fill_(100.0 + rank)           # ordinary kernel storing into the peer's memory
    hdl.barrier()
    log("local after peer store", t[510:514].tolist())
    g = torch.cuda.CUDAGraph()
    src = torch.full((512,), 7.0 + rank, device=dev)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        pb[:512].copy_(src)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); hdl.barrier()
    log("local after graph-replayed peer copy", t[:2].tolist())
except Exception as e:
    import traceback; traceback.print_exc(); log("FAILED", repr(e))
dist.barrier(); dist.destroy_process_group()
