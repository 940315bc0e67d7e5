#!/usr/bin/env python
"""Feasibility probe (2+ GPUs, torchrun): can this box map peer GPU memory into every rank
(torch.distributed._symmetric_memory) -- the plumbing of the fused exchange kernel."""
import os, sys, time
import torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs],
          "pad bytes", hdl.signal_pad_size, "multicast", hdl.has_multicast_support if hasattr(hdl, "has_multicast_support") else None,
          flush=True)
    t.fill_(float(rank + 1))
    dist.barrier(); torch.cuda.synchronize()
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (16,), torch.float32)
    print(rank, "peer", peer, "reads", remote[:4].tolist(), flush=True)
    remote[8:12] = 100.0 + rank                       # peer store
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(rank, "own buffer after peer store", t[8:12].tolist(), flush=True)
except Exception as e:
    print(rank, "SYMM_MEM FAILED", repr(e)[:500], flush=True)
dist.destroy_process_group()
