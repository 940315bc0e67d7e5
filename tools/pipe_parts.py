#!/usr/bin/env python
"""Which part of the host-buffer step costs what at N ranks: PipelinedStep with / without the host->device and
device->host copies, every variant synchronised per step.   torchrun --nproc-per-node N tools/pipe_parts.py"""
import os, sys, time, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
group = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
B, L, M = 4096, 256, 4096 * world
g = torch.Generator().manual_seed(0)
ph = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).pin_memory()
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
Th = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).contiguous().pin_memory()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


pipe = uq.PipelinedStep(B, L, M, sigma=(1.0, 0.05), seed=1, device=dev, group=group)
pipe(ph, Th, offset=0)
tr = ops.raw_target(pipe.d_target, torch.float32)
res = {}
for name, h2d, d2h in (("no copies", False, False), ("H2D only", True, False), ("D2H only", False, True), ("both", True, True)):
    def step(i):
        pipe._src_pulses, pipe._src_target = ph, Th
        pipe._enqueue(pipe.d_pulses, tr, i, h2d=h2d, d2h=d2h)
        torch.cuda.current_stream(dev).synchronize()
    for i in range(3):
        step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(10):
        step(i)
    barrier()
    res[name] = (time.perf_counter() - t0) / 10 * 1e3
# plain copies alone, all ranks at once
d = torch.empty_like(ph, device=dev)
hb = torch.empty_like(ph).pin_memory()
for name, fn in (("H2D 8.4 MB alone", lambda: d.copy_(ph, non_blocking=True)), ("D2H 8.4 MB alone", lambda: hb.copy_(d, non_blocking=True))):
    barrier()
    t0 = time.perf_counter()
    for i in range(10):
        fn()
        torch.cuda.current_stream(dev).synchronize()
    barrier()
    res[name] = (time.perf_counter() - t0) / 10 * 1e3
if rank == 0:
    for k, v in res.items():
        print(f"world={world} {k:20s} {v:.3f} ms/step", flush=True)
if world > 1:
    dist.destroy_process_group()
