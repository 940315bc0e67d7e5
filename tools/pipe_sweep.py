#!/usr/bin/env python
"""End-to-end (pinned host buffers in and out) and device-only time of PipelinedStep at the config-5 slice for several
chunkings:   python tools/pipe_sweep.py      or      torchrun --nproc-per-node N tools/pipe_sweep.py"""
import os, sys, time, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
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
pd, Td = ph.to(dev), Th.to(dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for chunks, ramp in ((1, False), (2, False), (4, False), (4, True), (6, True), (8, True), ("auto", True)):
    pipe = uq.PipelinedStep(B, L, M, chunks=chunks, sigma=(1.0, 0.05), seed=1, device=dev, ramp=ramp, group=group)
    n = 8
    for i in range(3):
        pipe(ph, Th, offset=i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n):
        pipe(ph, Th, offset=i)
    barrier()
    e2e = (time.perf_counter() - t0) / n * 1e3
    for i in range(3):
        pipe.run_device(pd, Td, offset=i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        pipe.run_device(pd, Td, offset=i)
    e1.record()
    barrier()
    devms = e0.elapsed_time(e1) / n
    if rank == 0:
        print(f"world={world} chunks={chunks} ramp={ramp} sizes={[b1 - b0 for b0, b1 in pipe.bounds]}: e2e {e2e:.3f} ms/step  device {devms:.3f} ms/step", flush=True)
if world > 1:
    dist.destroy_process_group()
