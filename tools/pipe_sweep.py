#!/usr/bin/env python
"""End-to-end (pinned host buffers in and out) time of PipelinedStep at the config-5 slice for several chunkings."""
import os, sys, time, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
B, L, M = 4096, 256, 4096
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ph = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).pin_memory()
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
Th = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).contiguous().pin_memory()
for chunks, ramp in ((1, False), (2, False), (3, False), (4, False), (4, True), (5, True), (6, True), (6, False), (8, True), (8, False)):
    pipe = uq.PipelinedStep(B, L, M, chunks=chunks, sigma=(1.0, 0.05), seed=1, device=dev, ramp=ramp)
    for i in range(3):
        pipe(ph, Th, offset=i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 8
    for i in range(n):
        pipe(ph, Th, offset=i)
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"chunks={chunks} ramp={ramp} sizes={[b1 - b0 for b0, b1 in pipe.bounds]}: {dt:.3f} ms/step", flush=True)
