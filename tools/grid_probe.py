#!/usr/bin/env python
"""BASELINE config 2 (L=64, 1000 x 1000 grid, forward only): kernel time against the split count."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import sweeps
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
L = 64
pulse = torch.stack([(torch.rand(L, generator=g) * 2 - 1) * math.pi, 0.1 + 0.4 * torch.rand(L, generator=g)], -1).to(dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X * (math.pi / 4)).to(dev)
ore, ple = torch.linspace(-3, 3, 1000).to(dev), torch.linspace(-0.15, 0.15, 1000).to(dev)
for sp in [0] + [int(a) for a in sys.argv[1:]]:
    fl = uq.tuning_flags(splits=sp)
    for _ in range(3): sweeps.fidelity_grid(pulse, T, ore, ple, flags=fl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): F = sweeps.fidelity_grid(pulse, T, ore, ple, flags=fl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"splits={sp:5d}: {ms*1e3:7.1f} us  {L*1e6/ms/1e6:.1f} Gprop/s  F.mean={F.mean().item():.6f}")
