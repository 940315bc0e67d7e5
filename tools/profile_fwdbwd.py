#!/usr/bin/env python
"""Small fixed launch sequence of the fused SU(2) kernel for ncu (one GPU):
    python tools/profile_fwdbwd.py [B] [M] [L] [flags] [f32|f64]      -> 2 warm-up + 3 profiled launches"""
import sys, os, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(sys.argv[3]) if len(sys.argv) > 3 else 256
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rdt = torch.float64 if (len(sys.argv) > 5 and sys.argv[5] == "f64") else torch.float32
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev, rdt)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)
tc = uq.target_coeffs(T, rdt)
Fsum = torch.empty(B, device=dev, dtype=rdt)
G = torch.empty(B, L, 2, device=dev, dtype=rdt)
for i in range(5):
    ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, i, None, None, Fsum, G, flags)
torch.cuda.synchronize()
print("ok", Fsum[:2].tolist())
