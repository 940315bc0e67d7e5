#!/usr/bin/env python
"""Time the fused SU(2) launch of the build selected by UQOC_LIB:  variant_time.py [B] [M] [L] [flags]"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(sys.argv[3]) if len(sys.argv) > 3 else 256
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)
tc = uq.target_coeffs(T, torch.float32)
Fsum = torch.empty(B, device=dev)
G = torch.empty(B, L, 2, device=dev)
def go(i):
    ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, i, None, None, Fsum, G, flags)
for i in range(3):
    go(i)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    go(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{os.environ.get('UQOC_LIB', 'default'):60s} B={B} M={M} L={L} flags={flags}: {ms:.4f} ms  {B * M * L / ms / 1e6:.1f} Gprop/s  {B * M * L * 116 / ms / 1e9 / 74.45 * 100:.1f}% peak  Fsum0={Fsum[0].item():.4f}")
