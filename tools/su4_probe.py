#!/usr/bin/env python
"""SU(4) kernels: time eigenframe vs Pade kernel and report the FP32 error against the FP64 kernel.
    python tools/su4_probe.py [B] [L] [M]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pulses = torch.stack([(torch.rand(B, L) * 2 - 1) * 3.15, (torch.rand(B, L) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L)], -1).to(dev)
T = torch.diag(torch.tensor([1, 1, 1, -1], dtype=torch.complex64)).to(dev)[None].expand(B, -1, -1)
err = uq.philox_errors_su4(B, M, (1.0, 0.05), seed=5)


def run(dtype, flags, with_F=True):
    p = pulses.to(dtype)
    tgt = ops._su4_target(T, dtype, B)
    F = torch.empty(B * M, dtype=dtype, device=dev) if with_F else None
    Fsum = torch.empty(B, dtype=dtype, device=dev)
    G = torch.empty(B, L, 3, dtype=dtype, device=dev)
    e = err.to(dtype)
    def go():
        ops._su4_launch(True, p, tgt, e, None, M, 0, 1.0, (1.0, 0.05), 0, 0, None, F, None, Fsum, G, flags)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n):
        go()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n, F, G


ms64, F64, G64 = run(torch.float64, 0)
ms64p, F64p, G64p = run(torch.float64, 64)
print(f"fp64 eig {ms64:.3f} ms  pade {ms64p:.3f} ms  |dF| {(F64 - F64p).abs().max().item():.2e} "
      f"rel dG {((G64 - G64p).abs().max() / G64.abs().max()).item():.2e}")
for name, fl in (("eig", 0), ("pade", 64)):
    ms, F, G = run(torch.float32, fl)
    print(f"fp32 {name}: {ms:.3f} ms  {B * M * L / ms / 1e6:.3f} G su4-prop/s  |dF| {(F.double() - F64).abs().max().item():.2e} "
          f"rel dG {((G.double() - G64).abs().max() / G64.abs().max()).item():.2e}")
