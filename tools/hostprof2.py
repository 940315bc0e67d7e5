#!/usr/bin/env python
"""Host-side cost of the fused step at the reference's shipped step size (B=200, L=100, M=1000), piece by piece."""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B, M, L = 200, 1000, 100
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)
N = 200


def t(fn, n=N):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


p = pulses.clone().requires_grad_(True)
tc = ops.raw_target(T, torch.float32)
buf = torch.empty(B * L * 2 + B, device=dev)
G, Fsum = buf[:B * L * 2], buf[B * L * 2:]
lo = torch.empty(3, device=dev)
print(f"raw C call (uqoc_su2_fwdbwd_loss, preallocated)   {t(lambda: ops._launch_fwdbwd_loss(pulses, tc, None, M, (1.0, 0.05), 1, 0, 'sharp', 0.99, 100, None, None, Fsum, G, lo, ops.FLAG_RAW_TARGET)):7.1f} us")
print(f"clone + requires_grad                              {t(lambda: pulses.clone().requires_grad_(True)):7.1f} us")
print(f"forward only (no_grad)                             {t(lambda: uq.fused_propagate_loss(pulses, T, monte_carlo=M, seed=1)):7.1f} us")
print(f"forward (grad)                                     {t(lambda: uq.fused_propagate_loss(p, T, monte_carlo=M, seed=1)):7.1f} us")


def fb():
    loss, _ = uq.fused_propagate_loss(p, T, monte_carlo=M, seed=1)
    loss.backward()


print(f"forward + backward (autograd)                      {t(fb):7.1f} us")
print(f"torch.empty                                        {t(lambda: torch.empty(B * L * 2 + B, device=dev)):7.1f} us")
print(f"slice                                              {t(lambda: buf[:100]):7.1f} us")
print(f"Fsum / M                                           {t(lambda: Fsum / M):7.1f} us")
print(f"current_stream().cuda_stream                       {t(lambda: torch.cuda.current_stream(dev).cuda_stream):7.1f} us")
print(f"raw_target                                         {t(lambda: ops.raw_target(T, torch.float32)):7.1f} us")
if hasattr(uq, "FusedStep"):
    st = uq.FusedStep(B, L, M, device=dev)
    print(f"FusedStep (one C call, pre-sized buffers)          {t(lambda: st(pulses, T, offset=1)):7.1f} us")
