#!/usr/bin/env python
"""Time the strict reference-signature path (per-sample pulse rows, three-call structure
generator -> fidelity -> loss -> backward, trainer.py:80-90) against the fused op on the same step."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def ref_style():
    p = pulses.clone().requires_grad_(True)
    pm = p.repeat_interleave(M, 0)                       # trainer.py:80
    tm = T.repeat_interleave(M, 0)                       # trainer.py:81
    err = uq.get_ore_ple_error_distribution(M * B, 1.0, 0.05).to(dev)   # trainer.py:82
    U = uq.batched_unitary_generator(pm, err)            # trainer.py:84
    loss = uq.sharp_loss(U, tm, uq.fidelity, 1)          # trainer.py:88
    loss.backward()
    return loss

def fused():
    p = pulses.clone().requires_grad_(True)
    loss, _ = uq.fused_propagate_loss(p, T, monte_carlo=M, sigma=(1.0, 0.05), seed=1, offset=0)
    loss.backward()
    return loss

ms_r, ms_f = timed(ref_style), timed(fused)
props = B * M * L
print(f"B={B} M={M} L={L}: reference-signature path {ms_r:.3f} ms ({props/ms_r/1e6:.1f} Gprop/s)   fused op {ms_f:.3f} ms ({props/ms_f/1e6:.1f} Gprop/s)")
