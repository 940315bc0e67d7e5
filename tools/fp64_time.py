import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B, M, L = 512, 4096, 256
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev).double()
T = torch.eye(2, dtype=torch.complex128, device=dev)[None].expand(B, -1, -1).contiguous()
tc = uq.target_coeffs(T, torch.float64)
Fsum = torch.empty(B, device=dev, dtype=torch.float64); G = torch.empty(B, L, 2, device=dev, dtype=torch.float64)
for fl in (0, 4):
    for i in range(2): ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, i, None, None, Fsum, G, fl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, i, None, None, Fsum, G, fl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"fp64 flags={fl}: {ms:.3f} ms {B*M*L/ms/1e6:.1f} Gprop/s Fsum0={Fsum[0].item():.12f} G00={G[0,0,0].item():.12e}")
