#!/usr/bin/env python
"""Breakdown of the in-kernel epilogue (build with -DUQOC_FIN_TIMING, select with UQOC_LIB): globaltimer stamps of the
LAST block written next to the ticket counter.  fin_timing.py [B M L]"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B, M, L = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 65536, 256)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.035 + 0.035 * torch.rand(B, L, generator=g)], -1).to(dev)
tc = torch.zeros(B, 8, device=dev); tc[:, 0] = 2.0
buf = torch.empty(B * L * 2 + B, device=dev)
lo = torch.empty(3, device=dev)
ws = ops.su2_workspace(B, L, M, torch.float32, 0, dev)
for i in range(8):
    ops._launch_fwdbwd_loss(pulses, tc, None, M, (1.0, 0.05), 7, i, "sharp", 0.99, 100, None, None, buf[B * L * 2:], buf[:B * L * 2], lo, 0, ws=ws)
    torch.cuda.synchronize()
    st = ws[:128].view(torch.int64).cpu().tolist()
    t0 = st[7]
    print(f"run {i}: kernel start -> last block enters epilogue {(st[2]-t0)/1e3:7.2f} us | fence+ticket {(st[3]-st[2])/1e3:6.2f} | "
          f"partials sum {(st[4]-st[3])/1e3:6.2f} (first pass {(st[8]-st[3])/1e3:6.2f}) | loss {(st[5]-st[4])/1e3:6.2f} | store {(st[6]-st[5])/1e3:6.2f} | total {(st[6]-t0)/1e3:7.2f} us  loss={lo[0].item():.5f}")
