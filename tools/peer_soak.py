#!/usr/bin/env python
"""torchrun soak test of the NVLink peer-memory exchange: N steps with changing pulses and mixed shapes, every step
checked against the NCCL path (same shard results, different summation order) and for bit-identity across ranks."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
shapes = [(1, 256, 8192), (7, 33, 1000), (200, 100, 125), (3, 64, 40000)]
g = torch.Generator().manual_seed(0)
worst, mism = 0.0, 0
for (B, L, M) in shapes:
    base = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
    tc = torch.zeros(B, 8, device=dev); tc[:, 0] = 2.0
    px = uq.PeerExchange(dist.group.WORLD, B, L, 2, torch.float32, dev)
    bufA = torch.empty(B + B * L * 2, device=dev); bufB = torch.empty_like(bufA)
    for i in range(steps):
        pulses = base + 1e-3 * i
        ops._launch_fwdbwd_peer(pulses, tc, None, M, rank * M, (1.0, 0.05), 7, i, None, None, bufA[:B], bufA[B:], 0, px)
        if i % 10 == 0:
            ops._launch_fwdbwd(pulses, tc, None, None, M, rank * M, (1.0, 0.05), 7, i, None, None, bufB[:B], bufB[B:], 0)
            dist.all_reduce(bufB)
            d = ((bufA - bufB).abs().max() / bufB.abs().max()).item()
            worst = max(worst, d)
            ref = bufA.clone(); dist.broadcast(ref, 0)
            if not torch.equal(ref, bufA):
                mism += 1
    torch.cuda.synchronize()
    if rank == 0:
        print(f"shape {(B, L, M)}: {steps} peer steps ok, worst rel diff vs NCCL {worst:.2e}, rank mismatches {mism}", flush=True)
t = torch.tensor([worst, float(mism)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("SOAK", "PASS" if (t[0].item() < 1e-5 and t[1].item() == 0 and not torch.isnan(t).any()) else "FAIL", t.tolist(), flush=True)
dist.destroy_process_group()
