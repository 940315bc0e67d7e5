#!/usr/bin/env python
"""torchrun probe: latency of the fused step with the NCCL exchange vs the NVLink peer-memory exchange.
    torchrun --nproc-per-node N tools/peer_probe.py [B] [M] [L]"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
L = int(sys.argv[3]) if len(sys.argv) > 3 else 256
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.035 + 0.035 * torch.rand(B, L, generator=g)], -1).to(dev)
tc = torch.zeros(B, 8, device=dev); tc[:, 0] = 2.0
buf = torch.empty(B + B * L * 2, device=dev)
G, Fsum = buf[:B * L * 2], buf[B * L * 2:]
px = uq.PeerExchange(dist.group.WORLD, B, L, 2, torch.float32, dev)
M_total = M * world

def step_nccl(i):
    ops._launch_fwdbwd(pulses, tc, None, None, M, rank * M, (1.0, 0.05), 7, i, None, None, Fsum, G, 0)
    dist.all_reduce(buf)
    return ops._finalize(Fsum, B * M_total, "sharp", 0.99, 100, G)

def step_peer(i):
    ops._launch_fwdbwd_peer(pulses, tc, None, M, rank * M, (1.0, 0.05), 7, i, None, None, Fsum, G, 0, px)
    return ops._finalize(Fsum, B * M_total, "sharp", 0.99, 100, G)

lo = torch.empty(3, device=dev)

def step_peer_loss(i):
    ops._launch_fwdbwd_peer_loss(pulses, tc, None, M, rank * M, M_total, (1.0, 0.05), 7, i, "sharp", 0.99, 100, None, None, Fsum, G,
                                 lo, 0, px)
    return lo

def step_local(i):
    ops._launch_fwdbwd(pulses, tc, None, None, M, rank * M, (1.0, 0.05), 7, i, None, None, Fsum, G, 0)
    return ops._finalize(Fsum, B * M_total, "sharp", 0.99, 100, G)

res = {}
for name, fn in (("local(no exchange)", step_local), ("nccl", step_nccl), ("peer", step_peer), ("peer+loss fused", step_peer_loss)):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    # (a) back-to-back throughput
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        out = fn(i)
    e1.record(); torch.cuda.synchronize()
    thr = e0.elapsed_time(e1) / 50 * 1e3
    # (b) aligned single-step latency
    lat = []
    for i in range(20):
        dist.barrier(); torch.cuda.synchronize()
        e0.record(); out = fn(i); e1.record(); torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1) * 1e3)
    lat.sort()
    res[name] = (thr, lat[len(lat) // 2], float(out[0].item()), float(G.abs().sum().item()))
if rank == 0:
    for k, v in res.items():
        print(f"{k:22s} back-to-back {v[0]:7.1f} us/step   aligned median {v[1]:7.1f} us   loss {v[2]:.6f}  sum|G| {v[3]:.6e}", flush=True)
dist.barrier()
dist.destroy_process_group()
