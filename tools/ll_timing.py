#!/usr/bin/env python
"""Where the multi-GPU small-vector step spends its time (timing build, -DUQOC_LL_TIMING; UQOC_LIB selects it):
    tools/variants.sh llt:"-DUQOC_LL_TIMING=1"
    UQOC_LIB=.../lib/variants/llt.so torchrun --nproc-per-node N tools/ll_timing.py
globaltimer stamps: main kernel block 0 start [6], last block end [7]; exchange kernel block 0: entry [0], after
griddepcontrol.wait [1], after the pushes [2], after the column polls [3], after the loss [4], end [5]."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
B, M, L = 1, 65536, 256
g = torch.Generator().manual_seed(0)
pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.035 + 0.035 * torch.rand(B, L, generator=g)], -1).to(dev)
tc = torch.zeros(B, 8, device=dev); tc[:, 0] = 2.0
buf = torch.empty(B + B * L * 2, device=dev)
G, Fsum = buf[:B * L * 2], buf[B * L * 2:]
lo = torch.empty(3, device=dev)
px = uq.PeerExchange(dist.group.WORLD, B, L, 2, torch.float32, dev)
ws = ops.su2_workspace(B, L, M, torch.float32, 0, dev)
rows = []
BACK_TO_BACK = len(sys.argv) > 1 and sys.argv[1] == "b2b"      # stamps of the LAST of 30 queued steps (steady state)
for rep in range(40):
    for i in range(30 if BACK_TO_BACK else 1):
        ops._launch_fwdbwd_peer_loss(pulses, tc, None, M, rank * M, M * world, (1.0, 0.05), 7, rep * 30 + i, "sharp", 0.99, 100, None,
                                     None, Fsum, G, lo, 0, px, ws=ws)
        if BACK_TO_BACK and i == 28:
            ws[128:192].zero_()          # stream-ordered: clears the atomicMax slot before the last step
    torch.cuda.synchronize()
    if rep >= 10:
        rows.append(ws[128:192].view(torch.int64).cpu().tolist())
    ws[128:192].zero_()
    torch.cuda.synchronize()
import statistics as S
def med(f):
    return S.median(f(r) for r in rows) / 1e3
out = {"main kernel (block 0 start -> last block end)": med(lambda r: r[7] - r[6]),
       "last block end -> exchange kernel past griddepcontrol.wait": med(lambda r: r[1] - r[7]),
       "reduce partial rows + push": med(lambda r: r[2] - r[1]),
       "column polls + rank-order sum": med(lambda r: r[3] - r[2]),
       "fidelity words + loss": med(lambda r: r[4] - r[3]),
       "scale + write": med(lambda r: r[5] - r[4]),
       "total": med(lambda r: r[5] - r[6])}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    for k in out:
        print(f"{k:62s} " + " ".join(f"{g_[k]:6.2f}" for g_ in gathered) + "  us (per rank)")
dist.destroy_process_group()
