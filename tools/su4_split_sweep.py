import sys, torch
sys.path.insert(0, "/root/repo")
import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for B, L, M in ((1, 128, 1024), (1, 128, 4096), (1, 128, 8192), (1, 128, 14000), (1, 128, 16384), (1, 128, 32768), (4, 128, 2048), (1, 400, 4096)):
    pulses = torch.stack([(torch.rand(B, L) * 2 - 1) * 3.15, (torch.rand(B, L) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L)], -1).to(dev)
    T = torch.diag(torch.tensor([1, 1, 1, -1], dtype=torch.complex64)).to(dev)[None].expand(B, -1, -1)
    tgt = ops._su4_target(T, torch.float32, B)
    buf = torch.empty(B + B * L * 3, device=dev)
    res = []
    for wps in (0, 1, 4):
        fl = uq.tuning_flags(wps=wps)
        def go():
            ops._su4_launch(True, pulses, tgt, None, None, M, 0, 1.0, (1.0, 0.05), 7, 0, None, None, None, buf[:B], buf[B:], fl)
        for _ in range(3): go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): go()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(f"B={B} L={L} M={M}: default {res[0]:8.1f} us   one-sample-per-thread {res[1]:8.1f} us   train split {res[2]:8.1f} us", flush=True)
