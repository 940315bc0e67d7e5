#!/usr/bin/env python
"""Sweep the launch-shape overrides (samples per thread x warps per sample group x sample-tile splits) of the fused
SU(2) step on given problem sizes and print them against the library heuristic (``make_plan``, csrc/uqoc_api.cu):

    plan_sweep.py [B,M,L ...]        default: the reference's shipped step shapes

One process, CUDA events around 20 back-to-back ``uqoc_su2_fwdbwd_loss`` steps per candidate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq  # noqa: E402
from universal_quantum_optimal_control_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
SHAPES = [(200, 1000, 100), (100, 1000, 400), (200, 1000, 400), (1000, 1000, 100), (32, 8192, 256), (4, 256, 16)]


def timed(fn, n=20):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    shapes = [tuple(map(int, a.split(","))) for a in sys.argv[1:]] or SHAPES
    for B, M, L in shapes:
        g = torch.Generator().manual_seed(0)
        pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
        tc = torch.zeros(B, 8, device=dev)
        tc[:, 0] = 2.0
        Fsum, G, lo = torch.empty(B, device=dev), torch.empty(B, L, 2, device=dev), torch.empty(3, device=dev)
        rows = []
        cands = [("default", 0)]
        for st in (2, 4):
            for wps in (1, 4):
                for sp in (0, 1, 2, 3, 4, 8):
                    cands.append((f"st={st} wps={wps} splits={sp or 'auto'}", ops.tuning_flags(st=st, wps=wps, splits=sp)))
        ref = None
        for name, fl in cands:
            def step(i, fl=fl):
                ops._launch_fwdbwd_loss(pulses, tc, None, M, (1.0, 0.05), 7, i, "sharp", 0.99, 100, None, None, Fsum, G, lo, fl)
            try:
                ms = timed(step)
            except Exception as e:                            # a shape the override does not support
                rows.append((float("inf"), name, f"rejected: {str(e)[:60]}"))
                continue
            loss = lo[0].item()
            if ref is None:
                ref = loss
            rows.append((ms, name, f"loss={loss:.6f}" + ("" if abs(loss - ref) < 1e-4 * abs(ref) else "  LOSS MISMATCH")))
        props = B * M * L
        print(f"--- B={B} M={M} L={L}  ({props / 1e6:.1f} Mprop)", flush=True)
        d_ms = rows[0][0]
        for ms, name, note in sorted(rows)[:8] + [rows[0]]:
            print(f"  {name:28s} {ms * 1e3:9.2f} us  {props / ms / 1e6:7.1f} Gprop/s  {props * 116 / ms / 1e9 / 74.45 * 100:5.1f}% peak  "
                  f"{d_ms / ms:5.2f}x default  {note}", flush=True)


if __name__ == "__main__":
    main()
