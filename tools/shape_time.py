#!/usr/bin/env python
"""Time the fused SU(2) step of the build selected by UQOC_LIB on the BASELINE / shipped shapes and check its FP32
accuracy against the FP64 kernel:   shape_time.py [acc] [shape ...]      (shape = name or B,M,L)

Per shape two figures: `fwdbwd` = uqoc_su2_fwdbwd + uqoc_loss_finalize (the multi-launch structure), `step` =
uqoc_su2_fwdbwd_loss (whatever the library fuses).  CUDA events around 20 back-to-back steps after warm-up."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import universal_quantum_optimal_control_b200 as uq  # noqa: E402
from universal_quantum_optimal_control_b200 import ops  # noqa: E402

SHAPES = {"c5": (4096, 4096, 256, (0.1, 0.5)), "c3": (1, 65536, 256, (0.035, 0.07)), "score": (200, 1000, 100, (0.1, 0.5)),
          "grape": (100, 1000, 400, (0.035, 0.07)), "c1": (4, 256, 16, (0.1, 0.5)), "b8": (8, 65536, 256, (0.1, 0.5)),
          "b32": (32, 8192, 256, (0.1, 0.5))}
dev = torch.device("cuda", 0)
tag = os.path.basename(os.environ.get("UQOC_LIB", "default"))


def workload(B, M, L, tau, dtype=torch.float32):
    g = torch.Generator().manual_seed(0)
    pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, tau[0] + (tau[1] - tau[0]) * torch.rand(B, L, generator=g)], -1)
    X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
    T = torch.matrix_exp(-1j * X[None] * (torch.rand(B, generator=g) * math.pi)[:, None, None]).to(dev)
    return pulses.to(dev, dtype), uq.target_coeffs(T, dtype)


def timed(fn, n=20):
    for i in range(5):
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
    global args_all
    args = sys.argv[1:]
    args_all = list(args)
    flags = 0
    for a in list(args):
        if a.startswith("flags="):
            flags = int(a[6:], 0)
            args.remove(a)
    names = [a for a in args if a not in ("acc", "fwd")] or ["c5", "c3", "score", "grape"]
    for nm in names:
        B, M, L, tau = SHAPES[nm] if nm in SHAPES else (*map(int, nm.split(",")), (0.1, 0.5))
        pulses, tc = workload(B, M, L, tau)
        buf = torch.empty(B + B * L * 2, device=dev)
        G, Fsum = buf[:B * L * 2], buf[B * L * 2:]
        lo = torch.empty(3, device=dev)

        def f_multi(i):
            ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, i, None, None, Fsum, G, flags)
            ops._finalize(Fsum, B * M, "sharp", 0.99, 100, G)

        def f_step(i):
            ops._launch_fwdbwd_loss(pulses, tc, None, M, (1.0, 0.05), 7, i, "sharp", 0.99, 100, None, None, Fsum, G, lo, flags)

        def f_fwd(i):
            ops._launch_forward(pulses, tc, None, M, 0, (1.0, 0.05), 7, i, None, None, None, Fsum, flags)

        ms_a, ms_b = timed(f_multi), timed(f_step)
        props = B * M * L
        if "fwd" in args_all:
            print(f"{tag:14s} {nm:6s} forward only: {timed(f_fwd) * 1e3:9.2f} us", flush=True)
        print(f"{tag:14s} {nm:6s} B={B} M={M} L={L} flags={flags:#x}: fwdbwd+finalize {ms_a * 1e3:9.2f} us {props / ms_a / 1e6:7.1f} Gprop/s | "
              f"step {ms_b * 1e3:9.2f} us {props / ms_b / 1e6:7.1f} Gprop/s {props * 116 / ms_b / 1e9 / 74.45 * 100:5.1f}% peak  "
              f"loss={lo[0].item():.6f}", flush=True)
    if "acc" in args:
        for L, tau in ((16, (0.1, 0.5)), (100, (0.1, 0.5)), (256, (0.1, 0.5)), (256, (0.035, 0.07)), (400, (0.1, 0.5))):
            B, M = 8, 8192
            p64, tc64 = workload(B, M, L, tau, torch.float64)
            err = uq.philox_errors(B, M, (1.0, 0.05), 3, 0, dtype=torch.float64)
            res = {}
            for dn, dtype, fl in (("f64", torch.float64, 0), ("f32", torch.float32, flags), ("f32_wps4", torch.float32, flags | 16),
                                  ("f32_st2", torch.float32, flags | (2 << 8))):
                F = torch.empty(B * M, dtype=dtype, device=dev)
                Fsum = torch.empty(B, dtype=dtype, device=dev)
                G = torch.empty(B, L, 2, dtype=dtype, device=dev)
                ops._launch_fwdbwd(p64.to(dtype), tc64.to(dtype), err.to(dtype), None, M, 0, (1.0, 0.05), 0, 0, F, None, Fsum, G, fl)
                res[dn] = (F.double(), G.double())
            for dn in ("f32", "f32_wps4", "f32_st2"):
                dF = (res[dn][0] - res["f64"][0]).abs().max().item()
                dG = ((res[dn][1] - res["f64"][1]).abs().max() / res["f64"][1].abs().max()).item()
                print(f"{tag:14s} acc L={L} tau={tau} {dn:9s}: max|dF|={dF:.2e} rel|dG|={dG:.2e}", flush=True)


if __name__ == "__main__":
    main()
