#!/usr/bin/env python
"""On-GPU experiment harness: FMA peak probes, launch-shape sweep, accuracy table.
Writes JSON lines to stdout; run under gpurun and redirect into gpurun_out/."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import universal_quantum_optimal_control_b200 as uq  # noqa: E402
from universal_quantum_optimal_control_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def time_kernel(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e30, 0.0
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = min(best, ms)
        tot += ms
    return best, tot / iters


def workload(B, L, M, dtype, tau=(0.1, 0.5)):
    g = torch.Generator().manual_seed(0)
    phi = (torch.rand(B, L, generator=g) * 2 - 1) * 3.15
    t = tau[0] + (tau[1] - tau[0]) * torch.rand(B, L, generator=g)
    pulses = torch.stack([phi, t], -1).to(dev, dtype).contiguous()
    ang = torch.rand(B, generator=g) * math.pi
    X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
    T = torch.matrix_exp(-1j * X[None] * ang[:, None, None]).to(dev)
    return pulses, uq.target_coeffs(T, dtype)


def main():
    which = sys.argv[1:] or ["peak", "sweep", "acc"]
    if "peak" in which:
        for mode, name in ((0, "ffma"), (2, "ffma2"), (1, "dfma")):
            tf, ms = uq.fp32_peak_tflops(8192, mode)
            emit(probe="peak", mode=name, tflops=tf, ms=ms)
    if "sweep" in which:
        cases = [("curriculum", 4096, 256, 4096, (0.1, 0.5)), ("grape", 1, 256, 65536, (0.035, 0.07)),
                 ("shipped", 200, 100, 1000, (0.1, 0.5))]
        if "small" in which:
            cases += [("c1", 4, 16, 256, (0.1, 0.5)), ("mid", 64, 100, 1000, (0.1, 0.5))]
        only = [w[5:] for w in which if w.startswith("case=")]
        if only:
            cases = [c for c in cases if c[0] in only]
        f32only = "f32only" in which
        for name, B, L, M, tau in cases:
            for dtype, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
                if f32only and dn == "f64":
                    continue
                pulses, tc = workload(B, L, M, dtype, tau)
                Fsum = torch.empty(B, dtype=dtype, device=dev)
                G = torch.empty(B, L, 2, dtype=dtype, device=dev)
                shapes = [(0, 0), (1, 1), (2, 1), (4, 1), (-2, 1), (-4, 1)] if dn == "f32" else [(0, 0), (1, 1), (2, 1)]
                if B * M < 40000 or name == "grape":
                    shapes += [(1, 2), (1, 4), (1, 8), (1, 16)]
                for st, lps in shapes:
                    for fast in ((False, True) if dn == "f32" else (False,)):
                        if dn == "f64" and name == "curriculum" and st == 0:
                            continue
                        flags = uq.tuning_flags(st=abs(st), lps=lps, fast_sincos=fast, no_packed=st < 0)
                        fn = lambda: ops._launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 7, 0, None, None, Fsum, G, flags)
                        it = 5 if (dn == "f64" and name == "curriculum") else 10
                        best, mean = time_kernel(fn, iters=it)
                        props = B * M * L
                        emit(probe="sweep", case=name, B=B, L=L, M=M, dtype=dn, st=st, lps=lps, fast=fast, best_ms=best,
                             mean_ms=mean, gprops=props / best / 1e6, tflops116=props * 116 / best / 1e9)
                # forward only
                for fast in ((False, True) if dn == "f32" else (False,)):
                    flags = uq.tuning_flags(fast_sincos=fast)
                    fn = lambda: ops._launch_forward(pulses, tc, None, M, 0, (1.0, 0.05), 7, 0, None, None, None, Fsum, flags)
                    best, mean = time_kernel(fn, iters=5)
                    emit(probe="sweep_fwd", case=name, dtype=dn, fast=fast, best_ms=best, gprops=B * M * L / best / 1e6)
    if "su4" in which:
        for dtype, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
            for (B, L, M) in ((1, 128, 32768), (8, 128, 32768)):
                g = torch.Generator().manual_seed(0)
                pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, (torch.rand(B, L, generator=g) * 2 - 1) * 3.15,
                                      0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev, dtype)
                cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
                tgt = ops._su4_target(torch.diag(torch.tensor([1, 1, 1, -1], dtype=cdt)).to(dev)[None].expand(B, -1, -1), dtype, B)
                Fsum = torch.empty(B, dtype=dtype, device=dev)
                G = torch.empty(B, L, 3, dtype=dtype, device=dev)
                for bwd in (True, False):
                    fn = lambda: ops._su4_launch(bwd, pulses, tgt, None, None, M, 0, 1.0, (1.0, 0.05), 7, 0, None, None, None, Fsum, G if bwd else None, 0)
                    best, mean = time_kernel(fn, iters=5)
                    emit(probe="su4", dtype=dn, B=B, L=L, M=M, bwd=bwd, best_ms=best, mprops=B * M * L / best / 1e3)
    if "grid" in which:
        from universal_quantum_optimal_control_b200 import sweeps
        g = torch.Generator().manual_seed(0)
        L = 64
        pulse = torch.stack([(torch.rand(L, generator=g) * 2 - 1) * math.pi, 0.1 + 0.4 * torch.rand(L, generator=g)], -1).to(dev)
        X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
        T = torch.matrix_exp(-1j * X * (math.pi / 4)).to(dev)
        ore, ple = torch.linspace(-3, 3, 1000).to(dev), torch.linspace(-0.15, 0.15, 1000).to(dev)
        for dtype, dn in ((torch.float32, "f32"), (torch.float64, "f64")):
            fn = lambda: sweeps.fidelity_grid(pulse, T, ore, ple, dtype=dtype)
            best, mean = time_kernel(fn, iters=10)
            emit(probe="grid", dtype=dn, L=L, N=10 ** 6, best_ms=best, gprops=L * 1e6 / best / 1e6)
    if "acc" in which:
        for L, tau in ((16, (0.1, 0.5)), (64, (0.1, 0.5)), (100, (0.1, 0.5)), (256, (0.1, 0.5)), (256, (0.035, 0.07)), (400, (0.1, 0.5))):
            B, M = 8, 8192
            p64, tc64 = workload(B, L, M, torch.float64, tau)
            err = uq.philox_errors(B, M, (1.0, 0.05), 3, 0, dtype=torch.float64)
            res = {}
            for dn, dtype, fast in (("f64", torch.float64, False), ("f32", torch.float32, False), ("f32fast", torch.float32, True)):
                p = p64.to(dtype)
                tc = tc64.to(dtype)
                F = torch.empty(B * M, dtype=dtype, device=dev)
                Fsum = torch.empty(B, dtype=dtype, device=dev)
                G = torch.empty(B, L, 2, dtype=dtype, device=dev)
                ops._launch_fwdbwd(p, tc, err.to(dtype), None, M, 0, (1.0, 0.05), 0, 0, F, None, Fsum, G, uq.tuning_flags(fast_sincos=fast))
                res[dn] = (F.double(), G.double())
            for dn in ("f32", "f32fast"):
                dF = (res[dn][0] - res["f64"][0]).abs()
                dG = (res[dn][1] - res["f64"][1]).abs().max() / res["f64"][1].abs().max()
                emit(probe="acc", L=L, tau=tau, mode=dn, max_dF=dF.max().item(), rms_dF=dF.pow(2).mean().sqrt().item(),
                     rel_dG=dG.item())


if __name__ == "__main__":
    main()
