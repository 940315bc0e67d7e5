#!/usr/bin/env python
"""Benchmark of the hot path: fused fwd+bwd disorder-sampled SU(2) propagation + fidelity loss.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl uqoc|reference] [--workload curriculum|grape]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): SU(2) propagations/s, fwd+bwd; 1 propagation = one (pulse, error sample)
pair, a step over B targets x M samples x L pulses is B*M*L propagations.

Workloads
  curriculum (default): BASELINE config 5 slice -- B=4096 targets x L=256 x M=4096 Philox samples per
      target PER GPU (the full config has 10^6 samples/target sharded over 8 GPUs = 131072 per GPU,
      minutes per step; this is a time-bounded slice of the same shape, weak scaling in the sample
      axis exactly like the full config).  One step = fused kernel + (N>1) one NCCL all-reduce of
      [Fsum | dSumF/dpulses] + loss finalize.
  grape: BASELINE config 3 -- B=1, L=256, M=65536 explicit eps per GPU.

One JSON line on stdout (rank 0).  `value` = device-resident whole-job throughput; `e2e` = the
same step through the public API with HOST buffers (pinned H2D of pulses/targets, D2H of loss and
gradient inside the timed region); `roofline` = the fused kernel against the FP32 FMA pipe;
`cpu_baseline` = the reference's CPU implementation (oracle/torch_port.py, same ATen op sequence)
on this host's cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PROP_FWDBWD = 116.0   # 35 forward + 81 backward, algorithmic


# ------------------------------------------------------------------------------------ workloads
def make_workload(name: str, device, seed: int = 0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if name == "curriculum":
        B, L, M = 4096, 256, 4096
        tau_lo, tau_hi = 0.1, 0.5            # train/unitary_single_qubit_gate/model_params.json:4-5
        explicit = False
    elif name == "grape":
        B, L, M = 1, 256, 65536
        tau_lo, tau_hi = 0.035, 0.07         # train/GRAPE/model_params.json:4
        explicit = True
    else:
        raise ValueError(name)
    phi = (torch.rand(B, L, generator=g) * 2 - 1) * 3.15
    tau = tau_lo + (tau_hi - tau_lo) * torch.rand(B, L, generator=g)
    pulses = torch.stack([phi, tau], -1).contiguous()
    # random SU(2) targets, as build_SU2_dataset(random=True) (SCORE.py:215-251)
    th = torch.rand(B, generator=g) * math.pi
    ph = torch.rand(B, generator=g) * 2 * math.pi
    al = torch.rand(B, generator=g) * 2 * math.pi
    n = torch.stack([th.sin() * ph.cos(), th.sin() * ph.sin(), th.cos()], 1)
    c, s = (al / 2).cos(), (al / 2).sin()
    U = torch.zeros(B, 2, 2, dtype=torch.complex64)
    U[:, 0, 0] = torch.complex(c, -s * n[:, 2])
    U[:, 0, 1] = torch.complex(-s * n[:, 1], -s * n[:, 0])
    U[:, 1, 0] = torch.complex(s * n[:, 1], -s * n[:, 0])
    U[:, 1, 1] = torch.complex(c, s * n[:, 2])
    return dict(name=name, B=B, L=L, M=M, pulses=pulses, U_target=U, explicit=explicit, sigma=(1.0, 0.05))


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        # "under load" = samples above half the max clock (idle samples before/after the region dropped)
        load = [x for x in sm if mx and x > 0.5 * mx] or sm
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(workload_name: str, L: int, budget_s: float = 12.0):
    """Reference CPU path (torch port, all host threads) on a bounded sample of the workload."""
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    B, M = 1, 4096
    phi = (torch.rand(B, L, generator=g) * 2 - 1) * 3.15
    tau = 0.1 + 0.4 * torch.rand(B, L, generator=g)
    pulses = torch.stack([phi, tau], -1)
    T = torch.eye(2, dtype=torch.complex64)[None]
    err = torch.stack([torch.randn(B * M, generator=g), 0.05 * torch.randn(B * M, generator=g)])
    tp.train_step_loss_and_grad(pulses, T, err, M)      # warm-up
    best, reps, t_all = float("inf"), 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t_all < budget_s and reps < 10):
        t0 = time.perf_counter()
        tp.train_step_loss_and_grad(pulses, T, err, M)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    props = B * M * L
    return {"value": props / best, "unit": "prop/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"B={B} target x M={M} samples x L={L} (complex64, tree product, sharp_loss, autograd backward), "
                      f"best of {reps} after warm-up; host has {os.cpu_count()} logical cores"}


def workload_config(name: str, B: int, L: int, M: int, M_total: int, world: int) -> dict:
    """The `config` object both arms print: names the workload (no model keys)."""
    return {"workload": ("BASELINE config 5 slice (curriculum, Philox eps on-chip)" if name == "curriculum"
                         else "BASELINE config 3 (GRAPE, explicit eps)"),
            "targets_B": B, "pulses_L": L, "samples_per_target_per_gpu": M, "samples_per_target_total": M_total,
            "props_per_step": float(B) * M_total * L, "loss": "sharp", "sharding": f"samples x{world}"}


def cpu_port_c(L: int):
    """Informative second CPU figure: the plain-C closed-form port (oracle/uqoc_oracle.c, OpenMP, FP64) of the same
    fwd+bwd step on all host cores.  The reference's own op sequence (cpu_baseline / --impl reference) stays THE baseline."""
    import numpy as np
    from oracle import c_oracle as co
    rng = np.random.default_rng(0)
    B, M = 1, 16384
    pulses = np.stack([rng.uniform(-3.15, 3.15, (B, L)), rng.uniform(0.1, 0.5, (B, L))], -1)
    T = np.eye(2, dtype=np.complex128)[None]
    err = np.stack([rng.normal(0, 1, B * M), rng.normal(0, 0.05, B * M)])
    co.fidelity_sum_and_grad(pulses, T, err, M)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        co.fidelity_sum_and_grad(pulses, T, err, M)
        best = min(best, time.perf_counter() - t0)
    return {"value": B * M * L / best, "unit": "prop/s", "cores": co.threads(), "kind": "port",
            "sample": f"B={B} x M={M} x L={L}, FP64 complex 2x2 closed form, OpenMP, best of 3"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = {"curriculum": (4096, 256, 4096), "grape": (1, 256, 65536)}[args.workload]
    L = wl[1]
    from oracle import torch_port as tp
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    B, M = 1, 2048        # bounded sample of the workload per step
    phi = (torch.rand(B, L, generator=g) * 2 - 1) * 3.15
    tau = 0.1 + 0.4 * torch.rand(B, L, generator=g)
    pulses = torch.stack([phi, tau], -1)
    T = torch.eye(2, dtype=torch.complex64)[None]
    err = torch.stack([torch.randn(B * M, generator=g), 0.05 * torch.randn(B * M, generator=g)])
    for _ in range(max(1, min(args.warmup, 3))):
        tp.train_step_loss_and_grad(pulses, T, err, M)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        tp.train_step_loss_and_grad(pulses, T, err, M)
    dt = (time.perf_counter() - t0) / steps
    val = B * M * L / dt
    sample = (f"per step B={B} x M={M} samples x L={L} of workload '{args.workload}' (complex64, tree product, "
              f"sharp_loss, autograd), {steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": "SU(2) propagations/s fwd+bwd", "value": val, "unit": "prop/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload description as the uqoc arm (workload_config); the bounded per-step sample is stated beside it
        "config": dict(workload_config(args.workload, wl[0], L, wl[2], wl[2] * max(1, args.gpus), max(1, args.gpus)),
                       reference_arm_sample=f"B={B} x M={M} x L={L} per step on the host CPU",
                       note="CPU reference path (torch port of the reference op sequence, oracle/torch_port.py)"),
        "cpu_baseline": {"value": val, "unit": "prop/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "prop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="uqoc", choices=["uqoc", "reference"])
    ap.add_argument("--workload", default="curriculum", choices=["curriculum", "grape"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--fast-sincos", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="tuning flags (tuning_flags())")
    ap.add_argument("--chunks", type=int, default=0,
                    help="N>1, large exchange vector: target chunks of the pipelined step (all-reduce of chunk n under kernel n+1); "
                         "0 = wave-sized chunks chosen by PipelinedStep")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="N>1: all-reduce of [Fsum|G] through NCCL, or fused into the partials reduction over NVLink peer "
                         "memory (uqoc_su2_fwdbwd_peer); auto = peer when the vector is small (<= 2^18 reals)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the uqoc ops have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    import universal_quantum_optimal_control_b200 as uq
    from universal_quantum_optimal_control_b200 import ops

    rdt = torch.float32 if args.dtype == "f32" else torch.float64
    wl = make_workload(args.workload, dev)
    B, L, M = wl["B"], wl["L"], wl["M"]
    M_total = M * world                                   # weak scaling in the sample axis
    flags = args.flags | (1 if args.fast_sincos else 0)
    pulses_h = wl["pulses"].to(rdt).pin_memory()
    target_h = wl["U_target"].pin_memory()
    pulses_d = pulses_h.to(dev)
    target_d = target_h.to(dev)
    tc = uq.target_coeffs(target_d, rdt)
    err_d = None
    if wl["explicit"]:
        # this rank's shard of the (2, B*M_total) error tensor, generated once on the device
        err_d = uq.philox_errors(B, M, wl["sigma"], seed=1234, offset=0, j0=rank * M, device=dev, dtype=rdt)
    n_g = B * L * 2
    buf = torch.empty(n_g + B, dtype=rdt, device=dev)      # [G | Fsum]: one exchange, G 16-byte aligned
    G, Fsum = buf[:n_g], buf[n_g:]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    # N>1 exchange step.  Small vectors (few targets): fused into the step over NVLink peer memory
    # (uqoc_su2_fwdbwd_peer_loss).  MB-sized vectors: NCCL all-reduce per TARGET CHUNK, overlapped with the next chunk's
    # kernel (PipelinedStep).
    px, pipe = None, None
    if group is not None and args.exchange != "nccl":
        from universal_quantum_optimal_control_b200 import peer as peer_mod
        if args.exchange == "peer" or (B + n_g) <= peer_mod.MAX_N:
            px = uq.PeerExchange(group, B, L, 2, rdt, dev)
    if px is None and err_d is None and (group is not None or args.chunks > 1):
        pipe = uq.PipelinedStep(B, L, M_total, chunks=(args.chunks if args.chunks > 0 else "auto"), dtype=rdt, sigma=wl["sigma"], seed=1234, group=group,
                                device=dev, flags=flags)
    loss_dev = torch.empty(3, dtype=rdt, device=dev)
    launches = {"n": 0}

    def step_device(i):
        """One step with device-resident inputs; counts the kernels of THIS library it launches."""
        if px is not None:         # multi-GPU, small vector: one call, exchange + loss inside it
            ops._launch_fwdbwd_peer_loss(pulses_d, tc, err_d, M, rank * M, M_total, wl["sigma"], 1234, i, "sharp", 0.99, 100, None, None,
                                         Fsum, G, loss_dev, flags, px)
            launches["n"] += 3     # fused kernel + reduction/exchange + loss epilogue (1 when the epilogue runs in-kernel)
            return loss_dev
        if pipe is not None:       # target chunks on two streams, all-reduce of chunk n under the kernel of chunk n+1
            launches["n"] += len(pipe.bounds) + 1
            return pipe.run_device(pulses_d, target_d, offset=i)
        if group is None:          # single GPU: ONE library call (fused kernel + dependent-launched epilogue)
            ops._launch_fwdbwd_loss(pulses_d, tc, err_d, M, wl["sigma"], 1234, i, "sharp", 0.99, 100, None, None, Fsum, G, loss_dev, flags)
            launches["n"] += 2
            return loss_dev
        ops._launch_fwdbwd(pulses_d, tc, err_d, None, M, rank * M, wl["sigma"], 1234, i, None, None, Fsum, G, flags)
        dist.all_reduce(buf, group=group)
        launches["n"] += 3
        return ops._finalize(Fsum, B * M_total, "sharp", 0.99, 100, G)

    def barrier():
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(args.warmup):
        step_device(i)
    barrier()

    # ---- timed region: K steps, each bracketed by events on the launching stream; L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches["n"] = 0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        loss_out = step_device(args.warmup + i)
        ev[i][1].record()
    barrier()
    n_launches = launches["n"]
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    loss_val = float(loss_out[0].item())

    # ---- the dominant kernel alone (roofline): this rank's un-chunked fused launch, no exchange, same events / flush
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        kev[i][0].record()
        ops._launch_fwdbwd(pulses_d, tc, err_d, None, M, rank * M, wl["sigma"], 1234, args.warmup + i, None, None, Fsum, G, flags)
        kev[i][1].record()
    barrier()
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # ---- end-to-end: public API, host buffers, H2D + D2H inside the timed region
    grad_h = torch.empty(B, L, 2, dtype=rdt).pin_memory()
    loss_h = torch.empty(1, dtype=rdt).pin_memory()
    err_h = err_d.cpu().pin_memory() if err_d is not None else None

    def step_e2e(i):
        p = pulses_h.to(dev, non_blocking=True).requires_grad_(True)
        T = target_h.to(dev, non_blocking=True)
        if err_h is None:
            val, _ = uq.fused_propagate_loss(p, T, monte_carlo=M_total, sigma=wl["sigma"], seed=1234, offset=i,
                                             loss="sharp", flags=flags, group=px if px is not None else group)
        elif world == 1:
            val, _ = uq.fused_propagate_loss(p, T, error=err_h.to(dev, non_blocking=True), monte_carlo=M_total,
                                             loss="sharp", flags=flags)
        else:   # explicit eps at N>1: every rank copies in only its own shard of the (2, B*M_total) tensor
            val, _ = _sharded_explicit(uq, ops, p, T, err_h.to(dev, non_blocking=True), M, M_total, rank,
                                       px if px is not None else group, flags)
        val.backward()
        grad_h.copy_(p.grad, non_blocking=True)
        loss_h.copy_(val.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def wall(fn):
        for i in range(args.warmup):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            fn(args.warmup + i)
        barrier()
        return (time.perf_counter() - t0) * 1e3 / args.steps     # host wall clock: host-side work counts end to end

    e2e = {"fused_propagate_loss + backward": wall(step_e2e)}
    # same step through the CUDA-graph API (single GPU: same H2D / D2H per step, one graph launch)
    if world == 1:
        gs = uq.GraphedFusedStep(B, L, M, dtype=rdt, loss="sharp", explicit_error=err_h is not None, sigma=wl["sigma"],
                                 seed=1234, device=dev, flags=flags)
        e2e["GraphedFusedStep"] = wall(lambda i: gs(pulses_h, target_h, err_h))
    # and through the target-chunked pipeline (copies and all-reduce of chunk n under the kernel of chunk n+1)
    if err_h is None and px is None:
        pe = pipe if pipe is not None else uq.PipelinedStep(B, L, M_total, chunks=(args.chunks if args.chunks > 0 else "auto"), dtype=rdt, sigma=wl["sigma"], seed=1234,
                                                            group=group, device=dev, flags=flags)
        e2e["PipelinedStep"] = wall(lambda i: pe(pulses_h, target_h, offset=i))
    clocks = sampler.stop() if rank == 0 else None

    # ---- the launch that was timed, against the oracle (rank 0; after the timed regions)
    parity = None
    if rank == 0 and rdt == torch.float32:
        parity = parity_check(uq, ops, wl, pulses_d, tc, err_d, M, rank * M, args.warmup, flags, dev)

    # ---- BASELINE config 3 (GRAPE, few targets) at THIS rank count: weak scaling in the sample axis, exchange named
    c3 = config3(uq, ops, dev, group, rank, world)

    # ---- max over ranks
    e2e_best = min(e2e, key=e2e.get)
    t = torch.tensor([step_ms, kern_ms] + [e2e[k] for k in sorted(e2e)], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    vals = [float(x) for x in t.tolist()]
    step_ms, kern_ms = vals[0], vals[1]
    e2e = dict(zip(sorted(e2e), vals[2:]))
    e2e_best = min(e2e, key=e2e.get)
    e2e_ms = e2e[e2e_best]

    if rank == 0:
        props_step = float(B) * M_total * L
        value = props_step / (step_ms * 1e-3)
        props_kernel = float(B) * M * L                       # one launch, one GPU
        ach_tflops = props_kernel * FLOP_PER_PROP_FWDBWD / (kern_ms * 1e-3) / 1e12
        peak_mode = 1 if rdt == torch.float64 else 0
        try:
            peak_meas, _ = uq.fp32_peak_tflops(4096, peak_mode)
            peak2, _ = uq.fp32_peak_tflops(4096, 2) if peak_mode == 0 else (None, None)
        except Exception:
            peak_meas, peak2 = None, None
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12 * (0.5 if rdt == torch.float64 else 1.0)
        peak = nominal
        h2d = pulses_h.numel() * pulses_h.element_size() + target_h.numel() * target_h.element_size()
        if err_h is not None:
            h2d += err_h.numel() * err_h.element_size()
        d2h = grad_h.numel() * grad_h.element_size() + loss_h.element_size()
        if e2e_best == "PipelinedStep":
            d2h = (n_g + B) * grad_h.element_size()            # [G | Fsum] chunks
        exchange = None
        if group is not None:
            exchange = ("nvlink peer memory, fused into the step" if px is not None else
                        (f"nccl all-reduce per target chunk ({len(pipe.bounds)} chunks), overlapped with the next chunk's kernel"
                         if pipe is not None else "nccl all-reduce"))
        line = {
            "metric": "SU(2) propagations/s fwd+bwd", "value": value, "unit": "prop/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(workload_config(args.workload, B, L, M, M_total, world),
                           sincos="mufu" if args.fast_sincos else ("poly" if (args.flags & 4) else "table"),
                           exchange=exchange, l2_flush_between_steps=True,
                           timing="CUDA events per step on the launching stream, max over ranks"),
            "e2e": {"value": props_step / (e2e_ms * 1e-3), "unit": "prop/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": e2e_best,
                    "ms_per_step_by_api": e2e,
                    "result": ("loss + G = d(sum F)/d pulses (B, L, 2) + the scalar chain factor, on the host (d loss/d pulses = scale * G)"
                               if e2e_best == "PipelinedStep" else "loss + d loss/d pulses (B, L, 2) on the host"),
                    "timing": "host wall clock around K steps, pinned host buffers in, pinned host buffers out"},
            "gpu_launches": n_launches,
            "roofline": {"bound": "fp32", "achieved": ach_tflops, "peak": peak, "unit": "TFLOP/s", "frac": ach_tflops / peak,
                         "traffic": ncu_dram_traffic(args.workload) if (rdt == torch.float32 and not args.flags and not args.fast_sincos) else None,
                         "traffic_unit": f"bytes per launch (ncu dram__bytes_read+write, profiles/{NCU_BENCH_PROFILE})",
                         "algorithmic_bytes_per_launch": float(B * L * 2 * 4 * 2 + B * 8 * 4 + B * 4),
                         "kernel": "su2_kernel_x2 (fused fwd+bwd, packed f32x2, table sin/cos)" if rdt == torch.float32 else "su2_kernel<double> (fused fwd+bwd)",
                         "kernel_ms": kern_ms, "kernel_ms_includes_exchange": False,
                         "flop_per_prop": FLOP_PER_PROP_FWDBWD,
                         "peak_source": f"nominal FP32 FMA: 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP32 figure)",
                         "measured_ffma_tflops": peak_meas, "measured_ffma2_tflops": peak2},
            "clocks": clocks, "loss": loss_val, "parity_check": parity,
        }
        line["other_configs"] = {"c3_grape_B1_L256_M65536_fwdbwd": c3}
        if world == 1:
            try:
                line["other_configs"].update(other_configs(uq, ops, dev))
                if "su4" in line["other_configs"]:                # first-class block for the two-qubit path
                    line["su4"] = line["other_configs"].pop("su4")
            except Exception as e:  # informative only
                line["other_configs"]["error"] = repr(e)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, L)
            try:
                line["cpu_port_c"] = cpu_port_c(L)
            except Exception as e:  # informative only (needs gcc or the prebuilt oracle/_build/liboracle_c.so)
                line["cpu_port_c"] = {"error": repr(e)[:200]}
        print(json.dumps(line))
    if group is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit(f"parity_check failed: {parity}")


NCU_BENCH_PROFILE = "r2_su2_fwdbwd_bench_launch_ncu.txt"


def parity_check(uq, ops, wl, pulses_d, tc, err_d, M, j0, offset, flags, dev, n_check: int = 8):
    """Re-run the exact launch the timed region ran (same shape, plan, flags, Philox key) with the per-sample fidelities
    and the errors it used written out, and compare `n_check` targets spread over the batch with the FP64 oracle
    (oracle/uqoc_oracle.py, the checker) on THOSE errors.  Bounds: BASELINE.json's 1e-5 abs on F, 1e-4 rel on the gradient."""
    import numpy as np
    from oracle import uqoc_oracle as orc
    B, L = wl["B"], wl["L"]
    n_g = B * L * 2
    buf = torch.empty(n_g + B, device=dev)
    F = torch.empty(B * M, device=dev)
    e_used = torch.empty(2, B * M, device=dev)
    ops._launch_fwdbwd(pulses_d, tc, err_d, None, M, j0, wl["sigma"], 1234, offset, F, e_used, buf[n_g:], buf[:n_g], flags)
    torch.cuda.synchronize()
    sel = sorted(set(int(round(x)) for x in np.linspace(0, B - 1, n_check)))
    U_t = wl["U_target"].numpy().astype(np.complex128)
    p64 = wl["pulses"].numpy().astype(np.float32).astype(np.float64)
    Fh = F.view(B, M)[sel].cpu().numpy()
    Gh = buf[:n_g].view(B, L, 2)[sel].cpu().numpy()
    Eh = e_used.view(2, B, M)[:, sel].cpu().numpy().astype(np.float64)
    dF = dG = 0.0
    for k, b in enumerate(sel):
        _, G_b, F_b = orc.fidelity_sum_and_grad(p64[b:b + 1], U_t[b:b + 1], Eh[:, k], M)
        dF = max(dF, float(np.abs(Fh[k] - F_b).max()))
        dG = max(dG, float(np.abs(Gh[k] - G_b[0]).max() / np.abs(G_b[0]).max()))
    return {"max_abs_dF": dF, "rel_dG": dG, "n": len(sel), "samples_per_target": M, "bounds": {"dF": 1e-5, "dG": 1e-4},
            "ok": bool(dF < 1e-5 and dG < 1e-4),
            "what": "the timed launch re-run with F_out / err_out; sampled targets vs oracle/uqoc_oracle.py (FP64) on the errors it used"}


def config3(uq, ops, dev, group, rank, world):
    """BASELINE config 3 (GRAPE: B = 1, L = 256, 65536 explicit eps PER GPU) as one library call per step, at this rank
    count; the exchange of the 513-real vector is named.  CUDA events on the launching stream, max over ranks."""
    wl = make_workload("grape", dev)
    B, L, M = wl["B"], wl["L"], wl["M"]
    p = wl["pulses"].to(dev)
    tc = uq.target_coeffs(wl["U_target"].to(dev), torch.float32)
    err = uq.philox_errors(B, M, wl["sigma"], 1, 0, j0=rank * M, device=dev)
    buf = torch.empty(B * L * 2 + B, device=dev)
    G, Fsum = buf[:B * L * 2], buf[B * L * 2:]
    lo = torch.empty(3, device=dev)
    px, exchange = None, None
    if group is not None:
        import torch.distributed as dist
        try:
            px = uq.PeerExchange(group, B, L, 2, torch.float32, dev)
            exchange = "nvlink peer memory, fused into the step (uqoc_su2_fwdbwd_peer_loss)"
        except Exception as e:  # no peer mapping on this box: still a GPU path
            exchange = f"nccl all-reduce (PeerExchange unavailable: {e!r})"[:160]

    def step():
        if group is None:
            ops._launch_fwdbwd_loss(p, tc, err, M, wl["sigma"], 1, 0, "sharp", 0.99, 100, None, None, Fsum, G, lo, 0)
        elif px is not None:
            ops._launch_fwdbwd_peer_loss(p, tc, err, M, rank * M, M * world, wl["sigma"], 1, 0, "sharp", 0.99, 100, None, None, Fsum, G,
                                         lo, 0, px)
        else:
            ops._launch_fwdbwd(p, tc, err, None, M, rank * M, wl["sigma"], 1, 0, None, None, Fsum, G, 0)
            dist.all_reduce(buf, group=group)
            ops._finalize(Fsum, B * M * world, "sharp", 0.99, 100, G)

    for _ in range(5):
        step()
    if group is not None:
        dist.barrier()
    torch.cuda.synchronize()
    iters = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t.item())
    return {"prop_per_s": float(B) * M * world * L / (ms * 1e-3), "ms": ms, "n_gpus": world, "samples_per_gpu": M,
            "exchange": exchange, "pct_fp32_peak_per_gpu": B * M * L * FLOP_PER_PROP_FWDBWD / (ms * 1e-3) / 74.45e12 * 100,
            "loss": float(lo[0].item())}


def ncu_dram_traffic(workload: str):
    """dram__bytes_read + dram__bytes_write of the fused kernel for THIS launch shape, from the committed
    `ncu --set full` capture of `tools/profile_fwdbwd.py 4096 4096 256` (profiles/r1_su2_fwdbwd_bench_launch_ncu.txt)."""
    if workload != "curriculum":
        return None
    path = os.path.join(ROOT, "profiles", NCU_BENCH_PROFILE)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_su2_fwdbwd_bench_launch_ncu.txt")
    try:
        tot, units = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * units.get(f[2], 1.0)
        return tot
    except OSError:
        return None


SU4_FLOP_PER_PROP = 1484.0      # executed FLOP per SU(4) (pulse, sample): 2 x 742 FMA-pipe lane operations (ncu)


def other_configs(uq, ops, dev):
    """Device-side rates of the other BASELINE configs (parity-test cases, not the bench line): informative."""
    from universal_quantum_optimal_control_b200 import sweeps

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    g = torch.Generator().manual_seed(0)
    L = 64
    pulse = torch.stack([(torch.rand(L, generator=g) * 2 - 1) * math.pi, 0.1 + 0.4 * torch.rand(L, generator=g)], -1).to(dev)
    X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
    T = torch.matrix_exp(-1j * X * (math.pi / 4)).to(dev)
    ore, ple = torch.linspace(-3, 3, 1000).to(dev), torch.linspace(-0.15, 0.15, 1000).to(dev)
    ms = timed(lambda: sweeps.fidelity_grid(pulse, T, ore, ple))
    out["c2_grid_L64_1e6pts_fwd"] = {"prop_per_s": L * 1e6 / (ms * 1e-3), "ms": ms}
    B, L, M = 1, 128, 32768
    p4 = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, (torch.rand(B, L, generator=g) * 2 - 1) * 3.15,
                      0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev)
    tgt = ops._su4_target(torch.diag(torch.tensor([1, 1, 1, -1], dtype=torch.complex64)).to(dev)[None], torch.float32, 1)
    buf4 = torch.empty(B + B * L * 3, device=dev)
    ms = timed(lambda: ops._su4_launch(True, p4, tgt, None, None, M, 0, 1.0, (1.0, 0.05), 7, 0, None, None, None, buf4[:B], buf4[B:], 0), 5)
    out["c4_su4_B1_L128_M32768_fwdbwd"] = {"su4_prop_per_s": B * M * L / (ms * 1e-3), "ms": ms}
    # the same launch shape with 8 targets (CZ-like target): enough blocks for every SM sub-partition
    B8 = 8
    p8 = p4.expand(B8, -1, -1).contiguous()
    tgt8 = tgt.expand(B8, -1, -1, -1).contiguous()
    buf8 = torch.empty(B8 + B8 * L * 3, device=dev)
    ms8 = timed(lambda: ops._su4_launch(True, p8, tgt8, None, None, M, 0, 1.0, (1.0, 0.05), 7, 0, None, None, None, buf8[:B8], buf8[B8:], 0), 5)
    fl = SU4_FLOP_PER_PROP
    out["su4"] = {
        "workload": "BASELINE config 4: two-qubit SU(4), L=128, M=32768 eps per target, fused fwd+bwd, eigenframe kernel, Philox eps on-chip",
        "unit": "SU(4) prop/s", "B1": {"value": B * M * L / (ms * 1e-3), "ms": ms}, "B8": {"value": B8 * M * L / (ms8 * 1e-3), "ms": ms8},
        "flop_per_prop": fl,
        "flop_count": "EXECUTED: 2 x 742 FMA-pipe lane operations per (pulse, sample) (ncu instruction counts, "
                      "profiles/r1_su4_eigenframe_fwdbwd_ncu.txt: 384 packed forward + ~358 backward / sin-cos); a dense "
                      "per-pulse scaling-and-squaring formulation (UQOC_FLAG_SU4_PADE) executes 8160",
        "roofline": {"bound": "fp32", "unit": "TFLOP/s", "peak": 74.44992,
                     "achieved_B1": B * M * L * fl / (ms * 1e-3) / 1e12, "frac_B1": B * M * L * fl / (ms * 1e-3) / 74.44992e12,
                     "achieved_B8": B8 * M * L * fl / (ms8 * 1e-3) / 1e12, "frac_B8": B8 * M * L * fl / (ms8 * 1e-3) / 74.44992e12}}
    # stock-torch-on-the-same-B200 (informative): the reference's ATen op sequence (oracle/torch_port.py) on cuda:0
    try:
        from oracle import torch_port as tp
        Bt, Lt, Mt = 1, 256, 16384
        pt = torch.stack([(torch.rand(Bt, Lt, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(Bt, Lt, generator=g)], -1).to(dev)
        Tt = torch.eye(2, dtype=torch.complex64, device=dev)[None]
        et = uq.philox_errors(Bt, Mt, (1.0, 0.05), 3, 0, device=dev)
        ms = timed(lambda: tp.train_step_loss_and_grad(pt, Tt, et, Mt), 3)
        out["reference_op_sequence_on_this_B200_B1_L256_M16384_fwdbwd"] = {"prop_per_s": Bt * Mt * Lt / (ms * 1e-3), "ms": ms,
                                                                           "note": "torch.linalg.matrix_exp + bmm tree + autograd, complex64, cuda"}
    except Exception as e:  # informative only
        out["reference_op_sequence_on_this_B200_B1_L256_M16384_fwdbwd"] = {"error": repr(e)[:200]}
    # the reference scripts' own step shapes (monte_carlo = 1000, trainer.py:34): shipped single-qubit config
    # (SCORE.py:316-328 batch 200, L = 100) and the GRAPE script (grape_train.py:306 batch 100, L = 400)
    for name, (Bs, Ls, Ms, tlo, thi) in (("shipped_score_B200_L100_M1000_fwdbwd", (200, 100, 1000, 0.1, 0.5)),
                                         ("shipped_grape_B100_L400_M1000_fwdbwd", (100, 400, 1000, 0.035, 0.07))):
        ps = torch.stack([(torch.rand(Bs, Ls, generator=g) * 2 - 1) * 3.15, tlo + (thi - tlo) * torch.rand(Bs, Ls, generator=g)], -1).to(dev)
        tcs = uq.target_coeffs(torch.eye(2, dtype=torch.complex64, device=dev)[None].expand(Bs, -1, -1).contiguous(), torch.float32)
        bufs = torch.empty(Bs * Ls * 2 + Bs, device=dev)
        ng = Bs * Ls * 2
        ms = timed(lambda: (ops._launch_fwdbwd(ps, tcs, None, None, Ms, 0, (1.0, 0.05), 1, 0, None, None, bufs[ng:], bufs[:ng], 0),
                            ops._finalize(bufs[ng:], Bs * Ms, "sharp", 0.99, 100, bufs[:ng])))
        out[name] = {"prop_per_s": Bs * Ms * Ls / (ms * 1e-3), "ms": ms}
        if name.startswith("shipped_score"):
            # HOST time per step at the reference's shipped step size: the autograd path (fused_propagate_loss + backward(),
            # torch's engine included) and the one-C-call path with pre-sized buffers (FusedStep)
            import time as _time
            Ts = torch.eye(2, dtype=torch.complex64, device=dev)[None].expand(Bs, -1, -1).contiguous()
            pg = ps.clone().requires_grad_(True)

            def host_us(fn, n=200):
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                t0 = _time.perf_counter()
                for _ in range(n):
                    fn()
                dt = _time.perf_counter() - t0
                torch.cuda.synchronize()
                return dt / n * 1e6

            def autograd_step():
                pg.grad = None
                l_, _ = uq.fused_propagate_loss(pg, Ts, monte_carlo=Ms, sigma=(1.0, 0.05), seed=1, offset=0)
                l_.backward()

            class _Ident(torch.autograd.Function):          # floor: what torch's autograd costs around ANY custom CUDA op
                @staticmethod
                def forward(ctx, x):
                    return x.sum()

                @staticmethod
                def backward(ctx, g):
                    return g.expand(Bs, Ls, 2)

            def autograd_floor():
                pg.grad = None
                _Ident.apply(pg).backward()

            fs = uq.FusedStep(Bs, Ls, Ms, sigma=(1.0, 0.05), seed=1, device=dev)
            out[name]["host_us_per_step"] = {"fused_propagate_loss + backward (autograd)": host_us(autograd_step),
                                             "torch autograd floor (trivial custom Function + backward)": host_us(autograd_floor),
                                             "FusedStep (one C call, pre-sized buffers)": host_us(lambda: fs(ps, Ts, offset=0))}
    wl = make_workload("curriculum", dev)
    B, L, M = 512, wl["L"], wl["M"]
    p = wl["pulses"][:B].double().to(dev)
    tc = uq.target_coeffs(wl["U_target"][:B].to(dev), torch.float64)
    buf = torch.empty(B * L * 2 + B, dtype=torch.float64, device=dev)
    ms = timed(lambda: ops._launch_fwdbwd(p, tc, None, None, M, 0, wl["sigma"], 1, 0, None, None, buf[B * L * 2:], buf[:B * L * 2], 0), 3)
    out["c5_slice_fp64_B512_fwdbwd"] = {"prop_per_s": B * M * L / (ms * 1e-3), "ms": ms}
    return out


def _sharded_explicit(uq, ops, p, T, e_local, M, M_total, rank, group, flags):
    """Explicit-eps workload at N>1: each rank already holds its own shard of the error tensor."""
    import torch.distributed as dist
    tc = uq.target_coeffs(T, p.dtype)
    return ops._FusedPropagateLoss.apply(p, tc, e_local, M, rank * M, M_total, (1.0, 0.05), 0, 0, "sharp", 0.99, 100, flags,
                                         group, None, None, None)


if __name__ == "__main__":
    main()
