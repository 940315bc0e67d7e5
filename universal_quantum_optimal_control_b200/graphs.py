"""CUDA-graph replay of the fused step for host-driven optimisers (launch-bound configurations).

At BASELINE config 3 size (1 target x 65536 samples x 256 pulses = 1.7e7 propagations) the fused
kernels take ~80 us, less than the Python / launch overhead of issuing them.  :class:`GraphedFusedStep`
captures the whole step once -- H2D of the pulses / targets / (optional) explicit errors from pinned
staging buffers, target coefficients, fused forward+backward kernel, partials reduction, loss
finalize, D2H of the loss and the pulse gradient -- and replays it with one ``cudaGraphLaunch``.
Philox (seed, offset) live in device memory (``UQOC_FLAG_RNG_FROM_DEVICE``) so every replay draws
fresh error samples.  This is the step of ``train/GRAPE/grape_train.py`` / ``trainer.py:80-94`` when the
pulse parameters themselves are the optimisation variables (GRAPE proper, dCRAB, line searches).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from ._lib import LOSS_KINDS

FLAG_RNG_FROM_DEVICE = 8


class GraphedFusedStep:
    def __init__(self, B: int, L: int, monte_carlo: int, *, dtype=torch.float32, loss: str = "sharp", tau: float = 0.99,
                 k: float = 100, explicit_error: bool = False, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0,
                 device="cuda", flags: int = 0):
        if loss not in LOSS_KINDS:
            raise ValueError(f"unknown loss {loss!r}")
        self.B, self.L, self.M = B, L, int(monte_carlo)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("GraphedFusedStep needs a CUDA device: the uqoc ops have no CPU fallback")
        self.loss, self.tau, self.k, self.sigma, self.flags = loss, tau, k, tuple(float(s) for s in sigma), flags
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        # pinned host staging + device buffers (static addresses: graph-safe)
        self.h_pulses = torch.zeros(B, L, 2, dtype=dtype).pin_memory()
        self.h_target = torch.zeros(B, 2, 2, dtype=cdt).pin_memory()
        self.h_err = torch.zeros(2, B * self.M, dtype=dtype).pin_memory() if explicit_error else None
        self.h_rng = torch.zeros(2, dtype=torch.int64).pin_memory()
        self.h_out = torch.zeros(B * L * 2 + B + 4, dtype=dtype).pin_memory()       # [grad | Fsum | loss, Fbar, dloss, -]
        self.d_pulses = torch.zeros(B, L, 2, dtype=dtype, device=self.dev)
        self.d_target = torch.zeros(B, 2, 2, dtype=cdt, device=self.dev)
        self.d_err = torch.zeros(2, B * self.M, dtype=dtype, device=self.dev) if explicit_error else None
        self.d_rng = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self.d_out = torch.zeros(B * L * 2 + B + 4, dtype=dtype, device=self.dev)
        self.h_rng[0] = seed
        self._flags = flags | FLAG_RNG_FROM_DEVICE | ops.FLAG_RAW_TARGET
        # private workspace: the captured launch bakes its address in, so it must not be the shared grow-only one
        self.ws = ops.su2_workspace(B, L, self.M, dtype, self._flags, self.dev)
        self._step = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._stream = torch.cuda.Stream(self.dev)

    def _body(self):
        B, L, M = self.B, self.L, self.M
        self.d_pulses.copy_(self.h_pulses, non_blocking=True)
        self.d_target.copy_(self.h_target, non_blocking=True)
        if self.d_err is not None:
            self.d_err.copy_(self.h_err, non_blocking=True)
        self.d_rng.copy_(self.h_rng, non_blocking=True)
        tc = ops.raw_target(self.d_target, self.d_pulses.dtype)        # a view: the kernel forms the trace coefficients
        n_g = B * L * 2
        G, Fsum, lo = self.d_out[:n_g], self.d_out[n_g:n_g + B], self.d_out[n_g + B:n_g + B + 3]
        ops._launch_fwdbwd_loss(self.d_pulses, tc, self.d_err, M, self.sigma, self.d_rng.data_ptr(), 0, self.loss, self.tau, self.k,
                                None, None, Fsum, G, lo, self._flags, ws=self.ws)
        self.h_out.copy_(self.d_out, non_blocking=True)

    def capture(self):
        with torch.cuda.stream(self._stream):
            for _ in range(2):                       # warm-up outside capture (lazy init, smem attributes)
                self._body()
            self._stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self._stream):
                self._body()
        return self

    def __call__(self, pulses: torch.Tensor, U_target: Optional[torch.Tensor] = None, error: Optional[torch.Tensor] = None):
        """Run one step.  ``pulses`` (B, L, 2) host or device tensor; returns ``(loss, grad, mean_fid)`` as views of
        a pinned host buffer (valid until the next call)."""
        if self.graph is None:
            self.capture()
        self.h_pulses.copy_(pulses)
        if U_target is not None:
            self.h_target.copy_(U_target)
        if self.h_err is not None:
            if error is None:
                raise ValueError("this step was built with explicit_error=True: pass error (2, B*M)")
            self.h_err.copy_(error)
        self._step += 1
        self.h_rng[1] = self._step
        with torch.cuda.stream(self._stream):        # CUDAGraph.replay() launches on the CURRENT stream
            self.graph.replay()
        self._stream.synchronize()                   # the D2H copy has landed: the returned views are valid
        B, L = self.B, self.L
        n_g = B * L * 2
        return self.h_out[n_g + B], self.h_out[:n_g].view(B, L, 2), self.h_out[n_g:n_g + B] / self.M
