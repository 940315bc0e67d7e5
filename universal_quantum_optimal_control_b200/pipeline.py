"""Target-chunked, stream-pipelined fused step ("overlap the all-reduce of target-chunk n with the
kernel of chunk n+1").

The sample-sharded step ends in ONE exchange of ``[G | Fsum]`` (8.4 MB at BASELINE config 5) and, for a host-driven
optimiser, one host->device copy of the pulses and one device->host copy of the gradient.  Issued after the fused
kernel they are serial: 0.1 ms of all-reduce and ~1 ms of PCIe traffic per 8.3 ms kernel when 8 ranks copy at once
(round 1: device scaling 0.988, end-to-end 0.903).  :class:`PipelinedStep` splits the TARGETS into chunks that are
independent until the loss epilogue and runs them on two alternating streams::

    copy in :  H2D(0) H2D(1)      H2D(2)      H2D(3)
    stream A:         K(0) AR(0)              K(2) AR(2)
    stream B:                     K(1) AR(1)              K(3) AR(3)
    copy out:                 D2H(0)      D2H(1)      D2H(2)      D2H(3 + Fsum)

(AR = all-reduce of the chunk's rows of ``G``, N > 1 only; the B per-target fidelity sums of all chunks sit right behind
the last chunk's rows in the ``[G | Fsum]`` buffer and travel with them: one all-reduce and one copy for both)

so that every copy and every all-reduce except the last chunk's runs under another chunk's kernel, and consecutive
kernels overlap at their tails (blocks of chunk n+1 fill the SMs chunk n drains).  ``uqoc_su2_fwdbwd_slice`` gives chunk
rows their GLOBAL target index in the Philox counter, so the chunked step draws exactly the samples of the un-chunked
one.  The loss couples the chunks only through the scalar pooled mean fidelity: ``d loss / d pulses = scale * G`` with
``scale = loss'(Fbar) / (B M)``, so

* :meth:`run_device` (device-resident inputs) applies the epilogue once at the end (``uqoc_loss_finalize``), and
* :meth:`__call__` (pinned host buffers in and out) ships the UNSCALED ``G`` chunks as they complete and evaluates the
  scalar epilogue on the host from the B per-target sums: it returns ``(loss, G, scale, mean_fid)`` -- a host optimiser
  folds ``scale`` into its step size instead of waiting for a scaled copy.

Every rank ends with bit-identical ``G`` / ``Fsum`` (NCCL all-reduce) and therefore the same loss and scale.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import ops
from ._lib import LOSS_KINDS
from .sharding import shard_range


def host_loss(Fbar: float, loss: str, tau: float = 0.99, k: float = 100.0):
    """Loss and d loss / d Fbar of the pooled mean fidelity (SCORE.py:185-198), in double on the host."""
    if loss == "sharp":
        z = math.exp(-k * (Fbar - tau))
        lg = math.log1p(z)
        return lg * (1.0 - Fbar), -k * z / (1.0 + z) * (1.0 - Fbar) - lg
    if loss == "nll":
        return -math.log(Fbar), -1.0 / Fbar
    if loss == "infidelity":
        return 1.0 - Fbar, -1.0
    return Fbar, 1.0


class PipelinedStep:
    def __init__(self, B: int, L: int, monte_carlo: int, *, chunks="auto", dtype=torch.float32, loss: str = "sharp",
                 tau: float = 0.99, k: float = 100, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, group=None,
                 device="cuda", flags: int = 0, ramp: bool = True):
        if loss not in LOSS_KINDS:
            raise ValueError(f"unknown loss {loss!r}")
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("PipelinedStep needs a CUDA device: the uqoc ops have no CPU fallback")
        self.B, self.L, self.M_total = int(B), int(L), int(monte_carlo)
        self.loss, self.tau, self.k, self.sigma, self.seed = loss, float(tau), float(k), tuple(float(s) for s in sigma), int(seed)
        self.flags = flags | ops.FLAG_RAW_TARGET
        self.group = group
        rank, world = 0, 1
        if group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.j0, self.M = shard_range(self.M_total, rank, world)
        if self.M < 1:
            raise ValueError(f"monte_carlo = {monte_carlo} leaves rank {rank} of {world} without samples")
        self.bounds = self.chunk_bounds(self.B, chunks, ramp, torch.cuda.get_device_properties(self.dev).multi_processor_count)
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        self.dtype = dtype
        # pinned host staging (static addresses) and device buffers, both laid out [G (B, L, 2) | Fsum (B)]: a chunk is a
        # contiguous row range of each part; its gradient rows are exchanged / copied out as soon as its kernel is done, the
        # fidelity sums of all chunks together with the LAST chunk's rows (they are adjacent)
        self.n_g = B * L * 2
        self.h_pulses = torch.zeros(B, L, 2, dtype=dtype).pin_memory()
        self.h_target = torch.zeros(B, 2, 2, dtype=cdt).pin_memory()
        self.h_out = torch.zeros(self.n_g + B, dtype=dtype).pin_memory()
        self.d_pulses = torch.zeros(B, L, 2, dtype=dtype, device=self.dev)
        self.d_target = torch.zeros(B, 2, 2, dtype=cdt, device=self.dev)
        self.d_out = torch.zeros(self.n_g + B, dtype=dtype, device=self.dev)
        self._ws = [ops.su2_workspace(b1 - b0, L, self.M, dtype, self.flags, self.dev) for b0, b1 in self.bounds]
        self._streams = [torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)]
        self._copy_stream = torch.cuda.Stream(self.dev)                  # all host->device copies, issued up front
        self._out_stream = torch.cuda.Stream(self.dev)                   # all device->host copies
        self._h2d_done = [torch.cuda.Event() for _ in self.bounds]
        self._g_done = [torch.cuda.Event() for _ in self.bounds]
        self._src_pulses, self._src_target = self.h_pulses, self.h_target
        self._step = 0

    @staticmethod
    def chunk_bounds(B: int, chunks, ramp: bool = True, n_sm: int = 148):
        """Target ranges of the chunks.  The first chunk's host->device copy and the last chunk's all-reduce /
        device->host copy are the only transfers nothing can hide, so the outer chunks are small; ``chunks="auto"`` makes
        the inner ones whole WAVES of the fused kernel (5 resident 128-thread blocks per SM, one block per target at
        config-5 sample counts) so that no chunk ends in a partly filled wave: 185 + 370 + n x 740 + remainder at B200's
        148 SMs (measured 8.33 ms end to end against 8.42 for four equal chunks of the 4096-target step)."""
        B = int(B)
        if chunks == "auto":
            W = 5 * int(n_sm)
            head = [max(W // 4, 1), max(W // 2, 1)]
            n_full = (B - sum(head) - (W // 2 + W // 4)) // W
            if n_full < 1:
                chunks = 4 if B >= 64 else (2 if B >= 2 else 1)
            else:
                rem = B - sum(head) - n_full * W
                sizes = head + [W] * n_full + [rem - rem // 3, rem // 3]
                edges = [0]
                for sz in sizes:
                    if sz > 0:
                        edges.append(edges[-1] + sz)
                return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
        chunks = max(1, min(int(chunks), B))
        w = [min(2 ** i, 2 ** (chunks - 1 - i), 4) if ramp else 1 for i in range(chunks)]
        acc, edges = 0, [0]
        for wi in w:
            acc += wi
            edges.append(round(acc * B / sum(w)))
        return [(edges[i], edges[i + 1]) for i in range(chunks) if edges[i + 1] > edges[i]]

    @staticmethod
    def _host_source(t: torch.Tensor, staging: torch.Tensor) -> torch.Tensor:
        if (not t.is_cuda) and t.is_pinned() and t.dtype == staging.dtype and t.shape == staging.shape and t.is_contiguous():
            return t
        staging.copy_(t)
        return staging

    # ------------------------------------------------------------------ views
    def _views(self, buf, c):
        b0, b1 = self.bounds[c]
        return buf[b0 * self.L * 2:b1 * self.L * 2], buf[self.n_g + b0:self.n_g + b1]

    def gradient(self, buf=None) -> torch.Tensor:
        """(B, L, 2) view of the gradient part of ``buf`` (default: the pinned host buffer)."""
        buf = self.h_out if buf is None else buf
        return buf[:self.n_g].view(self.B, self.L, 2)

    def fidelity_sums(self, buf=None) -> torch.Tensor:
        buf = self.h_out if buf is None else buf
        return buf[self.n_g:]

    # ------------------------------------------------------------------ the pipeline
    def _enqueue(self, pulses_d, target_raw, offset, *, h2d: bool, d2h: bool):
        cur = torch.cuda.current_stream(self.dev)
        for st in self._streams + [self._copy_stream, self._out_stream]:
            st.wait_stream(cur)
        n = len(self.bounds)

        def copy_in(c):
            # host->device copies run on their own stream, one chunk ahead of the kernels: chunk c's kernel waits for
            # its event only, and the host issues chunk c+1's copies after chunk c's kernel launch, not before it
            b0, b1 = self.bounds[c]
            with torch.cuda.stream(self._copy_stream):
                self.d_pulses[b0:b1].copy_(self._src_pulses[b0:b1], non_blocking=True)
                self.d_target[b0:b1].copy_(self._src_target[b0:b1], non_blocking=True)
                self._h2d_done[c].record(self._copy_stream)

        if h2d:
            copy_in(0)
        for c, (b0, b1) in enumerate(self.bounds):
            st = self._streams[c % 2]
            is_last = c == n - 1
            with torch.cuda.stream(st):
                if h2d:
                    st.wait_event(self._h2d_done[c])
                G, Fsum = self._views(self.d_out, c)
                ops._launch_fwdbwd_slice(pulses_d[b0:b1], target_raw[b0:b1], None, self.M, self.j0, b0, self.sigma, self.seed, offset,
                                         Fsum, G, self.flags, ws=self._ws[c])
                if h2d and c + 1 < n:
                    copy_in(c + 1)
                if is_last:
                    # the last chunk's gradient rows end where the fidelity sums of ALL chunks begin ([G | Fsum] layout): its
                    # all-reduce and its copy-out carry both in one transfer -- nothing hides them, so one NCCL call and one
                    # copy instead of two each on the step's critical path
                    st.wait_stream(self._streams[(c + 1) % 2])      # the other stream's chunks wrote their fidelity sums
                    G = self.d_out[b0 * self.L * 2:]
                if self.group is not None:
                    import torch.distributed as dist
                    dist.all_reduce(G, group=self.group)            # one all-reduce per chunk: its gradient rows
                if d2h:
                    self._g_done[c].record(st)
            if d2h:
                # device->host copies on their own stream too: a compute stream never waits for a copy (under N ranks
                # copying at once the PCIe rate per rank halves; kernels of later chunks must not queue behind that)
                with torch.cuda.stream(self._out_stream):
                    self._out_stream.wait_event(self._g_done[c])
                    (self.h_out[b0 * self.L * 2:] if is_last else self._views(self.h_out, c)[0]).copy_(G, non_blocking=True)
        for st in self._streams + [self._copy_stream, self._out_stream]:
            cur.wait_stream(st)

    def run_device(self, pulses_d: torch.Tensor, U_target_d: torch.Tensor, offset: Optional[int] = None) -> torch.Tensor:
        """Device-resident inputs: enqueue the chunked step on the current stream's timeline and apply the loss epilogue
        once at the end.  Returns the device tensor ``{loss, Fbar, dloss/dFbar}``; the scaled gradient is
        ``self.gradient(self.d_out)``.  No host synchronisation."""
        if offset is None:
            self._step += 1
            offset = self._step
        self._enqueue(pulses_d, ops.raw_target(U_target_d, self.dtype), offset, h2d=False, d2h=False)
        return ops._finalize(self.fidelity_sums(self.d_out), self.B * self.M_total, self.loss, self.tau, self.k, self.d_out[:self.n_g])

    def __call__(self, pulses: torch.Tensor, U_target: Optional[torch.Tensor] = None, offset: Optional[int] = None):
        """One step from host buffers.  Returns ``(loss, G, scale, mean_fid)``: ``G`` (B, L, 2) =
        ``d sum_s F / d pulses`` summed over all ranks' samples (views of a pinned buffer, valid until the next call),
        ``d loss / d pulses = scale * G``, ``mean_fid`` (B,) the per-target mean fidelity."""
        if offset is None:
            self._step += 1
            offset = self._step
        # inputs that already live in pinned host memory are copied to the device straight from where they are; anything
        # else is staged through the step's own pinned buffers first (a host memcpy: 0.4 ms for 8 MB)
        self._src_pulses = self._host_source(pulses, self.h_pulses)
        if U_target is not None:
            self._src_target = self._host_source(U_target, self.h_target)
        self._enqueue(self.d_pulses, ops.raw_target(self.d_target, self.dtype), offset, h2d=True, d2h=True)
        torch.cuda.current_stream(self.dev).synchronize()
        Fs = self.fidelity_sums()
        n = self.B * self.M_total
        val, dval = host_loss(float(Fs.double().sum().item()) / n, self.loss, self.tau, self.k)
        return val, self.gradient(), dval / n, Fs / self.M_total
