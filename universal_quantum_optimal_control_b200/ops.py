"""Host-side mirror of the reference's hot-path interface over libuqoc.so.

Two layers:

* ``fused_propagate_loss`` / ``propagate_fidelity`` -- the fused ops that carry the
  performance (pulses are NOT repeated per Monte-Carlo sample; errors optionally generated
  on-chip by Philox); they replace ``trainer.py:80-90`` in one call.
* reference-signature adapters -- ``batched_unitary_generator``, ``fidelity``,
  ``sharp_loss``, ``negative_log_loss``, ``infidelity_loss``, ``custom_loss``,
  ``get_ore_ple_error_distribution``, ``get_ore_error_distribution`` -- same names, argument
  meaning and error behaviour as
  ``train/unitary_single_qubit_gate/universal_single_qubit_SCORE.py:77-198`` so they can be
  passed unchanged to ``UniversalModelTrainer`` (``model/universal_model_trainer.py:27-33``).

Everything runs through the C ABI on CUDA tensors; CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import F32, F64, FLAG_FAST_SINCOS, LOSS_KINDS, check
from .peer import PeerExchange
from .sharding import shard_errors, shard_range

__all__ = [
    "fused_propagate_loss", "fused_head_propagate_loss", "FusedStep", "propagate_fidelity", "batched_unitary_generator", "fidelity",
    "sharp_loss", "negative_log_loss", "infidelity_loss", "custom_loss",
    "get_ore_ple_error_distribution", "get_ore_error_distribution", "philox_errors",
    "target_coeffs", "tuning_flags", "fp32_peak_tflops",
    "fused_propagate_loss_su4", "su4_unitary_generator", "philox_errors_su4", "autotune_flags",
]


# ----------------------------------------------------------------------------- helpers
def _dt(t: torch.Tensor) -> int:
    return F64 if t.dtype == torch.float64 else F32


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the uqoc ops have no CPU fallback")


def _real_dtype(t: torch.Tensor) -> torch.dtype:
    if t.dtype in (torch.float64, torch.complex128):
        return torch.float64
    return torch.float32


_WORKSPACES: dict = {}


def alloc_workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """A workspace buffer for the library: ZERO-filled (its first 256 bytes are the ticket counter of the in-kernel
    epilogue, include/uqoc.h) and at least 1 MiB.  Owners of CUDA graphs allocate their own so that the address a
    captured launch baked in stays alive and private."""
    return torch.zeros(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)


def _workspace(nbytes: int, device: torch.device) -> Optional[torch.Tensor]:
    """Grow-only scratch buffer per (device, stream) -- the library never allocates.  Not for captured launches
    (a later, larger request replaces the buffer): graph owners pass their own ``ws``."""
    if nbytes <= 0:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = alloc_workspace(nbytes, device)
        _WORKSPACES[key] = buf
    return buf


class _on_device:
    """``cudaSetDevice`` is the caller's job (include/uqoc.h): make the tensors' device current around a library
    call when it is not already (the library takes the SM count and launches on the CURRENT device)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device: torch.device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _call(name: str, device: torch.device, *args) -> None:
    with _on_device(device):
        check(getattr(_lib.lib(), name)(*args), name)


def tuning_flags(st: int = 0, lps: int = 0, splits: int = 0, fast_sincos: bool = False, no_packed: bool = False,
                 no_table: bool = False, wps: int = 0, su4_pade: bool = False) -> int:
    """Pack the launch-shape overrides of include/uqoc.h (0 = library heuristic)."""
    return ((FLAG_FAST_SINCOS if fast_sincos else 0) | (2 if no_packed else 0) | (4 if no_table else 0)
            | (16 if wps == 4 else 0) | (32 if wps == 1 else 0) | (64 if su4_pade else 0)
            | ((st & 0xF) << 8) | ((lps & 0x3F) << 12)
            | ((splits & 0xFFF) << 18))


_AUTOTUNE_CACHE: dict = {}


def autotune_flags(B: int, L: int, M: int, dtype: torch.dtype = torch.float32, device="cuda", iters: int = 5,
                   candidates: Optional[Sequence[int]] = None) -> int:
    """Measure the launch shapes of the fused SU(2) kernel on THIS device for one problem size and return the
    fastest ``flags`` word (cached per (B, L, M, dtype, device)).  A training loop calls the same shape every
    step (``trainer.py:80-90``), so one measurement at start-up replaces the library heuristic
    (``make_plan`` in csrc/uqoc_api.cu), which is tuned on a handful of shapes only.  Synchronises."""
    dev = torch.device(device)
    key = (int(B), int(L), int(M), dtype, dev.index if dev.index is not None else torch.cuda.current_device())
    if key in _AUTOTUNE_CACHE:
        return _AUTOTUNE_CACHE[key]
    if candidates is None:
        candidates = [0] if dtype == torch.float64 else [
            0, tuning_flags(st=4, wps=1), tuning_flags(st=2, wps=1), tuning_flags(st=4, wps=4), tuning_flags(st=2, wps=4)]
    g = torch.Generator().manual_seed(0)
    pulses = torch.stack([(torch.rand(B, L, generator=g) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L, generator=g)], -1).to(dev, dtype)
    tc = torch.zeros(B, 8, dtype=dtype, device=dev)
    tc[:, 0] = 2.0                                           # identity target
    Fsum = torch.empty(B, dtype=dtype, device=dev)
    G = torch.empty(B, L, 2, dtype=dtype, device=dev)
    best, best_ms = 0, float("inf")
    for fl in candidates:
        try:
            for i in range(2):
                _launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 1, i, None, None, Fsum, G, fl)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                _launch_fwdbwd(pulses, tc, None, None, M, 0, (1.0, 0.05), 1, i, None, None, Fsum, G, fl)
            e1.record()
            e1.synchronize()
        except Exception:                                    # a shape this size does not support (shared memory)
            continue
        ms = e0.elapsed_time(e1) / iters
        if ms < best_ms * 0.98:                              # prefer the earlier (default-first) candidate on ties
            best, best_ms = fl, ms
    _AUTOTUNE_CACHE[key] = best
    return best


FLAG_RAW_TARGET = 128


def raw_target(U_target: torch.Tensor, real_dtype: torch.dtype) -> torch.Tensor:
    """(B,2,2) complex targets as the interleaved real array the kernels read under ``UQOC_FLAG_RAW_TARGET``
    (no launch when the tensor already has the matching complex dtype and is contiguous)."""
    _require_cuda(U_target, "U_target")
    if U_target.ndim != 3 or U_target.shape[-2:] != (2, 2):
        raise ValueError("'U_target' must have shape (B, 2, 2)")
    cdt = torch.complex64 if real_dtype == torch.float32 else torch.complex128
    t = U_target if U_target.dtype == cdt else U_target.to(cdt)
    return torch.view_as_real(t.resolve_conj().contiguous())


def target_coeffs(U_target: torch.Tensor, real_dtype: torch.dtype) -> torch.Tensor:
    """(B,2,2) complex targets -> (B,8) trace coefficients (``uqoc_su2_target_coeffs``)."""
    _require_cuda(U_target, "U_target")
    if U_target.ndim != 3 or U_target.shape[-2:] != (2, 2):
        raise ValueError("'U_target' must have shape (B, 2, 2)")
    if not U_target.is_complex():
        U_target = U_target.to(torch.complex64 if real_dtype == torch.float32 else torch.complex128)
    ur = torch.view_as_real(U_target.resolve_conj()).to(real_dtype).contiguous()
    out = torch.empty(U_target.shape[0], 8, dtype=real_dtype, device=U_target.device)
    _call("uqoc_su2_target_coeffs", out.device, _ptr(ur), U_target.shape[0], _ptr(out), _dt(out), _stream(out.device))
    return out


def philox_errors(B: int, M: int, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, offset: int = 0, j0: int = 0,
                  device="cuda", dtype=torch.float32) -> torch.Tensor:
    """(2, B*M) errors from the library's counter-based stream (``uqoc_philox_errors``)."""
    out = torch.empty(2, B * M, dtype=dtype, device=device)
    _call("uqoc_philox_errors", out.device, B, M, j0, float(sigma[0]), float(sigma[1]), seed, offset, _ptr(out), _dt(out),
                                        _stream(out.device))
    return out


def fp32_peak_tflops(iters: int = 4096, mode: int = F32) -> Tuple[float, float]:
    """Measured dependent-free FMA throughput (TFLOP/s, ms): 0 = FFMA, 1 = DFMA, 2 = FFMA2."""
    tf, ms = C.c_double(0), C.c_double(0)
    check(_lib.lib().uqoc_fp32_peak_probe(iters, mode, C.byref(tf), C.byref(ms)), "uqoc_fp32_peak_probe")
    return tf.value, ms.value


# ----------------------------------------------------------------------------- fused op
_WS_BYTES: dict = {}


def _su2_ws_bytes(B, L, M, dt, flags, device) -> int:
    """uqoc_su2_workspace_bytes, memoised per device (the plan depends on the SM count only)."""
    key = (B, L, M, dt, flags, device.index)
    n = _WS_BYTES.get(key)
    if n is None:
        with _on_device(device):
            n = _WS_BYTES[key] = int(_lib.lib().uqoc_su2_workspace_bytes(B, L, M, dt, flags))
    return n


def su2_workspace(B: int, L: int, M: int, dtype: torch.dtype, flags: int, device) -> torch.Tensor:
    """A private, zero-initialised workspace for one (B, L, M) launch shape (CUDA-graph owners)."""
    dev = torch.device(device)
    return alloc_workspace(_su2_ws_bytes(B, L, M, F64 if dtype == torch.float64 else F32, flags, dev), dev)


def _ws_for(pulses, M, flags, ws):
    B, L, _ = pulses.shape
    n = _su2_ws_bytes(B, L, M, _dt(pulses), flags, pulses.device)
    if ws is None:
        ws = _workspace(n, pulses.device)
    elif ws.numel() < n:
        raise ValueError(f"workspace of {ws.numel()} bytes is too small for this launch shape ({n} bytes)")
    return ws, n


def _launch_fwdbwd(pulses, tc, error, weight, M, j0, sigma, seed, offset, F_out, err_out, Fsum, G, flags, ws=None):
    B, L, _ = pulses.shape
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_fwdbwd", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), _ptr(weight), B, L, M, j0, sigma[0], sigma[1],
          seed, offset, _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(ws), ws_bytes, _dt(pulses), flags,
          _stream(pulses.device))


def _launch_fwdbwd_slice(pulses, tc, error, M, j0, b0, sigma, seed, offset, Fsum, G, flags, ws=None):
    """Rows [b0, b0 + B) of a larger batch: b0 only enters the Philox counter (target-chunked pipelines)."""
    B, L, _ = pulses.shape
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_fwdbwd_slice", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), None, B, L, M, j0, b0, sigma[0], sigma[1],
          seed, offset, None, None, _ptr(Fsum), _ptr(G), _ptr(ws), ws_bytes, _dt(pulses), flags, _stream(pulses.device))


def _check_px(px, pulses):
    B, L, P = pulses.shape
    if not px.matches(B, L, P, pulses.dtype):
        raise ValueError(f"PeerExchange was built for (B, L, P, dtype) = {(px.B, px.L, px.P, px.dtype)}, "
                         f"got {(B, L, P, pulses.dtype)}")


def _launch_fwdbwd_peer(pulses, tc, error, M, j0, sigma, seed, offset, F_out, err_out, Fsum, G, flags, px, ws=None):
    """This rank's shard + fused [partials reduction | NVLink peer exchange]: Fsum / G come back summed over ranks."""
    B, L, _ = pulses.shape
    _check_px(px, pulses)
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_fwdbwd_peer", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), None, B, L, M, j0, sigma[0], sigma[1],
          seed, offset, _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(ws), ws_bytes, px.rank, px.world,
          px.data_ptrs, px.flag_ptrs, px.next_epoch(), _dt(pulses), flags, _stream(pulses.device))


def _launch_fwdbwd_peer_loss(pulses, tc, error, M, j0, M_total, sigma, seed, offset, loss, tau, k, F_out, err_out, Fsum, G,
                             loss_out, flags, px, ws=None):
    """Multi-GPU step in one call: shard + partials reduction + NVLink peer exchange + loss epilogue (one launch for
    small exchange vectors)."""
    B, L, _ = pulses.shape
    _check_px(px, pulses)
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_fwdbwd_peer_loss", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), B, L, M, j0, M_total, sigma[0],
          sigma[1], seed, offset, LOSS_KINDS[loss], tau, k, _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(loss_out),
          _ptr(ws), ws_bytes, px.rank, px.world, px.data_ptrs, px.flag_ptrs, px.next_epoch(), _dt(pulses), flags,
          _stream(pulses.device))


def _launch_fwdbwd_loss(pulses, tc, error, M, sigma, seed, offset, loss, tau, k, F_out, err_out, Fsum, G, loss_out, flags,
                        ws=None):
    """Single-GPU step: fused kernel with the partials reduction + loss epilogue in its last block (one launch) or, for
    large partial volumes, one or two follow-up launches."""
    B, L, _ = pulses.shape
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_fwdbwd_loss", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), B, L, M, sigma[0], sigma[1], seed, offset,
          LOSS_KINDS[loss], tau, k, _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(loss_out), _ptr(ws), ws_bytes,
          _dt(pulses), flags, _stream(pulses.device))


def _launch_forward(pulses, tc, error, M, j0, sigma, seed, offset, U_out, F_out, err_out, Fsum, flags, ws=None):
    B, L, _ = pulses.shape
    ws, ws_bytes = _ws_for(pulses, M, flags, ws)
    _call("uqoc_su2_forward", pulses.device, _ptr(pulses), _ptr(tc), _ptr(error), B, L, M, j0, sigma[0], sigma[1], seed,
          offset, _ptr(U_out), _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(ws), ws_bytes, _dt(pulses), flags,
          _stream(pulses.device))


def _finalize(Fsum, n_total, loss, tau, k, G):
    loss_out = torch.empty(3, dtype=Fsum.dtype, device=Fsum.device)
    _call("uqoc_loss_finalize", Fsum.device, _ptr(Fsum), Fsum.numel(), float(n_total), LOSS_KINDS[loss], float(tau), float(k),
          _ptr(G), 0 if G is None else G.numel(), _ptr(loss_out), _dt(Fsum), _stream(Fsum.device))
    return loss_out


class _FusedPropagateLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pulses, tc, error, M, j0, M_total, sigma, seed, offset, loss, tau, k, flags, group, F_out, err_out, ws=None):
        B, L, _ = pulses.shape
        need_grad = ctx.needs_input_grad[0]
        n_g = B * L * 2 if need_grad else 0          # one buffer [G | Fsum]: ONE exchange, G 16-byte aligned for any B
        buf = torch.empty(n_g + B, dtype=pulses.dtype, device=pulses.device)
        G, Fsum = (buf[:n_g] if need_grad else None), buf[n_g:]
        if need_grad and group is None:
            loss_out = torch.empty(3, dtype=pulses.dtype, device=pulses.device)
            _launch_fwdbwd_loss(pulses, tc, error, M, sigma, seed, offset, loss, tau, k, F_out, err_out, Fsum, G, loss_out, flags, ws)
        elif need_grad and isinstance(group, PeerExchange):
            # exchange (NVLink peer memory, no NCCL call) and loss epilogue inside the fused kernel's last block: peer.py
            loss_out = torch.empty(3, dtype=pulses.dtype, device=pulses.device)
            _launch_fwdbwd_peer_loss(pulses, tc, error, M, j0, M_total, sigma, seed, offset, loss, tau, k, F_out, err_out, Fsum,
                                     G, loss_out, flags, group, ws)
        else:
            if isinstance(group, PeerExchange):
                group = group.group
            if need_grad:
                _launch_fwdbwd(pulses, tc, error, None, M, j0, sigma, seed, offset, F_out, err_out, Fsum, G, flags, ws)
            else:
                _launch_forward(pulses, tc, error, M, j0, sigma, seed, offset, None, F_out, err_out, Fsum, flags, ws)
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)   # [G | Fsum]: the one exchange step
            loss_out = _finalize(Fsum, B * M_total, loss, tau, k, G)
        mean_fid = Fsum / M_total
        if need_grad:
            ctx.save_for_backward(G.view(B, L, 2))
        ctx.mark_non_differentiable(mean_fid)
        return loss_out[0], mean_fid

    @staticmethod
    def backward(ctx, g_loss, _g_mean):
        (G,) = ctx.saved_tensors
        return (g_loss * G,) + (None,) * 16


def fused_propagate_loss(pulses: torch.Tensor, U_target: torch.Tensor, *, error: Optional[torch.Tensor] = None,
                         monte_carlo: int, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, offset: int = 0,
                         loss: str = "sharp", tau: float = 0.99, k: float = 100, dtype: Optional[torch.dtype] = None,
                         fast_sincos: bool = False, flags: int = 0, group=None,
                         F_out: Optional[torch.Tensor] = None, err_out: Optional[torch.Tensor] = None,
                         workspace: Optional[torch.Tensor] = None):
    """Disorder-averaged loss of ``trainer.py:80-88`` and (through autograd) its pulse gradient.

    pulses (B, L, 2) real [phi, tau] (un-repeated); U_target (B, 2, 2) complex;
    error (2, B*monte_carlo) explicit [delta; eps] with sample s = b*M + j, or None for
    on-chip Philox N(0, sigma) samples keyed by (seed, offset).  With ``group`` (a
    torch.distributed process group) every rank handles samples
    j in [rank*M/R, (rank+1)*M/R) of every target and one all-reduce combines [Fsum | G]; with
    ``group`` a :class:`PeerExchange` that exchange is fused into the partials reduction over NVLink
    peer memory (small exchange vectors: few targets).  ``workspace``: a private buffer from :func:`su2_workspace`
    for callers that capture the launch in a CUDA graph (default: a shared grow-only buffer per stream).
    Returns ``(loss scalar, mean fidelity per target (B,))``.
    """
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")
    _require_cuda(pulses, "pulses")
    if loss not in LOSS_KINDS:
        raise ValueError(f"unknown loss {loss!r}; expected one of {sorted(LOSS_KINDS)}")
    rdt = dtype or _real_dtype(pulses)
    B, L, _ = pulses.shape
    M_total = int(monte_carlo)
    if U_target.shape[0] != B:
        raise ValueError(f"U_target batch {U_target.shape[0]} != pulses batch {B}")
    p = pulses.to(rdt).contiguous()
    tc = raw_target(U_target, rdt)                        # (B, 2, 2, 2): the kernels form the trace coefficients themselves
    flags = flags | FLAG_RAW_TARGET
    rank, world = 0, 1
    if isinstance(group, PeerExchange):                   # NVLink peer-memory exchange instead of NCCL (peer.py)
        rank, world = group.rank, group.world
    elif group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    j0, M = shard_range(M_total, rank, world)
    if error is not None:
        _require_cuda(error, "error")
        if error.shape != (2, B * M_total):
            raise ValueError(f"'error' must have shape (2, {B * M_total}), got {tuple(error.shape)}")
        error = shard_errors(error.to(rdt), B, M_total, j0, M).contiguous()
    fl = flags | (FLAG_FAST_SINCOS if fast_sincos else 0)
    return _FusedPropagateLoss.apply(p, tc, error, M, j0, M_total, tuple(float(s) for s in sigma), int(seed), int(offset),
                                     loss, tau, k, fl, group, F_out, err_out, workspace)


class FusedStep:
    """The fused step as ONE C call with pre-sized outputs, no autograd: for callers that own the pulse parameters
    (GRAPE proper ``train/GRAPE/grape_train.py``, dCRAB, line searches) or hand the gradient to ``pulses.backward(grad)``
    themselves.  Everything that does not change from step to step -- output buffers, the private workspace, the launch
    plan's workspace size, the ctypes argument list -- is built once; a call converts five arguments and enters
    ``uqoc_su2_fwdbwd_loss`` (host time ~25 us against ~150 us for ``fused_propagate_loss(...)`` + ``backward()``, whose
    cost is torch's autograd engine, tools/hostprof2.py).

        step = FusedStep(B, L, monte_carlo=1000, sigma=(0.4, 0.05), seed=0)
        loss, grad, fsum = step(pulses, U_target, offset=it)      # device views, OVERWRITTEN by the next call

    ``loss`` is a 0-dim view, ``grad`` = d loss / d pulses (B, L, 2), ``fsum`` = per-target fidelity sums (mean = fsum / M).
    """

    def __init__(self, B: int, L: int, monte_carlo: int, *, dtype: torch.dtype = torch.float32, loss: str = "sharp",
                 tau: float = 0.99, k: float = 100, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, device="cuda",
                 flags: int = 0):
        if loss not in LOSS_KINDS:
            raise ValueError(f"unknown loss {loss!r}; expected one of {sorted(LOSS_KINDS)}")
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("FusedStep needs a CUDA device: the uqoc ops have no CPU fallback")
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.B, self.L, self.M, self.dtype = int(B), int(L), int(monte_carlo), dtype
        self.cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        self.flags = int(flags) | FLAG_RAW_TARGET
        n_g = self.B * self.L * 2
        self.buf = torch.zeros(n_g + self.B + 3, dtype=dtype, device=self.dev)         # [G | Fsum | loss, Fbar, dloss/dFbar]
        self.grad, self.fsum = self.buf[:n_g].view(self.B, self.L, 2), self.buf[n_g:n_g + self.B]
        self.loss_out = self.buf[n_g + self.B:]
        self.loss = self.loss_out[0]
        self.ws = su2_workspace(self.B, self.L, self.M, dtype, self.flags, self.dev)
        self._fn = _lib.lib().uqoc_su2_fwdbwd_loss
        dt = F64 if dtype == torch.float64 else F32
        # (pulses, target, err) and (offset) and (stream) are filled per call; the rest is fixed
        self._tail = (C.c_void_p(self.fsum.data_ptr()), C.c_void_p(self.grad.data_ptr()), C.c_void_p(self.loss_out.data_ptr()),
                      C.c_void_p(self.ws.data_ptr()), self.ws.numel(), dt, self.flags)
        self._mid = (self.B, self.L, self.M, float(sigma[0]), float(sigma[1]), int(seed))
        self._loss = (LOSS_KINDS[loss], float(tau), float(k), None, None)              # F_out, err_out unused

    def __call__(self, pulses: torch.Tensor, U_target: torch.Tensor, *, error: Optional[torch.Tensor] = None, offset: int = 0):
        if pulses.shape != (self.B, self.L, 2) or pulses.dtype != self.dtype or not pulses.is_contiguous() or not pulses.is_cuda:
            raise ValueError(f"'pulses' must be a contiguous CUDA ({self.B}, {self.L}, 2) tensor of dtype {self.dtype}")
        if U_target.shape != (self.B, 2, 2) or U_target.dtype != self.cdtype or not U_target.is_contiguous():
            raise ValueError(f"'U_target' must be a contiguous ({self.B}, 2, 2) tensor of dtype {self.cdtype}")
        e_ptr = None
        if error is not None:
            if error.shape != (2, self.B * self.M) or error.dtype != self.dtype or not error.is_contiguous():
                raise ValueError(f"'error' must be a contiguous (2, {self.B * self.M}) tensor of dtype {self.dtype}")
            e_ptr = error.data_ptr()
        prev = torch.cuda.current_device()
        if prev != self.dev.index:
            torch.cuda.set_device(self.dev.index)
        try:
            rc = self._fn(pulses.data_ptr(), U_target.data_ptr(), e_ptr, *self._mid, int(offset), *self._loss, *self._tail,
                          torch.cuda.current_stream(self.dev).cuda_stream)
        finally:
            if prev != self.dev.index:
                torch.cuda.set_device(prev)
        if rc != 0:
            check(rc, "uqoc_su2_fwdbwd_loss")
        return self.loss, self.grad, self.fsum


class _FusedHeadPropagateLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, tc, error, M, sigma, seed, offset, loss, tau, k, flags, mode, ranges, scale, phi_offset, base,
                pulses_out, F_out, err_out, ws):
        B, L, P = logits.shape
        need_grad = ctx.needs_input_grad[0]
        n_g = B * L * P if need_grad else 0
        buf = torch.empty(n_g + B, dtype=logits.dtype, device=logits.device)
        G, Fsum = (buf[:n_g] if need_grad else None), buf[n_g:]
        loss_out = torch.empty(3, dtype=logits.dtype, device=logits.device)
        # the workspace rows are sized for the widest gradient row (uqoc_su2_workspace_bytes)
        n = _su2_ws_bytes(B, L, M, _dt(logits), flags, logits.device)
        if ws is None:
            ws = _workspace(n, logits.device)
        elif ws.numel() < n:
            raise ValueError(f"workspace of {ws.numel()} bytes is too small for this launch shape ({n} bytes)")
        rg = (C.c_double * 4)(*ranges)
        _call("uqoc_su2_head_step", logits.device, _ptr(logits), mode, rg, float(scale), _ptr(phi_offset), _ptr(base), _ptr(tc),
              _ptr(error), B, L, M, sigma[0], sigma[1], seed, offset, LOSS_KINDS[loss], float(tau), float(k), _ptr(pulses_out),
              _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(loss_out), _ptr(ws), n, _dt(logits), flags,
              _stream(logits.device))
        mean_fid = Fsum / M
        if need_grad:
            ctx.save_for_backward(G.view(B, L, P))
        ctx.mark_non_differentiable(mean_fid)
        return loss_out[0], mean_fid

    @staticmethod
    def backward(ctx, g_loss, _g_mean):
        (G,) = ctx.saved_tensors
        return (g_loss * G,) + (None,) * 19


def fused_head_propagate_loss(logits: torch.Tensor, U_target: torch.Tensor, *, head: str,
                              pulse_ranges: Sequence[Sequence[float]], phi_offset: Optional[torch.Tensor] = None,
                              base_pulse: Optional[torch.Tensor] = None, scale: float = 0.2,
                              error: Optional[torch.Tensor] = None, monte_carlo: int, sigma: Sequence[float] = (1.0, 0.05),
                              seed: int = 0, offset: int = 0, loss: str = "sharp", tau: float = 0.99, k: float = 100,
                              dtype: Optional[torch.dtype] = None, flags: int = 0, pulses_out: Optional[torch.Tensor] = None,
                              F_out: Optional[torch.Tensor] = None, err_out: Optional[torch.Tensor] = None,
                              workspace: Optional[torch.Tensor] = None):
    """:func:`fused_propagate_loss` with the pulse generator's element-wise head folded in:
    ``logits`` are the raw outputs of the model's last linear layer -- (B, L, 2) for ``head="transformer"``
    (``model/universal_model.py:126-143``: sigmoid, range map, optional finetune ``base_pulse`` with ``scale``, relu on
    tau, ``phi_offset`` = target azimuth, wrap) or (B, L, 3) for ``head="grape"`` (``model/GRAPE_model.py:76-89``).  The
    fused kernel applies the head while it stages the pulse train and the head's backward while it writes the gradient:
    ``loss.backward()`` delivers ``d loss / d logits`` and no pulses tensor exists in between (``pulses_out`` (B, L, 2)
    optionally receives the pulses for logging).  Returns ``(loss, mean fidelity per target)``."""
    if head not in ("transformer", "grape"):
        raise ValueError(f"unknown head {head!r}; expected 'transformer' or 'grape'")
    P = 2 if head == "transformer" else 3
    if logits.ndim != 3 or logits.shape[-1] != P:
        raise ValueError(f"'logits' must have shape (B, L, {P})")
    _require_cuda(logits, "logits")
    if loss not in LOSS_KINDS:
        raise ValueError(f"unknown loss {loss!r}; expected one of {sorted(LOSS_KINDS)}")
    rdt = dtype or _real_dtype(logits)
    B, L, _ = logits.shape
    if U_target.shape[0] != B:
        raise ValueError(f"U_target batch {U_target.shape[0]} != logits batch {B}")
    (lo0, hi0), (lo1, hi1) = pulse_ranges
    M = int(monte_carlo)
    if error is not None:
        _require_cuda(error, "error")
        if error.shape != (2, B * M):
            raise ValueError(f"'error' must have shape (2, {B * M}), got {tuple(error.shape)}")
        error = error.to(rdt).contiguous()
    off = base = None
    if head == "transformer":
        off = None if phi_offset is None else phi_offset.to(logits.device, rdt).contiguous()
        base = None if base_pulse is None else base_pulse.to(logits.device, rdt).contiguous()
        if off is not None and off.shape != (B,):
            raise ValueError(f"'phi_offset' must have shape ({B},)")
        if base is not None and base.shape != (L, 2):
            raise ValueError(f"'base_pulse' must have shape ({L}, 2)")
    elif phi_offset is not None or base_pulse is not None:
        raise ValueError("the GRAPE head takes no phi_offset / base_pulse")
    if pulses_out is not None and (pulses_out.shape != (B, L, 2) or pulses_out.dtype != rdt or not pulses_out.is_contiguous()):
        raise ValueError(f"'pulses_out' must be a contiguous ({B}, {L}, 2) tensor of dtype {rdt}")
    return _FusedHeadPropagateLoss.apply(logits.to(rdt).contiguous(), raw_target(U_target, rdt), error, M,
                                         tuple(float(x) for x in sigma), int(seed), int(offset), loss, tau, k,
                                         int(flags) | FLAG_RAW_TARGET, 0 if head == "transformer" else 1,
                                         (float(lo0), float(hi0), float(lo1), float(hi1)), scale, off, base, pulses_out, F_out,
                                         err_out, workspace)


class _PropagateFidelity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pulses, tc, error, M, sigma, seed, offset, flags):
        B, L, _ = pulses.shape
        F = torch.empty(B * M, dtype=pulses.dtype, device=pulses.device)
        _launch_forward(pulses, tc, error, M, 0, sigma, seed, offset, None, F, None, None, flags)
        ctx.save_for_backward(pulses, tc, error if error is not None else pulses.new_empty(0))
        ctx.args = (M, sigma, seed, offset, flags, error is not None)
        return F

    @staticmethod
    def backward(ctx, gF):
        pulses, tc, error = ctx.saved_tensors
        M, sigma, seed, offset, flags, has_err = ctx.args
        B, L, _ = pulses.shape
        Fsum = torch.empty(B, dtype=pulses.dtype, device=pulses.device)
        G = torch.empty(B, L, 2, dtype=pulses.dtype, device=pulses.device)
        _launch_fwdbwd(pulses, tc, error if has_err else None, gF.to(pulses.dtype).contiguous(), M, 0, sigma, seed, offset,
                       None, None, Fsum, G, flags)
        return (G,) + (None,) * 7


def propagate_fidelity(pulses: torch.Tensor, U_target: torch.Tensor, error: Optional[torch.Tensor], monte_carlo: int, *,
                       sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, offset: int = 0,
                       dtype: Optional[torch.dtype] = None, fast_sincos: bool = False, flags: int = 0) -> torch.Tensor:
    """Per-sample fidelity F (B*M,) of un-repeated pulses, differentiable w.r.t. ``pulses`` for an
    arbitrary downstream loss (the backward kernel takes dLoss/dF_s as per-sample weights)."""
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")
    _require_cuda(pulses, "pulses")
    rdt = dtype or _real_dtype(pulses)
    B = pulses.shape[0]
    M = int(monte_carlo)
    if error is not None:
        _require_cuda(error, "error")
        if error.shape != (2, B * M):
            raise ValueError(f"'error' must have shape (2, {B * M}), got {tuple(error.shape)}")
        error = error.to(rdt).contiguous()
    fl = flags | (FLAG_FAST_SINCOS if fast_sincos else 0)
    return _PropagateFidelity.apply(pulses.to(rdt).contiguous(), target_coeffs(U_target, rdt), error, M,
                                    tuple(float(s) for s in sigma), int(seed), int(offset), fl)


# ----------------------------------------------------------------------------- reference-signature adapters
class _Generator(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pulses, error):
        Bm, L, _ = pulses.shape
        U = torch.empty(Bm, 2, 2, 2, dtype=pulses.dtype, device=pulses.device)
        _call("uqoc_su2_generator_forward", pulses.device, _ptr(pulses), _ptr(error), Bm, L, _ptr(U), _dt(pulses), 0,
                                                    _stream(pulses.device))
        ctx.save_for_backward(pulses, error)
        return torch.view_as_complex(U)

    @staticmethod
    def backward(ctx, gU):
        pulses, error = ctx.saved_tensors
        Bm, L, _ = pulses.shape
        g = torch.view_as_real(gU.resolve_conj()).to(pulses.dtype).contiguous()
        gp = torch.empty_like(pulses)
        _call("uqoc_su2_generator_backward", pulses.device, _ptr(pulses), _ptr(error), _ptr(g), Bm, L, _ptr(gp), _dt(pulses), 0,
                                                     _stream(pulses.device))
        return gp, None


def batched_unitary_generator(pulses: torch.Tensor, error: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``batched_unitary_generator`` (SCORE.py:77-145 / grape_train.py:78-138).

    pulses (Bm, L, 2) [phi, tau]; error (2, Bm) [delta; eps] -> (Bm, 2, 2) complex64
    (complex128 for float64 inputs), differentiable w.r.t. ``pulses``.  A stride-0
    ``expand``-ed pulse sequence (visualize/util.py:245) takes the shared-pulse kernel.
    """
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")
    _require_cuda(pulses, "pulses")
    _require_cuda(error, "error")
    Bm, L, _ = pulses.shape
    if error.ndim != 2 or error.shape[0] != 2 or error.shape[1] != Bm:
        raise ValueError(f"'error' must have shape (2, {Bm})")
    rdt = torch.float64 if (pulses.dtype == torch.float64 or error.dtype == torch.float64) else torch.float32
    err = error.to(rdt).contiguous()
    grad_needed = pulses.requires_grad and torch.is_grad_enabled()
    if Bm > 1 and pulses.stride(0) == 0 and not grad_needed:
        shared = pulses[0:1].to(rdt).contiguous()
        U = torch.empty(Bm, 2, 2, 2, dtype=rdt, device=pulses.device)
        eye = torch.eye(2, dtype=torch.complex64 if rdt == torch.float32 else torch.complex128, device=pulses.device)[None]
        _launch_forward(shared, target_coeffs(eye, rdt), err, Bm, 0, (0.0, 0.0), 0, 0, U, None, None, None, 0)
        return torch.view_as_complex(U)
    return _Generator.apply(pulses.to(rdt).contiguous(), err)


class _Fidelity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, T, d, t_stride):
        Bm = U.shape[0]
        F = torch.empty(Bm, dtype=U.dtype, device=U.device)
        _call("uqoc_fidelity_forward", U.device, _ptr(U), _ptr(T), Bm, d, t_stride, _ptr(F), _dt(U), _stream(U.device))
        ctx.save_for_backward(U, T)
        ctx.args = (d, t_stride)
        return F

    @staticmethod
    def backward(ctx, gF):
        U, T = ctx.saved_tensors
        d, t_stride = ctx.args
        gU = torch.empty_like(U)
        _call("uqoc_fidelity_backward", U.device, _ptr(U), _ptr(T), _ptr(gF.to(U.dtype).contiguous()), U.shape[0], d, t_stride,
                                                _ptr(gU), _dt(U), _stream(U.device))
        return gU, None, None, None


def fidelity(U_out: torch.Tensor, U_target: torch.Tensor, num_qubits: int) -> torch.Tensor:
    """Drop-in for ``fidelity`` (SCORE.py:168-183): (|Tr(U_out^dagger U_target)|^2 + d) / (d(d+1))."""
    _require_cuda(U_out, "U_out")
    _require_cuda(U_target, "U_target")
    d = 2 ** num_qubits
    if U_out.ndim != 3 or U_out.shape[-2:] != (d, d):
        raise ValueError(f"'U_out' must have shape (B, {d}, {d})")
    rdt = _real_dtype(U_out)
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    Bm = U_out.shape[0]
    Ur = torch.view_as_real(U_out.to(cdt).resolve_conj().contiguous())
    if U_target.ndim == 2 or (U_target.ndim == 3 and (U_target.shape[0] == 1 or U_target.stride(0) == 0)):
        T = U_target.reshape(-1, d, d)[0:1].to(cdt).resolve_conj().contiguous()
        t_stride = 0
    else:
        if U_target.shape[0] != Bm:
            raise ValueError("U_target batch does not match U_out")
        T = U_target.to(cdt).resolve_conj().contiguous()
        t_stride = 2 * d * d
    return _Fidelity.apply(Ur, torch.view_as_real(T), d, t_stride)


class _MeanLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, F, kind, tau, k):
        n = F.numel()
        S = torch.empty(1, dtype=F.dtype, device=F.device)
        ws = _workspace(1024 * 8, F.device)
        _call("uqoc_sum", F.device, _ptr(F), n, _ptr(S), _ptr(ws), ws.numel(), _dt(F), _stream(F.device))
        out = _finalize(S, n, kind, tau, k, None)
        ctx.save_for_backward(out)
        ctx.shape = F.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        n = 1
        for s in ctx.shape:
            n *= s
        return (g * out[2] / n).expand(ctx.shape), None, None, None


def _mean_loss(F: torch.Tensor, kind: str, tau: float = 0.99, k: float = 100) -> torch.Tensor:
    _require_cuda(F, "fidelity")
    if F.dtype not in (torch.float32, torch.float64):
        F = F.float()
    return _MeanLoss.apply(F.contiguous(), kind, tau, k)


def negative_log_loss(U_out, U_target, fidelity_fn, num_qubits):
    """Drop-in for SCORE.py:185-186: -log(mean F)."""
    return _mean_loss(fidelity_fn(U_out, U_target, num_qubits), "nll")


def infidelity_loss(U_out, U_target, fidelity_fn, num_qubits):
    """Drop-in for SCORE.py:189-190: 1 - mean F."""
    return _mean_loss(fidelity_fn(U_out, U_target, num_qubits), "infidelity")


def sharp_loss(U_out, U_target, fidelity_fn, num_qubits, tau=0.99, k=100):
    """Drop-in for SCORE.py:193-195: custom_loss(mean F, tau, k)."""
    return _mean_loss(fidelity_fn(U_out, U_target, num_qubits), "sharp", tau, k)


def custom_loss(x: torch.Tensor, tau=0.99, k=100):
    """Drop-in for SCORE.py:197-198: log(1+exp(-k(x-tau))) * (1-x) of a scalar fidelity."""
    if x.numel() != 1:
        raise ValueError("custom_loss expects the scalar mean fidelity (SCORE.py:194)")
    return _mean_loss(x.reshape(1), "sharp", tau, k)


_SAMPLER_STATE = {"calls": 0}


def get_ore_ple_error_distribution(batch_size: int, delta_std=1.0, epsilon_std=0.05, *, device="cuda",
                                   dtype=torch.float32, seed: Optional[int] = None,
                                   offset: Optional[int] = None) -> torch.Tensor:
    """Drop-in for SCORE.py:158-161, generated on the device: (2, batch_size) with row 0 ~
    N(0, delta_std^2) and row 1 ~ N(0, epsilon_std^2).  The stream is Philox keyed by
    ``seed`` (default ``torch.initial_seed()``) and a per-call ``offset`` counter, so runs are
    reproducible under ``torch.manual_seed`` without a host RNG or an H2D copy."""
    if seed is None:
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if offset is None:
        offset = _SAMPLER_STATE["calls"]
        _SAMPLER_STATE["calls"] += 1
    return philox_errors(1, int(batch_size), (float(delta_std), float(epsilon_std)), seed, offset, 0, device, dtype)


def get_ore_error_distribution(batch_size: int, delta_std=1.0, *, device="cuda", dtype=torch.float32,
                               seed: Optional[int] = None, offset: Optional[int] = None) -> torch.Tensor:
    """Drop-in for SCORE.py:154-155 (legacy 1-D ORE sampler)."""
    return get_ore_ple_error_distribution(batch_size, delta_std, 0.0, device=device, dtype=dtype, seed=seed,
                                          offset=offset)[0]


# ----------------------------------------------------------------------------- two-qubit SU(4) path
# NOT in the reference (README.md:86 promises it); builder-defined contract, see include/uqoc.h.
def _su4_target(U_target: torch.Tensor, rdt: torch.dtype, B: int) -> torch.Tensor:
    _require_cuda(U_target, "U_target")
    if U_target.ndim != 3 or U_target.shape[-2:] != (4, 4) or U_target.shape[0] != B:
        raise ValueError(f"'U_target' must have shape ({B}, 4, 4)")
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    return torch.view_as_real(U_target.to(cdt).resolve_conj().contiguous()).contiguous()


def su4_workspace(B: int, L: int, M: int, dtype: torch.dtype, flags: int, device) -> torch.Tensor:
    """A private workspace for one (B, L, M) SU(4) launch shape (CUDA-graph owners)."""
    dev = torch.device(device)
    with _on_device(dev):
        n = int(_lib.lib().uqoc_su4_workspace_bytes(B, L, M, F64 if dtype == torch.float64 else F32, flags))
    return alloc_workspace(n, dev)


def _su4_launch(bwd, pulses, tgt, error, weight, M, j0, J, sigma, seed, offset, U_out, F_out, err_out, Fsum, G, flags, ws=None):
    B, L, _ = pulses.shape
    dt = _dt(pulses)
    dev = pulses.device
    with _on_device(dev):
        ws_bytes = int(_lib.lib().uqoc_su4_workspace_bytes(B, L, M, dt, flags))
    if ws is None:
        ws = _workspace(ws_bytes, dev)
    elif ws.numel() < ws_bytes:
        raise ValueError(f"workspace of {ws.numel()} bytes is too small for this launch shape ({ws_bytes} bytes)")
    if bwd:
        _call("uqoc_su4_fwdbwd", dev, _ptr(pulses), _ptr(tgt), _ptr(error), _ptr(weight), B, L, M, j0, float(J), float(sigma[0]),
              float(sigma[1]), seed, offset, _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(G), _ptr(ws), ws_bytes, dt, flags,
              _stream(dev))
    else:
        _call("uqoc_su4_forward", dev, _ptr(pulses), _ptr(tgt), _ptr(error), B, L, M, j0, float(J), float(sigma[0]),
              float(sigma[1]), seed, offset, _ptr(U_out), _ptr(F_out), _ptr(err_out), _ptr(Fsum), _ptr(ws), ws_bytes, dt, flags,
              _stream(dev))


class _FusedPropagateLossSU4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pulses, tgt, error, M, j0, M_total, J, sigma, seed, offset, loss, tau, k, flags, group, F_out, err_out,
                ws=None):
        B, L, _ = pulses.shape
        need_grad = ctx.needs_input_grad[0]
        n_g = B * L * 3 if need_grad else 0
        buf = torch.empty(n_g + B, dtype=pulses.dtype, device=pulses.device)
        G, Fsum = (buf[:n_g] if need_grad else None), buf[n_g:]
        _su4_launch(need_grad, pulses, tgt, error, None, M, j0, J, sigma, seed, offset, None, F_out, err_out, Fsum, G, flags, ws)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        loss_out = _finalize(Fsum, B * M_total, loss, tau, k, G)
        mean_fid = Fsum / M_total
        if need_grad:
            ctx.save_for_backward(G.view(B, L, 3))
        ctx.mark_non_differentiable(mean_fid)
        return loss_out[0], mean_fid

    @staticmethod
    def backward(ctx, g_loss, _g_mean):
        (G,) = ctx.saved_tensors
        return (g_loss * G,) + (None,) * 17


def fused_propagate_loss_su4(pulses: torch.Tensor, U_target: torch.Tensor, *, error: Optional[torch.Tensor] = None,
                             monte_carlo: int, J: float = 1.0, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0,
                             offset: int = 0, loss: str = "sharp", tau: float = 0.99, k: float = 100,
                             dtype: Optional[torch.dtype] = None, flags: int = 0, group=None,
                             F_out: Optional[torch.Tensor] = None, err_out: Optional[torch.Tensor] = None,
                             workspace: Optional[torch.Tensor] = None):
    """Two-qubit twin of :func:`fused_propagate_loss`: pulses (B, L, 3) = [phi1, phi2, tau], U_target (B, 4, 4),
    error (3, B*M) = [delta1; delta2; eps] or None (Philox), coupling ``J`` of the ZZ term.  ``workspace``: a private
    buffer from :func:`su4_workspace` for callers that capture the launch in a CUDA graph."""
    if pulses.ndim != 3 or pulses.shape[-1] != 3:
        raise ValueError("'pulses' must have shape (B, L, 3)")
    _require_cuda(pulses, "pulses")
    if loss not in LOSS_KINDS:
        raise ValueError(f"unknown loss {loss!r}; expected one of {sorted(LOSS_KINDS)}")
    rdt = dtype or _real_dtype(pulses)
    B = pulses.shape[0]
    M_total = int(monte_carlo)
    tgt = _su4_target(U_target, rdt, B)
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    j0, M = shard_range(M_total, rank, world)
    if error is not None:
        _require_cuda(error, "error")
        if error.shape != (3, B * M_total):
            raise ValueError(f"'error' must have shape (3, {B * M_total}), got {tuple(error.shape)}")
        error = shard_errors(error.to(rdt), B, M_total, j0, M).contiguous()
    return _FusedPropagateLossSU4.apply(pulses.to(rdt).contiguous(), tgt, error, M, j0, M_total, float(J),
                                        tuple(float(x) for x in sigma), int(seed), int(offset), loss, tau, k, flags, group,
                                        F_out, err_out, workspace)


class _GeneratorSU4(torch.autograd.Function):
    """Per-sample pulse rows: forward = uqoc_su4_forward with B = Bm, M = 1; backward = uqoc_su4_generator_backward."""

    @staticmethod
    def forward(ctx, pulses, err, J, flags):
        Bm = pulses.shape[0]
        rdt = pulses.dtype
        cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
        U = torch.empty(Bm, 4, 4, 2, dtype=rdt, device=pulses.device)
        tgt = _su4_target(torch.eye(4, dtype=cdt, device=pulses.device)[None].expand(Bm, -1, -1), rdt, Bm)
        _su4_launch(False, pulses, tgt, err, None, 1, 0, J, (0.0, 0.0), 0, 0, U, None, None, None, None, flags)
        ctx.save_for_backward(pulses, err)
        ctx.args = (J, flags)
        return torch.view_as_complex(U)

    @staticmethod
    def backward(ctx, gU):
        pulses, err = ctx.saved_tensors
        J, flags = ctx.args
        Bm, L, _ = pulses.shape
        g = torch.view_as_real(gU.resolve_conj()).to(pulses.dtype).contiguous()
        gp = torch.empty_like(pulses)
        _call("uqoc_su4_generator_backward", pulses.device, _ptr(pulses), _ptr(err), _ptr(g), Bm, L, float(J), _ptr(gp),
              _dt(pulses), flags, _stream(pulses.device))
        return gp, None, None, None


def su4_unitary_generator(pulses: torch.Tensor, error: torch.Tensor, J: float = 1.0, flags: int = 0) -> torch.Tensor:
    """Two-qubit generator with the reference's calling convention: pulses (Bm, L, 3), error (3, Bm) ->
    (Bm, 4, 4) complex, differentiable w.r.t. ``pulses`` (one block per sample row: correct, not tuned -- the fused
    :func:`fused_propagate_loss_su4` is the fast path).  ``expand``-ed (stride-0) pulses that need no gradient share one
    staged pulse train."""
    if pulses.ndim != 3 or pulses.shape[-1] != 3:
        raise ValueError("'pulses' must have shape (B, L, 3)")
    _require_cuda(pulses, "pulses")
    _require_cuda(error, "error")
    Bm = pulses.shape[0]
    if error.ndim != 2 or error.shape != (3, Bm):
        raise ValueError(f"'error' must have shape (3, {Bm})")
    rdt = torch.float64 if (pulses.dtype == torch.float64 or error.dtype == torch.float64) else torch.float32
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    err = error.to(rdt).contiguous()
    grad_needed = pulses.requires_grad and torch.is_grad_enabled()
    if Bm > 1 and pulses.stride(0) == 0 and not grad_needed:
        U = torch.empty(Bm, 4, 4, 2, dtype=rdt, device=pulses.device)
        tgt = _su4_target(torch.eye(4, dtype=cdt, device=pulses.device)[None], rdt, 1)
        _su4_launch(False, pulses[0:1].to(rdt).contiguous(), tgt, err, None, Bm, 0, J, (0.0, 0.0), 0, 0, U, None, None, None,
                    None, flags)
        return torch.view_as_complex(U)
    return _GeneratorSU4.apply(pulses.to(rdt).contiguous(), err, float(J), int(flags) & ~64)   # eigenframe kernel both ways


def philox_errors_su4(B: int, M: int, sigma: Sequence[float] = (1.0, 0.05), seed: int = 0, offset: int = 0, j0: int = 0,
                      device="cuda", dtype=torch.float32) -> torch.Tensor:
    """(3, B*M) = [delta1; delta2; eps] of the SU(4) kernels' Philox stream (reported through err_out)."""
    dev = torch.device(device)
    pulses = torch.zeros(B, 1, 3, dtype=dtype, device=dev)
    cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
    tgt = _su4_target(torch.eye(4, dtype=cdt, device=dev)[None].expand(B, -1, -1), dtype, B)
    out = torch.empty(3, B * M, dtype=dtype, device=dev)
    _su4_launch(False, pulses, tgt, None, None, M, j0, 0.0, sigma, seed, offset, None, None, out, None, None, 0)
    return out
