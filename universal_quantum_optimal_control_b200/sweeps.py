"""Forward-only disorder sweeps: the call patterns of
``visualize/util.py:209-271`` (``fidelity_contour_plot``) and ``:280-326`` (``get_avg_fidelity`` /
``plot_fidelity_by_std``) as single kernel launches."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib
from ._lib import check
from .ops import _dt, _ptr, _real_dtype, _require_cuda, _stream, _workspace, target_coeffs

__all__ = ["fidelity_grid", "fidelity_vs_sigma"]


def fidelity_grid(pulse: torch.Tensor, U_target: torch.Tensor, delta_axis: torch.Tensor, eps_axis: torch.Tensor,
                  *, dtype=None, flags: int = 0) -> torch.Tensor:
    """F on the grid ``meshgrid(delta_axis, eps_axis, indexing="ij")`` for one pulse sequence.

    Equivalent to ``util.py:231-249``: ``errors_grid = stack([ORE_grid.flatten(), PLE_grid.flatten()])``,
    ``g(pulse.expand(N,-1,-1), errors_grid)``, ``fidelity(...).reshape(Nd, Ne)`` -- without building
    the (2, N) error tensor or the (N, 2, 2) unitaries.  pulse (L, 2) or (B, L, 2); returns (Nd, Ne)
    or (B, Nd, Ne)."""
    _require_cuda(pulse, "pulse")
    squeeze = pulse.ndim == 2
    if squeeze:
        pulse = pulse[None]
        U_target = U_target.reshape(1, 2, 2)
    if pulse.ndim != 3 or pulse.shape[-1] != 2:
        raise ValueError("'pulse' must have shape (L, 2) or (B, L, 2)")
    rdt = dtype or _real_dtype(pulse)
    B, L, _ = pulse.shape
    nd, ne = int(delta_axis.numel()), int(eps_axis.numel())
    axes = torch.cat([delta_axis.reshape(-1).to(pulse.device, rdt), eps_axis.reshape(-1).to(pulse.device, rdt)]).contiguous()
    p = pulse.to(rdt).contiguous()
    tc = target_coeffs(U_target.to(pulse.device), rdt)
    F = torch.empty(B, nd, ne, dtype=rdt, device=pulse.device)
    lib = _lib.lib()
    ws_bytes = lib.uqoc_su2_workspace_bytes(B, L, nd * ne, _dt(p), flags)
    ws = _workspace(ws_bytes, pulse.device)
    esz = axes.element_size()
    check(lib.uqoc_su2_forward_grid(_ptr(p), _ptr(tc), axes.data_ptr(), nd, axes.data_ptr() + nd * esz, ne, B, L, None,
                                    _ptr(F), None, _ptr(ws), ws_bytes, _dt(p), flags, _stream(pulse.device)),
          "uqoc_su2_forward_grid")
    return F[0] if squeeze else F


def fidelity_vs_sigma(pulse: torch.Tensor, U_target: torch.Tensor, delta_stds: Sequence[float], epsilon_std: float = 0.05,
                      M: int = 10000, *, seed: int = 0, offset: int = 0, dtype=None,
                      flags: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean fidelity and its standard error for every ``delta_std`` in one launch
    (``util.py:313-326`` runs 199 Python iterations of M = 10000 samples each).
    Returns ``(F_mean (S,), F_err (S,))`` with ``F_err = std / sqrt(M)`` as ``util.py:296-297``."""
    _require_cuda(pulse, "pulse")
    if pulse.ndim != 2 or pulse.shape[-1] != 2:
        raise ValueError("'pulse' must have shape (L, 2)")
    rdt = dtype or _real_dtype(pulse)
    sig = torch.as_tensor(list(delta_stds), dtype=rdt, device=pulse.device).reshape(-1)
    S = sig.numel()
    table = torch.stack([sig, torch.full_like(sig, float(epsilon_std))], dim=1).contiguous()
    p = pulse.to(rdt)[None].expand(S, -1, -1).contiguous()
    tc = target_coeffs(U_target.reshape(1, 2, 2).to(pulse.device).expand(S, -1, -1), rdt)
    F = torch.empty(S, M, dtype=rdt, device=pulse.device)
    lib = _lib.lib()
    L = pulse.shape[0]
    ws_bytes = lib.uqoc_su2_workspace_bytes(S, L, M, _dt(p), flags)
    ws = _workspace(ws_bytes, pulse.device)
    check(lib.uqoc_su2_forward_sigmas(_ptr(p), _ptr(tc), _ptr(table), S, L, M, 0, seed, offset, _ptr(F), None, _ptr(ws),
                                      ws_bytes, _dt(p), flags, _stream(pulse.device)), "uqoc_su2_forward_sigmas")
    return F.mean(dim=1), F.std(dim=1) / (M ** 0.5)
