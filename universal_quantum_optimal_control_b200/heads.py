"""Pulse heads: the element-wise tail of the reference's two pulse generators
as one CUDA launch forward and one backward, instead of 6-8 ATen kernels per step.

* :func:`transformer_pulse_head` -- ``model/universal_model.py:131-143``: sigmoid, range map, optional
  finetune base pulse, relu on tau, add the target azimuth to phi, wrap phi to [-pi, pi).
* :func:`grape_pulse_head` -- ``model/GRAPE_model.py:76-89``: sigmoid, ``phi = atan2(u_y, u_x)``, range map,
  relu on tau.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .ops import _dt, _ptr, _require_cuda, _stream


def _ranges(r):
    (lo0, hi0), (lo1, hi1) = r
    return (C.c_double * 4)(float(lo0), float(hi0), float(lo1), float(hi1))


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, offset, base, mode, ranges, scale):
        B, L, _ = logits.shape
        out = torch.empty(B, L, 2, dtype=logits.dtype, device=logits.device)
        check(_lib.lib().uqoc_pulse_head_forward(_ptr(logits), _ptr(offset), _ptr(base), B, L, mode, _ranges(ranges), scale,
                                                 _ptr(out), _dt(logits), _stream(logits.device)), "uqoc_pulse_head_forward")
        ctx.save_for_backward(logits, base if base is not None else logits.new_empty(0))
        ctx.args = (mode, ranges, scale, base is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        logits, base = ctx.saved_tensors
        mode, ranges, scale, has_base = ctx.args
        B, L, _ = logits.shape
        gl = torch.empty_like(logits)
        check(_lib.lib().uqoc_pulse_head_backward(_ptr(logits), _ptr(base) if has_base else None, _ptr(g.contiguous()), B, L,
                                                  mode, _ranges(ranges), scale, _ptr(gl), _dt(logits), _stream(logits.device)),
              "uqoc_pulse_head_backward")
        return gl, None, None, None, None, None


def _prep(logits, P):
    _require_cuda(logits, "logits")
    if logits.ndim != 3 or logits.shape[-1] != P:
        raise ValueError(f"'logits' must have shape (B, L, {P})")
    if logits.dtype not in (torch.float32, torch.float64):
        logits = logits.float()
    return logits.contiguous()


def transformer_pulse_head(logits: torch.Tensor, pulse_ranges: Sequence[Sequence[float]],
                           phi_offset: Optional[torch.Tensor] = None, base_pulse: Optional[torch.Tensor] = None,
                           scale: float = 0.2) -> torch.Tensor:
    """(B, L, 2) head outputs -> (B, L, 2) pulses [phi, tau]; ``pulse_ranges = ((lo_phi, hi_phi), (lo_tau, hi_tau))``
    (``model_params.json`` ``pulse_space``), ``phi_offset`` (B,) the target azimuth (``umodel.py:141``),
    ``base_pulse`` (L, 2) the finetune base (``umodel.py:135-138``, pulses = scale * pulses + base)."""
    x = _prep(logits, 2)
    off = None if phi_offset is None else phi_offset.to(x.device, x.dtype).contiguous()
    base = None if base_pulse is None else base_pulse.to(x.device, x.dtype).contiguous()
    return _Head.apply(x, off, base, 0, tuple(tuple(r) for r in pulse_ranges), float(scale))


def grape_pulse_head(logits: torch.Tensor, pulse_ranges: Sequence[Sequence[float]]) -> torch.Tensor:
    """(B, L, 3) MLP outputs [u_x, u_y, u_tau] -> (B, L, 2) pulses (``GRAPE_model.py:76-89``)."""
    x = _prep(logits, 3)
    return _Head.apply(x, None, None, 1, tuple(tuple(r) for r in pulse_ranges), 1.0)
