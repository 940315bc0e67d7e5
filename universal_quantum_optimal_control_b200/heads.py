"""Pulse heads: the element-wise tail of the reference's two pulse generators
as one CUDA launch forward and one backward, instead of 6-8 ATen kernels per step.

* :func:`transformer_pulse_head` -- ``model/universal_model.py:131-143``: sigmoid, range map, optional
  finetune base pulse, relu on tau, add the target azimuth to phi, wrap phi to [-pi, pi).
* :func:`grape_pulse_head` -- ``model/GRAPE_model.py:76-89``: sigmoid, ``phi = atan2(u_y, u_x)``, range map,
  relu on tau.

Folded into the fused step (no pulses tensor between model and op): :func:`ops.fused_head_propagate_loss`.
:class:`HeadlessGRAPE` / :class:`HeadlessTransformer` wrap the reference's (unmodified) model objects so that a
trainer gets the LOGITS of their last linear layer plus the head description, which is what the fused step takes.
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


class HeadSpecPy:
    """What the fused step needs to know about a model's element-wise tail."""

    def __init__(self, kind: str, pulse_ranges, base_pulse: Optional[torch.Tensor] = None, scale: float = 0.2):
        self.kind, self.base_pulse, self.scale = kind, base_pulse, float(scale)
        self.pulse_ranges = tuple(tuple(float(v) for v in r) for r in pulse_ranges)


class HeadlessGRAPE(torch.nn.Module):
    """The reference's ``GRAPE`` (``model/GRAPE_model.py``) up to its last linear layer.  ``logits(rv)`` is
    ``model.layer(rv).reshape(B, L, 3)`` (``GRAPE_model.py:77-78``); ``forward`` reproduces ``GRAPE.forward`` through
    the one-launch head, so the wrapper is a drop-in for the wrapped model wherever pulses are wanted."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.num_qubits = getattr(model, "num_qubits", 1)
        self.uqoc_head = HeadSpecPy("grape", model.param_ranges.tolist())

    def logits(self, rotation_vector: torch.Tensor):
        x = self.model.layer(rotation_vector)
        return x.reshape(rotation_vector.shape[0], self.model.pulse_length, 3), None

    def forward(self, rotation_vector: torch.Tensor) -> torch.Tensor:
        return grape_pulse_head(self.logits(rotation_vector)[0], self.uqoc_head.pulse_ranges)


class HeadlessTransformer(torch.nn.Module):
    """The reference's ``UniversalQOCTransformer`` up to ``head`` (``model/universal_model.py:83-128``): the encoder
    input is assembled with the model's own static helpers, ``logits`` returns the last sequence position reshaped to
    (B, max_pulses, 2) and the target azimuth that the tail adds to phi (``:92``, ``:141``).  A ``finetune`` base pulse is
    loaded ONCE here (the reference ``torch.load``s it inside every forward, ``:135-138``)."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.num_qubits = getattr(model, "num_qubits", 1)
        base = None
        if getattr(model, "finetune", None):
            base = torch.load(model.finetune)
        self.uqoc_head = HeadSpecPy("transformer", model.param_ranges.tolist(), base, 0.2)

    def logits(self, rotation_vector: torch.Tensor):
        m = self.model
        rv = rotation_vector
        B = rv.shape[0]
        azimuth = torch.atan2(rv[:, 1], rv[:, 0])
        rescaled = torch.stack([torch.sqrt(rv[:, 0] ** 2 + rv[:, 1] ** 2), torch.zeros(B, device=rv.device), rv[:, 2], rv[:, 3]], 1)
        cls = type(m)
        seq = cls.score_sequence_from_yxy(cls.euler_yxy_from_rotation_vector(rescaled))          # (B, 9, 2, 2)
        emb = m.unitary_proj(cls._to_real_vector(seq).to(torch.float).to(rv.device))
        emb = emb + cls.sinusoidal_positional_encoding(9, m.d_model, device=emb.device).unsqueeze(0)
        x = m.head(m.encoder(emb))[:, -1, :]
        return x.view(B, m.max_pulses, m.param_dim), azimuth

    def forward(self, rotation_vector: torch.Tensor) -> torch.Tensor:
        x, az = self.logits(rotation_vector)
        return transformer_pulse_head(x, self.uqoc_head.pulse_ranges, az, self.uqoc_head.base_pulse, self.uqoc_head.scale)
