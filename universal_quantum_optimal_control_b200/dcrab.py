"""dCRAB objective on the GPU: NumPy-facing drop-in for
``train/dCRAB/dCRAB.py:26-59`` so that Nelder-Mead evaluates its 200-sample objective in one forward
launch instead of ``S x len(t)`` calls of ``scipy.linalg.expm``.

The same Hamiltonian as the trainer path (``dCRAB.py:41-42``: ``H = (cos phi X + sin phi Y + delta Z)(1+eps)/2``)
on a uniform time grid.  The reference's objective uses ``(|Tr| + 2)/6`` -- the trace modulus is NOT
squared (``dCRAB.py:58``); that variant is preserved by default and flagged (``squared=False``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def build_phi(params, t, omegas):
    """``dCRAB.py:26-34``: phi(t) = phi0 + sum_n a_n cos(w_n t) + b_n sin(w_n t)."""
    params, t, omegas = np.asarray(params, dtype=np.float64), np.asarray(t, dtype=np.float64), np.asarray(omegas, dtype=np.float64)
    N = len(omegas)
    a, b = params[1:1 + N], params[1 + N:1 + 2 * N]
    wt = np.outer(omegas, t)
    return params[0] + a @ np.cos(wt) + b @ np.sin(wt)


def propagate(phi_vals, t, deltas, epss, device="cuda", dtype=torch.float64):
    """Batched ``dCRAB.py:37-44`` for all error samples at once -> (S, 2, 2) complex tensor on the GPU."""
    dt = float(t[1] - t[0])
    phi = torch.as_tensor(np.asarray(phi_vals), dtype=dtype, device=device)
    pulse = torch.stack([phi, torch.full_like(phi, dt)], dim=-1)
    err = torch.as_tensor(np.stack([np.asarray(deltas), np.asarray(epss)]), dtype=dtype, device=device)
    S = err.shape[1]
    return ops.batched_unitary_generator(pulse[None].expand(S, -1, -1), err)


def average_infidelity(params, t, omegas, U_target, deltas, epss, X=None, Y=None, Z=None, *, squared: bool = False,
                       device="cuda", dtype=torch.float64) -> float:
    """``dCRAB.py:47-59`` (same positional signature; X, Y, Z are accepted and ignored -- the Paulis are
    baked into the kernel).  ``squared=False`` keeps the reference's ``(|Tr| + 2)/6``; ``squared=True`` gives
    the trainer's ``(|Tr|^2 + 2)/6`` (``SCORE.py:168-183``)."""
    U = propagate(build_phi(params, t, omegas), t, deltas, epss, device, dtype)
    T = torch.as_tensor(np.asarray(U_target), device=device).to(U.dtype)
    F2 = ops.fidelity(U, T[None].expand(U.shape[0], -1, -1), 1)          # (|tr|^2 + 2)/6
    if squared:
        return float(1.0 - F2.mean().item())
    tr_abs = torch.sqrt(torch.clamp(6.0 * F2 - 2.0, min=0.0))
    return float(1.0 - ((tr_abs + 2.0) / 6.0).mean().item())
