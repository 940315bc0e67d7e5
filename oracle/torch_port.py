"""Torch restatement of the reference's CPU implementation of the hot path.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/uqoc_oracle.py header): used by
``tests/`` as a second checker (float64 autograd) and by ``bench.py`` as the timed
CPU baseline (``cpu_baseline.kind == "port"`` and ``--impl reference``).  The
product never imports it.

Unlike oracle/uqoc_oracle.py (closed-form numpy), this port issues the SAME
sequence of ATen operations the reference does -- element-wise Hamiltonian build,
``torch.linalg.matrix_exp``, a log2(L)-deep batched-matmul tree, the fidelity
einsum, the softplus-weighted loss and autograd's backward through all of it --
so that its wall-clock on the host's cores is the reference's own cost for this
path.  Reference lines (paths relative to upstream root):

  SCORE.py  = train/unitary_single_qubit_gate/universal_single_qubit_SCORE.py
  trainer.py= model/universal_model_trainer.py
"""
from __future__ import annotations

import torch

_PAULI = None


def _paulis(cdtype, device):
    # SCORE.py:53-70 (I, X, Y, Z stack)
    global _PAULI
    if _PAULI is None:
        _PAULI = torch.tensor(
            [[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]],
            dtype=torch.complex128)
    return _PAULI.to(device=device, dtype=cdtype)


def generator_tree(pulses: torch.Tensor, error: torch.Tensor) -> torch.Tensor:
    """SCORE.py:77-145.  pulses (Bm,L,2) [phi,tau], error (2,Bm) -> (Bm,2,2)."""
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")
    Bm = pulses.shape[0]
    cdtype = torch.complex64 if pulses.dtype == torch.float32 else torch.complex128
    sig = _paulis(cdtype, pulses.device)
    phi = pulses[..., 0]
    tau = pulses[..., 1]
    delta, eps = error[0], error[1]
    ham = torch.cos(phi)[..., None, None] * sig[1] + torch.sin(phi)[..., None, None] * sig[2]   # :117-120
    ham = ham + delta[:, None, None, None] * sig[3]                                             # :122
    ham = 0.5 * ham * (1 + eps[:, None, None, None])                                            # :124
    steps = torch.linalg.matrix_exp(-1j * ham * tau[..., None, None])                           # :127
    eye = torch.eye(2, dtype=cdtype, device=pulses.device).expand(Bm, 1, 2, 2)
    level = steps
    while level.size(1) > 1:                                                                    # :134-140
        if level.size(1) % 2 == 1:
            level = torch.cat([level, eye], dim=1)
        level = level[:, 1::2] @ level[:, 0::2]
    return level[:, 0]


def generator_sequential(pulses: torch.Tensor, error: torch.Tensor) -> torch.Tensor:
    """train/GRAPE/grape_train.py:78-138 (running product, :133-136)."""
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")
    Bm, L, _ = pulses.shape
    cdtype = torch.complex64 if pulses.dtype == torch.float32 else torch.complex128
    sig = _paulis(cdtype, pulses.device)
    phi, tau = pulses[..., 0], pulses[..., 1]
    delta, eps = error[0], error[1]
    ham = torch.cos(phi)[..., None, None] * sig[1] + torch.sin(phi)[..., None, None] * sig[2]
    ham = ham + delta[:, None, None, None] * sig[3]
    ham = 0.5 * ham * (1 + eps[:, None, None, None])
    steps = torch.linalg.matrix_exp(-1j * ham * tau[..., None, None])
    out = torch.eye(2, dtype=cdtype, device=pulses.device).expand(Bm, 2, 2)
    for k in range(L):
        out = steps[:, k] @ out
    return out


def fidelity(U_out, U_target, num_qubits: int):
    """SCORE.py:168-183."""
    prod = U_out.conj().transpose(-1, -2) @ U_target
    tr = torch.einsum("bii->b", prod)
    d = 2 ** num_qubits
    return (tr.abs() ** 2 + d) / (d * (d + 1))


def softplus_weighted(x, tau=0.99, k=100):
    """SCORE.py:197-198."""
    return torch.log(1 + torch.exp(-k * (x - tau))) * (1 - x)


def loss_value(F, loss="sharp", tau=0.99, k=100):
    m = F.mean()
    if loss == "sharp":            # SCORE.py:193-195
        return softplus_weighted(m, tau, k)
    if loss == "nll":              # SCORE.py:185-186
        return -torch.log(m)
    if loss == "infidelity":       # SCORE.py:189-190
        return 1 - m
    if loss == "none":
        return m
    raise ValueError(loss)


def train_step_loss_and_grad(pulses, U_target, error, M, loss="sharp", tau=0.99, k=100,
                             generator=generator_tree):
    """trainer.py:80-90 around an explicit pulses leaf: repeat_interleave(M) of
    pulses and targets, generator, loss on the pooled mean fidelity, backward.
    Returns (loss (), grad (B,L,2), F (B*M,))."""
    leaf = pulses.detach().clone().requires_grad_(True)
    p_mc = leaf.repeat_interleave(M, dim=0)
    t_mc = U_target.repeat_interleave(M, dim=0)
    U = generator(p_mc, error)
    F = fidelity(U, t_mc, 1)
    val = loss_value(F, loss, tau, k)
    val.backward()
    return val.detach(), leaf.grad.detach(), F.detach()


def forward_fidelity(pulses, U_target, error, M, generator=generator_tree):
    """trainer.py:113-120 (evaluate) / visualize/util.py:244-249 (grid sweep)."""
    with torch.no_grad():
        p_mc = pulses.repeat_interleave(M, dim=0)
        t_mc = U_target.repeat_interleave(M, dim=0)
        return fidelity(generator(p_mc, error), t_mc, 1)
