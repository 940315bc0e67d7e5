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


# ----------------------------------------------------------------------------- two-qubit SU(4) cross-oracle
# NOT in the reference (README.md:86,122 promise train/two_qubit/ only).  SURVEY.md §8a A9 prescribes the oracle:
# torch.linalg.matrix_exp + the SCORE.py:131-142 tree on the builder-defined Hamiltonian (include/uqoc.h)
#   H = 1/2 [cos phi1 XI + sin phi1 YI + cos phi2 IX + sin phi2 IY + delta1 ZI + delta2 IZ + J ZZ],
#   U_k = exp(-i H_k tau_k (1 + eps)),
# in complex128 with autograd gradients.  Parity UNPINNED by the reference; this is a second, independent formulation
# next to oracle/uqoc_oracle.py::su4_* (numpy eigendecomposition) -- tests/golden/su4_cross.npz is generated from it.
def su4_generator_tree(pulses: torch.Tensor, error: torch.Tensor, J: float = 1.0) -> torch.Tensor:
    """pulses (Bm, L, 3) [phi1, phi2, tau], error (3, Bm) [delta1; delta2; eps] -> (Bm, 4, 4), following the op
    sequence of SCORE.py:116-142 (element-wise Hamiltonian build, matrix_exp, log2(L) batched-matmul tree)."""
    if pulses.ndim != 3 or pulses.shape[-1] != 3:
        raise ValueError("'pulses' must have shape (B, L, 3)")
    Bm = pulses.shape[0]
    cdtype = torch.complex64 if pulses.dtype == torch.float32 else torch.complex128
    s = _paulis(cdtype, pulses.device)
    I2, X, Y, Z = s[0], s[1], s[2], s[3]
    XI, YI, ZI = torch.kron(X, I2), torch.kron(Y, I2), torch.kron(Z, I2)
    IX, IY, IZ = torch.kron(I2, X), torch.kron(I2, Y), torch.kron(I2, Z)
    ZZ = torch.kron(Z, Z)
    p1, p2, tau = pulses[..., 0], pulses[..., 1], pulses[..., 2]
    d1, d2, eps = error[0], error[1], error[2]
    e4 = lambda v: v[..., None, None]
    ham = e4(torch.cos(p1)) * XI + e4(torch.sin(p1)) * YI + e4(torch.cos(p2)) * IX + e4(torch.sin(p2)) * IY
    ham = ham + d1[:, None, None, None] * ZI + d2[:, None, None, None] * IZ + J * ZZ
    ham = 0.5 * ham
    steps = torch.linalg.matrix_exp(-1j * ham * e4(tau) * (1 + eps)[:, None, None, None])
    eye = torch.eye(4, dtype=cdtype, device=pulses.device).expand(Bm, 1, 4, 4)
    level = steps
    while level.size(1) > 1:
        if level.size(1) % 2 == 1:
            level = torch.cat([level, eye], dim=1)
        level = level[:, 1::2] @ level[:, 0::2]
    return level[:, 0]


def su4_train_step(pulses, U_target, error, M, J=1.0, loss="sharp", tau=0.99, k=100):
    """trainer.py:80-88 with the SU(4) generator (num_qubits = 2): returns (loss (), F (B*M,), U (B*M, 4, 4)); call
    ``loss.backward()`` for d loss / d pulses on the ``pulses`` leaf."""
    p_mc = pulses.repeat_interleave(M, dim=0)
    t_mc = U_target.repeat_interleave(M, dim=0)
    U = su4_generator_tree(p_mc, error, J)
    F = fidelity(U, t_mc, 2)
    return loss_value(F, loss, tau, k), F, U
