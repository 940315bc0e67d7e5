"""CPU oracle for the disorder-sampled propagation + fidelity-loss path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``universal_quantum_optimal_control_b200``)
never imports this module and has no CPU fallback.

This file is an independent numpy restatement (complex matrix form, float64 by
default) of the reference algorithm.  Each function cites the reference lines it
follows (paths relative to the upstream repository root):

* ``SCORE.py``  = train/unitary_single_qubit_gate/universal_single_qubit_SCORE.py
* ``grape.py``  = train/GRAPE/grape_train.py
* ``trainer.py``= model/universal_model_trainer.py

Parity pinning: the reference ships no test for this path, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (imports the unmodified reference)
and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this module against every one of them.

The SU(4) two-qubit section has NO counterpart in the reference (SURVEY.md §8a
row A9): its Hamiltonian is builder-defined and its parity is *unpinned* by the
reference; the oracle there is scipy/numpy matrix exponentials of that definition.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# Pauli algebra (SCORE.py:53-70)
# --------------------------------------------------------------------------
I2 = np.eye(2, dtype=np.complex128)
SX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
SY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
SZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)


def _ctype(real_dtype):
    return np.complex64 if np.dtype(real_dtype) == np.float32 else np.complex128


# --------------------------------------------------------------------------
# A1 + A2: Hamiltonian and per-pulse exponential
# --------------------------------------------------------------------------
def pulse_hamiltonians(pulses: np.ndarray, error: np.ndarray) -> np.ndarray:
    """H[s,i] = 1/2 (1+eps_s) (cos(phi) X + sin(phi) Y + delta_s Z).

    SCORE.py:107-124 (twin grape.py:108-125).  Note that the detuning term is
    scaled by (1+eps) as well -- the code, not the README prose, is the contract.
    ``pulses`` (Bm, L, 2) = [phi, tau]; ``error`` (2, Bm) = [delta; eps].
    """
    pulses = np.asarray(pulses)
    error = np.asarray(error)
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        # SCORE.py:99-100
        raise ValueError("'pulses' must have shape (B, L, 2)")
    phi = pulses[..., 0]
    delta = error[0][:, None]
    eps = error[1][:, None]
    ct = _ctype(pulses.dtype)
    H = (np.cos(phi)[..., None, None] * SX.astype(ct)
         + np.sin(phi)[..., None, None] * SY.astype(ct)
         + delta[..., None, None] * SZ.astype(ct))
    return (0.5 * (1.0 + eps))[..., None, None].astype(pulses.dtype) * H


def pulse_unitaries(pulses: np.ndarray, error: np.ndarray) -> np.ndarray:
    """U[s,i] = exp(-i H[s,i] tau_i)  (SCORE.py:127, grape.py:128).

    The reference calls ``torch.linalg.matrix_exp``; for a traceless Hermitian
    2x2 generator G = H*tau with G^2 = g^2 I the exponential is exactly
    cos(g) I - i sin(g)/g G, which is what is evaluated here.
    """
    pulses = np.asarray(pulses)
    H = pulse_hamiltonians(pulses, error)
    tau = pulses[..., 1]
    G = H * tau[..., None, None]
    # g^2 = -det(G) for traceless Hermitian G
    g2 = (G[..., 0, 0] * G[..., 0, 0] + G[..., 0, 1] * G[..., 1, 0]).real
    g = np.sqrt(np.maximum(g2, 0.0))
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(g > 0, np.sin(g) / np.where(g > 0, g, 1.0), 1.0)
    ct = G.dtype
    eye = np.eye(2, dtype=ct)
    return (np.cos(g)[..., None, None] * eye - 1j * sinc[..., None, None] * G).astype(ct)


# --------------------------------------------------------------------------
# A3 / A4: ordered product
# --------------------------------------------------------------------------
def ordered_product_tree(U: np.ndarray) -> np.ndarray:
    """U_L ... U_1 by the pairwise tree of SCORE.py:131-142 (identity padded)."""
    X = U
    Bm = X.shape[0]
    eye = np.broadcast_to(np.eye(2, dtype=X.dtype), (Bm, 1, 2, 2))
    while X.shape[1] > 1:
        if X.shape[1] & 1:
            X = np.concatenate([X, eye], axis=1)
        X = np.matmul(X[:, 1::2], X[:, 0::2])
    return X[:, 0]


def ordered_product_sequential(U: np.ndarray) -> np.ndarray:
    """U_L ... U_1 by the running product of grape.py:133-136."""
    Bm, L = U.shape[:2]
    out = np.broadcast_to(np.eye(2, dtype=U.dtype), (Bm, 2, 2)).copy()
    for k in range(L):
        out = np.matmul(U[:, k], out)
    return out


def batched_unitary_generator(pulses, error, product: str = "tree") -> np.ndarray:
    """Full generator: SCORE.py:77-145 (``product='tree'``) or grape.py:78-138
    (``product='sequential'``).  Returns (Bm, 2, 2) complex."""
    U = pulse_unitaries(pulses, error)
    if U.shape[1] == 0:
        return np.broadcast_to(np.eye(2, dtype=U.dtype), (U.shape[0], 2, 2)).copy()
    return ordered_product_tree(U) if product == "tree" else ordered_product_sequential(U)


# --------------------------------------------------------------------------
# A5-A7: fidelity and losses
# --------------------------------------------------------------------------
def fidelity(U_out: np.ndarray, U_target: np.ndarray, num_qubits: int = 1) -> np.ndarray:
    """F = (|Tr(U_out^dagger U_target)|^2 + d) / (d (d+1))   (SCORE.py:168-183)."""
    tr = np.einsum("bji,bji->b", np.conj(U_out), U_target)
    d = 2 ** num_qubits
    return (np.abs(tr) ** 2 + d) / (d * (d + 1))


def custom_loss(x, tau=0.99, k=100):
    """log(1+exp(-k (x - tau))) * (1 - x)   (SCORE.py:197-198)."""
    return np.log(1.0 + np.exp(-k * (x - tau))) * (1.0 - x)


def custom_loss_grad(x, tau=0.99, k=100):
    """d custom_loss / d x."""
    z = np.exp(-k * (x - tau))
    return -k * z / (1.0 + z) * (1.0 - x) - np.log(1.0 + z)


def sharp_loss_from_F(F, tau=0.99, k=100):
    """SCORE.py:193-195 applied to a fidelity vector."""
    return custom_loss(np.mean(F), tau, k)


def negative_log_loss_from_F(F):
    """SCORE.py:185-186."""
    return -np.log(np.mean(F))


def infidelity_loss_from_F(F):
    """SCORE.py:189-190."""
    return 1.0 - np.mean(F)


LOSSES = ("sharp", "nll", "infidelity", "none")


def loss_and_dloss(Fbar, loss: str = "sharp", tau=0.99, k=100):
    """Scalar loss value and d loss / d Fbar for each loss the trainer accepts."""
    if loss == "sharp":
        return custom_loss(Fbar, tau, k), custom_loss_grad(Fbar, tau, k)
    if loss == "nll":
        return -np.log(Fbar), -1.0 / Fbar
    if loss == "infidelity":
        return 1.0 - Fbar, -1.0
    if loss == "none":  # plain mean fidelity
        return Fbar, 1.0
    raise ValueError(f"unknown loss {loss!r}")


# --------------------------------------------------------------------------
# A8: error samplers (distributional restatement; torch's RNG stream itself is
# not reproducible from numpy, golden eps come from the fixtures)
# --------------------------------------------------------------------------
def get_ore_ple_error_distribution(batch_size, delta_std=1.0, epsilon_std=0.05, rng=None):
    """SCORE.py:158-161: stack([N(0,1)*delta_std, N(0,1)*epsilon_std]) -> (2, n) f32."""
    rng = rng or np.random.default_rng()
    ore = rng.standard_normal(batch_size).astype(np.float32) * np.float32(delta_std)
    ple = rng.standard_normal(batch_size).astype(np.float32) * np.float32(epsilon_std)
    return np.stack([ore, ple])


def get_ore_error_distribution(batch_size, delta_std=1.0, rng=None):
    """SCORE.py:154-155."""
    rng = rng or np.random.default_rng()
    return rng.standard_normal(batch_size).astype(np.float32) * np.float32(delta_std)


# --------------------------------------------------------------------------
# A10: the trainer's Monte-Carlo layout + analytic backward
# --------------------------------------------------------------------------
def expand_mc(pulses, U_target, M):
    """trainer.py:80-81: repeat_interleave(M, dim=0) of pulses and targets, so
    sample index s = b*M + j."""
    return np.repeat(pulses, M, axis=0), np.repeat(U_target, M, axis=0)


def per_sample_fidelity(pulses, U_target, error, M, product="tree"):
    """F for every (target b, sample j): trainer.py:80-88 without the loss."""
    p_mc, t_mc = expand_mc(np.asarray(pulses), np.asarray(U_target), M)
    U = batched_unitary_generator(p_mc, error, product)
    return fidelity(U, t_mc, 1), U


def fidelity_sum_and_grad(pulses, U_target, error, M):
    """Sum_j F[b,j] per target and d(Sum_j F[b,j])/d pulses[b]  -- the linear
    part of trainer.py:80-90's backward, by an explicit prefix/suffix adjoint in
    complex-matrix form (independent of the kernels' quaternion formulation).

    Returns (Fsum (B,), grad (B, L, 2), F (B*M,)).
    """
    pulses = np.asarray(pulses, dtype=np.float64)
    error = np.asarray(error, dtype=np.float64)
    T = np.asarray(U_target, dtype=np.complex128)
    B, L, _ = pulses.shape
    p_mc, t_mc = expand_mc(pulses, T, M)
    Bm = B * M
    U = pulse_unitaries(p_mc, error)                     # (Bm, L, 2, 2)
    H = pulse_hamiltonians(p_mc, error)
    # prefix[i] = U_i ... U_1 (prefix[0] = I), suffix[i] = U_L ... U_{i+1}
    prefix = np.empty((Bm, L + 1, 2, 2), dtype=np.complex128)
    prefix[:, 0] = I2
    for i in range(L):
        prefix[:, i + 1] = U[:, i] @ prefix[:, i]
    suffix = np.empty((Bm, L + 1, 2, 2), dtype=np.complex128)
    suffix[:, L] = I2
    for i in range(L - 1, -1, -1):
        suffix[:, i] = suffix[:, i + 1] @ U[:, i]
    U_out = prefix[:, L]
    tr = np.einsum("bji,bji->b", np.conj(U_out), t_mc)
    F = (np.abs(tr) ** 2 + 2.0) / 6.0
    # dF = (1/6) * 2 Re( tr * Tr(T^dagger dU_out) ),  dU_out = suffix[i+1] dU_i prefix[i]
    Td = np.conj(np.swapaxes(t_mc, -1, -2))
    phi = p_mc[..., 0]
    delta = error[0][:, None]
    eps = error[1][:, None]
    w = np.sqrt(1.0 + delta ** 2)
    h = p_mc[..., 1] * 0.5 * (1.0 + eps) * w
    # dU_i/dtau = -i H_i U_i ;  dU_i/dphi = -i sin(h)/w (-sin phi X + cos phi Y)
    dU_dtau = -1j * (H @ U)
    dn = (-np.sin(phi))[..., None, None] * SX + np.cos(phi)[..., None, None] * SY
    dU_dphi = -1j * (np.sin(h) / w)[..., None, None] * dn
    # A_i = prefix[i] T^dagger suffix[i+1]  so that Tr(T^dagger dU_out) = Tr(A_i dU_i)
    A = prefix[:, :L] @ Td[:, None] @ suffix[:, 1:]
    g_tau = (2.0 / 6.0) * (tr[:, None] * np.einsum("blij,blji->bl", A, dU_dtau)).real
    g_phi = (2.0 / 6.0) * (tr[:, None] * np.einsum("blij,blji->bl", A, dU_dphi)).real
    grad = np.stack([g_phi, g_tau], axis=-1).reshape(B, M, L, 2).sum(axis=1)
    Fsum = F.reshape(B, M).sum(axis=1)
    return Fsum, grad, F


def loss_and_grad(pulses, U_target, error, M, loss="sharp", tau=0.99, k=100):
    """Scalar loss and d loss / d pulses (B, L, 2) exactly as trainer.py:80-90
    produces them: the loss acts on the mean fidelity pooled over all B*M
    samples (SCORE.py:194)."""
    Fsum, grad, F = fidelity_sum_and_grad(pulses, U_target, error, M)
    n = F.size
    Fbar = Fsum.sum() / n
    val, dval = loss_and_dloss(Fbar, loss, tau, k)
    return val, dval / n * grad, F


# --------------------------------------------------------------------------
# Counter-based RNG: Philox4x32-10 (Random123) + Box-Muller, the on-chip
# replacement of SCORE.py:158-161 (north_star: cuRAND-free Philox).
# --------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10.  ``ctr`` (..., 4) uint32, ``key`` (..., 2) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c[0].astype(np.uint64)
            p1 = PHILOX_M1 * c[2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_errors(B, M, sigma_delta, sigma_eps, seed, offset, j0=0, dtype=np.float64):
    """(delta, eps) for targets b < B and global sample indices j0 <= j < j0+M.

    Counter layout (shared with csrc/uqoc_philox.cuh): ctr = (j lo32, j hi32, b,
    offset lo32), key = (seed lo32, seed hi32).  u_k = (x_k + 0.5) * 2^-32,
    radius = sqrt(-2 ln u_0), angle = 2 pi u_1; delta = sigma_delta * radius *
    cos(angle), eps = sigma_eps * radius * sin(angle).  Returns (2, B*M) with
    sample index s = b*M + (j - j0).
    """
    j = (np.arange(M, dtype=np.uint64) + np.uint64(j0))[None, :].repeat(B, axis=0)
    b = np.arange(B, dtype=np.uint32)[:, None].repeat(M, axis=1)
    ctr = np.stack([(j & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    (j >> np.uint64(32)).astype(np.uint32),
                    b,
                    np.full((B, M), np.uint32(offset & 0xFFFFFFFF), dtype=np.uint32)], axis=-1)
    key = np.empty((B, M, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key)
    u0 = (x[..., 0].astype(np.float64) + 0.5) * 2.0 ** -32
    u1 = (x[..., 1].astype(np.float64) + 0.5) * 2.0 ** -32
    rad = np.sqrt(-2.0 * np.log(u0))
    ang = 2.0 * np.pi * u1
    delta = sigma_delta * rad * np.cos(ang)
    eps = sigma_eps * rad * np.sin(ang)
    return np.stack([delta.reshape(-1), eps.reshape(-1)]).astype(dtype)


def philox_errors_su4(B, M, sigma_delta, sigma_eps, seed, offset, j0=0, dtype=np.float64):
    """(delta1, delta2, eps) of the SU(4) kernels: first Box-Muller pair -> (delta1, eps) (the SU(2)
    stream), second pair's cosine branch -> delta2.  Returns (3, B*M)."""
    j = (np.arange(M, dtype=np.uint64) + np.uint64(j0))[None, :].repeat(B, axis=0)
    b = np.arange(B, dtype=np.uint32)[:, None].repeat(M, axis=1)
    ctr = np.stack([(j & np.uint64(0xFFFFFFFF)).astype(np.uint32), (j >> np.uint64(32)).astype(np.uint32), b,
                    np.full((B, M), np.uint32(offset & 0xFFFFFFFF), dtype=np.uint32)], axis=-1)
    key = np.empty((B, M, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key)
    u = (x.astype(np.float64) + 0.5) * 2.0 ** -32
    r0, r1 = np.sqrt(-2.0 * np.log(u[..., 0])), np.sqrt(-2.0 * np.log(u[..., 2]))
    d1 = sigma_delta * r0 * np.cos(2 * np.pi * u[..., 1])
    eps = sigma_eps * r0 * np.sin(2 * np.pi * u[..., 1])
    d2 = sigma_delta * r1 * np.cos(2 * np.pi * u[..., 3])
    return np.stack([d1.reshape(-1), d2.reshape(-1), eps.reshape(-1)]).astype(dtype)


# --------------------------------------------------------------------------
# A9: two-qubit SU(4) path.  NOT IN THE REFERENCE -- builder-defined (SURVEY.md
# §8a row A9), parity unpinned by the reference.  Same callable contract:
#   pulses (Bm, L, 3) = [phi1, phi2, tau], error (3, Bm) = [delta1, delta2, eps]
#   H = 1/2 [cos phi1 XI + sin phi1 YI + cos phi2 IX + sin phi2 IY
#            + delta1 ZI + delta2 IZ + J ZZ],   U_k = exp(-i H_k tau_k (1+eps)).
# --------------------------------------------------------------------------
def _kron(a, b):
    return np.kron(a, b)


XI, YI, ZI = _kron(SX, I2), _kron(SY, I2), _kron(SZ, I2)
IX, IY, IZ = _kron(I2, SX), _kron(I2, SY), _kron(I2, SZ)
ZZ = _kron(SZ, SZ)


def su4_hamiltonians(pulses, error, J=1.0):
    pulses = np.asarray(pulses, dtype=np.float64)
    error = np.asarray(error, dtype=np.float64)
    if pulses.ndim != 3 or pulses.shape[-1] != 3:
        raise ValueError("'pulses' must have shape (B, L, 3)")
    p1, p2 = pulses[..., 0], pulses[..., 1]
    d1 = error[0][:, None]
    d2 = error[1][:, None]
    e = lambda a: a[..., None, None]
    H = (e(np.cos(p1)) * XI + e(np.sin(p1)) * YI + e(np.cos(p2)) * IX + e(np.sin(p2)) * IY
         + e(d1 + 0 * p1) * ZI + e(d2 + 0 * p1) * IZ + J * ZZ)
    return 0.5 * H


def _expm_herm(G):
    """exp(-i G) for Hermitian G (..., n, n) through an eigendecomposition."""
    w, V = np.linalg.eigh(G)
    return (V * np.exp(-1j * w)[..., None, :]) @ np.conj(np.swapaxes(V, -1, -2))


def su4_pulse_unitaries(pulses, error, J=1.0):
    pulses = np.asarray(pulses, dtype=np.float64)
    error = np.asarray(error, dtype=np.float64)
    H = su4_hamiltonians(pulses, error, J)
    scale = pulses[..., 2] * (1.0 + error[2][:, None])
    return _expm_herm(H * scale[..., None, None])


def su4_unitary_generator(pulses, error, J=1.0):
    U = su4_pulse_unitaries(pulses, error, J)
    Bm, L = U.shape[:2]
    out = np.broadcast_to(np.eye(4, dtype=np.complex128), (Bm, 4, 4)).copy()
    for k in range(L):
        out = U[:, k] @ out
    return out


def su4_loss_and_grad(pulses, U_target, error, M, J=1.0, loss="sharp", tau=0.99, k=100):
    """Pooled-mean loss (SCORE.py:194 semantics with d = 4) and d loss / d pulses (B, L, 3)."""
    Fsum, grad, F = su4_fidelity_sum_and_grad(pulses, U_target, error, M, J)
    n = F.size
    val, dval = loss_and_dloss(Fsum.sum() / n, loss, tau, k)
    return val, dval / n * grad, F


def su4_fidelity_sum_and_grad(pulses, U_target, error, M, J=1.0, fd_step=None):
    """Sum_j F and its pulse gradient for the SU(4) path.  Gradient by an exact
    adjoint: dU_i = -i * int_0^1 exp(-i s G) dG exp(-i (1-s) G) ds evaluated in
    the eigenbasis of G (divided differences of exp)."""
    pulses = np.asarray(pulses, dtype=np.float64)
    error = np.asarray(error, dtype=np.float64)
    T = np.asarray(U_target, dtype=np.complex128)
    B, L, _ = pulses.shape
    p_mc = np.repeat(pulses, M, axis=0)
    t_mc = np.repeat(T, M, axis=0)
    Bm = B * M
    H = su4_hamiltonians(p_mc, error, J)
    scale = p_mc[..., 2] * (1.0 + error[2][:, None])
    G = H * scale[..., None, None]
    w, V = np.linalg.eigh(G)
    Vh = np.conj(np.swapaxes(V, -1, -2))
    ew = np.exp(-1j * w)
    U = (V * ew[..., None, :]) @ Vh
    prefix = np.empty((Bm, L + 1, 4, 4), dtype=np.complex128)
    prefix[:, 0] = np.eye(4)
    for i in range(L):
        prefix[:, i + 1] = U[:, i] @ prefix[:, i]
    suffix = np.empty((Bm, L + 1, 4, 4), dtype=np.complex128)
    suffix[:, L] = np.eye(4)
    for i in range(L - 1, -1, -1):
        suffix[:, i] = suffix[:, i + 1] @ U[:, i]
    U_out = prefix[:, L]
    tr = np.einsum("bji,bji->b", np.conj(U_out), t_mc)
    F = (np.abs(tr) ** 2 + 4.0) / 20.0
    Td = np.conj(np.swapaxes(t_mc, -1, -2))
    A = prefix[:, :L] @ Td[:, None] @ suffix[:, 1:]          # Tr(T^dag dU_out) = Tr(A_i dU_i)
    # divided differences  Phi[a,b] = (e^{-i w_a} - e^{-i w_b}) / (w_a - w_b)   (-> -i e^{-i w} on the diagonal)
    dw = w[..., :, None] - w[..., None, :]
    de = ew[..., :, None] - ew[..., None, :]
    small = np.abs(dw) < 1e-9
    Phi = np.where(small, -1j * ew[..., :, None], de / np.where(small, 1.0, dw))

    def dU_of(dG):
        inner = Vh @ dG @ V
        return V @ (Phi * inner) @ Vh

    e = lambda a: a[..., None, None]
    p1, p2 = p_mc[..., 0], p_mc[..., 1]
    dG_p1 = 0.5 * e(scale) * (e(-np.sin(p1)) * XI + e(np.cos(p1)) * YI)
    dG_p2 = 0.5 * e(scale) * (e(-np.sin(p2)) * IX + e(np.cos(p2)) * IY)
    dG_tau = H * e((1.0 + error[2][:, None]) + 0 * p1)
    grads = []
    for dG in (dG_p1, dG_p2, dG_tau):
        dU = dU_of(dG)
        g = (2.0 / 20.0) * (tr[:, None] * np.einsum("blij,blji->bl", A, dU)).real
        grads.append(g)
    grad = np.stack(grads, axis=-1).reshape(B, M, L, 3).sum(axis=1)
    return F.reshape(B, M).sum(axis=1), grad, F
