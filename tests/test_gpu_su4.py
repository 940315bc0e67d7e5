"""GPU tests of the two-qubit SU(4) path against oracle/uqoc_oracle.py::su4_*.  The reference has
no two-qubit code: parity here is against the builder's own oracle (unpinned)."""
import numpy as np
import pytest
import torch

import universal_quantum_optimal_control_b200 as uq
from oracle import uqoc_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _case(seed, B, L, M, J=1.0, sd=1.0):
    rng = np.random.default_rng(seed)
    pulses = np.stack([rng.uniform(-3.15, 3.15, (B, L)), rng.uniform(-3.15, 3.15, (B, L)), rng.uniform(0.1, 0.5, (B, L))], -1)
    err = np.stack([rng.normal(0, sd, B * M), rng.normal(0, sd, B * M), rng.normal(0, 0.05, B * M)])
    T = orc.su4_unitary_generator(np.stack([rng.uniform(-3, 3, (B, 5)), rng.uniform(-3, 3, (B, 5)), rng.uniform(0.3, 1.0, (B, 5))], -1),
                                  np.zeros((3, B)), J)
    return pulses, err, T


PADE = [False, True]      # default eigenframe kernel / per-pulse scaling-and-squaring kernel (UQOC_FLAG_SU4_PADE)


@pytest.mark.parametrize("pade,wps", [(False, 0), (False, 1), (True, 0)])     # wps 0: library plan (train split over the warps
@pytest.mark.parametrize("dtype,tolF,tolG", [(torch.float64, 1e-12, 1e-10), (torch.float32, 1e-5, 1e-4)])       # for L >= 32 here), 1: never
@pytest.mark.parametrize("L,M,splits", [(1, 5, 0), (7, 70, 0), (33, 130, 3), (34, 40, 0), (128, 64, 0), (400, 96, 0)])
def test_su4_fused_matches_oracle(dtype, tolF, tolG, L, M, splits, pade, wps):
    B, J = 3, 0.8
    pulses, err, T = _case(L, B, L, M, J)
    if dtype == torch.float32:
        pulses, err = pulses.astype(np.float32).astype(np.float64), err.astype(np.float32).astype(np.float64)
    want_l, want_g, want_F = orc.su4_loss_and_grad(pulses, T, err, M, J, "sharp")
    p = _t(pulses, dtype).requires_grad_(True)
    F = torch.empty(B * M, dtype=dtype, device=DEV)
    loss, mf = uq.fused_propagate_loss_su4(p, _t(T), error=_t(err, dtype), monte_carlo=M, J=J, F_out=F,
                                           flags=uq.tuning_flags(splits=splits, su4_pade=pade, wps=wps))
    loss.backward()
    assert np.abs(F.cpu().numpy() - want_F).max() < tolF
    assert abs(loss.item() - want_l) < (1e-11 if dtype == torch.float64 else 1e-4) * max(1, abs(want_l))
    g = p.grad.cpu().numpy().astype(np.float64)
    assert np.abs(g - want_g).max() / np.abs(want_g).max() < tolG
    assert np.abs(mf.cpu().numpy() - want_F.reshape(B, M).mean(1)).max() < 10 * tolF


@pytest.mark.parametrize("pade", PADE)
def test_su4_general_target_and_large_detuning(pade):
    B, L, M, J = 2, 20, 40, 1.3
    pulses, err, T = _case(3, B, L, M, J, sd=2.5)            # |delta| up to ~7: several squarings
    rng = np.random.default_rng(9)
    T = T + 0.3 * (rng.normal(size=T.shape) + 1j * rng.normal(size=T.shape))
    want_l, want_g, want_F = orc.su4_loss_and_grad(pulses, T, err, M, J, "nll")
    p = _t(pulses).requires_grad_(True)
    F = torch.empty(B * M, dtype=torch.float64, device=DEV)
    loss, _ = uq.fused_propagate_loss_su4(p, _t(T), error=_t(err), monte_carlo=M, J=J, loss="nll", F_out=F,
                                          flags=uq.tuning_flags(su4_pade=pade))
    loss.backward()
    assert np.abs(F.cpu().numpy() - want_F).max() < 1e-11
    assert abs(loss.item() - want_l) < 1e-11
    assert np.abs(p.grad.cpu().numpy() - want_g).max() / np.abs(want_g).max() < 1e-10


@pytest.mark.parametrize("pade", PADE)
def test_su4_generator_and_generic_fidelity(pade):
    fl = uq.tuning_flags(su4_pade=pade)
    B, L, M, J = 1, 16, 50, 1.0
    pulses, err, T = _case(4, B, L, M, J)
    want_U = orc.su4_unitary_generator(np.repeat(pulses, M, 0), err, J)
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 3e-6)):
        pl = _t(pulses, dtype)
        U = uq.su4_unitary_generator(pl.expand(M, -1, -1), _t(err, dtype), J, fl)      # shared pulse train
        assert U.shape == (M, 4, 4)
        assert np.abs(U.cpu().numpy() - want_U).max() < tol
        U2 = uq.su4_unitary_generator(pl.expand(M, -1, -1).contiguous(), _t(err, dtype), J, fl)   # per-sample rows
        assert np.abs(U2.cpu().numpy() - want_U).max() < tol
        F = uq.fidelity(U, _t(T).expand(M, -1, -1), 2)                                  # d = 4 (SCORE.py:181-183)
        want_F = orc.fidelity(want_U, np.repeat(T, M, 0), 2)
        assert np.abs(F.cpu().numpy() - want_F).max() < 10 * tol
    # J = 0 factorises into two independent SU(2) propagators
    U0 = uq.su4_unitary_generator(_t(pulses).expand(M, -1, -1), _t(err), 0.0, fl).cpu().numpy()
    p1 = np.repeat(pulses[:, :, [0, 2]], M, 0)
    p2 = np.repeat(pulses[:, :, [1, 2]], M, 0)
    # SU(2) oracle scales delta by (1+eps) like the SU(4) definition does
    Ua = orc.batched_unitary_generator(p1, err[[0, 2]])
    Ub = orc.batched_unitary_generator(p2, err[[1, 2]])
    kron = np.einsum("bij,bkl->bikjl", Ua, Ub).reshape(M, 4, 4)
    assert np.abs(U0 - kron).max() < 1e-12


def test_su4_philox_stream_and_sharding():
    B, M = 3, 500
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 3e-6)):
        e = uq.philox_errors_su4(B, M, (0.7, 0.05), seed=77, offset=4, j0=9, dtype=dtype).cpu().numpy()
        want = orc.philox_errors_su4(B, M, 0.7, 0.05, 77, 4, j0=9)
        assert np.abs(e - want).max() < tol
    # (delta1, eps) coincide with the SU(2) stream of the same (seed, offset, b, j)
    e2 = uq.philox_errors(B, M, (0.7, 0.05), seed=77, offset=4, j0=9, dtype=torch.float64).cpu().numpy()
    assert np.abs(e[[0, 2]] - e2).max() < 3e-6
    pulses, _, T = _case(5, B, 12, M)
    p = _t(pulses).requires_grad_(True)
    err_out = torch.empty(3, B * M, dtype=torch.float64, device=DEV)
    loss, _ = uq.fused_propagate_loss_su4(p, _t(T), monte_carlo=M, sigma=(0.7, 0.05), seed=77, offset=4, err_out=err_out)
    loss.backward()
    want_l, want_g, _ = orc.su4_loss_and_grad(pulses, T, err_out.cpu().numpy(), M, 1.0)
    assert abs(loss.item() - want_l) < 1e-11
    assert np.abs(p.grad.cpu().numpy() - want_g).max() / np.abs(want_g).max() < 1e-10


def test_su4_degenerate_spectrum():
    """delta1 = delta2 = 0 and J = 0 make 2H' = XI + IX doubly degenerate: the Jacobi basis is then arbitrary
    inside the degenerate subspace, the propagator must not be."""
    B, L, M = 2, 9, 6
    pulses, err, T = _case(11, B, L, M, 0.0)
    err[:2] = 0.0
    err[0, 3] = 1e-9                                          # nearly degenerate
    want_l, want_g, want_F = orc.su4_loss_and_grad(pulses, T, err, M, 0.0, "sharp")
    p = _t(pulses).requires_grad_(True)
    F = torch.empty(B * M, dtype=torch.float64, device=DEV)
    loss, _ = uq.fused_propagate_loss_su4(p, _t(T), error=_t(err), monte_carlo=M, J=0.0, F_out=F)
    loss.backward()
    assert np.abs(F.cpu().numpy() - want_F).max() < 1e-12
    assert np.abs(p.grad.cpu().numpy() - want_g).max() / np.abs(want_g).max() < 1e-10


def test_su4_kernels_agree_fp32_long_train():
    """FP32 eigenframe kernel against the FP64 kernel at L = 512 (coherent rounding of V / mu would show here)."""
    torch.manual_seed(1)
    B, L, M, J = 2, 512, 2048, 1.0
    pulses = torch.stack([(torch.rand(B, L) * 2 - 1) * 3.15, (torch.rand(B, L) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L)], -1).to(DEV)
    T = torch.diag(torch.tensor([1, 1, 1, -1], dtype=torch.complex64)).to(DEV)[None].expand(B, -1, -1)
    err = uq.philox_errors_su4(B, M, (1.0, 0.05), seed=5)
    F32 = torch.empty(B * M, device=DEV)
    F64 = torch.empty(B * M, device=DEV, dtype=torch.float64)
    p32 = pulses.clone().requires_grad_(True)
    l32, _ = uq.fused_propagate_loss_su4(p32, T, error=err, monte_carlo=M, J=J, F_out=F32)
    l32.backward()
    p64 = pulses.double().requires_grad_(True)
    l64, _ = uq.fused_propagate_loss_su4(p64, T, error=err.double(), monte_carlo=M, J=J, F_out=F64)
    l64.backward()
    assert (F32.double() - F64).abs().max().item() < 1e-5
    assert ((p32.grad.double() - p64.grad).abs().max() / p64.grad.abs().max()).item() < 1e-4


def test_su4_config4_sized_properties():
    """BASELINE config 4 shape (L=128, 32k eps): invariants + a sampled oracle check."""
    torch.manual_seed(0)
    B, L, M, J = 1, 128, 32768, 1.0
    pulses = torch.stack([(torch.rand(B, L) * 2 - 1) * 3.15, (torch.rand(B, L) * 2 - 1) * 3.15, 0.1 + 0.4 * torch.rand(B, L)], -1).to(DEV)
    CZ = torch.diag(torch.tensor([1, 1, 1, -1], dtype=torch.complex64)).to(DEV)[None]
    p = pulses.clone().requires_grad_(True)
    F = torch.empty(B * M, device=DEV)
    err_out = torch.empty(3, B * M, device=DEV)
    loss, mf = uq.fused_propagate_loss_su4(p, CZ, monte_carlo=M, J=J, seed=3, F_out=F, err_out=err_out)
    loss.backward()
    assert F.min().item() >= 0.2 - 1e-5 and F.max().item() <= 1 + 1e-5
    g1 = p.grad.clone()
    p.grad = None
    loss2, _ = uq.fused_propagate_loss_su4(p, CZ, monte_carlo=M, J=J, seed=3)
    loss2.backward()
    assert torch.equal(g1, p.grad)                                   # deterministic
    idx = np.arange(0, M, 512)
    U = orc.su4_unitary_generator(np.repeat(pulses.cpu().numpy().astype(np.float64), len(idx), 0),
                                  err_out.cpu().numpy()[:, idx].astype(np.float64), J)
    want = orc.fidelity(U, np.repeat(CZ.cpu().numpy(), len(idx), 0), 2)
    assert np.abs(F.cpu().numpy()[idx] - want).max() < 1e-5
    p64 = pulses.double().requires_grad_(True)
    l64, _ = uq.fused_propagate_loss_su4(p64, CZ, error=err_out.double(), monte_carlo=M, J=J)
    l64.backward()
    assert ((g1.double() - p64.grad).abs().max() / p64.grad.abs().max()).item() < 1e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-4)])
def test_su4_generator_backward_matches_torch_autograd(dtype, tol):
    """su4_unitary_generator is differentiable: an arbitrary real functional of U_out back-propagated through the
    cotangent-seeded eigenframe kernel equals autograd through the matrix_exp + tree cross-oracle (complex128, CPU)."""
    from oracle import torch_port as tp
    Bm, L, J = 37, 19, 0.9
    rng = np.random.default_rng(11)
    pulses = np.stack([rng.uniform(-3.15, 3.15, (Bm, L)), rng.uniform(-3.15, 3.15, (Bm, L)), rng.uniform(0.1, 0.5, (Bm, L))], -1)
    err = np.stack([rng.normal(0, 1.0, Bm), rng.normal(0, 1.0, Bm), rng.normal(0, 0.05, Bm)])
    if dtype == torch.float32:
        pulses, err = pulses.astype(np.float32).astype(np.float64), err.astype(np.float32).astype(np.float64)
    Cw = rng.normal(size=(Bm, 4, 4)) + 1j * rng.normal(size=(Bm, 4, 4))
    pr = torch.from_numpy(pulses).requires_grad_(True)
    Ur = tp.su4_generator_tree(pr, torch.from_numpy(err), J)
    lr = (torch.from_numpy(Cw).conj() * Ur).real.sum() + (Ur.abs() ** 4).sum()          # linear + non-linear part
    lr.backward()
    p = _t(pulses, dtype).requires_grad_(True)
    U = uq.su4_unitary_generator(p, _t(err, dtype), J)
    cdt = torch.complex128 if dtype == torch.float64 else torch.complex64
    l = (_t(Cw).to(cdt).conj() * U).real.sum() + (U.abs() ** 4).sum()
    l.backward()
    assert (U.detach().cpu().to(torch.complex128) - Ur.detach()).abs().max().item() < (1e-12 if dtype == torch.float64 else 2e-5)
    g, gr = p.grad.cpu().double(), pr.grad
    assert ((g - gr).abs().max() / gr.abs().max()).item() < tol


def test_su4_forward_only_fits_the_shape_workspace():
    """the forward-only launch of a shape whose forward kernel would pick more splits than the backward kernel
    used to fail with UQOC_E_WORKSPACE; the split count is now the same for both."""
    L, J = 16, 1.0
    for B in (100, 148 * 3, 148 * 5, 148 * 6 + 1, 1500):
        M = 200
        pulses, err, T = _case(B, B, L, M, J)
        p = _t(pulses, torch.float32)
        with torch.no_grad():
            loss, mf = uq.fused_propagate_loss_su4(p, _t(T), monte_carlo=M, J=J, sigma=(0.5, 0.05), seed=3)
        pg = p.clone().requires_grad_(True)
        loss2, mf2 = uq.fused_propagate_loss_su4(pg, _t(T), monte_carlo=M, J=J, sigma=(0.5, 0.05), seed=3)
        assert torch.allclose(mf, mf2, atol=2e-6), B
        assert abs(loss.item() - loss2.item()) < 1e-5 * max(1.0, abs(loss2.item()))


def test_su4_device_resident_philox_state():
    """UQOC_FLAG_RNG_FROM_DEVICE for the SU(4) kernels: (seed, offset) read from device memory give the same step as
    the immediate values (CUDA-graph replays advance the offset on the device)."""
    B, L, M, J = 2, 12, 300, 1.0
    pulses, err, T = _case(5, B, L, M, J)
    for pade in PADE:
        fl = uq.tuning_flags(su4_pade=pade)
        p1 = _t(pulses, torch.float32).requires_grad_(True)
        l1, f1 = uq.fused_propagate_loss_su4(p1, _t(T), monte_carlo=M, J=J, sigma=(0.7, 0.05), seed=17, offset=5, flags=fl)
        l1.backward()
        rng = torch.tensor([17, 5], dtype=torch.int64, device=DEV)
        p2 = _t(pulses, torch.float32).requires_grad_(True)
        l2, f2 = uq.fused_propagate_loss_su4(p2, _t(T), monte_carlo=M, J=J, sigma=(0.7, 0.05), seed=rng.data_ptr(), flags=fl | 8)
        l2.backward()
        assert torch.equal(f1, f2) and torch.equal(p1.grad, p2.grad) and l1.item() == l2.item()


def test_su4_graphed_train_step_matches_eager():
    """GraphedTrainStep with a 3-parameter pulse model (two-qubit path): replays reproduce the eager FusedTrainer-style
    steps (same Philox offsets, Adam, clipping)."""
    from universal_quantum_optimal_control_b200.trainer import GraphedTrainStep, SigmaSpec
    B, L, M = 4, 10, 256

    class Tiny2Q(torch.nn.Module):
        num_qubits = 2

        def __init__(self):
            super().__init__()
            self.net = torch.nn.Linear(6, 3 * L)

        def forward(self, x):
            y = self.net(x).view(-1, L, 3)
            return torch.stack([3.0 * torch.tanh(y[..., 0]), 3.0 * torch.tanh(y[..., 1]), 0.1 + 0.4 * torch.sigmoid(y[..., 2])], -1)

    torch.manual_seed(0)
    m_g, m_e = Tiny2Q().to(DEV), Tiny2Q().to(DEV)
    m_e.load_state_dict(m_g.state_dict())
    emb = torch.randn(B, 6, device=DEV)
    _, _, T = _case(2, B, L, 1, 1.0)
    Tt = _t(T).to(torch.complex64)
    step = GraphedTrainStep(m_g, B=B, emb_shape=(6,), monte_carlo=M, lr=1e-2, seed=9, target_dim=4)
    opt = torch.optim.Adam(m_e.parameters(), lr=1e-2)
    spec = SigmaSpec(0.6, 0.05)
    for it in range(1, 4):
        lg = step(emb, Tt, spec).item()
        opt.zero_grad(set_to_none=True)
        loss, _ = uq.fused_propagate_loss_su4(m_e(emb), Tt, monte_carlo=M, sigma=spec.sigma, seed=9, offset=it)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m_e.parameters(), max_norm=1.0)
        opt.step()
        assert abs(lg - loss.item()) < 1e-5 * max(1.0, abs(loss.item())), (it, lg, loss.item())
    for a, b in zip(m_g.parameters(), m_e.parameters()):
        assert torch.allclose(a, b, atol=1e-5)


def test_su4_train_split_kernel_equals_one_thread_per_sample_kernel():
    """BASELINE config 4 regime (one target, many samples, Philox): the train-split kernel (library plan) against the
    one-sample-per-thread kernel (UQOC_FLAG_WPS1) -- same samples, fidelities to FP32 rounding, gradients to 1e-5 --
    plus weights, explicit errors and an odd train length whose last warp gets a short chunk."""
    for B, L, M in ((1, 128, 32768), (2, 37, 1000), (1, 65, 100)):
        pulses, err, T = _case(7 + L, B, L, M, 1.0)
        p0 = _t(pulses, torch.float32)
        out = []
        for wps in (0, 1):
            p = p0.clone().requires_grad_(True)
            F = torch.empty(B * M, dtype=torch.float32, device=DEV)
            loss, mf = uq.fused_propagate_loss_su4(p, _t(T), monte_carlo=M, sigma=(0.8, 0.05), seed=3, offset=1, F_out=F,
                                                   flags=uq.tuning_flags(wps=wps))
            loss.backward()
            out.append((loss.item(), F.clone(), p.grad.clone()))
        assert abs(out[0][0] - out[1][0]) < 2e-6 * max(1.0, abs(out[1][0]))
        assert (out[0][1] - out[1][1]).abs().max().item() < 5e-6
        assert ((out[0][2] - out[1][2]).abs().max() / out[1][2].abs().max()).item() < 2e-5, (B, L, M)
