"""The plain-C oracle (oracle/uqoc_oracle.c: complex 2x2 arithmetic, stored prefixes, running suffix) against the
golden vectors generated from the unmodified reference and against the numpy oracle.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle as co
from oracle import uqoc_oracle as orc


def test_c_oracle_builds_and_reports_threads():
    assert co.threads() >= 1


@pytest.mark.parametrize("tag", ["sd04", "sd07", "sd10"])
def test_c_oracle_matches_reference_train_step_golden(tag):
    g = load_golden("c1_train_step.npz")
    pulses, T, err, M = g["pulses"], g["U_target"], g[f"error_{tag}"], int(g["M"])
    Fsum, grad, F, U = co.fidelity_sum_and_grad(pulses, T, err, M, want_U=True)
    assert np.abs(F - g[f"F64_{tag}"]).max() < 1e-12                       # reference FP64 fidelities (SCORE.py:168-183)
    assert np.abs(U - g[f"U64_{tag}"]).max() < 1e-12                       # reference FP64 unitaries (SCORE.py:77-145)
    n = F.size
    for loss in ("sharp", "nll", "infidelity"):
        val, dval = orc.loss_and_dloss(Fsum.sum() / n, loss)
        assert abs(val - float(g[f"loss64_{loss}_{tag}"])) < 1e-12
        want = g[f"grad64_{loss}_{tag}"]
        assert np.abs(dval / n * grad - want).max() < 1e-11 * max(1.0, np.abs(want).max())   # autograd of the reference


@pytest.mark.parametrize("name", ["grape_L256.npz", "general_target.npz"])
def test_c_oracle_matches_grape_and_general_target_goldens(name):
    g = load_golden(name)
    M = int(g["M"])
    Fsum, grad, F, U = co.fidelity_sum_and_grad(g["pulses"], g["U_target"], g["error"], M, want_U=True)
    assert np.abs(F - g["F64"]).max() < 1e-12
    assert np.abs(U - g["U64"]).max() < 1e-12
    n = F.size
    val, dval = orc.loss_and_dloss(Fsum.sum() / n, "sharp")
    assert abs(val - float(g["loss64"])) < 1e-12
    assert np.abs(dval / n * grad - g["grad64"]).max() < 1e-11 * max(1.0, np.abs(g["grad64"]).max())


@pytest.mark.parametrize("B,L,M", [(1, 1, 1), (3, 7, 5), (2, 64, 300), (5, 33, 257)])
def test_c_oracle_equals_numpy_oracle(B, L, M):
    rng = np.random.default_rng(B * 100 + L)
    pulses = np.stack([rng.uniform(-7, 7, (B, L)), rng.uniform(-1.0, 2.0, (B, L))], -1)
    T = rng.normal(size=(B, 2, 2)) + 1j * rng.normal(size=(B, 2, 2))
    err = np.stack([rng.normal(0, 2, B * M), rng.normal(0, 0.1, B * M)])
    Fs0, g0, F0 = orc.fidelity_sum_and_grad(pulses, T, err, M)
    Fs1, g1, F1, U1 = co.fidelity_sum_and_grad(pulses, T, err, M, want_U=True)
    U0 = orc.batched_unitary_generator(np.repeat(pulses, M, 0), err)
    assert np.abs(F0 - F1).max() < 1e-12 and np.abs(Fs0 - Fs1).max() < 1e-10
    assert np.abs(g0 - g1).max() < 1e-11 * max(1.0, np.abs(g0).max())
    assert np.abs(U0 - U1).max() < 1e-12
