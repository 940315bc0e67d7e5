"""Pin the CPU oracle (oracle/uqoc_oracle.py, oracle/torch_port.py) against the
golden vectors produced by the unmodified reference (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest
import torch

from oracle import uqoc_oracle as orc
from oracle import torch_port as tp
from conftest import load_golden

SDS = ("sd04", "sd07", "sd10")


@pytest.mark.parametrize("tag", SDS)
def test_c1_unitary_and_fidelity(tag):
    g = load_golden("c1_train_step.npz")
    M = int(g["M"])
    pulses = g["pulses"].astype(np.float64)
    err = g[f"error_{tag}"].astype(np.float64)
    F, U = orc.per_sample_fidelity(pulses, g["U_target"].astype(np.complex128), err, M)
    assert np.abs(U - g[f"U64_{tag}"]).max() < 1e-13
    assert np.abs(F - g[f"F64_{tag}"]).max() < 1e-13
    # sequential product (grape_train.py:133-136) is the same matrix
    Fs, Us = orc.per_sample_fidelity(pulses, g["U_target"].astype(np.complex128), err, M, "sequential")
    assert np.abs(Us - g[f"U64_{tag}"]).max() < 1e-13


@pytest.mark.parametrize("tag", SDS)
@pytest.mark.parametrize("loss", ["sharp", "nll", "infidelity"])
def test_c1_loss_and_grad(tag, loss):
    g = load_golden("c1_train_step.npz")
    M = int(g["M"])
    val, grad, _ = orc.loss_and_grad(g["pulses"], g["U_target"], g[f"error_{tag}"], M, loss)
    ref_l = g[f"loss64_{loss}_{tag}"]
    ref_g = g[f"grad64_{loss}_{tag}"]
    assert abs(val - ref_l) < 1e-12 * max(1.0, abs(ref_l))
    assert np.abs(grad - ref_g).max() < 1e-11 * max(1.0, np.abs(ref_g).max())


@pytest.mark.parametrize("L", [1, 2, 3, 7, 33])
def test_ragged_lengths(L):
    g = load_golden("ragged_lengths.npz")
    M = int(g["M"])
    val, grad, F = orc.loss_and_grad(g[f"L{L}_pulses"], g[f"L{L}_U_target"], g[f"L{L}_error"], M)
    assert np.abs(F - g[f"L{L}_F64"]).max() < 1e-13
    assert abs(val - g[f"L{L}_loss64"]) < 1e-12
    assert np.abs(grad - g[f"L{L}_grad64"]).max() < 1e-11 * max(1.0, np.abs(g[f"L{L}_grad64"]).max())


def test_general_target():
    g = load_golden("general_target.npz")
    val, grad, F = orc.loss_and_grad(g["pulses"], g["U_target"], g["error"], int(g["M"]))
    assert np.abs(F - g["F64"]).max() < 1e-12
    assert abs(val - g["loss64"]) < 1e-12 * max(1, abs(g["loss64"]))
    assert np.abs(grad - g["grad64"]).max() < 1e-11 * max(1.0, np.abs(g["grad64"]).max())


def test_grape_L256():
    g = load_golden("grape_L256.npz")
    val, grad, F = orc.loss_and_grad(g["pulses"], g["U_target"], g["error"], int(g["M"]))
    assert np.abs(F - g["F64"]).max() < 1e-12
    assert np.abs(grad - g["grad64"]).max() < 1e-10 * max(1.0, np.abs(g["grad64"]).max())
    # the GRAPE script's complex64 sequential generator agrees to fp32 noise
    assert np.abs(F - g["Fseq32"]).max() < 5e-5


def test_grid_sweep():
    g = load_golden("grid_sweep.npz")
    N = g["errors"].shape[1]
    pulses = np.broadcast_to(g["pulse"].astype(np.float64), (N,) + g["pulse"].shape)
    U = orc.batched_unitary_generator(pulses, g["errors"].astype(np.float64))
    F = orc.fidelity(U, np.broadcast_to(g["U_target"].astype(np.complex128), (N, 2, 2)), 1)
    assert np.abs(U - g["U64"]).max() < 1e-13
    assert np.abs(F - g["F64"]).max() < 1e-13
    # layout of util.py:231-240: row-major (ORE, PLE) with ORE slowest
    assert g["errors"][0, 0] == g["errors"][0, 22] and g["errors"][1, 0] != g["errors"][1, 1]


@pytest.mark.parametrize("key", ["n03", "n04", "n06", "n08", "n09", "n12"])
def test_score_pulses(key):
    g = load_golden("score_pulses.npz")
    pulse = g[f"{key}_pulse"].astype(np.float64)
    K = g["probe"].shape[1]
    U = orc.batched_unitary_generator(np.broadcast_to(pulse, (K,) + pulse.shape), g["probe"])
    F = orc.fidelity(U, np.broadcast_to(g[f"{key}_U_target"], (K, 2, 2)), 1)
    assert np.abs(F - g[f"{key}_F64"]).max() < 1e-12
    # physics known-answers: the composite pulse implements X(n pi)
    assert F[0] > 1 - 1e-6 and F[1] > 0.998 and F[2] > 0.995


def test_torch_port_matches_golden():
    g = load_golden("c1_train_step.npz")
    M = int(g["M"])
    pulses = torch.from_numpy(g["pulses"]).double()
    T = torch.from_numpy(g["U_target"]).to(torch.complex128)
    err = torch.from_numpy(g["error_sd07"]).double()
    val, grad, F = tp.train_step_loss_and_grad(pulses, T, err, M, "sharp")
    assert abs(val.item() - g["loss64_sharp_sd07"]) < 1e-13
    assert (grad.numpy() - g["grad64_sharp_sd07"]).__abs__().max() < 1e-12
    assert np.abs(F.numpy() - g["F64_sd07"]).max() < 1e-13
    # fp32 leg reproduces the reference's fp32 numbers bit-for-bit (same ATen ops)
    val32, grad32, F32 = tp.train_step_loss_and_grad(pulses.float(), T.to(torch.complex64), err.float(), M)
    assert np.abs(F32.numpy() - g["F32_sd07"]).max() < 2e-6
    # sequential generator (complex64 only in the reference)
    Fs = tp.fidelity(tp.generator_sequential(pulses.float().repeat_interleave(M, 0), err.float()),
                     T.to(torch.complex64).repeat_interleave(M, 0), 1)
    assert np.abs(Fs.numpy() - g["F64_sd07"]).max() < 2e-5


def test_bad_shape_raises():
    with pytest.raises(ValueError):
        orc.batched_unitary_generator(np.zeros((3, 4, 3)), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        tp.generator_tree(torch.zeros(3, 4), torch.zeros(2, 3))


def test_philox_known_answers():
    # Random123 KAT
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_philox_errors_statistics_and_sharding():
    e = orc.philox_errors(4, 20000, 0.7, 0.05, seed=1234, offset=3)
    assert e.shape == (2, 80000)
    assert abs(e[0].mean()) < 0.01 and abs(e[0].std() - 0.7) < 0.01
    assert abs(e[1].mean()) < 0.001 and abs(e[1].std() - 0.05) < 0.001
    assert abs(np.corrcoef(e[0], e[1])[0, 1]) < 0.02
    # sample set is independent of how j is split across ranks
    a = orc.philox_errors(4, 100, 0.7, 0.05, 1234, 3, j0=0).reshape(2, 4, 100)
    b = orc.philox_errors(4, 50, 0.7, 0.05, 1234, 3, j0=50).reshape(2, 4, 50)
    assert np.array_equal(a[:, :, 50:], b)


def test_su4_oracle_selfconsistency():
    """SU(4) is builder-defined (parity unpinned): check the adjoint gradient
    against central finite differences and basic identities."""
    rng = np.random.default_rng(0)
    B, L, M = 2, 5, 3
    pulses = np.stack([rng.uniform(-3, 3, (B, L)), rng.uniform(-3, 3, (B, L)), rng.uniform(0.1, 0.5, (B, L))], -1)
    err = np.stack([rng.normal(0, 1, B * M), rng.normal(0, 1, B * M), rng.normal(0, 0.05, B * M)])
    T = orc.su4_unitary_generator(np.stack([rng.uniform(-3, 3, (B, 4)), rng.uniform(-3, 3, (B, 4)),
                                            rng.uniform(0.2, 1.0, (B, 4))], -1), np.zeros((3, B)))
    Fsum, grad, F = orc.su4_fidelity_sum_and_grad(pulses, T, err, M)
    U = orc.su4_unitary_generator(np.repeat(pulses, M, 0), err)
    assert np.abs(U @ np.conj(np.swapaxes(U, -1, -2)) - np.eye(4)).max() < 1e-12
    assert np.all(F <= 1 + 1e-12) and np.all(F >= 0.2 - 1e-12)
    h = 1e-6
    for idx in [(0, 0, 0), (1, 3, 1), (0, 4, 2), (1, 2, 2)]:
        pp, pm = pulses.copy(), pulses.copy()
        pp[idx] += h
        pm[idx] -= h
        fd = (orc.su4_fidelity_sum_and_grad(pp, T, err, M)[0].sum()
              - orc.su4_fidelity_sum_and_grad(pm, T, err, M)[0].sum()) / (2 * h)
        assert abs(fd - grad[idx]) < 1e-7 * max(1, abs(fd))
