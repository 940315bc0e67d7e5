"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference is imported as-is; the only accommodation is a stub for
``matplotlib`` (absent in this image; ``model/universal_model_trainer.py:13``
imports ``matplotlib.pyplot`` at module top).  ``visualize/util.py`` cannot be
imported (pwlf/qutip absent), so its SCORE tables (``util.py:47-112``) are
exec'd from a source slice of the file in place -- nothing is copied into the
repository except the numeric outputs written to ``tests/golden/*.npz``.
"""
import math
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("UQOC_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    sys.path.insert(0, REF)
    import importlib
    score = importlib.import_module("train.unitary_single_qubit_gate.universal_single_qubit_SCORE")
    grape = importlib.import_module("train.GRAPE.grape_train")
    return score, grape


def score_tables():
    """exec visualize/util.py lines 47-112 (angle_vec_dict, SCOREn_config)."""
    with open(os.path.join(REF, "visualize", "util.py")) as f:
        lines = f.readlines()
    src = "".join(lines[46:112])
    ns = {"np": np, "torch": torch, "math": math}
    exec(compile(src, "util.py[47:112]", "exec"), ns)
    return ns["angle_vec_dict"], ns["SCOREn_config"]


def ref_step(score, pulses, U_target, error, M, loss_name, dtype):
    """trainer.py:80-90 with the reference's own callables."""
    loss_fn = {"sharp": score.sharp_loss, "nll": score.negative_log_loss,
               "infidelity": score.infidelity_loss}[loss_name]
    p = pulses.to(dtype).clone().requires_grad_(True)
    e = error.to(dtype)
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    T = U_target.to(cd)
    p_mc = p.repeat_interleave(M, dim=0)
    t_mc = T.repeat_interleave(M, dim=0)
    U = score.batched_unitary_generator(p_mc, e)
    F = score.fidelity(U, t_mc, 1)
    loss = loss_fn(U, t_mc, score.fidelity, 1)
    loss.backward()
    return U.detach(), F.detach(), loss.detach(), p.grad.detach()


def main():
    score, grape = import_reference()
    torch.set_num_threads(4)

    # ---------------------------------------------------------------- config 1
    torch.manual_seed(0)                                   # SCORE.py:332
    _, U_all = score.build_SU2_dataset(16)                 # SCORE.py:215-251
    U_target = U_all[:4]
    B, L, M = 4, 16, 256
    phi = (torch.rand(B, L) * 2 - 1) * 3.15                # model_params.json:4
    tau = 0.1 + 0.4 * torch.rand(B, L)                     # model_params.json:5
    pulses = torch.stack([phi, tau], dim=-1)
    out = {"pulses": pulses.numpy(), "U_target": U_target.numpy(), "M": M}
    for sd in (0.4, 0.7, 1.0):
        error = score.get_ore_ple_error_distribution(B * M, sd, 0.05)
        tag = f"sd{int(sd * 10):02d}"
        out[f"error_{tag}"] = error.numpy()
        for loss_name in ("sharp", "nll", "infidelity"):
            U64, F64, l64, g64 = ref_step(score, pulses, U_target, error, M, loss_name, torch.float64)
            out[f"loss64_{loss_name}_{tag}"] = l64.numpy()
            out[f"grad64_{loss_name}_{tag}"] = g64.numpy()
        out[f"U64_{tag}"] = U64.numpy()
        out[f"F64_{tag}"] = F64.numpy()
        U32, F32, l32, g32 = ref_step(score, pulses, U_target, error, M, "sharp", torch.float32)
        out[f"F32_{tag}"] = F32.numpy()
        out[f"loss32_sharp_{tag}"] = l32.numpy()
        out[f"grad32_sharp_{tag}"] = g32.numpy()
    np.savez_compressed(os.path.join(OUT, "c1_train_step.npz"), **out)

    # ------------------------------------------------- odd / tiny L, odd sizes
    torch.manual_seed(1)
    out = {}
    for L in (1, 2, 3, 7, 33):
        B, M = 3, 5
        phi = (torch.rand(B, L) * 2 - 1) * 3.15
        tau = 0.1 + 0.4 * torch.rand(B, L)
        pulses = torch.stack([phi, tau], dim=-1)
        _, U_t = score.build_SU2_dataset(B, random=True)
        error = score.get_ore_ple_error_distribution(B * M, 1.0, 0.05)
        U64, F64, l64, g64 = ref_step(score, pulses, U_t, error, M, "sharp", torch.float64)
        out[f"L{L}_pulses"] = pulses.numpy()
        out[f"L{L}_U_target"] = U_t.numpy()
        out[f"L{L}_error"] = error.numpy()
        out[f"L{L}_U64"] = U64.numpy()
        out[f"L{L}_F64"] = F64.numpy()
        out[f"L{L}_loss64"] = l64.numpy()
        out[f"L{L}_grad64"] = g64.numpy()
    out["M"] = 5
    np.savez_compressed(os.path.join(OUT, "ragged_lengths.npz"), **out)

    # ----------------------- general (non-unitary, complex) target, larger L
    torch.manual_seed(2)
    B, L, M = 2, 64, 64
    phi = (torch.rand(B, L) * 2 - 1) * 3.15
    tau = 0.1 + 0.4 * torch.rand(B, L)
    pulses = torch.stack([phi, tau], dim=-1)
    U_t = torch.randn(B, 2, 2, dtype=torch.complex128) * 0.7
    error = score.get_ore_ple_error_distribution(B * M, 1.0, 0.05)
    U64, F64, l64, g64 = ref_step(score, pulses, U_t, error, M, "sharp", torch.float64)
    np.savez_compressed(os.path.join(OUT, "general_target.npz"), pulses=pulses.numpy(),
                        U_target=U_t.numpy(), error=error.numpy(), M=M, U64=U64.numpy(),
                        F64=F64.numpy(), loss64=l64.numpy(), grad64=g64.numpy())

    # ------------------------------------- long sequence (config-3 shaped, cut)
    torch.manual_seed(42)                                  # grape_train.py:322
    B, L, M = 1, 256, 512
    phi = (torch.rand(B, L) * 2 - 1) * 3.15                # train/GRAPE/model_params.json:3
    tau = 0.035 + 0.035 * torch.rand(B, L)                 # train/GRAPE/model_params.json:4
    pulses = torch.stack([phi, tau], dim=-1)
    _, U_t = score.build_SU2_dataset(B, random=True)
    error = score.get_ore_ple_error_distribution(B * M, 1.0, 0.05)
    U64, F64, l64, g64 = ref_step(score, pulses, U_t, error, M, "sharp", torch.float64)
    # the GRAPE script's own sequential generator (complex64 only, grape_train.py:133-136)
    Useq = grape.batched_unitary_generator(pulses.repeat_interleave(M, 0), error)
    Fseq = grape.fidelity(Useq, U_t.repeat_interleave(M, 0), 1)
    np.savez_compressed(os.path.join(OUT, "grape_L256.npz"), pulses=pulses.numpy(),
                        U_target=U_t.numpy(), error=error.numpy(), M=M, F64=F64.numpy(),
                        U64=U64.numpy(), loss64=l64.numpy(), grad64=g64.numpy(),
                        Fseq32=Fseq.numpy(), Useq32=Useq.numpy())

    # --------------------------- contour grid sweep (util.py:231-249, cut down)
    torch.manual_seed(0)
    L = 64
    pulse = torch.stack([(torch.rand(L) * 2 - 1) * math.pi, 0.1 + 0.4 * torch.rand(L)], dim=-1)
    X = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex64)
    U_t = torch.matrix_exp(-1j * X * (math.pi / 4))        # X(pi/2), built as SCORE.py:246-248
    ORE = torch.linspace(-3, 3, 41)
    PLE = torch.linspace(-0.15, 0.15, 23)
    Og, Pg = torch.meshgrid(ORE, PLE, indexing="ij")
    errors = torch.stack([Og.flatten(), Pg.flatten()], dim=0)
    N = errors.shape[1]
    Ug = score.batched_unitary_generator(pulse.double().expand(N, -1, -1), errors.double())
    Fg = score.fidelity(Ug, U_t.to(torch.complex128).expand(N, -1, -1), 1)
    Ug32 = score.batched_unitary_generator(pulse.expand(N, -1, -1), errors)
    Fg32 = score.fidelity(Ug32, U_t.expand(N, -1, -1), 1)
    np.savez_compressed(os.path.join(OUT, "grid_sweep.npz"), pulse=pulse.numpy(), U_target=U_t.numpy(),
                        ore=ORE.numpy(), ple=PLE.numpy(), errors=errors.numpy(), F64=Fg.numpy(),
                        U64=Ug.numpy(), F32=Fg32.numpy())

    # ------------------------------------ SCORE composite pulses (util.py:47-112)
    angle_vec_dict, SCOREn_config = score_tables()
    out = {}
    probe = torch.tensor([[0.0, 0.3, 0.0, -0.3, 0.5], [0.0, 0.0, 0.05, -0.05, 0.02]], dtype=torch.float64)
    for n in angle_vec_dict:
        pulse = SCOREn_config(n, 0.0)                      # (L,2) f32 [phi, angle]
        U_t = torch.matrix_exp(-1j * X.to(torch.complex128) * (n * math.pi / 2))   # X(n pi)
        K = probe.shape[1]
        U = score.batched_unitary_generator(pulse.double().expand(K, -1, -1), probe)
        F = score.fidelity(U, U_t.expand(K, -1, -1), 1)
        key = f"n{round(n * 12):02d}"                      # 12 n: 3,4,6,8,9,12
        out[f"{key}_pulse"] = pulse.numpy()
        out[f"{key}_U_target"] = U_t.numpy()
        out[f"{key}_F64"] = F.numpy()
        out[f"{key}_U64"] = U.numpy()
    out["probe"] = probe.numpy()
    np.savez_compressed(os.path.join(OUT, "score_pulses.npz"), **out)

    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
