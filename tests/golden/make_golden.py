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

    golden_heads()
    golden_dcrab()
    golden_trajectory(score)
    golden_su4_cross()

    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


class _Returns(torch.nn.Module):
    """Stands in for a learned layer of a reference model: returns what ``fn`` says (our logits)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


# ------------------------------------------------------------------ f-3: pulse heads (GRAPE_model.py:64-91, universal_model.py:126-145)
def golden_heads():
    """The element-wise TAIL of the two reference pulse generators, run by the reference's own forward() on given
    logits: the learned layers are replaced by callables that return our logits (``GRAPE.layer``,
    ``UniversalQOCTransformer.head``), everything after them is the unmodified reference code.  The gradient of
    sum(w * pulses) with respect to the logits comes from the reference's autograd graph."""
    import importlib
    import tempfile
    gm = importlib.import_module("model.GRAPE_model")
    um = importlib.import_module("model.universal_model")
    out = {}
    torch.manual_seed(11)
    # ---- GRAPE head (train/GRAPE/model_params.json ranges)
    B, L = 3, 21
    space = {"phi": (-3.15, 3.15), "tau": (0.035, 0.07)}
    for dname, dt in (("f64", torch.float64), ("f32", torch.float32)):
        model = gm.GRAPE(space, L, device=torch.device("cpu"))
        logits = (torch.randn(B, L, 3, dtype=dt) * 2).requires_grad_(True)
        w = torch.randn(B, L, 2, dtype=dt)
        model.layer = _Returns(lambda rv: logits.reshape(B, L * 3))
        pulses = model(torch.zeros(B, 4, dtype=dt))
        (pulses * w).sum().backward()
        out[f"grape_{dname}_logits"] = logits.detach().numpy()
        out[f"grape_{dname}_w"] = w.numpy()
        out[f"grape_{dname}_pulses"] = pulses.detach().numpy()
        out[f"grape_{dname}_glogits"] = logits.grad.numpy()
    out["grape_ranges"] = model.param_ranges.numpy()        # float32, as the reference stores them (GRAPE_model.py:39-41)
    # ---- transformer head: demo L=400 config has tau range crossing zero (relu active) and a finetune base pulse
    B, L = 4, 19
    space = {"phi": (-3.15, 3.15), "tau": (-0.5, 0.5)}
    rv = torch.tensor([[0.6, 0.0, 0.8, 1.0], [0.0, -0.6, 0.8, 2.0], [-0.28, 0.96, 0.0, 3.0], [0.36, 0.48, -0.8, 0.5]])
    base = torch.stack([(torch.rand(L) * 2 - 1) * 3.0, 0.1 + 0.3 * torch.rand(L)], -1)
    with tempfile.TemporaryDirectory() as td:
        base_path = os.path.join(td, "base.pt")
        torch.save(base, base_path)
        for dname, dt in (("f64", torch.float64), ("f32", torch.float32)):
            for tag, fin in (("plain", False), ("finetune", base_path)):
                model = um.UniversalQOCTransformer(1, space, max_pulses=L, d_model=16, n_layers=1, n_heads=2, dropout=0.0,
                                                   finetune=fin)
                model.eval()
                logits = (torch.randn(B, L, 2, dtype=dt) * 2).requires_grad_(True)
                w = torch.randn(B, L, 2, dtype=dt)
                model.head = _Returns(lambda h: logits.reshape(B, 1, L * 2).expand(B, h.shape[1], L * 2))
                pulses = model(rv)
                (pulses * w).sum().backward()
                out[f"tf_{dname}_{tag}_logits"] = logits.detach().numpy()
                out[f"tf_{dname}_{tag}_w"] = w.numpy()
                out[f"tf_{dname}_{tag}_pulses"] = pulses.detach().numpy()
                out[f"tf_{dname}_{tag}_glogits"] = logits.grad.numpy()
    out["tf_ranges"] = model.param_ranges.numpy()           # float32 (universal_model.py:47-49)
    out["tf_rotation_vector"] = rv.numpy()
    out["tf_phi_offset"] = torch.atan2(rv[:, 1], rv[:, 0]).numpy()       # universal_model.py:93
    out["tf_base"] = base.numpy()
    np.savez_compressed(os.path.join(OUT, "heads.npz"), **out)


# ------------------------------------------------------------------ f-4: dCRAB objective (train/dCRAB/dCRAB.py:26-59)
def golden_dcrab():
    import importlib
    dc = importlib.import_module("train.dCRAB.dCRAB")
    X, Y, Z = dc.pauli_matrices()
    rng = np.random.default_rng(5)
    out = {}
    for tag, (T_total, dt, N, S) in (("a", (1.2, 0.01, 5, 24)), ("b", (0.6, 0.02, 3, 7))):
        t = np.arange(0, T_total, dt)
        omegas = rng.uniform(0.5, 8.0, N)
        params = rng.normal(0, 0.5, 1 + 2 * N)
        deltas, epss = rng.normal(0, 0.4, S), rng.normal(0, 0.05, S)
        Ut = np.array([[1, -1j], [-1j, 1]]) / np.sqrt(2)                  # X(pi/2), dCRAB.py:139
        out[f"{tag}_t"] = t
        out[f"{tag}_omegas"] = omegas
        out[f"{tag}_params"] = params
        out[f"{tag}_deltas"] = deltas
        out[f"{tag}_epss"] = epss
        out[f"{tag}_U_target"] = Ut
        out[f"{tag}_phi"] = dc.build_phi(params, t, omegas)
        out[f"{tag}_U"] = np.stack([dc.propagate(out[f"{tag}_phi"], t, d, e, X, Y, Z) for d, e in zip(deltas, epss)])
        out[f"{tag}_infidelity"] = dc.average_infidelity(params, t, omegas, Ut, deltas, epss, X, Y, Z)
    np.savez_compressed(os.path.join(OUT, "dcrab.npz"), **out)


# ------------------------------------------------------------------ f-1 / config 1: loss trajectory through the reference trainer
def golden_trajectory(score):
    """B = 4 targets of build_SU2_dataset(16), L = 16, M = 256, sigma = (0.4, 0.05), sharp_loss,
    5 steps of the UNMODIFIED UniversalModelTrainer.train_epoch (trainer.py:58-94) on the CPU with a fixed error
    closure, then evaluate() (trainer.py:101-121).  FP64 (tight) and FP32 (the reference's default) trajectories."""
    import importlib
    tm = importlib.import_module("model.universal_model_trainer")
    sys.path.insert(0, OUT)
    from tiny_model import TinyPulseModel
    torch.manual_seed(0)
    rv_all, U_all = score.build_SU2_dataset(16)
    rv, U_t = rv_all[:4].clone(), U_all[:4].clone()
    B, L, M, steps = 4, 16, 256, 5
    errors = torch.stack([score.get_ore_ple_error_distribution(B * M, 0.4, 0.05) for _ in range(steps + 1)])
    out = {"rotation_vector": rv.numpy(), "U_target": U_t.numpy(), "errors": errors.numpy(), "M": M, "lr": 1e-2}
    torch.manual_seed(3)
    init = TinyPulseModel(L)
    out.update({f"init.{k}": v.numpy() for k, v in init.state_dict().items()})
    for dname, dt in (("f64", torch.float64), ("f32", torch.float32)):
        model = TinyPulseModel(L)
        model.load_state_dict(init.state_dict())
        model = model.to(dt)
        cd = torch.complex128 if dt == torch.float64 else torch.complex64
        it = iter(errors.to(dt))
        closure = lambda n: next(it)
        trainer = tm.UniversalModelTrainer(model, score.batched_unitary_generator, score.get_ore_ple_error_distribution,
                                           fidelity_fn=score.fidelity, loss_fn=score.sharp_loss,
                                           optimizer=torch.optim.Adam(model.parameters(), lr=1e-2), monte_carlo=M, device="cpu")
        out[f"{dname}_pulses0"] = model(rv.to(dt)).detach().numpy()
        losses = [trainer.train_epoch(rv.to(dt), U_t.to(cd), closure) for _ in range(steps)]
        out[f"{dname}_losses"] = np.array(losses)
        out[f"{dname}_eval_fid"] = trainer.evaluate(rv.to(dt), U_t.to(cd), closure)
        out.update({f"{dname}_final.{k}": v.numpy() for k, v in model.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "trajectory.npz"), **out)


# ------------------------------------------------------------------ A9: SU(4) cross-oracle (torch.linalg.matrix_exp + tree, complex128)
def golden_su4_cross():
    """The two-qubit path has no reference implementation (README.md:86,122 promise only).  SURVEY.md A9 prescribes the
    oracle: torch.linalg.matrix_exp + the SCORE.py:131-142 tree on the builder-defined H in complex128, gradients from
    autograd -- oracle/torch_port.py::su4_* -- independent of the numpy eigendecomposition oracle and of the kernels."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import torch_port as tp
    rng = np.random.default_rng(21)
    out = {}
    for tag, (B, L, M, J, sd) in (("a", (2, 12, 9, 1.0, 1.0)), ("b", (1, 40, 16, 0.7, 2.0))):
        pulses = np.stack([rng.uniform(-3.15, 3.15, (B, L)), rng.uniform(-3.15, 3.15, (B, L)), rng.uniform(0.1, 0.5, (B, L))], -1)
        err = np.stack([rng.normal(0, sd, B * M), rng.normal(0, sd, B * M), rng.normal(0, 0.05, B * M)])
        T = np.linalg.qr(rng.normal(size=(B, 4, 4)) + 1j * rng.normal(size=(B, 4, 4)))[0]       # random unitary targets
        T[0] = np.diag([1, 1, 1, -1])                                                            # CZ
        p = torch.from_numpy(pulses).requires_grad_(True)
        loss, F, U = tp.su4_train_step(p, torch.from_numpy(T), torch.from_numpy(err), M, J, "sharp")
        loss.backward()
        out.update({f"{tag}_pulses": pulses, f"{tag}_error": err, f"{tag}_U_target": T, f"{tag}_M": M, f"{tag}_J": J,
                    f"{tag}_U": U.detach().numpy(), f"{tag}_F": F.detach().numpy(), f"{tag}_loss": loss.detach().numpy(),
                    f"{tag}_grad": p.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "su4_cross.npz"), **out)


if __name__ == "__main__":
    main()
