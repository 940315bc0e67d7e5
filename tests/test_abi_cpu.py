"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/uqoc.h declares, argument validation happens before any device work, and CPU tensors
are rejected (no fallback).  No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import universal_quantum_optimal_control_b200 as uq
from universal_quantum_optimal_control_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names |= set(re.findall(r"\b(uqoc_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported by libuqoc.so"
    # and the ctypes table covers the header one to one
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.uqoc_version() == 100
    # bad argument -> negative code + message, no device touched
    rc = lib.uqoc_su2_fwdbwd(None, None, None, None, 1, 1, 1, 0, 1.0, 0.05, 0, 0, None, None, None, None, None, 0, 7, 0, None)
    assert rc == -1
    assert b"dtype" in lib.uqoc_last_error()
    rc = lib.uqoc_su2_forward(None, None, None, 0, 4, 4, 0, 1.0, 0.05, 0, 0, None, None, None, None, None, 0, 0, 0, None)
    assert rc == -1 and b"B out of range" in lib.uqoc_last_error()


def test_workspace_query_is_pure():
    lib = _lib.lib()
    # B=1 with many samples splits across blocks -> needs workspace; many targets do not
    assert lib.uqoc_su2_workspace_bytes(1, 256, 65536, 0, 0) > 0
    assert lib.uqoc_su2_workspace_bytes(4096, 256, 4096, 0, 0) == 256      # the ticket area only
    assert lib.uqoc_su2_workspace_bytes(0, 256, 16, 0, 0) == 0


def test_shape_validation_matches_reference():
    # SCORE.py:99-100: ValueError("'pulses' must have shape (B, L, 2)")
    with pytest.raises(ValueError, match=r"'pulses' must have shape \(B, L, 2\)"):
        uq.batched_unitary_generator(torch.zeros(3, 4, 3), torch.zeros(2, 3))
    with pytest.raises(ValueError, match=r"'pulses' must have shape \(B, L, 2\)"):
        uq.batched_unitary_generator(torch.zeros(3, 4), torch.zeros(2, 3))
    with pytest.raises(ValueError):
        uq.fused_propagate_loss(torch.zeros(3, 4, 5), torch.zeros(3, 2, 2), monte_carlo=4)


def test_cpu_tensors_are_rejected_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uq.batched_unitary_generator(torch.zeros(3, 4, 2), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uq.fused_propagate_loss(torch.zeros(3, 4, 2), torch.zeros(3, 2, 2, dtype=torch.complex64), monte_carlo=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        uq.fidelity(torch.zeros(3, 2, 2, dtype=torch.complex64), torch.zeros(3, 2, 2, dtype=torch.complex64), 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "universal_quantum_optimal_control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "/root/reference" not in src, f


def test_tuning_flags_packing():
    f = uq.tuning_flags(st=4, lps=1, splits=7, fast_sincos=True)
    assert f & 1 and (f >> 8) & 0xF == 4 and (f >> 12) & 0x3F == 1 and (f >> 18) & 0xFFF == 7


def test_python_flag_words_match_the_header():
    """tuning_flags() / the module constants encode exactly the UQOC_FLAG_* bits of include/uqoc.h."""
    from universal_quantum_optimal_control_b200 import graphs, ops
    src = open(os.path.join(ROOT, "include", "uqoc.h")).read()
    bits = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+UQOC_FLAG_([A-Z0-9_]+)\s+(\d+)u", src)}
    assert bits["FAST_SINCOS"] == _lib.FLAG_FAST_SINCOS == uq.tuning_flags(fast_sincos=True)
    assert bits["NO_PACKED"] == uq.tuning_flags(no_packed=True)
    assert bits["NO_TABLE"] == uq.tuning_flags(no_table=True)
    assert bits["RNG_FROM_DEVICE"] == graphs.FLAG_RNG_FROM_DEVICE
    assert bits["WPS4"] == uq.tuning_flags(wps=4) and bits["WPS1"] == uq.tuning_flags(wps=1)
    assert bits["SU4_PADE"] == uq.tuning_flags(su4_pade=True)
    assert bits["RAW_TARGET"] == ops.FLAG_RAW_TARGET
    assert uq.tuning_flags(st=4, lps=8, splits=37) == (4 << 8) | (8 << 12) | (37 << 18)
    # peer-exchange sizing helpers are pure host functions
    lib = _lib.lib()
    assert lib.uqoc_peer_data_bytes(513, 8, _lib.F32) == 2 * 8 * 544 * 8     # {value, epoch} words: 8 bytes per real
    assert lib.uqoc_peer_flag_bytes(8) == 8 * 1024 * 4
    assert lib.uqoc_peer_data_bytes(0, 8, _lib.F32) == 0


def test_new_entry_points_reject_cpu_tensors_and_bad_shapes():
    """Folded head / single-call step (round 2): reference-style shape errors, no CPU fallback."""
    from universal_quantum_optimal_control_b200 import ops
    T = torch.eye(2, dtype=torch.complex64)[None].expand(3, -1, -1)
    with pytest.raises(ValueError, match=r"must have shape \(B, L, 3\)"):
        uq.fused_head_propagate_loss(torch.zeros(3, 5, 2), T, head="grape", pulse_ranges=((-3, 3), (0.1, 0.5)), monte_carlo=4)
    with pytest.raises(ValueError, match="unknown head"):
        uq.fused_head_propagate_loss(torch.zeros(3, 5, 2), T, head="mlp", pulse_ranges=((-3, 3), (0.1, 0.5)), monte_carlo=4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        uq.fused_head_propagate_loss(torch.zeros(3, 5, 2), T, head="transformer", pulse_ranges=((-3, 3), (0.1, 0.5)), monte_carlo=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.FusedStep(3, 5, 4, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        uq.su4_unitary_generator(torch.zeros(3, 5, 3), torch.zeros(3, 3))


def test_pipelined_step_chunk_bounds_cover_every_target_once():
    """PipelinedStep.chunk_bounds (pure host logic): contiguous, exhaustive, non-empty chunks for explicit counts, the
    ramped sizes and the wave-sized automatic choice (148 SMs x 5 blocks = 740 targets per wave)."""
    from universal_quantum_optimal_control_b200.pipeline import PipelinedStep as P
    for B in (1, 2, 5, 37, 64, 100, 739, 740, 1500, 2035, 2036, 2500, 4096, 10000):
        for chunks, ramp in (("auto", True), (1, False), (2, False), (3, True), (4, True), (6, True), (8, False), (50, True)):
            b = P.chunk_bounds(B, chunks, ramp, 148)
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(x1 == y0 for (_, x1), (y0, _) in zip(b, b[1:]))
            assert all(x1 > x0 for x0, x1 in b)
            if chunks != "auto":
                assert len(b) <= max(1, min(chunks, B))
    sizes = [x1 - x0 for x0, x1 in P.chunk_bounds(4096, "auto", True, 148)]
    assert sizes[:2] == [185, 370] and sizes[2:6] == [740] * 4 and sizes[-1] < sizes[-2] < 740      # small outer chunks, whole waves inside
    assert [x1 - x0 for x0, x1 in P.chunk_bounds(4096, 4, False)] == [1024] * 4
