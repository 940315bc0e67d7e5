"""ctypes loader of the plain-C oracle (oracle/uqoc_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of uqoc_oracle.c.  ``build()`` compiles the C file with gcc into
``oracle/_build/liboracle_c.so`` (git-ignored; it travels to the GPU box with the snapshot, and is rebuilt on
demand where gcc is present).  The functions mirror oracle/uqoc_oracle.py so the tests can run all three
formulations against the same golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "uqoc_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: the C oracle cannot be built")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [gcc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        h.uqoc_c_su2.restype = C.c_int
        h.uqoc_c_su2.argtypes = [dp, dp, dp, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        h.uqoc_c_threads.restype = C.c_int
        _lib = h
    return _lib


def threads() -> int:
    return int(lib().uqoc_c_threads())


def _target_reals(U_target: np.ndarray, B: int) -> np.ndarray:
    T = np.ascontiguousarray(np.asarray(U_target, dtype=np.complex128).reshape(B, 2, 2))
    return np.ascontiguousarray(T.view(np.float64).reshape(B, 8))


def fidelity_sum_and_grad(pulses, U_target, error, M: int, want_grad: bool = True, want_U: bool = False):
    """(Fsum (B,), grad (B, L, 2) or None, F (B*M,), [U_out (B*M, 2, 2)]) -- twin of
    oracle/uqoc_oracle.py::fidelity_sum_and_grad (trainer.py:80-90 sample layout)."""
    pulses = np.ascontiguousarray(np.asarray(pulses, dtype=np.float64))
    if pulses.ndim != 3 or pulses.shape[-1] != 2:
        raise ValueError("'pulses' must have shape (B, L, 2)")            # SCORE.py:99-100
    B, L, _ = pulses.shape
    error = np.ascontiguousarray(np.asarray(error, dtype=np.float64))
    if error.shape != (2, B * M):
        raise ValueError(f"'error' must have shape (2, {B * M})")
    T = _target_reals(U_target, B)
    F = np.empty(B * M)
    Fsum = np.empty(B)
    grad = np.empty((B, L, 2)) if want_grad else None
    U = np.empty((B * M, 2, 2, 2)) if want_U else None
    rc = lib().uqoc_c_su2(pulses, T, error, B, L, M, F.ctypes.data, Fsum.ctypes.data,
                          None if grad is None else grad.ctypes.data, None if U is None else U.ctypes.data)
    if rc != 0:
        raise MemoryError("uqoc_c_su2 failed")
    out = (Fsum, grad, F)
    if want_U:
        out += (U.view(np.complex128).reshape(B * M, 2, 2),)
    return out
