"""ctypes binding of libuqoc.so (include/uqoc.h).  No CPU fallback: if the shared library is
missing and cannot be built the import of any op raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# UQOC_LIB: load another build of the same ABI (kernel-tuning experiments, tools/variants.sh)
LIB_PATH = os.environ.get("UQOC_LIB") or os.path.join(HERE, "lib", "libuqoc.so")

F32, F64 = 0, 1
FLAG_FAST_SINCOS = 1
FLAG_NO_FIN = 1 << 30          # UQOC_FLAG_NO_FIN: no in-kernel epilogue (separate reduction / loss launches)
FLAG_NO_FAT = 1 << 31          # UQOC_FLAG_NO_FAT: no fat blocks for few-target shapes
LOSS_KINDS = {"sharp": 0, "nll": 1, "infidelity": 2, "none": 3}

_lib = None
_lock = threading.Lock()

_vp, _i64, _u64, _dbl, _int, _uint = C.c_void_p, C.c_int64, C.c_uint64, C.c_double, C.c_int, C.c_uint

# name -> (restype, argtypes); mirrors include/uqoc.h one to one
SIGNATURES = {
    "uqoc_version": (_int, []),
    "uqoc_last_error": (C.c_char_p, []),
    "uqoc_su2_target_coeffs": (_int, [_vp, _i64, _vp, _int, _vp]),
    "uqoc_su2_workspace_bytes": (_i64, [_i64, _i64, _i64, _int, _uint]),
    "uqoc_su2_fwdbwd": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64,
                               _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su2_fwdbwd_slice": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64,
                                     _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_peer_data_bytes": (_i64, [_i64, _int, _int]),
    "uqoc_peer_flag_bytes": (_i64, [_int]),
    "uqoc_su2_fwdbwd_peer": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64,
                                    _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.c_uint32, _int, _uint, _vp]),
    "uqoc_su2_fwdbwd_peer_loss": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64, _int, _dbl, _dbl,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64), C.c_uint32, _int, _uint, _vp]),
    "uqoc_su2_fwdbwd_loss": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64, _int, _dbl, _dbl,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su2_forward": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64,
                                _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su2_forward_grid": (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su2_forward_sigmas": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su2_generator_forward": (_int, [_vp, _vp, _i64, _i64, _vp, _int, _uint, _vp]),
    "uqoc_su2_generator_backward": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _int, _uint, _vp]),
    "uqoc_su4_workspace_bytes": (_i64, [_i64, _i64, _i64, _int, _uint]),
    "uqoc_su4_fwdbwd": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _dbl, _u64, _u64,
                               _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su4_forward": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl, _dbl, _u64, _u64,
                                _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_su4_generator_backward": (_int, [_vp, _vp, _vp, _i64, _i64, _dbl, _vp, _int, _uint, _vp]),
    "uqoc_pulse_head_forward": (_int, [_vp, _vp, _vp, _i64, _i64, _int, C.POINTER(_dbl), _dbl, _vp, _int, _vp]),
    "uqoc_pulse_head_backward": (_int, [_vp, _vp, _vp, _i64, _i64, _int, C.POINTER(_dbl), _dbl, _vp, _int, _vp]),
    "uqoc_su2_head_step": (_int, [_vp, _int, C.POINTER(_dbl), _dbl, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _dbl, _dbl, _u64, _u64,
                                  _int, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _uint, _vp]),
    "uqoc_loss_finalize": (_int, [_vp, _i64, _dbl, _int, _dbl, _dbl, _vp, _i64, _vp, _int, _vp]),
    "uqoc_fidelity_forward": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _int, _vp]),
    "uqoc_fidelity_backward": (_int, [_vp, _vp, _vp, _i64, _int, _i64, _vp, _int, _vp]),
    "uqoc_sum": (_int, [_vp, _i64, _vp, _vp, _i64, _int, _vp]),
    "uqoc_philox_errors": (_int, [_i64, _i64, _i64, _dbl, _dbl, _u64, _u64, _vp, _int, _vp]),
    "uqoc_fp32_peak_probe": (_int, [_int, _int, C.POINTER(_dbl), C.POINTER(_dbl)]),
}


class UqocError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (building in-tree first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                from .build import build_lib
                build_lib()
            except Exception as e:  # noqa: BLE001
                raise UqocError(
                    f"libuqoc.so not found at {LIB_PATH} and could not be built ({e}); "
                    "the CUDA extension is required -- there is no CPU fallback") from e
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                continue  # optional entry point not built yet
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def register(name, restype, argtypes):
    """Late registration hook for entry points declared by other modules (SU(4))."""
    SIGNATURES[name] = (restype, argtypes)
    if _lib is not None and hasattr(_lib, name):
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def check(rc: int, what: str = "uqoc") -> None:
    if rc != 0:
        msg = lib().uqoc_last_error().decode("utf-8", "replace")
        raise UqocError(f"{what} failed (rc={rc}): {msg}")
