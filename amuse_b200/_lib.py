"""ctypes binding of include/amuse_b200.h -- the only place the shared library is loaded.

Fails loudly: a missing library is an ImportError-class failure (``AmuseLibraryError``), never a
silent fallback to another implementation.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(os.environ.get("AMUSE_B200_LIB", Path(__file__).resolve().parent / "lib" / "libamuse_b200.so"))

# every symbol include/amuse_b200.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "amuse_create", "amuse_destroy", "amuse_last_error", "amuse_version", "amuse_load_weights",
    "amuse_finalize_weights", "amuse_reserve", "amuse_denoise", "amuse_denoiser_eps", "amuse_decode",
    "amuse_rot6d_to_axis_angle", "amuse_diffusion_backward", "amuse_diffusion_backward_host",
    "amuse_ast_features", "amuse_schedule", "amuse_launch_count", "amuse_profile_arm", "amuse_profile_read",
    "amuse_debug_tc_gemm", "amuse_debug_attn_profile", "amuse_debug_set_decode_plan", "amuse_fbank", "amuse_encode", "amuse_motion_to_feats",
    "amuse_debug_philox_normals",
]

AMUSE_OK = 0
SAMPLER = {"ddim": 0, "ddpm": 1}
ERRORS = {-1: "AMUSE_E_INVALID", -2: "AMUSE_E_STATE", -3: "AMUSE_E_CUDA", -4: "AMUSE_E_MISSING",
          -5: "AMUSE_E_UNSUPPORTED"}


class AmuseLibraryError(RuntimeError):
    pass


class AmuseError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libamuse_b200.so once and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.is_file():
        raise AmuseLibraryError(
            f"{LIB_PATH} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  amuse_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    p, i, f, u64, i64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_int64
    lib.amuse_create.argtypes = [C.POINTER(p), i]
    lib.amuse_destroy.argtypes = [p]
    lib.amuse_destroy.restype = None
    lib.amuse_last_error.argtypes = [p]
    lib.amuse_last_error.restype = C.c_char_p
    lib.amuse_version.argtypes = []
    lib.amuse_version.restype = C.c_char_p
    lib.amuse_load_weights.argtypes = [p, C.c_char_p, p, C.POINTER(i64), i, i]
    lib.amuse_finalize_weights.argtypes = [p, p]
    lib.amuse_reserve.argtypes = [p, i, i]
    lib.amuse_denoise.argtypes = [p, i, i, i, f, i, p, p, p, p, p, u64, u64, p, p]
    lib.amuse_denoiser_eps.argtypes = [p, i, i, p, p, p, p, p, p]
    lib.amuse_decode.argtypes = [p, i, p, p, p, p, p]
    lib.amuse_rot6d_to_axis_angle.argtypes = [p, i64, p, p, p]
    lib.amuse_diffusion_backward.argtypes = [p, i, i, i, f, i, p, p, p, p, p, u64, u64, p, p, p, p, p]
    lib.amuse_diffusion_backward_host.argtypes = [p, i, i, i, f, i, p, p, p, p, p, u64, u64, p, p, p]
    lib.amuse_ast_features.argtypes = [p, i, p, p, p, p, p]
    lib.amuse_schedule.argtypes = [p, i, i, f, C.POINTER(C.c_int32), C.POINTER(f)]
    lib.amuse_launch_count.argtypes = [p]
    lib.amuse_launch_count.restype = i64
    lib.amuse_profile_arm.argtypes = [p, i]
    lib.amuse_profile_read.argtypes = [p, C.POINTER(i64), i]
    lib.amuse_fbank.argtypes = [p, i, i, p, f, f, p, p]
    lib.amuse_debug_attn_profile.argtypes = [p, i, p, i]
    lib.amuse_debug_set_decode_plan.argtypes = [p, i, i]
    lib.amuse_debug_philox_normals.argtypes = [p, u64, u64, i, i, p, p]
    lib.amuse_encode.argtypes = [p, i, p, p, p, p]
    lib.amuse_motion_to_feats.argtypes = [p, i64, p, p, p, p]
    lib.amuse_debug_tc_gemm.argtypes = [p, i, i, i, i, p, p, i, p, p, p, p, p, i, p, p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("amuse_destroy",):
            fn.restype = C.c_int
    _lib = lib
    return lib
