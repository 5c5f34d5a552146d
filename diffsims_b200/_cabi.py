"""ctypes binding of ``libdiffsims_b200.so`` (declared in include/diffsims_b200.h).

This is the ONLY route from the Python API mirror to the sm_100a kernels.  There is no
CPU fallback: a missing library or a missing CUDA device raises.
"""
import ctypes
from ctypes import c_char_p, c_double, c_int32, c_void_p
from pathlib import Path

import os

# (DIFFSIMS_B200_LIB: an instrumented build of the same sources, used by tools/ only)
_LIB_PATH = Path(os.environ.get("DIFFSIMS_B200_LIB") or Path(__file__).resolve().parent / "libdiffsims_b200.so")
_lib = None

# every symbol include/diffsims_b200.h declares
SYMBOLS = ("ds_abi_version", "ds_last_error", "ds_set_option", "ds_get_option", "ds_structure_factors_scratch_bytes", "ds_structure_factors", "ds_pack_gtable",
           "ds_simulate", "ds_render_scratch_bytes", "ds_render_launch_count", "ds_render", "ds_pack_csr", "ds_quantize_u16", "ds_polar_flatten", "ds_library_pixel_coords",
           "ds_beam_grid_num_blocks", "ds_beam_grid", "ds_beam_points_num_blocks", "ds_beam_points", "ds_so3_grid_num_blocks", "ds_so3_grid")
ABI_VERSION = 2


class NativeLibraryError(ImportError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise NativeLibraryError(
            f"{_LIB_PATH} is missing: build it with `python -m diffsims_b200.build` "
            "(nvcc, sm_100a). diffsims_b200 has no CPU fallback.")
    L = ctypes.CDLL(str(_LIB_PATH))
    for s in SYMBOLS:
        if not hasattr(L, s):
            raise NativeLibraryError(f"{_LIB_PATH} does not export {s}")
    L.ds_abi_version.restype = c_int32
    L.ds_last_error.restype = c_char_p
    if L.ds_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version {L.ds_abi_version()} != {ABI_VERSION}: rebuild the library")
    P, I, D = c_void_p, c_int32, c_double
    L.ds_structure_factors.argtypes = [P, I, P, P, I, P, P, I, P, P, P, I, P, P, P, I, P]
    L.ds_structure_factors_scratch_bytes.argtypes = [I, I]
    L.ds_pack_gtable.argtypes = [P, I, P, P, P, I, D]
    L.ds_simulate.argtypes = [P, I, P, I, P, P, P, D, D, D, D, I, D, D, D, I, P, P, P, P, P, P, I, P, P, P]
    L.ds_render.argtypes = [P, I, I, P, P, P, I, I, D, D, D, D, I, I, D, I, D, I, P, P, D]
    L.ds_pack_csr.argtypes = [P, I, I, P, P, P, P, P, P, P, P]
    L.ds_quantize_u16.argtypes = [P, ctypes.c_int64, P, P]
    L.ds_polar_flatten.argtypes = [P, I, I, P, P, P, I, I, P, I, P, P, P, P]
    L.ds_library_pixel_coords.argtypes = [P, I, I, P, P, D, D, D, D, D, D, P]
    L.ds_beam_grid.argtypes = [P, I, I, P, I, P, D, P, P, P, P]
    L.ds_beam_grid_num_blocks.argtypes = [I]
    L.ds_beam_points.argtypes = [P, I, ctypes.c_int64, P, I, P, D, P, P, P, P]
    L.ds_beam_points_num_blocks.argtypes = [ctypes.c_int64]
    L.ds_so3_grid.argtypes = [P, I, I, I, I, P, D, P, P, P, P, P]
    L.ds_so3_grid_num_blocks.argtypes = [I]
    L.ds_set_option.argtypes = [c_char_p, I]
    L.ds_get_option.argtypes = [c_char_p, P]
    for s in SYMBOLS[2:]:
        getattr(L, s).restype = c_int32
    L.ds_render_launch_count.argtypes = [I, I, I, I, I, D]
    L.ds_render_scratch_bytes.argtypes = [I, I]
    L.ds_render_scratch_bytes.restype = ctypes.c_int64
    L.ds_structure_factors_scratch_bytes.restype = ctypes.c_int64
    L.ds_beam_grid_num_blocks.restype = ctypes.c_int64
    L.ds_beam_points_num_blocks.restype = ctypes.c_int64
    L.ds_so3_grid_num_blocks.restype = ctypes.c_int64
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().ds_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def set_option(name, value):
    """Process-wide schedule option of the native library (see include/diffsims_b200.h); -1 = default."""
    check(lib().ds_set_option(name.encode(), int(value)), "ds_set_option")


def get_option(name):
    v = c_int32(0)
    check(lib().ds_get_option(name.encode(), ctypes.byref(v)), "ds_get_option")
    return v.value
