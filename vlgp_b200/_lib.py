"""ctypes binding of libvlgp_b200.so (include/vlgp_b200.h) -- the only way the Python host reaches the GPU.

There is no CPU fallback: if the shared library is missing, cannot be loaded, or no sm_100 device is present, every
entry point of the package raises (``VlgpNativeError``) instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["VlgpNativeError", "load", "lib_path", "EXPORTS", "nccl_path"]

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)
ctx_p = C.c_void_p

# name -> (restype, argtypes); must list every VLGP_API symbol of include/vlgp_b200.h (tests/test_abi.py checks it)
EXPORTS = {
    "vlgp_create": (C.c_int, [C.c_int, C.POINTER(ctx_p)]),
    "vlgp_destroy": (C.c_int, [ctx_p]),
    "vlgp_last_error": (C.c_char_p, [ctx_p]),
    "vlgp_device_info": (C.c_int, [ctx_p, c_int_p, c_int_p, c_int_p, C.POINTER(C.c_uint64), C.c_char_p]),
    "vlgp_sync": (C.c_int, [ctx_p]),
    "vlgp_timer_start": (C.c_int, [ctx_p]),
    "vlgp_timer_stop": (C.c_int, [ctx_p, C.POINTER(C.c_float)]),
    "vlgp_counters": (C.c_int, [ctx_p, c_i64_p]),
    "vlgp_set_model": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_u8_p, C.c_double, C.c_double]),
    "vlgp_set_regressors": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_trials_set_x": (C.c_int, [ctx_p, C.c_int, c_double_p]),
    "vlgp_set_params": (C.c_int, [ctx_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "vlgp_get_params": (C.c_int, [ctx_p] + [c_double_p] * 7),
    "vlgp_trials_create": (C.c_int, [ctx_p, C.c_int, c_i32_p, c_int_p]),
    "vlgp_trials_free": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_trials_set_y": (C.c_int, [ctx_p, C.c_int, C.c_void_p, C.c_int]),
    # the two pointer/row tables are passed as packed byte strings (uint64 / int64), see _fastpack.pointers
    "vlgp_trials_set_y_parts": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, c_int_p]),
    "vlgp_trials_project_y": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "vlgp_trials_set_state": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "vlgp_trials_set_state_parts": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "vlgp_trials_get_state_parts": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "vlgp_trials_get_state": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "vlgp_make_cholesky": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_get_cholesky": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p, c_i32_p, c_i32_p]),
    "vlgp_set_cholesky": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p]),
    "vlgp_estep": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_double, C.c_int, c_int_p]),
    "vlgp_estep_subset": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_double, C.c_int, c_i32_p, C.c_int, c_int_p]),
    "vlgp_trials_copy_rows": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i64_p, c_i64_p, C.c_int64]),
    "vlgp_update_w": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_update_v": (C.c_int, [ctx_p, C.c_int, c_int_p]),
    "vlgp_mstep": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                             c_int_p]),
    "vlgp_mstep_begin": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "vlgp_mstep_end": (C.c_int, [ctx_p, c_int_p]),
    "vlgp_hstep_prepare": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_hstep_objective": (C.c_int, [ctx_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p]),
    "vlgp_hstep_objective_batch": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i32_p, c_double_p, c_double_p, c_double_p,
                                             c_i32_p]),
    "vlgp_hstep_optimize": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i32_p, c_double_p, c_double_p, c_i32_p, C.c_double,
                                      c_double_p, c_double_p, c_i32_p, c_i32_p, c_i32_p]),
    "vlgp_lbfgsb_new": (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int,
                                  C.c_double, C.POINTER(C.c_void_p)]),
    "vlgp_lbfgsb_advance": (C.c_int, [C.c_void_p, C.c_double, c_double_p, c_double_p, c_int_p]),
    "vlgp_lbfgsb_info": (C.c_int, [C.c_void_p, c_double_p, c_int_p, c_int_p, c_int_p]),
    "vlgp_lbfgsb_free": (C.c_int, [C.c_void_p]),
    "vlgp_posterior_cov": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_double, c_double_p]),
    "vlgp_latent_affine": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p]),
    "vlgp_latent_affine_rows": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_i64_p, C.c_int64]),
    "vlgp_norms": (C.c_int, [ctx_p, C.c_int, c_double_p]),
    "vlgp_latent_moments": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_i64_p]),
    "vlgp_comm_unique_id": (C.c_int, [ctx_p, C.c_char_p, C.c_char_p]),
    "vlgp_comm_init": (C.c_int, [ctx_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]),
    "vlgp_comm_allreduce": (C.c_int, [ctx_p, c_double_p, C.c_int, C.c_int]),
    "vlgp_comm_allreduce_bulk": (C.c_int, [ctx_p, c_double_p, C.c_int64]),
    "vlgp_shm_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vlgp_shm_allreduce": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int]),
    "vlgp_shm_close": (C.c_int, [C.c_void_p, C.c_int]),
    "vlgp_comm_attach_shm": (C.c_int, [ctx_p, C.c_void_p]),
    "vlgp_comm_enable_p2p": (C.c_int, [ctx_p, c_int_p]),
    "vlgp_host_f64_to_u8": (C.c_int, [c_double_p, C.POINTER(C.c_ubyte), C.c_int64]),
    "vlgp_host_pack_isa": (C.c_int, []),
    "vlgp_host_pool_selftest": (C.c_int, [C.c_int, C.c_int, c_int_p]),
    "vlgp_peak_fp64": (C.c_int, [ctx_p, c_double_p, c_double_p]),
    "vlgp_peak_hbm": (C.c_int, [ctx_p, C.c_uint64, c_double_p]),
    "vlgp_flush_l2": (C.c_int, [ctx_p]),
    "vlgp_gpfa_estep": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "vlgp_gpfa_stats": (C.c_int, [ctx_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "vlgp_set_precision": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_trials_prefetch_state": (C.c_int, [ctx_p, C.c_int, C.c_int]),
    "vlgp_trials_prefetch_state_into": (C.c_int, [ctx_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vlgp_trials_prefetch_take": (C.c_int, [ctx_p, C.c_int, C.c_int, c_int_p]),
    "vlgp_trials_prefetch_wait": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "vlgp_host_free": (C.c_int, [C.c_void_p]),
    "vlgp_profile_enable": (C.c_int, [ctx_p, C.c_int]),
    "vlgp_profile_get": (C.c_int, [ctx_p, C.c_int, c_double_p, c_i64_p]),
}


class VlgpNativeError(RuntimeError):
    """The native sm_100a library is unavailable or reported an error.  Never swallowed: there is no fallback path."""


def lib_path() -> str:
    return os.environ.get("VLGP_B200_LIB", os.path.join(HERE, "libvlgp_b200.so"))


def load():
    """dlopen the library once and attach prototypes.  Raises VlgpNativeError when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise VlgpNativeError(
            "%s not found: build it with `python -m vlgp_b200.build` (needs nvcc); there is no CPU fallback" % path)
    try:
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    except OSError as e:  # pragma: no cover
        raise VlgpNativeError("cannot load %s: %s" % (path, e)) from e
    for name, (res, args) in EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise VlgpNativeError("%s does not export %s (stale build?)" % (path, name)) from e
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def nccl_path() -> str:
    """Path of the libnccl to dlopen: the one bundled with torch's wheels when present (what the driver's launcher
    pairs with the installed driver), else the system library."""
    cand = os.environ.get("VLGP_NCCL_LIB")
    if cand:
        return cand
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:  # pragma: no cover
        pass
    return "libnccl.so.2"


def as_f64(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)
