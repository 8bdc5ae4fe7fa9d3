"""ctypes binding of the C ABI declared in include/spatpca_b200.h.

This is the Python-side twin of the Rcpp glue (r-package/src/rcpp_glue.cpp): it
only marshals column-major FP64 host buffers into the ABI and turns negative
statuses into exceptions.  There is no fallback of any kind: if the shared
library is missing, or there is no CUDA device, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspatpca_b200.so")

SPB_OK = 0
SPB_MAX_K = 32
ERR_NAMES = {
    -1: "SPB_ERR_INVALID_ARG", -2: "SPB_ERR_NO_DEVICE", -3: "SPB_ERR_CUDA", -4: "SPB_ERR_NOT_POSDEF",
    -5: "SPB_ERR_SINGULAR", -6: "SPB_ERR_POLAR", -7: "SPB_ERR_SVD", -8: "SPB_ERR_INTERNAL",
}


class SpatpcaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class CvStats(C.Structure):
    _fields_ = [
        ("n_admm_calls", C.c_int), ("admm_iters", C.c_longlong), ("n_inverses", C.c_int),
        ("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
        ("ms_omega", C.c_double), ("ms_prologue", C.c_double), ("ms_inverse", C.c_double),
        ("ms_admm", C.c_double), ("ms_cv_loss", C.c_double), ("ms_gamma", C.c_double),
        ("bytes_h2d", C.c_longlong), ("bytes_d2h", C.c_longlong), ("gpu_launches", C.c_int),
        ("n_cg_calls", C.c_int), ("cg_passes", C.c_longlong), ("ms_cg", C.c_double), ("admm_bytes", C.c_double),
        ("p3_flops", C.c_double),
        ("omega_route", C.c_int),
        ("omega_x0_resid", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/spatpca_b200.h one to one
SIGNATURES = {
    "spb_abi_version": (C.c_int, []),
    "spb_last_error": (C.c_char_p, []),
    "spb_device_count": (C.c_int, []),
    "spb_set_device": (C.c_int, [C.c_int]),
    "spb_trim_memory": (C.c_int, []),
    "spb_tps_matrix": (C.c_int, [_dp, C.c_int, C.c_int, _dp]),
    "spb_eigen_function": (C.c_int, [_dp, C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "spb_spatial_prediction": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int,
                                         _dp, _dp, _dp, _dp]),
    "spb_cv": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int,
                         _dp, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp,
                         C.POINTER(CvStats)]),
    "spb_cv_session": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int,
                                 _dp, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp,
                                 C.POINTER(CvStats)]),
    "spb_session_clear": (C.c_int, []),
    "spb_engine_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int,
                                    _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, _dp, C.c_int]),
    "spb_engine_destroy": (C.c_int, [_vp]),
    "spb_engine_upload": (C.c_int, [_vp, _dp, _dp, _dp]),
    "spb_engine_set_omega": (C.c_int, [_vp, _dp]),
    "spb_engine_prepare": (C.c_int, [_vp]),
    "spb_engine_get_omega": (C.c_int, [_vp, _dp]),
    "spb_engine_stage1_fold": (C.c_int, [_vp, C.c_int, _dp]),
    "spb_engine_stage1_full": (C.c_int, [_vp, C.c_int]),
    "spb_engine_stage2_fold": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "spb_engine_stage2_full": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spb_engine_stage3_fold": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp]),
    "spb_engine_stage1_folds": (C.c_int, [_vp, _ip, C.c_int, _dp]),
    "spb_engine_stage2_folds": (C.c_int, [_vp, _ip, C.c_int, C.c_int, _dp]),
    "spb_engine_stage3_folds": (C.c_int, [_vp, _ip, C.c_int, C.c_int, C.c_int, _dp]),
    "spb_engine_can_share_inverses": (C.c_int, [_vp]),
    "spb_engine_prepare_fold": (C.c_int, [_vp, C.c_int]),
    "spb_engine_prepare_starts": (C.c_int, [_vp, C.POINTER(C.c_int), C.c_int]),
    "spb_engine_inverse_buffer": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "spb_engine_build_inverse": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spb_engine_mark_inverse": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spb_engine_sync": (C.c_int, [_vp]),
    "spb_engine_omega_begin": (C.c_int, [_vp, _ip]),
    "spb_engine_omega_solve_cols": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spb_engine_omega_product_cols": (C.c_int, [_vp, C.c_int, C.c_int]),
    "spb_engine_omega_refine_cols": (C.c_int, [_vp, C.c_int, C.c_int, _ip]),
    "spb_engine_omega_buffer": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "spb_engine_omega_sync": (C.c_int, [_vp]),
    "spb_engine_omega_finish": (C.c_int, [_vp]),
    "spb_engine_reset_k": (C.c_int, [_vp, C.c_int]),
    "spb_engine_get_eigenfn": (C.c_int, [_vp, _dp]),
    "spb_engine_get_stats": (C.c_int, [_vp, C.POINTER(CvStats)]),
    "spb_engine_call_log": (C.c_int, [_vp, _ip, C.c_int]),
    "spb_engine_release_fold": (C.c_int, [_vp, C.c_int]),
    "spb_select_index": (C.c_int, [_dp, C.c_int, C.c_int, _dp]),
    "spb_inv_sympd": (C.c_int, [_dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp]),
    "spb_admm_core2": (C.c_int, [_dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp, _dp, _ip]),
    "spb_admm_core3": (C.c_int, [_dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                 _dp, _dp, _dp, _dp, _dp, _ip]),
    "spb_admm_system": (C.c_int, [_dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_int, C.c_double, _dp, _dp, _dp, _dp, _dp, _ip]),
    "spb_core2_matrix_free": (C.c_int, [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                        C.c_double, C.c_double, _dp, _dp, _dp, _ip, _ip]),
    "spb_bench_core2_cg": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _ip]),
    "spb_bench_core2_batched": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _ip]),
    "spb_dmma_matvec": (C.c_int, [_dp, C.c_int, _dp, C.c_int, _dp]),
    "spb_bench_dmma_matvec": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp]),
    "spb_core2_batched": (C.c_int, [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int,
                                    C.c_double, C.c_double, _dp, _dp, _dp, _ip, _dp, _ip]),
    "spb_core3_batched": (C.c_int, [_dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, C.c_double, _dp, C.c_int,
                                    C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip]),
    "spb_core2_batched_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "spb_default_blocks": (C.c_int, [C.c_int]),
    "spb_default_path_blocks": (C.c_int, [C.c_int]),
    "spb_block_offsets": (C.c_int, [C.c_int, C.c_int, _ip, C.c_int]),
    "spb_admm_phase_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, C.c_int]),
    "spb_svd_right": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "spb_cv_loss": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "spb_gamma_loss": (C.c_int, [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp]),
    "spb_bench_admm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "spb_bench_admm_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "spb_bench_dgemm": (C.c_int, [C.c_int, C.c_int, _dp]),
    "spb_bench_inverse": (C.c_int, [C.c_int, C.c_int, _dp]),
    "spb_bench_factor": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp]),
}

_lib = None


def load():
    """Load libspatpca_b200.so (built by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(spatpca_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        msg = load().spb_last_error()
        raise SpatpcaError(rc, msg.decode() if msg else "")
    return rc


def fmat(a, ncol=None) -> np.ndarray:
    """Column-major FP64 copy/view, the layout R hands to `.Call`."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1) if ncol is None else a.reshape(-1, ncol)
    return np.asfortranarray(a)


def fvec(a) -> np.ndarray:
    return np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)).ravel())


def ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)
