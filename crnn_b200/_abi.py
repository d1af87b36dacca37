"""ctypes mirror of include/crnn_b200.h and the loader of libcrnn_b200.so.

The structs here are the boundary definition (they restate the public header),
not an implementation.  The product path has NO CPU fallback: `load_library`
raises if the CUDA shared library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libcrnn_b200.so")

# enums (include/crnn_b200.h)
RHS_F0, RHS_F1, RHS_F2, RHS_F5, RHS_F4 = 0, 1, 2, 3, 4
ALG_TSIT5, ALG_ROSENBROCK23, ALG_KENCARP4, ALG_AUTO_TSIT5_ROS23, ALG_TRBDF2, ALG_AUTO_TSIT5_TRBDF2 = 0, 1, 2, 3, 4, 5
SENS_NONE, SENS_FORWARD, SENS_INTERP_ADJOINT, SENS_DISCRETE_ADJOINT = 0, 1, 2, 3
LOSS_MAE_SCALED, LOSS_MAE_LOG, LOSS_MSE = 0, 1, 2
RET_DEFAULT, RET_SUCCESS, RET_DTNAN, RET_MAXITERS, RET_DTLESSTHANMIN, RET_UNSTABLE = 0, 1, 3, 4, 5, 6
ERR_BAD_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class CModel(C.Structure):
    _fields_ = [
        ("n_state", C.c_int32), ("n_species", C.c_int32), ("n_in", C.c_int32), ("n_reac", C.c_int32),
        ("rhs_kind", C.c_int32), ("n_tab", C.c_int32),
        ("lb", C.c_double), ("ub", C.c_double), ("gas_R", C.c_double),
        ("out_scale", c_double_p), ("w_in", c_double_p), ("w_b", c_double_p), ("w_out", c_double_p),
        ("mw", c_double_p), ("tab_t", c_double_p), ("tab_T", c_double_p), ("tab_P", c_double_p),
        ("w_obs", c_double_p),
        ("mlp_n_layers", C.c_int32), ("mlp_act_out", C.c_int32),
        ("mlp_dims", c_int32_p), ("mlp_in_idx", c_int32_p), ("mlp_params", c_double_p), ("aug_src", c_int32_p), ("w_J", c_double_p),
    ]


class COpts(C.Structure):
    _fields_ = [
        ("alg", C.c_int32), ("sens_mode", C.c_int32), ("err_norm_includes_sens", C.c_int32),
        ("n_save", C.c_int32), ("n_obs", C.c_int32), ("n_abstol", C.c_int32), ("n_reltol", C.c_int32),
        ("buffers_on_device", C.c_int32),
        ("maxiters", C.c_int64),
        ("t0", C.c_double), ("t1", C.c_double),
        ("pred_clamp_lo", C.c_double), ("pred_clamp_hi", C.c_double),
        ("abstol", c_double_p), ("reltol", c_double_p), ("saveat", c_double_p), ("obs_idx", c_int32_p),
        ("qmin", C.c_double), ("qmax", C.c_double), ("gamma", C.c_double),
        ("beta1", C.c_double), ("beta2", C.c_double),
        ("stream", C.c_void_p),
        ("qsteady_min", C.c_double), ("qsteady_max", C.c_double),
        ("err_norm_mean_over_state_only", C.c_int32), ("reserved0", C.c_int32),
    ]


class CTrainOpts(C.Structure):
    _fields_ = [
        ("p2vec_kind", C.c_int32), ("optimiser", C.c_int32), ("batch", C.c_int32), ("reserved", C.c_int32),
        ("eta", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
        ("expdecay_eta", C.c_double), ("expdecay_decay", C.c_double), ("expdecay_clip", C.c_double),
        ("expdecay_step", C.c_int64), ("grad_max", C.c_double), ("p2vec_b0", C.c_double),
        ("n_save_used", C.c_void_p),
    ]


class CStats(C.Structure):
    _fields_ = [
        ("n_accept", C.c_int32), ("n_reject", C.c_int32), ("n_rhs", C.c_int32), ("n_jac", C.c_int32),
        ("t_reached", C.c_double), ("dt_last", C.c_double),
    ]


STATS_DTYPE = np.dtype(
    [("n_accept", "<i4"), ("n_reject", "<i4"), ("n_rhs", "<i4"), ("n_jac", "<i4"),
     ("t_reached", "<f8"), ("dt_last", "<f8")]
)
assert STATS_DTYPE.itemsize == C.sizeof(CStats) == 32

# every symbol include/crnn_b200.h declares
EXPORTS = (
    "crnn_create", "crnn_destroy", "crnn_last_error", "crnn_version", "crnn_launch_count",
    "crnn_solve_batch", "crnn_loss_grad_batch", "crnn_profile_begin", "crnn_profile_end", "crnn_copy_grad_each",
    "crnn_debug_lean_math", "crnn_create_multi", "crnn_device_count", "crnn_dataset_create", "crnn_dataset_destroy",
    "crnn_dataset_size", "crnn_loss_grad_indexed", "crnn_loss_grad_particles", "crnn_train_steps",
)

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libcrnn_b200.so (built by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(crnn_b200 has no CPU fallback)")
    lib = C.CDLL(p)
    lib.crnn_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.crnn_create.restype = C.c_int
    lib.crnn_destroy.argtypes = [C.c_void_p]
    lib.crnn_destroy.restype = None
    lib.crnn_last_error.argtypes = [C.c_void_p]
    lib.crnn_last_error.restype = C.c_char_p
    lib.crnn_version.argtypes = []
    lib.crnn_version.restype = C.c_int
    lib.crnn_launch_count.argtypes = [C.c_void_p]
    lib.crnn_launch_count.restype = C.c_int64
    lib.crnn_profile_begin.argtypes = [C.c_void_p]
    lib.crnn_profile_begin.restype = C.c_int
    lib.crnn_profile_end.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.crnn_profile_end.restype = C.c_int
    lib.crnn_copy_grad_each.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.crnn_copy_grad_each.restype = C.c_int
    lib.crnn_debug_lean_math.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.crnn_debug_lean_math.restype = C.c_int
    lib.crnn_create_multi.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int32]
    lib.crnn_create_multi.restype = C.c_int
    lib.crnn_device_count.argtypes = [C.c_void_p]
    lib.crnn_device_count.restype = C.c_int32
    lib.crnn_dataset_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                        C.POINTER(C.c_void_p)]
    lib.crnn_dataset_create.restype = C.c_int
    lib.crnn_dataset_destroy.argtypes = [C.c_void_p]
    lib.crnn_dataset_destroy.restype = None
    lib.crnn_dataset_size.argtypes = [C.c_void_p]
    lib.crnn_dataset_size.restype = C.c_int64
    lib.crnn_loss_grad_indexed.argtypes = [
        C.c_void_p, C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.crnn_loss_grad_indexed.restype = C.c_int
    lib.crnn_loss_grad_particles.argtypes = [
        C.c_void_p, C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p]
    lib.crnn_loss_grad_particles.restype = C.c_int
    lib.crnn_train_steps.argtypes = [
        C.c_void_p, C.POINTER(CModel), C.POINTER(COpts), C.POINTER(CTrainOpts), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.crnn_train_steps.restype = C.c_int
    lib.crnn_solve_batch.argtypes = [
        C.c_void_p, C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_int64, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.crnn_solve_batch.restype = C.c_int
    lib.crnn_loss_grad_batch.argtypes = [
        C.c_void_p, C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_int32,
        C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.crnn_loss_grad_batch.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def dptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_int32_p)
