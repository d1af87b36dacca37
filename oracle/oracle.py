"""ctypes wrapper of oracle/liboracle_crnn.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It shares the boundary structs of
include/crnn_b200.h (mirrored in crnn_b200/_abi.py) so that the same model /
option objects can be handed to the oracle and to the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from crnn_b200 import _abi
from crnn_b200._abi import CModel, COpts, STATS_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_crnn.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "crnn_oracle.c"), os.path.join(HERE, "lean_math_host.c"),
            os.path.join(HERE, "..", "crnn_b200", "csrc", "lean_math.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle_crnn.so"], check=True, capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.crnn_oracle_solve_batch.restype = C.c_int
        _lib.crnn_oracle_solve_batch.argtypes = [
            C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_int64, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.crnn_oracle_loss_grad_batch.restype = C.c_int
        _lib.crnn_oracle_loss_grad_batch.argtypes = [
            C.POINTER(CModel), C.POINTER(COpts), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.crnn_oracle_rhs.restype = C.c_int
        _lib.crnn_oracle_rhs.argtypes = [C.POINTER(CModel), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.crnn_oracle_rhs_t.restype = C.c_int
        _lib.crnn_oracle_rhs_t.argtypes = [C.POINTER(CModel), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.crnn_oracle_rhs_sens.restype = C.c_int
        _lib.crnn_oracle_rhs_sens.argtypes = [C.POINTER(CModel)] + [C.c_void_p] * 6
        _lib.crnn_oracle_tsit5_tableau.restype = None
        _lib.crnn_oracle_tsit5_tableau.argtypes = [C.c_void_p] * 3
        _lib.crnn_oracle_kencarp4_tableau.restype = None
        _lib.crnn_oracle_kencarp4_tableau.argtypes = [C.c_void_p] * 2
        _lib.crnn_oracle_set_lu_reciprocal.restype = None
        _lib.crnn_oracle_set_lu_reciprocal.argtypes = [C.c_int]
        _lib.crnn_oracle_get_lu_reciprocal.restype = C.c_int
        _lib.crnn_oracle_set_shared_math.restype = None
        _lib.crnn_oracle_set_shared_math.argtypes = [C.c_int]
        _lib.crnn_oracle_get_shared_math.restype = C.c_int
        for fn, na in (("crnn_lean_log", 1), ("crnn_lean_exp", 1), ("crnn_lean_pow", 2), ("crnn_lean_log10", 1), ("crnn_lean_exp10", 1)):
            getattr(_lib, fn).restype = C.c_double
            getattr(_lib, fn).argtypes = [C.c_double] * na
    return _lib


class kc4_inverse:
    """Context manager for the oracle's named KenCarp4 switch: W^{-1} by Gauss-Jordan + mat-vec solves (the CUDA kernel's form)."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        L = lib()
        L.crnn_oracle_get_kc4_inverse.restype = C.c_int
        self.prev = L.crnn_oracle_get_kc4_inverse()
        L.crnn_oracle_set_kc4_inverse(int(self.on))
        return self

    def __exit__(self, *exc):
        lib().crnn_oracle_set_kc4_inverse(self.prev)
        return False


class shared_math:
    """Context manager for the oracle's named math switch: inside it log/exp/pow are the lean functions of
    crnn_b200/csrc/lean_math.h (the kernels' own, bit-identical on host and device); outside, the C library."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        self.prev = lib().crnn_oracle_get_shared_math()
        lib().crnn_oracle_set_shared_math(int(self.on))
        return self

    def __exit__(self, *exc):
        lib().crnn_oracle_set_shared_math(self.prev)
        return False


class lu_reciprocal:
    """Context manager for the oracle's named LU switch: inside it the diagonal of U is stored inverted and the
    triangular solves multiply (the CUDA kernels' form); outside, the literal division form (default)."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        self.prev = lib().crnn_oracle_get_lu_reciprocal()
        lib().crnn_oracle_set_lu_reciprocal(int(self.on))
        return self

    def __exit__(self, *exc):
        lib().crnn_oracle_set_lu_reciprocal(self.prev)
        return False


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u0(model, u0):
    u0 = np.ascontiguousarray(np.asarray(u0, dtype=np.float64))
    if u0.ndim == 1:
        u0 = u0[None, :]
    assert u0.shape[1] == model.n_state, "u0 must be [N, n_state] (trajectory-major rows)"
    return u0


def solve_batch(model, opts, u0, n_save_used=None, n_threads=1):
    """u0 [N, n_state] -> dict(pred [N, n_save, n_obs], n_saved, retcode, stats)."""
    u0 = _u0(model, u0)
    N = u0.shape[0]
    cm, k1 = model.to_c()
    co, k2 = opts.to_c(model.n_state)
    n_obs = co.n_obs
    pred = np.zeros((N, co.n_save, n_obs))
    n_saved = np.zeros(N, dtype=np.int32); ret = np.zeros(N, dtype=np.int32)
    stats = np.zeros(N, dtype=STATS_DTYPE)
    nsu = None if n_save_used is None else np.ascontiguousarray(n_save_used, dtype=np.int32)
    rc = lib().crnn_oracle_solve_batch(C.byref(cm), C.byref(co), _p(u0), N, _p(nsu), _p(pred), _p(n_saved),
                                       _p(ret), _p(stats), int(n_threads))
    if rc:
        raise RuntimeError(f"oracle solve_batch failed: {rc}")
    return dict(pred=pred, n_saved=n_saved, retcode=ret, stats=stats)


def loss_grad_batch(model, opts, seed, u0, data, yscale, loss_kind=_abi.LOSS_MAE_SCALED, n_save_used=None,
                    n_threads=1, want_pred=False, want_grad_each=False):
    """data [N, n_save, n_obs]; seed [n_w, np] -> dict(loss [N], grad_sum [np], ...)."""
    u0 = _u0(model, u0)
    N = u0.shape[0]
    cm, k1 = model.to_c()
    co, k2 = opts.to_c(model.n_state)
    seed = np.asfortranarray(np.asarray(seed, dtype=np.float64))
    n_p = seed.shape[1]
    assert seed.shape[0] == model.n_w
    seed_flat = seed.reshape(-1, order="F").copy()
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
    assert data.shape == (N, co.n_save, co.n_obs)
    ys = np.ascontiguousarray(np.asarray(yscale, dtype=np.float64).reshape(-1))
    loss = np.zeros(N); grad = np.zeros(n_p)
    pred = np.zeros((N, co.n_save, co.n_obs)) if want_pred else None
    geach = np.zeros((N, n_p)) if want_grad_each else None
    n_saved = np.zeros(N, dtype=np.int32); ret = np.zeros(N, dtype=np.int32)
    stats = np.zeros(N, dtype=STATS_DTYPE)
    nsu = None if n_save_used is None else np.ascontiguousarray(n_save_used, dtype=np.int32)
    rc = lib().crnn_oracle_loss_grad_batch(C.byref(cm), C.byref(co), _p(seed_flat), n_p, _p(u0), N, _p(nsu),
                                           _p(data), _p(ys), int(loss_kind), _p(loss), _p(grad), _p(geach),
                                           _p(pred), _p(n_saved), _p(ret), _p(stats), int(n_threads))
    if rc:
        raise RuntimeError(f"oracle loss_grad_batch failed: {rc}")
    return dict(loss=loss, grad_sum=grad, grad_each=geach, pred=pred, n_saved=n_saved, retcode=ret, stats=stats)


def rhs(model, u, want_jac=False):
    cm, k = model.to_c()
    u = np.ascontiguousarray(u, dtype=np.float64)
    du = np.zeros(model.n_state)
    J = np.zeros((model.n_state, model.n_state)) if want_jac else None
    lib().crnn_oracle_rhs(C.byref(cm), _p(u), _p(du), _p(J))
    return (du, J) if want_jac else du


def rhs_t(model, t, u):
    """non-autonomous form: (f(u, t), J = df/du, dT = df/dt)"""
    cm, k = model.to_c()
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = model.n_state
    du = np.zeros(n); J = np.zeros((n, n)); dT = np.zeros(n)
    lib().crnn_oracle_rhs_t(C.byref(cm), float(t), _p(u), _p(du), _p(J), _p(dT))
    return du, J, dT


def rhs_sens(model, u, S, seedcol=None, v=None):
    cm, k = model.to_c()
    u = np.ascontiguousarray(u, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
    sc = None if seedcol is None else np.ascontiguousarray(seedcol, dtype=np.float64)
    vv = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    dS = np.zeros(model.n_state); dJv = np.zeros(model.n_state) if v is not None else None
    lib().crnn_oracle_rhs_sens(C.byref(cm), _p(u), _p(S), _p(sc), _p(vv), _p(dS), _p(dJv))
    return dS, dJv


def rhs_sens_t(model, t, u, S, seedcol=None, v=None, tau=0.0):
    """(f'[(S, seedcol)], D^2 f[(S, seedcol), (v, tau)]) at (u, t): first and mixed second directional derivatives"""
    cm, k = model.to_c()
    u = np.ascontiguousarray(u, dtype=np.float64); S = np.ascontiguousarray(S, dtype=np.float64)
    sc = None if seedcol is None else np.ascontiguousarray(seedcol, dtype=np.float64)
    vv = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
    dS = np.zeros(model.n_state); dJv = np.zeros(model.n_state)
    f = lib().crnn_oracle_rhs_sens_t
    f.restype = C.c_int
    f.argtypes = [C.POINTER(CModel), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    f(C.byref(cm), float(t), _p(u), _p(S), _p(sc), _p(vv), float(tau), _p(dS), _p(dJv))
    return dS, dJv


def tsit5_tableau():
    a = np.zeros((7, 6)); bt = np.zeros(7); r = np.zeros((7, 4))
    lib().crnn_oracle_tsit5_tableau(_p(a), _p(bt), _p(r))
    return a, bt, r


def kencarp4_tableau():
    a = np.zeros((6, 6)); bhat = np.zeros(6)
    lib().crnn_oracle_kencarp4_tableau(_p(a), _p(bhat))
    return a, bhat
