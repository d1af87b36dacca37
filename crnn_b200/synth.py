"""Synthetic inputs shaped like the reference scripts' data generation.

Initial conditions follow the scripts' distributions (case2/case2.jl:62-65,
robertson/rober_crnn.jl:44-46, case1/case1.jl:47-49, case3/case3.jl:106) drawn
from a counter-based Philox stream so every harness (CPU oracle, GPU engine, any
rank of a multi-GPU run) regenerates identical fp64 values for a trajectory
index without communication.  Targets are made by the caller: solve the
generating mechanism (cases.true_model_*) and apply `noisy_targets`.
"""
from __future__ import annotations

import numpy as np


BLOCK = 1024  # trajectories per independent Philox counter block


def _block_rng(seed: int, stream: int, block: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, stream, block]))


def _blocked(seed, stream, start, N, shape_tail, draw):
    """Rows start..start+N-1 of a virtual infinite array generated BLOCK rows at a time."""
    out = np.empty((N,) + tuple(shape_tail))
    b0, b1 = start // BLOCK, (start + N - 1) // BLOCK if N > 0 else start // BLOCK - 1
    for b in range(b0, b1 + 1):
        blk = draw(_block_rng(seed, stream, b), (BLOCK,) + tuple(shape_tail))
        lo, hi = max(start, b * BLOCK), min(start + N, (b + 1) * BLOCK)
        out[lo - start:hi - start] = blk[lo - b * BLOCK:hi - b * BLOCK]
    return out


def make_u0(case: str, N: int, seed: int = 1234, start: int = 0) -> np.ndarray:
    """[N, n_state] initial conditions for trajectories start..start+N-1 (rank-shardable)."""
    ncols = {"case1": 2, "case2": 3, "case3": 9, "robertson": 2}[case]
    r = _blocked(seed, 0, start, N, (ncols,), lambda g, shp: g.random(shp))
    if case == "case1":
        u0 = np.zeros((N, 5)); u0[:, 0:2] = r + 0.2
    elif case == "case2":
        u0 = np.zeros((N, 7)); u0[:, 0:2] = r[:, 0:2] * 2.0 + 0.2; u0[:, 6] = r[:, 2] * 20.0 + 323.0
    elif case == "case3":
        u0 = 10.0 ** (-3.0 * r)
    elif case == "robertson":
        u0 = np.empty((N, 3)); u0[:, 0] = r[:, 0] + 0.5; u0[:, 1] = 1e-8; u0[:, 2] = r[:, 1] + 0.5
    else:
        raise KeyError(case)
    return u0


def noisy_targets(pred: np.ndarray, noise: float, seed: int = 1234, start: int = 0) -> np.ndarray:
    """ode_data += randn .* ode_data .* noise (case2/case2.jl:79); pred [N, n_save, n_obs]."""
    z = _blocked(seed, 1, start, pred.shape[0], pred.shape[1:], lambda g, shp: g.standard_normal(shp))
    return pred * (1.0 + noise * z)


def yscale_from(data: np.ndarray, lb: float = 0.0) -> np.ndarray:
    """max over experiments of (max_t - min_t + lb) per species (case2/case2.jl:70-72,83)."""
    return (data.max(axis=1) - data.min(axis=1) + lb).max(axis=0)
