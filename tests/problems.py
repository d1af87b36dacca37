"""Seeded problem set-ups shared by the CPU and GPU tests (inputs only; the
expected values come from the oracle or from the golden fixtures)."""
from __future__ import annotations

import numpy as np

from crnn_b200 import cases, synth
from oracle import oracle

_TRUE = {"case1": cases.true_model_case1, "case2": cases.true_model_case2,
         "case3": cases.true_model_case3, "robertson": cases.true_model_robertson}


def trained_p(name, golden, seed=0):
    """Parameter vector for a case: the reference's committed checkpoint where one exists,
    otherwise the script's own random initialisation (case1.jl:86, case3.jl:35-36)."""
    c = cases.CASES[name]
    if name in ("case2", "robertson"):
        return np.array(golden[name]["p"])
    g = np.random.default_rng(seed)
    if name == "case1":
        return 0.1 * g.standard_normal(c.n_p)
    p = (g.random(c.n_p) - 0.5) * 2 * np.sqrt(6 / (c.ns + c.nr))
    p[-1] = 0.1
    return p


def make_problem(name, golden, N, seed=1234, noise=0.05, obs=None):
    """-> dict(case, model, seed, opts, u0, data, yscale, loss_kind)."""
    c = cases.CASES[name]
    u0 = synth.make_u0(name, N, seed)
    n_state = u0.shape[1]
    obs_idx = np.arange(c.ns) if obs is None else np.asarray(obs)
    # targets from the generating mechanism (always Rosenbrock23-safe tolerances for robertson)
    mt = _TRUE[name]()
    ot = c.opts(obs_idx=obs_idx, pred_clamp=(-np.inf, np.inf))
    truth = oracle.solve_batch(mt, ot, u0, n_threads=8)["pred"]
    data = synth.noisy_targets(truth, noise if name != "robertson" else 1e-4, seed)
    yscale = synth.yscale_from(data, c.lb if name != "robertson" else 0.0)
    out_scale = None
    if name == "robertson":   # dydt_scale = yscale ./ t_end (rober_crnn.jl:81-82)
        out_scale = yscale / c.tspan[1]
    if name == "case3":       # dy_std_ = y_std ./ tspan[2] (case3.jl:141-143)
        out_scale = yscale / c.tspan[1]
    model, seedm = c.model(trained_p(name, golden), out_scale=out_scale)
    opts = c.opts(obs_idx=obs_idx)
    return dict(case=c, model=model, seed=seedm, opts=opts, u0=u0, data=data, yscale=yscale,
                loss_kind=c.loss_kind, true_model=mt)
