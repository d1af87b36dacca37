"""GPU parity of the MLP-augmented RHS (F4: yeast_glycolysis.jl:128-142, rober_crnn_qssa.jl:111-126) on the predict path
(k_wide_solve<..., MLP = true>: the MLP evaluated lane-per-neuron, finite-difference Jacobian for the stiff steppers as the scripts'
autodiff=false) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases
from crnn_b200.engine import EngineError
from oracle import oracle
from test_f4_mlp_cpu import qssa_like_model, yeast_u0

pytestmark = pytest.mark.gpu


def _counts_equal(got, ref):
    for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
        bad = np.nonzero(got["stats"][k] != ref["stats"][k])[0]
        assert bad.size == 0, f"{k} differs from the oracle for trajectories {bad[:8]}"
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])


@pytest.mark.parametrize("alg", [_abi.ALG_TSIT5, _abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_TRBDF2, _abi.ALG_ROSENBROCK23])
def test_yeast_checkpoint_model_matches_the_oracle(engine, golden, alg):
    """the reference's committed yeast model (164 CRNN + 130 MLP parameters), 300 saves on [0, 5], AutoTsit5(TRBDF2) as written (:33)"""
    m = cases.yeast_model(np.array(golden["yeast"]["p"]))
    u0 = yeast_u0(192, seed=3)
    o = cases.yeast_opts(alg=alg)
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    err = np.abs(got["pred"] - ref["pred"]) / scale
    if alg in (_abi.ALG_TSIT5, _abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_AUTO_TSIT5_ROS23):
        _counts_equal(got, ref)
        # a trained oscillator: a few trajectories amplify the last-bit differences of the MLP's summation order over 130+ steps
        assert err.max() < 1e-5 and np.median(err.max(axis=(1, 2))) < 1e-9
        if alg != _abi.ALG_TSIT5:    # most trajectories never leave Tsit5; the ones that do take the oracle's stiff steps too
            assert (got["stats"]["n_jac"] == 0).mean() > 0.5 and (got["stats"]["n_jac"] > 0).any()
    else:
        # ~200 attempts per trajectory with a finite-difference Jacobian (its sqrt(eps) noise feeds Newton / error tests):
        # all but a few trajectories take the oracle's counts; every one solves the ODE to tolerance
        same = np.ones(len(u0), dtype=bool)
        for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
            same &= got["stats"][k] == ref["stats"][k]
        assert same.mean() > 0.9, (~same).sum()
        assert err[same].max() < 1e-6 and err.max() < 2e-2
        assert (got["stats"]["n_jac"] >= got["stats"]["n_accept"]).all()


@pytest.mark.parametrize("alg", [_abi.ALG_TSIT5, _abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_ROSENBROCK23])
def test_qssa_shaped_model(engine, alg):
    """rober_crnn_qssa.jl's shape: the MLP of (u1, u3) REPLACES input row 2, exp output layer, AutoTsit5(Rosenbrock23(autodiff=false))"""
    m = qssa_like_model()
    u0 = 0.2 + np.random.default_rng(5).random((128, 3))
    from crnn_b200.model import SolveOpts
    o = SolveOpts(saveat=np.linspace(0.0, 4.0, 30), t0=0.0, t1=4.0, alg=alg, abstol=1e-7, reltol=1e-4, obs_idx=np.arange(3))
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])
    ok = ref["retcode"] == _abi.RET_SUCCESS
    assert ok.mean() > 0.9
    same = got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]
    assert same[ok].mean() > (0.999 if alg == _abi.ALG_TSIT5 else 0.9)
    scale = np.abs(ref["pred"][ok]).max(axis=(0, 1)) + 1e-300
    assert (np.abs(got["pred"] - ref["pred"])[ok & same] / scale).max() < 1e-6


def test_f4_gradients_and_kencarp4_are_refused_loudly(engine, golden):
    m = cases.yeast_model(np.array(golden["yeast"]["p"]))
    u0 = yeast_u0(4)
    with pytest.raises(EngineError):
        engine.loss_grad_batch(m, cases.yeast_opts(alg=0), np.zeros((m.n_w, 2)), u0, np.zeros((4, 300, 7)), np.ones(7))
    with pytest.raises(EngineError):
        engine.solve_batch(m, cases.yeast_opts(alg=_abi.ALG_KENCARP4), u0)


def test_edge_cases_of_the_new_steppers_and_flavours(engine, golden):
    """empty and single-trajectory calls, per-trajectory truncation (`n_save_used`, rober_crnn.jl:218), `maxiters` and device buffers
    for TRBDF2 / AutoTsit5(TRBDF2) and the F4 flavour: retcodes, n_saved and counts as the oracle's"""
    import torch
    from dataclasses import replace
    from problems import make_problem
    m4 = cases.yeast_model(np.array(golden["yeast"]["p"]))
    pb = make_problem("robertson", golden, 24)
    jobs = [(m4, cases.yeast_opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2, n_save=40), yeast_u0(24, seed=9)),
            (pb["true_model"], pb["case"].opts(alg=_abi.ALG_TRBDF2, pred_clamp=(-np.inf, np.inf)), pb["u0"]),
            (pb["true_model"], pb["case"].opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2, pred_clamp=(-np.inf, np.inf)), pb["u0"])]
    for model, o, u0 in jobs:
        r0 = engine.solve_batch(model, o, u0[:0])
        assert r0["pred"].shape == (0, o.n_save, o.n_obs(model.n_state)) and r0["retcode"].shape == (0,)
        g1 = engine.solve_batch(model, o, u0[:1]); r1 = oracle.solve_batch(model, o, u0[:1])
        assert g1["retcode"][0] == r1["retcode"][0] == _abi.RET_SUCCESS and g1["stats"]["n_rhs"][0] == r1["stats"]["n_rhs"][0]
        nsu = np.random.default_rng(1).integers(1, o.n_save + 1, size=len(u0)).astype(np.int32)
        g = engine.solve_batch(model, o, u0, n_save_used=nsu); r = oracle.solve_batch(model, o, u0, n_save_used=nsu, n_threads=8)
        assert np.array_equal(g["n_saved"], nsu) and np.array_equal(g["retcode"], r["retcode"])
        assert (g["stats"]["n_accept"] == r["stats"]["n_accept"]).mean() > 0.9
        assert (g["pred"][np.arange(len(u0))[:, None] * 0 + np.arange(o.n_save)[None, :] >= nsu[:, None]] == 0.0).all()   # rows beyond n_saved are zero
        om = replace(o, maxiters=7)
        g = engine.solve_batch(model, om, u0); r = oracle.solve_batch(model, om, u0, n_threads=8)
        assert (g["retcode"] == _abi.RET_MAXITERS).all() and np.array_equal(g["retcode"], r["retcode"])
        assert np.array_equal(g["n_saved"], r["n_saved"]) and np.array_equal(g["stats"]["n_rhs"], r["stats"]["n_rhs"])
        # device buffers: same bits as the host-buffer call
        gh = engine.solve_batch(model, o, u0)
        gd = engine.solve_batch(model, o, torch.from_numpy(u0).cuda())
        assert np.array_equal(gd["pred"].cpu().numpy(), gh["pred"]) and np.array_equal(gd["retcode"].cpu().numpy(), gh["retcode"])
