"""GPU parity of the MLP-augmented RHS (F4: yeast_glycolysis.jl:128-142, rober_crnn_qssa.jl:111-126) against the CPU oracle, through
the C-ABI: the predict path (k_wide_solve<..., MLP = true>: the MLP evaluated lane-per-neuron, finite-difference Jacobian for the stiff
steppers as the scripts' autodiff=false) and the gradient path (k_tsit5_adjoint<..., MLP = true>: the adjoint RHS goes back through
the chain, all CRNN + MLP parameters in one backward pass)."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases
from crnn_b200.engine import EngineError
from oracle import oracle
from crnn_b200.model import SolveOpts
from test_f4_mlp_cpu import model_from_flat, qssa_like_model, yeast_u0

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


def _yeast_training_problem(golden, N, n_save=60, seed=4):
    from scipy.integrate import solve_ivp
    p = np.array(golden["yeast"]["p"])
    u0 = yeast_u0(N, seed=seed)
    ts = np.linspace(0.0, 5.0, n_save)
    data = np.array([solve_ivp(cases.yeast_true_rhs, (0, 5), u, method="Radau", rtol=1e-9, atol=1e-12, t_eval=ts).y.T for u in u0])
    return p, u0, data, data.std(axis=1).max(axis=0) + 1e-5


@pytest.mark.parametrize("mode", [_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT])
def test_f4_gradient_by_the_adjoints(engine, golden, mode):
    """loss + gradient of all 294 parameters of the yeast model (164 CRNN + 130 MLP) from the adjoint kernels: the adjoint RHS goes back
    through the Flux chain, the quadrature runs in the extended weight space [w_in; w_b; w_out; w_J; mlp]; against the oracle
    (whose gradient tests/test_f4_mlp_cpu.py checks against finite differences)"""
    p, u0, data, ys = _yeast_training_problem(golden, 48)
    m, seed = cases.yeast_model(p), cases.yeast_seed(p)
    assert seed.shape == (m.n_w, 294) and m.n_w == 377
    o = cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=60, sens_mode=mode)
    got = engine.loss_grad_batch(m, o, seed, u0, data, ys, _abi.LOSS_MAE_SCALED, want_pred=True)
    ref = oracle.loss_grad_batch(m, o, seed, u0, data, ys, _abi.LOSS_MAE_SCALED, want_pred=True, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    for k in ("n_accept", "n_reject"):                     # the forward pass: the oracle's steps
        assert np.array_equal(got["stats"][k], ref["stats"][k])
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-6)
    rel = np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"])
    assert rel < (1e-6 if mode == _abi.SENS_DISCRETE_ADJOINT else 1e-3), rel
    assert np.abs(got["grad_sum"][164:]).max() > 0 and np.abs(got["grad_sum"][156:163]).max() > 0   # MLP and w_J entries are live
    if mode == _abi.SENS_DISCRETE_ADJOINT:
        # an oracle-free check: central differences of the GPU's own loss at reltol 1e-6 (~400 steps, inside the kernel's record
        # capacity).  The discrete adjoint holds the step sizes fixed, the differences do not: agreement to ~tol/h, bar 1e-2
        tol = dict(n_save=60, abstol=1e-9, reltol=1e-6, pred_clamp=(-np.inf, np.inf))
        ov, og = cases.yeast_opts(alg=_abi.ALG_TSIT5, **tol), cases.yeast_opts(alg=_abi.ALG_TSIT5, sens_mode=mode, **tol)
        def L(q):
            pr = engine.solve_batch(cases.yeast_model(q), ov, u0[:4])["pred"]
            return float(np.sum(np.mean(np.abs(data[:4] / ys - pr / ys), axis=(1, 2))))
        r4 = engine.loss_grad_batch(m, og, seed, u0[:4], data[:4], ys, _abi.LOSS_MAE_SCALED)
        assert (r4["retcode"] == _abi.RET_SUCCESS).all()
        ks = [1, 40, 100, 158, 162, 163, 170, 260]
        fd = np.array([(L(p + 1e-6 * np.eye(294)[k]) - L(p - 1e-6 * np.eye(294)[k])) / 2e-6 for k in ks])
        assert np.abs(r4["grad_sum"][ks] - fd).max() < 1e-2 * np.abs(fd).max(), (r4["grad_sum"][ks], fd)


@pytest.mark.parametrize("mode", [_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT])
def test_f4_gradient_of_the_qssa_shape(engine, mode):
    """rober_crnn_qssa.jl's shape: exp output, two of the three states into the MLP, the MLP-fed input row BETWEEN state rows, no w_J;
    identity seed = the gradient with respect to every weight of the extended space"""
    q = qssa_like_model()
    w = q.flat_weights()
    u0 = 0.2 + np.random.default_rng(3).random((64, 3))
    # the script's loss looks at rows [1, 3] only (`loss_neuralode`, :150-155: mae over pred[[1, 3], :]): obs_idx = [0, 2]
    o = SolveOpts(saveat=np.linspace(0.0, 2.0, 21), t0=0.0, t1=2.0, alg=_abi.ALG_TSIT5, abstol=1e-8, reltol=1e-6, maxiters=100000, sens_mode=mode,
                  obs_idx=np.array([0, 2]))
    data = oracle.solve_batch(model_from_flat(q, w * (1.0 + 0.05 * np.random.default_rng(5).normal(size=w.size))), o, u0, n_threads=8)["pred"]
    assert data.shape == (64, 21, 2)
    got = engine.loss_grad_batch(q, o, np.eye(w.size), u0, data, np.ones(2), _abi.LOSS_MAE_SCALED)
    ref = oracle.loss_grad_batch(q, o, np.eye(w.size), u0, data, np.ones(2), _abi.LOSS_MAE_SCALED, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    for k in ("n_accept", "n_reject"):
        assert np.array_equal(got["stats"][k], ref["stats"][k])
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-9)
    tol = 1e-7 if mode == _abi.SENS_DISCRETE_ADJOINT else 1e-3
    assert np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) < tol * np.linalg.norm(ref["grad_sum"])
    assert np.abs(got["grad_sum"] - ref["grad_sum"]).max() < tol * np.abs(ref["grad_sum"]).max()
    assert np.abs(got["grad_sum"][42:45]).max() > 0 and np.abs(got["grad_sum"][45:]).max() > 0    # w_J and MLP entries are live
    with pytest.raises(EngineError):        # the adjoints carry the MAE losses
        engine.loss_grad_batch(q, o, np.eye(w.size), u0, data, np.ones(2), _abi.LOSS_MSE)


def test_f4_adjoint_record_capacity_is_reported(engine, golden):
    """the adjoint kernels keep the forward steps of a trajectory in shared memory plus 512 overflow slots: a solve that needs more
    (here: tolerances of 1e-10) comes back MaxIters for that trajectory — never a silently truncated gradient"""
    p, u0, data, ys = _yeast_training_problem(golden, 4)
    m, seed = cases.yeast_model(p), cases.yeast_seed(p)
    o = cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=60, sens_mode=_abi.SENS_DISCRETE_ADJOINT, abstol=1e-12, reltol=1e-10)
    got = engine.loss_grad_batch(m, o, seed, u0, data, ys, _abi.LOSS_MAE_SCALED)
    assert (got["retcode"] == _abi.RET_MAXITERS).all()


def test_f4_forward_mode_and_kencarp4_are_refused_loudly(engine, golden):
    m = cases.yeast_model(np.array(golden["yeast"]["p"]))
    u0 = yeast_u0(4)
    with pytest.raises(EngineError):     # forward sensitivities of F4 are not built: the adjoint sens_modes serve its gradients
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


def test_f4_gradient_edge_cases(engine, golden):
    """the F4 gradient path with an empty batch, one trajectory, the script's random truncation (`batch = rand(batch_min:ntotal)`,
    yeast_glycolysis.jl:245 -> n_save_used), device-resident buffers (same bits) and the dataset-indexed call of the training loop"""
    import torch
    p, u0, data, ys = _yeast_training_problem(golden, 24, seed=12)
    m, seed = cases.yeast_model(p), cases.yeast_seed(p)
    o = cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=60, sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    r0 = engine.loss_grad_batch(m, o, seed, u0[:0], data[:0], ys)
    assert r0["loss"].shape == (0,) and r0["grad_sum"].shape == (294,) and not r0["grad_sum"].any()
    g1 = engine.loss_grad_batch(m, o, seed, u0[:1], data[:1], ys); r1 = oracle.loss_grad_batch(m, o, seed, u0[:1], data[:1], ys)
    assert np.linalg.norm(g1["grad_sum"] - r1["grad_sum"]) < 1e-6 * np.linalg.norm(r1["grad_sum"])
    nsu = np.random.default_rng(2).integers(32, 61, size=24).astype(np.int32)
    g = engine.loss_grad_batch(m, o, seed, u0, data, ys, n_save_used=nsu)
    r = oracle.loss_grad_batch(m, o, seed, u0, data, ys, n_save_used=nsu, n_threads=8)
    assert np.array_equal(g["n_saved"], nsu) and np.array_equal(g["stats"]["n_accept"], r["stats"]["n_accept"])
    np.testing.assert_allclose(g["loss"], r["loss"], rtol=1e-6)
    assert np.linalg.norm(g["grad_sum"] - r["grad_sum"]) < 1e-6 * np.linalg.norm(r["grad_sum"])
    gd = engine.loss_grad_batch(m, o, seed, torch.from_numpy(u0).cuda(), torch.from_numpy(data).cuda(), ys, n_save_used=torch.from_numpy(nsu).cuda())
    assert np.array_equal(gd["grad_sum"].cpu().numpy(), g["grad_sum"]) and np.array_equal(gd["loss"].cpu().numpy(), g["loss"])
    ds = engine.dataset(u0, data)
    idx = np.array([5, 3, 3, 17, 0])
    gi = engine.loss_grad_indexed(m, o, seed, ds, ys, _abi.LOSS_MAE_SCALED, idx=idx, n_save_used=nsu[idx], want_loss=True)
    gb = engine.loss_grad_batch(m, o, seed, u0[idx], data[idx], ys, n_save_used=nsu[idx])
    np.testing.assert_allclose(gi["grad_sum"], gb["grad_sum"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gi["loss"], gb["loss"], rtol=1e-13)
    ds.close()
