"""RHS flavour F4 — a CRNN whose hidden input rows come from a Flux MLP of the state (yeast-glycolysis/yeast_glycolysis.jl:128-142,
robertson/rober_crnn_qssa.jl:111-126) — in the oracle: against a numpy transcription of the script's `crnn`, the finite-difference
Jacobian the scripts' stiff steppers use (autodiff=false), and the reference's COMMITTED yeast checkpoint (p[294] = 164 CRNN + 130 MLP
parameters): through p2vec + the F4 RHS it must follow the generating glycolysis oscillator at the loss level of its own history."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from crnn_b200 import _abi, cases
from crnn_b200.model import CRNNModel, SolveOpts
from oracle import oracle


def yeast_literal(p):
    w_in, w_b, w_out, w_J, pnn = cases.p2vec_yeast(p)

    def crnn(u):   # yeast_glycolysis.jl:128-132
        u_ = np.concatenate([u, cases.mlp_reference(cases.YEAST_MLP_DIMS, pnn, u)])
        w_in_x = w_in.T @ np.log(np.clip(u_, 1e-5, 100.0))
        return (w_out @ np.exp(w_in_x + w_b))[:7] + w_J
    return crnn


def yeast_u0(N, seed=0):
    g = np.random.default_rng(seed)
    return cases.YEAST_IC_LB + g.random((N, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)   # yeast_glycolysis.jl:69-73


def qssa_like_model(seed=0):
    """the SHAPE of rober_crnn_qssa.jl:111-126: u_ = [u1; mlp(u1, u3); u3], Chain(Dense(2,4,gelu), Dense(4,4,gelu) x 2, Dense(4,1,exp))"""
    g = np.random.default_rng(seed)
    dims = (2, 4, 4, 4, 1)
    n_par = sum(dims[l] * dims[l + 1] + dims[l + 1] for l in range(4))
    w_in = np.clip(g.normal(0.5, 0.6, (3, 6)), 0.0, 2.5)
    w_out = -w_in * 10.0 ** g.normal(-0.5, 0.3, (3, 6)) + np.abs(g.normal(0.0, 0.2, (3, 6))) * (w_in == 0.0)
    return CRNNModel(w_in=w_in, w_b=g.normal(-1.0, 1.0, 6), w_out=w_out, rhs_kind=_abi.RHS_F4, lb=1e-8, ub=np.inf,
                     mlp_dims=np.array(dims), mlp_in_idx=np.array([0, 2]), mlp_params=g.normal(0.0, 0.5, n_par) - 0.2,
                     mlp_act_out=1, aug_src=np.array([0, -1, 2]))


def test_f4_rhs_is_the_scripts_formula_and_its_fd_jacobian(golden):
    p = np.array(golden["yeast"]["p"])
    assert p.size == 294
    m = cases.yeast_model(p)
    assert m.n_state == 7 and m.n_in == 12 and m.n_reac == 12 and m.mlp_params.size == 130
    lit = yeast_literal(p)
    for u in yeast_u0(8, seed=1):
        f, J, dT = oracle.rhs_t(m, 0.3, u)
        np.testing.assert_allclose(f, lit(u), rtol=1e-13, atol=1e-13)
        Jc = np.array([(lit(u + 1e-6 * e) - lit(u - 1e-6 * e)) / 2e-6 for e in np.eye(7)]).T
        assert np.abs(J - Jc).max() < 1e-5 * np.abs(Jc).max()        # forward differences at sqrt(eps)
        assert np.all(dT == 0.0)
    q = qssa_like_model()
    mlp = lambda u: cases.mlp_reference((2, 4, 4, 4, 1), q.mlp_params, u[[0, 2]], act_out=1)
    for u in 0.2 + np.random.default_rng(2).random((5, 3)):
        u_ = np.array([u[0], mlp(u)[0], u[2]])
        want = q.w_out @ np.exp(q.w_in.T @ np.log(np.clip(u_, 1e-8, np.inf)) + q.w_b)
        np.testing.assert_allclose(oracle.rhs_t(q, 0.0, u)[0], want, rtol=1e-13)


def test_committed_yeast_checkpoint_follows_the_glycolysis_oscillator(golden):
    """the reference's own trained model (checkpoint/mymodel.bson: loss history 1.87 -> min 0.12, last 0.27-0.28 on 0.1 %-noisy data)
    against fresh trajectories of trueODEfunc (:47-66) from the script's initial-condition box: same loss level"""
    p = np.array(golden["yeast"]["p"])
    m = cases.yeast_model(p)
    N = 12
    u0 = yeast_u0(N)
    ts = np.linspace(0.0, 5.0, 300)
    data = np.array([solve_ivp(cases.yeast_true_rhs, (0, 5), u, method="Radau", rtol=1e-9, atol=1e-12, t_eval=ts).y.T for u in u0])
    yscale = data.std(axis=1).max(axis=0) + 1e-5            # y_scale = maximum(std(ode_data, dims=2)) .+ lb (:98,101)
    losses = {}
    for alg in (_abi.ALG_TSIT5, _abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_TRBDF2, _abi.ALG_ROSENBROCK23):
        r = oracle.solve_batch(m, cases.yeast_opts(alg=alg), u0, n_threads=4)
        assert (r["retcode"] == _abi.RET_SUCCESS).all() and (r["n_saved"] == 300).all()
        losses[alg] = float(np.mean(np.abs(np.clip(data, 1e-5, 100.0) / yscale - r["pred"] / yscale)))   # loss_neuralode (:159-164)
    hist = golden["yeast"]["l_loss_val"]
    assert 0.5 * hist["min"] < losses[_abi.ALG_AUTO_TSIT5_TRBDF2] < 2.0 * hist["last"], (losses, hist)
    # the script's algorithm: the composite never leaves Tsit5 on this model; the stiff steppers alone agree to tolerance
    assert losses[_abi.ALG_AUTO_TSIT5_TRBDF2] == losses[_abi.ALG_TSIT5]
    for alg in (_abi.ALG_TRBDF2, _abi.ALG_ROSENBROCK23):
        assert abs(losses[alg] - losses[_abi.ALG_TSIT5]) < 0.02
    # a tight Tsit5 solve of the literal numpy RHS by scipy = the oracle's tight solve
    lit = yeast_literal(p)
    sol = solve_ivp(lambda t, y: lit(y), (0, 5), u0[0], method="DOP853", rtol=1e-10, atol=1e-12, t_eval=ts)
    tight = oracle.solve_batch(m, cases.yeast_opts(alg=_abi.ALG_TSIT5, abstol=1e-11, reltol=1e-9, pred_clamp=(-np.inf, np.inf)), u0[:1])
    assert np.abs(tight["pred"][0] - sol.y.T).max() < 1e-6
    with pytest.raises(RuntimeError):      # gradients of F4: the adjoint sens_modes only
        oracle.loss_grad_batch(m, cases.yeast_opts(alg=0), np.zeros((m.n_w, 1)), u0, data, yscale)


def model_from_flat(m, w):
    """inverse of CRNNModel.flat_weights for an F4 model without an observable row"""
    import dataclasses
    nin, nr, ns = m.n_in, m.n_reac, m.n_species
    o = np.cumsum([0, nin * nr, nr, ns * nr, ns, m.mlp_params.size])
    return dataclasses.replace(m, w_in=w[o[0]:o[1]].reshape(nin, nr, order="F"), w_b=w[o[1]:o[2]], w_out=w[o[2]:o[3]].reshape(ns, nr, order="F"),
                               w_J=w[o[3]:o[4]], mlp_params=w[o[4]:o[5]])


def test_f4_adjoint_gradients_in_the_oracle(golden):
    """the oracle's adjoint RHS for F4 goes back through the Flux chain (reverse mode: gelu' / softplus' / exp') and integrates the
    gradient in the extended weight space [vec(w_in); w_b; vec(w_out); w_J; mlp]: (a) cases.yeast_seed = d flat_weights / dp,
    (b) discrete and interpolating adjoint agree at tight tolerance, (c) both are the central differences of the loss,
    for the yeast checkpoint (softplus output, all states into the MLP, w_J) and for the QSSA shape (exp output, a state subset
    into the MLP, an augmented row BETWEEN state rows, no w_J) with an identity seed"""
    p = np.array(golden["yeast"]["p"])
    m, seed = cases.yeast_model(p), cases.yeast_seed(p)
    assert m.n_w == 377 and seed.shape == (377, 294)
    for k in (0, 5, 20, 100, 157, 163, 170, 293):
        e = 1e-6 * np.eye(294)[k]
        fd = (cases.yeast_model(p + e).flat_weights() - cases.yeast_model(p - e).flat_weights()) / 2e-6
        assert np.abs(seed[:, k] - fd).max() < 1e-8 * max(1.0, np.abs(fd).max())
    tight = dict(n_save=40, abstol=1e-12, reltol=1e-10, pred_clamp=(-np.inf, np.inf))
    u0 = yeast_u0(2)
    ts = np.linspace(0.0, 5.0, 40)
    data = np.array([solve_ivp(cases.yeast_true_rhs, (0, 5), u, method="Radau", rtol=1e-9, atol=1e-12, t_eval=ts).y.T for u in u0])
    ys = data.std(axis=1).max(axis=0) + 1e-5
    g = {mode: oracle.loss_grad_batch(m, cases.yeast_opts(alg=0, sens_mode=mode, **tight), seed, u0, data, ys, _abi.LOSS_MAE_SCALED, n_threads=2)
         for mode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT)}
    gd, gi = g[_abi.SENS_DISCRETE_ADJOINT]["grad_sum"], g[_abi.SENS_INTERP_ADJOINT]["grad_sum"]
    assert np.linalg.norm(gd - gi) < 1e-7 * np.linalg.norm(gd)
    ov = cases.yeast_opts(alg=0, **tight)
    def L(q):
        pr = oracle.solve_batch(cases.yeast_model(q), ov, u0, n_threads=2)["pred"]
        return np.sum(np.mean(np.abs(data / ys - pr / ys), axis=(1, 2)))
    ks = [0, 15, 60, 130, 157, 160, 163, 165, 200, 293]
    fd = np.array([(L(p + 1e-6 * np.eye(294)[k]) - L(p - 1e-6 * np.eye(294)[k])) / 2e-6 for k in ks])
    assert np.abs(gd[ks] - fd).max() < 1e-5 * np.abs(fd).max(), (gd[ks], fd)

    q = qssa_like_model()
    w = q.flat_weights()
    assert w.size == q.n_w == 6 * 7 + 3 + 57
    u0 = 0.2 + np.random.default_rng(3).random((2, 3))
    so = lambda **kw: SolveOpts(saveat=np.linspace(0.0, 2.0, 21), t0=0.0, t1=2.0, alg=0, abstol=1e-12, reltol=1e-10, maxiters=100000, **kw)
    data = oracle.solve_batch(model_from_flat(q, w * (1.0 + 0.05 * np.random.default_rng(5).normal(size=w.size))), so(), u0)["pred"]
    gq = {mode: oracle.loss_grad_batch(q, so(sens_mode=mode), np.eye(w.size), u0, data, np.ones(3), _abi.LOSS_MAE_SCALED)["grad_sum"]
          for mode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT)}
    assert np.linalg.norm(gq[_abi.SENS_DISCRETE_ADJOINT] - gq[_abi.SENS_INTERP_ADJOINT]) < 1e-4 * np.linalg.norm(gq[_abi.SENS_DISCRETE_ADJOINT])
    def Lq(v):
        pr = oracle.solve_batch(model_from_flat(q, v), so(), u0)["pred"]
        return np.sum(np.mean(np.abs(data - pr), axis=(1, 2)))
    ks = [0, 4, 17, 20, 30, 41, 43, 46, 60, 80, 101]    # w_in (incl. the MLP-fed row 1), w_b, w_out, w_J, MLP weights and biases
    fd = np.array([(Lq(w + 1e-6 * np.eye(w.size)[k]) - Lq(w - 1e-6 * np.eye(w.size)[k])) / 2e-6 for k in ks])
    gd = gq[_abi.SENS_INTERP_ADJOINT]
    assert np.abs(gd[ks] - fd).max() < 1e-5 * np.abs(fd).max(), (gd[ks], fd)
    with pytest.raises(RuntimeError):      # the adjoints carry the MAE losses
        oracle.loss_grad_batch(q, so(sens_mode=_abi.SENS_INTERP_ADJOINT), np.eye(w.size), u0, data, np.ones(3), _abi.LOSS_MSE)
    assert np.abs(gd[42:45]).max() > 0 and np.abs(gd[45:]).max() > 0


def test_mlp_rows_postmap_is_the_qssa_scripts_line(golden):
    """rober_crnn_qssa.jl:139 `pred[2:2, :] .= rep(pred[[1, 3], :])` and the yeast script's hidden-species read-out, host side"""
    q = qssa_like_model()
    pred = 0.2 + np.random.default_rng(0).random((3, 5, 3))
    out, mapped = cases.mlp_rows_postmap(q, pred)
    assert out.shape == (3, 5, 1) and np.array_equal(mapped[..., [0, 2]], pred[..., [0, 2]])
    for n in range(3):
        for k in range(5):
            want = cases.mlp_reference((2, 4, 4, 4, 1), q.mlp_params, pred[n, k, [0, 2]], act_out=1)
            assert mapped[n, k, 1] == want[0] == out[n, k, 0]
    m = cases.yeast_model(np.array(golden["yeast"]["p"]))
    pr = 0.5 + np.random.default_rng(1).random((2, 4, 7))
    hidden, same = cases.mlp_rows_postmap(m, pr)
    assert hidden.shape == (2, 4, 5) and np.array_equal(same, pr)       # yeast: the MLP rows are extra inputs, no state row is replaced
    with pytest.raises(ValueError):
        cases.mlp_rows_postmap(q, pred[..., :2], obs_idx=[0, 1])          # the chain reads rows 0 and 2
