"""RHS flavour F4 — a CRNN whose hidden input rows come from a Flux MLP of the state (yeast-glycolysis/yeast_glycolysis.jl:128-142,
robertson/rober_crnn_qssa.jl:111-126) — in the oracle: against a numpy transcription of the script's `crnn`, the finite-difference
Jacobian the scripts' stiff steppers use (autodiff=false), and the reference's COMMITTED yeast checkpoint (p[294] = 164 CRNN + 130 MLP
parameters): through p2vec + the F4 RHS it must follow the generating glycolysis oscillator at the loss level of its own history."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from crnn_b200 import _abi, cases
from crnn_b200.model import CRNNModel
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
    with pytest.raises(RuntimeError):      # predict path only
        oracle.loss_grad_batch(m, cases.yeast_opts(alg=0), np.zeros((m.n_w, 1)), u0, data, yscale)
