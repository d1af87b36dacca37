"""CPU tests of the oracle's F2 (HyChem mass-fraction) RHS and AutoTsit5(Rosenbrock23) composite, against the
literal reference formulas, finite differences and scipy's Radau — what pins the oracle where the reference
itself cannot run (its HyChem data file is not in its tree, Julia is not installed)."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from crnn_b200 import _abi, cases, synth
from oracle import oracle
from problems import make_problem

YS = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
ALG = {"tsit5": _abi.ALG_TSIT5, "ros23": _abi.ALG_ROSENBROCK23, "auto": _abi.ALG_AUTO_TSIT5_ROS23,
       "kc4": _abi.ALG_KENCARP4}


def crnn_hychem_literal(u, t, m):
    """HyChem/crnn_pyrolysis_mass.jl:107-114,121-131 transcribed line by line in numpy."""
    P = np.interp(t, m.tab_t, m.tab_P)
    T = np.interp(t, m.tab_t, m.tab_T)
    Y = np.clip(u, m.lb, 10.0)
    density = P / (8.31446261815324e3 * T * np.sum(Y / m.mw))
    C = density * (Y / m.mw) * 1e3
    logX = np.log(np.clip(C, m.lb, 10.0))
    w_in_x = m.w_in.T @ np.concatenate([logX, [-1.0 / m.gas_R / T, np.log(T)]])
    wdot = m.w_out @ np.exp(w_in_x + m.w_b)
    return wdot * m.mw / density * m.out_scale


def _state(i=1):
    u = cases.hychem_u0(8)[i].copy()
    u[1:5] = [1e-3, 2e-3, 5e-4, 3e-3]
    u[5] = 1.0 - u[:5].sum() - u[6:].sum()
    return u


def test_p2vec_hychem_shapes_and_seed():
    p = cases.hychem_p(0)
    assert p.size == 211                                   # np = nr*(2ns+3)+1, crnn_pyrolysis_mass.jl:74
    w_in, w_b, w_out, seed = cases.p2vec_hychem(p)
    assert w_in.shape == (11, 10) and w_out.shape == (9, 10) and seed.shape == (10 * (11 + 1 + 9), 211)
    slope = p[-1] * 10.0
    np.testing.assert_allclose(w_b, p[:10] * slope)
    np.testing.assert_allclose(w_in[10], p[10:20])          # log T exponents
    np.testing.assert_allclose(w_in[9], p[20:30] * slope)   # Ea row
    w_in_raw = p[120:210].reshape(10, 9).T
    np.testing.assert_allclose(w_in[:9], np.clip(w_in_raw, 0.0, 2.5))
    np.testing.assert_allclose(w_out, -w_in_raw * 10.0 ** p[30:120].reshape(10, 9).T)
    flat = lambda q: np.concatenate([a.reshape(-1, order="F") for a in cases.p2vec_hychem(q)[:3]])
    for k in (0, 15, 25, 47, 130, 210):
        h = 1e-6
        pp, pm = p.copy(), p.copy(); pp[k] += h; pm[k] -= h
        np.testing.assert_allclose(seed[:, k], (flat(pp) - flat(pm)) / (2 * h), rtol=1e-6, atol=1e-9)


def test_f2_rhs_is_the_literal_formula_and_derivatives_match_fd():
    m, seed = cases.hychem_model(cases.hychem_p(0), YS)
    u = _state()
    for t in (0.0, 1.7e-6, 0.003, float(m.tab_t[7]), float(m.tab_t[-1])):
        f, J, dT = oracle.rhs_t(m, t, u)
        np.testing.assert_allclose(f, crnn_hychem_literal(u, t, m), rtol=1e-13)
    t = 0.003
    f, J, dT = oracle.rhs_t(m, t, u)
    Jfd = np.zeros_like(J)
    for l in range(9):
        h = 1e-6 * u[l]
        up, um = u.copy(), u.copy(); up[l] += h; um[l] -= h
        Jfd[:, l] = (crnn_hychem_literal(up, t, m) - crnn_hychem_literal(um, t, m)) / (2 * h)
    assert np.abs(J - Jfd).max() < 1e-6 * np.abs(J).max()
    h = 1e-8
    dTfd = (crnn_hychem_literal(u, t + h, m) - crnn_hychem_literal(u, t - h, m)) / (2 * h)
    assert np.abs(dT - dTfd).max() < 1e-6 * np.abs(dT).max()
    # clamp edges: a species below lb (chi = 0) and a concentration above ub (chiC = 0: N2 at 10 atm)
    ue = u.copy(); ue[7] = 1e-9
    fe, Je, _ = oracle.rhs_t(m, t, ue)
    np.testing.assert_allclose(fe, crnn_hychem_literal(ue, t, m), rtol=1e-13)
    assert np.all(Je[:, 7] == 0.0)
    # directional derivative along (S, seed column) = what the duals carry
    g = np.random.default_rng(0)
    S = g.standard_normal(9) * u
    for col in (3, 17, 28, 60, 150, 210):
        dS, _ = oracle.rhs_sens(m, u, S, seedcol=seed[:, col])
        eps = 1e-6
        def f_at(sg):
            p2 = cases.hychem_p(0); p2[col] += sg * eps
            m2, _ = cases.hychem_model(p2, YS)
            return crnn_hychem_literal(u + sg * eps * S, float(m.tab_t[0]), m2)
        fd = (f_at(+1) - f_at(-1)) / (2 * eps)
        assert np.abs(dS - fd).max() < 2e-6 * max(np.abs(fd).max(), 1e-12), col


@pytest.mark.parametrize("alg", ["tsit5", "ros23", "auto", "kc4"])
def test_f2_solvers_against_radau(alg):
    m, _ = cases.hychem_model(cases.hychem_p(0), YS)
    o = cases.hychem_opts(alg=ALG[alg])
    u0 = cases.hychem_u0(3)
    r = oracle.solve_batch(m, o, u0)
    assert (r["retcode"] == _abi.RET_SUCCESS).all() and (r["n_saved"] == 40).all()
    for i in range(3):
        sol = solve_ivp(lambda t, y: crnn_hychem_literal(y, min(t, m.tab_t[-1]), m), (0.0, o.t1), u0[i], method="Radau",
                        rtol=1e-10, atol=1e-14, t_eval=o.saveat)
        err = np.abs(r["pred"][i] - sol.y.T).max(axis=0) / YS
        assert err.max() < 5e-3, (alg, err)     # reltol 1e-3 solvers, errors relative to the species' scale
    if alg == "auto":
        att = r["stats"]["n_accept"] + r["stats"]["n_reject"]
        assert (r["stats"]["n_jac"] > 0).all() and (r["stats"]["n_jac"] < att).all()


def test_f2_forward_gradient_against_finite_differences():
    """ForwardDiff.gradient(x -> loss_n_ode(x, sample), p) (crnn_pyrolysis_mass.jl:143-147,201) on the F2 model: all 211
    columns through Tsit5, vs central differences of tight solves"""
    p = cases.hychem_p(2)
    m, seed = cases.hychem_model(p, YS)
    u0 = cases.hychem_u0(1)
    truth = oracle.solve_batch(cases.hychem_model(cases.hychem_p(5), YS)[0], cases.hychem_opts(alg=ALG["ros23"], abstol=1e-12, reltol=1e-9), u0)
    data = truth["pred"]
    tight = dict(alg=ALG["tsit5"], abstol=1e-13, reltol=1e-10, maxiters=1000000)
    o = cases.hychem_opts(**tight)
    r = oracle.loss_grad_batch(m, o, seed, u0, data, YS)
    assert r["retcode"][0] == _abi.RET_SUCCESS

    def loss_at(pv):
        mm, _ = cases.hychem_model(pv, YS)
        pr = oracle.solve_batch(mm, o, u0)["pred"]
        return np.mean(np.abs(pr / YS - data / YS))
    for k in (0, 7, 12, 24, 33, 101, 125, 199, 210):
        h = 1e-4    # the adaptive solve is only piecewise smooth in p (step-sequence noise ~1e-9): smaller h drowns in it
        pp, pm = p.copy(), p.copy(); pp[k] += h; pm[k] -= h
        fd = (loss_at(pp) - loss_at(pm)) / (2 * h)
        assert abs(r["grad_sum"][k] - fd) < 2e-3 * max(abs(fd), np.abs(r["grad_sum"]).max() * 1e-2), (k, r["grad_sum"][k], fd)


def test_autoswitch_is_inert_on_nonstiff_and_engages_on_stiff(golden):
    # non-stiff: the trained case2 CRNN never trips the detector -> identical to plain Tsit5 (case2.jl:26 as written)
    pb = make_problem("case2", golden, 64)
    c = pb["case"]
    a = oracle.solve_batch(pb["model"], c.opts(alg=ALG["auto"]), pb["u0"])
    b = oracle.solve_batch(pb["model"], c.opts(alg=ALG["tsit5"]), pb["u0"])
    assert np.array_equal(a["pred"], b["pred"]) and (a["stats"]["n_jac"] == 0).all()
    for k in ("n_accept", "n_reject", "n_rhs"):
        assert np.array_equal(a["stats"][k], b["stats"][k])
    # stiff: the true Robertson mechanism (rober_crnn.jl:56-63) on [0, 1e5]
    pr = make_problem("robertson", golden, 8)
    cr = pr["case"]
    o = cr.opts(alg=ALG["auto"], pred_clamp=(-np.inf, np.inf))
    r = oracle.solve_batch(pr["true_model"], o, pr["u0"])
    assert (r["retcode"] == _abi.RET_SUCCESS).all()
    att = r["stats"]["n_accept"] + r["stats"]["n_reject"]
    assert (r["stats"]["n_jac"] > 10).all() and (att - r["stats"]["n_jac"] >= 11).all()   # >= 11 Tsit5 attempts before the switch
    # a pure Tsit5 run of the same problem needs orders of magnitude more steps
    t5 = oracle.solve_batch(pr["true_model"], cr.opts(alg=ALG["tsit5"], pred_clamp=(-np.inf, np.inf), maxiters=3000), pr["u0"][:1])
    assert t5["retcode"][0] == _abi.RET_MAXITERS
    k = np.array([4e-2, 3e7, 1e4])
    def rob(t, y):
        return [-k[0] * y[0] + k[2] * y[1] * y[2], k[0] * y[0] - k[2] * y[1] * y[2] - k[1] * y[1] ** 2, k[1] * y[1] ** 2]
    sol = solve_ivp(rob, (0.0, 1e5), pr["u0"][0], method="Radau", rtol=1e-10, atol=1e-14, t_eval=o.saveat)
    scale = np.abs(sol.y).max(axis=1)
    assert (np.abs(r["pred"][0] - sol.y.T) / scale).max() < 2e-2


def test_autoswitch_forward_gradient_equals_components_when_no_switch(golden):
    """under forward sensitivities the composite carries the duals through whichever half runs"""
    pb = make_problem("case2", golden, 16)
    c = pb["case"]
    a = oracle.loss_grad_batch(pb["model"], c.opts(alg=ALG["auto"], obs_idx=pb["opts"].obs_idx), pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    b = oracle.loss_grad_batch(pb["model"], c.opts(alg=ALG["tsit5"], obs_idx=pb["opts"].obs_idx), pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    np.testing.assert_array_equal(a["grad_sum"], b["grad_sum"])
    pr = make_problem("robertson", golden, 4)
    cr = pr["case"]
    g_auto = oracle.loss_grad_batch(pr["model"], cr.opts(alg=ALG["auto"], abstol=1e-10, reltol=1e-8), pr["seed"], pr["u0"], pr["data"], pr["yscale"])
    g_ros = oracle.loss_grad_batch(pr["model"], cr.opts(alg=ALG["ros23"], abstol=1e-10, reltol=1e-8), pr["seed"], pr["u0"], pr["data"], pr["yscale"])
    assert (g_auto["stats"]["n_jac"] > 0).all()
    np.testing.assert_allclose(g_auto["grad_sum"], g_ros["grad_sum"], rtol=2e-3, atol=1e-4 * np.abs(g_ros["grad_sum"]).max())


def test_f2_argument_validation():
    m, _ = cases.hychem_model(cases.hychem_p(0), YS)
    o = cases.hychem_opts(alg=ALG["tsit5"])
    o.t1 = 1.0; o.saveat = np.array([0.0, 1.0])       # beyond the T(t), P(t) tables
    with pytest.raises(RuntimeError):
        oracle.solve_batch(m, o, cases.hychem_u0(1))


def test_reversible_crnn_is_an_f0_model_with_twice_the_reactions():
    """`case1 rev/case1.jl:80-89` transcribed literally vs the F0 oracle RHS on p2vec_case1_rev's weights; seed vs FD"""
    g = np.random.default_rng(4)
    ns, nr = 5, 10
    p = g.standard_normal(nr * (ns + 1)) * 0.5
    p[12] = 3.1; p[30] = -2.9                       # beyond the [-2.5, 2.5] clamp of w_out
    w_in, w_b, w_out, seed = cases.p2vec_case1_rev(p)
    assert w_in.shape == (5, 20) and w_out.shape == (5, 20) and seed.shape == (20 * 11, 60)
    w_kf = p[:nr]; wo = np.clip(p[nr:].reshape(nr, ns).T, -2.5, 2.5)
    u = 0.1 + g.random(ns)
    u_in = np.log(np.clip(u, 1e-5, np.inf))
    lit = wo @ (np.exp(np.clip(-wo, 0, 2.5).T @ u_in + w_kf) - np.exp(np.clip(wo, 0, 2.5).T @ u_in + w_kf))
    m, _ = cases.CASES["case1_rev"].model(p)
    np.testing.assert_allclose(oracle.rhs(m, u), lit, rtol=1e-13, atol=1e-15)
    flat = lambda q: np.concatenate([a.reshape(-1, order="F") for a in cases.p2vec_case1_rev(q)[:3]])
    for k in (0, 9, 11, 12, 30, 59):
        h = 1e-6
        pp, pm = p.copy(), p.copy(); pp[k] += h; pm[k] -= h
        np.testing.assert_allclose(seed[:, k], (flat(pp) - flat(pm)) / (2 * h), rtol=1e-6, atol=1e-9)
    # the generating network conserves A+B+C+D+E... (every reaction is 1 <-> 1 or 2 <-> 2 molecules)
    mt = cases.true_model_case1_rev()
    u0 = g.random((4, 5)); u0[:, :2] += 0.2
    r = oracle.solve_batch(mt, cases.CASES["case1_rev"].opts(), u0)
    assert (r["retcode"] == _abi.RET_SUCCESS).all()
    np.testing.assert_allclose(r["pred"].sum(axis=2), u0.sum(axis=1)[:, None] * np.ones((1, 100)), rtol=1e-9)


def test_golden_oracle_vectors_f2_auto():
    """Regression pin of the oracle's F2 / AutoTsit5 / adjoint-with-F2 parts (tests/golden/make_oracle_vectors_f2.py)."""
    import json, os, importlib.util
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("mk_f2", os.path.join(here, "make_oracle_vectors_f2.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    gv = json.load(open(os.path.join(here, "oracle_vectors_f2_auto.json")))
    for name, (m, o, u0) in mk.problems().items():
        r = oracle.solve_batch(m, o, u0)
        for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
            assert r["stats"][k].tolist() == gv[name][k], (name, k)
        np.testing.assert_allclose(r["pred"][:, ::7, :], np.array(gv[name]["pred_every7"]), rtol=1e-6, atol=1e-12)
    m, seed, u0, data = mk.gradient_problem()
    for mode, sm in (("forward", _abi.SENS_FORWARD), ("discrete", _abi.SENS_DISCRETE_ADJOINT), ("interp", _abi.SENS_INTERP_ADJOINT)):
        r = oracle.loss_grad_batch(m, cases.hychem_opts(alg=ALG["tsit5"], sens_mode=sm), seed, u0, data, YS)
        v = gv[f"hychem_grad_{mode}"]
        np.testing.assert_allclose(r["loss"], v["loss"], rtol=1e-9)
        np.testing.assert_allclose(r["grad_sum"], v["grad_sum"], rtol=1e-6, atol=1e-9 * np.abs(v["grad_sum"]).max())


def test_f2_model_validation_and_front_end_case():
    """host-side checks of the F2 model description (shapes, tables) and the HyChem `Case` of the front-end mirror"""
    from crnn_b200.model import CRNNModel
    w_in = np.zeros((11, 10)); w_out = np.zeros((9, 10)); w_b = np.zeros(10)
    tab = dict(tab_t=[0.0, 1.0], tab_T=[1300.0, 1200.0], tab_P=[1e6, 1e6])
    with pytest.raises(ValueError):
        CRNNModel(w_in=w_in[:10], w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F2, mw=cases.HYCHEM_MW, **tab)      # n_in != ns + 2
    with pytest.raises(ValueError):
        CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F2, **tab)                               # no molar masses
    with pytest.raises(ValueError):
        CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F2, mw=cases.HYCHEM_MW,
                  tab_t=[0.0, 0.0], tab_T=[1.0, 1.0], tab_P=[1.0, 1.0])                                       # knots not ascending
    m = CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F2, mw=cases.HYCHEM_MW, **tab)
    assert (m.n_state, m.n_species, m.n_in, m.n_w) == (9, 9, 11, 10 * (11 + 1 + 9))
    cm, keep = m.to_c()
    assert cm.n_tab == 2 and cm.n_state == 9 and cm.n_in == 11
    c = cases.hychem_case()
    mc, seed = c.model(cases.hychem_p(0), out_scale=YS / 0.01)
    mh, _ = cases.hychem_model(cases.hychem_p(0), YS)
    assert np.array_equal(mc.w_out, mh.w_out) and np.array_equal(mc.tab_T, mh.tab_T) and mc.gas_R == mh.gas_R
    o = c.opts()
    assert o.sens_mode == _abi.SENS_DISCRETE_ADJOINT and o.n_save == 40 and o.saveat[0] == 0.0
    assert abs(o.t1 - 0.01 / 1.01) < 1e-15 and np.all(np.diff(o.saveat) > 0)
    assert seed.shape == (mc.n_w, 211)


def test_gene_regulatory_checkpoint_reproduces_the_generating_mechanism(golden):
    """The reference's committed gene-regulatory checkpoint (p[285], iter 1290, train loss 4.28e-3 on 1 %-noisy data:
    SURVEY App. D.3) pushed through p2vec_gene (gene-regulatory.jl:39-50) and the oracle must reproduce the trajectories
    of the generating mechanism (:75-131) at that loss level — pins the p2vec orientation, the masked DNA rows, the RHS
    and the solver against a reference artifact."""
    c = cases.CASES["gene"]
    p = np.array(golden["gene"]["p"])
    assert p.size == 285 == c.n_p
    m, seed = c.model(p)
    assert np.all(m.w_out[[0, 3, 6], :] == 0.0) and seed.shape == (15 * 19, 285)
    mt = cases.true_model_gene()
    # the generating mechanism as a CRNN == trueODEfunc written out
    kk = np.exp(mt.w_b)
    y = 0.1 + np.random.default_rng(1).random(9)
    R = np.array([kk[0] * y[0], kk[1] * y[1], kk[2] * y[1], kk[3] * y[2], kk[4] * y[3], kk[5] * y[4], kk[6] * y[4], kk[7] * y[5],
                  kk[8] * y[6], kk[9] * y[7], kk[10] * y[7], kk[11] * y[8], kk[12] * y[7] * y[2], kk[13] * y[4] * y[8], kk[14] * y[1] * y[5]])
    lit = np.array([0, R[0] - R[2] - R[14], R[1] - R[3], 0, R[4] - R[6] - R[13], R[5] - R[7], 0, R[8] - R[10] - R[12], R[9] - R[11]])
    np.testing.assert_allclose(oracle.rhs(mt, y), lit, rtol=1e-13, atol=1e-15)
    u0 = np.random.default_rng(0).random((30, 9))                     # u0_list = rand(Float32, (n_exp, ns)) (:134)
    truth = oracle.solve_batch(mt, c.opts(pred_clamp=(-np.inf, np.inf)), u0, n_threads=8)
    pred = oracle.solve_batch(m, c.opts(), u0, n_threads=8)
    assert (truth["retcode"] == _abi.RET_SUCCESS).all() and (pred["retcode"] == _abi.RET_SUCCESS).all()
    mae = np.mean(np.abs(np.clip(truth["pred"], c.lb, c.ub) - pred["pred"]))    # loss_neuralode (:184-190)
    assert mae < 4.3e-3, mae                                          # measured 1.9e-3 on noise-free targets
    # the DNA species are constants of the trained model too
    np.testing.assert_allclose(pred["pred"][:, :, [0, 3, 6]], np.repeat(np.clip(u0[:, None, [0, 3, 6]], c.lb, c.ub), 40, axis=1), rtol=1e-12)


def test_mixed_second_derivative_of_the_f2_rhs_against_finite_differences():
    """D^2 f[(S, dW), (v, tau)] - the partials a nested-dual Jacobian / time derivative carry when
    ForwardDiff.gradient runs through Rosenbrock23 (crnn_pyrolysis_mass.jl:29,201) - against central differences of the
    first directional derivative, state AND time direction, on the non-autonomous density-coupled F2 RHS."""
    m, seed = cases.hychem_model(cases.hychem_p(3, stiff=2.0), YS)
    g = np.random.default_rng(0)
    u = cases.hychem_u0(4)[2]
    t = 0.37 * m.tab_t[-1]
    for col in (0, 5, 33, 120, 210):
        S = g.standard_normal(9) * u
        v = g.standard_normal(9) * u
        sc = seed[:, col]
        for tau in (0.0, 1.0, 2.5):
            _, d2 = oracle.rhs_sens_t(m, t, u, S, sc, v, tau)
            h = 1e-6
            hp = oracle.rhs_sens_t(m, t + h * tau * m.tab_t[-1] * 0 + h * tau, u + h * v, S, sc)[0]
            hm = oracle.rhs_sens_t(m, t - h * tau, u - h * v, S, sc)[0]
            fd = (hp - hm) / (2 * h)
            np.testing.assert_allclose(d2, fd, rtol=2e-5, atol=2e-6 * np.abs(fd).max())


@pytest.mark.parametrize("alg", ["ros23", "auto"])
def test_f2_stiff_forward_gradient_through_rosenbrock23_and_the_composite(alg):
    """What the HyChem script really runs: ForwardDiff.gradient through AutoTsit5(Rosenbrock23) on the F2 model
    (crnn_pyrolysis_mass.jl:29,201).  All 211 columns, vs central differences of tight solves."""
    p = cases.hychem_p(2, stiff=4.0)
    m, seed = cases.hychem_model(p, YS)
    u0 = cases.hychem_u0(1)
    data = oracle.solve_batch(cases.hychem_model(cases.hychem_p(5, stiff=4.0), YS)[0],
                              cases.hychem_opts(alg=ALG["ros23"], abstol=1e-12, reltol=1e-9), u0)["pred"]
    o = cases.hychem_opts(alg=ALG[alg], abstol=1e-11, reltol=1e-7, maxiters=1000000)
    r = oracle.loss_grad_batch(m, o, seed, u0, data, YS)
    assert r["retcode"][0] == _abi.RET_SUCCESS
    if alg == "auto":
        assert r["stats"]["n_jac"][0] > 0          # the composite did switch to Rosenbrock23

    def loss_at(pv):
        mm, _ = cases.hychem_model(pv, YS)
        pr = oracle.solve_batch(mm, o, u0)["pred"]
        return np.mean(np.abs(pr / YS - data / YS))
    # the composite's switch points move discretely with p: its loss is only piecewise smooth, a wider difference
    # averages over the jumps
    h, tol = (1e-4, 5e-3) if alg == "ros23" else (1e-3, 2e-2)
    for k in (0, 7, 12, 24, 33, 101, 125, 199, 210):
        pp, pm = p.copy(), p.copy(); pp[k] += h; pm[k] -= h
        fd = (loss_at(pp) - loss_at(pm)) / (2 * h)
        assert abs(r["grad_sum"][k] - fd) < tol * max(abs(fd), np.abs(r["grad_sum"]).max() * 1e-2), (k, r["grad_sum"][k], fd)
    if alg == "auto":   # and it is the derivative Rosenbrock23 alone gives, to the integration tolerance
        rr = oracle.loss_grad_batch(m, cases.hychem_opts(alg=ALG["ros23"], abstol=1e-11, reltol=1e-7, maxiters=1000000), seed, u0, data, YS)
        np.testing.assert_allclose(r["grad_sum"], rr["grad_sum"], rtol=1e-3, atol=1e-4 * np.abs(rr["grad_sum"]).max())
