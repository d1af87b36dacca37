"""CPU tests that pin the oracle: against the reference's committed checkpoints (golden
fixtures), the generating mechanisms written in the scripts, published order conditions of the
tableaus, an independent integrator (scipy Radau at 1e-12) and finite differences."""
import json
import os

import numpy as np
from dataclasses import replace as dataclasses_replace
import pytest
from scipy.integrate import solve_ivp

from crnn_b200 import _abi, cases
from oracle import oracle
from problems import make_problem

HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- tableaus
def test_tsit5_order_conditions():
    a, bt, r = oracle.tsit5_tableau()
    A = np.zeros((7, 7)); A[:, :6] = a
    b = A[6].copy()                       # FSAL: b = a7.
    c = A.sum(axis=1)
    assert np.allclose(c, [0, 0.161, 0.327, 0.9, 0.9800255409045097, 1, 1], atol=1e-15)
    one = np.ones(7)
    Ac = A @ c; Ac2 = A @ c**2; AAc = A @ Ac
    conds = [  # all 17 rooted-tree conditions up to order 5
        (b @ one, 1), (b @ c, 1 / 2), (b @ c**2, 1 / 3), (b @ Ac, 1 / 6),
        (b @ c**3, 1 / 4), (b @ (c * Ac), 1 / 8), (b @ Ac2, 1 / 12), (b @ AAc, 1 / 24),
        (b @ c**4, 1 / 5), (b @ (c**2 * Ac), 1 / 10), (b @ (c * Ac2), 1 / 15), (b @ (c * AAc), 1 / 30),
        (b @ (Ac * Ac), 1 / 20), (b @ (A @ c**3), 1 / 20), (b @ (A @ (c * Ac)), 1 / 40),
        (b @ (A @ Ac2), 1 / 60), (b @ (A @ AAc), 1 / 120)]
    for got, want in conds:
        assert abs(got - want) < 2e-14   # coefficients up to 12.9: a few ulp of cancellation
    # embedded error weights: sum 0, and b - btilde is an order-4 method
    assert abs(bt.sum()) < 1e-15
    bh = b - bt
    for got, want in [(bh @ one, 1), (bh @ c, 1 / 2), (bh @ c**2, 1 / 3), (bh @ Ac, 1 / 6),
                      (bh @ c**3, 1 / 4), (bh @ (c * Ac), 1 / 8), (bh @ Ac2, 1 / 12), (bh @ AAc, 1 / 24)]:
        assert abs(got - want) < 1e-14
    # dense output: b(1) = b, b(0) = 0, order-4 conditions at interior points
    bth = lambda th: np.array([th * (r[j, 0] + th * (r[j, 1] + th * (r[j, 2] + th * r[j, 3]))) for j in range(7)])
    assert np.allclose(bth(1.0), b, atol=1e-14)
    for th in (0.25, 0.5, 0.8):
        w = bth(th)
        assert abs(w @ one - th) < 1e-14 and abs(w @ c - th**2 / 2) < 1e-14
        assert abs(w @ c**2 - th**3 / 3) < 1e-13 and abs(w @ Ac - th**3 / 6) < 1e-13
        assert abs(w @ c**3 - th**4 / 4) < 1e-13 and abs(w @ AAc - th**4 / 24) < 1e-13


def test_rosenbrock23_order_two_on_linear_problem():
    """ode23s is exact-Jacobian order 2: halving the tolerance-driven step on u' = -u."""
    m = cases.CRNNModel(w_in=np.ones((1, 1)), w_b=np.zeros(1), w_out=-np.ones((1, 1)), lb=1e-300, ub=np.inf)
    # CRNN with w_in = 1, w_out = -1: du = -exp(log u) = -u.  Cfg (1,1) is oracle-only.
    errs = []
    for rt in (1e-4, 1e-6):
        o = cases.SolveOpts(saveat=np.array([1.0]), t0=0.0, t1=1.0, alg=_abi.ALG_ROSENBROCK23, abstol=rt * 1e-3, reltol=rt)
        r = oracle.solve_batch(m, o, np.array([[1.0]]))
        errs.append(abs(r["pred"][0, 0, 0] - np.exp(-1.0)))
    # error-per-step control with an order-2 method: global error ~ tol^(2/3) (100x tol -> ~20x)
    assert errs[0] < 3e-4 and errs[1] < errs[0] / 8


# ---------------------------------------------------------------- golden fixtures (reference checkpoints)
def test_p2vec_case2_checkpoint_recovers_generating_mechanism(golden):
    """the trained checkpoint decodes to the transesterification kinetics."""
    p = np.array(golden["case2"]["p"])
    assert len(p) == 25 and golden["case2"]["iter"] == 3700
    w_in, w_b, w_out, seed = cases.p2vec_case2(p)
    assert abs(100 * p[-1] - 14.0743) < 1e-3                     # slope
    np.testing.assert_allclose(w_b, [19.3145, 18.3490, 7.8600], atol=1e-3)
    np.testing.assert_allclose(w_in[6], [14.5410, 14.4022, 6.4345], atol=1e-3)
    true = golden["case2_true"]
    # reactions come out permuted: (DG+ROH), (TG+ROH), (MG+ROH)
    np.testing.assert_allclose(np.sort(w_in[6]), np.sort(true["Ea"]), atol=0.05)
    np.testing.assert_allclose(np.sort(w_b), np.sort(true["logA"]), atol=0.85)
    stoich = np.array([[0, -1, -1, 1, 0, 1], [-1, -1, 1, 0, 0, 1], [0, -1, 0, -1, 1, 1]], dtype=float).T
    np.testing.assert_allclose(w_out, stoich, atol=0.08)
    assert seed.shape == (3 * (7 + 1 + 6), 25)


def test_p2vec_robertson_checkpoint_table(golden):
    """rows [w_in' | w_b | w_out'] of the stiff trained CRNN."""
    p = np.array(golden["robertson"]["p"])
    assert len(p) == 43 and golden["robertson"]["iter"] == 10850
    w_in, w_b, w_out, _ = cases.p2vec_robertson(p)
    assert abs(abs(p[-1]) - 1.28268) < 1e-5
    np.testing.assert_allclose(w_b, [3.995, 28.75, 5.847, 29.03, -13.45, -3.982], rtol=2e-3)
    np.testing.assert_allclose(w_in[:, 0], [0.2118, 0.3527, 0.0], atol=1e-4)
    np.testing.assert_allclose(w_in[:, 2], [2.5, 0.0, 0.0], atol=1e-12)      # clamp(.,0,2.5) active
    np.testing.assert_allclose(w_out[1], [-2.44e-3, -107.06, 791.9, -532.4, 4.21e-2, 1.42e-3], rtol=5e-3)
    # readme table (another trained model) has the same layout: 6 reactions x (3 + 1 + 3)
    assert np.array(golden["robertson_readme"]["table"]).shape == (6, 7)


def test_p2vec_seed_matches_finite_differences(golden):
    rng = np.random.default_rng(0)
    for name, fn, n_p in (("case1", cases.p2vec_case1, 24), ("case2", cases.p2vec_case2, 25),
                          ("case3", cases.p2vec_case3, 153), ("robertson", cases.p2vec_robertson, 43)):
        p = rng.uniform(-1.5, 1.5, n_p)
        p[np.abs(p) < 0.05] = 0.3            # stay off the clamp/abs kinks
        *_, seed = fn(p)
        flat = lambda q: np.concatenate([w.reshape(-1, order="F") for w in fn(q)[:3]])
        fd = np.stack([(flat(p + 1e-6 * e) - flat(p - 1e-6 * e)) / 2e-6 for e in np.eye(n_p)], axis=1)
        np.testing.assert_allclose(seed, fd, atol=2e-7 * max(1.0, np.abs(fd).max()))


def test_trained_case2_crnn_reproduces_generating_mechanism(golden):
    """Physics pin: the checkpointed CRNN tracks trueODEfunc (case2.jl:38-59) to ~1-2 % normalised
    MAE — the level of the reference's own final loss (1.4e-2..1.7e-2)."""
    pb = make_problem("case2", golden, 64, noise=0.0)
    r = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    assert (r["retcode"] == 1).all()
    assert 0.002 < r["loss"].mean() < 0.02
    noisy = make_problem("case2", golden, 64, noise=0.05)
    rn = oracle.loss_grad_batch(noisy["model"], noisy["opts"], noisy["seed"], noisy["u0"], noisy["data"], noisy["yscale"])
    last = golden["case2"]["l_loss_val"]["last"]
    assert 0.5 * last < rn["loss"].mean() < 2.0 * last


# ---------------------------------------------------------------- RHS / Jacobian / directional derivatives
def _true_rhs_case2(u, g):
    R = g["R"]; k = np.exp(np.array(g["logA"])) * np.exp(-np.array(g["Ea"]) / R / u[6])
    r1, r2, r3 = k[0] * u[0] * u[1], k[1] * u[2] * u[1], k[2] * u[3] * u[1]
    return np.array([-r1, -r1 - r2 - r3, r1 - r2, r2 - r3, r3, r1 + r2 + r3, 0.0])


def test_true_mechanisms_as_crnn(golden):
    rng = np.random.default_rng(1)
    m = cases.true_model_case2()
    for _ in range(5):
        u = np.r_[rng.uniform(0.05, 2, 6), rng.uniform(323, 343)]
        np.testing.assert_allclose(oracle.rhs(m, u), _true_rhs_case2(u, golden["case2_true"]), rtol=1e-12)
    k = golden["robertson_true"]["k"]
    y = np.array([0.9, 2e-5, 0.3])
    want = [-k[0] * y[0] + k[2] * y[1] * y[2], k[0] * y[0] - k[1] * y[1]**2 - k[2] * y[1] * y[2], k[1] * y[1]**2]
    np.testing.assert_allclose(oracle.rhs(cases.true_model_robertson(), y), want, rtol=1e-12)
    k = golden["case1_true"]["k"]
    y = rng.uniform(0.1, 1, 5)
    want = [-2 * k[0] * y[0]**2 - k[1] * y[0], k[0] * y[0]**2 - k[3] * y[1] * y[3], k[1] * y[0] - k[2] * y[2],
            k[2] * y[2] - k[3] * y[1] * y[3], k[3] * y[1] * y[3]]
    np.testing.assert_allclose(oracle.rhs(cases.true_model_case1(), y), want, rtol=1e-12)
    y = rng.uniform(0.1, 1, 9)
    r = [y[0] * y[1], y[2] * y[3], y[4] * y[5], y[6] * y[7], y[2], y[4], y[6], y[8]]
    want = [0, -r[0] + r[4], r[0] - r[4], -r[1] + r[5], r[1] - r[5], -r[2] + r[6], r[2] - r[6], -r[3] + r[7], r[3] - r[7]]
    np.testing.assert_allclose(oracle.rhs(cases.true_model_case3(), y), want, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", ["case2", "robertson", "case3"])
def test_jacobian_and_directional_derivatives_vs_fd(golden, name):
    pb = make_problem(name, golden, 2)
    m, seed = pb["model"], pb["seed"]
    rng = np.random.default_rng(2)
    u = pb["u0"][0].copy()
    u[:m.n_species] = rng.uniform(0.05, 1.0, m.n_species)
    f, J = oracle.rhs(m, u, want_jac=True)
    h = 1e-6
    Jfd = np.stack([(oracle.rhs(m, u + h * u[i] * e) - oracle.rhs(m, u - h * u[i] * e)) / (2 * h * u[i])
                    for i, e in enumerate(np.eye(m.n_state))], axis=1)
    np.testing.assert_allclose(J, Jfd, rtol=1e-6, atol=1e-7 * np.abs(Jfd).max())
    # df along (S, seed column) and d(J v) along the same direction (nested-dual terms)
    c = pb["case"]
    col = 3
    S = rng.standard_normal(m.n_state) * 0.1; S[m.n_species:] = 0
    v = rng.standard_normal(m.n_state); v[m.n_species:] = 0
    dS, dJv = oracle.rhs_sens(m, u, S, seed[:, col], v)
    p0 = np.array(golden[name]["p"]) if name in golden and "p" in golden[name] else None
    if p0 is None:
        from problems import trained_p
        p0 = trained_p(name, golden)
    e = np.zeros_like(p0); e[col] = 1.0
    eps = 1e-6
    mp, _ = c.model(p0 + eps * e, m.out_scale); mm, _ = c.model(p0 - eps * e, m.out_scale)
    fd = (oracle.rhs(mp, u + eps * S) - oracle.rhs(mm, u - eps * S)) / (2 * eps)
    np.testing.assert_allclose(dS, fd, rtol=2e-5, atol=1e-7 * max(1.0, np.abs(fd).max()))
    Jp = oracle.rhs(mp, u + eps * S, want_jac=True)[1]; Jm = oracle.rhs(mm, u - eps * S, want_jac=True)[1]
    fd2 = (Jp - Jm) @ v / (2 * eps)
    np.testing.assert_allclose(dJv, fd2, rtol=5e-5, atol=1e-6 * max(1.0, np.abs(fd2).max()))


def test_clamp_edges_of_the_rhs():
    m = cases.CRNNModel(w_in=np.eye(2), w_b=np.zeros(2), w_out=-np.eye(2), lb=1e-5, ub=10.0)
    f, J = oracle.rhs(m, np.array([0.0, 20.0]), want_jac=True)     # below lb / above ub
    np.testing.assert_allclose(f, [-1e-5, -10.0], rtol=1e-14)
    assert np.all(J == 0)                                          # dual clamp: derivative 0 outside
    f, J = oracle.rhs(m, np.array([1e-5, 10.0]), want_jac=True)    # on the closed interval: derivative 1
    np.testing.assert_allclose(np.diag(J), [-1.0, -1.0], rtol=1e-12)


# ---------------------------------------------------------------- integrator vs independent truth
def test_tsit5_case2_against_radau(golden):
    pb = make_problem("case2", golden, 6)
    r = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    for i in range(6):
        sol = solve_ivp(lambda t, u: oracle.rhs(pb["model"], u), (0, 50), pb["u0"][i], method="Radau",
                        rtol=1e-12, atol=1e-14, t_eval=pb["opts"].saveat)
        assert np.abs(sol.y[:6].T - r["pred"][i]).max() < 2e-4         # reltol 1e-3 run
    tight = pb["case"].opts(obs_idx=np.arange(6), abstol=1e-12, reltol=1e-10)
    rt = oracle.solve_batch(pb["model"], tight, pb["u0"][:1])
    sol = solve_ivp(lambda t, u: oracle.rhs(pb["model"], u), (0, 50), pb["u0"][0], method="Radau",
                    rtol=1e-12, atol=1e-14, t_eval=tight.saveat)
    assert np.abs(sol.y[:6].T - rt["pred"][0]).max() < 2e-9
    # survey scratch value: IC [1.2,1.5,0,0,0,0,333] takes 17 steps / 104 RHS evaluations
    ic = np.array([[1.2, 1.5, 0, 0, 0, 0, 333.0]])
    s = oracle.solve_batch(pb["model"], pb["opts"], ic)["stats"]
    assert (s["n_accept"][0], s["n_reject"][0], s["n_rhs"][0]) == (17, 0, 104)
    # ... and, with the 25 forward-sensitivity columns in DiffEqBase's dual norm (mean over n_state*(1+np) numbers,
    # SURVEY App. C.3), 20 steps / 122 RHS evaluations
    counts = lambda st: (int(st["n_accept"][0]), int(st["n_reject"][0]), int(st["n_rhs"][0]))
    lg = lambda **kw: oracle.loss_grad_batch(pb["model"], pb["case"].opts(obs_idx=np.arange(6), **kw), pb["seed"], ic,
                                             pb["data"][:1], pb["yscale"])["stats"]
    assert counts(lg()) == (20, 0, 122)
    # named switches: the mean over the n_state rows only (releases before 0.2.0), and the value-only norm
    assert counts(lg(err_norm_mean_over_partials=False)) == (23, 0, 140)
    assert counts(lg(err_norm_includes_sens=False)) == (17, 0, 104)


def test_rosenbrock23_robertson_against_radau_and_mass(golden):
    mt = cases.true_model_robertson()
    c = cases.CASES["robertson"]
    o = c.opts()
    u0 = np.array([[1.0, 1e-8, 1.0], [0.6, 1e-8, 1.4]])
    r = oracle.solve_batch(mt, o, u0)
    assert (r["retcode"] == 1).all() and (r["n_saved"] == 40).all()
    for i in range(2):
        sol = solve_ivp(lambda t, u: oracle.rhs(mt, u), (0, 1e5), u0[i], method="Radau", rtol=1e-12, atol=1e-16,
                        t_eval=o.saveat)
        scale = np.abs(sol.y).max(axis=1)
        assert (np.abs(sol.y.T - r["pred"][i]) / scale).max() < 2e-3
        # Robertson conserves y1+y2+y3 (rober_crnn.jl:56-63); linear invariants survive the W-method
        assert np.abs(r["pred"][i].sum(axis=1) - u0[i].sum()).max() < 1e-9
    assert 40 <= r["stats"]["n_accept"][0] <= 90 and r["stats"]["n_jac"][0] == r["stats"]["n_accept"][0] + r["stats"]["n_reject"][0]


def test_case2_element_balances(golden):
    pb = make_problem("case2", golden, 16)
    o = pb["case"].opts(obs_idx=np.arange(6), pred_clamp=(-np.inf, np.inf))
    y = oracle.solve_batch(pb["true_model"], o, pb["u0"])["pred"]
    for w in ([1, 0, 1, 1, 1, 0], [3, 0, 2, 1, 0, 1], [0, 1, 0, 0, 0, 1]):
        inv = y @ np.array(w, dtype=float)
        assert np.abs(inv - inv[:, :1]).max() < 1e-3     # linear invariants: exact up to the lb clamp


# ---------------------------------------------------------------- sensitivities
@pytest.mark.parametrize("name,tol", [("case2", 1e-6), ("case1", 1e-6)])
def test_forward_gradient_vs_finite_differences(golden, name, tol):
    pb = make_problem(name, golden, 2)
    c = pb["case"]
    p = np.array(golden[name]["p"]) if name == "case2" else None
    if p is None:
        from problems import trained_p
        p = trained_p(name, golden)
    ot = c.opts(obs_idx=np.arange(c.ns), abstol=1e-12, reltol=1e-10)
    args = (pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    g = oracle.loss_grad_batch(pb["model"], ot, pb["seed"], *args)["grad_sum"]
    L = lambda q: oracle.loss_grad_batch(c.model(q, pb["model"].out_scale)[0], ot, pb["seed"], *args)["loss"].sum()
    fd = np.array([(L(p + 1e-6 * e) - L(p - 1e-6 * e)) / 2e-6 for e in np.eye(len(p))])
    assert np.linalg.norm(g - fd) / np.linalg.norm(fd) < tol
    # run tolerances: discrete sensitivities of the coarse step sequence stay within 1e-4
    g2 = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], *args)["grad_sum"]
    assert np.linalg.norm(g2 - fd) / np.linalg.norm(fd) < 2e-4


def test_rosenbrock_forward_gradient_vs_finite_differences(golden):
    """nested-dual terms (dJ) of Rosenbrock23(autodiff=true) under ForwardDiff (rober_crnn.jl:33,219)."""
    pb = make_problem("robertson", golden, 1)
    c = pb["case"]; p = np.array(golden["robertson"]["p"])
    ot = c.opts(abstol=np.array([1e-10, 1e-12, 1e-10]), reltol=np.full(3, 1e-6), maxiters=10**7)
    args = (pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    g = oracle.loss_grad_batch(pb["model"], ot, pb["seed"], *args)["grad_sum"]
    L = lambda q: oracle.loss_grad_batch(c.model(q, pb["model"].out_scale)[0], ot, pb["seed"], *args)["loss"].sum()
    fd = np.array([(L(p + 1e-5 * e) - L(p - 1e-5 * e)) / 2e-5 for e in np.eye(len(p))])
    assert np.linalg.norm(g - fd) / np.linalg.norm(fd) < 2e-5


def test_partials_in_error_norm_switch(golden):
    pb = make_problem("case2", golden, 8)
    on = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    off_o = pb["case"].opts(obs_idx=np.arange(6), err_norm_includes_sens=False)
    off = oracle.loss_grad_batch(pb["model"], off_o, pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    val = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    assert np.array_equal(off["stats"]["n_accept"], val["stats"]["n_accept"])     # value-only norm
    assert (on["stats"]["n_accept"] >= off["stats"]["n_accept"]).all()            # partials tighten steps
    assert (on["stats"]["n_accept"] > off["stats"]["n_accept"]).any()


def test_config1_case1_as_written_and_effective_tolerances(golden):
    """BASELINE config 1: case1.jl:29-30 writes atol 1e-5 / rtol 1e-2 (as `atol=`/`rtol=`, which the
    Julia-1.6-era solver ignored, so the effective values were 1e-6 / 1e-3): both are run, against Radau."""
    pb = make_problem("case1", golden, 3)
    c = pb["case"]
    at, rt = cases.AS_WRITTEN_TOL["case1"]
    written = c.opts(abstol=at, reltol=rt)
    rw = oracle.solve_batch(pb["true_model"], written, pb["u0"])
    re = oracle.solve_batch(pb["true_model"], c.opts(), pb["u0"])
    assert (rw["retcode"] == 1).all() and (re["retcode"] == 1).all()
    assert (rw["stats"]["n_accept"] <= re["stats"]["n_accept"]).all() and (rw["stats"]["n_accept"] < re["stats"]["n_accept"]).any()
    for i in range(3):
        sol = solve_ivp(lambda t, u: oracle.rhs(pb["true_model"], u), c.tspan, pb["u0"][i], method="Radau",
                        rtol=1e-12, atol=1e-14, t_eval=written.saveat)
        ew, ee = np.abs(sol.y.T - rw["pred"][i]).max(), np.abs(sol.y.T - re["pred"][i]).max()
        assert ew < 5e-3 and ee < 5e-4, (ew, ee)
    # loss + gradient at both settings: forward mode vs the discrete adjoint of the value-norm solve
    for o in (written, c.opts()):
        ov = c.opts(abstol=o.abstol, reltol=o.reltol, err_norm_includes_sens=False)
        oa = c.opts(abstol=o.abstol, reltol=o.reltol, sens_mode=_abi.SENS_DISCRETE_ADJOINT)
        f = oracle.loss_grad_batch(pb["model"], ov, pb["seed"], pb["u0"], pb["data"], pb["yscale"])
        a = oracle.loss_grad_batch(pb["model"], oa, pb["seed"], pb["u0"], pb["data"], pb["yscale"])
        np.testing.assert_allclose(a["loss"], f["loss"], rtol=1e-12)
        np.testing.assert_allclose(a["grad_sum"], f["grad_sum"], rtol=1e-8, atol=1e-11)


def test_qsteady_dead_band_defaults_and_switch(golden):
    """step_accept_controller!: qsteady_min <= q <= qsteady_max keeps dt.  Defaults 1 / 1.2 for Rosenbrock23 and
    KenCarp4 (adaptive implicit), 1 / 1 for Tsit5 and the composite; both are overridable."""
    c = cases.CASES["robertson"]
    pbr = make_problem("robertson", golden, 64)
    mt, u0 = pbr["model"], pbr["u0"]
    dflt = oracle.solve_batch(mt, c.opts(), u0)
    explicit = oracle.solve_batch(mt, c.opts(controller=dict(qsteady_min=1.0, qsteady_max=1.2)), u0)
    none = oracle.solve_batch(mt, c.opts(controller=dict(qsteady_min=1.0, qsteady_max=1.0)), u0)
    assert np.array_equal(dflt["pred"], explicit["pred"])
    assert not np.array_equal(dflt["pred"], none["pred"])          # the dead-band is hit on the trained stiff CRNN
    assert np.abs(dflt["pred"] - none["pred"]).max() < 5e-3       # same solution within the tolerance
    # Tsit5: no dead-band by default; switching one on changes the step sequence
    pb = make_problem("case2", golden, 4)
    a = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    b = oracle.solve_batch(pb["model"], pb["case"].opts(obs_idx=np.arange(6), controller=dict(qsteady_min=1.0, qsteady_max=1.0)), pb["u0"])
    d = oracle.solve_batch(pb["model"], pb["case"].opts(obs_idx=np.arange(6), controller=dict(qsteady_min=0.5, qsteady_max=1.3)), pb["u0"])
    assert np.array_equal(a["pred"], b["pred"]) and not np.array_equal(a["pred"], d["pred"])


def test_lu_literal_division_is_the_default_and_reciprocal_is_a_switch(golden):
    """The oracle's LU follows the generic lu!/ldiv! (divisions) by default; the CUDA kernels' reciprocal-diagonal form
    is a named switch.  The two differ by rounding only: same step counts, states to ~1e-9."""
    c = cases.CASES["robertson"]
    pb = make_problem("robertson", golden, 16)
    assert oracle.lib().crnn_oracle_get_lu_reciprocal() == 0
    lit = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    with oracle.lu_reciprocal():
        rec = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    assert oracle.lib().crnn_oracle_get_lu_reciprocal() == 0
    assert np.array_equal(lit["stats"]["n_accept"], rec["stats"]["n_accept"])
    assert np.array_equal(lit["stats"]["n_reject"], rec["stats"]["n_reject"])
    assert not np.array_equal(lit["pred"], rec["pred"])           # the switch does something
    np.testing.assert_allclose(rec["pred"], lit["pred"], rtol=1e-8, atol=1e-14)


# ---------------------------------------------------------------- failure / truncation paths
def test_retcodes_and_truncation(golden):
    pb = make_problem("case2", golden, 5)
    c = pb["case"]
    r = oracle.solve_batch(pb["model"], c.opts(obs_idx=np.arange(6), maxiters=4), pb["u0"])
    assert (r["retcode"] == _abi.RET_MAXITERS).all()
    assert ((r["stats"]["n_accept"] + r["stats"]["n_reject"]) == 4).all() and (r["n_saved"] < 50).all()
    for i in range(5):
        assert np.all(r["pred"][i, r["n_saved"][i]:] == 0)
    m = cases.CRNNModel(w_in=pb["model"].w_in, w_b=pb["model"].w_b * np.nan, w_out=pb["model"].w_out,
                        rhs_kind=_abi.RHS_F1, lb=1e-6, ub=10.0)
    assert (oracle.solve_batch(m, pb["opts"], pb["u0"])["retcode"] == _abi.RET_DTNAN).all()
    lg = oracle.loss_grad_batch(pb["model"], c.opts(obs_idx=np.arange(6), maxiters=1, saveat=np.linspace(1, 50, 50)),
                                pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    assert np.isnan(lg["loss"]).all() and np.all(lg["grad_sum"] == 0)   # nothing saved: NaN loss, zero grad


def test_random_time_truncation(golden):
    """rober_crnn.jl:125,218: tspan = [0, tsteps[sample]], loss over 1:sample."""
    pb = make_problem("case2", golden, 4)
    nsu = np.array([50, 10, 1, 33], dtype=np.int32)
    r = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], n_save_used=nsu,
                               want_pred=True)
    assert np.array_equal(r["n_saved"], nsu)
    full = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    assert np.abs(r["pred"][1, :10] - full["pred"][1, :10]).max() < 5e-3   # different step sequence, same solution
    assert np.all(r["pred"][1, 10:] == 0) and r["stats"]["t_reached"][2] == 0.0


def test_golden_oracle_vectors():
    """Regression pin of the oracle itself: vectors written by tests/golden/make_oracle_vectors.py."""
    with open(os.path.join(HERE, "golden", "oracle_vectors.json")) as f:
        gv = json.load(f)
    with open(os.path.join(HERE, "golden", "checkpoints.json")) as f:
        golden = json.load(f)
    for name, v in gv.items():
        pb = make_problem(name, golden, v["N"])
        if pb["opts"].alg == _abi.ALG_TSIT5 and pb["seed"].shape[1] <= 63:
            r = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"],
                                       pb["loss_kind"], want_pred=True)
            np.testing.assert_allclose(r["loss"], v["loss"], rtol=1e-9)
            np.testing.assert_allclose(r["grad_sum"], v["grad_sum"], rtol=1e-7, atol=1e-9 * np.abs(v["grad_sum"]).max())
        else:
            r = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
        assert r["stats"]["n_accept"].tolist() == v["n_accept"] and r["stats"]["n_reject"].tolist() == v["n_reject"]
        np.testing.assert_allclose(r["pred"][:, ::7, :], np.array(v["pred_every7"]), rtol=1e-6, atol=1e-10)


# ---------------------------------------------------------------- KenCarp4 (BASELINE config 5; not in the reference)
def test_kencarp4_tableau_order_conditions():
    """ESDIRK4(3)6L[2]SA: stiffly accurate, stage order 2, main method order 4, embedded order 3."""
    A, bhat = oracle.kencarp4_tableau()
    b = A[5].copy(); c = A.sum(axis=1); one = np.ones(6)
    assert np.allclose(c, [0, 1 / 2, 83 / 250, 31 / 50, 17 / 20, 1], atol=1e-14)
    assert np.allclose(np.diag(A)[1:], 0.25) and A[0, 0] == 0
    Ac = A @ c
    for got, want in [(b @ one, 1), (b @ c, 1 / 2), (b @ c**2, 1 / 3), (b @ Ac, 1 / 6), (b @ c**3, 1 / 4),
                      (b @ (c * Ac), 1 / 8), (b @ (A @ c**2), 1 / 12), (b @ (A @ Ac), 1 / 24)]:
        assert abs(got - want) < 1e-13
    for got, want in [(bhat @ one, 1), (bhat @ c, 1 / 2), (bhat @ c**2, 1 / 3), (bhat @ Ac, 1 / 6)]:
        assert abs(got - want) < 1e-13
    assert np.allclose(A[1:] @ c, c[1:]**2 / 2, atol=1e-13)       # stage order 2


def test_kencarp4_against_radau_and_conservation(golden):
    c = cases.CASES["robertson"]
    pb = make_problem("robertson", golden, 6)
    o = c.opts(alg=_abi.ALG_KENCARP4)
    r = oracle.solve_batch(pb["model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["model"], c.opts(abstol=np.array([1e-10, 1e-12, 1e-10]), reltol=np.full(3, 1e-7),
                                                  maxiters=10**7), pb["u0"])          # tight Rosenbrock23
    assert (r["retcode"] == 1).all() and (r["stats"]["n_accept"] < 60).all()
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    assert (np.abs(r["pred"][:, -1] - ref["pred"][:, -1]) / scale).max() < 2e-3       # step endpoints: the method itself
    # HyChem-sized synthetic model (29 species + T, 30 reactions): sum(u) is conserved by construction
    m = cases.synthetic_stiff_model(); u0 = cases.synthetic_stiff_u0(8); so = cases.synthetic_stiff_opts()
    s = oracle.solve_batch(m, so, u0, n_threads=4)
    assert (s["retcode"] == 1).all()
    assert np.abs(s["pred"][:, -1].sum(axis=1) - u0[:, :29].sum(axis=1)).max() < 1e-6
    tight = cases.synthetic_stiff_opts(); tight.abstol = 1e-12; tight.reltol = 1e-8; tight.maxiters = 10**6
    st = oracle.solve_batch(m, tight, u0, n_threads=4)
    assert np.abs(st["pred"][:, -1] - s["pred"][:, -1]).max() < 2e-3          # run tolerance vs tight, step endpoint
    assert (s["stats"]["n_accept"] < 120).all() and (st["stats"]["n_accept"] > s["stats"]["n_accept"]).all()
    from crnn_b200.engine import EngineError  # noqa: F401  (KenCarp4 has no sensitivity path: value only)
    with pytest.raises(RuntimeError):
        oracle.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"])


# ---------------------------------------------------------------- TRBDF2 and AutoTsit5(TRBDF2) (Cathode/src/network.jl:102)
def test_trbdf2_is_second_order_and_l_stable():
    """the tableau the oracle hard-codes (gamma = 2 - sqrt2, d = gamma/2, w = sqrt2/4): order conditions up to 2, the
    embedded weights sum to zero with a vanishing first moment (3rd-order companion), R(infinity) = 0; and the measured
    convergence order of the fixed-tolerance-free method on the linear test problem through the oracle's own stepping"""
    s2 = np.sqrt(2.0); gam = 2 - s2; d = gam / 2; w = s2 / 4
    A = np.array([[0, 0, 0], [d, d, 0], [w, w, d]]); b = A[2]; c = A.sum(axis=1)
    assert np.allclose(c, [0, gam, 1]) and abs(b.sum() - 1) < 1e-15 and abs(b @ c - 0.5) < 1e-15
    bt = np.array([(1 - s2) / 3, 1 / 3, (s2 - 2) / 3])
    assert abs(bt.sum()) < 1e-15
    bhat = b + bt                                            # Hosea-Shampine companion: order 3
    for got, want in [(bhat @ c, 1 / 2), (bhat @ c**2, 1 / 3), (bhat @ (A @ c), 1 / 6)]:
        assert abs(got - want) < 1e-14
    e = np.ones(3)                                           # stability function at z -> -inf: 1 - b^T A^{-1} e on the implicit part
    R = lambda z: 1 + z * b @ np.linalg.solve(np.eye(3) - z * A, e)
    assert abs(R(-1e4)) < 1e-3 and abs(R(-1e7)) < 1e-6 and abs(R(-1.0) - np.exp(-1.0)) < 2e-2


@pytest.mark.parametrize("alg", [_abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2])
def test_trbdf2_against_tight_reference(golden, alg):
    c = cases.CASES["robertson"]
    pb = make_problem("robertson", golden, 6)
    tight = c.opts(abstol=np.array([1e-10, 1e-12, 1e-10]), reltol=np.full(3, 1e-7), maxiters=10**7)
    for model in (pb["model"], pb["true_model"]):
        o = c.opts(alg=alg, pred_clamp=(-np.inf, np.inf))
        r = oracle.solve_batch(model, o, pb["u0"])
        ref = oracle.solve_batch(model, dataclasses_replace(tight, pred_clamp=(-np.inf, np.inf)), pb["u0"])
        assert (r["retcode"] == 1).all() and (r["n_saved"] == c.n_save).all()
        scale = np.abs(ref["pred"]).max(axis=(0, 1))
        assert (np.abs(r["pred"] - ref["pred"]) / scale).max() < 5e-3
        att = r["stats"]["n_accept"] + r["stats"]["n_reject"]
        if alg == _abi.ALG_TRBDF2:
            assert (r["stats"]["n_jac"] >= att).all()                     # one factorisation per attempt (+ refreshes)
        else:
            assert (r["stats"]["n_jac"] > 0).all() and (r["stats"]["n_jac"] < att).all()   # both halves ran
    # tighter tolerances converge towards the reference (second order: error ~ tol^(2/3)-ish, at least 10x better here)
    o1 = c.opts(alg=alg, abstol=np.array([1e-9, 1e-11, 1e-9]), reltol=np.full(3, 1e-6), pred_clamp=(-np.inf, np.inf))
    r1 = oracle.solve_batch(pb["true_model"], o1, pb["u0"])
    assert (np.abs(r1["pred"] - ref["pred"]) / scale).max() < 5e-5
    # a non-stiff model: the composite never leaves Tsit5 and equals it bit for bit
    p2 = make_problem("case2", golden, 4)
    a = oracle.solve_batch(p2["model"], p2["case"].opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2), p2["u0"])
    b = oracle.solve_batch(p2["model"], p2["case"].opts(alg=_abi.ALG_TSIT5), p2["u0"])
    assert np.array_equal(a["pred"], b["pred"]) and (a["stats"]["n_jac"] == 0).all()
    # forward sensitivities = duals through the Newton iterations: at tight tolerances the gradient is Rosenbrock23's,
    # at run tolerances it is within the solver tolerance of it
    args = (pb["seed"][:, :], pb["u0"][:3], pb["data"][:3], pb["yscale"], pb["loss_kind"])
    tight_o = dict(abstol=np.array([1e-12, 1e-14, 1e-12]), reltol=np.full(3, 1e-9), maxiters=10**7)
    g_ref = oracle.loss_grad_batch(pb["model"], c.opts(alg=_abi.ALG_ROSENBROCK23, **tight_o), *args)["grad_sum"]
    g_t = oracle.loss_grad_batch(pb["model"], c.opts(alg=alg, **tight_o), *args)
    assert (g_t["retcode"] == 1).all()
    assert np.linalg.norm(g_t["grad_sum"] - g_ref) / np.linalg.norm(g_ref) < 2e-6
    g_r = oracle.loss_grad_batch(pb["model"], c.opts(alg=alg), *args)
    assert np.linalg.norm(g_r["grad_sum"] - g_ref) / np.linalg.norm(g_ref) < 5e-3
    # partials in the norm change the step sequence (more attempts than the value-only solve), switched off they do not
    v = oracle.solve_batch(pb["model"], c.opts(alg=alg), pb["u0"][:3])
    g0 = oracle.loss_grad_batch(pb["model"], c.opts(alg=alg, err_norm_includes_sens=False), *args)
    assert np.array_equal(g0["stats"]["n_rhs"], v["stats"]["n_rhs"])


# ---------------------------------------------------------------- interpolating adjoint (BASELINE config 4; not in the reference)
@pytest.mark.parametrize("name", ["case2", "case3", "case1"])
def test_adjoint_gradient_matches_forward_mode_and_fd(golden, name):
    """Continuous adjoint vs the discrete forward-mode gradient: O(tolerance) apart at run tolerances,
    identical in the limit; and vs central finite differences of a tight solve."""
    pb = make_problem(name, golden, 3)
    c = pb["case"]
    data = np.abs(pb["data"]) + 1e-6 if name == "case3" else pb["data"]
    args = (pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    # case1's random weights with b0 = -10 give O(1e-3) gradients: the absolute tolerance shows
    for (at, rt), tol in (((1e-6, 1e-3), 1e-3), ((1e-11, 1e-9), 1e-5 if name == "case1" else 2e-8)):
        of = c.opts(obs_idx=np.arange(c.ns), abstol=at, reltol=rt)
        oa = c.opts(obs_idx=np.arange(c.ns), abstol=at, reltol=rt, sens_mode=_abi.SENS_INTERP_ADJOINT)
        rf = oracle.loss_grad_batch(pb["model"], of, *args)
        ra = oracle.loss_grad_batch(pb["model"], oa, *args, want_pred=True)
        np.testing.assert_allclose(ra["loss"], oracle.loss_grad_batch(pb["model"], c.opts(
            obs_idx=np.arange(c.ns), abstol=at, reltol=rt, err_norm_includes_sens=False), *args)["loss"], rtol=1e-12)
        assert np.linalg.norm(ra["grad_sum"] - rf["grad_sum"]) / np.linalg.norm(rf["grad_sum"]) < tol
        # the adjoint's forward pass is the plain value solve
        val = oracle.solve_batch(pb["model"], of, pb["u0"])
        assert np.array_equal(ra["stats"]["n_accept"], val["stats"]["n_accept"])
        np.testing.assert_allclose(ra["pred"], val["pred"], rtol=1e-12, atol=1e-15)
    if name == "case2":
        p = np.array(golden["case2"]["p"])
        ot = c.opts(obs_idx=np.arange(c.ns), abstol=1e-12, reltol=1e-10)
        L = lambda q: oracle.loss_grad_batch(c.model(q)[0], ot, *args)["loss"].sum()
        fd = np.array([(L(p + 1e-6 * e) - L(p - 1e-6 * e)) / 2e-6 for e in np.eye(len(p))])
        assert np.linalg.norm(ra["grad_sum"] - fd) / np.linalg.norm(fd) < 1e-6


def test_adjoint_truncation_and_missing_species(golden):
    pb = make_problem("case2", golden, 4, obs=np.array([0, 1, 3, 4, 5]))
    nsu = np.array([50, 7, 1, 30], dtype=np.int32)
    c = pb["case"]
    oa = c.opts(obs_idx=np.array([0, 1, 3, 4, 5]), sens_mode=_abi.SENS_INTERP_ADJOINT)
    of = c.opts(obs_idx=np.array([0, 1, 3, 4, 5]))
    args = (pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    ra = oracle.loss_grad_batch(pb["model"], oa, *args, n_save_used=nsu, want_grad_each=True)
    rf = oracle.loss_grad_batch(pb["model"], of, *args, n_save_used=nsu, want_grad_each=True)
    assert np.array_equal(ra["n_saved"], nsu)
    for i in range(4):
        gn = np.linalg.norm(rf["grad_each"][i])
        assert np.linalg.norm(ra["grad_each"][i] - rf["grad_each"][i]) <= 3e-4 * gn + 1e-12


@pytest.mark.parametrize("name", ["case2", "case3", "case1"])
def test_discrete_adjoint_equals_forward_mode(golden, name):
    """Reverse-mode through the recorded Tsit5 steps + dense output = the forward-mode (dual) gradient of the
    same discrete solve (value-only error norm), to rounding; truncation and missing species included."""
    obs = np.array([0, 1, 3, 4, 5]) if name == "case2" else None
    pb = make_problem(name, golden, 6, obs=obs)
    c = pb["case"]
    oi = np.arange(c.ns) if obs is None else obs
    data = np.abs(pb["data"]) + 1e-6 if name == "case3" else pb["data"]
    nsu = np.array([c.n_save, 3, 1, c.n_save // 2, c.n_save, 2], dtype=np.int32)
    of = c.opts(obs_idx=oi, err_norm_includes_sens=False)
    od = c.opts(obs_idx=oi, sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    args = (pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    rf = oracle.loss_grad_batch(pb["model"], of, *args, n_save_used=nsu, want_grad_each=True, want_pred=True)
    rd = oracle.loss_grad_batch(pb["model"], od, *args, n_save_used=nsu, want_grad_each=True, want_pred=True)
    np.testing.assert_allclose(rd["loss"], rf["loss"], rtol=1e-13)
    np.testing.assert_allclose(rd["pred"], rf["pred"], rtol=1e-13, atol=1e-300)
    assert np.array_equal(rd["stats"]["n_accept"], rf["stats"]["n_accept"])
    for i in range(6):
        gn = np.linalg.norm(rf["grad_each"][i])
        assert np.linalg.norm(rd["grad_each"][i] - rf["grad_each"][i]) <= 1e-11 * gn + 1e-300
