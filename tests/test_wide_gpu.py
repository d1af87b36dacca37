"""GPU parity tests of the GENERIC predict path (kernel_wide_solve.cuh: runtime dimensions, F0/F1/F2,
Tsit5 / Rosenbrock23 / AutoTsit5(Rosenbrock23)) against the CPU oracle, through the C-ABI.
Bar: step / RHS / Jacobian counts, retcodes and n_saved identical; saved states within the written tolerance."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases, synth
from crnn_b200.model import CRNNModel, SolveOpts
from oracle import oracle
from problems import make_problem

pytestmark = pytest.mark.gpu

ALG = {"tsit5": _abi.ALG_TSIT5, "ros23": _abi.ALG_ROSENBROCK23, "auto": _abi.ALG_AUTO_TSIT5_ROS23}
YS_HYCHEM = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])


def _counts_equal(got, ref):
    for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
        bad = np.nonzero(got["stats"][k] != ref["stats"][k])[0]
        assert bad.size == 0, f"{k} differs from the oracle for trajectories {bad[:8]}"
    assert np.array_equal(got["retcode"], ref["retcode"])
    assert np.array_equal(got["n_saved"], ref["n_saved"])


def _rel_err(got, ref):
    scale = np.abs(ref).max(axis=(0, 1)) + 1e-300
    return (np.abs(got - ref) / scale).max()


@pytest.mark.parametrize("name,alg,N,tol", [
    ("case2", "tsit5", 512, 1e-9), ("case2", "ros23", 256, 1e-9), ("case2", "auto", 512, 1e-9),
    ("case1", "auto", 64, 1e-9), ("robertson", "ros23", 256, 1e-5), ("robertson", "auto", 256, 1e-5)])
def test_generic_path_matches_oracle_on_the_reference_models(engine, golden, monkeypatch, name, alg, N, tol):
    """same models as the dimension-specialised kernels, forced through the generic kernel
    (robertson's trained CRNN, |w_out| up to 800, amplifies rounding: 1e-5)"""
    monkeypatch.setenv("CRNN_B200_FORCE_WIDE", "1")
    pb = make_problem(name, golden, N)
    o = pb["case"].opts(alg=ALG[alg], obs_idx=pb["opts"].obs_idx)
    got = engine.solve_batch(pb["model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["model"], o, pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < tol
    assert (got["retcode"] == _abi.RET_SUCCESS).all()


@pytest.mark.parametrize("name,N,tol", [("case2", 512, 1e-9), ("case1", 64, 1e-9), ("robertson", 256, 1e-5), ("case3", 96, 1e-5)])
def test_autoswitch_thread_per_trajectory_kernel_matches_oracle(engine, golden, name, N, tol):
    """AutoTsit5(Rosenbrock23()) on the dimension-specialised path (k_auto_value: the Tsit5 and Rosenbrock23 steps of the
    thread-per-trajectory kernels behind OrdinaryDiffEq's AutoSwitch counter) — case2.jl:26 as written"""
    pb = make_problem(name, golden, N)
    o = pb["case"].opts(alg=ALG["auto"], obs_idx=pb["opts"].obs_idx)
    got = engine.solve_batch(pb["model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["model"], o, pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < tol
    if name == "robertson":   # the trained stiff CRNN: both halves ran
        att = got["stats"]["n_accept"] + got["stats"]["n_reject"]
        assert (got["stats"]["n_jac"] > 0).all() and (got["stats"]["n_jac"] < att).all()
    else:
        assert (got["stats"]["n_jac"] == 0).all()


@pytest.mark.parametrize("alg", ["ros23", "auto"])
def test_autoswitch_on_the_true_robertson_mechanism(engine, golden, alg):
    """AutoTsit5(Rosenbrock23()) on the stiff generating mechanism (rober_crnn.jl:56-63): starts with Tsit5, detects
    stiffness, finishes with Rosenbrock23 — counts identical to the oracle, sum(y) conserved"""
    pb = make_problem("robertson", golden, 256)
    c = pb["case"]
    o = c.opts(alg=ALG[alg], pred_clamp=(-np.inf, np.inf))
    got = engine.solve_batch(pb["true_model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["true_model"], o, pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    att = got["stats"]["n_accept"] + got["stats"]["n_reject"]
    if alg == "auto":
        assert (got["stats"]["n_jac"] > 0).all() and (got["stats"]["n_jac"] < att).all()   # both halves ran
    mass = got["pred"].sum(axis=2)
    np.testing.assert_allclose(mass, pb["u0"].sum(axis=1)[:, None] * np.ones_like(mass), rtol=2e-3)


def test_dimensions_without_a_specialised_kernel(engine):
    """(n_species, n_reac) = (4, 5) has no template instantiation: solve_batch serves it from the generic kernel"""
    g = np.random.default_rng(3)
    ns, nr = 4, 5
    w_in = np.clip(g.normal(0.6, 0.6, (ns, nr)), 0.0, 2.0)
    # consumption only of a reaction's own reactants (rate -> 0 with them), production elsewhere: states stay positive
    w_out = -w_in * 10.0 ** g.normal(0.0, 0.2, (ns, nr)) + np.abs(g.normal(0.0, 0.3, (ns, nr))) * (w_in == 0.0)
    m = CRNNModel(w_in=w_in, w_b=g.normal(-1.0, 1.0, nr), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=1e-6, ub=10.0,
                  out_scale=np.array([1.0, 0.5, 2.0, 1.0]))
    u0 = 0.2 + g.random((200, ns))
    for alg in ALG.values():
        o = SolveOpts(saveat=np.linspace(0.0, 5.0, 30), t0=0.0, t1=5.0, alg=alg, abstol=1e-7, reltol=1e-4)
        got = engine.solve_batch(m, o, u0)
        ref = oracle.solve_batch(m, o, u0, n_threads=8)
        _counts_equal(got, ref)
        # a trajectory whose species cross the lb clamp (a kink of the RHS) amplifies rounding: two CPU builds of
        # the oracle (with / without FMA contraction) differ by 3e-9 there and by 1e-14 elsewhere
        smooth = ref["pred"].min(axis=(1, 2)) > 1e-3
        assert smooth.sum() > 100
        assert _rel_err(got["pred"][smooth], ref["pred"][smooth]) < 1e-11
        assert _rel_err(got["pred"], ref["pred"]) < 1e-6


@pytest.mark.parametrize("alg", ["tsit5", "ros23", "auto"])
def test_hychem_f2_matches_oracle(engine, alg):
    """HyChem/crnn_pyrolysis_mass.jl's RHS (mass fractions, density coupling, tabulated T(t), P(t), non-autonomous
    Rosenbrock23 with df/dt) with the script's own initialisation p = 0.1 randn, slope 0.1"""
    N = 384
    m, _ = cases.hychem_model(cases.hychem_p(0), YS_HYCHEM)
    o = cases.hychem_opts(alg=ALG[alg])
    u0 = cases.hychem_u0(N)
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    if alg == "auto":   # this problem is mildly stiff (the N2 row is scaled by 90/t_end): the switch happens
        att = got["stats"]["n_accept"] + got["stats"]["n_reject"]
        assert (got["stats"]["n_jac"] > 0).all() and (got["stats"]["n_jac"] < att).all()


def test_hychem_truncation_device_buffers_and_maxiters(engine):
    """random time truncation `sample = rand(batch_size:ntotal)` (crnn_pyrolysis_mass.jl:199), torch device
    buffers, and the maxiters retcode on the generic path"""
    import torch
    N = 200
    m, _ = cases.hychem_model(cases.hychem_p(1), YS_HYCHEM)
    o = cases.hychem_opts(alg=ALG["auto"])
    u0 = cases.hychem_u0(N, seed=7)
    nsu = np.random.default_rng(0).integers(32, 41, N).astype(np.int32)
    got = engine.solve_batch(m, o, u0, n_save_used=nsu)
    ref = oracle.solve_batch(m, o, u0, n_save_used=nsu, n_threads=8)
    _counts_equal(got, ref)
    assert np.array_equal(got["n_saved"], nsu)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    dev = engine.solve_batch(m, o, torch.from_numpy(u0).cuda(), n_save_used=torch.from_numpy(nsu).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dev["pred"].cpu().numpy(), got["pred"])
    o2 = cases.hychem_opts(alg=ALG["auto"], maxiters=12)
    got2 = engine.solve_batch(m, o2, u0)
    ref2 = oracle.solve_batch(m, o2, u0, n_threads=8)
    _counts_equal(got2, ref2)
    assert (got2["retcode"] == _abi.RET_MAXITERS).all() and (got2["n_saved"] < o2.n_save).all()
    for i in range(0, N, 37):
        k = got2["n_saved"][i]
        np.testing.assert_allclose(got2["pred"][i, :k], ref2["pred"][i, :k], rtol=1e-8, atol=1e-14)


def test_generic_path_full_size_properties(engine, golden, monkeypatch):
    """65 536 case2 trajectories through the generic kernel: equal to the specialised kernel's saved states
    (two different CUDA formulations of the same algorithm) and every trajectory successful"""
    pb = make_problem("case2", golden, 64)
    u0 = synth.make_u0("case2", 65536)
    fast = engine.solve_batch(pb["model"], pb["opts"], u0)
    monkeypatch.setenv("CRNN_B200_FORCE_WIDE", "1")
    wide = engine.solve_batch(pb["model"], pb["opts"], u0)
    assert (wide["retcode"] == _abi.RET_SUCCESS).all()
    for k in ("n_accept", "n_reject", "n_rhs"):
        assert np.array_equal(wide["stats"][k], fast["stats"][k])
    assert _rel_err(wide["pred"], fast["pred"]) < 1e-9


@pytest.mark.parametrize("mode", ["discrete", "interp"])
def test_hychem_f2_gradient_np211_by_the_adjoints(engine, mode):
    """loss + gradient of the HyChem model (np = 211, crnn_pyrolysis_mass.jl:143-147,201) on the GPU by the adjoint
    kernels (cost independent of np), against the oracle's adjoint of the same kind and against the oracle's
    FORWARD-mode gradient (211 dual columns through Tsit5 — what ForwardDiff.gradient computes)"""
    N = 96
    # the NON-stiff variant (every reaction two e-folds slower): with the script's own initialisation explicit Tsit5
    # runs at its stability limit, where accept/reject decisions sit within rounding of EEst = 1 (no step-count parity)
    # and value-only-controlled tangents are unstable.  Seeds 0 / 1: random CRNN weights whose trajectories do not sit on
    # a clamp kink (for most other seeds two CPU builds of the oracle, with / without FMA contraction, differ by 1e-6 .. 0.5)
    m, seed = cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS_HYCHEM)
    u0 = cases.hychem_u0(N)
    data = oracle.solve_batch(cases.hychem_model(cases.hychem_p(1, lnA_shift=-2.0), YS_HYCHEM)[0],
                              cases.hychem_opts(alg=ALG["ros23"]), u0, n_threads=8)["pred"]
    sm = _abi.SENS_DISCRETE_ADJOINT if mode == "discrete" else _abi.SENS_INTERP_ADJOINT
    tol = {}
    o = cases.hychem_opts(alg=ALG["tsit5"], sens_mode=sm, maxiters=100000, **tol)
    got = engine.loss_grad_batch(m, o, seed, u0, data, YS_HYCHEM, want_pred=True)
    ref = oracle.loss_grad_batch(m, o, seed, u0, data, YS_HYCHEM, want_pred=True, n_threads=8)
    assert np.array_equal(got["retcode"], ref["retcode"]) and (got["retcode"] == _abi.RET_SUCCESS).all()
    for k in ("n_accept", "n_reject"):
        assert np.array_equal(got["stats"][k], ref["stats"][k])
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-9)
    gmax = np.abs(ref["grad_sum"]).max()
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=1e-6, atol=1e-8 * gmax)
    if mode == "discrete":
        # the discrete adjoint IS the forward-mode derivative of the same step sequence (value-only error norm)
        of = cases.hychem_opts(alg=ALG["tsit5"], sens_mode=_abi.SENS_FORWARD, err_norm_includes_sens=False, maxiters=100000, **tol)
        fwd = oracle.loss_grad_batch(m, of, seed, u0[:16], data[:16], YS_HYCHEM, n_threads=8)
        g16 = engine.loss_grad_batch(m, o, seed, u0[:16], data[:16], YS_HYCHEM)
        np.testing.assert_allclose(g16["grad_sum"], fwd["grad_sum"], rtol=1e-6, atol=1e-8 * np.abs(fwd["grad_sum"]).max())


def test_kencarp4_on_the_hychem_f2_model(engine):
    """BASELINE config 5 as named: the HyChem pyrolysis RHS (F2, non-autonomous: the implicit stages are evaluated at
    t + c_i dt) integrated by KenCarp4.  Newton iteration counts may flip on rounding for a few trajectories."""
    N = 256
    m, _ = cases.hychem_model(cases.hychem_p(0), YS_HYCHEM)
    o = cases.hychem_opts(alg=_abi.ALG_KENCARP4)
    u0 = cases.hychem_u0(N)
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all() and np.array_equal(got["n_saved"], ref["n_saved"])
    same = (got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]) & (got["stats"]["n_accept"] == ref["stats"]["n_accept"]) & \
           (got["stats"]["n_reject"] == ref["stats"]["n_reject"])
    assert same.mean() >= 0.95, f"{(~same).sum()} of {N} trajectories differ in step/RHS counts"
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    err = np.abs(got["pred"] - ref["pred"]) / scale
    assert err[same].max() < 1e-7 and err.max() < 5e-3


def _trb_close(got, ref, frac=0.02, tol_same=1e-7):
    """TRBDF2's simplified-Newton iteration counts can flip on last-bit differences (like KenCarp4's): all but a few
    trajectories must have identical counts, and those agree to rounding; the rest still solve the ODE to tolerance"""
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])
    same = np.ones(got["retcode"].shape, dtype=bool)
    for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
        same &= got["stats"][k] == ref["stats"][k]
    assert same.mean() >= 1.0 - frac, f"{(~same).sum()} of {same.size} trajectories differ in step / RHS / Jacobian counts"
    scale = np.abs(ref["pred"]).max(axis=(0, 1)) + 1e-300
    err = np.abs(got["pred"] - ref["pred"]) / scale
    assert err[same].max() < tol_same and err.max() < 5e-3
    return same


@pytest.mark.parametrize("alg", [_abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2])
def test_trbdf2_and_its_composite_match_the_oracle(engine, golden, alg):
    """TRBDF2 as an ESDIRK and AutoTsit5(TRBDF2()) (Cathode/src/network.jl:102, yeast_glycolysis.jl:33) on the stiff models:
    the true Robertson mechanism, the trained stiff CRNN, the HyChem F2 model (non-autonomous: stage times matter)"""
    pb = make_problem("robertson", golden, 256)
    c = pb["case"]
    o = c.opts(alg=alg, pred_clamp=(-np.inf, np.inf))
    for model, tol in ((pb["true_model"], 1e-7), (pb["model"], 1e-4)):
        got = engine.solve_batch(model, o, pb["u0"])
        ref = oracle.solve_batch(model, o, pb["u0"], n_threads=8)
        assert (got["retcode"] == _abi.RET_SUCCESS).all()
        _trb_close(got, ref, tol_same=tol)
        att = got["stats"]["n_accept"] + got["stats"]["n_reject"]
        if alg == _abi.ALG_AUTO_TSIT5_TRBDF2:
            assert (got["stats"]["n_jac"] > 0).all() and (got["stats"]["n_jac"] < att).all()   # both halves ran
    m, _ = cases.hychem_model(cases.hychem_p(0), YS_HYCHEM)
    oh = cases.hychem_opts(alg=alg)
    u0 = cases.hychem_u0(256)
    got = engine.solve_batch(m, oh, u0)
    ref = oracle.solve_batch(m, oh, u0, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    _trb_close(got, ref, frac=0.05)
    # a non-stiff model: the composite never leaves Tsit5 and is the Tsit5 solve of the same kernel bit for bit
    p2 = make_problem("case2", golden, 128)
    a = engine.solve_batch(p2["model"], p2["case"].opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2), p2["u0"])
    b = oracle.solve_batch(p2["model"], p2["case"].opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2), p2["u0"], n_threads=8)
    _counts_equal(a, b)
    assert (a["stats"]["n_jac"] == 0).all() and _rel_err(a["pred"], b["pred"]) < 1e-9
    # (gradients through TRBDF2: tests/test_gen_sens_gpu.py; the HyChem flavour keeps Rosenbrock23 and is refused loudly)
    from crnn_b200.engine import EngineError
    _, sh = cases.hychem_model(cases.hychem_p(0), YS_HYCHEM)
    with pytest.raises(EngineError):
        engine.loss_grad_batch(m, oh, sh, u0[:4], ref["pred"][:4], YS_HYCHEM)


def test_cathode_predict_as_written_autotsit5_trbdf2_with_heat_release(engine):
    """pred_n_ode of Cathode/src/network.jl:104-131: AutoTsit5(TRBDF2) on the F5 model under the five temperature programmes,
    the saved states mapped to the heat release HRR_getter(ts, sol) * w_delH inside the predict kernel"""
    import cathode_problem as cp
    pb = cp.make(2, seed=1)
    for alg in (_abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_ROSENBROCK23):
        o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
        for e, beta in enumerate(cp.BETAS):
            m, _ = cp.model_for(pb["particles"][0], beta, pb["t_hi"])
            u0 = np.tile(pb["u0"][e], (8, 1)) * (1.0 - 0.01 * np.arange(8))[:, None]
            got = engine.solve_batch(m, o, u0)
            ref = oracle.solve_batch(m, o, u0, n_threads=8)
            assert got["pred"].shape == (8, cp.N_SAVE, 1) and (got["retcode"] == _abi.RET_SUCCESS).all()
            if alg in (_abi.ALG_TRBDF2, _abi.ALG_ROSENBROCK23):
                _counts_equal(got, ref)             # the stiff steppers alone: identical counts (TRBDF2's Newton counts too)
                assert _rel_err(got["pred"], ref["pred"]) < 1e-8
            else:
                # the composites spend the ramp's stiffening phase with Tsit5 AT its stability limit (up to 20 % of the
                # attempts rejected): accept / reject decisions there flip on summation order, for either stiff half
                assert np.array_equal(got["n_saved"], ref["n_saved"])
                assert _rel_err(got["pred"], ref["pred"]) < 2e-2
                assert abs(got["stats"]["n_rhs"].sum() / ref["stats"]["n_rhs"].sum() - 1.0) < 0.03
                att = got["stats"]["n_accept"] + got["stats"]["n_reject"]
                assert (got["stats"]["n_jac"] <= att).all()
                if beta <= 2.0:                      # slow ramp, never near the limit: exact
                    _counts_equal(got, ref)


def test_reversible_crnn_case1_rev_on_the_generic_path(engine):
    """`case1 rev/case1.jl`: 5 species, 10 reversible reactions = an F0 CRNN with 20 reactions, np = 60.  No specialised
    kernel has these dimensions: predict runs on the generic kernel, the gradient on the discrete adjoint, which
    equals the oracle's forward-mode (ForwardDiff-semantics) gradient with the value-only error norm."""
    c = cases.CASES["case1_rev"]
    g = np.random.default_rng(11)
    N = 128
    u0 = g.random((N, 5)); u0[:, :2] += 0.2
    data = oracle.solve_batch(cases.true_model_case1_rev(), c.opts(), u0, n_threads=8)["pred"]
    ys = synth.yscale_from(data, c.lb)
    m, seed = c.model(g.standard_normal(c.n_p) * 0.5)
    got = engine.solve_batch(m, c.opts(), u0)
    ref = oracle.solve_batch(m, c.opts(), u0, n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-9
    o = c.opts(sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    ga = engine.loss_grad_batch(m, o, seed, u0, data, ys, c.loss_kind)
    of = c.opts(sens_mode=_abi.SENS_FORWARD, err_norm_includes_sens=False)
    fwd = oracle.loss_grad_batch(m, of, seed, u0, data, ys, c.loss_kind, n_threads=8)
    np.testing.assert_allclose(ga["loss"], fwd["loss"], rtol=1e-9)
    np.testing.assert_allclose(ga["grad_sum"], fwd["grad_sum"], rtol=1e-6, atol=1e-8 * np.abs(fwd["grad_sum"]).max())
    # FORWARD with the value-only norm on dimensions without a forward kernel is served by the same discrete adjoint ...
    gf = engine.loss_grad_batch(m, of, seed, u0, data, ys, c.loss_kind)
    np.testing.assert_array_equal(gf["grad_sum"], ga["grad_sum"])
    # ... and with the partials in the norm (a different step sequence) it is refused, not approximated
    from crnn_b200.engine import EngineError
    with pytest.raises(EngineError):
        engine.loss_grad_batch(m, c.opts(sens_mode=_abi.SENS_FORWARD), seed, u0, data, ys, c.loss_kind)


def test_generic_path_edge_cases_empty_single_nan_weights(engine):
    """N = 0, N = 1, and a model whose weights contain NaN (retcode, no abort) on the generic kernels"""
    m, seed = cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS_HYCHEM)
    for alg in ALG.values():
        o = cases.hychem_opts(alg=alg)
        r0 = engine.solve_batch(m, o, np.zeros((0, 9)))
        assert r0["pred"].shape == (0, 40, 9) and r0["retcode"].shape == (0,)
        u1 = cases.hychem_u0(1)
        r1 = engine.solve_batch(m, o, u1)
        ref1 = oracle.solve_batch(m, o, u1)
        _counts_equal(r1, ref1)
        assert _rel_err(r1["pred"], ref1["pred"]) < 1e-9
    data = ref1["pred"] * 1.02
    for sm in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT):
        o = cases.hychem_opts(alg=ALG["tsit5"], sens_mode=sm)
        g0 = engine.loss_grad_batch(m, o, seed, np.zeros((0, 9)), np.zeros((0, 40, 9)), YS_HYCHEM)
        assert g0["loss"].shape == (0,) and np.all(g0["grad_sum"] == 0.0)
        g1 = engine.loss_grad_batch(m, o, seed, u1, data, YS_HYCHEM)
        gr = oracle.loss_grad_batch(m, o, seed, u1, data, YS_HYCHEM)
        np.testing.assert_allclose(g1["loss"], gr["loss"], rtol=1e-9)
        np.testing.assert_allclose(g1["grad_sum"], gr["grad_sum"], rtol=1e-6, atol=1e-9 * np.abs(gr["grad_sum"]).max())
    # NaN in the weights: every trajectory reports a failure code, the batch still returns (rober_crnn.jl:130-134 prints and goes on)
    import copy
    mb = copy.deepcopy(m); mb.w_b = mb.w_b.copy(); mb.w_b[3] = np.nan
    u8 = cases.hychem_u0(8)
    for alg in ALG.values():
        o = cases.hychem_opts(alg=alg)
        rb = engine.solve_batch(mb, o, u8)
        rr = oracle.solve_batch(mb, o, u8)
        assert np.array_equal(rb["retcode"], rr["retcode"]) and (rb["retcode"] != _abi.RET_SUCCESS).all()
        assert np.array_equal(rb["n_saved"], rr["n_saved"])


@pytest.mark.parametrize("force_wide", [False, True])
def test_autoswitch_back_to_tsit5_when_the_stiffness_goes_away(engine, monkeypatch, force_wide):
    """A -> B slow, B -> C fast: B sits in quasi-steady state (stiff: Tsit5 alone needs 13 000 steps, the composite hands
    over to Rosenbrock23) until it falls below the lb clamp, where its Jacobian column vanishes and the composite must
    switch BACK to Tsit5 (dt /= 2, controller exponents of order 5 again).  Both GPU formulations against the oracle."""
    if force_wide:
        monkeypatch.setenv("CRNN_B200_FORCE_WIDE", "1")
    w_out = np.zeros((3, 3)); w_out[:, 0] = [-1, 1, 0]; w_out[:, 1] = [0, -1, 1]
    m = CRNNModel(w_in=np.eye(3), w_b=np.log([1.0, 1e4, 1e-30]), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=1e-6, ub=np.inf)
    g = np.random.default_rng(0)
    u0 = np.c_[0.5 + g.random(192), 1e-4 * g.random(192), 0.1 * g.random(192)]
    o = SolveOpts(saveat=np.linspace(0.0, 10.0, 26), t0=0.0, t1=10.0, alg=ALG["auto"], abstol=1e-8, reltol=1e-4, maxiters=200000)
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    tsit5_attempts = got["stats"]["n_accept"] + got["stats"]["n_reject"] - got["stats"]["n_jac"]
    assert (got["stats"]["n_jac"] > 20).all() and (tsit5_attempts > 25).all()   # > the ~12 attempts before the first switch


def test_adjoint_on_a_large_model_keeps_two_blocks_per_sm(engine):
    """n_w = 496 of the 512 the adjoint kernel supports (15 species, 16 reactions): the quadrature accumulators leave a
    four-block build no room for the step record, the host falls back to two blocks per SM; gradient vs the oracle"""
    g = np.random.default_rng(8)
    ns, nr = 15, 16
    w_in = np.clip(g.normal(0.4, 0.5, (ns, nr)), 0.0, 2.0) * (g.random((ns, nr)) < 0.3)
    w_out = -w_in * 10.0 ** g.normal(0.0, 0.2, (ns, nr)) + np.abs(g.normal(0.0, 0.3, (ns, nr))) * (w_in == 0.0) * (g.random((ns, nr)) < 0.2)
    m = CRNNModel(w_in=w_in, w_b=g.normal(-2.5, 0.5, nr), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=1e-6, ub=10.0)
    assert m.n_w == 496
    u0 = 0.5 + g.random((64, ns))
    n_p = 40
    seed = np.zeros((m.n_w, n_p)); seed[g.choice(m.n_w, n_p, replace=False), np.arange(n_p)] = 1.0   # 40 of the weights are trainable
    o = SolveOpts(saveat=np.linspace(0.0, 2.0, 21), t0=0.0, t1=2.0, abstol=1e-7, reltol=1e-4, sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    data = oracle.solve_batch(m, o, u0, n_threads=8)["pred"] * (1.0 + 0.05 * g.standard_normal((64, 21, ns)))
    ys = np.ones(ns)
    got = engine.loss_grad_batch(m, o, seed, u0, data, ys)
    of = SolveOpts(saveat=o.saveat, t0=0.0, t1=2.0, abstol=1e-7, reltol=1e-4, sens_mode=_abi.SENS_FORWARD, err_norm_includes_sens=False)
    ref = oracle.loss_grad_batch(m, of, seed, u0, data, ys, n_threads=8)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-9)
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=1e-6, atol=1e-8 * np.abs(ref["grad_sum"]).max())


def test_gene_regulatory_checkpoint_on_the_gpu(engine, golden):
    """gene-regulatory.jl (9 species, 15 reactions, np = 285 > the forward kernel's 255 columns): the committed checkpoint
    on the dimension-specialised predict kernel, its gradient by the discrete adjoint vs the oracle's forward mode"""
    c = cases.CASES["gene"]
    m, seed = c.model(np.array(golden["gene"]["p"]))
    u0 = np.random.default_rng(0).random((160, 9))
    got = engine.solve_batch(m, c.opts(), u0)
    ref = oracle.solve_batch(m, c.opts(), u0, n_threads=8)
    _counts_equal(got, ref)
    assert _rel_err(got["pred"], ref["pred"]) < 1e-8
    truth = oracle.solve_batch(cases.true_model_gene(), c.opts(pred_clamp=(-np.inf, np.inf)), u0, n_threads=8)["pred"]
    data = np.clip(truth * (1.0 + 0.01 * np.random.default_rng(2).standard_normal(truth.shape)), c.lb, c.ub)
    ys = np.ones(9)
    ga = engine.loss_grad_batch(m, c.opts(), seed, u0, data, ys, c.loss_kind)
    assert 1e-3 < ga["loss"].mean() < 1.5e-2                          # the checkpoint's loss level on 1 %-noisy data
    of = c.opts(sens_mode=_abi.SENS_FORWARD, err_norm_includes_sens=False)
    fwd = oracle.loss_grad_batch(m, of, seed, u0[:48], data[:48], ys, c.loss_kind, n_threads=8)
    g48 = engine.loss_grad_batch(m, c.opts(), seed, u0[:48], data[:48], ys, c.loss_kind)
    np.testing.assert_allclose(g48["loss"], fwd["loss"], rtol=1e-9)
    np.testing.assert_allclose(g48["grad_sum"], fwd["grad_sum"], rtol=1e-6, atol=1e-8 * np.abs(fwd["grad_sum"]).max())
