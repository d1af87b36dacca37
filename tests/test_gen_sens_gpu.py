"""The generic block-per-trajectory forward-sensitivity kernel (kernel_gen_sens.cuh) against the oracle's forward mode:
gradients through AutoTsit5(Rosenbrock23) (what case2.jl:26,195 and crnn_pyrolysis_mass.jl:29,201 run), stiff gradients
of the F2 model and of models with more than 6 species, and every reference model on the generic path."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases
from oracle import oracle
from problems import make_problem

pytestmark = pytest.mark.gpu
ALG = {"tsit5": _abi.ALG_TSIT5, "ros23": _abi.ALG_ROSENBROCK23, "auto": _abi.ALG_AUTO_TSIT5_ROS23}
YS_HYCHEM = np.array([0.05, 0.02, 0.01, 0.02, 0.01, 0.02, 0.01, 0.01, 0.9])


def _compare(got, ref, rtol_state=1e-9, rtol_loss=1e-10, rtol_grad=1e-8, counts=True):
    if counts:
        for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
            bad = np.nonzero(got["stats"][k] != ref["stats"][k])[0]
            assert bad.size == 0, f"{k} differs from the oracle for trajectories {bad[:8]}"
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])
    scale = np.abs(ref["pred"]).max(axis=(0, 1), keepdims=True)
    assert (np.abs(got["pred"] - ref["pred"]) <= rtol_state * np.abs(ref["pred"]) + 1e-12 * scale).all()
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=rtol_loss)
    gmax = np.abs(ref["grad_sum"]).max()
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=rtol_grad, atol=1e-9 * gmax)


@pytest.mark.parametrize("name,alg,N", [("case2", "tsit5", 96), ("case1", "tsit5", 48), ("robertson", "ros23", 64),
                                        ("case2", "ros23", 64), ("case2", "auto", 96), ("robertson", "auto", 64)])
def test_generic_kernel_on_the_reference_models(engine, golden, monkeypatch, name, alg, N):
    """same models as the dimension-specialised kernels serve, forced onto the generic one; partials in the error norm"""
    monkeypatch.setenv("CRNN_B200_FORCE_GENERIC", "1")
    pb = make_problem(name, golden, N)
    c = pb["case"]
    o = c.opts(obs_idx=np.arange(c.ns), alg=ALG[alg])
    args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    stiff = alg != "tsit5"
    _compare(got, ref, rtol_state=1e-7 if stiff else 1e-9, rtol_loss=1e-7 if stiff else 1e-10, rtol_grad=1e-6 if stiff else 1e-8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    if name == "case2" and alg == "auto":      # never stiff: the composite is bit-identical to Tsit5 (case2.jl:26 as written)
        monkeypatch.delenv("CRNN_B200_FORCE_GENERIC")
        fast = engine.loss_grad_batch(pb["model"], c.opts(obs_idx=np.arange(c.ns)), *args[2:], want_pred=True)
        assert np.array_equal(fast["stats"]["n_accept"], got["stats"]["n_accept"])
        np.testing.assert_allclose(got["grad_sum"], fast["grad_sum"], rtol=1e-10)
    if name == "robertson" and alg == "auto":
        assert (got["stats"]["n_jac"] > 0).any()          # the composite did switch


def test_generic_kernel_norm_switches_and_truncation(engine, golden, monkeypatch):
    monkeypatch.setenv("CRNN_B200_FORCE_GENERIC", "1")
    pb = make_problem("case2", golden, 64, obs=np.array([0, 1, 3, 4, 5]))
    nsu = np.random.default_rng(0).integers(1, 51, size=64).astype(np.int32)
    for kw in (dict(err_norm_includes_sens=False), dict(err_norm_mean_over_partials=False), dict(alg=ALG["auto"])):
        o = pb["case"].opts(obs_idx=np.array([0, 1, 3, 4, 5]), **kw)
        args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
        got = engine.loss_grad_batch(*args, n_save_used=nsu, want_pred=True)
        ref = oracle.loss_grad_batch(*args, n_save_used=nsu, want_pred=True, n_threads=8)
        _compare(got, ref)
        assert np.array_equal(got["n_saved"], nsu)


def test_case3_np153_and_stiff_gradient_beyond_six_species(engine, golden, monkeypatch):
    """Rosenbrock23 forward sensitivities for n_species = 9 (the register-LU kernel stops at 6) and the 153-column
    Tsit5 gradient on the generic kernel; near-true weights (see tests/test_full_size_gpu.py)"""
    from test_full_batch_parity_gpu import case3_near_true_model
    c = cases.CASES["case3"]
    model, seed = case3_near_true_model(c)
    pb = make_problem("case3", golden, 48)
    data = np.abs(pb["data"]) + 1e-6
    for alg in ("ros23", "auto", "tsit5"):
        if alg == "tsit5":
            monkeypatch.setenv("CRNN_B200_FORCE_GENERIC", "1")
        o = c.opts(obs_idx=np.arange(c.ns), alg=ALG[alg])
        args = (model, o, seed, pb["u0"], data, pb["yscale"], c.loss_kind)
        got = engine.loss_grad_batch(*args, want_pred=True)
        ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
        _compare(got, ref, rtol_state=1e-6, rtol_loss=1e-6, rtol_grad=1e-5)


@pytest.mark.parametrize("alg", ["tsit5", "ros23", "auto"])
def test_hychem_f2_forward_gradient_np211(engine, alg):
    """ForwardDiff.gradient through the solver on the HyChem model (crnn_pyrolysis_mass.jl:143-147,201): 211 dual columns,
    the non-autonomous density-coupled F2 RHS, Tsit5 / Rosenbrock23 / the composite the script names"""
    N = 48
    stiff = 0.0 if alg == "tsit5" else 4.0
    kw = dict(lnA_shift=-2.0) if alg == "tsit5" else dict(stiff=stiff)
    m, seed = cases.hychem_model(cases.hychem_p(0, **kw), YS_HYCHEM)
    u0 = cases.hychem_u0(N)
    data = oracle.solve_batch(cases.hychem_model(cases.hychem_p(1, **kw), YS_HYCHEM)[0],
                              cases.hychem_opts(alg=ALG["ros23"]), u0, n_threads=8)["pred"]
    o = cases.hychem_opts(alg=ALG[alg], maxiters=100000)
    got = engine.loss_grad_batch(m, o, seed, u0, data, YS_HYCHEM, want_pred=True)
    ref = oracle.loss_grad_batch(m, o, seed, u0, data, YS_HYCHEM, want_pred=True, n_threads=8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    _compare(got, ref, rtol_state=1e-6, rtol_loss=1e-7, rtol_grad=1e-5)
    if alg == "auto":
        assert (got["stats"]["n_jac"] > 0).any()


def test_indexed_dataset_call_on_the_generic_kernel(engine, golden):
    pb = make_problem("case2", golden, 150)
    o = pb["case"].opts(obs_idx=np.arange(6), alg=ALG["auto"])
    ds = engine.dataset(pb["u0"], pb["data"])
    idx = np.random.default_rng(3).permutation(150)[:40]
    a = engine.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"])
    b = engine.loss_grad_indexed(pb["model"], o, pb["seed"], ds, pb["yscale"], pb["loss_kind"], idx=idx, want_loss=True)
    assert np.array_equal(a["loss"], b["loss"]) and np.array_equal(a["grad_sum"], b["grad_sum"])
    ds.close()


@pytest.mark.parametrize("name,N", [("case2", 512), ("robertson", 96), ("case1", 64)])
def test_autotsit5_fast_path_specialised_kernel_plus_hand_over(engine, golden, name, N):
    """AutoTsit5(Rosenbrock23) without forcing the generic kernel: the specialised Tsit5 kernel carries the AutoSwitch
    monitor and hands the trajectories that would switch over to the composite kernel.  Same results as the oracle's
    composite either way: case2 / case1 never switch (case2.jl:26 as written), the stiff Robertson CRNN always does."""
    pb = make_problem(name, golden, N)
    c = pb["case"]
    o = c.opts(obs_idx=np.arange(c.ns), alg=ALG["auto"])
    args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    stiff = name == "robertson"
    _compare(got, ref, rtol_state=1e-7 if stiff else 1e-9, rtol_loss=1e-7 if stiff else 1e-10, rtol_grad=1e-6 if stiff else 1e-8)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    if stiff:
        assert (got["stats"]["n_jac"] > 0).all()
    else:
        assert (got["stats"]["n_jac"] == 0).all()
    # ragged + indexed dataset call through the same path
    ds = engine.dataset(pb["u0"], pb["data"])
    idx = np.random.default_rng(5).permutation(N)[:N // 3]
    nsu = np.random.default_rng(6).integers(1, o.n_save + 1, size=idx.size).astype(np.int32)
    a = engine.loss_grad_indexed(pb["model"], o, pb["seed"], ds, pb["yscale"], pb["loss_kind"], idx=idx, n_save_used=nsu, want_loss=True)
    b = oracle.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"], n_save_used=nsu, n_threads=8)
    np.testing.assert_allclose(a["loss"], b["loss"], rtol=1e-7 if stiff else 1e-10)
    np.testing.assert_allclose(a["grad_sum"], b["grad_sum"], rtol=1e-6 if stiff else 1e-8, atol=1e-9 * np.abs(b["grad_sum"]).max())
    ds.close()


def test_autotsit5_fast_path_mixed_batch(engine, golden):
    """a batch in which SOME trajectories switch: the true Robertson mechanism's rates scaled down for half of the batch is
    not expressible with shared weights, so mix by horizon instead - short horizons never reach the switch, long ones do"""
    pb = make_problem("robertson", golden, 64)
    c = pb["case"]
    o = c.opts(obs_idx=np.arange(c.ns), alg=ALG["auto"])
    nsu = np.where(np.arange(64) % 2 == 0, 2, 40).astype(np.int32)      # tspan = [0, tsteps[sample]] (rober_crnn.jl:125)
    args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, n_save_used=nsu, want_pred=True)
    ref = oracle.loss_grad_batch(*args, n_save_used=nsu, want_pred=True, n_threads=8)
    _compare(got, ref, rtol_state=1e-7, rtol_loss=1e-7, rtol_grad=1e-6)
    assert np.array_equal(got["n_saved"], nsu)


@pytest.mark.parametrize("alg", [_abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2])
def test_gradients_through_trbdf2_and_its_composite(engine, golden, alg):
    """ForwardDiff through AutoTsit5(TRBDF2) (the Cathode scripts' training path, Cathode/src/network.jl:102 +
    Cathode_NCM333_UQ/src_333/network.jl:232): dual columns through the simplified-Newton iterations, against the oracle"""
    for name, N in (("robertson", 64), ("case2", 48)):
        pb = make_problem(name, golden, N)
        c = pb["case"]
        o = c.opts(obs_idx=np.arange(c.ns), alg=alg)
        args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
        got = engine.loss_grad_batch(*args, want_pred=True)
        ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
        assert (got["retcode"] == _abi.RET_SUCCESS).all()
        if name == "case2" and alg == _abi.ALG_AUTO_TSIT5_TRBDF2:
            _compare(got, ref)                                    # never stiff: Tsit5, exact counts
            assert (got["stats"]["n_jac"] == 0).all()
            continue
        # Newton iteration counts can flip on summation order for a few trajectories; the rest agree to rounding
        same = np.ones(N, dtype=bool)
        for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
            same &= got["stats"][k] == ref["stats"][k]
        assert same.mean() >= 0.9, f"{(~same).sum()} of {N} trajectories differ in counts"
        np.testing.assert_allclose(got["loss"][same], ref["loss"][same], rtol=1e-6)
        scale = np.abs(ref["pred"]).max(axis=(0, 1))
        assert (np.abs(got["pred"] - ref["pred"])[same] / scale).max() < 1e-6
        assert np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"]) < (1e-6 if same.all() else 2e-3)
        if alg == _abi.ALG_AUTO_TSIT5_TRBDF2 and name == "robertson":
            assert (got["stats"]["n_jac"] > 0).any()
    # and against Rosenbrock23's gradient at tight tolerances (two different discretisations of the same sensitivities)
    pb = make_problem("robertson", golden, 8)
    c = pb["case"]
    tight = dict(abstol=np.array([1e-11, 1e-13, 1e-11]), reltol=np.full(3, 1e-8), maxiters=10 ** 7)
    args = (pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    g_t = engine.loss_grad_batch(pb["model"], c.opts(alg=alg, **tight), *args)
    g_r = engine.loss_grad_batch(pb["model"], c.opts(alg=_abi.ALG_ROSENBROCK23, **tight), *args)
    assert np.linalg.norm(g_t["grad_sum"] - g_r["grad_sum"]) / np.linalg.norm(g_r["grad_sum"]) < 1e-5


def test_cathode_training_path_as_written_trbdf2_gradients(engine):
    """loss + gradient of the heat-release MSE through AutoTsit5(TRBDF2) on the F5 model, batched and particle-batched
    (src_333/network.jl:222-260 with the script's own algorithm)"""
    import cathode_problem as cp
    pb = cp.make(4, seed=1, alg=_abi.ALG_AUTO_TSIT5_TRBDF2)
    assert pb["opts"].alg == _abi.ALG_AUTO_TSIT5_TRBDF2
    for alg in (_abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2):
        o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
        e = 1                                         # 5 K/min: TRBDF2 alone equals the oracle count for count
        m, sd = cp.model_for(pb["particles"][0], cp.BETAS[e], pb["t_hi"])
        u0 = np.tile(pb["u0"][e], (8, 1)) * (1.0 - 0.01 * np.arange(8))[:, None]
        data = np.tile(pb["data"][e], (8, 1, 1))
        got = engine.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True)
        ref = oracle.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True, n_threads=8)
        assert (got["retcode"] == _abi.RET_SUCCESS).all()
        np.testing.assert_allclose(got["pred"], ref["pred"], rtol=2e-2, atol=1e-3 * np.abs(ref["pred"]).max())
        assert np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"]) < 2e-2
        if alg == _abi.ALG_TRBDF2:
            same = (got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]) & (got["stats"]["n_accept"] == ref["stats"]["n_accept"])
            assert same.mean() >= 0.75
            np.testing.assert_allclose(got["loss"][same], ref["loss"][same], rtol=1e-5)
    # the SVGD form with the script's algorithm: one launch of particles x experiments
    o5 = cases.cathode_opts(pb["opts"].saveat, alg=_abi.ALG_TRBDF2, pred_clamp=(-np.inf, np.inf))
    got = engine.loss_grad_particles(pb["model"], o5, pb["weights"], pb["seeds"], pb["u0"], pb["data"], pb["yscale"],
                                     _abi.LOSS_MSE, tab_T=pb["tab_T"], want_stats=True)
    pb5 = dict(pb, opts=o5)
    loss, grad, nacc = cp.oracle_particles(pb5, _abi.LOSS_MSE)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()
    np.testing.assert_allclose(got["loss"], loss, rtol=2e-2)
    assert np.abs(got["grad"] - grad).max() < 2e-2 * np.abs(grad).max()
