"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs.  Bar: accepted/rejected step counts and retcodes identical; saved states,
losses and gradients within the fp64 tolerances written below."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases, synth
from oracle import oracle
from problems import make_problem

pytestmark = pytest.mark.gpu

RTOL_STATE = 1e-9     # saved states vs oracle (north star asks <= 1e-6)
RTOL_LOSS = 1e-10
RTOL_GRAD = 1e-8


def _states_close(got, ref, rtol=RTOL_STATE, scaled=1e-12):
    """Saved states vs the oracle: elementwise rtol plus an absolute floor relative to each
    observed row's range (rounding differences — FMA contraction, libm — are amplified by the
    dynamics; two CPU builds of the oracle itself differ by the same amount)."""
    scale = np.abs(ref).max(axis=(0, 1), keepdims=True)
    err = np.abs(got - ref)
    ok = err <= rtol * np.abs(ref) + scaled * scale
    assert ok.all(), f"max abs err {err.max():.3e}, worst rel {np.max(err / np.maximum(np.abs(ref), 1e-300)):.3e}"


def _counts_equal(got, ref):
    for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
        bad = np.nonzero(got["stats"][k] != ref["stats"][k])[0]
        assert bad.size == 0, f"{k} differs from the oracle for trajectories {bad[:8]}"
    assert np.array_equal(got["retcode"], ref["retcode"])
    assert np.array_equal(got["n_saved"], ref["n_saved"])


@pytest.mark.parametrize("name,N", [("case1", 64), ("case2", 512), ("case3", 128)])
def test_tsit5_value_matches_oracle(engine, golden, name, N):
    pb = make_problem(name, golden, N)
    got = engine.solve_batch(pb["model"], pb["opts"], pb["u0"])
    ref = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    # case3's random-initialised weights (w_in up to 4 on 1e-5-sized states) amplify rounding:
    # the oracle compiled with and without FMA contraction differs from itself by 1e-6 relative
    _states_close(got["pred"], ref["pred"], rtol=(1e-5 if name == "case3" else RTOL_STATE), scaled=1e-9 if name == "case3" else 1e-12)
    assert (got["retcode"] == _abi.RET_SUCCESS).all()


@pytest.mark.parametrize("name", ["case1", "case2", "case3", "robertson"])
def test_true_mechanism_value(engine, golden, name):
    """generating mechanisms written as CRNNs (the synthetic-target path of bench.py)"""
    pb = make_problem(name, golden, 96)
    c = pb["case"]
    o = c.opts(obs_idx=np.arange(c.ns), pred_clamp=(-np.inf, np.inf))
    got = engine.solve_batch(pb["true_model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["true_model"], o, pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    assert (np.abs(got["pred"] - ref["pred"]) / scale).max() < 1e-9


def test_rosenbrock23_value_matches_oracle(engine, golden):
    pb = make_problem("robertson", golden, 256)
    got = engine.solve_batch(pb["model"], pb["opts"], pb["u0"])
    ref = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"], n_threads=8)
    _counts_equal(got, ref)
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    assert (np.abs(got["pred"] - ref["pred"]) / scale).max() < 1e-8


@pytest.mark.parametrize("name,N", [("case1", 64), ("case2", 384)])
def test_tsit5_forward_sens_loss_grad(engine, golden, name, N):
    pb = make_problem(name, golden, N)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["pred"], ref["pred"], rtol=RTOL_STATE, atol=1e-13)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=RTOL_LOSS)
    gmax = np.abs(ref["grad_sum"]).max()
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=RTOL_GRAD, atol=1e-10 * gmax)


def test_sens_value_only_norm_switch(engine, golden):
    """err_norm_includes_sens = 0: the step sequence must equal the value-only solve's."""
    pb = make_problem("case2", golden, 128)
    o = pb["case"].opts(obs_idx=np.arange(6), err_norm_includes_sens=False)
    got = engine.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    val = engine.solve_batch(pb["model"], o, pb["u0"])
    assert np.array_equal(got["stats"]["n_accept"], val["stats"]["n_accept"])
    ref = oracle.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"], n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=RTOL_GRAD, atol=1e-10 * np.abs(ref["grad_sum"]).max())


def test_mae_log_loss_kind(engine, golden):
    """case3's loss (log-MAE with (lb,ub) clamps) on a case1-sized model (np=20 fits forward mode)."""
    pb = make_problem("case1", golden, 64)
    o = pb["case"].opts(obs_idx=np.arange(5), pred_clamp=(1e-5, 10.0))
    args = (pb["model"], o, pb["seed"], pb["u0"], np.abs(pb["data"]) + 1e-7, pb["yscale"], _abi.LOSS_MAE_LOG)
    got = engine.loss_grad_batch(*args)
    ref = oracle.loss_grad_batch(*args, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=RTOL_LOSS)
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=RTOL_GRAD, atol=1e-10 * np.abs(ref["grad_sum"]).max())


def test_missing_species_obs_idx(engine, golden):
    """case2_missing.jl:165: i_obs = [1,2,4,5,6] (species 3 unobserved)."""
    obs = np.array([0, 1, 3, 4, 5])
    pb = make_problem("case2", golden, 96, obs=obs)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["pred"], ref["pred"], rtol=RTOL_STATE, atol=1e-13)
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=RTOL_GRAD, atol=1e-10 * np.abs(ref["grad_sum"]).max())


def test_device_buffers_equal_host_buffers(engine, golden):
    import torch
    pb = make_problem("case2", golden, 300)
    host = engine.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"],
                                  pb["loss_kind"], want_pred=True)
    u0 = torch.from_numpy(pb["u0"]).cuda(); data = torch.from_numpy(pb["data"]).cuda()
    dev = engine.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], u0, data, pb["yscale"], pb["loss_kind"],
                                 want_pred=True)
    torch.cuda.synchronize()
    assert np.array_equal(dev["pred"].cpu().numpy(), host["pred"])
    assert np.array_equal(dev["loss"].cpu().numpy(), host["loss"])
    assert np.array_equal(dev["retcode"].cpu().numpy(), host["retcode"])
    # the reduction order differs between the chunked host path and the single-launch device path
    np.testing.assert_allclose(dev["grad_sum"].cpu().numpy(), host["grad_sum"], rtol=1e-12)


def test_edge_empty_single_and_ragged(engine, golden):
    pb = make_problem("case2", golden, 33)
    # N = 0
    r0 = engine.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"][:0], pb["data"][:0], pb["yscale"])
    assert r0["loss"].shape == (0,) and np.all(r0["grad_sum"] == 0)
    # N = 1
    r1 = engine.solve_batch(pb["model"], pb["opts"], pb["u0"][:1])
    ref1 = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"][:1])
    np.testing.assert_allclose(r1["pred"], ref1["pred"], rtol=RTOL_STATE, atol=1e-13)
    # ragged time truncation (rober_crnn.jl:218: sample = rand(batchsize:datasize))
    nsu = np.random.default_rng(3).integers(5, 51, size=33).astype(np.int32)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, n_save_used=nsu, want_pred=True)
    ref = oracle.loss_grad_batch(*args, n_save_used=nsu, want_pred=True, n_threads=4)
    _counts_equal(got, ref)
    assert np.array_equal(got["n_saved"], nsu)
    np.testing.assert_allclose(got["pred"], ref["pred"], rtol=RTOL_STATE, atol=1e-13)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=RTOL_LOSS)


def test_retcode_paths(engine, golden):
    pb = make_problem("case2", golden, 40)
    c = pb["case"]
    # maxiters truncation: n_saved < n_save, MaxIters, batch does not abort
    o = c.opts(obs_idx=np.arange(6), maxiters=5)
    got = engine.solve_batch(pb["model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["model"], o, pb["u0"])
    _counts_equal(got, ref)
    assert (got["retcode"] == _abi.RET_MAXITERS).all() and (got["n_saved"] < 50).all()
    for i in range(40):
        k = got["n_saved"][i]
        np.testing.assert_allclose(got["pred"][i, :k], ref["pred"][i, :k], rtol=RTOL_STATE, atol=1e-13)
        assert np.all(got["pred"][i, k:] == 0)
    # NaN weights: flagged per trajectory, never an engine error
    m = cases.CRNNModel(w_in=pb["model"].w_in, w_b=pb["model"].w_b * np.nan, w_out=pb["model"].w_out,
                        rhs_kind=pb["model"].rhs_kind, lb=pb["model"].lb, ub=pb["model"].ub)
    bad = engine.solve_batch(m, pb["opts"], pb["u0"])
    refb = oracle.solve_batch(m, pb["opts"], pb["u0"])
    assert np.array_equal(bad["retcode"], refb["retcode"])
    assert np.isin(bad["retcode"], [_abi.RET_DTNAN, _abi.RET_UNSTABLE]).all()


def test_bad_arguments_are_engine_errors(engine, golden):
    from crnn_b200.engine import EngineError
    pb = make_problem("case2", golden, 4)
    with pytest.raises(EngineError):
        engine.solve_batch(pb["model"], pb["case"].opts(obs_idx=np.array([0, 0])), pb["u0"])
    with pytest.raises(EngineError):   # beyond the generic kernel's 32 state components
        m = cases.CRNNModel(w_in=np.ones((40, 2)), w_b=np.zeros(2), w_out=np.ones((40, 2)))
        engine.solve_batch(m, pb["case"].opts(obs_idx=np.arange(40)), np.ones((2, 40)))
    m = cases.CRNNModel(w_in=np.ones((4, 2)), w_b=np.zeros(2), w_out=np.ones((4, 2)))
    with pytest.raises(EngineError):   # a DENSE seed column (several w_in rows) on dimensions without a specialised kernel,
        # partials in the error norm: neither the generic forward kernel nor the discrete adjoint computes that
        engine.loss_grad_batch(m, pb["case"].opts(obs_idx=np.arange(4)), np.ones((m.n_w, 3)), np.ones((2, 4)),
                               np.ones((2, 50, 4)), np.ones(4))
    # structured columns on the same dimensions are served by the generic forward kernel
    sd = np.zeros((m.n_w, 3)); sd[0, 0] = 1.0; sd[8, 1] = 1.0; sd[10, 2] = 1.0
    r = engine.loss_grad_batch(m, pb["case"].opts(obs_idx=np.arange(4), t1=1.0, saveat=np.linspace(0, 1, 50)), sd,
                               np.full((2, 4), 0.5), np.ones((2, 50, 4)), np.ones(4))
    assert (r["retcode"] == 1).all() and np.isfinite(r["grad_sum"]).all()


def test_full_size_properties_case2(engine, golden):
    """BASELINE configs[1] size (65 536 ICs): size-independent properties instead of the oracle."""
    import torch
    c = cases.CASES["case2"]
    N = 65536
    u0 = synth.make_u0("case2", N)
    tm = cases.true_model_case2()
    o = c.opts(obs_idx=np.arange(6), pred_clamp=(-np.inf, np.inf))
    tr = engine.solve_batch(tm, o, u0)
    assert (tr["retcode"] == 1).all() and (tr["n_saved"] == 50).all()
    y = tr["pred"]
    # element balances of the transesterification mechanism (case2.jl:40-48):
    # glycerol backbone TG+DG+MG+GL and acyl groups 3TG+2DG+MG+ester are conserved, ROH+ester too
    for w in ([1, 0, 1, 1, 1, 0], [3, 0, 2, 1, 0, 1], [0, 1, 0, 0, 0, 1]):
        inv = y @ np.array(w, dtype=float)
        assert np.abs(inv - inv[:, :1]).max() < 2e-3
    assert np.array_equal(y[:, 0, :], u0[:, :6])          # t0 is in saveat: saved exactly
    # CRNN loss/gradient at full size: sharded sum == whole-batch sum (linearity of grad_sum)
    data = synth.noisy_targets(y, 0.05)
    ys = synth.yscale_from(data, c.lb)
    model, seed = c.model(np.array(golden["case2"]["p"]))
    opts = c.opts(obs_idx=np.arange(6))
    whole = engine.loss_grad_batch(model, opts, seed, u0, data, ys)
    assert (whole["retcode"] == 1).all()
    parts = [engine.loss_grad_batch(model, opts, seed, u0[a:b], data[a:b], ys)
             for a, b in ((0, 20000), (20000, 65536))]
    np.testing.assert_allclose(sum(p["grad_sum"] for p in parts), whole["grad_sum"], rtol=1e-11)
    assert np.array_equal(np.concatenate([p["loss"] for p in parts]), whole["loss"])
    # trained checkpoint reproduces the generating mechanism: normalised MAE of the order the
    # reference's own loss history ends at (1.4e-2..1.7e-2 incl. 5 % noise)
    assert 0.01 < whole["loss"].mean() < 0.06
    # a random sample against the oracle
    idx = np.random.default_rng(0).choice(N, 64, replace=False)
    ref = oracle.loss_grad_batch(model, opts, seed, u0[idx], data[idx], ys, n_threads=8)
    np.testing.assert_allclose(whole["loss"][idx], ref["loss"], rtol=RTOL_LOSS)
    assert np.array_equal(whole["stats"]["n_accept"][idx], ref["stats"]["n_accept"])


# ---------------------------------------------------------------- wide gradients: several warps per trajectory
def _grad_close(got, ref, rtol=RTOL_GRAD):
    gmax = np.abs(ref["grad_sum"]).max()
    np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=rtol, atol=1e-9 * gmax)


def test_case3_forward_sens_np153_five_warps_per_trajectory(engine, golden):
    """case3/case3.jl:265-270: Zygote.forwarddiff over all 153 parameters, log-MAE loss (:183-190).
    154 dual columns = 5 warps sharing one trajectory (named-barrier group)."""
    pb = make_problem("case3", golden, 48)
    # targets must be positive for the log loss; the script clamps them to [lb, ub] (case3.jl:185)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], np.abs(pb["data"]) + 1e-6, pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    _counts_equal(got, ref)
    _states_close(got["pred"], ref["pred"], rtol=1e-5, scaled=1e-9)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-7)
    _grad_close(got, ref, rtol=1e-5)       # rounding-amplifying random weights (see test_tsit5_value)


def test_three_warps_per_trajectory_column_subset(engine, golden):
    """80 of case3's seed columns (np = 80 -> 3 tiles): also how ForwardDiff-style chunking is expressed."""
    pb = make_problem("case3", golden, 32)
    seed = pb["seed"][:, :80]
    args = (pb["model"], pb["opts"], seed, pb["u0"], np.abs(pb["data"]) + 1e-6, pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args)
    ref = oracle.loss_grad_batch(*args, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-7)
    _grad_close(got, ref, rtol=1e-5)


def test_two_tiles_per_lane_np43(engine, golden):
    """robertson's p2vec (np = 43 -> two column tiles per lane) integrated with Tsit5 on a short span."""
    pb = make_problem("robertson", golden, 24)
    c = pb["case"]
    o = c.opts(alg=_abi.ALG_TSIT5, t1=2.0, saveat=np.linspace(0.1, 2.0, 12), abstol=1e-8, reltol=1e-4, maxiters=200000)
    data = pb["data"][:, :12, :]
    args = (pb["model"], o, pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args)
    ref = oracle.loss_grad_batch(*args, n_threads=8)
    _counts_equal(got, ref)
    assert (got["retcode"] == 1).all()
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-8)
    _grad_close(got, ref, rtol=1e-6)


def test_dense_seed_fallback(engine, golden):
    """A seed that is not 'one w_in row + one w_out entry per column' takes the dense layout."""
    pb = make_problem("case2", golden, 64)
    rng = np.random.default_rng(5)
    seed = rng.standard_normal((pb["model"].n_w, 9)) * (rng.random((pb["model"].n_w, 9)) < 0.5)
    args = (pb["model"], pb["opts"], seed, pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args)
    ref = oracle.loss_grad_batch(*args, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=RTOL_LOSS)
    _grad_close(got, ref)


def test_rosenbrock23_forward_sens_robertson(engine, golden):
    """The reference's stiff training path: ForwardDiff.gradient through Rosenbrock23(autodiff=true)
    (rober_crnn.jl:33,139-144,219), vector tolerances (:34-35), random time truncation (:218)."""
    pb = make_problem("robertson", golden, 96)
    nsu = np.random.default_rng(7).integers(32, 41, size=96).astype(np.int32)   # sample = rand(batchsize:datasize)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, n_save_used=nsu, want_pred=True)
    ref = oracle.loss_grad_batch(*args, n_save_used=nsu, want_pred=True, n_threads=8)
    _counts_equal(got, ref)
    assert (got["retcode"] == 1).all() and np.array_equal(got["n_saved"], nsu)
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    assert (np.abs(got["pred"] - ref["pred"]) / scale).max() < 1e-7
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-6)
    _grad_close(got, ref, rtol=1e-5)


def test_rosenbrock23_forward_sens_case2(engine, golden):
    """stiff half of AutoTsit5(Rosenbrock23) on the case2 model (F1: Arrhenius row, T as state)."""
    pb = make_problem("case2", golden, 64)
    o = pb["case"].opts(obs_idx=np.arange(6), alg=_abi.ALG_ROSENBROCK23)
    args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args)
    ref = oracle.loss_grad_batch(*args, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-8)
    _grad_close(got, ref, rtol=1e-6)


# ---------------------------------------------------------------- KenCarp4: generic-dimension, lane = state component
def _kc4_close(got, ref, frac_counts=0.02):
    """Newton iteration counts can flip on a rounding-level difference at the eta*|dz| < kappa test;
    allow that for a small fraction of trajectories, the rest must match exactly."""
    same = (got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]) & (got["stats"]["n_accept"] == ref["stats"]["n_accept"]) & \
           (got["stats"]["n_reject"] == ref["stats"]["n_reject"])
    assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])
    assert same.mean() >= 1.0 - frac_counts, f"{(~same).sum()} of {same.size} trajectories differ in step/RHS counts"
    scale = np.abs(ref["pred"]).max(axis=(0, 1))
    err = np.abs(got["pred"] - ref["pred"]) / scale
    assert err[same].max() < 1e-7
    assert err.max() < 5e-3           # a flipped Newton count still solves the same ODE to tolerance


@pytest.mark.parametrize("name", ["robertson", "case2"])
def test_kencarp4_small_models(engine, golden, name):
    pb = make_problem(name, golden, 128)
    c = pb["case"]
    o = c.opts(obs_idx=np.arange(c.ns), alg=_abi.ALG_KENCARP4)
    got = engine.solve_batch(pb["model"], o, pb["u0"])
    ref = oracle.solve_batch(pb["model"], o, pb["u0"], n_threads=8)
    assert (got["retcode"] == 1).all()
    _kc4_close(got, ref)


def test_kencarp4_hychem_sized(engine):
    """BASELINE config 5 shape: 29 species + T, 30 reactions, stiff, 40 log-spaced saves."""
    m = cases.synthetic_stiff_model(); u0 = cases.synthetic_stiff_u0(256); o = cases.synthetic_stiff_opts()
    got = engine.solve_batch(m, o, u0)
    ref = oracle.solve_batch(m, o, u0, n_threads=8)
    assert (got["retcode"] == 1).all() and np.array_equal(got["n_saved"], ref["n_saved"])
    # Newton decisions can flip on last-ulp differences between CUDA's and glibc's exp/log; allow a few
    # trajectories to differ in counts (they still solve the same ODE to tolerance).
    # trace species sit at the abstol (1e-8) level: measure against max(species range, 1e-4)
    scale = np.maximum(np.abs(ref["pred"]).max(axis=(0, 1)), 1e-4)
    assert (np.abs(got["pred"] - ref["pred"]) / scale).max() < 5e-3
    same = got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]
    assert same.mean() > 0.9 and (np.abs(got["pred"] - ref["pred"])[same] / scale).max() < 2e-6
    assert abs(got["stats"]["n_rhs"].sum() / ref["stats"]["n_rhs"].sum() - 1.0) < 0.1
    assert np.abs(got["pred"][:, -1].sum(axis=1) - u0[:, :29].sum(axis=1)).max() < 1e-6   # sum(u) conserved
    nsu = np.random.default_rng(1).integers(5, 41, size=256).astype(np.int32)
    g2 = engine.solve_batch(m, o, u0, n_save_used=nsu)
    assert np.array_equal(g2["n_saved"], nsu)


# ---------------------------------------------------------------- interpolating adjoint (BASELINE config 4)
@pytest.mark.parametrize("name,N", [("case2", 192), ("case3", 96), ("case1", 48)])
def test_adjoint_matches_oracle_and_forward_mode(engine, golden, name, N):
    pb = make_problem(name, golden, N)
    c = pb["case"]
    data = np.abs(pb["data"]) + 1e-6 if name == "case3" else pb["data"]
    oa = c.opts(obs_idx=np.arange(c.ns), sens_mode=_abi.SENS_INTERP_ADJOINT)
    args = (pb["model"], oa, pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, want_pred=True)
    ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
    amp = name == "case3"                        # rounding-amplifying random weights (see test_tsit5_value)
    if amp:
        # ~100 backward segments per trajectory with an accept test each: a last-ulp difference can flip one;
        # forward counts, retcodes and saves must still match, backward counts for all but a few
        for k in ("n_accept", "n_reject"):
            assert np.array_equal(got["stats"][k], ref["stats"][k])
        assert np.array_equal(got["retcode"], ref["retcode"]) and np.array_equal(got["n_saved"], ref["n_saved"])
        assert (got["stats"]["n_jac"] == ref["stats"]["n_jac"]).mean() >= 0.95
    else:
        _counts_equal(got, ref)                  # forward steps, backward steps (n_jac), RHS counts, retcodes
    _states_close(got["pred"], ref["pred"], rtol=1e-5 if amp else RTOL_STATE, scaled=1e-9 if amp else 1e-12)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-7 if amp else RTOL_LOSS)
    if amp:   # a flipped backward step changes that trajectory's gradient at tolerance level
        assert np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"]) < 1e-4
    else:
        _grad_close(got, ref, rtol=1e-7)
    # and the continuous adjoint agrees with the discrete forward-mode gradient to O(tolerance)
    of = c.opts(obs_idx=np.arange(c.ns))
    fwd = engine.loss_grad_batch(pb["model"], of, pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    # (case3's random weights: a few trajectories are integrated coarsely by the value-only forward pass
    #  and their continuous-adjoint gradient is tens of % off the discrete one at reltol 1e-3 — the oracle
    #  shows the same; the tight-tolerance agreement is asserted in test_oracle_cpu.py)
    assert np.linalg.norm(got["grad_sum"] - fwd["grad_sum"]) / np.linalg.norm(fwd["grad_sum"]) < (0.2 if amp else 1e-3)


def test_adjoint_ragged_missing_species_and_device_buffers(engine, golden):
    import torch
    obs = np.array([0, 1, 3, 4, 5])
    pb = make_problem("case2", golden, 130, obs=obs)
    oa = pb["case"].opts(obs_idx=obs, sens_mode=_abi.SENS_INTERP_ADJOINT)
    nsu = np.random.default_rng(11).integers(1, 51, size=130).astype(np.int32)
    args = (pb["model"], oa, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args, n_save_used=nsu)
    ref = oracle.loss_grad_batch(*args, n_save_used=nsu, n_threads=8)
    _counts_equal(got, ref)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=RTOL_LOSS)
    _grad_close(got, ref, rtol=1e-7)
    dev = engine.loss_grad_batch(pb["model"], oa, pb["seed"], torch.from_numpy(pb["u0"]).cuda(),
                                 torch.from_numpy(pb["data"]).cuda(), pb["yscale"], pb["loss_kind"],
                                 n_save_used=torch.from_numpy(nsu).cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(dev["grad_sum"].cpu().numpy(), got["grad_sum"], rtol=1e-11)
    assert np.array_equal(dev["loss"].cpu().numpy(), got["loss"])


@pytest.mark.parametrize("name,N", [("case2", 256), ("case3", 128), ("case1", 64)])
def test_discrete_adjoint_equals_forward_mode_on_gpu(engine, golden, name, N):
    """CRNN_SENS_DISCRETE_ADJOINT: same gradient as the forward-mode kernel with the value-only error norm
    (two independent CUDA code paths), and parity with the oracle."""
    pb = make_problem(name, golden, N)
    c = pb["case"]
    data = np.abs(pb["data"]) + 1e-6 if name == "case3" else pb["data"]
    nsu = np.random.default_rng(5).integers(1, c.n_save + 1, size=N).astype(np.int32)
    od = c.opts(obs_idx=np.arange(c.ns), sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    of = c.opts(obs_idx=np.arange(c.ns), err_norm_includes_sens=False)
    args = (pb["seed"], pb["u0"], data, pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(pb["model"], od, *args, n_save_used=nsu, want_pred=True)
    ref = oracle.loss_grad_batch(pb["model"], od, *args, n_save_used=nsu, want_pred=True, n_threads=8)
    fwd = engine.loss_grad_batch(pb["model"], of, *args, n_save_used=nsu, want_pred=True)
    _counts_equal(got, ref)
    amp = name == "case3"
    _states_close(got["pred"], ref["pred"], rtol=1e-5 if amp else RTOL_STATE, scaled=1e-9 if amp else 1e-12)
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=1e-7 if amp else RTOL_LOSS)
    _grad_close(got, ref, rtol=1e-5 if amp else 1e-8)
    assert np.array_equal(got["stats"]["n_accept"], fwd["stats"]["n_accept"])
    # two different CUDA code paths for the same discrete solve: rounding only (amplified for case3)
    np.testing.assert_allclose(got["loss"], fwd["loss"], rtol=1e-6 if amp else 1e-11)
    if amp:
        assert np.linalg.norm(got["grad_sum"] - fwd["grad_sum"]) / np.linalg.norm(fwd["grad_sum"]) < 1e-5
    else:
        _grad_close(got, fwd, rtol=1e-8)


def test_per_trajectory_gradients(engine, golden):
    """crnn_copy_grad_each: the per-experiment gradients robertson/rober_crnn_lm.jl's LM variant needs."""
    pb = make_problem("case2", golden, 50)
    args = (pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    got = engine.loss_grad_batch(*args)
    ge = engine.grad_each(50, pb["seed"].shape[1])
    ref = oracle.loss_grad_batch(*args, want_grad_each=True, n_threads=4)
    np.testing.assert_allclose(ge, ref["grad_each"], rtol=1e-7, atol=1e-10 * np.abs(ref["grad_each"]).max())
    np.testing.assert_allclose(ge.sum(axis=0), got["grad_sum"], rtol=1e-11)
