"""Parameter-batched solves: every trajectory its own weights - the SVGD loop of
Cathode_NCM333_UQ/src_333/network.jl:222-260 (100 particles x 5 data sets, sequential in the reference) as ONE launch,
with RHS flavour F5, per-experiment temperature programmes and the heat-release observable; against the oracle run
particle by particle, experiment by experiment."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases
from oracle import oracle
import cathode_problem as cp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [_abi.LOSS_MSE, _abi.LOSS_MAE_SCALED])
def test_particles_equal_the_per_particle_loop(engine, kind):
    pb = cp.make(6, seed=1)
    got = engine.loss_grad_particles(pb["model"], pb["opts"], pb["weights"], pb["seeds"], pb["u0"], pb["data"], pb["yscale"],
                                     kind, tab_T=pb["tab_T"], want_stats=True)
    loss, grad, nacc = cp.oracle_particles(pb, kind)
    assert (got["retcode"] == _abi.RET_SUCCESS).all() and (got["n_saved"] == cp.N_SAVE).all()
    assert np.array_equal(got["stats"]["n_accept"], nacc)                  # same step sequence, every (particle, experiment)
    np.testing.assert_allclose(got["loss"], loss, rtol=1e-8)
    np.testing.assert_allclose(got["grad"], grad, rtol=1e-6, atol=1e-8 * np.abs(grad).max())


def test_svgd_shape_100_particles_x_5_datasets(engine):
    """the reference's shape: num_particles = 100 (config.yaml:36), five heating rates; one launch of 500 trajectories"""
    pb = cp.make(100, seed=2)
    got = engine.loss_grad_particles(pb["model"], pb["opts"], pb["weights"], pb["seeds"], pb["u0"], pb["data"], pb["yscale"],
                                     _abi.LOSS_MSE, tab_T=pb["tab_T"], want_stats=True)
    assert got["loss"].shape == (100, 5) and got["grad"].shape == (100, 17)
    assert (got["retcode"] == _abi.RET_SUCCESS).all() and np.isfinite(got["grad"]).all()
    idx = [0, 17, 42, 99]
    loss, grad, nacc = cp.oracle_particles(pb, _abi.LOSS_MSE, idx=idx)
    assert np.array_equal(got["stats"]["n_accept"][idx], nacc)
    np.testing.assert_allclose(got["loss"][idx], loss, rtol=1e-8)
    np.testing.assert_allclose(got["grad"][idx], grad, rtol=1e-6, atol=1e-8 * np.abs(grad).max())
    # dlnprob's outputs (network.jl:222-260): mean loss over the particles and -grad with the Normalizer scaling stay host work
    assert got["loss"].sum(axis=1).mean() > 0


def test_f5_with_shared_weights_batched_and_predict(engine):
    """one parameter set, many experiments: crnn_loss_grad_batch / crnn_solve_batch on the F5 model"""
    pb = cp.make(1, seed=4)
    g = np.random.default_rng(0)
    N = 64
    u0 = np.tile(np.array([1.0, 0.0, 0.0]), (N, 1)) * (1.0 + 0.05 * g.random((N, 1)))
    m, sd = cp.model_for(pb["particles"][0], 10.0, pb["t_hi"])
    data = oracle.solve_batch(m, pb["opts"], u0, n_threads=8)["pred"] * 1.05
    for alg in (_abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_ROS23):
        o = cases.cathode_opts(pb["opts"].saveat, alg=alg)
        got = engine.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True)
        ref = oracle.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True, n_threads=8)
        # a stiff model: one accept test in ~3000 may flip on rounding (kernel and oracle order their sums differently)
        same = np.ones(N, dtype=bool)
        for k in ("n_accept", "n_reject", "n_jac"):
            same &= got["stats"][k] == ref["stats"][k]
        if alg == _abi.ALG_ROSENBROCK23:
            assert same.mean() >= 0.95
            np.testing.assert_allclose(got["pred"][same], ref["pred"][same], rtol=1e-7, atol=1e-10)
            np.testing.assert_allclose(got["loss"][same], ref["loss"][same], rtol=1e-7)
        # (the composite stays on Tsit5 here, stepping at its stability limit: accept tests sit within rounding of
        #  EEst = 1 and the two implementations take different, equally valid step sequences)
        np.testing.assert_allclose(got["pred"], ref["pred"], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(got["grad_sum"], ref["grad_sum"], rtol=1e-3, atol=1e-5 * np.abs(ref["grad_sum"]).max())
    # species trajectories (no observable map) on the generic predict kernel
    ms = cases.CRNNModel(w_in=m.w_in, w_b=m.w_b, w_out=m.w_out, rhs_kind=_abi.RHS_F5, lb=m.lb, ub=m.ub, gas_R=m.gas_R,
                         tab_t=m.tab_t, tab_T=m.tab_T)
    os_ = cases.cathode_opts(pb["opts"].saveat, alg=_abi.ALG_ROSENBROCK23, obs_idx=np.arange(3))
    gs = engine.solve_batch(ms, os_, u0)
    rs = oracle.solve_batch(ms, os_, u0, n_threads=8)
    same = gs["stats"]["n_accept"] == rs["stats"]["n_accept"]
    assert same.mean() >= 0.95
    np.testing.assert_allclose(gs["pred"][same], rs["pred"][same], rtol=1e-7, atol=1e-10)
    # mass balance of the sequential scheme with unit stoichiometry would be c1 + c2 + c3 non-increasing
    assert (np.diff(gs["pred"][:, :, 0], axis=1) <= 1e-9).all()
