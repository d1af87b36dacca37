"""RHS flavour F5 (Cathode/src/network.jl:68-80), the heat-release observable (:82-91,121) and the MSE loss of the SVGD
variant (Cathode_NCM333_UQ/src_333/network.jl:262-275) in the oracle: against a line-by-line numpy transcription of the
scripts' functions, finite differences, and Radau."""
import numpy as np
from scipy.integrate import solve_ivp

from crnn_b200 import _abi, cases
from oracle import oracle
import cathode_problem as cp

R = -1.0 / 8.314


def literal(p, beta, t, u, lb=1e-8):
    """crnn! and HRR_getter of src_333/network.jl:153-180 with p already in physical units"""
    logX = np.log(np.clip(u, lb, 10.0))
    T = cases.CATHODE_T0 + beta / 60.0 * t
    temp_term = np.log(T) * p[6:9] + R / T * (p[3:6] * 1e5)
    rates = np.exp(temp_term + p[12:15] * logX + p[0:3])
    du = -rates.copy()
    du[1] += p[15] * rates[0]; du[2] += p[16] * rates[1]
    return du, rates @ p[9:12]


def test_f5_rhs_and_heat_release_equal_the_scripts_functions():
    g = np.random.default_rng(0)
    p_phys = cases.cathode_p_true()
    for beta in (2.0, 10.0, 20.0):
        m, _ = cp.model_for(p_phys / cp.P_SCALES, beta, cp.t_end(beta))
        for _ in range(5):
            u = g.random(3) * np.array([1.0, 0.5, 0.2]); u[g.integers(3)] = 1e-12      # one species below the clamp
            t = g.random() * cp.t_end(beta)
            f, J, dT = oracle.rhs_t(m, t, u)
            np.testing.assert_allclose(f, literal(p_phys, beta, t, u)[0], rtol=1e-13)
            h = 1e-6
            Jfd = np.array([(literal(p_phys, beta, t, u + h * e)[0] - literal(p_phys, beta, t, u - h * e)[0]) / (2 * h) for e in np.eye(3)]).T
            inside = u >= 1e-8
            np.testing.assert_allclose(J[:, inside], Jfd[:, inside], rtol=1e-5, atol=1e-9 * np.abs(Jfd).max())
            assert np.all(J[:, ~inside] == 0)
            dTfd = (literal(p_phys, beta, t + 1e-3, u)[0] - literal(p_phys, beta, t - 1e-3, u)[0]) / 2e-3
            np.testing.assert_allclose(dT, dTfd, rtol=1e-6, atol=1e-12)


def test_p2vec_cathode_maps_and_seeds():
    g = np.random.default_rng(1)
    for fn, p in ((cases.p2vec_cathode, np.concatenate([1 + 0.01 * g.standard_normal(3), [1.0, 1.1, 1.2] + 0.01 * g.standard_normal(3),
                                                         0.01 * g.standard_normal(3), [1.0, 0.2, 0.3] + 0.01 * g.standard_normal(3),
                                                         1 + 0.01 * g.standard_normal(3), 1 + 0.01 * g.standard_normal(2), [0.1]])),
                  (lambda q: cases.p2vec_cathode_uq(q, cp.P_SCALES), cases.cathode_p_true() / cp.P_SCALES)):
        w_in, w_b, w_out, w_obs, seed = fn(p)
        assert w_in.shape == (5, 3) and seed.shape == (30, p.size)
        flat = lambda q: np.concatenate([a.reshape(-1, order="F") for a in fn(q)[:4]])
        h = 1e-7
        fd = np.array([(flat(p + h * e) - flat(p - h * e)) / (2 * h) for e in np.eye(p.size)]).T
        np.testing.assert_allclose(seed, fd, rtol=1e-6, atol=1e-6)
    # the deterministic script's initialisation (network.jl:9-24): ln A = p*slope*20 = 20, orders 1, stoichiometry 1
    p0 = np.zeros(18); p0[0:3] = 1; p0[3:6] = [1.0, 1.1, 1.2]; p0[9:12] = [1.0, 0.2, 0.3]; p0[12:15] = 1; p0[15:17] = 1; p0[17] = 0.1
    w_in, w_b, w_out, w_obs, _ = cases.p2vec_cathode(p0)
    np.testing.assert_allclose(w_b, 20.0); np.testing.assert_allclose(np.diag(w_in[:3]), 1.0)
    np.testing.assert_allclose(w_in[3], [1.0e5, 1.1e5, 1.2e5]); np.testing.assert_allclose(w_obs, [100.0, 20.0, 30.0])
    np.testing.assert_allclose(w_out, [[-1, 0, 0], [1, -1, 0], [0, 1, -1]])


def test_heat_release_trajectory_against_radau():
    p_phys = cases.cathode_p_true()
    beta = 10.0; te = cp.t_end(beta)
    m, _ = cp.model_for(p_phys / cp.P_SCALES, beta, te)
    ts = np.linspace(0, te, 60)
    sol = solve_ivp(lambda t, y: literal(p_phys, beta, t, y)[0], (0, te), [1.0, 0, 0], method="Radau", rtol=1e-10, atol=1e-13, t_eval=ts)
    hr = np.array([literal(p_phys, beta, t, y)[1] for t, y in zip(ts, sol.y.T)])
    # AutoTsit5(TRBDF2(autodiff = true)) is what the script runs (Cathode/src/network.jl:102)
    for alg in (_abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2):
        r = oracle.solve_batch(m, cases.cathode_opts(ts, alg=alg), np.array([[1.0, 0, 0]]))
        assert r["retcode"][0] == _abi.RET_SUCCESS and r["pred"].shape == (1, 60, 1)
        assert np.abs(r["pred"][0, :, 0] - hr).max() < 2e-3 * hr.max()


def test_gradient_of_the_heat_release_loss_against_finite_differences():
    """ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p_temp) (src_333/network.jl:232): forward sensitivities through
    Rosenbrock23 with the observable post-map, MSE and MAE losses, vs central differences of tight solves"""
    pb = cp.make(1, seed=3)
    p = pb["particles"][0]
    e = 2
    o = cases.cathode_opts(pb["opts"].saveat, alg=_abi.ALG_ROSENBROCK23, abstol=1e-13, reltol=1e-9, maxiters=10 ** 7)
    for kind in (_abi.LOSS_MSE, _abi.LOSS_MAE_SCALED):
        m, sd = cp.model_for(p, cp.BETAS[e], pb["t_hi"])
        r = oracle.loss_grad_batch(m, o, sd, pb["u0"][e:e + 1], pb["data"][e:e + 1], pb["yscale"], kind, want_pred=True)

        def loss_at(q):
            mm, _ = cp.model_for(q, cp.BETAS[e], pb["t_hi"])
            pr = oracle.solve_batch(mm, o, pb["u0"][e:e + 1])["pred"]
            d = pr - pb["data"][e:e + 1]
            return np.mean(d ** 2) if kind == _abi.LOSS_MSE else np.mean(np.abs(d))
        assert abs(r["loss"][0] - loss_at(p)) < 1e-9      # partials in the norm: another step sequence
        for k in range(17):
            h = 1e-4 * max(abs(p[k]), 0.1)
            pp, pm = p.copy(), p.copy(); pp[k] += h; pm[k] -= h
            fd = (loss_at(pp) - loss_at(pm)) / (2 * h)
            assert abs(r["grad_sum"][k] - fd) < 1e-3 * max(abs(fd), 1e-3 * np.abs(r["grad_sum"]).max()), (kind, k, r["grad_sum"][k], fd)
