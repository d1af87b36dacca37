"""Seeded Cathode (DSC) problem shared by the CPU and GPU tests: five heating rates like
Cathode_NCM333_UQ/exp_data/UNCERT_cath_1_{2,5,10,15,20}.csv on a common time grid (in units of the ramp), targets from
a 'true' parameter set with 2 % noise, particles scattered around it (the SVGD initialisation of
src_333/crnn_cathode.jl)."""
import numpy as np

from crnn_b200 import _abi, cases
from oracle import oracle

BETAS = np.array([2.0, 5.0, 10.0, 15.0, 20.0])
P_SCALES = np.array([30.0, 30.0, 30.0, 1.5, 1.5, 1.5, 1.0, 1.0, 1.0, 100.0, 100.0, 100.0, 1.0, 1.0, 1.0, 1.0, 1.0])
N_SAVE = 48


def t_end(beta):
    return 330.0 / (beta / 60.0)     # 100 C -> 430 C


def model_for(p, beta, t_hi):
    w_in, w_b, w_out, w_obs, seed = cases.p2vec_cathode_uq(p, P_SCALES)
    return cases.cathode_model(w_in, w_b, w_out, w_obs, beta, t_hi), seed


def make(P, seed=0, alg=_abi.ALG_ROSENBROCK23, beta_common_grid=True):
    """-> dict(opts, models [E] for particle 0, particles p [P,17], weights [P,n_w], seeds [P,n_w,17], u0 [E,3],
    data [E,N_SAVE,1], tab_T [E,2], yscale)."""
    g = np.random.default_rng(seed)
    p_true = cases.cathode_p_true() / P_SCALES
    # one common grid in RAMP units: experiment e runs to t_end(beta_e); the engine wants one saveat per call, so every
    # experiment is integrated over the same [0, t_hi] with its own heating rate scaled: T(t) = T0 + (330 / t_hi) t * s_e
    t_hi = t_end(BETAS.max())
    ts = np.linspace(0.0, t_hi, N_SAVE)
    opts = cases.cathode_opts(ts, alg=alg, pred_clamp=(-np.inf, np.inf))
    E = BETAS.size
    # experiments differ by their heating rate: over the common window the slower ramps reach lower temperatures
    tab_T = np.stack([cases.cathode_ramp(b, t_hi)[1] for b in BETAS])
    u0 = np.tile(np.array([1.0, 0.0, 0.0]), (E, 1))
    data = np.zeros((E, N_SAVE, 1))
    for e, b in enumerate(BETAS):
        m, _ = model_for(p_true, b, t_hi)
        data[e] = oracle.solve_batch(m, opts, u0[e:e + 1])["pred"][0]
    data *= 1.0 + 0.02 * g.standard_normal(data.shape)
    parts = p_true[None, :] * (1.0 + 0.03 * g.standard_normal((P, 17)))
    ws, sds = [], []
    for q in range(P):
        m, sd = model_for(parts[q], BETAS[0], t_hi)
        ws.append(m.flat_weights()); sds.append(sd)
    m0, _ = model_for(parts[0], BETAS[0], t_hi)
    return dict(opts=opts, model=m0, particles=parts, weights=np.array(ws), seeds=np.array(sds), u0=u0, data=data,
                tab_T=tab_T, yscale=np.array([1.0]), t_hi=t_hi)


def oracle_particles(pb, loss_kind, idx=None, n_threads=8):
    """the reference's loop: for each particle, for each experiment, loss + ForwardDiff.gradient"""
    P = pb["particles"].shape[0]
    idx = range(P) if idx is None else idx
    E = BETAS.size
    loss = np.zeros((len(idx), E)); grad = np.zeros((len(idx), 17)); nacc = np.zeros((len(idx), E), dtype=int)
    for a, q in enumerate(idx):
        for e, b in enumerate(BETAS):
            m, sd = model_for(pb["particles"][q], b, pb["t_hi"])
            r = oracle.loss_grad_batch(m, pb["opts"], sd, pb["u0"][e:e + 1], pb["data"][e:e + 1], pb["yscale"], loss_kind)
            loss[a, e] = r["loss"][0]; grad[a] += r["grad_sum"]; nacc[a, e] = r["stats"]["n_accept"][0]
    return loss, grad, nacc
