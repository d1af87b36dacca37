"""The reference scripts' `p2vec` maps (flat trainable vector p -> physical weights)
and their Jacobians dW/dp (the seed matrix `crnn_loss_grad_batch` takes).

In the Julia front end this part STAYS in Julia: `W = p2vec(p)` and
`S = ForwardDiff.jacobian(p2vec_flat, p)`.  Here the same maps are written once
over a tiny forward-mode dual array (`_D`) whose derivative conventions are
ForwardDiff's: clamp passes derivative 1 on the closed interval, abs uses
signbit (derivative +1 at +0).

  case1      case1/case1.jl:72-78      (b0 = -10, :70)
  case2      case2/case2.jl:91-99
  case3      case3/case3.jl:42-53
  robertson  robertson/rober_crnn.jl:85-96
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi
from .model import CRNNModel, SolveOpts


class _D:
    """value [..] + jacobian [.., np] w.r.t. the flat parameter vector."""

    def __init__(self, v, j):
        self.v = np.asarray(v, dtype=np.float64)
        self.j = np.asarray(j, dtype=np.float64)

    @staticmethod
    def seed(p):
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        return _D(p, np.eye(p.size))

    def __getitem__(self, idx):
        return _D(self.v[idx], self.j[idx])

    def reshape_f(self, *shape):  # Julia reshape = column-major
        n_p = self.j.shape[-1]
        return _D(self.v.reshape(shape, order="F"), self.j.reshape(shape + (n_p,), order="F"))

    def __mul__(self, o):
        if isinstance(o, _D):
            return _D(self.v * o.v, self.j * o.v[..., None] + o.j * self.v[..., None])
        return _D(self.v * o, self.j * o)

    __rmul__ = __mul__

    def __add__(self, o):
        if isinstance(o, _D):
            return _D(self.v + o.v, self.j + o.j)
        return _D(self.v + o, self.j)

    def __neg__(self):
        return _D(-self.v, -self.j)

    def abs(self):
        s = np.where(np.signbit(self.v), -1.0, 1.0)
        return _D(self.v * s, self.j * s[..., None])

    def clamp(self, lo, hi):
        inside = (self.v >= lo) & (self.v <= hi)
        return _D(np.clip(self.v, lo, hi), self.j * inside[..., None])

    def pow10(self):
        v = 10.0 ** self.v
        return _D(v, self.j * (v * np.log(10.0))[..., None])

    def broadcast_scalar(self, shape):
        return _D(np.broadcast_to(self.v, shape), np.broadcast_to(self.j, shape + self.j.shape[-1:]))


def _pack(w_in: _D, w_b: _D, w_out: _D):
    """-> (w_in, w_b, w_out, dW_dp[n_w, np]) with rows [vec(w_in); w_b; vec(w_out)] col-major."""
    n_p = w_b.j.shape[-1]
    seed = np.concatenate([
        w_in.j.reshape(-1, n_p, order="F"), w_b.j.reshape(-1, n_p), w_out.j.reshape(-1, n_p, order="F")], axis=0)
    return w_in.v.copy(), w_b.v.copy(), w_out.v.copy(), np.asfortranarray(seed)


def _hard_threshold(x: "_D", cutoff: float, ref=None) -> "_D":
    """`w[findall(abs.(w) .< p_cutoff)] .= 0` of the pruning scripts (case1_hardthreshhold.jl:76-77, case2_pruning.jl:105-106);
    `ref` is the array the threshold is tested on when it is not `x` itself (case3_pruning.jl:243-247)."""
    if not cutoff:
        return x
    keep = (np.abs(x.v if ref is None else ref) >= cutoff).astype(np.float64)
    return _D(x.v * keep, x.j * keep[..., None])


def p2vec_case1(p, ns=5, nr=4, b0=-10.0, p_cutoff=0.0):
    """case1/case1.jl:70-78; `p_cutoff` > 0: the evaluate-only pruned variant, case1_hardthreshhold.jl:72-81."""
    d = _D.seed(p)
    w_b = d[0:nr] + b0
    w_out = _hard_threshold(d[nr:nr * (ns + 1)].reshape_f(ns, nr), p_cutoff)
    w_in = (-w_out).clamp(0.0, 2.5)
    return _pack(w_in, w_b, w_out)


def p2vec_case2(p, ns=6, nr=3, p_cutoff=0.0):
    """case2/case2.jl:91-99; last row of w_in is the Arrhenius Ea row.  `p_cutoff` > 0: case2_pruning.jl:98-117."""
    d = _D.seed(p)
    slope = d[nr * (ns + 2)] * 100.0
    slope_v = slope.broadcast_scalar((nr,))
    w_b = d[0:nr] * slope_v
    w_out = _hard_threshold(d[nr:nr * (ns + 1)].reshape_f(ns, nr), p_cutoff)
    w_in_Ea = (d[nr * (ns + 1):nr * (ns + 2)] * slope_v).abs()
    w_in = (-w_out).clamp(0.0, 4.0)
    w_in = _D(np.vstack([w_in.v, w_in_Ea.v[None, :]]), np.concatenate([w_in.j, w_in_Ea.j[None, :, :]], axis=0))
    return _pack(w_in, w_b, w_out)


def p2vec_case3(p, ns=9, nr=8, p_cutoff=0.0, dy_std=None):
    """case3/case3.jl:42-53 (p[end] is unused by the weights).  `p_cutoff` > 0: case3_pruning.jl:233-250 — w_out is tested
    after scaling each reaction's column by dy_std and normalising by its largest entry, w_in on its own magnitude."""
    d = _D.seed(p)
    w_b = d[0:nr]
    w_in_raw = d[nr * (ns + 1):nr * (2 * ns + 1)].reshape_f(ns, nr)
    w_out_raw = d[nr:nr * (ns + 1)].reshape_f(ns, nr)
    w_out = (-w_in_raw) * w_out_raw.abs()
    w_in = w_in_raw.clamp(0.0, 4.0)
    if p_cutoff:
        scaled = w_out.v * (np.ones(ns) if dy_std is None else np.asarray(dy_std, dtype=np.float64).reshape(-1))[:, None]
        scaled = scaled / scaled.max(axis=0, keepdims=True)      # maximum(w_out_, dims=2) per reaction (signed, as written)
        w_out = _hard_threshold(w_out, p_cutoff, ref=scaled)
        w_in = _hard_threshold(w_in, p_cutoff)
    return _pack(w_in, w_b, w_out)


def p2vec_robertson(p, ns=3, nr=6):
    """robertson/rober_crnn.jl:85-96."""
    d = _D.seed(p)
    slope = d[nr * (2 * ns + 1)].abs()
    w_b = d[0:nr] * (slope * 10.0).broadcast_scalar((nr,))
    w_in_raw = d[nr * (ns + 1):nr * (2 * ns + 1)].reshape_f(ns, nr)
    w_out_raw = d[nr:nr * (ns + 1)].reshape_f(ns, nr)
    w_out = (-w_in_raw) * w_out_raw.pow10()
    w_in = w_in_raw.clamp(0.0, 2.5)
    return _pack(w_in, w_b, w_out)


@dataclass
class Case:
    """Constants of one reference script."""
    name: str
    ns: int
    nr: int
    n_p: int
    rhs_kind: int
    lb: float
    ub: float
    alg: int
    abstol: object
    reltol: object
    tspan: tuple
    n_save: int
    p2vec: object
    pred_clamp: tuple
    loss_kind: int
    maxiters: int = 100000
    log_saveat: bool = False
    saveat_fn: object = None      # custom save grid (HyChem's `_tsteps`)
    model_extra: object = None    # extra CRNNModel fields (F2: gas_R, mw, tab_t, tab_T, tab_P)
    sens_mode: int = _abi.SENS_FORWARD

    def saveat(self) -> np.ndarray:
        if self.saveat_fn is not None:
            return self.saveat_fn()
        if self.log_saveat:  # tsteps = 10 .^ range(0, 5, length=datasize), rober_crnn.jl:48
            return 10.0 ** np.linspace(0.0, 5.0, self.n_save)
        return np.linspace(self.tspan[0], self.tspan[1], self.n_save)

    def model(self, p, out_scale=None):
        w_in, w_b, w_out, seed = self.p2vec(p)
        m = CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=self.rhs_kind, lb=self.lb, ub=self.ub,
                      out_scale=out_scale, **(self.model_extra or {}))
        return m, seed

    def opts(self, **kw) -> SolveOpts:
        base = dict(saveat=self.saveat(), t0=self.tspan[0], t1=self.tspan[1], alg=self.alg,
                    abstol=self.abstol, reltol=self.reltol, maxiters=self.maxiters,
                    pred_clamp=self.pred_clamp, sens_mode=self.sens_mode)
        base.update(kw)
        return SolveOpts(**base)


INF = float("inf")

# Tolerances: the scripts pass `atol=`/`rtol=`, which Julia-1.6-era OrdinaryDiffEq
# silently ignored, so the effective tolerances are the defaults
# abstol=1e-6 / reltol=1e-3.  `as_written` keeps the values in the scripts.
CASES = {
    "case1": Case("case1", 5, 4, 24, _abi.RHS_F0, 1e-5, 10.0, _abi.ALG_TSIT5, 1e-6, 1e-3,
                  (0.0, 40.0), 100, p2vec_case1, (-10.0, 10.0), _abi.LOSS_MAE_SCALED, maxiters=10000),
    "case2": Case("case2", 6, 3, 25, _abi.RHS_F1, 1e-6, 10.0, _abi.ALG_TSIT5, 1e-6, 1e-3,
                  (0.0, 50.0), 50, p2vec_case2, (-10.0, 10.0), _abi.LOSS_MAE_SCALED),
    "case3": Case("case3", 9, 8, 153, _abi.RHS_F0, 1e-5, 100.0, _abi.ALG_TSIT5, 1e-6, 1e-3,
                  (0.0, 10.0), 100, p2vec_case3, (1e-5, 100.0), _abi.LOSS_MAE_LOG),
    "robertson": Case("robertson", 3, 6, 43, _abi.RHS_F0, 1e-8, INF, _abi.ALG_ROSENBROCK23,
                      np.array([1e-6, 1e-8, 1e-6]), np.array([1e-3, 1e-3, 1e-3]),
                      (0.0, 1e5), 40, p2vec_robertson, (-INF, INF), _abi.LOSS_MAE_SCALED,
                      maxiters=10000, log_saveat=True),
}
AS_WRITTEN_TOL = {"case1": (1e-5, 1e-2), "case2": (1e-6, 1e-3), "case3": (1e-5, 1e-2)}


# ---- generating ("true") mechanisms written as CRNNs, used to make synthetic targets ----

def true_model_case2(lb=1e-30) -> CRNNModel:
    """trueODEfunc + Arrhenius of case2/case2.jl:38-59 as an F1 CRNN."""
    ns, nr = 6, 3
    logA = np.array([18.60, 19.13, 7.93]); Ea = np.array([14.54, 14.42, 6.47])
    w_in = np.zeros((ns + 1, nr)); w_out = np.zeros((ns, nr))
    # r1 = k1*TG*ROH ; r2 = k2*DG*ROH ; r3 = k3*MG*ROH
    w_in[[0, 1], 0] = 1; w_in[[2, 1], 1] = 1; w_in[[3, 1], 2] = 1
    w_in[ns, :] = Ea
    w_out[:, 0] = [-1, -1, 1, 0, 0, 1]
    w_out[:, 1] = [0, -1, -1, 1, 0, 1]
    w_out[:, 2] = [0, -1, 0, -1, 1, 1]
    return CRNNModel(w_in=w_in, w_b=logA, w_out=w_out, rhs_kind=_abi.RHS_F1, lb=lb, ub=INF)


def true_model_robertson(lb=1e-300) -> CRNNModel:
    """trueODEfunc of robertson/rober_crnn.jl:52,56-63 as an F0 CRNN (no scaling)."""
    k = np.array([4e-2, 3e7, 1e4])
    w_in = np.zeros((3, 3)); w_out = np.zeros((3, 3))
    w_in[0, 0] = 1; w_in[1, 1] = 2; w_in[[1, 2], 2] = 1
    w_out[:, 0] = [-1, 1, 0]; w_out[:, 1] = [0, -1, 1]; w_out[:, 2] = [1, -1, 0]
    return CRNNModel(w_in=w_in, w_b=np.log(k), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=lb, ub=INF)


def true_model_case1(lb=1e-30) -> CRNNModel:
    """trueODEfunc of case1/case1.jl:27,38-44 as an F0 CRNN."""
    k = np.array([0.1, 0.2, 0.13, 0.3])
    w_in = np.zeros((5, 4)); w_out = np.zeros((5, 4))
    w_in[0, 0] = 2; w_in[0, 1] = 1; w_in[2, 2] = 1; w_in[[1, 3], 3] = 1
    w_out[:, 0] = [-2, 1, 0, 0, 0]; w_out[:, 1] = [-1, 0, 1, 0, 0]
    w_out[:, 2] = [0, 0, -1, 1, 0]; w_out[:, 3] = [0, -1, 0, -1, 1]
    return CRNNModel(w_in=w_in, w_b=np.log(k), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=lb, ub=INF)


def true_model_case3(lb=1e-30) -> CRNNModel:
    """trueODEfunc of case3/case3.jl:83-103 (MAPK cascade, k = ones(8)) as an F0 CRNN."""
    ns, nr = 9, 8
    w_in = np.zeros((ns, nr)); w_out = np.zeros((ns, nr))
    pairs = [(0, 1), (2, 3), (4, 5), (6, 7)]
    for j, (a, b) in enumerate(pairs):          # r_{j+1} = y_a * y_b ; b -> b*
        w_in[[a, b], j] = 1
        w_out[b, j] = -1; w_out[b + 1, j] = 1
    for j, a in enumerate([2, 4, 6, 8]):        # r_{5..8} = y_a ; a* -> a
        w_in[a, 4 + j] = 1
        w_out[a, 4 + j] = -1; w_out[a - 1, 4 + j] = 1
    return CRNNModel(w_in=w_in, w_b=np.zeros(nr), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=lb, ub=INF)


def synthetic_stiff_model(ns=29, nr=30, seed=0, lb=1e-12) -> CRNNModel:
    """HyChem-sized synthetic CRNN (BASELINE config 5; the reference's HyChem data file is not in its
    repository, HyChem/crnn_pyrolysis_mass.jl:32): `ns` species + temperature as the last state (F1),
    `nr`/2 REVERSIBLE mass-action reactions (1 or 2 reactants <-> as many products: sum(u) is conserved,
    the state stays bounded and — being reversible — away from the lb clamp), pre-exponentials spread over
    9 decades and Arrhenius rows of 0-12 kcal/mol: a stiff system."""
    g = np.random.default_rng(seed)
    w_in = np.zeros((ns + 1, nr)); w_out = np.zeros((ns, nr))
    w_b = np.zeros(nr)
    for jf in range(0, nr - 1, 2):
        k = 1 if g.random() < 0.4 else 2
        reac = g.choice(ns, size=k, replace=False)
        prod = g.choice(np.setdiff1d(np.arange(ns), reac), size=k, replace=False)
        for j, (a_, b_) in ((jf, (reac, prod)), (jf + 1, (prod, reac))):
            for a in a_:
                w_in[a, j] += 1.0; w_out[a, j] -= 1.0
            for c in b_:
                w_out[c, j] += 1.0
        w_b[jf] = g.uniform(2.0, 19.0)                  # ln A forward
        w_b[jf + 1] = w_b[jf] - g.uniform(0.0, 5.0)     # ln A reverse
        w_in[ns, jf] = g.uniform(0.0, 12.0)             # Ea [kcal/mol]
        w_in[ns, jf + 1] = g.uniform(0.0, 12.0)
    return CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F1, lb=lb, ub=INF)


def synthetic_stiff_u0(N, ns=29, seed=1234, start=0) -> np.ndarray:
    """ICs for synthetic_stiff_model: six species at U(0.05,1), the rest at 1e-8, T ~ U(1000,1400) K."""
    from . import synth
    r = synth._blocked(seed, 7, start, N, (7,), lambda g, shp: g.random(shp))
    u0 = np.full((N, ns + 1), 1e-8)
    u0[:, :6] = 0.05 + 0.95 * r[:, :6]
    u0[:, ns] = 1000.0 + 400.0 * r[:, 6]
    return u0


def synthetic_stiff_opts(alg=_abi.ALG_KENCARP4, ns=29, n_save=40, t1=1.0) -> SolveOpts:
    """HyChem-like settings: abstol 1e-8 / reltol 1e-3 (crnn_pyrolysis_mass.jl:26-27), 40 log-spaced saves."""
    return SolveOpts(saveat=t1 * 10.0 ** np.linspace(-6.0, 0.0, n_save), t0=0.0, t1=t1, alg=alg, abstol=1e-8,
                     reltol=1e-3, maxiters=100000, obs_idx=np.arange(ns))


# ---- HyChem (JP-10 pyrolysis on mass fractions): HyChem/crnn_pyrolysis_mass.jl ----

HYCHEM_MW = np.array([136.238, 2.016, 16.043, 26.038, 28.054, 28.014, 56.108, 1.008, 15.035])  # :58
HYCHEM_GAS_R = float(np.float32(1.98720425864083e-3))  # a Float32 literal in the script (:106)


def p2vec_hychem(p, ns=9, nr=10):
    """HyChem/crnn_pyrolysis_mass.jl:78-90: w_in = [clamp(w_in_raw, 0, 2.5); Ea'; b'] (n_in = ns + 2)."""
    d = _D.seed(p)
    slope = d[nr * (2 * ns + 3)] * 10.0
    slope_v = slope.broadcast_scalar((nr,))
    w_b = d[0:nr] * slope_v
    w_in_b = d[nr:2 * nr]
    w_in_Ea = d[2 * nr:3 * nr] * slope_v
    w_out_raw = d[3 * nr:nr * (ns + 3)].reshape_f(ns, nr)
    w_in_raw = d[nr * (ns + 3):nr * (2 * ns + 3)].reshape_f(ns, nr)
    w_out = (-w_in_raw) * w_out_raw.pow10()
    w_in_c = w_in_raw.clamp(0.0, 2.5)
    w_in = _D(np.vstack([w_in_c.v, w_in_Ea.v[None, :], w_in_b.v[None, :]]),
              np.concatenate([w_in_c.j, w_in_Ea.j[None, :, :], w_in_b.j[None, :, :]], axis=0))
    return _pack(w_in, w_b, w_out)


def hychem_tables(t_end=0.01, n_tab=48):
    """Synthetic stand-in for the script's T(t), P(t) columns (its data file `data/10atm_1300K_0.01.txt`, :32, is not
    in the reference tree): a 10 atm / 1300 K pyrolysis history cooling by ~12 % as the endothermic cracking proceeds,
    on non-uniform knots (0 and a log-spaced grid, like the script's resampled `_tsteps`, :41-42)."""
    tab_t = np.concatenate([[0.0], t_end * 10.0 ** np.linspace(-5.0, 0.0, n_tab - 1)])
    s = np.sqrt(tab_t / t_end)
    tab_T = 1300.0 - 160.0 * s
    tab_P = 10.0 * 101325.0 * (1.0 + 0.02 * np.sin(3.0 * s))
    return tab_t, tab_T, tab_P


def hychem_saveat(t_end=0.01, n_save=40):
    """`_tsteps` of crnn_pyrolysis_mass.jl:41-42: 40 log-spaced points in [t_end/100, t_end/1.01], the first forced to 0."""
    ts = 10.0 ** np.linspace(np.log10(t_end / 100.0), np.log10(t_end / 1.01), n_save)
    ts[0] = 0.0
    return ts


def hychem_model(p, yscale, t_end=0.01, lb=1e-8, ns=9, nr=10):
    """(CRNNModel, seed) of the HyChem script for a parameter vector p[211]; dydt_scale = yscale / t_end (:119)."""
    w_in, w_b, w_out, seed = p2vec_hychem(p, ns, nr)
    tab_t, tab_T, tab_P = hychem_tables(t_end)
    m = CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F2, lb=lb, ub=10.0,
                  out_scale=np.asarray(yscale, dtype=np.float64) / t_end, gas_R=HYCHEM_GAS_R,
                  mw=HYCHEM_MW[:ns], tab_t=tab_t, tab_T=tab_T, tab_P=tab_P)
    return m, seed


def hychem_opts(alg=_abi.ALG_AUTO_TSIT5_ROS23, t_end=0.01, **kw) -> SolveOpts:
    """atol = lb = 1e-8, rtol = 1e-3, maxiters = 10000 (:21,26-28); tspan = [0, tsteps[sample]] (:137)."""
    ts = hychem_saveat(t_end)
    base = dict(saveat=ts, t0=0.0, t1=float(ts[-1]), alg=alg, abstol=1e-8, reltol=1e-3, maxiters=10000)
    base.update(kw)
    return SolveOpts(**base)


def hychem_u0(N, seed=1234, start=0, ns=9) -> np.ndarray:
    """Fuel (C10H16) at 3-8 % in N2, traces of the products (the script has ONE measured initial state, :73)."""
    from . import synth
    r = synth._blocked(seed, 11, start, N, (ns,), lambda g, shp: g.random(shp))
    u0 = 1e-6 * (1.0 + r)
    u0[:, 0] = 0.03 + 0.05 * r[:, 0]
    u0[:, 5] = 1.0 - u0[:, 0] - (u0[:, 1:5].sum(1) + u0[:, 6:].sum(1))
    return u0


def hychem_p(seed=0, ns=9, nr=10, sigma=0.1, slope=0.1, stiff=0.0, lnA_shift=0.0):
    """`p = randn(np) .* 0.1; p[end] = 0.1` (:75-76).  `stiff` > 0 spreads ln A over that many e-folds and switches a few
    fast consumption channels on — a stand-in for a trained stiff pyrolysis model; `lnA_shift` < 0 slows every reaction
    down by that many e-folds (a NON-stiff variant: with the script's initialisation the fastest time scale is
    ~1/50 of t_end and explicit Tsit5 runs at its stability limit)."""
    g = np.random.default_rng(seed)
    n_p = nr * (2 * ns + 3) + 1
    p = g.standard_normal(n_p) * sigma
    p[-1] = slope
    p[0:nr] += lnA_shift / (10.0 * slope)
    if stiff > 0:
        p[0:nr] += np.linspace(0.0, stiff, nr)            # ln A / slope
        w_in_raw = p[nr * (ns + 3):nr * (2 * ns + 3)].reshape(nr, ns)
        w_in_raw[:, 0] = np.abs(w_in_raw[:, 0]) + 0.5      # every reaction consumes fuel
    return p


# ---- reversible CRNN ("case1 rev/case1.jl"): an F0 model with 2*nr reactions ----

def p2vec_case1_rev(p, ns=5, nr=10):
    """`case1 rev/case1.jl:72-89`: du = w_out * (exp(w_in_f' log u + w_kf) - exp(w_in_b' log u + w_kb)) with
    w_in_f = clamp(-w_out, 0, 2.5), w_in_b = clamp(w_out, 0, 2.5), w_kb = w_kf (Kc = 1) and w_out clamped to
    [-2.5, 2.5].  That is the plain F0 CRNN with the 2*nr reactions [forward; backward]:
    w_in = [w_in_f  w_in_b], w_b = [w_kf; w_kb], w_out = [w_out  -w_out] — no extra RHS flavour is needed."""
    d = _D.seed(p)
    w_kf = d[0:nr]
    w_o = d[nr:nr * (ns + 1)].reshape_f(ns, nr).clamp(-2.5, 2.5)
    w_in_f = (-w_o).clamp(0.0, 2.5)
    w_in_b = w_o.clamp(0.0, 2.5)
    cat = lambda a, b, ax: _D(np.concatenate([a.v, b.v], axis=ax), np.concatenate([a.j, b.j], axis=ax))
    return _pack(cat(w_in_f, w_in_b, 1), cat(w_kf, w_kf, 0), cat(w_o, -w_o, 1))


def true_model_case1_rev(lb=1e-30) -> CRNNModel:
    """the generating network A<->B, B<->C, C<->D, 2C<->D+E with unit rate constants (`case1 rev/case1.jl:31-38`)"""
    ns = 5
    fw = [({0: 1}, {1: 1}), ({1: 1}, {2: 1}), ({2: 1}, {3: 1}), ({2: 2}, {3: 1, 4: 1})]
    cols_in, cols_out = [], []
    for reac, prod in fw:
        for a, b in ((reac, prod), (prod, reac)):
            wi = np.zeros(ns); wo = np.zeros(ns)
            for i, nu in a.items():
                wi[i] = nu; wo[i] -= nu
            for i, nu in b.items():
                wo[i] += nu
            cols_in.append(wi); cols_out.append(wo)
    return CRNNModel(w_in=np.array(cols_in).T, w_b=np.zeros(len(cols_in)), w_out=np.array(cols_out).T,
                     rhs_kind=_abi.RHS_F0, lb=lb, ub=INF)


CASES["case1_rev"] = Case("case1_rev", 5, 20, 60, _abi.RHS_F0, 1e-5, INF, _abi.ALG_TSIT5, 1e-6, 1e-3,
                          (0.0, 10.0), 100, p2vec_case1_rev, (-INF, INF), _abi.LOSS_MAE_SCALED, maxiters=10000,
                          # a parameter of the reversible CRNN touches TWO w_out entries (forward and reverse reaction): its
                          # seed columns are not of the structured shape the forward kernels take, so the front end asks for
                          # the discrete adjoint (the forward-mode derivative of the value-norm solve)
                          sens_mode=_abi.SENS_DISCRETE_ADJOINT)


def hychem_case(t_end=0.01, alg=_abi.ALG_TSIT5, sens_mode=_abi.SENS_DISCRETE_ADJOINT) -> Case:
    """HyChem/crnn_pyrolysis_mass.jl as a `Case` for the front-end mirror (`CRNNProblem(hychem_case(), ...)`, with
    `out_scale = yscale / t_end`): F2 RHS under the synthetic T(t), P(t) tables, the script's tolerances and save grid.
    The gradient of its 211 parameters comes from the adjoint kernels (Tsit5), see DESIGN.md §3.2d / §8."""
    tab_t, tab_T, tab_P = hychem_tables(t_end)
    ts = hychem_saveat(t_end)
    return Case("hychem", 9, 10, 211, _abi.RHS_F2, 1e-8, 10.0, alg, 1e-8, 1e-3, (0.0, float(ts[-1])), 40, p2vec_hychem,
                (-INF, INF), _abi.LOSS_MAE_SCALED, maxiters=10000, saveat_fn=lambda: hychem_saveat(t_end),
                model_extra=dict(gas_R=HYCHEM_GAS_R, mw=HYCHEM_MW, tab_t=tab_t, tab_T=tab_T, tab_P=tab_P),
                sens_mode=sens_mode)


# ---- gene-regulatory network (gene-regulatory-network/gene-regulatory.jl): 9 species, 15 reactions, np = 285 ----

def p2vec_gene(p, ns=9, nr=15):
    """gene-regulatory.jl:39-50: like case3 (w_out = -w_in_raw * |w_out_raw|, w_in = clamp(w_in_raw, 0, 4)) with the
    three DNA rows of w_out_raw zeroed (`w_out[[1, 4, 7], :] .= 0`: the catalysts are never consumed or produced)."""
    d = _D.seed(p)
    w_b = d[0:nr]
    w_in_raw = d[nr * (ns + 1):nr * (2 * ns + 1)].reshape_f(ns, nr)
    w_out_raw = d[nr:nr * (ns + 1)].reshape_f(ns, nr)
    mask = np.ones((ns, nr)); mask[[0, 3, 6], :] = 0.0
    w_out_raw = _D(w_out_raw.v * mask, w_out_raw.j * mask[..., None])
    w_out = (-w_in_raw) * w_out_raw.abs()
    w_in = w_in_raw.clamp(0.0, 4.0)
    return _pack(w_in, w_b, w_out)


def true_model_gene(lb=1e-30) -> CRNNModel:
    """trueODEfunc of gene-regulatory.jl:75-131 (k at :141) as an F0 CRNN: transcription / translation with the template
    as a catalyst, first-order decays, and the cyclic repression mRNA_i + protein_j -> protein_j."""
    ns, nr = 9, 15
    k = np.array([1.8, 2.1, 1.3, 1.5, 2.2, 2, 2, 2.5, 3.2, 3, 2.3, 2.5, 6, 4, 3])
    w_in = np.zeros((ns, nr)); w_out = np.zeros((ns, nr))
    for g in range(3):                      # gene g: DNA 3g, mRNA 3g+1, protein 3g+2; reactions 4g .. 4g+3
        dna, mrna, prot = 3 * g, 3 * g + 1, 3 * g + 2
        w_in[dna, 4 * g] = 1; w_out[mrna, 4 * g] = 1            # DNA -> DNA + mRNA
        w_in[mrna, 4 * g + 1] = 1; w_out[prot, 4 * g + 1] = 1   # mRNA -> mRNA + protein
        w_in[mrna, 4 * g + 2] = 1; w_out[mrna, 4 * g + 2] = -1  # mRNA -> 0
        w_in[prot, 4 * g + 3] = 1; w_out[prot, 4 * g + 3] = -1  # protein -> 0
    for j, (mrna, prot) in enumerate(((7, 2), (4, 8), (1, 5))):  # R13 = k y8 y3, R14 = k y5 y9, R15 = k y2 y6 (1-based)
        w_in[[mrna, prot], 12 + j] = 1; w_out[mrna, 12 + j] = -1
    return CRNNModel(w_in=w_in, w_b=np.log(k), w_out=w_out, rhs_kind=_abi.RHS_F0, lb=lb, ub=INF)


# as written: atol 1e-5 / rtol 1e-2 under the silently ignored keywords -> the solver defaults
CASES["gene"] = Case("gene", 9, 15, 285, _abi.RHS_F0, 1e-5, 100.0, _abi.ALG_TSIT5, 1e-6, 1e-3,
                     (0.0, 4.0), 40, p2vec_gene, (1e-5, 100.0), _abi.LOSS_MAE_SCALED, sens_mode=_abi.SENS_DISCRETE_ADJOINT)


# ---------------------------------------------------------------------------------------------------------------------
# Cathode thermal decomposition (DSC): three sequential reactions c1 -> c2 -> c3 -> products under a linear temperature
# ramp, observed through the heat release.  RHS flavour F5 + observable post-map.
#   Cathode/src/network.jl                     deterministic fit, MAE loss, clamped p2vec (:27-50)
#   Cathode_NCM333_UQ/src_333/network.jl       SVGD: 100 particles x 5 heating rates, linear p2vec with p_scales, MSE loss
# ---------------------------------------------------------------------------------------------------------------------
CATHODE_GAS_R = 8.314          # `const R = -1.0 / 8.314` (Cathode/src/network.jl:67): x = R/T = -1/(8.314 T)
CATHODE_T0 = 100.0 + 273.15    # K (network.jl:106, src_333/network.jl:196)


def _pack_cathode(order: _D, Ea: _D, b: _D, lnA: _D, nu: _D, delH: _D):
    """-> (w_in [5,3], w_b [3], w_out [3,3], w_obs [3], dW_dp [n_w = 30, np]) of the F5 form of `crnn!` / `HRR_getter`:
    w_in = [diag(order); (Ea 1e5)'; b'], w_b = ln A, w_out = [[-1,0,0],[nu2,-1,0],[0,nu3,-1]], w_obs = delH."""
    n_p = lnA.j.shape[-1]
    w_in = np.zeros((5, 3)); j_in = np.zeros((5, 3, n_p))
    for k in range(3):
        w_in[k, k] = order.v[k]; j_in[k, k] = order.j[k]
        w_in[3, k] = Ea.v[k] * 1e5; j_in[3, k] = Ea.j[k] * 1e5
        w_in[4, k] = b.v[k]; j_in[4, k] = b.j[k]
    w_out = -np.eye(3); j_out = np.zeros((3, 3, n_p))
    w_out[1, 0] = nu.v[0]; j_out[1, 0] = nu.j[0]     # du[2] += w_out[2] * rxn_rates[1]
    w_out[2, 1] = nu.v[1]; j_out[2, 1] = nu.j[1]     # du[3] += w_out[3] * rxn_rates[2]
    seed = np.concatenate([j_in.reshape(-1, n_p, order="F"), lnA.j.reshape(-1, n_p), j_out.reshape(-1, n_p, order="F"),
                           delH.j.reshape(-1, n_p)], axis=0)
    return w_in, lnA.v.copy(), w_out, delH.v.copy(), np.asfortranarray(seed)


def p2vec_cathode(p):
    """Cathode/src/network.jl:27-50 (18 parameters, slope last, clamps as written)."""
    d = _D.seed(p)
    slope = d[17] * 10.0
    lnA = (d[0:3] * (slope * 20.0).broadcast_scalar((3,))).clamp(0.0, 50.0)
    nu = d[15:17].clamp(0.01, 5.0)
    order = d[12:15].clamp(0.01, 10.0)
    Ea = d[3:6].abs().clamp(0.0, 3.0)
    b = d[6:9]
    delH = (d[9:12].abs() * 100.0).clamp(10.0, 300.0)
    return _pack_cathode(order, Ea, b, lnA, nu, delH)


def p2vec_cathode_uq(p, p_scales):
    """Cathode_NCM333_UQ/src_333/network.jl:93-108,153-168: linear in p, every entry scaled by p_scales (17 parameters)."""
    d = _D.seed(p)
    sc = np.asarray(p_scales, dtype=np.float64)
    scaled = lambda a, b_: _D(d.v[a:b_] * sc[a:b_], d.j[a:b_] * sc[a:b_, None])
    lnA, Ea, b, delH, order, nu = scaled(0, 3), scaled(3, 6), scaled(6, 9), scaled(9, 12), scaled(12, 15), scaled(15, 17)
    return _pack_cathode(order, Ea, b, lnA, nu, delH)


def cathode_ramp(beta_K_per_min, t_end, T0=CATHODE_T0):
    """`getsampletemp`: T = T0 + beta/60 t (network.jl:59-64) as a two-knot table."""
    return np.array([0.0, t_end]), np.array([T0, T0 + beta_K_per_min / 60.0 * t_end])


def cathode_model(w_in, w_b, w_out, w_obs, beta, t_end, lb=1e-8):
    tab_t, tab_T = cathode_ramp(beta, t_end)
    return CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, w_obs=w_obs, rhs_kind=_abi.RHS_F5, lb=lb, ub=10.0,
                     gas_R=CATHODE_GAS_R, tab_t=tab_t, tab_T=tab_T)


def cathode_opts(ts, alg=_abi.ALG_AUTO_TSIT5_ROS23, lb=1e-8, **kw) -> SolveOpts:
    """`ODEProblem(crnn!, u0, tspan, p, abstol = lb)`, `saveat = ts` (network.jl:96,103-116); `alg = ALG_AUTO_TSIT5_TRBDF2` is the
    script's own `AutoTsit5(TRBDF2(autodiff = true))` (:102), the default its Rosenbrock23 sibling."""
    ts = np.asarray(ts, dtype=np.float64)
    base = dict(saveat=ts, t0=float(ts[0]), t1=float(ts[-1]), alg=alg, abstol=lb, reltol=1e-3, maxiters=100000,
                obs_idx=np.array([0]))
    base.update(kw)
    return SolveOpts(**base)


def cathode_p_true():
    """A physically plausible parameter set in the UQ script's (scaled) coordinates: ln A, Ea [1e5 J/mol], b, delH, orders, nu."""
    return np.array([28.0, 30.0, 33.0, 1.25, 1.40, 1.60, 0.0, 0.0, 0.0, 120.0, 40.0, 60.0, 1.0, 1.2, 1.0, 0.9, 0.8])


# ---- yeast glycolysis (yeast-glycolysis/yeast_glycolysis.jl): 7 observed + 5 hidden species, the hidden ones from an MLP (F4) ----

YEAST_NS, YEAST_NS_, YEAST_NR = 7, 12, 12          # :29-31
YEAST_MLP_DIMS = (7, 5, 5, 5, 5)                   # Chain(Dense(ns, node, gelu), Dense(node, node, gelu) x 2, Dense(node, ns_ - ns, softplus)), :137-142
YEAST_IC_LB = np.array([0.15, 1.19, 0.04, 0.10, 0.08, 0.14, 0.05])   # :70-71
YEAST_IC_UB = np.array([1.60, 2.16, 0.20, 0.35, 0.30, 2.67, 0.10])


def yeast_true_rhs(t, s, k=(100.0, 6.0, 16.0, 100.0, 1.28, 12.0)):
    """trueODEfunc, yeast_glycolysis.jl:47-66 (q = 4, K1 = 0.52, A = 4, N = 1, J0 = 2.5, phi = 0.1; k at :78)"""
    q, K1, A, N, J0, phi = 4, 0.52, 4.0, 1.0, 2.5, 0.1
    r1 = k[0] * s[0] * s[5] / (1 + (s[5] / K1) ** q)
    r2 = k[1] * s[1] * (N - s[4])
    r3 = k[2] * s[2] * (A - s[5])
    r4 = k[3] * s[3] * s[4]
    r5 = k[4] * s[5]
    r6 = k[5] * s[1] * s[4]
    r7 = 13 * s[6]
    r8 = 13 * (s[3] - s[6])
    return np.array([J0 - r1, 2 * r1 - r2 - r6, r2 - r3, r3 - r4 - r8, r2 - r4 - r6, -2 * r1 + 2 * r3 - r5, phi * r8 - r7])


def p2vec_yeast(p):
    """yeast_glycolysis.jl:112-123,145: p = vcat(pcrnn [164], pnn [130]); slope = pcrnn[end] * 100, w_b = pcrnn[1:nr] * slope,
    w_out = reshape(pcrnn[nr+1 : nr (ns_+1)], ns_, nr), w_in = clamp(-w_out, 0, 4), w_J = pcrnn[nr (ns_+1)+1 : end-1];
    pnn in Flux.destructure order.  -> (w_in [12,12], w_b, w_out [12,12], w_J [7], pnn)."""
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    ns, ns_, nr = YEAST_NS, YEAST_NS_, YEAST_NR
    n_crnn = nr * (ns_ + 1) + ns + 1
    pc, pnn = p[:n_crnn], p[n_crnn:]
    slope = pc[-1] * 100.0
    w_b = pc[:nr] * slope
    w_out = pc[nr:nr * (ns_ + 1)].reshape(ns_, nr, order="F")
    w_in = np.clip(-w_out, 0.0, 4.0)
    w_J = pc[nr * (ns_ + 1):n_crnn - 1]
    return w_in, w_b, w_out, w_J, pnn


def yeast_seed(p):
    """dW/dp of the yeast script's parameter vector in the adjoint's weight space [vec(w_in) 12x12; w_b; vec(w_out[1:ns, :]); w_J;
    pnn]: p2vec (:112-123) differentiated with ForwardDiff's clamp convention, the MLP parameters map to themselves."""
    ns, ns_, nr = YEAST_NS, YEAST_NS_, YEAST_NR
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    n_crnn = nr * (ns_ + 1) + ns + 1
    d = _D.seed(p)
    slope = d[n_crnn - 1] * 100.0
    w_b = d[0:nr] * slope.broadcast_scalar((nr,))
    w_out = d[nr:nr * (ns_ + 1)].reshape_f(ns_, nr)
    w_in = (-w_out).clamp(0.0, 4.0)
    w_J = d[nr * (ns_ + 1):n_crnn - 1]
    pnn = d[n_crnn:]
    n_p = p.size
    return np.concatenate([w_in.j.reshape(-1, n_p, order="F"), w_b.j.reshape(-1, n_p), w_out.j[:ns].reshape(-1, n_p, order="F"),
                           w_J.j.reshape(-1, n_p), pnn.j.reshape(-1, n_p)], axis=0)


def yeast_model(p, lb=1e-5, ub=100.0) -> CRNNModel:
    """`crnn` of yeast_glycolysis.jl:128-132 as an F4 model: u_ = vcat(u, rep(u)), du = (w_out * exp(w_in' log clamp(u_) + w_b))[1:ns]
    .+ w_J — only the first ns rows of w_out enter (lb = atol = 1e-5, ub = 100, :34-37)."""
    w_in, w_b, w_out, w_J, pnn = p2vec_yeast(p)
    ns, ns_ = YEAST_NS, YEAST_NS_
    return CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out[:ns], rhs_kind=_abi.RHS_F4, lb=lb, ub=ub, mlp_dims=np.array(YEAST_MLP_DIMS),
                     mlp_in_idx=np.arange(ns), mlp_params=pnn, mlp_act_out=0,
                     aug_src=np.concatenate([np.arange(ns), -1 - np.arange(ns_ - ns)]), w_J=w_J)


def yeast_opts(alg=_abi.ALG_AUTO_TSIT5_TRBDF2, n_save=300, t1=5.0, abstol=1e-6, reltol=1e-3, **kw) -> SolveOpts:
    """tspan [0, 5], 300 saves (:24-27,74-75); AutoTsit5(TRBDF2(autodiff=false)) (:33); the script's `atol=`/`rtol=` keywords
    fall back to the solver defaults; predictions clamped to [lb, ub] (:155)"""
    kw.setdefault("pred_clamp", (1e-5, 100.0))
    return SolveOpts(saveat=np.linspace(0.0, t1, n_save), t0=0.0, t1=t1, alg=alg, abstol=abstol, reltol=reltol,
                     maxiters=100000, obs_idx=np.arange(YEAST_NS), **kw)


@dataclass
class _YeastCase(Case):
    """yeast_glycolysis.jl as a Case: F4 model, gradient of all 294 parameters (164 CRNN + 130 MLP) by the adjoint kernels over the
    extended weight space; `alg` is Tsit5 — the non-stiff half of the script's AutoTsit5(TRBDF2), which the gradient path serves
    (predictions with the composite: `opts(alg=ALG_AUTO_TSIT5_TRBDF2)`)."""

    def model(self, p, out_scale=None):
        return yeast_model(p, self.lb, self.ub), yeast_seed(p)


CASES["yeast"] = _YeastCase("yeast", YEAST_NS, YEAST_NR, 294, _abi.RHS_F4, 1e-5, 100.0, _abi.ALG_TSIT5, 1e-6, 1e-3, (0.0, 5.0), 300,
                            p2vec=lambda p: (*p2vec_yeast(p)[:3], None), pred_clamp=(1e-5, 100.0), loss_kind=_abi.LOSS_MAE_SCALED,
                            sens_mode=_abi.SENS_DISCRETE_ADJOINT)


def mlp_reference(dims, params, x, act_out=0):
    """numpy restatement of the Flux chain (for tests): gelu hidden layers (NNlib's tanh form), softplus / exp output"""
    a = np.asarray(x, dtype=np.float64)
    off = 0
    L = len(dims) - 1
    for l in range(L):
        din, dout = dims[l], dims[l + 1]
        W = params[off:off + din * dout].reshape(dout, din, order="F"); off += din * dout
        b = params[off:off + dout]; off += dout
        s = W @ a + b
        if l + 1 < L:
            a = 0.5 * s * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (s + 0.044715 * s ** 3)))
        else:
            a = np.log1p(np.exp(-np.abs(s))) + np.maximum(s, 0.0) if act_out == 0 else np.exp(s)
    return a


def mlp_rows_postmap(model, pred, obs_idx=None):
    """`pred[2:2, :] .= rep(pred[[1, 3], :])` of rober_crnn_qssa.jl:139 (and the hidden-species read-out `rep(pred)` of
    yeast_glycolysis.jl:170-171): the Flux chain evaluated on the saved states of an F4 model, host side.
    pred [N, n_save, n_obs] with all state rows observed (or `obs_idx` naming them) -> (mlp_out [N, n_save, n_mlp_out], pred with
    every state row that `aug_src` feeds from the MLP at the SAME position replaced by that output — the QSSA convention)."""
    pred = np.asarray(pred, dtype=np.float64)
    obs = np.arange(model.n_state) if obs_idx is None else np.asarray(obs_idx)
    col = {int(r): k for k, r in enumerate(obs)}
    if any(int(r) not in col for r in model.mlp_in_idx):
        raise ValueError("the MLP's input rows must be among the observed rows")
    x = pred[..., [col[int(r)] for r in model.mlp_in_idx]]
    flat = x.reshape(-1, x.shape[-1])
    out = np.array([mlp_reference(tuple(int(d) for d in model.mlp_dims), model.mlp_params, v, model.mlp_act_out) for v in flat])
    out = out.reshape(pred.shape[:-1] + (out.shape[-1],))
    mapped = pred.copy()
    for q, src in enumerate(model.aug_src):      # input row q of the CRNN sits at state position q when n_in == n_state (QSSA)
        if src < 0 and model.n_in == model.n_state and q in col:
            mapped[..., col[q]] = out[..., -1 - int(src)]
    return out, mapped
