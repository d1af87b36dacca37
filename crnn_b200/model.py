"""Host-side descriptions of a CRNN model and of one `solve` call.

`CRNNModel` is what the reference scripts' `p2vec` returns (w_in, w_b, w_out —
column-major, e.g. case2/case2.jl:91-99) plus the constants their RHS closes
over (lb, ub, dydt_scale / dy_std_, the gas constant).  `SolveOpts` is the
keyword set of `ODEProblem(...)`/`solve(...)` (case2/case2.jl:121,126).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import CModel, COpts, _f64, dptr, iptr

GAS_R = 1.98720425864083e-3  # case2/case2.jl:113


@dataclass
class CRNNModel:
    w_in: np.ndarray          # [n_in, n_reac]
    w_b: np.ndarray           # [n_reac]
    w_out: np.ndarray         # [n_species, n_reac]
    rhs_kind: int = _abi.RHS_F0
    lb: float = 1e-5
    ub: float = 10.0
    out_scale: np.ndarray | None = None
    gas_R: float = GAS_R
    # F2 only (HyChem/crnn_pyrolysis_mass.jl:58,103-104): molar masses and the T(t), P(t) tables
    mw: np.ndarray | None = None
    tab_t: np.ndarray | None = None
    tab_T: np.ndarray | None = None
    tab_P: np.ndarray | None = None
    # observable post-map: y = sum_j w_obs[j] r_j (heat release, Cathode/src/network.jl:82-91,121); None: rows of u
    w_obs: np.ndarray | None = None
    # F4 (yeast_glycolysis.jl:128-142, rober_crnn_qssa.jl:111-126): a Flux MLP of the state supplies the hidden input rows
    mlp_dims: np.ndarray | None = None      # [L + 1] layer widths d0 .. d_L
    mlp_in_idx: np.ndarray | None = None    # [d0] state rows fed to the MLP
    mlp_params: np.ndarray | None = None    # Flux.destructure order: per layer W (d_out x d_in, column-major), b
    mlp_act_out: int = 0                    # 0 softplus, 1 exp; hidden layers: gelu
    aug_src: np.ndarray | None = None       # [n_in]: >= 0 state row, < 0: MLP output -1 - value
    w_J: np.ndarray | None = None           # [n_species] additive source term

    def __post_init__(self):
        self.w_in = np.asarray(self.w_in, dtype=np.float64)
        self.w_b = np.asarray(self.w_b, dtype=np.float64).reshape(-1)
        self.w_out = np.asarray(self.w_out, dtype=np.float64)
        if self.out_scale is not None:
            self.out_scale = np.asarray(self.out_scale, dtype=np.float64).reshape(-1)
        n_in, nr = self.w_in.shape
        ns, nr2 = self.w_out.shape
        if nr != nr2 or self.w_b.shape[0] != nr:
            raise ValueError("w_in, w_b, w_out disagree on n_reac")
        if self.rhs_kind == _abi.RHS_F0 and n_in != ns:
            raise ValueError("F0 needs n_in == n_species")
        if self.rhs_kind == _abi.RHS_F1 and n_in != ns + 1:
            raise ValueError("F1 needs n_in == n_species + 1 (Arrhenius row)")
        if self.out_scale is not None and self.out_scale.shape[0] != ns:
            raise ValueError("out_scale must have n_species entries")
        if self.w_obs is not None:
            self.w_obs = np.ascontiguousarray(self.w_obs, dtype=np.float64).reshape(-1)
            if self.w_obs.shape[0] != nr:
                raise ValueError("w_obs must have n_reac entries")
        if self.rhs_kind == _abi.RHS_F4:
            self.mlp_dims = np.ascontiguousarray(self.mlp_dims, dtype=np.int32).reshape(-1)
            self.mlp_in_idx = np.ascontiguousarray(self.mlp_in_idx, dtype=np.int32).reshape(-1)
            self.mlp_params = np.ascontiguousarray(self.mlp_params, dtype=np.float64).reshape(-1)
            self.aug_src = np.ascontiguousarray(self.aug_src, dtype=np.int32).reshape(-1)
            d = self.mlp_dims
            if d.size < 2 or d.max() > 32 or self.mlp_in_idx.size != d[0] or self.aug_src.size != n_in:
                raise ValueError("F4: mlp_dims [L+1] (<= 32 each), mlp_in_idx [d0], aug_src [n_in]")
            if self.mlp_params.size != int(sum(d[l] * d[l + 1] + d[l + 1] for l in range(d.size - 1))):
                raise ValueError("F4: mlp_params must hold W and b of every layer (Flux.destructure order)")
            if self.aug_src.max() >= ns or self.aug_src.min() < -int(d[-1]) or self.mlp_in_idx.max() >= ns:
                raise ValueError("F4: aug_src / mlp_in_idx out of range")
            if self.w_J is not None:
                self.w_J = np.ascontiguousarray(self.w_J, dtype=np.float64).reshape(-1)
        if self.rhs_kind == _abi.RHS_F5:
            if n_in != ns + 2:
                raise ValueError("F5 needs n_in == n_species + 2 (Arrhenius and log T rows)")
            if self.tab_t is None or self.tab_T is None:
                raise ValueError("F5 needs the tab_t / tab_T temperature programme")
            self.tab_t, self.tab_T = (np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (self.tab_t, self.tab_T))
            self.tab_P = np.ones_like(self.tab_T)
            self.mw = np.ones(ns)
            if not (self.tab_t.size == self.tab_T.size >= 2) or np.any(np.diff(self.tab_t) <= 0):
                raise ValueError("F5: tab_t / tab_T must have equal length >= 2, tab_t strictly ascending")
        if self.rhs_kind == _abi.RHS_F2:
            if n_in != ns + 2:
                raise ValueError("F2 needs n_in == n_species + 2 (Arrhenius and log T rows)")
            if self.mw is None or self.tab_t is None or self.tab_T is None or self.tab_P is None:
                raise ValueError("F2 needs mw and the tab_t / tab_T / tab_P tables")
            self.mw = np.ascontiguousarray(self.mw, dtype=np.float64).reshape(-1)
            self.tab_t, self.tab_T, self.tab_P = (np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
                                                  for a in (self.tab_t, self.tab_T, self.tab_P))
            if self.mw.shape[0] != ns or not (self.tab_t.size == self.tab_T.size == self.tab_P.size >= 2):
                raise ValueError("F2: mw must have n_species entries and the tables equal length >= 2")
            if np.any(np.diff(self.tab_t) <= 0):
                raise ValueError("F2: tab_t must be strictly ascending")

    @property
    def n_species(self) -> int:
        return self.w_out.shape[0]

    @property
    def n_reac(self) -> int:
        return self.w_out.shape[1]

    @property
    def n_in(self) -> int:
        return self.w_in.shape[0]

    @property
    def n_state(self) -> int:
        return self.n_species + (1 if self.rhs_kind == _abi.RHS_F1 else 0)

    @property
    def n_w(self) -> int:
        n = self.n_reac * (self.n_in + 1 + self.n_species) + (self.n_reac if self.w_obs is not None else 0)
        if self.rhs_kind == _abi.RHS_F4:   # the adjoint's extended weight space: + w_J + the MLP parameters
            n += self.n_species + self.mlp_params.size
        return n

    def flat_weights(self) -> np.ndarray:
        """[vec(w_in); w_b; vec(w_out); (w_obs) | (F4: w_J; mlp_params)] column-major: the row order of dW_dp."""
        parts = [self.w_in.reshape(-1, order="F"), self.w_b, self.w_out.reshape(-1, order="F")]
        if self.w_obs is not None:
            parts.append(self.w_obs)
        if self.rhs_kind == _abi.RHS_F4:
            parts += [np.zeros(self.n_species) if self.w_J is None else self.w_J, self.mlp_params]
        return np.concatenate(parts)

    def to_c(self):
        """Returns (CModel, keepalive)."""
        keep = [
            np.asfortranarray(self.w_in).reshape(-1, order="F").copy(),
            _f64(self.w_b),
            np.asfortranarray(self.w_out).reshape(-1, order="F").copy(),
            None if self.out_scale is None else _f64(self.out_scale),
        ]
        m = CModel()
        m.n_state, m.n_species, m.n_in, m.n_reac = self.n_state, self.n_species, self.n_in, self.n_reac
        m.rhs_kind = self.rhs_kind
        m.lb, m.ub, m.gas_R = float(self.lb), float(self.ub), float(self.gas_R)
        m.w_in, m.w_b, m.w_out = dptr(keep[0]), dptr(keep[1]), dptr(keep[2])
        m.out_scale = dptr(keep[3])
        if self.w_obs is not None:
            keep.append(self.w_obs)
            m.w_obs = dptr(self.w_obs)
        if self.rhs_kind == _abi.RHS_F4:
            keep += [self.mlp_dims, self.mlp_in_idx, self.mlp_params, self.aug_src, self.w_J]
            m.mlp_n_layers, m.mlp_act_out = int(self.mlp_dims.size - 1), int(self.mlp_act_out)
            m.mlp_dims, m.mlp_in_idx, m.aug_src = iptr(self.mlp_dims), iptr(self.mlp_in_idx), iptr(self.aug_src)
            m.mlp_params, m.w_J = dptr(self.mlp_params), dptr(self.w_J)
        if self.rhs_kind in (_abi.RHS_F2, _abi.RHS_F5):
            keep += [self.mw, self.tab_t, self.tab_T, self.tab_P]
            m.n_tab = int(self.tab_t.size)
            m.mw, m.tab_t, m.tab_T, m.tab_P = dptr(self.mw), dptr(self.tab_t), dptr(self.tab_T), dptr(self.tab_P)
        return m, keep


@dataclass
class SolveOpts:
    saveat: np.ndarray
    t0: float
    t1: float
    alg: int = _abi.ALG_TSIT5
    abstol: float | np.ndarray = 1e-6   # OrdinaryDiffEq defaults (what `atol=`/`rtol=` silently fell back to)
    reltol: float | np.ndarray = 1e-3
    maxiters: int = 100000
    obs_idx: np.ndarray | None = None   # default: every row of u
    pred_clamp: tuple[float, float] = (-np.inf, np.inf)
    sens_mode: int = _abi.SENS_FORWARD
    err_norm_includes_sens: bool = True
    # True: DiffEqBase's norm over Dual arrays, mean over n_state*(1+np) numbers; False: over n_state rows
    err_norm_mean_over_partials: bool = True
    controller: dict = field(default_factory=dict)  # qmin,qmax,gamma,beta1,beta2,qsteady_min,qsteady_max overrides

    def to_c(self, n_state: int, buffers_on_device: bool = False, stream: int = 0):
        saveat = _f64(self.saveat).reshape(-1)
        if saveat.size and (np.any(np.diff(saveat) < 0) or saveat[0] < self.t0 or saveat[-1] > self.t1):
            raise ValueError("saveat must be ascending and inside [t0, t1]")
        abstol = _f64(np.atleast_1d(self.abstol))
        reltol = _f64(np.atleast_1d(self.reltol))
        for name, a in (("abstol", abstol), ("reltol", reltol)):
            if a.size not in (1, n_state):
                raise ValueError(f"{name} must have 1 or n_state entries")
        obs = np.arange(n_state, dtype=np.int32) if self.obs_idx is None else \
            np.ascontiguousarray(np.asarray(self.obs_idx, dtype=np.int32))
        if obs.size and (obs.min() < 0 or obs.max() >= n_state):
            raise ValueError("obs_idx out of range")
        o = COpts()
        o.alg, o.sens_mode = int(self.alg), int(self.sens_mode)
        o.err_norm_includes_sens = int(bool(self.err_norm_includes_sens))
        o.n_save, o.n_obs = saveat.size, obs.size
        o.n_abstol, o.n_reltol = abstol.size, reltol.size
        o.buffers_on_device = int(buffers_on_device)
        o.maxiters = int(self.maxiters)
        o.t0, o.t1 = float(self.t0), float(self.t1)
        o.pred_clamp_lo, o.pred_clamp_hi = float(self.pred_clamp[0]), float(self.pred_clamp[1])
        o.abstol, o.reltol, o.saveat, o.obs_idx = dptr(abstol), dptr(reltol), dptr(saveat), iptr(obs)
        for k in ("qmin", "qmax", "gamma", "beta1", "beta2", "qsteady_min", "qsteady_max"):
            setattr(o, k, float(self.controller.get(k, 0.0)))
        o.err_norm_mean_over_state_only = int(not self.err_norm_mean_over_partials)
        o.stream = C.c_void_p(stream)
        return o, [saveat, abstol, reltol, obs]

    @property
    def n_save(self) -> int:
        return int(np.asarray(self.saveat).size)

    def n_obs(self, n_state: int) -> int:
        return n_state if self.obs_idx is None else int(np.asarray(self.obs_idx).size)
