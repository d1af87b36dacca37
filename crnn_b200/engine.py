"""Python face of the C-ABI (libcrnn_b200.so): `Engine.solve_batch` / `Engine.loss_grad_batch`.

Host-side mirror of the two reference call sites the engine replaces:
`predict_neuralode(u0, p)` (case2/case2.jl:124-128) and
`ForwardDiff.gradient(x -> loss_neuralode(x, i_exp), p)` (case2/case2.jl:195), batched.

Inputs are either numpy arrays (host buffers: the library runs the chunked
H2D -> kernel -> D2H pipeline itself) or torch CUDA tensors (device buffers: the
call only enqueues kernels on the current torch stream).  Layout is
trajectory-major: u0 [N, n_state], data/pred [N, n_save, n_obs] (C order), which
is the C-ABI's column-major u0[n_state, N], data[n_obs, n_save, N].
There is no CPU fallback: constructing an Engine without the CUDA library or a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import STATS_DTYPE
from .model import CRNNModel, SolveOpts


class EngineError(RuntimeError):
    pass


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Dataset:
    """Device-resident u0_list / ode_data_list of a script (case2/case2.jl:62-83), uploaded once
    (`crnn_dataset_create`); split into contiguous shards on a multi-device Engine."""

    def __init__(self, engine: "Engine", u0, data):
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        data = np.ascontiguousarray(data, dtype=np.float64)
        if u0.ndim != 2 or data.ndim != 3 or data.shape[0] != u0.shape[0]:
            raise ValueError("u0 must be [N, n_state] and data [N, n_save, n_obs]")
        self.engine, self.N = engine, u0.shape[0]
        self.n_state, self.n_save, self.n_obs = u0.shape[1], data.shape[1], data.shape[2]
        d = C.c_void_p()
        engine._check(engine._lib.crnn_dataset_create(engine._h, u0.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                                                      self.n_state, self.n_obs, self.n_save, self.N, C.byref(d)))
        self._d = d

    def close(self):
        if getattr(self, "_d", None) and getattr(self.engine, "_h", None):
            self.engine._lib.crnn_dataset_destroy(self._d)
        self._d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.N


class Engine:
    def __init__(self, device: int = -1, devices=None):
        """`device`: one CUDA device (-1 = current).  `devices`: a list of device ids (or an int n = devices 0..n-1) for
        ONE handle over several GPUs of this process (`crnn_create_multi`: shards + NCCL all-reduce of the gradient)."""
        self._lib = _abi.load_library()
        h = C.c_void_p()
        if devices is not None:
            ids = np.arange(int(devices), dtype=np.int32) if np.isscalar(devices) else np.ascontiguousarray(devices, dtype=np.int32)
            rc = self._lib.crnn_create_multi(C.byref(h), ids.ctypes.data_as(C.c_void_p), int(ids.size))
            what = f"crnn_create_multi({ids.tolist()})"
        else:
            rc = self._lib.crnn_create(C.byref(h), int(device))
            what = "crnn_create"
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): no usable CUDA device (or NCCL); crnn_b200 has no CPU path")
        self._h = h

    @property
    def n_devices(self) -> int:
        return int(self._lib.crnn_device_count(self._h))

    def dataset(self, u0, data) -> Dataset:
        return Dataset(self, u0, data)

    def loss_grad_indexed(self, model: CRNNModel, opts: SolveOpts, seed, ds: Dataset, yscale,
                          loss_kind=_abi.LOSS_MAE_SCALED, idx=None, n_save_used=None, want_loss=False, want_stats=False):
        """`crnn_loss_grad_indexed`: loss + gradient over rows `idx` (None = all) of a device-resident dataset.
        -> dict(loss_sum, n_ok, loss_mean, grad_sum [np], and loss / n_saved / retcode / stats when asked for)."""
        n_obs, n = opts.n_obs(model.n_state), model.n_state
        cm, k1 = model.to_c()
        co, k2 = opts.to_c(n, False)
        seed = np.asarray(seed, dtype=np.float64)
        if seed.ndim != 2 or seed.shape[0] != model.n_w:
            raise ValueError(f"seed must be [n_w={model.n_w}, np]")
        n_p = seed.shape[1]
        seed_flat = np.ascontiguousarray(seed.reshape(-1, order="F"))
        ys = self._host(np.asarray(yscale).reshape(-1), np.float64, (n_obs,), "yscale")
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.int64).reshape(-1)
        M = ds.N if ix is None else ix.size
        nsu = None if n_save_used is None else self._host(n_save_used, np.int32, (M,), "n_save_used")
        ls, grad = np.zeros(2), np.zeros(n_p)
        loss = np.empty(M) if want_loss else None
        n_saved = np.empty(M, dtype=np.int32) if want_loss else None
        ret = np.empty(M, dtype=np.int32) if want_loss else None
        stats = np.empty(M, dtype=STATS_DTYPE) if want_stats else None
        hp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._check(self._lib.crnn_loss_grad_indexed(
            self._h, C.byref(cm), C.byref(co), hp(seed_flat), n_p, ds._d, hp(ix), M, hp(nsu), hp(ys), int(loss_kind),
            hp(ls), hp(grad), hp(loss), hp(n_saved), hp(ret), hp(stats)))
        return dict(loss_sum=float(ls[0]), n_ok=int(ls[1]), loss_mean=float(ls[0] / max(ls[1], 1.0)), grad_sum=grad,
                    loss=loss, n_saved=n_saved, retcode=ret, stats=stats)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.crnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.crnn_launch_count(self._h))

    def profile_begin(self):
        self._check(self._lib.crnn_profile_begin(self._h))

    def profile_end(self):
        """-> (summed solver-kernel milliseconds, number of solver-kernel launches)"""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self._check(self._lib.crnn_profile_end(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def grad_each(self, N: int, n_p: int) -> np.ndarray:
        """[N, np] per-trajectory gradients d loss_i / d p of the last forward-mode loss_grad_batch call
        (rober_crnn_lm.jl:216-218 takes the Jacobian of the per-experiment losses)."""
        out = np.empty((N, n_p))
        self._check(self._lib.crnn_copy_grad_each(self._h, out.ctypes.data_as(C.c_void_p), N, n_p, 0, None))
        return out

    def lean_math(self, op: str, x, x2=None) -> np.ndarray:
        """The device copy of crnn_b200/csrc/lean_math.h evaluated elementwise (diagnostic; op in log, exp, pow, log10, exp10)."""
        code = {"log": 0, "exp": 1, "pow": 2, "log10": 3, "exp10": 4}[op]
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        x2 = None if x2 is None else np.ascontiguousarray(x2, dtype=np.float64).reshape(-1)
        y = np.empty_like(x)
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._check(self._lib.crnn_debug_lean_math(self._h, code, vp(x), vp(x2), vp(y), x.size))
        return y

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(f"crnn_b200 error {rc}: {self._lib.crnn_last_error(self._h).decode()}")

    # ------------------------------------------------------------------
    @staticmethod
    def _host(a, dtype, shape=None, name="array"):
        a = np.ascontiguousarray(a, dtype=dtype)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {shape}, got {a.shape}")
        return a

    @staticmethod
    def _dev(t, dtype, shape, name):
        import torch
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
            raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
        return t

    def solve_batch(self, model: CRNNModel, opts: SolveOpts, u0, n_save_used=None, want_stats=True, out=None):
        """predict_neuralode for N initial conditions -> dict(pred, n_saved, retcode, stats)."""
        n_obs, n_save, n = opts.n_obs(model.n_state), opts.n_save, model.n_state
        cm, k1 = model.to_c()
        if _is_torch(u0):
            import torch
            N = u0.shape[0]
            self._dev(u0, torch.float64, (N, n), "u0")
            dev = u0.device
            co, k2 = opts.to_c(n, True, torch.cuda.current_stream(dev).cuda_stream)
            pred = torch.empty((N, n_save, n_obs), dtype=torch.float64, device=dev) if out is None else out
            n_saved = torch.empty(N, dtype=torch.int32, device=dev)
            ret = torch.empty(N, dtype=torch.int32, device=dev)
            stats = torch.empty((N, STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev) if want_stats else None
            nsu = None if n_save_used is None else self._dev(n_save_used, torch.int32, (N,), "n_save_used")
            ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            self._check(self._lib.crnn_solve_batch(self._h, C.byref(cm), C.byref(co), ptr(u0), N, ptr(nsu),
                                                   ptr(pred), ptr(n_saved), ptr(ret), ptr(stats)))
            return dict(pred=pred, n_saved=n_saved, retcode=ret, stats=stats)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        if u0.ndim == 1:
            u0 = u0[None, :]
        N = u0.shape[0]
        self._host(u0, np.float64, (N, n), "u0")
        co, k2 = opts.to_c(n, False)
        pred = np.empty((N, n_save, n_obs)) if out is None else out
        n_saved = np.empty(N, dtype=np.int32); ret = np.empty(N, dtype=np.int32)
        stats = np.empty(N, dtype=STATS_DTYPE) if want_stats else None
        nsu = None if n_save_used is None else self._host(n_save_used, np.int32, (N,), "n_save_used")
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        self._check(self._lib.crnn_solve_batch(self._h, C.byref(cm), C.byref(co), ptr(u0), N, ptr(nsu),
                                               ptr(pred), ptr(n_saved), ptr(ret), ptr(stats)))
        return dict(pred=pred, n_saved=n_saved, retcode=ret, stats=stats)

    def loss_grad_batch(self, model: CRNNModel, opts: SolveOpts, seed, u0, data, yscale,
                        loss_kind=_abi.LOSS_MAE_SCALED, n_save_used=None, want_pred=False, want_stats=True):
        """Batched loss_neuralode and its gradient -> dict(loss [N], grad_sum [np], ...).

        seed = dW/dp [n_w, np] (Jacobian of p2vec; rows [vec(w_in); w_b; vec(w_out)])."""
        n_obs, n_save, n = opts.n_obs(model.n_state), opts.n_save, model.n_state
        cm, k1 = model.to_c()
        seed = np.asarray(seed, dtype=np.float64)
        if seed.ndim != 2 or seed.shape[0] != model.n_w:
            raise ValueError(f"seed must be [n_w={model.n_w}, np]")
        n_p = seed.shape[1]
        seed_flat = np.ascontiguousarray(seed.reshape(-1, order="F"))
        ys = self._host(np.asarray(yscale).reshape(-1), np.float64, (n_obs,), "yscale")
        hp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        if _is_torch(u0):
            import torch
            N = u0.shape[0]
            dev = u0.device
            self._dev(u0, torch.float64, (N, n), "u0")
            self._dev(data, torch.float64, (N, n_save, n_obs), "data")
            co, k2 = opts.to_c(n, True, torch.cuda.current_stream(dev).cuda_stream)
            loss = torch.empty(N, dtype=torch.float64, device=dev)
            grad = torch.empty(n_p, dtype=torch.float64, device=dev)
            pred = torch.empty((N, n_save, n_obs), dtype=torch.float64, device=dev) if want_pred else None
            n_saved = torch.empty(N, dtype=torch.int32, device=dev)
            ret = torch.empty(N, dtype=torch.int32, device=dev)
            stats = torch.empty((N, STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev) if want_stats else None
            nsu = None if n_save_used is None else self._dev(n_save_used, torch.int32, (N,), "n_save_used")
            ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            self._check(self._lib.crnn_loss_grad_batch(
                self._h, C.byref(cm), C.byref(co), hp(seed_flat), n_p, ptr(u0), N, ptr(nsu), ptr(data), hp(ys),
                int(loss_kind), ptr(loss), ptr(grad), ptr(pred), ptr(n_saved), ptr(ret), ptr(stats)))
            return dict(loss=loss, grad_sum=grad, pred=pred, n_saved=n_saved, retcode=ret, stats=stats)
        u0 = np.ascontiguousarray(u0, dtype=np.float64)
        if u0.ndim == 1:
            u0 = u0[None, :]
        N = u0.shape[0]
        self._host(u0, np.float64, (N, n), "u0")
        data = self._host(data, np.float64, (N, n_save, n_obs), "data")
        co, k2 = opts.to_c(n, False)
        loss = np.empty(N); grad = np.zeros(n_p)
        pred = np.empty((N, n_save, n_obs)) if want_pred else None
        n_saved = np.empty(N, dtype=np.int32); ret = np.empty(N, dtype=np.int32)
        stats = np.empty(N, dtype=STATS_DTYPE) if want_stats else None
        nsu = None if n_save_used is None else self._host(n_save_used, np.int32, (N,), "n_save_used")
        self._check(self._lib.crnn_loss_grad_batch(
            self._h, C.byref(cm), C.byref(co), hp(seed_flat), n_p, hp(u0), N, hp(nsu), hp(data), hp(ys),
            int(loss_kind), hp(loss), hp(grad), hp(pred), hp(n_saved), hp(ret), hp(stats)))
        return dict(loss=loss, grad_sum=grad, pred=pred, n_saved=n_saved, retcode=ret, stats=stats)


def _particles(self, model: CRNNModel, opts: SolveOpts, weights, seeds, u0, data, yscale, loss_kind=_abi.LOSS_MSE,
               tab_T=None, tab_P=None, n_save_used=None, want_stats=False):
    """`crnn_loss_grad_particles`: P parameter sets ("particles") x E experiments in one launch
    (Cathode_NCM333_UQ/src_333/network.jl:222-260).  weights [P, n_w] (flat [vec(w_in); w_b; vec(w_out); (w_obs)] per
    particle), seeds [P, n_w, np], u0 [E, n_state], data [E, n_save, n_obs], tab_T / tab_P [E, n_tab] or None.
    -> dict(loss [P, E], grad [P, np], n_saved, retcode [P, E], stats)."""
    n = model.n_state
    cm, k1 = model.to_c()
    co, k2 = opts.to_c(n, False)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    P, nw = weights.shape
    if nw != model.n_w:
        raise ValueError(f"weights must be [P, n_w={model.n_w}]")
    seeds = np.asarray(seeds, dtype=np.float64)
    if seeds.ndim != 3 or seeds.shape[0] != P or seeds.shape[1] != nw:
        raise ValueError("seeds must be [P, n_w, np]")
    n_p = seeds.shape[2]
    seeds_f = np.ascontiguousarray(np.transpose(seeds, (0, 2, 1)))       # per particle [np][n_w] = column-major [n_w, np]
    u0 = self._host(u0, np.float64, None, "u0")
    E = u0.shape[0]
    data = self._host(data, np.float64, (E, opts.n_save, opts.n_obs(n)), "data")
    ys = self._host(np.asarray(yscale).reshape(-1), np.float64, (opts.n_obs(n),), "yscale")
    tT = None if tab_T is None else self._host(tab_T, np.float64, (E, model.tab_t.size), "tab_T")
    tP = None if tab_P is None else self._host(tab_P, np.float64, (E, model.tab_t.size), "tab_P")
    nsu = None if n_save_used is None else self._host(n_save_used, np.int32, (E,), "n_save_used")
    loss = np.empty((P, E)); grad = np.empty((P, n_p))
    n_saved = np.empty((P, E), dtype=np.int32); ret = np.empty((P, E), dtype=np.int32)
    stats = np.empty((P, E), dtype=STATS_DTYPE) if want_stats else None
    hp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    self._check(self._lib.crnn_loss_grad_particles(
        self._h, C.byref(cm), C.byref(co), hp(weights), hp(seeds_f), n_p, P, hp(u0), E, hp(nsu), hp(data), hp(tT), hp(tP),
        hp(ys), int(loss_kind), hp(loss), hp(grad), hp(n_saved), hp(ret), hp(stats)))
    return dict(loss=loss, grad=grad, n_saved=n_saved, retcode=ret, stats=stats)


Engine.loss_grad_particles = _particles


def _train_steps(self, model: CRNNModel, opts: SolveOpts, ds: Dataset, order, yscale, p, opt_state=None,
                 loss_kind=_abi.LOSS_MAE_SCALED, p2vec_kind=2, optimiser="adam", batch=1, eta=1e-3, beta=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, expdecay=None, grad_max=None, p2vec_b0=-10.0, n_save_used=None):
    """`crnn_train_steps`: the scripts' epoch loop (case2/case2.jl:192-198) entirely on the device - p2vec, solve +
    forward sensitivities, gradient reduction and the Flux optimiser chain enqueued back to back, nothing returning to the
    host between optimiser steps.  `order` [n_steps * batch] dataset rows (the host's randperm), `expdecay` =
    (eta, decay, step, clip) or None, `opt_state` [2 np + 4] from a previous call (None: fresh), `n_save_used` [n_steps * batch]
    the per-visit random truncation `sample = rand(batchsize:datasize)` of rober_crnn.jl:218 (None: all save points).
    -> dict(p, opt_state, step_loss [n_steps], step_gnorm [n_steps])."""
    cm, k1 = model.to_c()
    co, k2 = opts.to_c(model.n_state, False)
    order = np.ascontiguousarray(order, dtype=np.int64).reshape(-1)
    if order.size % batch:
        raise ValueError("order must hold n_steps * batch dataset rows")
    n_steps = order.size // batch
    p = np.array(p, dtype=np.float64).reshape(-1)
    n_p = p.size
    ed = expdecay or (0.0, 1.0, 0, 0.0)
    if opt_state is None:
        opt_state = np.concatenate([np.zeros(2 * n_p), [beta[0], beta[1], ed[0], 0.0]])
    opt_state = np.array(opt_state, dtype=np.float64).reshape(-1)
    if opt_state.size != 2 * n_p + 4:
        raise ValueError("opt_state must be [2*np + 4]")
    nsu = None
    if n_save_used is not None:
        nsu = np.ascontiguousarray(n_save_used, dtype=np.int32).reshape(-1)
        if nsu.size != order.size:
            raise ValueError("n_save_used must hold one entry per visited experiment (n_steps * batch)")
    t = _abi.CTrainOpts(int(p2vec_kind), {"adam": 0, "nadam": 1}[optimiser], int(batch), 0, eta, beta[0], beta[1], eps, weight_decay,
                        ed[0], ed[1], ed[3], int(ed[2]), 0.0 if grad_max is None else float(grad_max), float(p2vec_b0),
                        None if nsu is None else nsu.ctypes.data)
    ys = self._host(np.asarray(yscale).reshape(-1), np.float64, (opts.n_obs(model.n_state),), "yscale")
    sl, sg = np.empty(n_steps), np.empty(n_steps)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    self._check(self._lib.crnn_train_steps(self._h, C.byref(cm), C.byref(co), C.byref(t), ds._d, hp(order), n_steps, hp(ys),
                                           int(loss_kind), hp(p), hp(opt_state), hp(sl), hp(sg)))
    return dict(p=p, opt_state=opt_state, step_loss=sl, step_gnorm=sg)


Engine.train_steps = _train_steps


def stats_from_torch(stats_u8) -> np.ndarray:
    """uint8 [N, 32] CUDA tensor of crnn_stats -> numpy structured array."""
    return stats_u8.cpu().numpy().view(STATS_DTYPE).reshape(-1)
