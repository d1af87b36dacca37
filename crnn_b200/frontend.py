"""Python mirror of the reference scripts' front end, batched.

The Julia scripts keep `p2vec`, `predict_neuralode`, `loss_neuralode` and the ADAM epoch loop
(case2/case2.jl:91-99,124-128,132-137,192-207); only `solve` and `ForwardDiff.gradient` are
replaced by the C-ABI.  Julia is not available in this image, so this module restates that
front end in Python with the same names and argument meaning; `julia/CRNNB200.jl` is the shim a
maintainer would drop into the scripts.

Semantics that change by batching: the scripts take one optimiser step per experiment
(batch size 1, `for i_exp in randperm(n_exp_train)`); `train` here takes one step per
mini-batch with the mean gradient, `batch=1` reproduces the per-experiment loop.
"""
from __future__ import annotations

import numpy as np

from . import _abi, optim as _optim
from .cases import CASES, Case
from .engine import Engine


class CRNNProblem:
    """One script's globals: u0_list, ode_data_list, tsteps, yscale (+ dydt_scale), i_obs."""

    def __init__(self, case: str | Case, u0_list, ode_data_list, yscale, i_obs=None, out_scale=None,
                 engine: Engine | None = None, **opt_overrides):
        self.case = CASES[case] if isinstance(case, str) else case
        self.u0_list = np.ascontiguousarray(u0_list, dtype=np.float64)            # [n_exp, n_state]
        self.i_obs = np.arange(self.case.ns) if i_obs is None else np.asarray(i_obs)
        self.ode_data_list = np.ascontiguousarray(ode_data_list, dtype=np.float64)  # [n_exp, n_save, n_obs]
        self.yscale = np.asarray(yscale, dtype=np.float64).reshape(-1)
        self.out_scale = out_scale
        self.engine = engine or Engine()
        self.opts = self.case.opts(obs_idx=self.i_obs, **opt_overrides)
        # the training set lives on the device(s): uploaded once, every optimiser step then moves only weights in and
        # [loss, gradient] out (crnn_dataset_create / crnn_loss_grad_indexed)
        self.dataset = self.engine.dataset(self.u0_list, self.ode_data_list)

    # -- the scripts' functions ------------------------------------------------------------
    def p2vec(self, p):
        w_in, w_b, w_out, _ = self.case.p2vec(p)
        return w_in, w_b, w_out

    def predict_neuralode(self, u0, p, sample=None):
        """`predict_neuralode(u0, p)`: clamped saved states [n_obs, n_saved] (Julia orientation)
        for one IC, or [N, n_save, n_obs] for a batch of ICs."""
        model, _ = self.case.model(p, self.out_scale)
        u0 = np.asarray(u0, dtype=np.float64)
        single = u0.ndim == 1
        nsu = None if sample is None else np.full(1 if single else u0.shape[0], sample, dtype=np.int32)
        r = self.engine.solve_batch(model, self.opts, u0, n_save_used=nsu)
        if single:
            if r["retcode"][0] != _abi.RET_SUCCESS:
                print("ode solver failed")   # robertson/rober_crnn.jl:130-134
            return r["pred"][0, :r["n_saved"][0]].T
        return r["pred"]

    def loss_neuralode(self, p, i_exp, sample=None):
        """`loss_neuralode(p, i_exp)` (0-based i_exp); an index array gives the per-experiment losses."""
        return self.loss_grad(p, i_exp, sample=sample)[0]

    def loss_grad(self, p, idx, sample=None):
        """(loss, grad) for experiments `idx`: `ForwardDiff.gradient(x -> loss_neuralode(x, i), p)`
        batched; the loss / gradient are means over the batch."""
        idx = np.atleast_1d(np.asarray(idx))
        model, seed = self.case.model(p, self.out_scale)
        nsu = None if sample is None else np.broadcast_to(np.asarray(sample, dtype=np.int32), idx.shape).copy()
        r = self.engine.loss_grad_indexed(model, self.opts, seed, self.dataset, self.yscale, self.case.loss_kind,
                                          idx=idx, n_save_used=nsu)
        n = max(r["n_ok"], 1)
        return r["loss_sum"] / n, r["grad_sum"] / n

    # -- the epoch loop ---------------------------------------------------------------------
    def train(self, p, opt: _optim.Optimiser, n_epoch, n_exp_train, batch=None, grad_max=None, rng=None,
              sample_range=None, callback=None):
        """The scripts' training loop (case2/case2.jl:192-207, robertson/rober_crnn.jl:215-234)."""
        rng = rng or np.random.default_rng(0)
        p = np.array(p, dtype=np.float64)
        batch = batch or n_exp_train
        history = []
        for epoch in range(n_epoch):
            perm = rng.permutation(n_exp_train)
            gnorms = []
            for lo in range(0, n_exp_train, batch):
                idx = perm[lo:lo + batch]
                sample = None if sample_range is None else rng.integers(sample_range[0], sample_range[1] + 1, size=idx.size)
                _, grad = self.loss_grad(p, idx, sample=sample)
                if grad_max is not None:
                    grad, gn = _optim.clip_by_norm(grad, grad_max)
                else:
                    gn = float(np.linalg.norm(grad))
                gnorms.append(gn)
                opt.update(p, grad)
            n_exp = self.u0_list.shape[0]
            model, seed = self.case.model(p, self.out_scale)
            losses = self.engine.loss_grad_indexed(model, self.opts, seed, self.dataset, self.yscale,
                                                   self.case.loss_kind, want_loss=True)["loss"]
            loss_train = float(np.mean(losses[:n_exp_train]))
            loss_val = float(np.mean(losses[n_exp_train:])) if n_exp > n_exp_train else float("nan")
            history.append((loss_train, loss_val, float(np.mean(gnorms))))
            if callback is not None:
                callback(p, loss_train, loss_val)
        return p, history


    def train_on_device(self, p, n_epoch, n_exp_train, rng=None, batch=1, sample_range=None, **optimiser):
        """The same epoch loop with every optimiser step ON the device (`crnn_train_steps`): one C call per epoch, the
        visiting order `randperm(n_exp_train)` (case2/case2.jl:194) drawn here.  `optimiser`: Engine.train_steps keywords
        (optimiser, eta, beta, weight_decay, expdecay, grad_max).  -> (p, history) like `train`."""
        rng = rng or np.random.default_rng(0)
        p = np.array(p, dtype=np.float64)
        model, _ = self.case.model(p, self.out_scale)
        state, history = None, []
        n_steps = n_exp_train // batch
        optimiser.setdefault("p2vec_kind", {"case1": 1, "case2": 2, "case3": 3, "robertson": 4}.get(self.case.name, 0))   # the device p2vec kernels built
        for epoch in range(n_epoch):
            order = rng.permutation(n_exp_train)[:n_steps * batch]
            sample = None if sample_range is None else rng.integers(sample_range[0], sample_range[1] + 1, size=order.size)
            r = self.engine.train_steps(model, self.opts, self.dataset, order, self.yscale, p, state, self.case.loss_kind,
                                        batch=batch, n_save_used=sample, **optimiser)
            p, state = r["p"], r["opt_state"]
            model, seed = self.case.model(p, self.out_scale)
            losses = self.engine.loss_grad_indexed(model, self.opts, seed, self.dataset, self.yscale,
                                                   self.case.loss_kind, want_loss=True)["loss"]
            n_exp = self.u0_list.shape[0]
            history.append((float(np.mean(losses[:n_exp_train])),
                            float(np.mean(losses[n_exp_train:])) if n_exp > n_exp_train else float("nan"),
                            float(np.mean(r["step_gnorm"]))))
        return p, history


def save_checkpoint(path, p, history, iter_=None):
    """`@save "./checkpoint/mymodel.bson" p opt l_loss_train l_loss_val iter` (case2/case2.jl:178) for a `train` history
    (list of (loss_train, loss_val, grad_norm)); see crnn_b200/checkpoint.py for what is and is not written."""
    from . import checkpoint
    checkpoint.save(path, p, len(history) if iter_ is None else iter_,
                    l_loss_train=[h[0] for h in history], l_loss_val=[h[1] for h in history])


def load_checkpoint(path):
    """`@load` (case2/case2.jl:183-186): -> (p, iter, l_loss_train, l_loss_val); also reads the reference's own files."""
    from . import checkpoint
    c = checkpoint.load(path)
    tr = c.get("l_loss_train", c.get("list_loss_train", []))
    va = c.get("l_loss_val", c.get("list_loss_val", []))
    return c["p"], c["iter"], tr, va
