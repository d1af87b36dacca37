#!/usr/bin/env python
"""case2/case2.jl's training loop on the engine (see tests/test_training_gpu.py): prints the loss history.
usage: python tools/train_demo.py [n_epoch] [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crnn_b200 import cases, optim, synth
from crnn_b200.engine import Engine
from crnn_b200.frontend import CRNNProblem

n_epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 300
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
c = cases.CASES["case2"]
n_tr, n_val = 20, 10
u0 = synth.make_u0("case2", n_tr + n_val, seed=1234)
truth = eng.solve_batch(cases.true_model_case2(), c.opts(obs_idx=np.arange(c.ns), pred_clamp=(-np.inf, np.inf)), u0)
data = synth.noisy_targets(truth["pred"], 0.05)
prob = CRNNProblem("case2", u0, data, synth.yscale_from(data, c.lb), engine=eng)
g = np.random.default_rng(1234)
p = g.standard_normal(c.n_p) * 0.1
p[:c.nr] += 0.8; p[c.nr * (c.ns + 1):c.nr * (c.ns + 2)] += 0.8; p[-1] = 0.1
opt = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 500 * n_tr, 1e-4), *optim.ADAMW(0.005, (0.9, 0.999), 1e-6).chain)
t = time.perf_counter()
marks = sorted(set([0, 1, 2, 5, 10, 20, 50, 100, 200, 300, 500, 1000, 2000, n_epoch - 1]))
def cb(pp, ltr, lval, _s={"e": 0}):
    if _s["e"] in marks:
        print(f"epoch {_s['e']:5d}  train {ltr:.4f}  val {lval:.4f}  ({time.perf_counter() - t:.1f} s)", flush=True)
    _s["e"] += 1
p_end, hist = prob.train(p, opt, n_epoch=n_epoch, n_exp_train=n_tr, batch=batch, rng=g, callback=cb)
w_in, w_b, w_out = prob.p2vec(p_end)
print("ln A  ", np.round(w_b, 2), " (generating mechanism: 18.60 19.13 7.93)")
print("Ea    ", np.round(w_in[c.ns], 2), " (generating mechanism: 14.54 14.42 6.47)")
print("w_out'"); print(np.round(w_out.T, 2))
print(f"{n_epoch} epochs x {n_tr // batch} optimiser steps in {time.perf_counter() - t:.1f} s")
