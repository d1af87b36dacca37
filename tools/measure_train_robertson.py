"""rober_crnn.jl's batch-1 epoch loop: optimiser steps per second on the device (crnn_train_steps, p2vec_kind 4) and driven from the host"""
import sys, json, time, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from crnn_b200 import _abi, cases, optim
from crnn_b200.engine import Engine
from crnn_b200.frontend import CRNNProblem
from problems import make_problem
golden = json.load(open(os.path.join(R, 'tests', 'golden', 'checkpoints.json')))
eng = Engine(0)
pb = make_problem("robertson", golden, 30)
prob = CRNNProblem("robertson", pb["u0"], pb["data"], pb["yscale"], out_scale=pb["model"].out_scale, engine=eng)
p = np.array(golden["robertson"]["p"])
model, _ = prob.case.model(p, prob.out_scale)
g = np.random.default_rng(0)
order = np.concatenate([g.permutation(30) for _ in range(40)])
sample = g.integers(32, 41, size=order.size)
kw = dict(p2vec_kind=4, optimiser="adam", eta=0.005, weight_decay=1e-6, grad_max=10.0)
eng.train_steps(model, prob.opts, prob.dataset, order[:100], prob.yscale, p, None, prob.case.loss_kind, n_save_used=sample[:100], **kw)
t0 = time.perf_counter(); r = eng.train_steps(model, prob.opts, prob.dataset, order, prob.yscale, p, None, prob.case.loss_kind, n_save_used=sample, **kw); t1 = time.perf_counter()
print("device loop steps/s", order.size / (t1 - t0))
opt = optim.ADAMW(0.005, (0.9, 0.999), 1e-6); q = p.copy()
t0 = time.perf_counter()
for k, i in enumerate(order[:300]):
    l, gr = prob.loss_grad(q, np.array([i]), sample=sample[k:k + 1]); gr, _ = optim.clip_by_norm(gr, 10.0); opt.update(q, gr)
t1 = time.perf_counter()
print("host loop steps/s", 300 / (t1 - t0))
