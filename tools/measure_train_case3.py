"""case3.jl's batch-1 epoch loop: optimiser steps per second on the device (crnn_train_steps, p2vec_kind 3) and driven from the host"""
import sys, json, time; import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from crnn_b200.frontend import CRNNProblem
from problems import make_problem, trained_p
golden = json.load(open(os.path.join(R, 'tests', 'golden', 'checkpoints.json')))
eng = Engine(0)
pb = make_problem("case3", golden, 20)
prob = CRNNProblem("case3", pb["u0"], np.abs(pb["data"]) + 1e-6, pb["yscale"], out_scale=pb["model"].out_scale, engine=eng)
p = trained_p("case3", golden)
model, _ = prob.case.model(p, prob.out_scale)
order = np.concatenate([np.random.default_rng(i).permutation(20) for i in range(50)])
kw = dict(p2vec_kind=3, optimiser="nadam", eta=0.001)
eng.train_steps(model, prob.opts, prob.dataset, order[:100], prob.yscale, p, None, prob.case.loss_kind, **kw)
t0 = time.perf_counter(); r = eng.train_steps(model, prob.opts, prob.dataset, order, prob.yscale, p, None, prob.case.loss_kind, **kw); t1 = time.perf_counter()
print("device loop steps/s", order.size / (t1 - t0))
from crnn_b200 import optim
opt = optim.Optimiser(optim.NADAM(0.001)); q = p.copy()
t0 = time.perf_counter()
for i in order[:300]:
    l, g = prob.loss_grad(q, np.array([i])); opt.update(q, g)
t1 = time.perf_counter()
print("host loop steps/s", 300 / (t1 - t0))
