"""diagnostic: on-device training loop against the host loop, step by step (prints the first diverging step)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from crnn_b200 import optim
from crnn_b200.engine import Engine
from test_training_gpu import _case2_problem

eng = Engine(0)
prob, p0, g = _case2_problem(eng)
batch, n_steps = 1, 60
kw = dict(optimiser="adam", eta=0.005, beta=(0.9, 0.999), weight_decay=1e-6, expdecay=(5e-3, 0.5, 25, 1e-4))
order = np.concatenate([g.permutation(20) for _ in range(3)])[:60]
model, _ = prob.case.model(p0)
opt = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 25, 1e-4), optim.ADAMW(0.005, (0.9, 0.999), 1e-6))
p = p0.copy(); pd = p0.copy(); st = None
for s in range(n_steps):
    loss, grad = prob.loss_grad(p, order[s:s + 1])
    lossd_h, gradd_h = prob.loss_grad(pd, order[s:s + 1])      # host gradient at the device's p
    r = eng.train_steps(model, prob.opts, prob.dataset, order[s:s + 1], prob.yscale, pd, st, prob.case.loss_kind, batch=1, **kw)
    opt.update(p, grad)
    print(s, loss, r["step_loss"][0], lossd_h, np.linalg.norm(grad), r["step_gnorm"][0], np.linalg.norm(gradd_h),
          np.abs(r["p"] - p).max(), flush=True)
    pd, st = r["p"], r["opt_state"]
