#!/usr/bin/env python
"""Throughput of the generic forward-sensitivity kernel (kernel_gen_sens.cuh) on the HyChem model (np = 211) and on case2
through the composite algorithm; JSON to stdout."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crnn_b200 import _abi, cases, synth
from crnn_b200.engine import Engine

YS = np.array([0.05, 0.02, 0.01, 0.02, 0.01, 0.02, 0.01, 0.01, 0.9])
eng = Engine(0)
dev = torch.device("cuda", 0)
out = {}

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

N = 16384
for alg, nm in ((_abi.ALG_TSIT5, "tsit5"), (_abi.ALG_ROSENBROCK23, "ros23"), (_abi.ALG_AUTO_TSIT5_ROS23, "auto")):
    kw = dict(lnA_shift=-2.0) if nm == "tsit5" else dict(stiff=4.0)
    m, seed = cases.hychem_model(cases.hychem_p(0, **kw), YS)
    u0 = torch.from_numpy(cases.hychem_u0(N)).to(dev)
    data = eng.solve_batch(cases.hychem_model(cases.hychem_p(1, **kw), YS)[0], cases.hychem_opts(alg=_abi.ALG_ROSENBROCK23), u0, want_stats=False)["pred"]
    o = cases.hychem_opts(alg=alg, maxiters=100000)
    ms, r = timed(lambda: eng.loss_grad_batch(m, o, seed, u0, data, YS, want_stats=False))
    out[f"hychem_f2_np211_{nm}"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3, "loss": float(r["loss"].nanmean())}
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "checkpoints.json")))
c = cases.CASES["case2"]
model, seed = c.model(np.array(golden["case2"]["p"]))
N = 65536
u0 = torch.from_numpy(synth.make_u0("case2", N)).to(dev)
data = eng.solve_batch(cases.true_model_case2(), c.opts(obs_idx=np.arange(6), pred_clamp=(-np.inf, np.inf)), u0, want_stats=False)["pred"]
ys = np.ones(6)
for alg, nm in ((_abi.ALG_TSIT5, "tsit5_specialised"), (_abi.ALG_AUTO_TSIT5_ROS23, "auto_generic")):
    o = c.opts(obs_idx=np.arange(6), alg=alg)
    ms, r = timed(lambda: eng.loss_grad_batch(model, o, seed, u0, data, ys, want_stats=False))
    out[f"case2_np25_{nm}"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3}
print(json.dumps(out, indent=1))
