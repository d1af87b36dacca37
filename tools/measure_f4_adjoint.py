#!/usr/bin/env python
"""Throughput of the F4 gradient path (k_tsit5_adjoint<..., MLP>) on the reference's yeast checkpoint: loss + gradient of all 294
parameters (164 CRNN + 130 MLP), 300 saves on [0, 5] as yeast_glycolysis.jl:24-27, device-resident buffers; the predict path of the
same model beside it, and the oracle's rate on a small sample.  JSON to stdout."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from oracle import oracle

eng = Engine(0)
dev = torch.device("cuda", 0)
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "checkpoints.json")))
p = np.array(golden["yeast"]["p"])
m, seed = cases.yeast_model(p), cases.yeast_seed(p)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

N = 16384
g = np.random.default_rng(0)
u0h = cases.YEAST_IC_LB + g.random((N, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
u0 = torch.from_numpy(u0h).to(dev)
ov = cases.yeast_opts(alg=_abi.ALG_TSIT5)
# targets: the checkpoint's own predictions from perturbed initial conditions (the gradient is then non-trivial)
data = eng.solve_batch(m, ov, torch.from_numpy(u0h * (1.0 + 0.02 * g.normal(size=u0h.shape))).to(dev), want_stats=False)["pred"]
ys = np.ones(7)
out = {}
ms, _ = timed(lambda: eng.solve_batch(m, ov, u0, want_stats=False))
out["yeast_predict_tsit5"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3}
ms, _ = timed(lambda: eng.solve_batch(m, cases.yeast_opts(), u0, want_stats=False))
out["yeast_predict_auto_trbdf2"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3}
for mode, nm in ((_abi.SENS_DISCRETE_ADJOINT, "discrete"), (_abi.SENS_INTERP_ADJOINT, "interpolating")):
    o = cases.yeast_opts(alg=_abi.ALG_TSIT5, sens_mode=mode)
    ms, r = timed(lambda: eng.loss_grad_batch(m, o, seed, u0, data, ys, want_stats=False))
    out[f"yeast_grad_np294_{nm}_adjoint"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3, "loss": float(np.nanmean(r["loss"].cpu().numpy() if hasattr(r["loss"], "cpu") else r["loss"]))}
    if "--no-oracle" in sys.argv:
        continue
    n = 64
    t0 = time.perf_counter()
    oracle.loss_grad_batch(m, o, seed, u0h[:n], (data[:n].cpu().numpy() if hasattr(data, "cpu") else data[:n]), ys, n_threads=os.cpu_count())
    out[f"yeast_grad_np294_{nm}_adjoint"]["oracle_traj_per_s"] = n / (time.perf_counter() - t0)
    out[f"yeast_grad_np294_{nm}_adjoint"]["oracle_threads"] = os.cpu_count()
print(json.dumps(out, indent=1))
