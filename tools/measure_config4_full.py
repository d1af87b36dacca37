#!/usr/bin/env python
"""BASELINE config 4 at its FULL size on one B200: case3 (np = 153), 1 048 576 ICs resident in HBM (7.5 GB of targets),
loss + gradient by the interpolating adjoint, the discrete adjoint and the 153-column forward mode."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine
from problems import trained_p

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
c = cases.CASES["case3"]
g = torch.Generator(device="cuda"); g.manual_seed(1234)
u0 = torch.pow(10.0, -3.0 * torch.rand((N, 9), dtype=torch.float64, device="cuda", generator=g))   # 10^(-3 U(0,1)), case3.jl:106
obs = np.arange(c.ns)
truth = eng.solve_batch(cases.true_model_case3(), c.opts(obs_idx=obs, pred_clamp=(-np.inf, np.inf)), u0, want_stats=False)["pred"]
data = (truth * (1.0 + 0.05 * torch.randn(truth.shape, dtype=torch.float64, device="cuda", generator=g))).abs_() + 1e-6
del truth
ys = (data[:4096].amax(dim=(0, 1)) - data[:4096].amin(dim=(0, 1)) + c.lb).cpu().numpy()
model, seed = c.model(trained_p("case3", golden), out_scale=ys / c.tspan[1])
out = {"N": N, "targets_GB": data.numel() * 8 / 1e9}
for mode, sm in (("interp_adjoint", _abi.SENS_INTERP_ADJOINT), ("discrete_adjoint", _abi.SENS_DISCRETE_ADJOINT), ("forward_153_columns", _abi.SENS_FORWARD)):
    o = c.opts(obs_idx=obs, sens_mode=sm)
    r = eng.loss_grad_batch(model, o, seed, u0, data, ys, c.loss_kind, want_stats=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = eng.loss_grad_batch(model, o, seed, u0, data, ys, c.loss_kind, want_stats=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[mode] = {"ms": ms, "traj_per_s": N / ms * 1e3, "success_frac": float((r["retcode"] == 1).float().mean().item()),
                 "loss_mean": float(torch.nanmean(r["loss"]).item())}
    print(mode, out[mode], flush=True)
print(json.dumps(out))
