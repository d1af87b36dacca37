#!/usr/bin/env python
"""Times loss+gradient (crnn_loss_grad_batch, device buffers) on the training shapes:
case2 Tsit5 np=25 (65 536), robertson Rosenbrock23 np=43 (262 144), case3 Tsit5 np=153 (131 072).
usage: python tools/measure_train.py [out.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine, stats_from_torch

TRUE = {"case2": cases.true_model_case2, "robertson": cases.true_model_robertson, "case3": cases.true_model_case3}


def main():
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
    eng = Engine(0)
    out = {}
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["case2", "robertson", "case3"]
    for name, N, mode in (("case2", 65536, "forward"), ("case2", 65536, "adjoint"), ("case2", 65536, "discrete"),
                          ("robertson", 65536, "forward"), ("case3", 65536, "forward"), ("case3", 65536, "adjoint"),
                          ("case3", 65536, "discrete")):
        if name not in which:
            continue
        c = cases.CASES[name]
        u0 = synth.make_u0(name, N)
        obs = np.arange(c.ns)
        truth = eng.solve_batch(TRUE[name](), c.opts(obs_idx=obs, pred_clamp=(-np.inf, np.inf)), u0, want_stats=False)["pred"]
        g = np.random.default_rng(1)
        noise = 0.05 if name != "robertson" else 1e-4
        data = np.abs(truth * (1.0 + noise * g.standard_normal(truth.shape, dtype=np.float32))) + (1e-6 if name == "case3" else 0.0)
        ys = synth.yscale_from(data[:4096], c.lb if name != "robertson" else 0.0)
        out_scale = ys / c.tspan[1] if name in ("robertson", "case3") else None
        if name in ("case2", "robertson"):
            p = np.array(golden[name]["p"])
        else:
            g = np.random.default_rng(0); p = (g.random(c.n_p) - 0.5) * 2 * np.sqrt(6 / (c.ns + c.nr)); p[-1] = 0.1
        model, seed = c.model(p, out_scale)
        opts = c.opts(obs_idx=obs, sens_mode={"adjoint": _abi.SENS_INTERP_ADJOINT, "discrete": _abi.SENS_DISCRETE_ADJOINT, "forward": _abi.SENS_FORWARD}[mode])
        u0_d = torch.from_numpy(u0).cuda(); data_d = torch.from_numpy(data).cuda()
        for _ in range(2):
            r = eng.loss_grad_batch(model, opts, seed, u0_d, data_d, ys, c.loss_kind)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for _ in range(K):
            r = eng.loss_grad_batch(model, opts, seed, u0_d, data_d, ys, c.loss_kind)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        st = stats_from_torch(r["stats"])
        out[name + "_" + mode] = {"N": N, "np": int(seed.shape[1]), "ms": ms, "traj_per_s": N / ms * 1e3,
                     "rhs_per_s": float(st["n_rhs"].sum()) / ms * 1e3,
                     "steps_mean": float((st["n_accept"] + st["n_reject"]).mean()), "back_steps_mean": float(st["n_jac"].mean()),
                     "success_frac": float((r["retcode"] == 1).float().mean().item()),
                     "loss_mean": float(torch.nanmean(r["loss"]).item())}
        print(name, mode, out[name + "_" + mode])
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
