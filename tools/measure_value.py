#!/usr/bin/env python
"""Times the predict path (crnn_solve_batch, device buffers) on the BASELINE configs' shapes:
case2 Tsit5 65 536 ICs, robertson Rosenbrock23 262 144 ICs, case3 Tsit5 1 048 576/8 ICs.
usage: python tools/measure_value.py [out.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from crnn_b200 import cases, synth
from crnn_b200.engine import Engine, stats_from_torch


def main():
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
    eng = Engine(0)
    out = {}
    for name, N in (("case2", 65536), ("robertson", 262144), ("case3", 131072)):
        c = cases.CASES[name]
        u0 = synth.make_u0(name, N)
        out_scale = None
        if name == "robertson":
            out_scale = np.array([1.35, 7.5e-6, 1.35]) / 1e5      # dydt_scale
            p = np.array(golden["robertson"]["p"])
        elif name == "case2":
            p = np.array(golden["case2"]["p"])
        else:
            g = np.random.default_rng(0); p = (g.random(c.n_p) - 0.5) * 2 * np.sqrt(6 / (c.ns + c.nr)); p[-1] = 0.1
            out_scale = np.full(c.ns, 0.1)
        model, _ = c.model(p, out_scale)
        opts = c.opts(obs_idx=np.arange(c.ns))
        u0_d = torch.from_numpy(u0).cuda()
        pred = torch.empty((N, opts.n_save, c.ns), dtype=torch.float64, device="cuda")
        for _ in range(3):
            r = eng.solve_batch(model, opts, u0_d, out=pred)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for _ in range(K):
            r = eng.solve_batch(model, opts, u0_d, out=pred)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        st = stats_from_torch(r["stats"])
        ok = float((r["retcode"] == 1).float().mean().item())
        bytes_traj = 8 * (model.n_state + c.ns * opts.n_save)
        out[name] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3, "rhs_per_s": float(st["n_rhs"].sum()) / ms * 1e3,
                     "steps_mean": float((st["n_accept"] + st["n_reject"]).mean()), "success_frac": ok,
                     "hbm_GBps_algorithmic": N * bytes_traj / ms / 1e6}
        print(name, out[name])
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
