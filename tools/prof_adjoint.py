"""One launch of k_tsit5_adjoint (discrete or interp) on case2 / case3 for ncu.  usage: prof_adjoint.py case2|case3 discrete|interp [N]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine
from problems import make_problem
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
name, mode = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
eng = Engine(0)
pb = make_problem(name, golden, 256)
c = pb["case"]
u0 = synth.make_u0(name, N)
truth = eng.solve_batch(pb["true_model"], c.opts(obs_idx=np.arange(c.ns), pred_clamp=(-np.inf, np.inf)), u0, want_stats=False)["pred"]
data = np.abs(synth.noisy_targets(truth, 0.05)) + 1e-6
o = c.opts(obs_idx=np.arange(c.ns), sens_mode=_abi.SENS_DISCRETE_ADJOINT if mode == "discrete" else _abi.SENS_INTERP_ADJOINT)
ud = torch.from_numpy(u0).cuda(); dd = torch.from_numpy(data).cuda()
for _ in range(2):
    r = eng.loss_grad_batch(pb["model"], o, pb["seed"], ud, dd, pb["yscale"], c.loss_kind, want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
