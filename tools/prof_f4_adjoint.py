"""One launch of k_tsit5_adjoint<..., MLP> (discrete or interp) on the yeast checkpoint for ncu.  usage: prof_f4_adjoint.py discrete|interp [N]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from crnn_b200 import cases, _abi
from crnn_b200.engine import Engine
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
mode = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
eng = Engine(0)
p = np.array(golden["yeast"]["p"])
m, seed = cases.yeast_model(p), cases.yeast_seed(p)
g = np.random.default_rng(0)
u0h = cases.YEAST_IC_LB + g.random((N, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
ud = torch.from_numpy(u0h).cuda()
dd = eng.solve_batch(m, cases.yeast_opts(alg=_abi.ALG_TSIT5), torch.from_numpy(u0h * (1.0 + 0.02 * g.normal(size=u0h.shape))).cuda(), want_stats=False)["pred"]
o = cases.yeast_opts(alg=_abi.ALG_TSIT5, sens_mode=_abi.SENS_DISCRETE_ADJOINT if mode == "discrete" else _abi.SENS_INTERP_ADJOINT)
for _ in range(2):
    r = eng.loss_grad_batch(m, o, seed, ud, dd, np.ones(7), want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
