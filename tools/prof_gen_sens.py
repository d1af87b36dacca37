"""One launch of k_gen_sens on the HyChem F2 model (np = 211, AutoTsit5(Rosenbrock23)) for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
YS = np.array([0.05, 0.02, 0.01, 0.02, 0.01, 0.02, 0.01, 0.01, 0.9])
eng = Engine(0)
N = int(os.environ.get("PROF_N", "2048"))
alg = {"auto": _abi.ALG_AUTO_TSIT5_ROS23, "ros23": _abi.ALG_ROSENBROCK23, "tsit5": _abi.ALG_TSIT5}[os.environ.get("PROF_ALG", "auto")]
m, seed = cases.hychem_model(cases.hychem_p(0, stiff=4.0), YS)
u0 = torch.from_numpy(cases.hychem_u0(N)).cuda()
data = eng.solve_batch(cases.hychem_model(cases.hychem_p(1, stiff=4.0), YS)[0], cases.hychem_opts(alg=_abi.ALG_ROSENBROCK23), u0, want_stats=False)["pred"]
r = eng.loss_grad_batch(m, cases.hychem_opts(alg=alg, maxiters=100000), seed, u0, data, YS, want_stats=False)
torch.cuda.synchronize()
print("ok", float(r["loss"].nanmean()))
