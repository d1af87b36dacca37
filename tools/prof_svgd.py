"""One crnn_loss_grad_particles call on the SVGD shape (100 x 5) for ncu.  usage: prof_svgd.py [alg]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
import cathode_problem as cp
eng = Engine(0)
pb = cp.make(100, seed=2)
alg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
for _ in range(2):
    r = eng.loss_grad_particles(pb["model"], o, pb["weights"], pb["seeds"], pb["u0"], pb["data"], pb["yscale"], _abi.LOSS_MSE, tab_T=pb["tab_T"])
print("ok", int((r["retcode"] == 1).sum()))
