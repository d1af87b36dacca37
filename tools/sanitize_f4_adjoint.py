"""The F4 adjoint kernels alone, small (for compute-sanitizer memcheck / racecheck / initcheck / synccheck)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from crnn_b200 import cases, _abi
from crnn_b200.engine import Engine
from test_f4_mlp_cpu import qssa_like_model

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
p = np.array(golden["yeast"]["p"])
my = cases.yeast_model(p)
uy = cases.YEAST_IC_LB + np.random.default_rng(0).random((12, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
dy = eng.solve_batch(my, cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=40), uy * 1.02)["pred"]
for smode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT):
    r = eng.loss_grad_batch(my, cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=40, sens_mode=smode), cases.yeast_seed(p), uy, dy, np.ones(7))
    assert (r["retcode"] == 1).all() and np.isfinite(r["grad_sum"]).all()
q = qssa_like_model()
from crnn_b200.model import SolveOpts
u0 = 0.2 + np.random.default_rng(3).random((12, 3))
for smode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT):
    o = SolveOpts(saveat=np.linspace(0.0, 2.0, 21), t0=0.0, t1=2.0, alg=0, abstol=1e-8, reltol=1e-6, maxiters=100000, sens_mode=smode)
    r = eng.loss_grad_batch(q, o, np.eye(q.n_w), u0, np.ones((12, 21, 3)), np.ones(3))
    assert (r["retcode"] == 1).all() and np.isfinite(r["grad_sum"]).all()
eng.close()
print("sanitize_f4_adjoint ok")
