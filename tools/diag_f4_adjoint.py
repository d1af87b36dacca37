"""F4 (yeast) adjoint gradients, GPU vs oracle, at the default and at tight tolerances (development aid)"""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.integrate import solve_ivp
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from oracle import oracle

g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "checkpoints.json")))
p = np.array(g["yeast"]["p"])
m, seed = cases.yeast_model(p), cases.yeast_seed(p)
rng = np.random.default_rng(4)
N = 4
u0 = cases.YEAST_IC_LB + rng.random((N, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
ts = np.linspace(0, 5, 60)
data = np.array([solve_ivp(cases.yeast_true_rhs, (0, 5), u, method="Radau", rtol=1e-9, atol=1e-12, t_eval=ts).y.T for u in u0])
ys = data.std(axis=1).max(axis=0) + 1e-5
eng = Engine()
for name, kw in (("default", {}), ("1e-8", dict(abstol=1e-10, reltol=1e-8)), ("tight", dict(abstol=1e-12, reltol=1e-10, pred_clamp=(-np.inf, np.inf)))):
    for mode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT):
        o = cases.yeast_opts(alg=0, n_save=60, sens_mode=mode, **kw)
        a = eng.loss_grad_batch(m, o, seed, u0, data, ys, _abi.LOSS_MAE_SCALED)
        b = oracle.loss_grad_batch(m, o, seed, u0, data, ys, _abi.LOSS_MAE_SCALED)
        rel = np.linalg.norm(a["grad_sum"] - b["grad_sum"]) / np.linalg.norm(b["grad_sum"])
        print(name, mode, "ret", a["retcode"], b["retcode"], "acc", a["stats"]["n_accept"], b["stats"]["n_accept"], "rej", a["stats"]["n_reject"],
              b["stats"]["n_reject"], "loss", a["loss"], b["loss"], "rel", rel)
        blocks = {"w_in": slice(0, 144), "w_b": slice(144, 156), "w_out/wJ/slope": slice(156, 164), "mlp": slice(164, 294)}
        for k, s in blocks.items():
            print("   ", k, np.abs(a["grad_sum"][s] - b["grad_sum"][s]).max(), np.abs(b["grad_sum"][s]).max())
