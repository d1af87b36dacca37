"""diagnostic: gradients through TRBDF2 / AutoTsit5(TRBDF2) on the GPU against the oracle (count agreement, errors)"""
import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from oracle import oracle
from problems import make_problem
import cathode_problem as cp
golden = json.load(open(os.path.join(R, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
with oracle.lu_reciprocal(True), oracle.shared_math(True), oracle.kc4_inverse(True):
    for name, N in (("robertson", 256), ("case2", 128)):
        pb = make_problem(name, golden, N); c = pb["case"]
        for alg in (4, 5):
            o = c.opts(obs_idx=np.arange(c.ns), alg=alg)
            args = (pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
            got = eng.loss_grad_batch(*args, want_pred=True); ref = oracle.loss_grad_batch(*args, want_pred=True, n_threads=8)
            same = np.ones(N, dtype=bool)
            for k in ("n_accept", "n_reject", "n_rhs", "n_jac"): same &= got["stats"][k] == ref["stats"][k]
            print(name, alg, "same", same.mean(), "loss rel", np.abs(got["loss"] / ref["loss"] - 1)[same].max(),
                  "grad rel", np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"]),
                  "rhs/traj", got["stats"]["n_rhs"].mean(), "njac", got["stats"]["n_jac"].mean())
    pb = cp.make(2, seed=1)
    for alg in (4, 5, 1):
        o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
        for e, beta in enumerate(cp.BETAS):
            m, sd = cp.model_for(pb["particles"][0], beta, pb["t_hi"])
            u0 = np.tile(pb["u0"][e], (8, 1)) * (1.0 - 0.01 * np.arange(8))[:, None]; data = np.tile(pb["data"][e], (8, 1, 1))
            got = eng.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True)
            ref = oracle.loss_grad_batch(m, o, sd, u0, data, pb["yscale"], _abi.LOSS_MSE, want_pred=True, n_threads=8)
            same = (got["stats"]["n_rhs"] == ref["stats"]["n_rhs"]) & (got["stats"]["n_accept"] == ref["stats"]["n_accept"])
            print("cathode", alg, beta, "same", same.mean(), "grad rel", np.linalg.norm(got["grad_sum"] - ref["grad_sum"]) / np.linalg.norm(ref["grad_sum"]),
                  "acc", got["stats"]["n_accept"].mean(), ref["stats"]["n_accept"].mean(), "rej", got["stats"]["n_reject"].mean())
