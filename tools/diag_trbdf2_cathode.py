"""diagnostic: TRBDF2 / AutoTsit5(TRBDF2) on the Cathode F5 model, GPU vs oracle counts"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from oracle import oracle
import cathode_problem as cp

eng = Engine(0)
pb = cp.make(2, seed=1)
with oracle.lu_reciprocal(True), oracle.shared_math(True), oracle.kc4_inverse(True):
    for alg in (_abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_TRBDF2):
        o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
        for e, beta in enumerate(cp.BETAS):
            m, _ = cp.model_for(pb["particles"][0], beta, pb["t_hi"])
            u0 = np.tile(pb["u0"][e], (8, 1)) * (1.0 - 0.01 * np.arange(8))[:, None]
            got = eng.solve_batch(m, o, u0)
            ref = oracle.solve_batch(m, o, u0, n_threads=8)
            print(alg, beta)
            for k in ("n_accept", "n_reject", "n_rhs", "n_jac"):
                print("  ", k, got["stats"][k], ref["stats"][k])
            scale = np.abs(ref["pred"]).max(axis=(0, 1))
            print("   err", (np.abs(got["pred"] - ref["pred"]) / scale).max(axis=(1, 2)))
