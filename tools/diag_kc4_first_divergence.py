"""diagnostic: first step at which k_kencarp4_wide and the oracle differ in (t_reached, dt_last, counts) on a mismatching trajectory"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from dataclasses import replace
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
from oracle import oracle
eng = Engine(0)
m = cases.synthetic_stiff_model(); o = cases.synthetic_stiff_opts(); u0 = cases.synthetic_stiff_u0(4096)
with oracle.lu_reciprocal(True), oracle.shared_math(True), oracle.kc4_inverse(True):
    got = eng.solve_batch(m, o, u0); ref = oracle.solve_batch(m, o, u0, n_threads=16)
    bad = np.nonzero((got["stats"]["n_rhs"] != ref["stats"]["n_rhs"]) | (got["stats"]["n_accept"] != ref["stats"]["n_accept"]))[0]
    print("mismatching", bad[:10], len(bad))
    for tr in bad[:4]:
        u = u0[tr:tr + 1]
        for k in range(1, 80):
            ok = replace(o, maxiters=k)
            g = eng.solve_batch(m, ok, u); r = oracle.solve_batch(m, ok, u)
            gs, rs = g["stats"][0], r["stats"][0]
            same = all(gs[f] == rs[f] for f in ("n_accept", "n_reject", "n_rhs", "n_jac")) and gs["t_reached"] == rs["t_reached"] and gs["dt_last"] == rs["dt_last"]
            if not same:
                print("traj", tr, "first difference at attempt", k, "gpu", gs, "oracle", rs)
                if k > 1:
                    ok1 = replace(o, maxiters=k - 1)
                    g1 = eng.solve_batch(m, ok1, u); r1 = oracle.solve_batch(m, ok1, u)
                    print("   previous attempt: gpu", g1["stats"][0], "oracle", r1["stats"][0])
                break
