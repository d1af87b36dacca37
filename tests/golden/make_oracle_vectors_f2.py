"""Writes tests/golden/oracle_vectors_f2_auto.json: outputs of the CPU oracle for the F2 (HyChem) RHS, the composite
AutoTsit5(Rosenbrock23) algorithm and the reversible CRNN — the regression pin of those parts of the oracle
(their own pinning — literal formulas, finite differences, Radau — lives in test_oracle_f2_auto_cpu.py).
Run:  python tests/golden/make_oracle_vectors_f2.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from crnn_b200 import _abi, cases, synth  # noqa: E402
from oracle import oracle  # noqa: E402
from problems import make_problem  # noqa: E402

YS = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
ALG = {"tsit5": _abi.ALG_TSIT5, "ros23": _abi.ALG_ROSENBROCK23, "auto": _abi.ALG_AUTO_TSIT5_ROS23, "kc4": _abi.ALG_KENCARP4}


def problems():
    """name -> (model, opts, u0)"""
    with open(os.path.join(HERE, "checkpoints.json")) as f:
        golden = json.load(f)
    out = {}
    mh, _ = cases.hychem_model(cases.hychem_p(0), YS)
    for a, alg in ALG.items():
        out[f"hychem_{a}"] = (mh, cases.hychem_opts(alg=alg), cases.hychem_u0(4))
    pr = make_problem("robertson", golden, 4)
    out["robertson_true_auto"] = (pr["true_model"], pr["case"].opts(alg=ALG["auto"], pred_clamp=(-np.inf, np.inf)), pr["u0"])
    out["robertson_crnn_auto"] = (pr["model"], pr["case"].opts(alg=ALG["auto"]), pr["u0"])
    return out


def gradient_problem():
    m, seed = cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS)
    u0 = cases.hychem_u0(3)
    data = oracle.solve_batch(cases.hychem_model(cases.hychem_p(1, lnA_shift=-2.0), YS)[0], cases.hychem_opts(alg=ALG["ros23"]), u0)["pred"]
    return m, seed, u0, data


def main():
    out = {}
    for name, (m, o, u0) in problems().items():
        r = oracle.solve_batch(m, o, u0)
        out[name] = {k: r["stats"][k].tolist() for k in ("n_accept", "n_reject", "n_rhs", "n_jac")}
        out[name]["pred_every7"] = r["pred"][:, ::7, :].tolist()
    m, seed, u0, data = gradient_problem()
    for mode, sm in (("forward", _abi.SENS_FORWARD), ("discrete", _abi.SENS_DISCRETE_ADJOINT), ("interp", _abi.SENS_INTERP_ADJOINT)):
        r = oracle.loss_grad_batch(m, cases.hychem_opts(alg=ALG["tsit5"], sens_mode=sm), seed, u0, data, YS)
        out[f"hychem_grad_{mode}"] = {"loss": r["loss"].tolist(), "grad_sum": r["grad_sum"].tolist()}
    with open(os.path.join(HERE, "oracle_vectors_f2_auto.json"), "w") as f:
        json.dump(out, f)
    print("wrote oracle_vectors_f2_auto.json", {k: v.get("n_accept") for k, v in out.items()})


if __name__ == "__main__":
    main()
