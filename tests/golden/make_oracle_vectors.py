"""Writes tests/golden/oracle_vectors.json: outputs of the CPU oracle on small seeded problems.

Regression pin for the oracle itself and a box-independent reference for the GPU tests (the
oracle's own pinning — checkpoints, Radau, finite differences — lives in test_oracle_cpu.py).
Run:  python tests/golden/make_oracle_vectors.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from crnn_b200 import _abi  # noqa: E402
from oracle import oracle  # noqa: E402
from problems import make_problem  # noqa: E402


def main():
    with open(os.path.join(HERE, "checkpoints.json")) as f:
        golden = json.load(f)
    out = {}
    for name, N in (("case1", 6), ("case2", 8), ("case3", 6), ("robertson", 6)):
        pb = make_problem(name, golden, N)
        v = {"N": N}
        if pb["opts"].alg == _abi.ALG_TSIT5 and pb["seed"].shape[1] <= 63:
            r = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"],
                                       pb["loss_kind"], want_pred=True)
            v["loss"] = r["loss"].tolist(); v["grad_sum"] = r["grad_sum"].tolist()
        else:
            r = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
        v["n_accept"] = r["stats"]["n_accept"].tolist(); v["n_reject"] = r["stats"]["n_reject"].tolist()
        v["pred_every7"] = r["pred"][:, ::7, :].tolist()
        out[name] = v
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote oracle_vectors.json", {k: v["n_accept"] for k, v in out.items()})


if __name__ == "__main__":
    main()
