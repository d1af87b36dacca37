"""The on-device training loops alone (case1 / case2 / case3 / robertson p2vec kernels, DEVW sensitivity kernels, optimiser kernel),
small — for compute-sanitizer memcheck / racecheck."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from crnn_b200.engine import Engine
from problems import make_problem, trained_p

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
jobs = [("case2", 2, "adam", {}), ("case1", 1, "adam", {}), ("case3", 3, "nadam", {}), ("robertson", 4, "adam", dict(grad_max=10.0))]
for name, kind, optimiser, extra in jobs:
    pb = make_problem(name, golden, 6)
    data = np.abs(pb["data"]) + 1e-6 if name == "case3" else pb["data"]
    ds = eng.dataset(pb["u0"], data)
    nsu = np.array([35, 40, 32, 33, 36, 38]) if name == "robertson" else None
    r = eng.train_steps(pb["model"], pb["opts"], ds, np.arange(6), pb["yscale"], trained_p(name, golden), None, pb["loss_kind"],
                        p2vec_kind=kind, optimiser=optimiser, batch=2, n_save_used=nsu, **extra)
    assert np.isfinite(r["p"]).all() and np.isfinite(r["step_loss"]).all(), name
    ds.close()
eng.close()
print("sanitize_train ok")
