"""Times crnn_loss_grad_particles on the SVGD shape (100 particles x 5 heating rates, 17 parameters, heat-release MSE) per algorithm,
with the oracle's particle-by-particle loop beside it (a sample of particles)."""
import sys, os, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine
import cathode_problem as cp
eng = Engine(0)
pb = cp.make(100, seed=2)
out = {}
for name, alg in (("Rosenbrock23", 1), ("AutoTsit5(Rosenbrock23)", 3), ("TRBDF2", 4), ("AutoTsit5(TRBDF2)", 5)):
    o = cases.cathode_opts(pb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf))
    f = lambda: eng.loss_grad_particles(pb["model"], o, pb["weights"], pb["seeds"], pb["u0"], pb["data"], pb["yscale"], _abi.LOSS_MSE, tab_T=pb["tab_T"], want_stats=True)
    r = f(); f()
    t0 = time.perf_counter()
    for _ in range(5): r = f()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    cp.oracle_particles(dict(pb, opts=o), _abi.LOSS_MSE, idx=list(range(4)))
    cpu_ms = (time.perf_counter() - t0) * 1e3 * 25      # 4 of 100 particles, one core (the reference's loop is sequential)
    out[name] = {"ms_per_svgd_gradient_call": ms, "oracle_one_core_ms_extrapolated": cpu_ms, "ok": bool((r["retcode"] == 1).all()),
                 "attempts_mean": float((r["stats"]["n_accept"] + r["stats"]["n_reject"]).mean())}
    print(name, out[name])
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
