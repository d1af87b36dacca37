"""Quick GPU check of the generic (wide) predict path against the CPU oracle (development aid; the pytest
versions live in tests/test_parity_gpu.py)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CRNN_B200_FORCE_WIDE"] = "1"
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine
from oracle import oracle

eng = Engine(0)
golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "checkpoints.json")))
ALG = {"tsit5": _abi.ALG_TSIT5, "ros23": _abi.ALG_ROSENBROCK23, "auto": _abi.ALG_AUTO_TSIT5_ROS23}


def compare(name, model, opts, u0):
    got = eng.solve_batch(model, opts, u0)
    ref = oracle.solve_batch(model, opts, u0, n_threads=8)
    same = lambda k: int((got["stats"][k] != ref["stats"][k]).sum())
    scale = np.abs(ref["pred"]).max(axis=(0, 1)) + 1e-300
    err = (np.abs(got["pred"] - ref["pred"]) / scale).max()
    print(f"{name:28s} N={u0.shape[0]:5d} ret_mismatch={int((got['retcode'] != ref['retcode']).sum())} "
          f"acc_mismatch={same('n_accept')} rej_mismatch={same('n_reject')} rhs_mismatch={same('n_rhs')} "
          f"jac_mismatch={same('n_jac')} nsaved_mismatch={int((got['n_saved'] != ref['n_saved']).sum())} "
          f"pred_err={err:.2e} steps={ref['stats']['n_accept'].mean():.1f} jac={ref['stats']['n_jac'].mean():.1f}",
          flush=True)


N = 2048
c = cases.CASES["case2"]
m2, _ = c.model(np.array(golden["case2"]["p"]))
u0 = synth.make_u0("case2", N)
for a in ALG:
    compare(f"case2 {a}", m2, c.opts(alg=ALG[a]), u0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from problems import make_problem
cr = cases.CASES["robertson"]
pr = make_problem("robertson", golden, N)
mr, u0r = pr["model"], pr["u0"]
for a in ALG:
    if a == "tsit5":
        continue
    compare(f"robertson crnn {a}", mr, cr.opts(alg=ALG[a]), u0r)
mt = cases.true_model_robertson()
for a in ("ros23", "auto"):
    compare(f"robertson true {a}", mt, cr.opts(alg=ALG[a]), u0r)
c3 = cases.CASES["case3"]
compare("case3 true tsit5", cases.true_model_case3(), c3.opts(), synth.make_u0("case3", N))
ysh = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
u0h = cases.hychem_u0(N)
for stiff in (0.0, 6.0):
    mh, _ = cases.hychem_model(cases.hychem_p(0, stiff=stiff), ysh)
    for a in ALG:
        compare(f"hychem stiff={stiff} {a}", mh, cases.hychem_opts(alg=ALG[a]), u0h)
# throughput of the generic path
import torch
for name, model, opts, mk in (("case2 tsit5", m2, c.opts(), lambda n: synth.make_u0("case2", n)),
                              ("hychem auto", mh, cases.hychem_opts(), cases.hychem_u0)):
    Nb = 65536
    ud = torch.from_numpy(mk(Nb)).cuda()
    for _ in range(2):
        eng.solve_batch(model, opts, ud, want_stats=False)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5):
        eng.solve_batch(model, opts, ud, want_stats=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"wide throughput {name}: {Nb / dt / 1e6:.2f} M traj/s ({dt * 1e3:.2f} ms / {Nb})")
