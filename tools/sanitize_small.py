"""One small call of every kernel family (run under compute-sanitizer: memcheck / racecheck / initcheck)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine
from problems import make_problem

N = int(os.environ.get("SAN_N", "48"))
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
YS = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
done = []
for name in ("case2", "case1", "case3", "robertson"):
    pb = make_problem(name, golden, N)
    eng.solve_batch(pb["model"], pb["opts"], pb["u0"]); done.append(f"{name} value")
    r = eng.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"], want_pred=True)
    done.append(f"{name} forward sens")
    if name != "robertson":
        for sm in (_abi.SENS_INTERP_ADJOINT, _abi.SENS_DISCRETE_ADJOINT):
            o = pb["case"].opts(obs_idx=pb["opts"].obs_idx, sens_mode=sm)
            eng.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
        done.append(f"{name} adjoints")
    o = pb["case"].opts(obs_idx=pb["opts"].obs_idx, alg=_abi.ALG_KENCARP4)
    eng.solve_batch(pb["model"], o, pb["u0"]); done.append(f"{name} kencarp4")
    eng.solve_batch(pb["model"], pb["case"].opts(obs_idx=pb["opts"].obs_idx, alg=_abi.ALG_AUTO_TSIT5_ROS23), pb["u0"])
    done.append(f"{name} auto (thread per trajectory)")
    os.environ["CRNN_B200_FORCE_WIDE"] = "1"
    for alg in (_abi.ALG_TSIT5, _abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_ROS23):
        if name == "robertson" and alg == _abi.ALG_TSIT5:
            continue
        eng.solve_batch(pb["model"], pb["case"].opts(obs_idx=pb["opts"].obs_idx, alg=alg), pb["u0"])
    os.environ["CRNN_B200_FORCE_WIDE"] = "0"
    done.append(f"{name} generic path")
m, seed = cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS)
u0 = cases.hychem_u0(N)
for alg in (_abi.ALG_TSIT5, _abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_KENCARP4):
    pr = eng.solve_batch(m, cases.hychem_opts(alg=alg), u0)
for smode in (_abi.SENS_INTERP_ADJOINT, _abi.SENS_DISCRETE_ADJOINT):
    eng.loss_grad_batch(m, cases.hychem_opts(alg=_abi.ALG_TSIT5, sens_mode=smode), seed, u0, pr["pred"] * 1.01, YS)
done.append("hychem F2 all")
# round 2: generic forward-sensitivity kernel (all three algorithms, F2 with 211 columns), datasets, particles
os.environ["CRNN_B200_FORCE_GENERIC"] = "1"
for name, alg in (("case2", _abi.ALG_AUTO_TSIT5_ROS23), ("robertson", _abi.ALG_ROSENBROCK23), ("case1", _abi.ALG_TSIT5)):
    pb = make_problem(name, golden, min(N, 16))
    o = pb["case"].opts(obs_idx=pb["opts"].obs_idx, alg=alg)
    eng.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"], want_pred=True)
os.environ["CRNN_B200_FORCE_GENERIC"] = "0"
mh, sh = cases.hychem_model(cases.hychem_p(0, stiff=4.0), YS)
for alg in (_abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_ROS23):
    eng.loss_grad_batch(mh, cases.hychem_opts(alg=alg), sh, u0[:4], pr["pred"][:4] * 1.01, YS)
done.append("generic forward sens")
pb = make_problem("case2", golden, N)
ds = eng.dataset(pb["u0"], pb["data"])
eng.loss_grad_indexed(pb["model"], pb["opts"], pb["seed"], ds, pb["yscale"], pb["loss_kind"], idx=np.arange(N)[::-2].copy(), want_loss=True)
ds.close(); done.append("dataset indexed")
import cathode_problem as cp
cpb = cp.make(3, seed=1)
eng.loss_grad_particles(cpb["model"], cpb["opts"], cpb["weights"], cpb["seeds"], cpb["u0"], cpb["data"], cpb["yscale"], _abi.LOSS_MSE, tab_T=cpb["tab_T"])
done.append("particles F5 + observable")
ms = cases.synthetic_stiff_model()
eng.solve_batch(ms, cases.synthetic_stiff_opts(), cases.synthetic_stiff_u0(N)); done.append("kencarp4 30-state")
# this session: TRBDF2 / AutoTsit5(TRBDF2) incl. the heat-release post-map, F4 (MLP inputs, finite-difference Jacobian), on-device training loop
for alg in (_abi.ALG_TRBDF2, _abi.ALG_AUTO_TSIT5_TRBDF2):
    eng.solve_batch(m, cases.hychem_opts(alg=alg), u0)
    mc, _ = cp.model_for(cpb["particles"][0], 10.0, cpb["t_hi"])
    eng.solve_batch(mc, cases.cathode_opts(cpb["opts"].saveat, alg=alg, pred_clamp=(-np.inf, np.inf)), np.tile(cpb["u0"][0], (8, 1)))
    eng.solve_batch(make_problem("robertson", golden, 8)["true_model"], cases.CASES["robertson"].opts(alg=alg), make_problem("robertson", golden, 8)["u0"])
done.append("trbdf2 + composite + observable")
my = cases.yeast_model(np.array(golden["yeast"]["p"]))
uy = cases.YEAST_IC_LB + np.random.default_rng(0).random((min(N, 16), 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
for alg in (_abi.ALG_TSIT5, _abi.ALG_ROSENBROCK23, _abi.ALG_AUTO_TSIT5_TRBDF2, _abi.ALG_TRBDF2):
    eng.solve_batch(my, cases.yeast_opts(alg=alg, n_save=40), uy)
done.append("F4 yeast")
dy = eng.solve_batch(my, cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=40), uy * 1.02)["pred"]
for smode in (_abi.SENS_DISCRETE_ADJOINT, _abi.SENS_INTERP_ADJOINT):   # incl. forward steps beyond the shared-memory record
    eng.loss_grad_batch(my, cases.yeast_opts(alg=_abi.ALG_TSIT5, n_save=40, sens_mode=smode), cases.yeast_seed(np.array(golden["yeast"]["p"])),
                        uy, dy, np.ones(7))
done.append("F4 yeast adjoint gradients")
pb = make_problem("case2", golden, 20)
ds = eng.dataset(pb["u0"], pb["data"])
eng.train_steps(pb["model"], pb["opts"], ds, np.arange(20)[::-1].copy(), pb["yscale"], np.array(golden["case2"]["p"]), None, pb["loss_kind"],
                batch=2, eta=5e-3, weight_decay=1e-6, expdecay=(5e-3, 0.5, 3, 1e-4), grad_max=0.5)
ds.close(); done.append("train_steps")
pb3 = make_problem("case3", golden, 6)
ds3 = eng.dataset(pb3["u0"], np.abs(pb3["data"]) + 1e-6)
from problems import trained_p
eng.train_steps(pb3["model"], pb3["opts"], ds3, np.arange(6), pb3["yscale"], trained_p("case3", golden), None, pb3["loss_kind"],
                p2vec_kind=3, optimiser="nadam", batch=2)
ds3.close(); done.append("train_steps case3 (five warps per trajectory)")
pbr = make_problem("robertson", golden, 6)
dsr = eng.dataset(pbr["u0"], pbr["data"])
eng.train_steps(pbr["model"], pbr["opts"], dsr, np.arange(6), pbr["yscale"], np.array(golden["robertson"]["p"]), None, pbr["loss_kind"],
                p2vec_kind=4, batch=2, grad_max=10.0, n_save_used=np.array([35, 40, 32, 33, 36, 38]))
dsr.close(); done.append("train_steps robertson (Rosenbrock23 sensitivities)")
eng.close()
print("sanitize_small ok:", "; ".join(done))
