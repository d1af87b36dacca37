#!/usr/bin/env python
"""Times the generic-dimension kernels on BASELINE config 5's shapes and on the HyChem F2 model (device buffers,
CUDA events), with the CPU oracle on all host cores beside each (bounded sample).
usage: python tools/measure_configs.py [out.json]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine, stats_from_torch
from oracle import oracle

YS = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
ALGN = {_abi.ALG_TSIT5: "Tsit5", _abi.ALG_ROSENBROCK23: "Rosenbrock23", _abi.ALG_KENCARP4: "KenCarp4",
        _abi.ALG_AUTO_TSIT5_ROS23: "AutoTsit5(Rosenbrock23)"}


def time_gpu(fn, K=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, r


def main():
    eng = Engine(0)
    cores = os.cpu_count() or 1
    out = {}

    def value(tag, model, mk_u0, opts, N, n_cpu):
        u0 = mk_u0(N)
        ud = torch.from_numpy(u0).cuda()
        ms, r = time_gpu(lambda: eng.solve_batch(model, opts, ud))
        st = stats_from_torch(r["stats"])
        t = time.perf_counter()
        oracle.solve_batch(model, opts, u0[:n_cpu], n_threads=cores)
        cpu = n_cpu / (time.perf_counter() - t)
        out[tag] = {"N": N, "alg": ALGN[opts.alg], "ms": ms, "traj_per_s": N / ms * 1e3, "rhs_per_s": float(st["n_rhs"].sum()) / ms * 1e3,
                    "steps_mean": float((st["n_accept"] + st["n_reject"]).mean()), "n_jac_mean": float(st["n_jac"].mean()),
                    "success_frac": float((r["retcode"] == 1).float().mean().item()),
                    "cpu_oracle_traj_per_s": cpu, "cpu_cores": cores, "cpu_sample": n_cpu}
        print(tag, out[tag], flush=True)

    # BASELINE config 5: "HyChem-sized" 30-state stiff model, KenCarp4, 131 072 ICs (16 384 per GPU on 8)
    ms30 = cases.synthetic_stiff_model()
    for N in (16384, 131072):
        value(f"stiff30_kencarp4_{N}", ms30, cases.synthetic_stiff_u0, cases.synthetic_stiff_opts(), N, 4096)
    # the reference-shaped HyChem model (ns = 9, nr = 10, F2): the script's own initialisation (mildly stiff)
    mh, seedh = cases.hychem_model(cases.hychem_p(0), YS)
    for alg in (_abi.ALG_AUTO_TSIT5_ROS23, _abi.ALG_ROSENBROCK23, _abi.ALG_KENCARP4, _abi.ALG_TSIT5):
        value(f"hychem_f2_{ALGN[alg]}", mh, cases.hychem_u0, cases.hychem_opts(alg=alg), 65536, 8192)
    # the composite algorithm on the dimension-specialised thread-per-trajectory kernel (k_auto_value) and, for contrast,
    # on the generic lane-per-component kernel (3 of 32 lanes busy for robertson)
    cr = cases.CASES["robertson"]
    value("robertson_true_auto", cases.true_model_robertson(), lambda n: synth.make_u0("robertson", n),
          cr.opts(alg=_abi.ALG_AUTO_TSIT5_ROS23, pred_clamp=(-np.inf, np.inf)), 262144, 16384)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
    c2 = cases.CASES["case2"]
    m2, _ = c2.model(np.array(golden["case2"]["p"]))
    value("case2_auto", m2, lambda n: synth.make_u0("case2", n), c2.opts(alg=_abi.ALG_AUTO_TSIT5_ROS23), 65536, 16384)
    os.environ["CRNN_B200_FORCE_WIDE"] = "1"
    value("robertson_true_auto_generic_kernel", cases.true_model_robertson(), lambda n: synth.make_u0("robertson", n),
          cr.opts(alg=_abi.ALG_AUTO_TSIT5_ROS23, pred_clamp=(-np.inf, np.inf)), 262144, 16384)
    value("case2_auto_generic_kernel", m2, lambda n: synth.make_u0("case2", n), c2.opts(alg=_abi.ALG_AUTO_TSIT5_ROS23), 65536, 16384)
    os.environ["CRNN_B200_FORCE_WIDE"] = "0"
    # HyChem gradient (np = 211) by the adjoints, non-stiff variant
    mg, seedg = cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS)
    N = 65536
    u0 = cases.hychem_u0(N); ud = torch.from_numpy(u0).cuda()
    data = eng.solve_batch(cases.hychem_model(cases.hychem_p(1, lnA_shift=-2.0), YS)[0], cases.hychem_opts(alg=_abi.ALG_ROSENBROCK23), ud)["pred"]
    for mode, sm in (("discrete", _abi.SENS_DISCRETE_ADJOINT), ("interp", _abi.SENS_INTERP_ADJOINT)):
        o = cases.hychem_opts(alg=_abi.ALG_TSIT5, sens_mode=sm)
        ms, r = time_gpu(lambda: eng.loss_grad_batch(mg, o, seedg, ud, data, YS))
        n_cpu = 2048
        of = cases.hychem_opts(alg=_abi.ALG_TSIT5, sens_mode=_abi.SENS_FORWARD)
        t = time.perf_counter()
        oracle.loss_grad_batch(mg, of, seedg, u0[:n_cpu], data[:n_cpu].cpu().numpy(), YS, n_threads=cores)
        cpu = n_cpu / (time.perf_counter() - t)
        out[f"hychem_f2_grad_np211_{mode}"] = {"N": N, "ms": ms, "traj_per_s": N / ms * 1e3,
                                               "cpu_oracle_forward_mode_traj_per_s": cpu, "cpu_cores": cores, "cpu_sample": n_cpu}
        print(mode, out[f"hychem_f2_grad_np211_{mode}"], flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
