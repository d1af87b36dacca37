#!/usr/bin/env python
"""gpurun_out/r1_{train,value,configs}.json + a bench line -> profiles/r1_throughput_all_configs.txt"""
import json, sys
t = json.load(open('gpurun_out/r1_train.json')); v = json.load(open('gpurun_out/r1_value.json')); c = json.load(open('gpurun_out/r1_configs.json'))
b = json.loads(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r1c_bench.json').read().strip().splitlines()[-1])
L = ["# one B200, device-resident inputs, CUDA-event timed, after warm-up (tools/measure_train.py, measure_value.py, measure_configs.py; bench.py for the headline)",
     f"# headline (bench.py): {b['value']/1e6:.2f} M traj/s device-resident, e2e {b['e2e']['value']/1e6:.2f} M traj/s, {b['rhs_evals_per_s']/1e9:.2f}e9 RHS evals/s, "
     f"cpu oracle {b['cpu_baseline']['value']/1e3:.1f} K traj/s on {b['cpu_baseline']['cores']} cores; clocks {b['clocks']}",
     "# loss+gradient kernels (65 536 trajectories)"]
names = {'case2_forward': 'case2 forward (np=25, Tsit5 + fwd sens)', 'case2_adjoint': 'case2 interpolating adjoint', 'case2_discrete': 'case2 discrete adjoint',
         'robertson_forward': 'robertson forward (np=43, Rosenbrock23 sens)', 'case3_forward': 'case3 forward (np=153, 5 warps/trajectory)',
         'case3_adjoint': 'case3 interpolating adjoint', 'case3_discrete': 'case3 discrete adjoint'}
for k, n in names.items():
    r = t[k]; L.append(f"{n:46s} {r['ms']:8.2f} ms {r['traj_per_s']/1e6:6.2f} M traj/s  {r['rhs_per_s']:.2e} RHS evals/s  {r['steps_mean']:.1f} steps/traj")
L.append("# predict kernels (thread per trajectory, dimension-specialised)")
for k in ('case2', 'robertson', 'case3'):
    r = v[k]; L.append(f"{k:10s} {r['N']:8d} {r['ms']:7.2f} ms {r['traj_per_s']/1e6:7.1f} M traj/s  {r['rhs_per_s']:.2e} RHS evals/s  {r['hbm_GBps_algorithmic']:.0f} GB/s algorithmic")
L.append("# other algorithms / the generic-dimension kernels (warp per trajectory, lane = state component), CPU oracle on the box's host cores beside them")
for k, r in c.items():
    if 'alg' in r:
        L.append(f"{k:36s} {r['N']:7d} {r['ms']:8.2f} ms {r['traj_per_s']/1e6:7.3f} M traj/s  {r['rhs_per_s']:.2e} RHS evals/s  {r['steps_mean']:.1f} steps ({r['n_jac_mean']:.1f} stiff)  "
                 f"cpu oracle {r['cpu_oracle_traj_per_s']/1e3:.1f} K traj/s on {r['cpu_cores']} cores ({r['traj_per_s']/r['cpu_oracle_traj_per_s']:.0f}x)")
    else:
        L.append(f"{k:36s} {r['N']:7d} {r['ms']:8.2f} ms {r['traj_per_s']/1e6:7.3f} M traj/s  cpu oracle forward mode (211 columns) {r['cpu_oracle_forward_mode_traj_per_s']/1e3:.2f} K traj/s "
                 f"on {r['cpu_cores']} cores ({r['traj_per_s']/r['cpu_oracle_forward_mode_traj_per_s']:.0f}x)")
import os
if os.path.exists('gpurun_out/r1_config4_full.json'):
    f4 = json.load(open('gpurun_out/r1_config4_full.json'))
    L.append(f"# BASELINE config 4 at its full size on ONE B200: case3 (np = 153), {f4['N']} ICs, {f4['targets_GB']:.2f} GB of targets resident in HBM (tools/measure_config4_full.py)")
    for k in ('interp_adjoint', 'discrete_adjoint', 'forward_153_columns'):
        r = f4[k]; L.append(f"case3 {k:22s} {f4['N']:8d} {r['ms']:9.1f} ms {r['traj_per_s']/1e6:6.2f} M traj/s  success {r['success_frac']:.3f}")
open('profiles/r1_throughput_all_configs.txt', 'w').write("\n".join(L) + "\n")
print("\n".join(L))
