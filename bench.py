#!/usr/bin/env python
"""bench.py — trajectories/s of the CRNN hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] — case2 (6 species + T, 3 reactions,
np = 25), Tsit5 + forward sensitivities + fused MAE loss, 65 536 seeded random ICs PER GPU
(weak scaling), trained checkpoint weights, targets = generating mechanism x (1 + 5 % noise).
One "step" = one pass of loss + gradient over the batch (one optimiser step's worth of work);
with N > 1 ranks each rank owns its own 65 536 trajectories and the only exchange is the NCCL
all-reduce of [sum loss, grad_sum] (26 doubles).

  value  : whole-job trajectories/s with inputs resident in HBM (device buffers)
  e2e    : same metric through the public API with HOST (pinned) buffers; the H2D copy of
           u0 + targets and the D2H read of loss/grad are inside the timed region
  roofline: dominant kernel (k_tsit5_sens) vs the measured HBM peak; algorithmic bytes are
           4 864 B/trajectory.  The kernel is fp64-ALU bound, so the fp64 FMA
           fraction is reported next to it.
  cpu_baseline / --impl reference: the CPU oracle (a port: Julia cannot run here) on the
           host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 65536
BYTES_PER_TRAJ = 8 * (7 + 2 * 6 * 50 + 1)   # u0 + targets read + saved states written + loss
FP64_PEAK_TFLOPS = 37.106   # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r1_fp64_peak.json)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_tsit5_sens launch of this exact workload, from the committed
# `ncu --set full` capture (profiles/r1_sens_v11.txt): 161.9 MB + 129.1 MB; algorithmic bytes are 318.8 MB
NCU_DRAM_BYTES_PER_LAUNCH = 291.04e6
METRIC = "trajectories/sec (case2: Tsit5 + 25 forward sensitivities + fused loss, 65 536 ICs per B200)"


def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")) as f:
        return json.load(f)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clocks + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The
    timed region is ~150 ms, shorter than nvidia-smi's start-up, so NVML is polled in-process every
    5 ms from a thread; `nvidia-smi -lms` is the fallback when the binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.sm, self.reasons, self.smax, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thr = None
        self.nv = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            except Exception:
                pass
            hnd = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        hnd = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        hnd = None
            if hnd is None:
                hnd = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv, self.hnd = nv, hnd
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(hnd, nv.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, hnd = self.nv, self.hnd
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(hnd))
                for nm, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(nm)
                self.power.append(nv.nvmlDeviceGetPowerUsage(hnd) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nv is not None:
            self._stop.set()
            self._thr.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "power_w_max": max(self.power) if self.power else None, "source": "nvml, 5 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


def build_inputs(eng, rank):
    """Seeded ICs for this rank's shard, targets from the generating mechanism (solved on the
    GPU by the engine itself), trained-checkpoint model + seed."""
    from crnn_b200 import cases, synth
    c = cases.CASES["case2"]
    start = rank * N_PER_GPU
    u0 = synth.make_u0("case2", N_PER_GPU, start=start)
    obs = np.arange(c.ns)
    truth = eng.solve_batch(cases.true_model_case2(), c.opts(obs_idx=obs, pred_clamp=(-np.inf, np.inf)), u0,
                            want_stats=False)
    assert (truth["retcode"] == 1).all()
    data = synth.noisy_targets(truth["pred"], 0.05, start=start)
    # yscale is a dataset constant: every rank derives it from the same first 1024 trajectories
    u0_ys = synth.make_u0("case2", 1024)
    tr_ys = eng.solve_batch(cases.true_model_case2(), c.opts(obs_idx=obs, pred_clamp=(-np.inf, np.inf)), u0_ys,
                            want_stats=False)
    yscale = synth.yscale_from(synth.noisy_targets(tr_ys["pred"], 0.05), c.lb)
    model, seed = c.model(np.array(load_golden()["case2"]["p"]))
    return c, model, seed, c.opts(obs_idx=obs), u0, data, yscale


def cpu_reference(steps, warmup, sample):
    """The reference arm / cpu_baseline: the CPU oracle (port) on all host cores."""
    from crnn_b200 import cases, synth
    from oracle import oracle
    c = cases.CASES["case2"]
    cores = os.cpu_count() or 1
    u0 = synth.make_u0("case2", sample)
    obs = np.arange(c.ns)
    truth = oracle.solve_batch(cases.true_model_case2(), c.opts(obs_idx=obs, pred_clamp=(-np.inf, np.inf)), u0,
                               n_threads=cores)["pred"]
    data = synth.noisy_targets(truth, 0.05)
    ys = synth.yscale_from(data, c.lb)
    model, seed = c.model(np.array(load_golden()["case2"]["p"]))
    opts = c.opts(obs_idx=obs)
    for _ in range(warmup):
        oracle.loss_grad_batch(model, opts, seed, u0[:256], data[:256], ys, n_threads=cores)
    t = time.perf_counter()
    for _ in range(steps):
        oracle.loss_grad_batch(model, opts, seed, u0, data, ys, n_threads=cores)
    dt = (time.perf_counter() - t) / steps
    return sample / dt, dt * 1e3, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=N_PER_GPU)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "case2 (ns=6+T, nr=3, np=25) Tsit5 + forward sensitivities + fused MAE loss, "
                          f"{N_PER_GPU} ICs per GPU, abstol 1e-6 reltol 1e-3, t in [0,50], 50 saves",
              "n_traj_per_gpu": N_PER_GPU, "parallelism": f"dp{a.gpus} (trajectory shards, all-reduce of [loss, grad])",
              "l2": "inputs larger than L2 (157 MB targets read + 157 MB saved states written per step)"}

    if a.impl == "reference":
        if rank != 0:
            return
        sample = min(a.cpu_sample, N_PER_GPU)
        v, ms, cores = cpu_reference(max(1, a.steps), a.warmup, sample)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "trajectories/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of the {N_PER_GPU} case2 trajectories per step, OpenMP over all host cores "
                                       "(CPU restatement of the reference's algorithm, not Julia)"},
            "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # stdout carries ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on the GPU boxes; WARN prints it too)
    # off it; NCCL caches the setting at its first call, so this has to happen before torch is imported
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]
    import torch
    import torch.distributed as dist
    from crnn_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (crnn_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    c, model, seed, opts, u0_h, data_h, yscale = build_inputs(eng, rank)
    dev = torch.device("cuda", local_rank)
    u0_d = torch.from_numpy(u0_h).to(dev)
    data_d = torch.from_numpy(data_h).to(dev)
    red = torch.zeros(seed.shape[1] + 1, dtype=torch.float64, device=dev)

    def step_device():
        r = eng.loss_grad_batch(model, opts, seed, u0_d, data_d, yscale, c.loss_kind, want_stats=False, want_pred=True)
        if world > 1:   # the one exchange of the path: [sum loss, grad_sum] over NVLink
            red[0] = r["loss"].sum(); red[1:] = r["grad_sum"]
            dist.all_reduce(red)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        r = step_device()
    barrier()
    # ---- timed region 1: device-resident inputs ----
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = eng.launch_count
    eng.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        r = step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    kern_ms, kern_n = eng.profile_end()
    launches = eng.launch_count - l0
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / a.steps
    value = N_PER_GPU * world / (ms_step * 1e-3)
    # RHS evaluations per trajectory from the engine's own counters (one extra untimed pass)
    rs = eng.loss_grad_batch(model, opts, seed, u0_d, data_d, yscale, c.loss_kind, want_stats=True)
    from crnn_b200.engine import stats_from_torch
    st = stats_from_torch(rs["stats"])
    rhs_per_traj = float(st["n_rhs"].mean())
    attempts_per_traj = float((st["n_accept"] + st["n_reject"]).mean())

    # ---- timed region 2: end to end with host (pinned) buffers through the public API ----
    u0_p = torch.from_numpy(u0_h).pin_memory(); data_p = torch.from_numpy(data_h).pin_memory()
    u0_n, data_n = u0_p.numpy(), data_p.numpy()

    def step_host():
        r = eng.loss_grad_batch(model, opts, seed, u0_n, data_n, yscale, c.loss_kind, want_stats=False)
        if world > 1:
            red.copy_(torch.from_numpy(np.concatenate([[r["loss"].sum()], r["grad_sum"]])))
            dist.all_reduce(red)
            red.cpu()
        return r

    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(a.steps, 10))
    for _ in range(e2e_steps):
        rh = step_host()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N_PER_GPU * world / (float(te.item()) / e2e_steps * 1e-3)
    h2d = u0_h.nbytes + data_h.nbytes
    d2h = N_PER_GPU * (8 + 4 + 4) + 8 * seed.shape[1]

    if rank == 0:
        peak, which = measured_peaks()
        kms = kern_ms / max(1, kern_n)
        achieved = BYTES_PER_TRAJ * N_PER_GPU / (kms * 1e-3) / 1e9
        # secondary (and binding) roofline: algorithmic fp64 flops of the path (DESIGN.md "flop model") against the
        # measured DFMA peak (profiles/r1_fp64_peak.json, tools/fp64_peak.cu)
        ns, nr, nin, ncol, nsave = 6, 3, 7, seed.shape[1] + 1, 50
        f_rhs = (2 * nin * nr + 2 * ns * nr + 40 * ns + 30 * nr) + ncol * (ns + 4 * ns * nr + 3 * nr + 2)
        f_step = ncol * ns * 2 * (21 + 6 + 7 + 2)
        f_save = ncol * ns * 2 * 9
        flop_traj = rhs_per_traj * f_rhs + attempts_per_traj * f_step + nsave * f_save
        fp64_tflops = flop_traj * N_PER_GPU / (kms * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "rhs_evals_per_s": value * rhs_per_traj, "rhs_evals_per_traj": rhs_per_traj,
            "loss_mean": float(r["loss"].mean().item()),
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "peak_source": which,
                         "kernel": "k_tsit5_sens<Cfg<6,3,1>,1,...>", "kernel_ms": kms, "kernel_launches": int(kern_n),
                         "fp64": {"achieved": fp64_tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                  "frac": fp64_tflops / FP64_PEAK_TFLOPS, "flop_per_trajectory": flop_traj,
                                  "peak_source": "measured DFMA microbenchmark (profiles/r1_fp64_peak.json)"},
                         "note": "HBM is the nominal roofline of the task's taxonomy; the kernel is fp64-issue bound by "
                                 "construction (~2e2 flop/B), so the fp64 fraction is the meaningful one; see DESIGN.md"},
        }
        if a.gpus == 1:
            v, ms, cores = cpu_reference(1, 1, min(a.cpu_sample, N_PER_GPU))
            out["cpu_baseline"] = {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port",
                                   "sample": f"{min(a.cpu_sample, N_PER_GPU)} of the {N_PER_GPU} trajectories, one pass, "
                                             "OpenMP over all host cores"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
