#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, subprocess, sys, io

KEYS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
 "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
 "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
 "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        out.append(f"kernel: {d.get('Kernel Name', ('?',''))[0][:110]}")
        for k in KEYS:
            if k in d:
                out.append(f"  {k} = {d[k][0]} {d[k][1]}")
        st = sorted(((float(v[0]), h[len(STALL):-len('_per_issue_active.ratio')]) for h, v in d.items()
                     if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')), reverse=True)
        out.append("  stall cycles per issued instruction: " + ", ".join(f"{n}={x:.2f}" for x, n in st[:8]))
    txt = "\n".join(out)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")

if __name__ == "__main__":
    main()
