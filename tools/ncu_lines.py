#!/usr/bin/env python
"""Per-CUDA-source-line executed-instruction / stall-sample / smem-wavefront shares from an .ncu-rep
captured with --import-source on (kernels built with -lineinfo).
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    agg = {}; fname = "?"; hdr = None; cur = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No":
            hdr = r; ia = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
            iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else -1; continue
        if hdr is None: continue
        if r[0] != "":
            cur = (fname, r[0], r[1].strip()[:90])
            a = agg.setdefault(cur, [0, 0, 0])
            try:
                a[0] += int(r[ia]); a[1] += int(r[ism]); a[2] += int(r[iw] or 0) if iw >= 0 else 0
            except (ValueError, IndexError):
                pass
    tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot}")
    for (f, ln, src), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{v[0]/tot*100:5.1f}% inst {v[1]/ts*100:5.1f}% smp {v[2]/1e6:8.1f}M wf  {f}:{ln}: {src}")

if __name__ == "__main__":
    main()
