#!/usr/bin/env python
"""Which source lines the overhead opcodes (UMOV, MOV, FSEL, BRA, ...) of a profiled kernel come from.
usage: python tools/ncu_opcodes.py rep.ncu-rep [OP ...]"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict

def main():
    rep = sys.argv[1]
    ops = sys.argv[2:] or ["UMOV", "IMAD.MOV.U32", "FSEL", "BRA", "ISETP.NE.AND", "BSYNC.RECONVERGENT", "LDCU.128", "IMAD", "VIADD"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    hdr = None; cur = None; fname = "?"; per = defaultdict(Counter); tot = 0; allops = Counter()
    for r in csv.reader(io.StringIO(raw)):
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; ia = hdr.index("Instructions Executed"); continue
        if hdr is None or r[0] == "Function Name": continue
        if r[0] != "": cur = (fname, r[0]); continue
        toks = [x for x in r[3].split() if not x.startswith("@")]
        if not toks or toks[0] == "...": continue
        try: n = int(r[ia])
        except ValueError: continue
        op = toks[0].rstrip(";"); per[op][cur] += n; allops[op] += n; tot += n
    for op in ops:
        print(f"{op}: {allops[op] / tot * 100:.1f}% of executed warp instructions")
        for k, n in per[op].most_common(8): print(f"    {n / tot * 100:5.2f}%  {k[0]}:{k[1]}")

if __name__ == "__main__":
    main()
