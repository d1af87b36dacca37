#!/usr/bin/env python
"""Executed-instruction mix of a profiled kernel by opcode (top N) and by coarse class.
usage: python tools/ncu_opmix.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
from collections import Counter

CLASSES = [("fp64", ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")), ("smem", ("LDS", "STS")), ("gmem", ("LDG", "STG", "LD.", "ST.", "ATOM", "RED")),
           ("const", ("LDC", "ULDC", "LDCU")), ("int/addr", ("IMAD", "IADD", "LEA", "SHF", "LOP3", "VIADD", "ISETP", "IABS", "SHL", "SHR", "PRMT", "I2F", "F2I", "I2I", "UIADD", "ULEA", "USHF", "ULOP", "UISETP", "UIMAD", "VIMNMX")),
           ("mov/sel", ("MOV", "UMOV", "SEL", "FSEL", "USEL", "R2UR", "S2R", "CS2R", "S2UR", "P2R", "R2P", "PLOP3", "UPLOP3", "UP2UR")), ("branch/sync", ("BRA", "BSSY", "BSYNC", "WARPSYNC", "BAR", "EXIT", "CALL", "RET", "NOP", "BRX", "JMP", "YIELD", "DEPBAR", "BREAK", "BMOV", "ELECT", "VOTE", "SHFL", "MATCH")),
           ("mufu/f32", ("MUFU", "FFMA", "FMUL", "FADD", "FSETP", "F2F", "FCHK", "FMNMX", "I2FP", "F2FP"))]

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    hdr = None; ops = Counter(); tot = 0
    for r in csv.reader(io.StringIO(raw)):
        if not r: continue
        if "Instructions Executed" in r and "Source" in r:
            hdr = r; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); continue
        if hdr is None or len(r) <= max(ia, isrc): continue
        toks = [x for x in r[isrc].split() if not x.startswith("@")]
        if not toks: continue
        try: n = int(r[ia])
        except ValueError: continue
        op = toks[0].rstrip(";"); ops[op] += n; tot += n
    print(f"total executed warp instructions {tot}")
    cls = Counter()
    for op, n in ops.items():
        for name, pre in CLASSES:
            if any(op.startswith(p) for p in pre): cls[name] += n; break
        else: cls["other"] += n
    for name, n in cls.most_common(): print(f"  {name:12s} {n / tot * 100:5.1f}%")
    for op, n in ops.most_common(top): print(f"    {op:24s} {n / tot * 100:5.2f}%")

if __name__ == "__main__":
    main()
