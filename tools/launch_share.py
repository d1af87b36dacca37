#!/usr/bin/env python
"""Per-kernel launch counts / summed duration / share from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python bench.py ...`).
usage: python tools/launch_share.py profiles/r1_launches_vNN.csv [out.txt]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r[ik])).replace("void ", "").strip()
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", "")) * scale
tot = sum(v[1] for v in agg.values())
out = [f"# {sys.argv[1]}: every kernel launched by the profiled command (warm-up, timed steps, stats pass, e2e pass, set-up solves);",
       "# per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"{k:44s} launches {v[0]:4d}  total {v[1]:9.3f} ms  mean {v[1] / v[0]:8.3f} ms  share {v[1] / tot * 100:5.1f} %")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
