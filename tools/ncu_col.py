#!/usr/bin/env python
"""Top CUDA source lines of one kernel by an arbitrary source-page column.
usage: tools/ncu_col.py report.ncu-rep <kernel regex> <launch-skip> "<column name>" [topN]"""
import csv, io, subprocess, sys
rep, kern, skip, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, agg, fname = None, [], ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-": agg.append((fname, r))
k = hdr.index(col)
def f(x):
    try: return float(x)
    except: return 0.0
tot = sum(f(r[k]) for _, r in agg) or 1
print(f"{col}: total {tot:.0f}")
for fn, r in sorted(agg, key=lambda x: -f(x[1][k]))[:top]:
    print(f"{100*f(r[k])/tot:6.2f}% {f(r[k]):>12.0f}  {fn}:{r[0]}: {r[1].strip()[:100]}")
