#!/usr/bin/env python
"""Top stall lines (CUDA source level) of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_src.py report.ncu-rep <kernel regex> [launch-skip] [topN]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""
agg = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        agg.append((fname, r))
k = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.lower().startswith("stall")]
tot = sum(float(r[k] or 0) for _, r in agg) or 1
print(f"kernel {kern} launch {skip}: {int(tot)} samples, {sum(float(r[ie] or 0) for _, r in agg):.0f} warp-instructions")
for f, r in sorted(agg, key=lambda x: -float(x[1][k] or 0))[:top]:
    stalls = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols if r[i] not in ("", "-")), reverse=True)[:2]
    st = " ".join(f"{n.replace('stall_', '')}:{int(v)}" for v, n in stalls if v > 0)
    print(f"{100 * float(r[k] or 0) / tot:6.2f}%  inst {float(r[ie] or 0):>10.0f}  {f}:{r[0]}: {r[1].strip()[:90]}   [{st}]")
