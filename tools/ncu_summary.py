#!/usr/bin/env python
"""Summarise one round's profile artefacts (gpurun_out/<tag>_*) into profiles/<tag>_summary.md.
usage: tools/ncu_summary.py <tag>"""
import csv, io, json, subprocess, sys, collections
tag = sys.argv[1]
base = f"gpurun_out/{tag}"
bench = json.load(open(f"{base}_bench.json"))
raw = subprocess.run(["ncu", "-i", f"{base}_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def short(name):
    for k in ("k_project", "k_emit_cull", "k_emit_coarse", "k_emit", "k_sort_hist", "k_onesweep_pass", "k_tile_ranges", "k_frame_init", "k_blend2", "k_blend"):
        if k in name:
            return k + ("<u64>" if "IyE" in name or "unsigned long long *, " in name and k == "k_onesweep_pass" and "unsigned long long *, const unsigned int *, unsigned long long" in name else "")
    return name[:30]
cols = [("gpu__time_duration.sum", "time us", lambda v: f"{float(v)*1e3 if float(v)<10 else float(v):.1f}"),
        ("launch__grid_size", "grid", str), ("launch__registers_per_thread", "regs", str),
        ("dram__bytes_read.sum", "DRAM rd MB", lambda v: f"{float(v):.1f}"), ("dram__bytes_write.sum", "DRAM wr MB", lambda v: f"{float(v):.1f}"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % (ncu peak)", lambda v: f"{float(v):.1f}"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", lambda v: f"{float(v):.1f}"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", lambda v: f"{float(v):.1f}"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", lambda v: f"{float(v):.1f}"),
        ("smsp__inst_executed.sum", "warp-inst M", lambda v: f"{float(v)/1e6:.1f}")]
units = rows[1]
out = [f"# Profile summary {tag}", "",
       f"Workload: {bench['config']['workload']} ({bench['config']['note']}); N={bench['config']['n_gaussians']}, "
       f"{bench['config']['width']}x{bench['config']['height']}, visible {bench['config']['visible']:.0f}, instances M={bench['config']['instances_M']:.0f} "
       f"(AABB instances {bench['config'].get('instances_aabb', 0):.0f}, tile_cull={bench['config'].get('tile_cull')}), blend mode {bench['config']['blend_mode']}.",
       "", f"bench.py (not under a profiler): value **{bench['value']:.1f} frames/s** ({bench['ms_per_step']:.3f} ms/frame), e2e {bench['e2e']['value']:.1f} frames/s, "
       f"clocks {bench['clocks']}.", "", "## Per-kernel roofline from bench.py (cudaEvent stage timers, algorithmic bytes of DESIGN.md section 3)", "",
       "| kernel | ms/launch | launches/frame | algorithmic MB/launch | achieved GB/s | frac of measured 6580 GB/s | share of frame |", "|---|---|---|---|---|---|---|"]
for k, v in bench["kernels"].items():
    out.append(f"| {k} | {v['ms_per_launch']:.4f} | {v['launches_per_step']} | {v['alg_bytes_per_launch']/1e6:.1f} | {v['achieved_gbs']:.0f} | {v['achieved_gbs']/6580.3:.3f} | {v['share_of_step']:.3f} |")
out += ["", f"sort keys/s (M / t_sort): {bench['sort_keys_per_s']:.3e}; blend warp visits/s: {bench.get('blend_warp_visits_per_s') or 0:.3e} (one visit = 64 pixel x Gaussian pairs evaluated); lane utilisation {bench.get('blend_lane_utilisation')}", "",
        "## ncu --set full, one frame (cold cache, serialised; compare shares, not absolutes)", "",
        "| kernel | " + " | ".join(c[1] for c in cols) + " |", "|---|" + "---|" * len(cols)]
for r in data:
    vals = []
    for key, _, fmt in cols:
        v = r[ix[key]] if key in ix else ""
        if key == "gpu__time_duration.sum":
            u = units[ix[key]]
            f = float(v)
            v = f * 1e3 if u.startswith("ms") else (f if u.startswith("us") else f / 1e3)
            vals.append(f"{v:.1f}")
        else:
            try: vals.append(fmt(v))
            except Exception: vals.append(v)
    out.append(f"| {short(r[ix['Kernel Name']])} | " + " | ".join(vals) + " |")
# launch list shares
tot = collections.OrderedDict()
for r in csv.reader(open(f"{base}_launches.csv")):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        n = short(r[4]); tot[n] = tot.get(n, 0) + float(r[14])
s = sum(tot.values()) or 1
out += ["", "## ncu launch list (gpu__time_duration.sum over the captured launches): share of device time", "", "| kernel | total us | share |", "|---|---|---|"]
for k, v in tot.items():
    out.append(f"| {k} | {v/1e3:.1f} | {v/s:.3f} |")
# per-kernel DRAM traffic per launch (largest launch of each kernel) for bench.py's roofline.traffic
traffic = {}
for r in data:
    n = short(r[ix["Kernel Name"]])
    def mb(key):
        v, u = float(r[ix[key]]), units[ix[key]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1e6)
    b = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    traffic[n] = max(traffic.get(n, 0), b)
json.dump({"source": f"{tag}_full.ncu-rep (ncu --set full --clock-control none, one frame)", "dram_bytes_per_launch": traffic},
          open("profiles/ncu_traffic.json", "w"), indent=1)
open(f"profiles/{tag}_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
