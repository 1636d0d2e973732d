"""Scene load times: PLY read + activation + upload through the C++ host (headless viewer JSON) for the
garden stand-in, and gsb_scene_upload alone (pageable and page-locked source) for the 50 M-Gaussian scene."""
import json, subprocess, sys, time, tempfile, os
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python"))
import bench, gs_b200 as g
import torch

out = {}
wl = bench.WORKLOADS["garden-standin"]
p = g.synth_params(center=(0, 0, 0), half_extent=wl["half"], log_scale_min=np.log(wl["ls"][0]), log_scale_max=np.log(wl["ls"][1]))
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    ply = Path(d) / "garden_standin.ply"
    rec = np.concatenate([g.synth_records(wl["seed"], min(1 << 20, wl["n"] - off), p, first=off) for off in range(0, wl["n"], 1 << 20)])
    g.write_ply(ply, rec)
    del rec
    exe = ROOT / "3dgs.cpp_b200" / "gs_viewer_headless"
    for _ in range(2):  # second run: file in the page cache
        r = subprocess.run([str(exe), "-w", str(wl["w"]), "-h", str(wl["h"]), "--frames", "20", "--camera", "0,0,14", "--cull", "2", str(ply)],
                           capture_output=True, text=True, timeout=600)
        info = json.loads(r.stdout.strip().splitlines()[-1])
    out["garden_ply_bytes"] = ply.stat().st_size
    out["garden_cli"] = {k: info[k] for k in ("gaussians", "load_ms", "read_activate_ms", "upload_ms", "fps_wall", "frame_ms")}
wl5 = bench.WORKLOADS["synthetic-50m"]
t0 = time.perf_counter()
vtx = bench.make_scene(g, wl5)
out["c5_synth_activate_s"] = time.perf_counter() - t0
c = g.Context(0)
t0 = time.perf_counter(); c.upload(vtx); out["c5_upload_pageable_s"] = time.perf_counter() - t0
pinned = torch.from_numpy(vtx).pin_memory()
t0 = time.perf_counter(); c.upload(pinned.numpy()); out["c5_upload_pinned_s"] = time.perf_counter() - t0
out["c5_bytes"] = int(vtx.nbytes)
c.close()
print(json.dumps(out))
