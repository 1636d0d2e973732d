"""Image equality of gsb_set_tile_cull levels 0 / 1 / 2 on the bench scene + stage times of each."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python"))
import bench, gs_b200 as g
name = sys.argv[1] if len(sys.argv) > 1 else "garden-standin"
wl = bench.WORKLOADS[name]
vtx = bench.make_scene(g, wl)
cams = bench.cameras(g, wl)
c = g.Context(0); c.upload(vtx)
ref = None
for level in (0, 1, 2):
    c.set_tile_cull(level)
    for cam in (cams[0], cams[5]):
        c.render(cam, g.FORMAT_RGBA32F)
    img = c.render(cams[0], g.FORMAT_RGBA32F)
    s = c.stats()
    if ref is None:
        ref = img
    print(f"{name} level {level}: equal {np.array_equal(img, ref)} M {s.num_instances} aabb {s.num_instances_aabb} emit {s.preprocess_sort_ms:.3f} sort_tile {s.sort_tile_ms:.3f} blend {s.render_ms:.3f} frame {s.frame_ms:.3f} consumed {s.blend_consumed} visits {s.blend_warp_visits}")
c.close()
