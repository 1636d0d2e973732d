"""Per-rank stage times of a sharded garden frame with every rank on cuda:0 (Group) and of plain band renders."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python"))
import bench, gs_b200 as g
wl = bench.WORKLOADS["garden-standin"]
vtx = bench.make_scene(g, wl)
cams = bench.cameras(g, wl)
c = g.Context(0); c.set_tile_cull(True); c.upload(vtx)
for world in (2, 4, 8):
    tiles_y = (wl["h"] + 15) // 16; R = (tiles_y + world - 1) // world
    for r in range(world):
        rows = (min(tiles_y, r * R), min(tiles_y, (r + 1) * R))
        for _ in range(3):
            c.render(cams[0], g.FORMAT_BGRA8, rows=rows)
        s = c.stats()
        print(f"plain band world {world} rank {r} rows {rows}: M {s.num_instances} nv {s.num_visible} blend {s.render_ms:.3f} frame {s.frame_ms:.3f} consumed {s.blend_consumed} visits {s.blend_warp_visits}")
c.close()
for world in (2, 4):
    grp = g.Group([0] * world)
    grp.upload(vtx)
    for r in range(world):
        grp.context(r).set_tile_cull(True)
    for _ in range(4):
        grp.render(cams[0], g.FORMAT_BGRA8)
    for r in range(world):
        s = grp.context(r).stats()
        print(f"group world {world} rank {r}: M {s.num_instances} nv {s.num_visible} proj+xchg {s.preprocess_ms:.3f} blend {s.shard_blend_ms:.3f} wait {s.shard_wait_ms:.3f} frame {s.frame_ms:.3f} consumed {s.blend_consumed} visits {s.blend_warp_visits}")
    grp.close()
