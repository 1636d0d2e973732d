"""Small frames for compute-sanitizer: every kernel of the frame path at every gsb_set_tile_cull level, the stand-alone
sort, and a 3-rank sharded group on cuda:0 (routing into "peer" memory, mailbox flags, peer-store blend)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python")); sys.path.insert(0, str(ROOT / "tests"))
import gs_b200 as g
import scenes

_, vtx, _ = scenes.c1(n=4000)
c = g.Context(0)
c.upload(vtx)
ref = None
for level in (0, 1, 2):
    c.set_tile_cull(level)
    for timers in (True, False):
        c.set_timers(timers)
        for cam in ("c1", "odd_size"):
            u = scenes.camera(cam)
            img = c.render(u, g.FORMAT_RGBA32F)
            img8 = c.render(u, g.FORMAT_BGRA8)
            if cam == "c1":
                if ref is None:
                    ref = img
                assert np.array_equal(img, ref), (level, timers)
c.close()
grp = g.Group([0, 0, 0])
grp.upload(vtx)
for level in (1, 2):
    for r in range(3):
        grp.context(r).set_tile_cull(level)
    for _ in range(2):
        assert np.array_equal(grp.render(scenes.camera("c1"), g.FORMAT_RGBA32F), ref), level
grp.close()
print("sanitize target ok")
