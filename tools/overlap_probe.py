#!/usr/bin/env python
"""Probe for DESIGN.md section 9 item 1: do two frames in flight raise throughput?
Two independent contexts (each with its own copy of the bench scene and its own stream) render alternate frames of
the bench orbit; compared with one context rendering all of them back to back.  Not part of bench.py: a measurement
to decide whether a shared-scene two-frame arena is worth building.  usage: python tools/overlap_probe.py [frames]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python"))
import bench  # noqa: E402
import gs_b200 as g  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    wl = bench.WORKLOADS["garden-standin"]
    vtx = bench.make_scene(g, wl)
    cams = bench.cameras(g, wl)
    W, H = wl["w"], wl["h"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctxs, streams, fbs = [], [], []
    for _ in range(2):
        c = g.Context(0)
        c.set_mode(g.MODE_EXACT)
        c.set_tile_cull(True)
        c.set_timers(False)
        c.upload(vtx)
        s = torch.cuda.Stream(device=dev)
        fb = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
        c.render_into(cams[0], fb.data_ptr(), g.FORMAT_BGRA8, stream=s, sync=True)
        c.reserve(int(c.stats().num_instances * 1.3) + 65536)
        for i in range(4):
            c.render_into(cams[i % len(cams)], fb.data_ptr(), g.FORMAT_BGRA8, stream=s, sync=True)
        ctxs.append(c); streams.append(s); fbs.append(fb)

    def run(nctx):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(frames):
            k = i % nctx
            ctxs[k].render_into(cams[i % len(cams)], fbs[k].data_ptr(), g.FORMAT_BGRA8, stream=streams[k], sync=False)
        torch.cuda.synchronize()
        return frames / (time.perf_counter() - t0)

    run(1); run(2)
    one = max(run(1) for _ in range(3))
    two = max(run(2) for _ in range(3))
    for c in ctxs:
        c.stats()
    print(json.dumps({"frames": frames, "one_context_fps": one, "two_contexts_fps": two, "ratio": two / one}))


if __name__ == "__main__":
    main()
