#!/usr/bin/env python
"""Generates tests/golden/c1_small.npz from the CPU oracle.

The reference (shg8/3DGS.cpp) has no golden vectors and cannot be run here (PARITY UNPINNED), so these
fixtures pin the ORACLE itself against accidental change and give the GPU tests a committed target:
  inputs : gs_b200.synth_records(seed=42, n=2000) (config-1 distribution), camera (0,0,5) identity, fov 45, 160x120
  outputs: float image with the shared-definition exp (bit-exact target of the CUDA path) and with libm exp,
           M, per-Gaussian tile counts, sorted keys/payloads, tile ranges.
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "3dgs.cpp_b200" / "python"))
sys.path.insert(0, str(ROOT / "oracle"))
import gs_b200 as g  # noqa: E402
import oracle as o  # noqa: E402

SEED, N, W, H = 42, 2000, 160, 120


def main():
    vtx = g.activate_records(g.synth_records(SEED, N))
    u = g.uniforms_from_camera([0, 0, 5], [1, 0, 0, 0], 45.0, 0.1, 1000.0, W, H)
    cov = o.cov3d(vtx)
    f0 = o.render_frame(vtx, cov, u)
    o.set_exp_mode(1)
    f1 = o.render_frame(vtx, cov, u)
    o.set_exp_mode(0)
    np.savez_compressed(Path(__file__).with_name("c1_small.npz"), seed=SEED, n=N, width=W, height=H,
                        uniforms=np.frombuffer(bytes(u), np.uint8), m=f1["m"], tiles=f1["tiles"], keys=f1["keys"],
                        vals=f1["vals"], ranges=f1["ranges"], rgba_exp_shared=f1["rgba"], rgba_exp_libm=f0["rgba"],
                        cov3d=cov)
    print("wrote", Path(__file__).with_name("c1_small.npz"), "M =", f1["m"])


if __name__ == "__main__":
    main()
