#!/usr/bin/env python
"""Generates tests/golden/*.npz (c1_small, c1_odd) from the CPU oracle.

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

# name: seed, n, width, height, camera position, camera quaternion (w, x, y, z), fov, synth params
FIXTURES = {
    "c1_small": dict(seed=42, n=2000, w=160, h=120, pos=[0, 0, 5], quat=[1, 0, 0, 0], fov=45.0, synth={}),
    # odd image size (partial tiles on both edges), off-axis rotated camera inside the cloud (near-plane culls,
    # Gaussians straddling the image border), large anisotropic splats (AABB clamping, long tile runs)
    "c1_odd": dict(seed=7, n=1500, w=203, h=117, pos=[0.7, -0.4, 2.2], quat=[0.9689124, 0.0, 0.2474040, 0.0], fov=70.0,
                   synth=dict(log_scale_min=-4.0, log_scale_max=-0.7)),
}


def make(name):
    fx = FIXTURES[name]
    p = g.synth_params(**fx["synth"]) if fx["synth"] else None
    rec = g.synth_records(fx["seed"], fx["n"], p) if p is not None else g.synth_records(fx["seed"], fx["n"])
    vtx = g.activate_records(rec)
    u = g.uniforms_from_camera(fx["pos"], fx["quat"], fx["fov"], 0.1, 1000.0, fx["w"], fx["h"])
    cov = o.cov3d(vtx)
    f0 = o.render_frame(vtx, cov, u)
    o.set_exp_mode(1)
    f1 = o.render_frame(vtx, cov, u)
    o.set_exp_mode(0)
    out = Path(__file__).with_name(name + ".npz")
    np.savez_compressed(out, seed=fx["seed"], n=fx["n"], width=fx["w"], height=fx["h"],
                        uniforms=np.frombuffer(bytes(u), np.uint8), m=f1["m"], tiles=f1["tiles"], keys=f1["keys"],
                        vals=f1["vals"], ranges=f1["ranges"], rgba_exp_shared=f1["rgba"], rgba_exp_libm=f0["rgba"],
                        cov3d=cov)
    print("wrote", out, "M =", f1["m"])


def main():
    for name in (sys.argv[1:] or FIXTURES):
        make(name)


if __name__ == "__main__":
    main()
