"""Shared seeded inputs for the parity tests (product-side generator: gs_b200.synth_records)."""
import numpy as np

import gs_b200 as g


def quat_axis_angle(axis, deg):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    h = np.radians(deg) / 2
    return np.array([np.cos(h), *(a * np.sin(h))], np.float32)


def c1(n=10_000, seed=42):
    """BASELINE config 1: synthetic 10 k Gaussians, 640x480, camera (0,0,5), identity rotation, fov 45."""
    rec = g.synth_records(seed, n)
    vtx = g.activate_records(rec)
    u = g.uniforms_from_camera([0, 0, 5], [1, 0, 0, 0], 45.0, 0.1, 1000.0, 640, 480)
    return rec, vtx, u


CAMERAS = {
    # name: (pos, quat, fov, W, H)
    "c1": ([0, 0, 5], [1, 0, 0, 0], 45.0, 640, 480),
    "inside": ([0.3, -0.2, 0.5], quat_axis_angle([0, 1, 0], 30), 60.0, 640, 480),       # camera inside the cloud: huge splats, many culled
    "odd_size": ([1.0, 0.5, 6.0], quat_axis_angle([1, 0, 0], -8), 45.0, 333, 217),      # W,H not multiples of 16
    "tiny": ([0, 0, 5], [1, 0, 0, 0], 45.0, 16, 16),                                      # a single tile
    "wide": ([0, 0, 9], quat_axis_angle([0, 0, 1], 17), 70.0, 1024, 64),                 # few tile rows
    "away": ([0, 0, 5], quat_axis_angle([0, 1, 0], 180), 45.0, 320, 240),                # looking away: everything culled
}


def camera(name):
    pos, q, fov, w, h = CAMERAS[name]
    return g.uniforms_from_camera(pos, q, fov, 0.1, 1000.0, w, h)
