"""Committed golden fixtures (tests/golden/make_golden.py).  CPU: the oracle still reproduces them.
GPU: the CUDA path reproduces them bit for bit."""
import sys
from pathlib import Path

import numpy as np
import pytest

GOLD_DIR = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLD_DIR))
from make_golden import FIXTURES  # noqa: E402  (the fixture specs: seeds, cameras, synth parameters)


def load(gs, name="c1_small"):
    z = np.load(GOLD_DIR / f"{name}.npz")
    fx = FIXTURES[name]
    rec = gs.synth_records(int(z["seed"]), int(z["n"]), gs.synth_params(**fx["synth"])) if fx["synth"] else gs.synth_records(int(z["seed"]), int(z["n"]))
    vtx = gs.activate_records(rec)
    u = gs.Uniforms.from_buffer_copy(z["uniforms"].tobytes())
    return z, vtx, u


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_oracle_reproduces_golden(gs, oracle, name):
    z, vtx, u = load(gs, name)
    fx = FIXTURES[name]
    assert bytes(gs.uniforms_from_camera(fx["pos"], fx["quat"], fx["fov"], 0.1, 1000.0, int(z["width"]), int(z["height"]))) == bytes(u)
    cov = oracle.cov3d(vtx)
    assert np.array_equal(cov, z["cov3d"])
    oracle.set_exp_mode(1)
    try:
        f = oracle.render_frame(vtx, cov, u)
    finally:
        oracle.set_exp_mode(0)
    assert f["m"] == int(z["m"])
    for k in ("tiles", "keys", "vals", "ranges"):
        assert np.array_equal(f[k], z[k]), k
    assert np.array_equal(f["rgba"], z["rgba_exp_shared"])
    f0 = oracle.render_frame(vtx, cov, u)
    assert np.abs(f0["rgba"] - z["rgba_exp_libm"]).max() <= 1e-6   # libm expf may differ by an ulp across glibc builds
    assert np.abs(z["rgba_exp_libm"] - z["rgba_exp_shared"]).max() <= 1e-4


@pytest.mark.gpu
def test_cuda_reproduces_golden(gs, ctx):
    z, vtx, u = load(gs)
    ctx.set_mode(gs.MODE_EXACT)
    ctx.set_debug(True)
    try:
        ctx.upload(vtx)
        img = ctx.render(u, gs.FORMAT_RGBA32F)
        assert ctx.stats().num_instances == int(z["m"])
        assert np.array_equal(ctx.download(gs.BUF_COV3D), z["cov3d"])
        assert np.array_equal(ctx.download(gs.BUF_TILES_OVERLAP), z["tiles"])
        assert np.array_equal(ctx.download(gs.BUF_KEYS_SORTED), z["keys"])
        assert np.array_equal(ctx.download(gs.BUF_VALS_SORTED), z["vals"])
        assert np.array_equal(ctx.download(gs.BUF_TILE_BOUNDARY), z["ranges"])
        assert np.array_equal(img, z["rgba_exp_shared"])
        assert np.abs(img - z["rgba_exp_libm"]).max() <= 1e-4
    finally:
        ctx.set_debug(False)
