"""Full-size checks on BASELINE.json's headline configuration (synthetic stand-in: 5.8 M Gaussians, 3200x1400):
size-independent properties of the CUDA path + a band of tile rows against the oracle (the oracle cannot do the
whole frame in seconds, a band it can)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, W, H = 5_800_000, 3200, 1400


@pytest.fixture(scope="module")
def garden(gs):
    p = gs.synth_params(center=(0, 0, 0), half_extent=(10.0, 4.0, 10.0), log_scale_min=math.log(0.003), log_scale_max=math.log(0.05))
    vtx = np.empty((N, 60), np.float32)
    for off in range(0, N, 1 << 20):
        cnt = min(1 << 20, N - off)
        vtx[off:off + cnt] = gs.activate_records(gs.synth_records(3, cnt, p, first=off))
    u = gs.uniforms_from_camera([0, 0, 14], [1, 0, 0, 0], 45.0, 0.1, 1000.0, W, H)
    c = gs.Context(0)
    c.upload(vtx)
    yield c, vtx, u
    c.close()


def test_fullsize_properties(gs, garden):
    c, vtx, u = garden
    c.set_mode(gs.MODE_EXACT)
    c.set_debug(True)
    try:
        img = c.render(u, gs.FORMAT_RGBA32F)
        st = c.stats()
        keys, vals = c.download(gs.BUF_KEYS_SORTED), c.download(gs.BUF_VALS_SORTED)
        tiles, ranges = c.download(gs.BUF_TILES_OVERLAP), c.download(gs.BUF_TILE_BOUNDARY)
    finally:
        c.set_debug(False)
    M = int(st.num_instances)
    assert M == int(tiles.astype(np.uint64).sum()) == keys.size == vals.size and M > 20_000_000
    assert st.num_visible == int((tiles > 0).sum())
    assert np.all(keys[1:] >= keys[:-1])                                   # sortedness of the full 64-bit keys
    # payload multiset: every Gaussian appears exactly tiles_overlap[i] times (a permutation of the emission)
    assert np.array_equal(np.bincount(vals, minlength=N).astype(np.uint32), tiles)
    # stability: equal keys keep Gaussian-index order
    same = keys[1:] == keys[:-1]
    assert np.all(vals[1:][same] > vals[:-1][same])
    # tile ranges partition the sorted list
    tid = (keys >> np.uint64(32)).astype(np.int64)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    start, end = np.searchsorted(tid, np.arange(T), "left"), np.searchsorted(tid, np.arange(T), "right")
    nonempty = start < end
    assert np.array_equal(ranges[nonempty], np.stack([start, end], -1)[nonempty]) and np.all(ranges[~nonempty] == 0)
    assert np.isfinite(img).all() and np.all(img[..., 3] == 1.0) and img[..., :3].max() < 50
    # idempotence
    assert np.array_equal(img, c.render(u, gs.FORMAT_RGBA32F))
    # tile-row bands (the multi-GPU shards) reproduce the frame
    tiles_y = (H + 15) // 16
    bands = [c.render(u, gs.FORMAT_RGBA32F, rows=(tiles_y * k // 4, tiles_y * (k + 1) // 4)) for k in range(4)]
    assert np.array_equal(np.concatenate(bands, axis=0), img)
    # exact instance culling: same pixels, fewer instances
    c.set_tile_cull(True)
    try:
        img_c = c.render(u, gs.FORMAT_RGBA32F)
        st_c = c.stats()
    finally:
        c.set_tile_cull(False)
    assert np.array_equal(img_c, img)
    assert st_c.num_instances_aabb == M and st_c.num_instances < 0.8 * M
    # FAST mode (FMA + ex2.approx) stays within the north_star tolerance except at the shader's own step functions
    c.set_mode(gs.MODE_FAST)
    try:
        err = np.abs(c.render(u, gs.FORMAT_RGBA32F) - img)
    finally:
        c.set_mode(gs.MODE_EXACT)
    assert (err > 1e-4).any(axis=-1).mean() < 1e-5, (err.max(), (err > 1e-4).sum())
    # packed swapchain format == quantised float image
    bgra = c.render(u, gs.FORMAT_BGRA8)
    q = np.rint(np.clip(img[..., [2, 1, 0, 3]], 0, 1) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(bgra, q)


def test_fullsize_band_matches_oracle_bit_exact(gs, oracle, garden):
    c, vtx, u = garden
    rows = (43, 45)  # two tile rows through the densest part of the frame
    c.set_mode(gs.MODE_EXACT)
    band = c.render(u, gs.FORMAT_RGBA32F, rows=rows)
    m_gpu = c.stats().num_instances
    oracle.set_exp_mode(1)
    try:
        ref = oracle.render_frame(vtx, oracle.cov3d(vtx), u, rows=rows)
    finally:
        oracle.set_exp_mode(0)
    assert m_gpu == ref["m"]
    assert np.array_equal(band, ref["rgba"][rows[0] * 16:rows[1] * 16])
