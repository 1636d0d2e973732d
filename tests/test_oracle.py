"""The oracle (oracle/gs_oracle.c) checked against independent numpy restatements and its own invariants.
The reference has no tests / golden vectors for this path (PARITY UNPINNED), so the oracle
is pinned by (a) a second, independent float64 numpy implementation of each stage, (b) the shaders' own
debug asserts (preprocess_sort.comp:41,60, preprocess.comp:173) and (c) committed golden fixtures."""
import math

import numpy as np
import pytest

import scenes


def test_exp_shared_accuracy_and_monotone(oracle):
    xs = np.linspace(-5.55, 0.0, 20001, dtype=np.float32)
    got = np.array([oracle.exp_shared(float(x)) for x in xs], np.float32)
    ref = np.exp(xs.astype(np.float64))
    ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert ulp.max() <= 2.0, ulp.max()
    assert np.all(np.diff(got) >= 0)
    assert oracle.exp_shared(0.0) == 1.0


def test_cov3d_matches_float64_rs2rt(gs, oracle):
    _, vtx, _ = scenes.c1(n=2000)
    cov = oracle.cov3d(vtx).astype(np.float64)
    s = vtx[:, 4:7].astype(np.float64)
    w, x, y, z = [vtx[:, 8 + k].astype(np.float64) for k in range(4)]
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    S = R @ (s[:, :, None] ** 2 * np.swapaxes(R, 1, 2))      # R diag(s^2) R^T
    ref = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)
    assert np.allclose(cov, ref, rtol=2e-5, atol=1e-9)


def numpy_preprocess(vtx, cov, u):
    """Independent float64 restatement of preprocess.comp for cross-checking."""
    V = np.array(u.view_mat, np.float64).reshape(4, 4).T
    P = np.array(u.proj_mat, np.float64).reshape(4, 4).T
    W, H = u.width, u.height
    p = np.concatenate([vtx[:, :3].astype(np.float64), np.ones((len(vtx), 1))], 1)
    ph = p @ P.T
    pv = p @ V.T
    ndc = ph[:, :3] / ph[:, 3:4]
    z = pv[:, 2]
    vis = z > 0.2
    limx, limy = 1.3 * u.tan_fovx, 1.3 * u.tan_fovy
    tx = np.clip(pv[:, 0] / z, -limx, limx) * z
    ty = np.clip(pv[:, 1] / z, -limy, limy) * z
    fx, fy = W / (2 * u.tan_fovx), H / (2 * u.tan_fovy)
    J = np.zeros((len(vtx), 2, 3))
    J[:, 0, 0], J[:, 0, 2] = fx / z, -fx * tx / z**2
    J[:, 1, 1], J[:, 1, 2] = fy / z, -fy * ty / z**2
    c = cov.astype(np.float64)
    Sig = np.stack([np.stack([c[:, 0], c[:, 1], c[:, 2]], -1), np.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                    np.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], -2)
    T = J @ V[:3, :3]
    c2 = T @ Sig @ np.swapaxes(T, 1, 2)
    a, b, d = c2[:, 0, 0] + 0.3, c2[:, 0, 1], c2[:, 1, 1] + 0.3
    det = a * d - b * b
    vis &= det > 0
    mid = 0.5 * (a + d)
    lam = mid + np.sqrt(np.maximum(0.1, mid * mid - det))
    rad = np.ceil(3 * np.sqrt(lam))
    uvx, uvy = ((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5
    tiles_x, tiles_y = (W + 15) // 16, (H + 15) // 16
    with np.errstate(invalid="ignore"):
        bx0 = np.clip(np.trunc((uvx - rad) / 16), 0, tiles_x)
        by0 = np.clip(np.trunc((uvy - rad) / 16), 0, tiles_y)
        bx1 = np.clip(np.trunc((uvx + rad + 15) / 16), 0, tiles_x)
        by1 = np.clip(np.trunc((uvy + rad + 15) / 16), 0, tiles_y)
    nt = (bx1 - bx0) * (by1 - by0)
    vis &= nt > 0
    return dict(vis=vis, rad=rad, conic=np.stack([d / det, -b / det, a / det], -1), uv=np.stack([uvx, uvy], -1), depth=z,
                aabb=np.stack([bx0, by0, bx1, by1], -1), nt=nt)


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size", "wide"])
def test_preprocess_matches_independent_numpy(gs, oracle, cam):
    _, vtx, _ = scenes.c1(n=4000)
    u = scenes.camera(cam)
    cov = oracle.cov3d(vtx)
    attr, tiles = oracle.preprocess(vtx, cov, u)
    ref = numpy_preprocess(vtx, cov, u)
    vis = tiles > 0
    # step functions (cull, ceil, trunc) may flip on fp32-vs-fp64 rounding for a handful of Gaussians
    assert (vis != ref["vis"]).mean() < 2e-3
    both = vis & ref["vis"]
    assert both.sum() > 100
    assert np.allclose(attr["depth"][both], ref["depth"][both], rtol=1e-5)
    assert np.allclose(attr["uv"][both], ref["uv"][both], rtol=1e-4, atol=2e-3)
    assert np.allclose(attr["conic_opacity"][both, :3], ref["conic"][both], rtol=2e-3, atol=1e-6)
    assert (attr["color_radii"][both, 3] != ref["rad"][both]).mean() < 5e-3
    assert (attr["aabb"][both] != ref["aabb"][both]).any(axis=1).mean() < 1e-2
    assert np.array_equal(attr["conic_opacity"][vis, 3], vtx[vis, 7])      # opacity passthrough
    assert np.all(attr["magic"][vis] == 0x4D415449) and np.all(attr["magic"][~vis] == 0)


def test_sh_colour_matches_numpy_and_only_red_is_clamped(gs, oracle):
    _, vtx, u = scenes.c1(n=3000)
    vtx = vtx.copy()
    vtx[:, 12:15] -= 2.0  # push DC down so that some channels go negative
    attr, tiles = oracle.preprocess(vtx, oracle.cov3d(vtx), u)
    vis = tiles > 0
    d = vtx[:, :3].astype(np.float64) - np.array(u.camera_position[:3], np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x, y, z = d.T
    sh = vtx[:, 12:].astype(np.float64).reshape(-1, 16, 3)
    C1, C2, C3 = 0.4886025119029199, [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396], \
        [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277, -0.5900435899266435]
    basis = [0.28209479177387814 * np.ones_like(x), -C1 * y, C1 * z, -C1 * x, C2[0] * x * y, C2[1] * y * z,
             C2[2] * (2 * z * z - x * x - y * y), C2[3] * x * z, C2[4] * (x * x - y * y), C3[0] * y * (3 * x * x - y * y),
             C3[1] * x * y * z, C3[2] * y * (4 * z * z - x * x - y * y), C3[3] * z * (2 * z * z - 3 * x * x - 3 * y * y),
             C3[4] * x * (4 * z * z - x * x - y * y), C3[5] * z * (x * x - y * y), C3[6] * x * (x * x - 3 * y * y)]
    col = sum(b[:, None] * sh[:, k, :] for k, b in enumerate(basis)) + 0.5
    col[:, 0] = np.maximum(col[:, 0], 0.0)
    got = attr["color_radii"][vis, :3]
    assert np.allclose(got, col[vis], rtol=1e-4, atol=2e-5)
    assert (got[:, 0] >= 0).all() and (got[:, 1] < 0).any() and (got[:, 2] < 0).any()   # preprocess.comp:102-104


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size", "tiny"])
def test_frame_invariants(gs, oracle, cam):
    _, vtx, _ = scenes.c1(n=3000)
    u = scenes.camera(cam)
    f = oracle.render_frame(vtx, oracle.cov3d(vtx), u)
    tiles, scan = f["tiles"], f["scan"]
    assert np.array_equal(scan, np.cumsum(tiles, dtype=np.uint64).astype(np.uint32))   # inclusive scan
    assert f["m"] == int(tiles.sum())
    # emission order: Gaussian-major, x outer, y inner (preprocess_sort.comp:47-48)
    ku, vu = f["keys_unsorted"], f["vals_unsorted"]
    assert np.all(np.diff(vu.astype(np.int64)) >= 0)
    tx = f["tiles_x"]
    for i in np.nonzero(tiles)[0][:50]:
        a = f["attr"][i]["aabb"]
        off = 0 if i == 0 else int(scan[i - 1])
        exp = [(x + y * tx) for x in range(a[0], a[2]) for y in range(a[1], a[3])]
        assert list(ku[off:off + tiles[i]] >> np.uint64(32)) == exp
        assert np.all((ku[off:off + tiles[i]] & np.uint64(0xFFFFFFFF)) == f["attr"][i]["depth"].view(np.uint32))
    # sort == numpy stable sort on the full 64-bit key (ties keep Gaussian order)
    order = np.argsort(ku, kind="stable")
    assert np.array_equal(f["keys"], ku[order]) and np.array_equal(f["vals"], vu[order])
    # ranges == searchsorted on the tile ids; untouched tiles stay (0, 0)
    tid = (f["keys"] >> np.uint64(32)).astype(np.int64)
    T = f["tiles_x"] * f["tiles_y"]
    start, end = np.searchsorted(tid, np.arange(T), "left"), np.searchsorted(tid, np.arange(T), "right")
    empty = start == end
    assert np.array_equal(f["ranges"][~empty], np.stack([start, end], -1)[~empty])
    assert np.all(f["ranges"][empty] == 0)
    assert np.all(f["rgba"][..., 3] == 1.0) and np.isfinite(f["rgba"]).all()


def numpy_blend_pixel(f, px, py):
    t = (px // 16) + (py // 16) * f["tiles_x"]
    s, e = f["ranges"][t]
    T, c = np.float32(1.0), np.zeros(3, np.float32)
    for i in range(s, e):
        a = f["attr"][f["vals"][i]]
        dx, dy = np.float32(a["uv"][0] - np.float32(px)), np.float32(a["uv"][1] - np.float32(py))
        co = a["conic_opacity"]
        power = np.float32(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy
        if power > 0:
            continue
        alpha = min(np.float32(0.99), np.float32(co[3] * np.float32(math.exp(power))))
        if alpha < np.float32(1.0 / 255.0):
            continue
        test_T = np.float32(T * (np.float32(1) - alpha))
        if test_T < np.float32(0.0001):
            break
        c = c + a["color_radii"][:3] * alpha * T
        T = test_T
    return c


def test_blend_matches_numpy_on_sampled_pixels(gs, oracle):
    _, vtx, u = scenes.c1(n=3000)
    f = oracle.render_frame(vtx, oracle.cov3d(vtx), u)
    rng = np.random.default_rng(3)
    for _ in range(60):
        px, py = int(rng.integers(0, u.width)), int(rng.integers(0, u.height))
        assert np.allclose(f["rgba"][py, px, :3], numpy_blend_pixel(f, px, py), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size"])
def test_exp_modes_agree_within_tolerance(gs, oracle, cam):
    """libm exp (the plain restatement) vs the shared-definition exp the CUDA kernel reproduces bit for bit."""
    _, vtx, _ = scenes.c1()
    u = scenes.camera(cam)
    cov = oracle.cov3d(vtx)
    a = oracle.render_frame(vtx, cov, u)["rgba"]
    oracle.set_exp_mode(1)
    try:
        b = oracle.render_frame(vtx, cov, u)["rgba"]
    finally:
        oracle.set_exp_mode(0)
    assert np.abs(a - b).max() <= 1e-4


def test_band_rows_reproduce_the_full_frame(gs, oracle):
    _, vtx, u = scenes.c1(n=3000)
    cov = oracle.cov3d(vtx)
    full = oracle.render_frame(vtx, cov, u)
    tiles_y = full["tiles_y"]
    for rb, re in [(0, 7), (7, 19), (19, tiles_y)]:
        band = oracle.render_frame(vtx, cov, u, rows=(rb, re))
        assert np.array_equal(band["rgba"][rb * 16:min(u.height, re * 16)], full["rgba"][rb * 16:min(u.height, re * 16)])
        assert band["m"] <= full["m"]


def test_pack_unorm8(oracle):
    x = np.array([[-1.0, 0.0, 0.5, 1.0], [2.0, 0.00196, 0.998, np.nan]], np.float32).reshape(2, 1, 4)
    rgba = oracle.pack_unorm8(x)
    assert rgba.reshape(2, 4).tolist() == [[0, 0, 128, 255], [255, 0, 254, 0]]   # 127.5 -> 128 (RNE), NaN -> 0
    bgra = oracle.pack_unorm8(x, bgra=True)
    assert bgra.reshape(2, 4).tolist() == [[128, 0, 0, 255], [254, 0, 255, 0]]


def test_oracle_scene_generator_equals_the_products(gs, oracle):
    """bench.py's reference arm builds its scene with the oracle's restatement of the SURVEY 8d generator so that it
    loads no product library; both generators (and both activations) must agree bit for bit."""
    import math
    assert np.array_equal(gs.synth_records(42, 70_000), oracle.synth_records(42, 70_000))
    kw = dict(center=(0, 0, 0), half_extent=(10.0, 4.0, 10.0), log_scale_min=math.log(0.003), log_scale_max=math.log(0.05))
    a = gs.synth_records(3, 100_000, gs.synth_params(**kw), first=5_000_000)
    b = oracle.synth_records(3, 100_000, oracle.synth_params(**kw), first=5_000_000)
    assert np.array_equal(a, b)
    assert np.array_equal(gs.activate_records(a), oracle.load_records(b))


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_parallel_sort_scan_are_the_serial_ones(oracle, threads):
    """gso_sort / gso_scan_inclusive run OpenMP-parallel above 64 Ki elements: still the stable LSD sort and the exact
    inclusive scan (uint32 wraparound) of the reference."""
    rng = np.random.default_rng(5)
    m = 300_007
    keys = (rng.integers(0, 700, m).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 50, m).astype(np.uint64) * np.uint64(0x01010101)
    vals = np.arange(m, dtype=np.uint32)
    oracle.set_num_threads(threads)
    try:
        k, v = oracle.sort_pairs(keys, vals)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
        tiles = rng.integers(0, 2**31, 200_000).astype(np.uint32)  # sums wrap around uint32
        scan = np.empty_like(tiles)
        oracle.lib.gso_scan_inclusive.restype = oracle.C.c_uint64
        oracle.lib.gso_scan_inclusive(tiles.ctypes.data_as(oracle.C.c_void_p), oracle.C.c_uint64(tiles.size), scan.ctypes.data_as(oracle.C.c_void_p))
        assert np.array_equal(scan, np.cumsum(tiles.astype(np.uint64)).astype(np.uint32))
    finally:
        oracle.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))


@pytest.mark.parametrize("cam,shift", [("c1", 2), ("odd_size", 2), ("inside", 2), ("wide", 1)])
def test_coarse_bin_model_reproduces_every_tile_list(gs, oracle, cam, shift):
    """The claim behind gsb_set_tile_cull level 2, checked on the oracle's own buffers (no GPU): emit one entry per
    (Gaussian, block of 2^shift x 2^shift tiles) in depth order with the mask of the block's tiles inside the Gaussian's
    tile AABB, sort the entries stably by block id, and let every tile keep the entries of its block whose mask has its
    bit -- that is exactly the tile's own (depth, index)-ordered list of the reference's 64-bit key sort."""
    _, vtx, _ = scenes.c1(n=3000)
    u = scenes.camera(cam)
    f = oracle.render_frame(vtx, oracle.cov3d(vtx), u)
    tiles_x, tiles_y = f["tiles_x"], f["tiles_y"]
    bins_x = (tiles_x + (1 << shift) - 1) >> shift
    aabb = f["attr"]["aabb"].astype(np.int64)
    depth_bits = f["attr"]["depth"].view(np.uint32).astype(np.uint64)
    live = np.nonzero(f["tiles"])[0]
    order = live[np.argsort(depth_bits[live], kind="stable")]  # the Gaussian-level sort: (depth bits, index)
    keys, vals = [], []
    for i in order:  # k_emit_coarse: block id | mask << 16, x outer / y inner
        x0, y0, x1, y1 = aabb[i]
        for bx in range(x0 >> shift, ((x1 - 1) >> shift) + 1):
            for by in range(y0 >> shift, ((y1 - 1) >> shift) + 1):
                m = 0
                for ly in range(1 << shift):
                    for lx in range(1 << shift):
                        tx, ty = (bx << shift) + lx, (by << shift) + ly
                        if x0 <= tx < x1 and y0 <= ty < y1:
                            m |= 1 << ((ly << shift) | lx)
                assert m != 0
                keys.append((bx + by * bins_x) | (m << 16))
                vals.append(i)
    keys, vals = np.array(keys, np.uint64), np.array(vals, np.int64)
    perm = np.argsort(keys & np.uint64(0xFFFF), kind="stable")  # the instance sort looks at the low 16 bits only
    keys, vals = keys[perm], vals[perm]
    blocks = (keys & np.uint64(0xFFFF)).astype(np.int64)
    ref_tile = (f["keys"] >> np.uint64(32)).astype(np.int64)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            t = tx + ty * tiles_x
            want = f["vals"][ref_tile == t].astype(np.int64)  # the reference's list of this tile, in sorted order
            sel = blocks == (tx >> shift) + (ty >> shift) * bins_x
            bit = 16 + (((ty & ((1 << shift) - 1)) << shift) | (tx & ((1 << shift) - 1)))
            got = vals[sel][((keys[sel] >> np.uint64(bit)) & np.uint64(1)).astype(bool)]
            assert np.array_equal(got, want), (cam, tx, ty)
    # and the entry count the GPU test pins: (Gaussian, block) pairs of the AABBs
    nb = ((((aabb[live, 2] - 1) >> shift) - (aabb[live, 0] >> shift) + 1) * (((aabb[live, 3] - 1) >> shift) - (aabb[live, 1] >> shift) + 1)).sum()
    assert len(keys) == nb
