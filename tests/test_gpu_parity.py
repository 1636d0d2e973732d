"""GPU parity tests: the CUDA path (through the C ABI, libgsb200.so) against the CPU oracle on the
same seeded inputs.  Integer / index stages are bit-exact; the float image is bit-exact against the
oracle's shared-definition exp (mode 1) and within 1e-4 L-inf of the libm-exp oracle (mode 0), the
tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

TOL = 1e-4  # per-channel L-inf, north_star


def oracle_frame(o, vtx, u, exp_mode, rows=None):
    o.set_exp_mode(exp_mode)
    try:
        return o.render_frame(vtx, o.cov3d(vtx), u, rows)
    finally:
        o.set_exp_mode(0)


def test_cov3d_exact(gs, oracle, ctx):
    _, vtx, _ = scenes.c1()
    ctx.upload(vtx)
    assert np.array_equal(ctx.download(gs.BUF_COV3D), oracle.cov3d(vtx))


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size", "tiny", "wide", "away"])
def test_frame_intermediates_and_image_exact(gs, oracle, ctx, cam):
    _, vtx, _ = scenes.c1()
    u = scenes.camera(cam)
    ctx.set_mode(gs.MODE_EXACT)
    ctx.set_debug(True)
    ctx.upload(vtx)
    img = ctx.render(u, gs.FORMAT_RGBA32F)
    st = ctx.stats()
    ref = oracle_frame(oracle, vtx, u, 1)
    try:
        assert st.num_instances == ref["m"]
        assert st.num_visible == int((ref["tiles"] > 0).sum())
        # P0: preprocess outputs
        assert np.array_equal(ctx.download(gs.BUF_TILES_OVERLAP), ref["tiles"])
        attr = ctx.download(gs.BUF_ATTR)
        for field in ["conic_opacity", "color_radii", "aabb", "uv", "depth", "magic"]:
            assert np.array_equal(attr[field], ref["attr"][field]), field
        # P1: scan.  BUF_PREFIX_SUM is the reference's layout (index order), derived on the host from the device tile
        # counts; the scan the DEVICE runs is over the depth-sorted survivors inside k_emit and is pinned directly:
        assert np.array_equal(ctx.download(gs.BUF_PREFIX_SUM), ref["scan"])
        vis = np.nonzero(ref["tiles"] > 0)[0]
        dorder = vis[np.argsort(ref["attr"]["depth"][vis].view(np.uint32), kind="stable")]  # (depth bits, index) order
        assert np.array_equal(ctx.download(gs.BUF_DEPTH_ORDER), dorder.astype(np.uint32))   # the Gaussian-level sort
        excl = np.concatenate([[0], np.cumsum(ref["tiles"][dorder].astype(np.uint64))[:-1]]) if dorder.size else np.empty(0, np.uint64)
        assert np.array_equal(ctx.download(gs.BUF_EMIT_OFFSETS), excl.astype(np.uint64))    # the device scan
        # P2 + the reference's radix passes 0-3: the device emits the instances in (depth, Gaussian) order, x outer /
        # y inner inside a Gaussian == preprocess_sort.comp's output stably sorted by its low 32 key bits
        order = np.argsort(ref["keys_unsorted"] & np.uint64(0xFFFFFFFF), kind="stable")
        assert np.array_equal(ctx.download(gs.BUF_KEYS_UNSORTED), ref["keys_unsorted"][order])
        assert np.array_equal(ctx.download(gs.BUF_VALS_UNSORTED), ref["vals_unsorted"][order])
        # P3: sort (stable => ties keep Gaussian-index order)
        assert np.array_equal(ctx.download(gs.BUF_KEYS_SORTED), ref["keys"])
        assert np.array_equal(ctx.download(gs.BUF_VALS_SORTED), ref["vals"])
        # P4: tile ranges
        assert np.array_equal(ctx.download(gs.BUF_TILE_BOUNDARY), ref["ranges"])
        # P5: blend, bit-exact against the shared-definition exp ...
        assert np.array_equal(img, ref["rgba"])
        assert st.blend_consumed == int(ref["consumed"].sum())
        # ... and within the north_star tolerance of the libm-exp restatement
        ref0 = oracle_frame(oracle, vtx, u, 0)
        assert np.abs(img - ref0["rgba"]).max() <= TOL
    finally:
        ctx.set_debug(False)


def test_fast_mode_within_tolerance(gs, oracle, ctx):
    _, vtx, u = scenes.c1()
    ctx.upload(vtx)
    ctx.set_mode(gs.MODE_FAST)
    try:
        img = ctx.render(u, gs.FORMAT_RGBA32F)
    finally:
        ctx.set_mode(gs.MODE_EXACT)
    ref0 = oracle_frame(oracle, vtx, u, 0)
    err = np.abs(img - ref0["rgba"])
    assert err.max() <= TOL, f"Linf {err.max()} pixels over tol {(err > TOL).any(axis=-1).sum()}"


@pytest.mark.parametrize("fmt_name,bgra", [("FORMAT_RGBA8", False), ("FORMAT_BGRA8", True)])
def test_unorm8_formats(gs, oracle, ctx, fmt_name, bgra):
    _, vtx, _ = scenes.c1()
    u = scenes.camera("odd_size")
    ctx.upload(vtx)
    img8 = ctx.render(u, getattr(gs, fmt_name))
    ref = oracle_frame(oracle, vtx, u, 1)
    assert np.array_equal(img8, oracle.pack_unorm8(ref["rgba"], bgra=bgra))


def test_band_render_equals_full_frame(gs, oracle, ctx):
    """Multi-GPU sharding unit: tile-row bands rendered separately == the full frame."""
    _, vtx, u = scenes.c1()
    ctx.upload(vtx)
    full = ctx.render(u, gs.FORMAT_RGBA32F)
    tiles_y = (u.height + 15) // 16
    for parts in (2, 3, 4):
        rows = [(tiles_y * k // parts, tiles_y * (k + 1) // parts) for k in range(parts)]
        bands = [ctx.render(u, gs.FORMAT_RGBA32F, rows=r) for r in rows]
        assert np.array_equal(np.concatenate(bands, axis=0), full)
    # and the band itself matches the oracle's band semantics (AABB clipped before counting)
    r = (tiles_y // 3, 2 * tiles_y // 3)
    band = ctx.render(u, gs.FORMAT_RGBA32F, rows=r)
    st = ctx.stats()
    ref = oracle_frame(oracle, vtx, u, 1, rows=r)
    assert st.num_instances == ref["m"]
    assert np.array_equal(band, ref["rgba"][r[0] * 16:min(u.height, r[1] * 16)])


def test_arena_regrow(gs, oracle):
    """Instance arena smaller than M: gsb_render regrows and re-renders (Renderer.cpp:541-563 analogue)."""
    _, vtx, u = scenes.c1(n=3000)
    c = gs.Context(0)
    try:
        c.upload(vtx)  # arena starts at max(N, 1024) = 3000 < M
        img = c.render(u, gs.FORMAT_RGBA32F)
        st = c.stats()
        ref = oracle_frame(oracle, vtx, u, 1)
        assert ref["m"] > 3000 and st.regrow_count >= 1 and st.instance_capacity >= ref["m"]
        assert np.array_equal(img, ref["rgba"])
    finally:
        c.close()


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257])
def test_small_and_empty_scenes(gs, oracle, n):
    _, vtx, u = scenes.c1(n=max(n, 1))
    vtx = vtx[:n]
    c = gs.Context(0)
    try:
        c.upload(vtx)
        img = c.render(u, gs.FORMAT_RGBA32F)
        ref = oracle_frame(oracle, vtx, u, 1)
        assert np.array_equal(img, ref["rgba"])
        assert c.stats().num_instances == ref["m"]
    finally:
        c.close()


def test_render_is_idempotent_and_mode_switch(gs, ctx):
    _, vtx, u = scenes.c1()
    ctx.upload(vtx)
    a = ctx.render(u, gs.FORMAT_RGBA32F)
    b = ctx.render(u, gs.FORMAT_RGBA32F)
    assert np.array_equal(a, b)


def test_errors(gs):
    c = gs.Context(0)
    try:
        u = scenes.camera("c1")
        with pytest.raises(gs.GsbError) as e:
            c.render(u)
        assert e.value.code == gs.ERR_NO_SCENE
        c.upload(scenes.c1(n=10)[1])
        u.width = 0
        with pytest.raises(gs.GsbError) as e:
            c.render(u)
        assert e.value.code == gs.ERR_INVALID
    finally:
        c.close()
    with pytest.raises(gs.GsbError) as e:
        gs.Context(9999)
    assert e.value.code == gs.ERR_NO_DEVICE


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size", "wide"])
def test_tile_cull_keeps_the_image_bit_exact(gs, oracle, ctx, cam):
    """gsb_set_tile_cull: instances that provably stay below alpha < 1/255 on a whole tile are dropped at emission.
    The image must stay bit-identical; the sorted (key, payload) list must be an ordered subset of the reference's;
    every dropped instance must be one that contributes to no pixel of its tile in the oracle."""
    _, vtx, _ = scenes.c1()
    u = scenes.camera(cam)
    ctx.set_mode(gs.MODE_EXACT)
    ctx.set_debug(True)
    ctx.set_tile_cull(True)
    try:
        ctx.upload(vtx)
        img = ctx.render(u, gs.FORMAT_RGBA32F)
        st = ctx.stats()
        ref = oracle_frame(oracle, vtx, u, 1)
        assert np.array_equal(img, ref["rgba"])
        assert st.num_instances_aabb == ref["m"] and st.num_instances <= ref["m"]
        keys, vals = ctx.download(gs.BUF_KEYS_SORTED), ctx.download(gs.BUF_VALS_SORTED)
        # ordered subset: (key, val) pairs are unique, so membership + order can be checked via a merge
        full = {(int(k), int(v)): i for i, (k, v) in enumerate(zip(ref["keys"], ref["vals"]))}
        pos = np.array([full[(int(k), int(v))] for k, v in zip(keys, vals)], dtype=np.int64)
        assert np.all(np.diff(pos) > 0)
        if ref["m"]:
            # dropped instances: max alpha over the tile's pixels < 1/255 (checked in float64 on a sample)
            kept = np.zeros(ref["m"], bool)
            kept[pos] = True
            dropped = np.nonzero(~kept)[0]
            rng = np.random.default_rng(1)
            tiles_x = (u.width + 15) // 16
            for i in rng.choice(dropped, size=min(300, dropped.size), replace=False) if dropped.size else []:
                a = ref["attr"][ref["vals"][i]]
                t = int(ref["keys"][i] >> np.uint64(32))
                xs = (t % tiles_x) * 16 + np.arange(16)
                ys = (t // tiles_x) * 16 + np.arange(16)
                dx, dy = a["uv"][0] - xs[None, :].astype(np.float64), a["uv"][1] - ys[:, None].astype(np.float64)
                co = a["conic_opacity"].astype(np.float64)
                power = -0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy
                assert (co[3] * np.exp(power)).max() < 1.0 / 255.0
        if cam == "c1":
            assert st.num_instances < ref["m"]  # the cull actually removes something
    finally:
        ctx.set_tile_cull(False)
        ctx.set_debug(False)


@pytest.mark.parametrize("cam", ["c1", "inside", "odd_size", "wide", "tiny", "away"])
def test_coarse_bins_keep_the_image_bit_exact(gs, oracle, ctx, cam, monkeypatch):
    """gsb_set_tile_cull level 2: the instance sort runs over blocks of 4x4 (or 2x2) tiles, the key carries the mask of the
    block's tiles inside the Gaussian's tile AABB, and every tile keeps the entries with its bit while walking the block's
    list.  The image must equal the oracle's bit for bit in every format, with timers (direct launches) and without (graph
    replay), and the instance count must be the number of (Gaussian, block) pairs of the oracle's AABBs."""
    _, vtx, _ = scenes.c1()
    u = scenes.camera(cam)
    ref = oracle_frame(oracle, vtx, u, 1)
    for shift in ("2", "1"):
        monkeypatch.setenv("GSB_COARSE_SHIFT", shift)
        c = gs.Context(0)
        try:
            c.set_mode(gs.MODE_EXACT)
            c.set_tile_cull(2)
            c.upload(vtx)
            for timers in (True, False, False):
                c.set_timers(timers)
                assert np.array_equal(c.render(u, gs.FORMAT_RGBA32F), ref["rgba"]), (cam, shift, timers)
            c.set_timers(True)
            assert np.array_equal(c.render(u, gs.FORMAT_BGRA8), oracle.pack_unorm8(ref["rgba"], bgra=True)), (cam, shift)
            st = c.stats()
            a = ref["attr"]["aabb"].astype(np.int64)
            live = (a[:, 2] > a[:, 0]) & (a[:, 3] > a[:, 1])
            s = int(shift)
            blocks = ((((a[:, 2] - 1) >> s) - (a[:, 0] >> s) + 1) * (((a[:, 3] - 1) >> s) - (a[:, 1] >> s) + 1))[live].sum()
            assert st.num_instances == blocks and st.num_instances_aabb == ref["m"], (cam, shift)
            c.set_mode(gs.MODE_FAST)
            assert np.abs(c.render(u, gs.FORMAT_RGBA32F) - ref["rgba"]).max() <= 1e-4
        finally:
            c.close()


def test_fp16_sh_storage_is_a_bounded_non_parity_mode(gs, oracle):
    """gsb_set_sh_storage(1): SH coefficients are stored as fp16 (half of k_project's dominant traffic).  It is flagged
    non-parity: geometry, culling and instance lists are untouched (same counts, same alpha channel), colours move by the
    half-precision rounding of the coefficients -- well below 8-bit resolution but above the 1e-4 parity tolerance."""
    _, vtx, u = scenes.c1()
    ref = oracle_frame(oracle, vtx, u, 1)
    c = gs.Context(0)
    try:
        c.upload(vtx)
        exact = c.render(u, gs.FORMAT_RGBA32F)
        st0 = c.stats()
        assert np.array_equal(exact, ref["rgba"])
        c.set_sh_storage(True)
        c.upload(vtx)
        for level in (0, 2):
            c.set_tile_cull(level)
            img = c.render(u, gs.FORMAT_RGBA32F)
            st = c.stats()
            assert (st.num_visible, st.num_instances_aabb) == (st0.num_visible, st0.num_instances_aabb)
            err = np.abs(img - exact).max()
            assert 0.0 < err < 4e-3, err
            assert np.isfinite(img).all() and np.array_equal(img[..., 3], exact[..., 3])
        c.set_sh_storage(False)  # back to fp32 at the next upload: bit-exact again
        c.set_tile_cull(0)
        c.upload(vtx)
        assert np.array_equal(c.render(u, gs.FORMAT_RGBA32F), ref["rgba"])
    finally:
        c.close()


def test_device_scan_with_tile_cull(gs, oracle, ctx):
    """With the exact instance cull on, k_emit scans the CULLED per-Gaussian counts: its offsets must partition the
    emitted list exactly (offset[j+1] - offset[j] instances of Gaussian depth_order[j], contiguous, in that order)."""
    _, vtx, u = scenes.c1()
    ctx.set_debug(True)
    ctx.set_tile_cull(True)
    try:
        ctx.upload(vtx)
        ctx.render(u, gs.FORMAT_RGBA32F)
        m = ctx.stats().num_instances
        order, offs = ctx.download(gs.BUF_DEPTH_ORDER), ctx.download(gs.BUF_EMIT_OFFSETS)
        vals = ctx.download(gs.BUF_VALS_UNSORTED)
        assert offs[0] == 0 and np.all(np.diff(offs.astype(np.int64)) >= 0) and offs[-1] <= m
        counts = np.diff(np.concatenate([offs, [m]]).astype(np.int64))
        assert np.array_equal(np.repeat(order, counts), vals)
    finally:
        ctx.set_tile_cull(False)
        ctx.set_debug(False)


def test_graph_replay_equals_direct_launches(gs, ctx):
    """gsb_set_graph: the captured middle of the frame (both sorts + emission) replays bit-identically, across cameras,
    frame sizes and a mid-sequence arena regrow (which invalidates the captured pointers)."""
    _, vtx, _ = scenes.c1()
    c = gs.Context(0)
    try:
        c.upload(vtx)
        c.set_timers(False)
        frames = {}
        for graph in (False, True):
            c.set_graph(graph)
            for rep in range(2):
                for cam in ("c1", "inside", "odd_size", "wide", "tiny", "c1"):
                    img = c.render(scenes.camera(cam), gs.FORMAT_RGBA32F)
                    assert np.array_equal(frames.setdefault(cam, img), img), (graph, cam)
        c.set_timers(True)
        assert np.array_equal(c.render(scenes.camera("c1"), gs.FORMAT_RGBA32F), frames["c1"])
        assert c.stats().frame_ms > 0
    finally:
        c.close()


def test_async_overflow_is_sticky(gs, ctx):
    """an overflow in ANY gsb_render_async frame must be reported by the next gsb_get_stats, even if later
    frames fit (each frame rewrites the per-frame flag)."""
    import torch
    _, vtx, _ = scenes.c1(n=3000)
    c = gs.Context(0)
    try:
        c.upload(vtx)  # arena = 3000 entries
        big, small = scenes.camera("c1"), scenes.camera("away")  # "away": everything culled, M = 0
        out = torch.zeros((480, 640, 4), dtype=torch.float32, device="cuda:0")
        c.render_into(big, out.data_ptr(), gs.FORMAT_RGBA32F, sync=False)    # overflows (M > 3000)
        small_out = torch.zeros((240, 320, 4), dtype=torch.float32, device="cuda:0")
        c.render_into(small, small_out.data_ptr(), gs.FORMAT_RGBA32F, sync=False)  # fits
        with pytest.raises(gs.GsbError) as e:
            c.stats()
        assert e.value.code == gs.ERR_OVERFLOW
        c.render_into(small, small_out.data_ptr(), gs.FORMAT_RGBA32F, sync=False)
        assert c.stats().num_instances == 0  # reported once, then cleared
    finally:
        c.close()


def test_opacity_above_one_matches_the_shader(gs, oracle):
    """gsb_scene_upload takes raw GSScene::Vertex arrays; an opacity > 1 lowers the alpha >= 1/255 cut below
    -5.55 and the kernel must keep those pairs exactly like render.comp:77-80 (min(0.99, opacity * exp(power)))."""
    _, vtx, u = scenes.c1(n=2000)
    vtx = vtx.copy()
    vtx[::3, 7] = 40.0   # scale_opacity.w
    vtx[1::7, 7] = 3.0
    c = gs.Context(0)
    try:
        c.upload(vtx)
        for cull in (False, True):
            c.set_tile_cull(cull)
            img = c.render(u, gs.FORMAT_RGBA32F)
            ref = oracle_frame(oracle, vtx, u, 1)
            assert np.array_equal(img, ref["rgba"]), cull
            assert np.abs(img - oracle_frame(oracle, vtx, u, 0)["rgba"]).max() <= TOL
    finally:
        c.close()


@pytest.mark.parametrize("fmt_name", ["FORMAT_RGBA32F", "FORMAT_BGRA8"])
def test_host_output_bands_equal_device_output(gs, ctx, fmt_name):
    """gsb_render to HOST memory blends in row bands and copies each band while the next ones blend; pageable, pinned
    and pitched destinations must all equal the single-kernel device-output frame."""
    import ctypes as C
    import torch
    _, vtx, _ = scenes.c1()
    fmt = getattr(gs, fmt_name)
    ctx.upload(vtx)
    for cam in ("c1", "odd_size", "tiny", "wide"):
        u = scenes.camera(cam)
        bpp = 16 if fmt == gs.FORMAT_RGBA32F else 4
        dt = torch.float32 if fmt == gs.FORMAT_RGBA32F else torch.uint8
        dev = torch.zeros((u.height, u.width, 4), dtype=dt, device="cuda:0")
        ctx.render_into(u, dev.data_ptr(), fmt, sync=True)
        ref = dev.cpu().numpy()
        assert np.array_equal(ctx.render(u, fmt), ref)                       # pageable numpy destination
        p = C.c_void_p()
        pitch = u.width * bpp + 64                                           # pinned + padded rows
        assert gs.lib.gsb_host_alloc(C.byref(p), pitch * u.height) == 0
        try:
            ctx._ck(gs.lib.gsb_render(ctx.h, C.byref(u), 0, gs.ALL_ROWS, p, pitch, gs.MEM_HOST, fmt, None))
            raw = np.frombuffer((C.c_char * (pitch * u.height)).from_address(p.value), np.uint8).reshape(u.height, pitch)
            got = raw[:, :u.width * bpp].copy().view(ref.dtype).reshape(ref.shape)
            assert np.array_equal(got, ref)
        finally:
            gs.lib.gsb_host_free(p)
