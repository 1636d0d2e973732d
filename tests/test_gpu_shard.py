"""Frame sharding over several GPUs through the C ABI: gsb_group_* with every rank on cuda:0.  The scene is
sharded by Gaussian index, the frame by tile rows, survivors and framebuffer bands travel by stores into peer-mapped memory
ordered by mailbox flags -- the same kernels and protocol the multi-GPU runs use, with "peer" pointers that happen to live on
one device, so the single-GPU suite covers routing, slot determinism, the gather, the peer-store blend and the regrow path.
The result must be BIT-IDENTICAL to the single-context frame (and so to the oracle)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import scenes

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))  # bench.py (workload table)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_frame_is_bit_identical_to_single_gpu(gs, ctx, world):
    _, vtx, _ = scenes.c1()
    ctx.set_mode(gs.MODE_EXACT)
    ctx.upload(vtx)
    g = gs.Group([0] * world)
    try:
        g.upload(vtx)
        for cam in ("c1", "inside", "odd_size", "wide", "tiny", "away", "c1"):
            u = scenes.camera(cam)
            for fmt in (gs.FORMAT_RGBA32F, gs.FORMAT_BGRA8):
                assert np.array_equal(g.render(u, fmt), ctx.render(u, fmt)), (world, cam, fmt)
        # per-rank counts: bands partition the instances exactly like single-context band renders do
        u = scenes.camera("c1")
        g.render(u, gs.FORMAT_RGBA32F)
        tiles_y = (u.height + 15) // 16
        R = (tiles_y + world - 1) // world
        for r in range(world):
            st = g.context(r).stats()
            rows = (min(tiles_y, r * R), min(tiles_y, (r + 1) * R))
            if rows[0] < rows[1]:
                ctx.render(u, gs.FORMAT_RGBA32F, rows=rows)
                ref = ctx.stats()
                assert (st.num_visible, st.num_instances) == (ref.num_visible, ref.num_instances), (world, r)
                # the tile ranges too: a range that starts too early still blends to the same image (the extra records are
                # culled per block) but shows up as extra consumed entries
                assert st.blend_consumed == ref.blend_consumed, (world, r)
            else:
                assert st.num_instances == 0
    finally:
        g.close()


def test_sharded_matches_oracle_with_cull_and_fast_paths(gs, oracle):
    _, vtx, u = scenes.c1()
    oracle.set_exp_mode(1)
    try:
        ref = oracle.render_frame(vtx, oracle.cov3d(vtx), u)["rgba"]
    finally:
        oracle.set_exp_mode(0)
    g = gs.Group([0, 0, 0, 0])
    try:
        g.upload(vtx)
        for cull in (0, 1, 2):  # reference lists, exact instance culling, coarse 4x4-tile bins
            for timers in (True, False):  # timers off: the middle of the frame replays from per-parity CUDA graphs
                for r in range(4):
                    c = g.context(r)
                    c.set_tile_cull(cull)
                    c.set_timers(timers)
                for _ in range(3):  # both exchange parities, graph capture + replay
                    assert np.array_equal(g.render(u, gs.FORMAT_RGBA32F), ref), (cull, timers)
    finally:
        g.close()


def test_sharded_arena_regrow_is_collective(gs, ctx):
    _, vtx, u = scenes.c1(n=3000)
    ctx.upload(vtx)
    want = ctx.render(u, gs.FORMAT_RGBA32F)
    g = gs.Group([0, 0])
    try:
        g.upload(vtx)  # arenas start at 2 * slice = 3000 entries < the band's M
        got = g.render(u, gs.FORMAT_RGBA32F)
        assert np.array_equal(got, want)
        assert sum(g.context(r).stats().regrow_count for r in range(2)) >= 1
        # async frames + a size change (window re-exchange) keep working
        for _ in range(5):
            g.render_async(u, gs.FORMAT_BGRA8)
        assert np.array_equal(g.render(scenes.camera("odd_size"), gs.FORMAT_RGBA32F), ctx.render(scenes.camera("odd_size"), gs.FORMAT_RGBA32F))
    finally:
        g.close()


def test_sharded_full_size_garden(gs):
    """BASELINE's headline workload over 4 ranks == the single-context frame, bit for bit (cull on, BGRA8 and float)."""
    import bench
    wl = bench.WORKLOADS["garden-standin"]
    vtx = bench.make_scene(gs, wl)
    u = bench.cameras(gs, wl)[2]
    c = gs.Context(0)
    g = gs.Group([0, 0, 0, 0])
    try:
        c.set_tile_cull(True)
        c.upload(vtx)
        want = c.render(u, gs.FORMAT_RGBA32F)
        want8 = c.render(u, gs.FORMAT_BGRA8)
        c.close()
        c = None
        g.upload(vtx)
        for level in (1, 2):
            for r in range(4):
                g.context(r).set_tile_cull(level)
            assert np.array_equal(g.render(u, gs.FORMAT_RGBA32F), want), level
            assert np.array_equal(g.render(u, gs.FORMAT_BGRA8), want8), level
    finally:
        if c is not None:
            c.close()
        g.close()
