"""Multi-GPU host logic on CPU: world_size-2/3 gloo, the CPU oracle standing in for the GPU kernels.
(1) The framebuffer all-gather of the equal-height bands (the GSB_SHARD_GATHER=nccl baseline of gsb_render_sharded).
(2) The frame-sharding exchange of csrc/gsb_shard.cu: the scene is sliced by gsb_shard_slice (the C ABI's own partition),
    every rank projects only its slice, delivers each survivor to the bands its tile AABB touches -- into the region
    reserved for this source, in Gaussian-index order -- and every band blends the concatenation of its regions.  The
    model checks the two claims the CUDA path relies on: the assembled per-band survivor list is the single-process
    band's list in the same order, and the band's pixels are bit-identical to the full frame's."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "3dgs.cpp_b200" / "python"))
    sys.path.insert(0, str(root / "oracle"))
    sys.path.insert(0, str(root / "tests"))
    import gs_b200 as g
    import oracle as o
    import scenes

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, vtx, _ = scenes.c1(n=1500)
        u = scenes.camera("odd_size")  # 333 x 217: H is not a multiple of 16, the last band is short
        W, H = u.width, u.height
        rb, re, rows_per = g.band_for_rank(H, rank, world)
        band = torch.zeros((rows_per * 16, W, 4), dtype=torch.float32)
        if rb < re:
            f = o.render_frame(vtx, o.cov3d(vtx), u, rows=(rb, re))
            nrows = min(H, re * 16) - rb * 16
            band[:nrows] = torch.from_numpy(f["rgba"][rb * 16:rb * 16 + nrows])
        full = torch.zeros((world * rows_per * 16, W, 4), dtype=torch.float32)
        dist.all_gather_into_tensor(full.view(-1), band.view(-1))
        if rank == 0:
            ref = o.render_frame(vtx, o.cov3d(vtx), u)["rgba"]
            out.put(bool(np.array_equal(full[:H].numpy(), ref)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_allgather_reassembles_the_frame(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def _exchange_worker(rank, world, port, out):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "3dgs.cpp_b200" / "python"))
    sys.path.insert(0, str(root / "oracle"))
    sys.path.insert(0, str(root / "tests"))
    import gs_b200 as g
    import oracle as o
    import scenes

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, vtx, _ = scenes.c1(n=3001)  # not a multiple of the world size: the last slice is short
        u = scenes.camera("odd_size")
        n = vtx.shape[0]
        first, count = g.shard_slice(n, rank, world)
        slices = [g.shard_slice(n, r, world) for r in range(world)]
        assert sum(c for _, c in slices) == n and all(slices[r][0] + slices[r][1] == slices[r + 1][0] for r in range(world - 1))
        # source side: project the local slice (whole frame), route by band
        attr, tiles = o.preprocess(vtx[first:first + count], o.cov3d(vtx[first:first + count]), u)
        aabb = attr["aabb"].astype(np.int64)
        surv = np.nonzero(tiles)[0]
        regions = []  # regions[d] = global indices of my survivors whose AABB touches band d, in index order
        for d in range(world):
            b0, b1, _ = g.band_for_rank(u.height, d, world)
            touch = np.maximum(aabb[surv, 1], b0) < np.minimum(aabb[surv, 3], b1)
            regions.append((first + surv[touch]).tolist())
        gathered = [None] * world
        dist.all_gather_object(gathered, regions)  # stands in for the stores into peer memory
        # destination side: my band's survivor list = my regions of every source, in source order
        mine = np.array([i for src in range(world) for i in gathered[src][rank]], dtype=np.int64)
        rb, re, _ = g.band_for_rank(u.height, rank, world)
        ok = True
        if rb < re:
            attr_b, tiles_b = o.preprocess(vtx, o.cov3d(vtx), u, rows=(rb, re))  # what ONE process would see for this band
            ok = np.array_equal(mine, np.nonzero(tiles_b)[0])
            # blending exactly those Gaussians (band-clipped) reproduces the full frame's rows
            sub = o.render_frame(vtx[mine], o.cov3d(vtx[mine]), u, rows=(rb, re))["rgba"]
            full = o.render_frame(vtx, o.cov3d(vtx), u)["rgba"]
            r0, r1 = rb * 16, min(u.height, re * 16)
            ok = ok and np.array_equal(sub[r0:r1], full[r0:r1])
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            out.put(all(flags))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slice_route_assemble_model_of_the_sharded_exchange(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(timeout=5) is True
