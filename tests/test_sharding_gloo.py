"""Multi-GPU host logic on CPU: world_size-2 gloo.  Each rank produces its tile-row band (here with the
CPU oracle standing in for the GPU render), pads it to the equal band height and all-gathers; rank 0
checks the reassembled framebuffer against the full-frame render.  This is exactly the data path
bench.py uses under torchrun with the nccl backend (one all_gather_into_tensor after the blend)."""
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
