"""world_size-2 gloo test of the multi-GPU host logic on CPU: interleaved tile ownership, the rank-major
gather layout and the un-tile step.  Each rank fills its tile-major local buffer with the C oracle's
pixels for the tiles it owns (the oracle stands in for the CUDA kernels here — this test is about the
plumbing, not the renderer), rank 0 gathers over gloo and un-tiles; the result must equal the full frame."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden_scene


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from cutrace_b200.distributed import local_tile_count, tiles_of_rank, untile_host
    from cutrace_b200.scene import TILE, FlatScene
    from oracle import pyoracle as po

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = FlatScene.load(os.path.join(ROOT, "tests", "golden", "scenes", "mirror.npz")).with_resolution(w, h)
    tx = (w + TILE - 1) // TILE
    nlt = local_tile_count(w, h, world)
    depth = torch.full((nlt * TILE * TILE,), float("inf"))
    color = torch.zeros((nlt * TILE * TILE, 3))
    px, dst = [], []
    for lt, gt in enumerate(tiles_of_rank(w, h, rank, world)):
        ty, txi = divmod(gt, tx)
        for py in range(TILE):
            for pxx in range(TILE):
                x, y = txi * TILE + pxx, ty * TILE + py
                if x < w and y < h:
                    px.append(y * w + x)
                    dst.append(lt * TILE * TILE + py * TILE + pxx)
    o = po.oracle_render(s, px=np.asarray(px, np.uint64), threads=2)
    depth[dst] = torch.from_numpy(o["depth"])
    color[dst] = torch.from_numpy(o["color"])
    gd = [torch.empty_like(depth) for _ in range(world)] if rank == 0 else None
    gc = [torch.empty_like(color) for _ in range(world)] if rank == 0 else None
    dist.gather(depth, gd, dst=0)
    dist.gather(color, gc, dst=0)
    mx = torch.tensor([float(o["depth"][np.isfinite(o["depth"])].max())])
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        D = untile_host([g.numpy().reshape(-1, 1) for g in gd], w, h, 1)
        Cc = untile_host([g.numpy() for g in gc], w, h, 3)
        np.savez(out_path, depth=D, color=Cc, max_depth=mx.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(100, 70)])
def test_tile_gather_untile_world2_gloo(tmp_path, oracle, w, h):
    import torch.multiprocessing as mp

    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), w, h, out), nprocs=2, join=True)
    got = np.load(out)
    s = load_golden_scene("mirror").with_resolution(w, h)
    full = oracle.oracle_render(s)
    assert np.array_equal(got["depth"].reshape(-1).view(np.uint32), full["depth"].view(np.uint32))
    assert np.array_equal(got["color"].reshape(-1, 3).view(np.uint32), full["color"].view(np.uint32))
    assert float(got["max_depth"][0]) == oracle.max_depth(full["depth"])


def test_tile_ownership_is_a_partition():
    from cutrace_b200.distributed import local_tile_count, tiles_of_rank

    for (w, h, world) in [(20, 20, 1), (3840, 2160, 8), (100, 70, 3), (7680, 4320, 4)]:
        allt = sorted(t for r in range(world) for t in tiles_of_rank(w, h, r, world))
        from cutrace_b200.scene import TILE
        n = ((w + TILE - 1) // TILE) * ((h + TILE - 1) // TILE)
        assert allt == list(range(n))
        assert all(len(tiles_of_rank(w, h, r, world)) <= local_tile_count(w, h, world) for r in range(world))


def _host_frame_worker(rank, world, port, w, h, out_path):
    """exchange="host" without a GPU: every rank writes the oracle's pixels of ITS tiles straight into the shared host frame
    (what the kernels do over PCIe after cutrace_frame_attach), a barrier, rank 0 reads the assembled frame."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from cutrace_b200.distributed import SharedHostFrame, tiles_of_rank
    from cutrace_b200.scene import TILE, FlatScene
    from oracle import pyoracle as po

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = FlatScene.load(os.path.join(ROOT, "tests", "golden", "scenes", "sphere_plane.npz")).with_resolution(w, h)
    frame = SharedHostFrame(w, h, rank, world, device=None, register=False)
    tx = (w + TILE - 1) // TILE
    px = []
    for gt in tiles_of_rank(w, h, rank, world):
        ty, txi = divmod(gt, tx)
        for py in range(TILE):
            for pxx in range(TILE):
                x, y = txi * TILE + pxx, ty * TILE + py
                if x < w and y < h:
                    px.append(y * w + x)
    px = np.asarray(px, np.uint64)
    o = po.oracle_render(s, px=px, threads=2)
    idx = px.astype(np.int64)
    frame.depth[idx] = o["depth"]; frame.normal[idx] = o["normal"]; frame.color[idx] = o["color"]; frame.hit_id[idx] = o["hit_id"]
    dist.barrier()
    if rank == 0:
        np.savez(out_path, **{k: np.array(v) for k, v in frame.as_dict().items()})
    dist.barrier()
    frame.close()
    dist.destroy_process_group()


def test_world2_shared_host_frame(tmp_path, oracle):
    import torch.multiprocessing as mp

    w, h = 70, 45
    out = str(tmp_path / "hostframe.npz")
    port = _free_port()
    mp.spawn(_host_frame_worker, args=(2, port, w, h, out), nprocs=2, join=True)
    got = np.load(out)
    ref = oracle.oracle_render(load_golden_scene("sphere_plane").with_resolution(w, h))
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(got[k].reshape(-1).view(np.uint32), ref[k].reshape(-1).view(np.uint32)), k


def _barrier_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import time

    import torch.distributed as dist

    from cutrace_b200.distributed import HostBarrier

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    hb = HostBarrier(rank, world)
    got = []
    for k in range(200):
        if rank == k % world:
            time.sleep(0.0005)           # a different straggler every round
        got.append(hb.max(float(k * 10 + rank)))
    hb.close()
    if rank == 0:
        np.save(out_path, np.asarray(got))
    dist.destroy_process_group()


def test_host_barrier_max_world3(tmp_path):
    """HostBarrier: 200 rounds with a rotating straggler; every round returns the max over the ranks of THAT round."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "hb.npy")
    mp.spawn(_barrier_worker, args=(3, _free_port(), out), nprocs=3, join=True)
    got = np.load(out)
    assert np.array_equal(got, np.arange(200) * 10.0 + 2.0)
