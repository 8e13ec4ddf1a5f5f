"""Screen-space tile sharding across GPUs: one process per GPU, scene replicated, tiles interleaved.

This is the only place the render path shards: pixels are independent
(/root/reference/inc/kernel.hpp:37-59 has no inter-thread communication), so rank r renders the
16x16 tiles whose slot s has s % world == r; slots walk the frame super-tile by super-tile (9 x 8 tiles, TileMap in
csrc/common.cuh), so a rank's tiles are a diagonal lattice over the whole image and its consecutive tiles are neighbours.
ONE exchange step follows.  Default ("peer"): every rank's kernels store their tiles straight into rank 0's
row-major frame over NVLink (CUDA IPC), so the exchange is fused into the producing kernels and only a
barrier remains.  Fallback ("gather"): tile-major local buffers, an NCCL gather of the float framebuffers
to rank 0, and an un-tile kernel there (cutrace_untile_device).  torch.distributed is plumbing only.
"""
from __future__ import annotations

import numpy as np

from . import Renderer
from .scene import TILE


def n_tiles(width, height):
    return ((width + TILE - 1) // TILE) * ((height + TILE - 1) // TILE)


def local_tile_count(width, height, world):
    return (n_tiles(width, height) + world - 1) // world


SUPER_W, SUPER_H = 9, 8   # CTB_SUPER_W / CTB_SUPER_H of csrc/common.cuh


def tile_of_slot(slot, width, height):
    """Screen tile (tx, ty) that tile slot ``slot`` shows — the Python mirror of tile_of_slot(curve = 1) in csrc/common.cuh
    (tests/test_abi.py checks it against the library's own function): super-tiles of SUPER_W x SUPER_H tiles in row-major order,
    row-major inside; the super-tiles of the last row / column are cut to what the frame has."""
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    sr, rem = divmod(slot, tiles_x * SUPER_H)
    h = min(SUPER_H, tiles_y - sr * SUPER_H)
    sc, rem2 = divmod(rem, SUPER_W * h)
    w = min(SUPER_W, tiles_x - sc * SUPER_W)
    iy, ix = divmod(rem2, w)
    return sc * SUPER_W + ix, sr * SUPER_H + iy


def slot_of_tile(tx, ty, width, height):
    """Inverse of tile_of_slot."""
    tiles_x, tiles_y = (width + TILE - 1) // TILE, (height + TILE - 1) // TILE
    sr, sc = ty // SUPER_H, tx // SUPER_W
    h = min(SUPER_H, tiles_y - sr * SUPER_H)
    w = min(SUPER_W, tiles_x - sc * SUPER_W)
    return sr * tiles_x * SUPER_H + sc * SUPER_W * h + (ty - sr * SUPER_H) * w + (tx - sc * SUPER_W)


def tiles_of_rank(width, height, rank, world):
    """Screen tile indices (row-major tile order) owned by ``rank``, in local-tile order."""
    tiles_x = (width + TILE - 1) // TILE
    out = []
    for s in range(rank, n_tiles(width, height), world):
        tx, ty = tile_of_slot(s, width, height)
        out.append(ty * tiles_x + tx)
    return out


def untile_host(parts, width, height, channels):
    """CPU reference of the un-tile step (used by the gloo tests): ``parts[r]`` is rank r's tile-major
    buffer of shape (n_local_tiles*TILE*TILE, channels)."""
    world = len(parts)
    out = np.zeros((height, width, channels), parts[0].dtype)
    for y in range(height):
        for x0 in range(0, width, TILE):
            slot = slot_of_tile(x0 // TILE, y // TILE, width, height)
            r, lt = slot % world, slot // world
            cnt = min(TILE, width - x0)
            src = lt * TILE * TILE + (y % TILE) * TILE
            out[y, x0:x0 + cnt] = parts[r].reshape(-1, channels)[src:src + cnt]
    return out


class SharedHostFrame:
    """One row-major frame block (depth n | normal 3n | colour 3n | hit id n, 32 bytes per pixel) in host memory that every
    rank process of one box maps and registers with CUDA (cutrace_host_register): the ranks' kernels store their tiles
    straight into it over their own PCIe links (cutrace_frame_attach), so a multi-GPU frame reaches the host without being
    funnelled through GPU 0.  Rank 0 creates an anonymous memory file (memfd; /dev/shm as a fallback), the other ranks open it
    through /proc/<pid>/fd/<n> — torch.distributed only carries the path."""

    def __init__(self, width, height, rank, world, device=None, register=True):
        import ctypes as C
        import mmap
        import os

        import torch
        import torch.distributed as dist

        from . import _lib

        self.width, self.height, self.rank = width, height, rank
        n = width * height
        self.n = n
        self.nbytes = 32 * n
        self._unlink = None
        path = [None]
        if rank == 0:
            try:
                self._fd = os.memfd_create("cutrace_b200_frame")
                os.ftruncate(self._fd, self.nbytes)
                path[0] = f"/proc/{os.getpid()}/fd/{self._fd}"
            except (AttributeError, OSError):
                import tempfile

                fd, name = tempfile.mkstemp(prefix="cutrace_b200_frame_", dir="/dev/shm")
                os.ftruncate(fd, self.nbytes)
                self._fd, path[0], self._unlink = fd, name, name
        if world > 1:
            dist.broadcast_object_list(path, src=0, **({"device": torch.device("cuda", device)} if device is not None else {}))
        if rank != 0:
            self._fd = os.open(path[0], os.O_RDWR)
        self._map = mmap.mmap(self._fd, self.nbytes)
        self._buf = (C.c_char * self.nbytes).from_buffer(self._map)
        self.ptr = C.addressof(self._buf)
        self._registered = False
        if register:   # (register=False: host-side logic only, for the CPU tests)
            self._lib = _lib.load()
            torch.cuda.set_device(device)
            _lib.check(self._lib.cutrace_host_register(self.ptr, self.nbytes))
            self._registered = True
        base = np.frombuffer(self._map, dtype=np.float32)
        self.depth = base[:n]
        self.normal = base[n:4 * n].reshape(n, 3)
        self.color = base[4 * n:7 * n].reshape(n, 3)
        self.hit_id = base[7 * n:8 * n].view(np.uint32)
        if world > 1:
            dist.barrier()   # every rank has the file open before rank 0 may drop the name

    def as_dict(self):
        return dict(depth=self.depth, normal=self.normal, color=self.color, hit_id=self.hit_id)

    def close(self):
        import os

        if getattr(self, "_registered", False):
            self._lib.cutrace_host_unregister(self.ptr)
            self._registered = False
        for k in ("depth", "normal", "color", "hit_id", "_buf"):
            self.__dict__.pop(k, None)
        try:
            self._map.close()
        except (BufferError, AttributeError):
            pass   # numpy views still alive somewhere: the mapping goes away with them
        if getattr(self, "_fd", None) is not None:
            os.close(self._fd)
            self._fd = None
        if self._unlink and self.rank == 0:
            try:
                os.unlink(self._unlink)
            except OSError:
                pass


class HostBarrier:
    """Barrier + max-reduce of one float between the rank processes of one box through a few cache lines of shared host memory
    (memfd opened through /proc/<pid>/fd, like SharedHostFrame).  The exchanges that are fused into the kernels (peer frame, host
    frame) need exactly this per frame — "every rank's kernel has finished" and the frame-wide max depth (kernel.hpp:120-125).
    Each rank's cutrace_render returns after its stream is idle, so a host-side barrier is sufficient; it takes ~2 us where the
    NCCL all-reduce + the read-back of its result took 90 us — 6 % of a 1.4 ms frame at N = 8 (profiles/r02_scaling.md).
    x86-64 keeps stores in order, so a rank's value is visible before its sequence number."""

    LINE = 16   # 128-byte slots (float64 view)

    def __init__(self, rank, world, device=None):
        import mmap
        import os

        import torch
        import torch.distributed as dist

        self.rank, self.world, self.seq = rank, world, 0
        nbytes = 8 * self.LINE * max(world, 1)
        path = [None]
        if rank == 0:
            self._fd = os.memfd_create("cutrace_b200_barrier")
            os.ftruncate(self._fd, nbytes)
            path[0] = f"/proc/{os.getpid()}/fd/{self._fd}"
        if world > 1:
            dist.broadcast_object_list(path, src=0, **({"device": torch.device("cuda", device)} if device is not None else {}))
        if rank != 0:
            self._fd = os.open(path[0], os.O_RDWR)
        self._map = mmap.mmap(self._fd, nbytes)
        self.slots = np.frombuffer(self._map, dtype=np.float64).reshape(max(world, 1), self.LINE)
        if world > 1:
            dist.barrier()

    def max(self, value, timeout_s=60.0):
        """Collective: returns the max of ``value`` over the ranks once every rank has called it."""
        import time

        self.seq += 1
        col = 1 + (self.seq & 1)            # values of even / odd rounds live in different words: a rank that is already in round
        mine = self.slots[self.rank]        # k+1 (it cannot be further ahead) never overwrites what a slower rank still reads for k
        mine[col] = float(value)
        mine[0] = float(self.seq)           # published last
        seqs = self.slots[:, 0]
        t0 = time.perf_counter()
        while float(seqs.min()) < self.seq:
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError("HostBarrier: a rank did not arrive")
        return float(self.slots[:, col].max())

    def close(self):
        import os

        self.slots = None
        try:
            self._map.close()
        except (BufferError, AttributeError):
            pass
        if getattr(self, "_fd", None) is not None:
            os.close(self._fd)
            self._fd = None


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class TileShardedRenderer:
    """Rank-local renderer of a tile-sharded frame.  Requires an initialised torch.distributed process group when
    world > 1.

    exchange="host": every rank's kernels store their tiles into a SharedHostFrame over their own PCIe link (the
    multi-GPU download path: nothing is funnelled through GPU 0).
    Default exchange ("peer"): rank 0 exports its row-major frame through CUDA IPC, the other ranks import it and
    their kernels store the G-buffer / final colour of their tiles straight into rank 0's HBM over NVLink — the
    transfer is fused into the producing kernels, and the only collective left is the 1-float max-depth all-reduce,
    which doubles as the barrier.  Fallback ("gather", also selectable): NCCL gather of the tile-major rank buffers
    to rank 0 followed by the device un-tile kernel.
    """

    def __init__(self, scene, rank, world, device, exchange="peer", host_frame=None, host_barrier=None, **kw):
        import torch

        self.host_barrier = host_barrier    # a HostBarrier shared by the ranks: replaces the NCCL all-reduce that ends a frame
        self.torch = torch
        self.rank, self.world = rank, world
        self.scene = scene
        self.device = torch.device("cuda", device)
        self.r = Renderer(scene, device=device, tile_rank=rank, tile_world=world, **kw)
        self.exchange = "none" if world == 1 else exchange
        n = scene.width * scene.height
        as_t = lambda p, cnt, ts: torch.as_tensor(_DevArray(p, cnt, ts), device=self.device)  # noqa: E731
        if self.exchange == "host":
            # every rank (rank 0 included) stores its tiles into the shared pinned host frame; nothing to set up per ctx
            # beyond the attach, so a ctx per frame costs no IPC round trip
            if host_frame is None:
                raise ValueError('exchange="host" needs a SharedHostFrame')
            self.host_frame = host_frame
            self.r.frame_attach(host_frame.ptr, host_frame.width, host_frame.height)
        if self.exchange == "peer":
            self.exchange = "peer" if self._setup_peer() else "gather"
        if self.exchange == "gather":
            (pd, pn, pc, pi), npx = self.r.device_buffers()
            self.npx = npx
            self.depth = as_t(pd, npx, "<f4")
            self.normal = as_t(pn, 3 * npx, "<f4")
            self.color = as_t(pc, 3 * npx, "<f4")
            self.hit_id = as_t(pi, npx, "<i4")
            if rank == 0:
                f32, i32 = torch.float32, torch.int32
                self.g_depth = torch.empty(world * npx, dtype=f32, device=self.device)
                self.g_normal = torch.empty(world * 3 * npx, dtype=f32, device=self.device)
                self.g_color = torch.empty(world * 3 * npx, dtype=f32, device=self.device)
                self.g_id = torch.empty(world * npx, dtype=i32, device=self.device)
                self.out_depth = torch.empty(n, dtype=f32, device=self.device)
                self.out_normal = torch.empty(3 * n, dtype=f32, device=self.device)
                self.out_color = torch.empty(3 * n, dtype=f32, device=self.device)
                self.out_id = torch.empty(n, dtype=i32, device=self.device)
        elif self.exchange in ("peer", "none") and rank == 0:
            pd, pn, pc, pi = self.r.frame_device()
            self.out_depth, self.out_normal = as_t(pd, n, "<f4"), as_t(pn, 3 * n, "<f4")
            self.out_color, self.out_id = as_t(pc, 3 * n, "<f4"), as_t(pi, n, "<i4")

    def _setup_peer(self):
        """rank 0 exports, everybody else imports; all ranks agree on success (else every rank falls back)."""
        import torch.distributed as dist

        torch = self.torch
        from ._lib import IPC_HANDLE_BYTES

        h = torch.zeros(IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        ok = 1
        try:
            if self.rank == 0:
                h.copy_(torch.frombuffer(bytearray(self.r.frame_ipc_export()), dtype=torch.uint8))
        except Exception:  # noqa: BLE001
            ok = 0
        dist.broadcast(h, src=0)
        if self.rank != 0 and ok:
            try:
                self.r.frame_ipc_import(bytes(h.cpu().numpy().tobytes()))
            except Exception:  # noqa: BLE001
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and self.rank != 0:
            self.r.frame_attach(None)
        return int(flag.item()) == 1

    def render(self):
        return self.r.render()

    def gather(self):
        """Completes the frame on rank 0.  peer: the stores already happened, only the ranks are synchronised (by the
        max-depth all-reduce the caller issues, or this barrier).  gather: NCCL gather + device un-tile on rank 0."""
        import torch.distributed as dist

        torch = self.torch
        if self.world == 1:
            return
        if self.exchange in ("peer", "host"):
            # the only collective of these exchanges: the 1-float max-depth reduction (kernel.hpp:120-125), which is also the
            # "every rank's tiles are stored" barrier (each rank's cutrace_render returned before it joined)
            if self.host_barrier is not None:
                self.frame_max_depth = self.host_barrier.max(self.r.stats()["max_depth"])
                return
            if not hasattr(self, "_md"):
                self._md = torch.zeros(1, dtype=torch.float32, device=self.device)
            self._md.fill_(float(self.r.stats()["max_depth"]))
            dist.all_reduce(self._md, op=dist.ReduceOp.MAX)
            self.frame_max_depth = float(self._md.item())
            return
        pairs = [(self.depth, getattr(self, "g_depth", None), 1), (self.normal, getattr(self, "g_normal", None), 3),
                 (self.color, getattr(self, "g_color", None), 3), (self.hit_id, getattr(self, "g_id", None), 1)]
        for src, dst, k in pairs:
            lst = list(dst.split(k * self.npx)) if self.rank == 0 else None
            dist.gather(src, lst, dst=0)
        if self.rank == 0:
            torch.cuda.current_stream(self.device).synchronize()
            self.r.untile_device(self.world, self.g_depth.data_ptr(), self.g_normal.data_ptr(), self.g_color.data_ptr(),
                                 self.g_id.data_ptr(), self.npx, self.out_depth.data_ptr(), self.out_normal.data_ptr(),
                                 self.out_color.data_ptr(), self.out_id.data_ptr())

    def max_depth(self, local_max):
        """max over ranks of the largest finite depth (kernel.hpp:120-125) — a 1-float all-reduce."""
        import torch.distributed as dist

        if self.host_barrier is not None and self.world > 1:
            return self.host_barrier.max(local_max)
        t = self.torch.tensor([local_max], dtype=self.torch.float32, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def release(self):
        """Collective: the owner of the frame has finished reading it.  In the peer / host exchanges the other ranks' kernels
        store straight into the frame, so the next render() must not start before its reader is done."""
        import torch.distributed as dist

        if self.world > 1 and self.exchange in ("peer", "host"):
            dist.barrier()

    def close(self):
        self.r.close()
