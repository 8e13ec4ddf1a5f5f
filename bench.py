#!/usr/bin/env python
"""bench.py — headline benchmark of the cutrace render path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own kernel (sm_100a rebuild)

Metric (BASELINE.json): Mrays/s (primary + secondary) on scene/bunny.json at 3840x2160; ms/frame is
`ms_per_step`.  Unique rays = primaries + reflection + transmission + shadow rays (one per light per
shaded hit); the same numerator is used for every implementation.

A step = one frame.  `value` times cutrace_render (+ the rank barrier when N > 1: the ranks' kernels store
their tiles into rank 0's frame over NVLink) with the scene already resident in HBM.  `e2e` times the whole
call sequence with host buffers, every step: upload (H2D + LBVH build) + render + results in pinned host
memory — N = 1: cutrace_render_download (copy engine), N > 1: every rank's kernels store their tiles into ONE
shared pinned host frame over their own PCIe link (cutrace_frame_attach on registered shared memory).
One process per GPU; timing: CUDA events / device time, max over ranks.

Besides the headline workload the line carries `extra_workloads` (BASELINE configs 1-3 and 5, same
measurements, fewer frames) and, for N > 1, `parity_check`: the assembled N-GPU frame against the same frame
rendered by one GPU, bit for bit.

The reference arm runs the UNMODIFIED reference kernel `render_kernel<default_gpu_scene,5>` rebuilt for
sm_100a from the reference headers in place (oracle/_ref/libcutrace_ref_gpu.so, built by oracle/Makefile in
the container) — the comparator BASELINE.json's north_star names.  It never loads this repo's CUDA library.
The reference has no CPU renderer; its device functions compiled for the host through a qualifier-erasing
shim (oracle/_ref/libcutrace_ref_host.so) are timed on a bounded pixel sample as `cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "triangle": dict(scene="triangle", width=20, height=20, label="scene/triangle.json@20x20"),
    "bunny4k": dict(scene="bunny", width=3840, height=2160, label="scene/bunny.json@3840x2160"),
    "mirror1080": dict(scene="mirror", width=1920, height=1080, label="scene/mirror.json@1920x1080"),
    "spheres1080": dict(scene="sphere_plane", width=1920, height=1080, label="scene/sphere_plane.json@1920x1080"),
    "synthetic10m": dict(scene="grid106", width=7680, height=4320, label="synthetic grid 106x106 (10,112,400 triangles)@7680x4320"),
}
# Unique rays per frame, a property of scene + resolution, identical for every implementation.  The reference arm
# uses these constants (it must not run this repo's renderer); tests/test_gpu_parity.py::test_bench_ray_constants pins them to what
# the renderer counts.
RAYS_PER_FRAME = {"triangle": 438, "bunny4k": 248_825_266, "mirror1080": 6_638_544, "spheres1080": 10_524_880, "synthetic10m": 406_265_436}
BOUNCES = 5


def load_workload(name):
    from cutrace_b200 import synth
    from cutrace_b200.scene import FlatScene

    w = dict(WORKLOADS[name], name=name)
    gold = os.path.join(ROOT, "tests", "golden", "scenes")
    if w["scene"].startswith("grid"):
        meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(gold, "bunny.npz")), FlatScene.load(os.path.join(gold, "mirror.npz")))[:2]
        return synth.grid_scene(meshes, grid=int(w["scene"][4:]), width=w["width"], height=w["height"]), w
    return FlatScene.load(os.path.join(gold, w["scene"] + ".npz")).with_resolution(w["width"], w["height"]), w


def n_primitives(scene):
    return scene.n_triangles + scene.n_spheres + scene.n_planes


def config_of(scene, wl):
    """The `config` object — the SAME keys and values in both arms."""
    n_px = scene.width * scene.height
    return {"workload": wl["label"], "rays_per_frame": RAYS_PER_FRAME[wl["name"]], "primitives": n_primitives(scene), "bounces": BOUNCES,
            "l2": "queues + framebuffer per frame > L2 (126 MB)" if n_px * 28 > 126e6 else "working set < L2; frames are re-rendered back to back"}


def survey_bytes_per_ray(n_prims):
    """SURVEY.md §8d's model figure: 64 B queue traffic + one 64-byte two-child node per level of a balanced tree + one 48-byte triangle."""
    import math

    return 64 + 64 * math.ceil(math.log2(max(n_prims, 2))) + 48


class ClockSampler:
    """SM clocks + throttle reasons during the timed region: ONE `nvidia-smi -lms 100` child (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, enabled=True):
        self.index, self.enabled, self.proc, self.lines = index, enabled, None, []

    def __enter__(self):
        if self.enabled:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                time.sleep(0.25)   # let the first sample land before the timed region starts
            except Exception:  # noqa: BLE001
                self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
                self.lines = [ln for ln in out.splitlines() if ln.strip()]
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        mhz, mx, reasons = [], None, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            try:
                mhz.append(float(p[0])); mx = float(p[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(self.NAMES, p[2:]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}


def cpu_baseline(scene, label, seconds_target=12.0):
    """The reference's device functions compiled for the host (oracle/_ref) on a strided pixel sample."""
    from oracle import pyoracle as po

    kind = "reference" if po.have_ref_host() else "port"
    if kind == "port" and not po.have_oracle():
        po.build()
    cores = os.cpu_count() or 1
    w, h = scene.width, scene.height
    render = (lambda px: po.ref_host_render(scene, px=px, threads=cores)) if kind == "reference" else \
        (lambda px: po.oracle_render(scene, px=px, threads=cores))

    def grid(nx, ny):   # a strided nx x ny grid of pixels of the frame
        xs = (np.arange(nx) * w) // nx
        ys = (np.arange(ny) * h) // ny
        return (ys[:, None] * w + xs[None, :]).reshape(-1).astype(np.uint64)

    px = grid(min(64, w), min(36, h))   # calibrate, then size the sample for ~seconds_target
    t = time.perf_counter(); render(px); dt = time.perf_counter() - t
    per_px = dt / len(px)
    n = int(min(w * h, max(len(px), seconds_target / max(per_px, 1e-9))))
    ny = max(1, int((n * h / w) ** 0.5)); nx = max(1, n // ny)
    px = grid(min(nx, w), min(ny, h))
    cnt = po.oracle_render(scene, px=px[:2048], threads=cores)["counters"]   # unique rays per pixel of the sample
    rays_per_px = (cnt["rays_primary"] + cnt["rays_reflect"] + cnt["rays_transmit"] + cnt["rays_shadow"]) / min(len(px), 2048)
    t = time.perf_counter(); render(px); dt = time.perf_counter() - t
    return {"value": rays_per_px * len(px) / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"{len(px)} px strided {min(nx, w)}x{min(ny, h)} grid of {label} ({dt:.1f} s, {rays_per_px:.2f} rays/px)"}


def ensure_library():
    """The product has no fallback: a snapshot without the built .so compiles it (nvcc is in the image) before anything runs.
    Under torchrun local rank 0 builds and the other ranks wait for the file."""
    lib_path = os.path.join(ROOT, "cutrace_b200", "lib", "libcutrace_b200.so")
    if os.path.exists(lib_path):
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        subprocess.run(["make", "-C", ROOT, "-j4", "lib"], check=True, stdout=sys.stderr)
    else:
        t0, last, stable = time.time(), -1, 0
        while stable < 3:   # present and unchanged for three seconds: the linker has finished writing it
            if time.time() - t0 > 900:
                raise RuntimeError(f"{lib_path} was not built")
            time.sleep(1.0)
            size = os.path.getsize(lib_path) if os.path.exists(lib_path) else -1
            stable = stable + 1 if (size > 0 and size == last) else 0
            last = size


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CUDA kernel rebuilt for sm_100a, rank 0 only.  Scene arrays come from the same
    fixture (pure numpy host code); this repo's CUDA library is never loaded in this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes as C

    from oracle import pyoracle as po

    scene, wl = load_workload(args.workload)
    base = {"impl": "reference", "metric": "Mrays/s (primary+secondary)", "unit": "Mrays/s", "higher_is_better": True}
    cfg = config_of(scene, wl)
    rays = RAYS_PER_FRAME[wl["name"]]
    if not po.have_ref_gpu():
        # the reference could not be compiled in the container: fall back to its host build / the port
        cb = cpu_baseline(scene, wl["label"], seconds_target=20.0)
        emit(dict(base, value=cb["value"], n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=None, scaling="strong",
                  vs_baseline=None, dtype="f32", data="synthetic", config=cfg, cpu_baseline=cb,
                  e2e={"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                  reference_arm="host build of the reference's device functions (no sm_100a rebuild available)"))
        return
    lib = po._load(po.REF_GPU_SO)
    fn = lib.cutrace_ref_gpu_render_e2e
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def ref_e2e(s, frames):
        """the reference's whole operator gpu::render<S,5,256>: managed allocs, launch + sync, 3*h row copies, host max scan"""
        n = s.width * s.height
        d, nm, c = np.empty(n, np.float32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        tot = C.c_float(); rms = C.c_float(); mx = C.c_float()
        desc = s.as_desc()
        out = []
        for i in range(1 + frames):
            rc = fn(C.byref(desc), 1e-3, d.ctypes.data, nm.ctypes.data, c.ctypes.data, C.byref(rms), C.byref(tot), C.byref(mx))
            if rc:
                raise RuntimeError(f"reference e2e failed: {rc}")
            if i:
                out.append(tot.value)
        return float(np.mean(out))

    with ClockSampler(0) as clk:
        ref = po.ref_gpu_render(scene, iters=args.steps, warmup=args.warmup)
        e2e = ref_e2e(scene, max(1, min(args.steps, 3)))
    ms = ref["render_ms"]
    # BASELINE configs 1-3 on the reference kernel (full frames) and config 5 on the oracle-side subset kernel
    extra = {}
    if not args.no_extra:
        for name in ("triangle", "spheres1080", "mirror1080"):
            if name == wl["name"]:
                continue
            s2, w2 = load_workload(name)
            r2 = po.ref_gpu_render(s2, iters=5, warmup=2)
            extra[name] = {"workload": w2["label"], "ms_per_frame": r2["render_ms"], "value": RAYS_PER_FRAME[name] / r2["render_ms"] / 1e3, "unit": "Mrays/s",
                           "e2e_ms_per_frame": ref_e2e(s2, 3)}
        if wl["name"] != "synthetic10m":
            s5, w5 = load_workload("synthetic10m")
            px = np.random.default_rng(0).choice(s5.width * s5.height, 4096, replace=False).astype(np.uint64)
            r5 = po.ref_gpu_render(s5, px=px)
            ms5 = r5["render_ms"] / 4096 * s5.width * s5.height
            extra["synthetic10m"] = {"workload": w5["label"], "ms_per_frame": ms5, "value": RAYS_PER_FRAME["synthetic10m"] / ms5 / 1e3, "unit": "Mrays/s",
                                     "extrapolated": f"brute force cannot render 33 M pixels: the reference's ray_cast / ray_color on a seeded 4096-pixel subset took "
                                                     f"{r5['render_ms']:.1f} ms, scaled linearly in the pixel count (4096 threads fill 16 of 148 SMs: an upper bound, up to ~9x high)"}
    cb = cpu_baseline(scene, wl["label"], seconds_target=10.0)
    import cutrace_b200._lib as product_lib
    from cutrace_b200.scene import _ARRAY_FIELDS

    scene_bytes = sum(getattr(scene, k).nbytes for k, _, _ in _ARRAY_FIELDS)
    n = scene.width * scene.height
    emit(dict(base, value=rays / ms / 1e3, n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms, scaling="strong",
              vs_baseline=None, dtype="f32", data="synthetic (reference scene geometry, fixture tests/golden/scenes)", config=cfg,
              launch="<<<w*h/256+1,256>>> render_kernel<S,5>, sm_100a rebuild", clocks=clk.summary(), gpu_launches=args.steps,
              e2e={"value": rays / e2e / 1e3, "unit": "Mrays/s", "ms_per_frame": e2e, "h2d_bytes_per_step": int(scene_bytes),
                   "d2h_bytes_per_step": int(28 * n), "what": "cutrace::gpu::render<S,5,256> total bracket (inc/kernel.hpp:88-126)"},
              extra_workloads=extra, cpu_baseline=cb, product_library_loaded=product_lib._lib is not None,
              reference_arm="reference CUDA kernel rebuilt for sm_100a (the comparator north_star names); cpu_baseline is the reference's device code compiled for the host"))


_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to fd 1 when
    NCCL_DEBUG is set on the box), so fd 1 is pointed at stderr for the rest of the process — library banners, child
    processes, stray prints — and the result line goes to a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        ensure_library()
        import cutrace_b200 as ct

        self.torch, self.dist, self.ct, self.args = torch, dist, ct, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.failed = False
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: cutrace_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        if self.world != args.gpus and self.rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={self.world}; using WORLD_SIZE", file=sys.stderr)
        # the ctx launches its trace chain on this (high-priority) stream; torch events see its kernels
        self.stream = torch.cuda.Stream(device=self.local_rank, priority=-1)
        torch.cuda.set_stream(self.stream)
        self.lib = ct._lib.load()
        # the per-frame "all tiles stored" barrier + max-depth reduction of the fused exchanges: shared host memory instead of an
        # NCCL all-reduce (--nccl-frame-barrier restores the latter)
        self.host_barrier = None
        if self.world > 1 and not args.nccl_frame_barrier:
            from cutrace_b200.distributed import HostBarrier

            self.host_barrier = HostBarrier(self.rank, self.world, self.local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return [float(x) for x in t]

    # ---- device-resident frames: K frames between two CUDA events on the launching stream ----
    def resident(self, scene, steps, warmup, sampler=False):
        from cutrace_b200.distributed import TileShardedRenderer

        torch, args = self.torch, self.args
        tsr = TileShardedRenderer(scene, rank=self.rank, world=self.world, device=self.local_rank, flags=args.flags, stream=self.stream.cuda_stream,
                                  exchange=args.exchange, host_barrier=self.host_barrier)

        host = {"render": 0.0, "gather": 0.0}

        def step():
            t0 = time.perf_counter()
            st = tsr.render()
            t1 = time.perf_counter()
            if self.world > 1:
                tsr.gather()
            host["render"] += t1 - t0
            host["gather"] += time.perf_counter() - t1
            return st

        for _ in range(warmup):
            st = step()
        self.barrier()
        host["render"] = host["gather"] = 0.0
        dev_ms = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(self.local_rank, enabled=(sampler and self.rank == 0)) as clk:
            if sampler:
                self.barrier()   # rank 0 spent 0.25 s starting the sampler: line the ranks up again before the timed region
            ev0.record(self.stream)
            for _ in range(steps):
                st = step()
                dev_ms.append(st["render_ms"])
            ev1.record(self.stream)
            self.barrier()
        ms_local = ev0.elapsed_time(ev1) / steps
        if args.verbose:
            print(f"[rank {self.rank}] step {ms_local:.3f} ms  render {np.mean(dev_ms):.3f}  rays {st['rays_total']}  px {st['local_pixels']}  "
                  f"scheduler {st['scheduler']}  host ms/step: render call {host['render'] / steps * 1e3:.3f} gather {host['gather'] / steps * 1e3:.3f}  "
                  f"phases {' '.join(f'{x:.3f}' for x in tsr.r.phase_ms())}", file=sys.stderr, flush=True)
        ms_step, render_dev_ms = self.reduce([ms_local, float(np.mean(dev_ms))], "max")
        rays, rays_shadow = self.reduce([st["rays_total"], st["rays_shadow"]], "sum")
        launches = int(st["kernel_launches"]) + (1 if self.world > 1 and self.rank == 0 and tsr.exchange == "gather" else 0)
        res = dict(ms_per_step=ms_step, render_device_ms=render_dev_ms, rays=rays, rays_shadow=rays_shadow, launches_per_step=launches,
                   scheduler={0: "wavefront: one launch per level and kind (CUDA graph)", 1: "wavefront: persistent frame kernel",
                              2: "pixel kernel (persistent, one thread per pixel path)"}[int(st["scheduler"])], exchange=tsr.exchange,
                   clocks=clk.summary() if sampler else None)
        # ---- the N-GPU frame against the same frame rendered by ONE GPU, bit for bit (pixels are independent in the reference,
        # /root/reference/inc/kernel.hpp:37-59: sharding changes who renders a pixel, not what) ----
        if self.world > 1 and not args.no_parity_check:
            res["parity_check"] = self.parity_check(scene, tsr, st)
        tsr.close()
        return res

    def parity_check(self, scene, tsr, st, host_frame=None):
        ct = self.ct
        ok = 1
        info = {}
        if self.rank == 0:
            if host_frame is not None:
                got = host_frame.as_dict()
            elif tsr.exchange == "gather":
                got = {k: getattr(tsr, "out_" + a).cpu().numpy() for k, a in (("depth", "depth"), ("normal", "normal"), ("color", "color"), ("hit_id", "id"))}
            else:
                got = tsr.r.download()
            # same scheduler as the shards used: the frame kernel and the per-level kernels are separate compilations (last-bit colours)
            forced = {0: ct.FLAG_LAUNCHES, 1: ct.FLAG_FRAME_KERNEL, 2: ct.FLAG_PIXEL_KERNEL}[int(st["scheduler"])]
            flags = (self.args.flags & ~(ct.FLAG_FRAME_KERNEL | ct.FLAG_LAUNCHES | ct.FLAG_PIXEL_KERNEL)) | forced
            with ct.Renderer(scene, device=self.local_rank, flags=flags) as r1:
                r1.render()
                one = r1.download()
            branching = scene.max_children() >= 2   # a material reflects AND transmits: float atomics, order-dependent last bits
            info = {"pixels": int(scene.width * scene.height), "world": self.world,
                    "source": "shared pinned host frame" if host_frame is not None else "rank 0's device frame"}
            for k in ("depth", "normal", "hit_id", "color"):
                a, b = np.asarray(got[k]).reshape(-1), np.asarray(one[k]).reshape(-1)
                same = bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
                if k == "color" and branching and not same:
                    same = bool(np.abs(a - b).max() < 1e-5)
                    info["color_note"] = "branching scene: float atomics, compared to 1e-5"
                info[k + "_equal"] = same
                if not same:
                    info[k + "_differing"] = int((a.view(np.uint32) != b.view(np.uint32)).sum())
                    ok = 0
            info["n_gpu_equals_1_gpu"] = bool(ok)
        ok = self.reduce([ok], "min")[0] >= 1   # every rank learns the verdict
        self.failed = self.failed or not ok
        return info

    # ---- per-kernel times: two extra frames with CUTRACE_FLAG_SERIALIZE (one stream, CUDA events around every kernel) ----
    def kernel_ms(self, scene):
        from cutrace_b200.distributed import TileShardedRenderer

        ser = TileShardedRenderer(scene, rank=self.rank, world=self.world, device=self.local_rank, flags=self.args.flags | self.ct.FLAG_SERIALIZE,
                                  stream=self.stream.cuda_stream, exchange="gather")
        ser.render()
        sst = ser.render()
        ser.close()
        shade, trace, serial = self.reduce([sst["shade_ms"], sst["trace_ms"], sst["render_ms"]], "max")
        return {"trace": trace, "shade": shade, "serialized_frame": serial,
                "note": "per-kernel times of the WAVEFRONT scheduler (CUTRACE_FLAG_SERIALIZE: one launch per level and kind on one stream), for comparison; the timed frames run the scheduler named in details.scheduler"}

    # ---- end to end: host scene in, host frame out, every step ----
    def e2e(self, scene, steps, check_parity=False, flags=None):
        import ctypes as C

        from cutrace_b200.distributed import SharedHostFrame, TileShardedRenderer
        from cutrace_b200.scene import _ARRAY_FIELDS

        ct, lib, args = self.ct, self.lib, self.args
        flags = args.flags if flags is None else flags
        n_px = scene.width * scene.height
        scene_bytes = sum(getattr(scene, k).nbytes for k, _, _ in _ARRAY_FIELDS if k != "obj_kind") + 64
        ms, parity = [], None
        # "inputs from pinned host memory": page-lock the big geometry arrays of the scene (a 10 M-triangle scene is 600 MB)
        locked = []
        for k, _, _ in _ARRAY_FIELDS:
            a = getattr(scene, k)
            if a.nbytes >= (1 << 20) and lib.cutrace_host_register(a.ctypes.data, a.nbytes) == 0:
                locked.append(a)
        if self.world == 1:
            pinned, ptrs = {}, []
            for k, m in (("depth", 1), ("normal", 3), ("color", 3)):
                p = lib.cutrace_host_alloc(n_px * m * 4)
                ptrs.append(p)
                pinned[k] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n_px * m,))
            for i in range(2 + steps):
                self.barrier()
                t0 = time.perf_counter()
                r = ct.Renderer(scene, device=self.local_rank, flags=flags, stream=self.stream.cuda_stream)
                md = C.c_float()
                ct._lib.check(lib.cutrace_render_download(r._ctx, pinned["depth"].ctypes.data, pinned["normal"].ctypes.data, pinned["color"].ctypes.data,
                                                          None, C.byref(md), None))
                r.close()
                self.barrier()
                if i >= 2:
                    ms.append((time.perf_counter() - t0) * 1e3)
            del pinned
            for p in ptrs:
                lib.cutrace_host_free(p)
            what = ("cutrace_upload_scene (H2D + BVH build: LBVH + SAH treelets) + cutrace_render_download (render; depth/normal D2H under the bounce levels, colour D2H at "
                    "the end) into pinned host buffers + cutrace_free, every frame")
            d2h = 28 * n_px
        else:
            frame = SharedHostFrame(scene.width, scene.height, self.rank, self.world, self.local_rank)
            for i in range(2 + steps):
                self.barrier()
                t0 = time.perf_counter()
                tr = TileShardedRenderer(scene, rank=self.rank, world=self.world, device=self.local_rank, flags=flags, stream=self.stream.cuda_stream,
                                         exchange="host", host_frame=frame, host_barrier=self.host_barrier)
                st = tr.render()                      # returns when this rank's tiles are in the host frame
                tr.max_depth(st["max_depth"])         # 1-float all-reduce (kernel.hpp:120-125); doubles as the "frame complete" barrier
                if i >= 2:
                    ms.append((time.perf_counter() - t0) * 1e3)
                if check_parity and i == 1 + steps:   # outside the timed part of the last frame
                    parity = self.parity_check(scene, tr, st, host_frame=frame)
                tr.close()
                self.barrier()
            frame.close()
            what = ("per rank: cutrace_upload_scene (H2D + BVH build: LBVH + SAH treelets) + cutrace_frame_attach(shared pinned host frame) + cutrace_render — every rank's "
                    "kernels store their tiles (G-buffer under the bounce levels, colour at the end) straight into ONE host frame over their own PCIe link — "
                    "+ max-depth all-reduce, every frame (cutrace_free of the frame's ctx follows outside the bracket)")
            d2h = 32 * n_px
            scene_bytes *= self.world
        for a in locked:
            lib.cutrace_host_unregister(a.ctypes.data)
        med, mean, mx = self.reduce([float(np.median(ms)), float(np.mean(ms)), float(np.max(ms))], "max")
        out = {"ms_per_frame": med, "ms_per_frame_mean": mean, "ms_per_frame_max": mx, "frames": len(ms), "statistic": "median",
               "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(d2h), "what": what}
        if parity is not None:
            out["parity_check"] = parity
        return out

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def compulsory_bytes_per_frame(n_px, rays_total, rays_shadow, n_lights, scene_bytes, pixel_kernel=True):
    """DRAM bytes a frame cannot avoid with this data layout.  Pixel kernel: the 32-byte frame record of every pixel out
    and the scene in once — a path lives in registers.  Wavefront schedulers: per traced ray a 32-byte ray record in (levels >= 1), a
    48-byte shade record and a 32-byte child ray out; per shaded hit 48 B in, 12 B level colour out and 12 B read again by the ordered
    sum; per pixel 20 B G-buffer, 4 B level count, 12 B final colour."""
    if pixel_kernel:
        return n_px * 32 + scene_bytes
    traced = rays_total - rays_shadow
    shaded = rays_shadow / max(1, n_lights)
    return traced * (32 + 48 + 32) - n_px * 32 + shaded * (48 + 12 + 12) + n_px * (20 + 4 + 12)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cutrace_b200", choices=["cutrace_b200", "reference"])
    ap.add_argument("--workload", default="bunny4k", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads (BASELINE configs 1-3 and 5)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--verbose", action="store_true", help="per-rank timing lines on stderr")
    ap.add_argument("--nccl-frame-barrier", action="store_true", help="N > 1: end every frame with the NCCL max-depth all-reduce instead of the shared-memory barrier")
    ap.add_argument("--exchange", default="peer", choices=["peer", "gather"],
                    help="N > 1, device-resident frames: peer = kernels store into rank 0's frame over NVLink (CUDA IPC), gather = NCCL gather + un-tile")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    b = Bench(args)
    scene, wl = load_workload(args.workload)
    n_px = scene.width * scene.height
    main_res = b.resident(scene, args.steps, args.warmup, sampler=True)
    kms = b.kernel_ms(scene)
    e2e = b.e2e(scene, max(5, min(args.steps, 20)), check_parity=not args.no_parity_check)
    # the same call with CUTRACE_FLAG_FAST_BUILD (LBVH only): what a one-frame-per-upload caller of a primitive-heavy scene would set
    e2e["fast_build_ms_per_frame"] = b.e2e(scene, 5, flags=args.flags | b.ct.FLAG_FAST_BUILD)["ms_per_frame"]
    rays, ms_step = main_res["rays"], main_res["ms_per_step"]

    extra = {}
    if not args.no_extra:
        for name in ("triangle", "spheres1080", "mirror1080", "synthetic10m"):
            if name == wl["name"]:
                continue
            s2, w2 = load_workload(name)
            big = name == "synthetic10m"
            r2 = b.resident(s2, 5 if big else 10, 3)
            e2 = b.e2e(s2, 3 if big else 5, check_parity=(big and not args.no_parity_check))
            e2f = b.e2e(s2, 3 if big else 5, flags=args.flags | b.ct.FLAG_FAST_BUILD)
            # ms_per_frame: K frames between two CUDA events, i.e. including cutrace_render's per-frame host synchronisation (what a caller
            # sees per frame; on a 20 x 20 frame that round trip is longer than the kernel).  render_device_ms: events around each frame's
            # kernel — the quantity the reference arm reports for its kernel (one event pair per launch, oracle/ref_gpu.cu).
            extra[name] = {"workload": w2["label"], "ms_per_frame": r2["ms_per_step"], "value": r2["rays"] / r2["ms_per_step"] / 1e3, "unit": "Mrays/s",
                           "rays_per_frame": int(r2["rays"]), "scheduler": r2["scheduler"], "render_device_ms": r2["render_device_ms"],
                           "e2e_ms_per_frame": e2["ms_per_frame"], "e2e_value": r2["rays"] / e2["ms_per_frame"] / 1e3,
                           "e2e_fast_build_ms_per_frame": e2f["ms_per_frame"],
                           "timed": "ms_per_frame: K frames between two CUDA events incl. cutrace_render's per-frame host synchronisation (and the rank "
                                    "barrier at N > 1); render_device_ms: events around each frame's kernel, the reference arm's per-launch quantity"}
            if "parity_check" in r2:
                extra[name]["parity_check"] = r2["parity_check"]
            if "parity_check" in e2:
                extra[name]["e2e_parity_check"] = e2["parity_check"]
            del s2

    if b.rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        clocks = main_res["clocks"]
        from cutrace_b200.scene import _ARRAY_FIELDS

        scene_bytes = sum(getattr(scene, k).nbytes for k, _, _ in _ARRAY_FIELDS if k != "obj_kind")
        comp = compulsory_bytes_per_frame(n_px, rays, main_res["rays_shadow"], scene.n_lights, scene_bytes,
                                          pixel_kernel=main_res["scheduler"].startswith("pixel"))
        achieved = comp / (ms_step * 1e-3) / 1e9
        # what ncu measured for one frame of this workload (committed: profiles/ncu_frame.json, written by tools/ncu_frame.py)
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_frame.json"))).get(wl["label"], {}).get(f"n{b.world}", {})
        except Exception:  # noqa: BLE001
            pass
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        sms = b.torch.cuda.get_device_properties(b.local_rank).multi_processor_count
        issue = None
        if ncu.get("thread_inst_per_frame"):
            lanes_peak = sms * 4 * 32 * sm_hz    # thread-instructions per second the SMs can issue
            issue = {"thread_inst_per_frame": ncu["thread_inst_per_frame"], "warp_inst_per_frame": ncu.get("warp_inst_per_frame"),
                     "active_lanes_per_warp_inst": ncu.get("active_lanes"), "frac": ncu["thread_inst_per_frame"] / (ms_step * 1e-3) / lanes_peak,
                     "peak": "SMs x 4 schedulers x 32 lanes x SM clock", "sms": sms, "sm_mhz": sm_hz / 1e6,
                     "source": "instruction counts: committed ncu capture of the same (deterministic) frame; time: this run"}
        bpr = survey_bytes_per_ray(n_primitives(scene))
        line = {
            "metric": "Mrays/s (primary+secondary)", "value": rays / ms_step / 1e3, "unit": "Mrays/s", "n_gpus": b.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (reference scene geometry from the committed fixture tests/golden/scenes; no image inputs)",
            "config": config_of(scene, wl),
            "details": {"parallelism": f"tiles{b.world}/{main_res['exchange']}" if b.world > 1 else "single", "scheduler": main_res["scheduler"],
                        "timed": "K frames between two CUDA events on the launching stream" + (
                            "" if b.world == 1 else ", incl. the per-frame rank barrier + max-depth reduction (" + ("shared host memory" if b.host_barrier is not None else "NCCL all-reduce") +
                            "; tiles are stored into rank 0's frame over NVLink by the kernels)"
                            if main_res["exchange"] == "peer" else ", incl. NCCL gather + un-tile"),
                        "render_device_ms": main_res["render_device_ms"], "rays_counted": int(rays)},
            "clocks": clocks,
            "e2e": dict(e2e, value=rays / e2e["ms_per_frame"] / 1e3, unit="Mrays/s"),
            "gpu_launches": main_res["launches_per_step"] * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "what": "compulsory DRAM bytes of one frame (pixel kernel: the 32-byte frame record per pixel + the scene once; wavefront: plus queue "
                                 "records and level colours — DESIGN.md 3) / ms_per_step",
                         "compulsory_bytes_per_frame": comp, "traffic": ncu.get("dram_bytes_per_frame"),
                         "traffic_unit": "bytes per frame, all kernels (dram__bytes_read.sum + dram__bytes_write.sum)",
                         "binding_resource": "issue slots x active lanes (the BVH is cache-resident; DRAM runs at `frac` of its peak)", "issue": issue,
                         "kernels": ncu.get("kernels"),
                         "survey_model": {"bytes_per_ray": bpr, "achieved": rays * bpr / (ms_step * 1e-3) / 1e9, "frac": rays * bpr / (ms_step * 1e-3) / 1e9 / peak,
                                          "note": "SURVEY 8d's model (one 64-byte node per level of a balanced tree per ray, as if from HBM): these bytes are served "
                                                  "from shared memory / L1, so this is NOT a bandwidth measurement and can exceed 1"},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "kernel_ms": kms,
            "extra_workloads": extra,
        }
        if "parity_check" in main_res:
            line["parity_check"] = main_res["parity_check"]
        if not args.no_cpu_baseline and b.world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(scene, wl["label"])
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "absent", "sample": f"failed: {e}"}
        emit(line)
    failed = b.failed
    b.finish()
    if failed:
        raise SystemExit("parity_check failed: the N-GPU frame differs from the 1-GPU frame")


if __name__ == "__main__":
    main()
