"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against
  (1) the committed goldens produced by the reference's own source (tests/golden/*.npz),
  (2) the C oracle on the same seeded inputs,
  (3) the reference's own CUDA kernel rebuilt for sm_100a (oracle/_ref/libcutrace_ref_gpu.so) when it
      was built in the container,
plus size-independent properties at BASELINE.json's full sizes.  Run with `-m gpu` on a B200.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, load_golden_scene
from parity import assert_parity, assert_parity_mesh_scene, assert_same_frame, compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ct():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (the product has no CPU path)")
    import subprocess

    import cutrace_b200
    from conftest import ROOT

    if not os.path.exists(cutrace_b200._lib.LIB_PATH):   # snapshot without the built .so: compile it, never fall back
        subprocess.run(["make", "-C", ROOT, "-j4", "lib"], check=True, stdout=subprocess.DEVNULL)
    cutrace_b200._lib.load()
    return cutrace_b200


def gpu_render(ct, scene, **kw):
    flags = kw.pop("flags", 0) | ct.FLAG_VALIDATE_BVH
    with ct.Renderer(scene, flags=flags, **kw) as r:
        st = r.render()
        out = r.download()
    return out, st


# ---- (1) reference goldens -------------------------------------------------------------------------
SCHEDULERS = ["launches", "frame", "pixel"]   # CUTRACE_SCHEDULER: per-level launches, the persistent frame kernel, the per-pixel kernel


@pytest.mark.parametrize("sched", SCHEDULERS)
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
@pytest.mark.parametrize("mode", ["smem", "global", "brute"])
def test_reference_goldens(ct, name, mode, sched, monkeypatch):
    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    g = dict(np.load(os.path.join(GOLDEN, GOLDEN_CASES[name])))
    s = load_golden_scene(name).with_resolution(int(g["width"]), int(g["height"]))
    flags = {"smem": 0, "global": ct.FLAG_NO_SMEM_TOP, "brute": ct.FLAG_BRUTE_FORCE}[mode]
    out, st = gpu_render(ct, s, flags=flags)
    m = compare(out, g, s.width, s.height)
    assert_parity(m, f"{name}/{mode} vs reference golden", oracle_is_host=True)
    assert st["max_depth"] == pytest.approx(float(np.max(g["depth"][np.isfinite(g["depth"])], initial=0.0)), rel=1e-4)


# ---- (2) C oracle on seeded inputs -------------------------------------------------------------------
@pytest.mark.parametrize("sched", SCHEDULERS)
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("translucent", [False, True])
def test_random_soup_vs_oracle(ct, oracle, seed, translucent, sched, monkeypatch):
    from cutrace_b200 import synth

    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    s = synth.random_soup(n_tri=300, n_sph=6, n_planes=2, n_lights=3, width=128, height=96, seed=seed, translucent=translucent,
                          duplicates=True)
    out, st = gpu_render(ct, s)
    ref = oracle.oracle_render(s)
    assert_parity(compare(out, ref, s.width, s.height), f"soup seed={seed} translucent={translucent} vs host oracle",
                  oracle_is_host=True, chaotic=True)
    c = ref["counters"]
    # same unique-ray bookkeeping (edge pixels and chaotic second bounces may shift a few rays)
    for k in ("rays_primary", "rays_reflect", "rays_transmit", "rays_shadow", "shadow_casts"):
        assert abs(st[k] - c[k]) <= max(8, 0.005 * c[k]), (k, st[k], c[k])
    if oracle.have_ref_gpu():   # the tight yardstick: the reference's own kernel, same FMA contraction
        gref = oracle.ref_gpu_render(s)
        m = compare(out, gref, s.width, s.height)
        assert_parity(m, f"soup seed={seed} translucent={translucent} vs reference sm_100a kernel")


def _empty_like(s, **kw):
    import copy

    e = copy.copy(s)
    for k, v in kw.items():
        setattr(e, k, v)
    e.__post_init__()
    return e


@pytest.mark.parametrize("sched", SCHEDULERS)
def test_edge_cases_vs_oracle(ct, oracle, sched, monkeypatch):
    from cutrace_b200 import synth

    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    base = synth.random_soup(n_tri=8, n_sph=1, n_planes=1, n_lights=2, width=97, height=61, seed=5)
    z3, zu = np.zeros((0, 3), np.float32), np.zeros(0, np.uint32)
    cases = {
        "odd resolution, tiny scene (one leaf)": synth.random_soup(n_tri=3, n_sph=0, n_planes=0, n_lights=1, width=97, height=61, seed=7),
        "single primitive": synth.random_soup(n_tri=1, n_sph=0, n_planes=0, n_lights=1, width=33, height=31, seed=8),
        "spheres only": synth.random_soup(n_tri=0, n_sph=9, n_planes=0, n_lights=2, width=64, height=48, seed=9),
        "planes only": synth.random_soup(n_tri=0, n_sph=0, n_planes=2, n_lights=2, width=64, height=48, seed=10),
        "no lights": synth.random_soup(n_tri=50, n_sph=2, n_planes=1, n_lights=0, width=64, height=48, seed=11),
        "1 x 1 frame": base.with_resolution(1, 1),
        "wide frame": base.with_resolution(300, 2),
    }
    # empty scene: every pixel is a miss
    empty = _empty_like(base, tri_p1=z3, tri_p2=z3, tri_p3=z3, tri_object=zu, sph_center=z3, sph_radius=np.zeros(0, np.float32),
                        sph_object=zu, pl_point=z3, pl_normal=z3, pl_object=zu, obj_material=zu, obj_kind=zu)
    cases["empty scene"] = empty
    # degenerate (zero-area) triangles are kept by the reference loader and simply never hit
    deg = synth.random_soup(n_tri=40, n_sph=0, n_planes=1, n_lights=1, width=64, height=48, seed=12)
    deg.tri_p2[::3] = deg.tri_p1[::3]
    cases["degenerate triangles"] = deg
    # many coincident centroids: Morton duplicates
    dup = synth.random_soup(n_tri=64, n_sph=0, n_planes=0, n_lights=1, width=64, height=48, seed=13)
    dup.tri_p1[:] = dup.tri_p1[0]; dup.tri_p2[:] = dup.tri_p2[0]; dup.tri_p3[:] = dup.tri_p3[0]
    cases["64 identical triangles (tie -> lowest object, first in file order)"] = dup
    for what, s in cases.items():
        out, st = gpu_render(ct, s)
        ref = oracle.oracle_render(s)
        assert_parity(compare(out, ref, s.width, s.height), what, oracle_is_host=True, chaotic=True)
        if oracle.have_ref_gpu() and what != "empty scene":   # the tight yardstick: same FMA contraction, same libm
            gref = oracle.ref_gpu_render(s)
            assert_parity(compare(out, gref, s.width, s.height), what + " vs reference sm_100a kernel")
        if what == "empty scene":
            assert np.all(out["hit_id"] == 0xFFFFFFFF) and np.all(np.isposinf(out["depth"])) and not out["color"].any()
            assert st["max_depth"] == 0.0


@pytest.mark.parametrize("sched", SCHEDULERS)
def test_bounce_budget_and_fudge(ct, oracle, sched, monkeypatch):
    """bounces = 0..3 and a different fudge go through the same code as the reference's template arguments."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    s = load_golden_scene("sphere_plane").with_resolution(96, 54)
    for b in (0, 1, 3):
        out, _ = gpu_render(ct, s, bounces=b)
        ref = oracle.oracle_render(s, bounces=b)
        assert_parity(compare(out, ref, s.width, s.height), f"bounces={b}", oracle_is_host=True)
    out, _ = gpu_render(ct, s, fudge=0.05)
    ref = oracle.oracle_render(s, fudge=0.05)
    assert_parity(compare(out, ref, s.width, s.height), "fudge=0.05", oracle_is_host=True)


# ---- (3) the reference's own kernel rebuilt for sm_100a ---------------------------------------------
@pytest.mark.parametrize("name,res", [("triangle", None), ("sphere_plane", (1920, 1080)), ("mirror", (1920, 1080)), ("bunny", (480, 270)),
                                      ("bunny", (3840, 2160))])
def test_vs_reference_cuda_kernel(ct, oracle, name, res):
    _vs_reference_cuda_kernel(ct, oracle, name, res)


@pytest.mark.parametrize("sched", ["frame", "pixel"])
@pytest.mark.parametrize("name,res", [("triangle", None), ("sphere_plane", (1920, 1080)), ("mirror", (640, 360)), ("bunny", (480, 270))])
def test_vs_reference_cuda_kernel_other_schedulers(ct, oracle, name, res, sched, monkeypatch):
    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    _vs_reference_cuda_kernel(ct, oracle, name, res)


def _vs_reference_cuda_kernel(ct, oracle, name, res):
    """BASELINE configs 1-4 at their full resolutions, FULL frames, against the reference's own kernel launched like
    /root/reference/inc/kernel.hpp:103-106 (bunny.json at 4K costs the reference ~0.9 s)."""
    if not oracle.have_ref_gpu():
        pytest.skip("oracle/_ref/libcutrace_ref_gpu.so was not built (reference tree absent at build time)")
    s = load_golden_scene(name)
    if res:
        s = s.with_resolution(*res)
    ref = oracle.ref_gpu_render(s)
    out, st = gpu_render(ct, s)
    m = compare(out, ref, s.width, s.height)
    assert_parity(m, f"{name} {s.width}x{s.height} vs reference sm_100a kernel")


def test_bunny_4k_subset_vs_reference_cuda_kernel(ct, oracle):
    """BASELINE config 4 at full size: 4096 seeded pixels of the 3840x2160 frame through the reference's
    ray_cast/ray_color (oracle-side subset kernel) against the same pixels of the full render."""
    if not oracle.have_ref_gpu():
        pytest.skip("reference CUDA oracle not built")
    s = load_golden_scene("bunny").with_resolution(3840, 2160)
    px = np.random.default_rng(0).choice(3840 * 2160, 4096, replace=False).astype(np.uint64)
    ref = oracle.ref_gpu_render(s, px=px)
    out, st = gpu_render(ct, s)
    sub = {k: out[k][px.astype(np.int64)] for k in ("depth", "normal", "color", "hit_id")}
    assert_parity(compare(sub, ref), "bunny 4K subset vs reference sm_100a kernel")
    # every ray hits in the closed box: 30 unique rays per pixel minus the rays that leak through cracks
    assert 29.9 * 3840 * 2160 <= st["rays_total"] <= 30 * 3840 * 2160
    # four shadow rays per shaded hit; a ray that leaks through a crack between triangles shades nothing
    assert st["rays_shadow"] % 4 == 0
    assert 0.9999 * 4 * (st["rays_primary"] + st["rays_reflect"]) <= st["rays_shadow"] <= 4 * (st["rays_primary"] + st["rays_reflect"])


def test_config5_full_size_subset_vs_reference_cuda_kernel(ct, oracle):
    """BASELINE config 5 at FULL size: the 106 x 106 instance hall (10,112,400 triangles) at 7680x4320 rendered
    completely by this path; a seeded 4096-pixel subset of it against the reference's ray_cast / ray_color run by the
    oracle-side subset kernel (brute force cannot render 33 M pixels: SURVEY.md 8d "Reference on config 5")."""
    if not oracle.have_ref_gpu():
        pytest.skip("reference CUDA oracle not built")
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=106, width=7680, height=4320)
    assert s.n_triangles == 10_112_400
    px = np.random.default_rng(0).choice(s.width * s.height, 4096, replace=False).astype(np.uint64)
    ref = oracle.ref_gpu_render(s, px=px)
    with ct.Renderer(s) as r:      # no FLAG_VALIDATE_BVH here: the validation walk of 3.6 M nodes is covered at G = 24
        st = r.render()
        out = r.download(want=("depth", "normal", "color", "hit_id"))
    sub = {k: out[k][px.astype(np.int64)] for k in ("depth", "normal", "color", "hit_id")}
    m = compare(sub, ref)
    assert_parity_mesh_scene(m, "config 5 (10.1 M triangles @ 7680x4320) subset vs reference sm_100a kernel")
    assert m["id_mismatch"] == 0, m
    assert np.all(out["hit_id"] != 0xFFFFFFFF)        # closed hall: every primary ray hits
    assert st["rays_primary"] == 7680 * 4320 and st["rays_shadow"] % 3 == 0
    import bench

    assert st["rays_total"] == bench.RAYS_PER_FRAME["synthetic10m"]


def test_grid_scene_full_frame_vs_reference_kernel(ct, oracle):
    """config 5 in small (6 x 6 instances, 32,400 triangles, mirrors) as a FULL frame against the reference's real
    render_kernel (brute force is affordable at this size), and the oracle-side subset kernel against that same kernel on
    the same pixels — the subset kernel is what the full-size config-5 test has to rely on."""
    if not oracle.have_ref_gpu():
        pytest.skip("reference CUDA oracle not built")
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=6, width=480, height=270)
    ref = oracle.ref_gpu_render(s)
    out, st = gpu_render(ct, s)
    m = compare(out, ref, s.width, s.height)
    assert_parity_mesh_scene(m, "6x6 grid full frame vs reference sm_100a kernel")
    brute, _ = gpu_render(ct, s, flags=ct.FLAG_BRUTE_FORCE)       # the BVH walk never changes which triangle wins or its t
    assert_same_frame(out, brute, "LBVH walk vs brute-force loop")
    px = np.random.default_rng(3).choice(s.width * s.height, 4096, replace=False).astype(np.uint64)
    sub = oracle.ref_gpu_render(s, px=px)
    full_sub = {k: ref[k][px.astype(np.int64)] for k in ("depth", "normal", "color", "hit_id")}
    ms = compare(sub, full_sub)
    print("subset kernel vs render_kernel:", ms, " this path vs render_kernel:", m)
    assert ms["id_mismatch"] == 0 and ms["depth_max_rel"] <= 1e-4, ms


def test_mesh_deviations_documented_in_design_8(ct, oracle):
    """DESIGN.md 8 lists two places where the LBVH walk is not the reference's mesh::intersect; both are constructed
    here and compared with the reference's own kernel.
    (1) per-mesh AABB pre-test (inc/default_schema.hpp:99-114): a FLAT mesh (all vertices in the plane z = 0, so its
        box has zero thickness) seen by rays that travel inside that plane and by axis-parallel rays through box faces.
        The reference's slab test produces 0 * inf = NaN there; fminf/fmaxf drop the NaN, so the mesh is still entered
        and the Cramer test decides (alpha = 0 -> non-finite -> miss).  Same image expected.
    (2) a mesh whose nearest triangle sits at exactly t == min_dist is rejected as a WHOLE by the reference
        (mesh::intersect keeps the closest triangle, ray_cast then drops it); here only that triangle is dropped.
        Constructed with fudge = 1.0 and an axis-aligned quad at distance exactly 1.0 in front of a second quad of the
        same mesh: a measure-zero case in which the two paths are allowed to differ, recorded rather than asserted equal."""
    if not oracle.have_ref_gpu():
        pytest.skip("reference CUDA oracle not built")
    from cutrace_b200.scene import FlatScene, LIGHT_POINT, OBJ_MESH

    def scene(tris, eye, look, fudge_scene=None, width=64, height=48):
        tris = np.asarray(tris, np.float32)
        fwd, right, up = ct.look_at(np.asarray(eye, np.float32), [0, 1, 0], np.asarray(look, np.float32))
        return FlatScene(cam_pos=np.asarray(eye, np.float32), cam_up=up, cam_forward=fwd, cam_right=right, ambient=0.1, width=width,
                         height=height, tri_p1=tris[:, 0], tri_p2=tris[:, 1], tri_p3=tris[:, 2], tri_object=np.zeros(len(tris), np.uint32),
                         obj_material=np.zeros(1, np.uint32), obj_kind=np.asarray([OBJ_MESH], np.uint32),
                         mat_color=np.asarray([[0.8, 0.5, 0.2]], np.float32), mat_specular=np.asarray([0.3], np.float32),
                         mat_reflect=np.asarray([0.2], np.float32), mat_phong=np.asarray([32], np.float32),
                         mat_transparency=np.zeros(1, np.float32), light_kind=np.asarray([LIGHT_POINT], np.uint32),
                         light_vec=np.asarray([[1, 3, -4]], np.float32), light_color=np.ones((1, 3), np.float32))

    flat = [[[-1, -1, 0], [1, -1, 0], [1, 1, 0]], [[-1, -1, 0], [1, 1, 0], [-1, 1, 0]], [[1, -1, 0], [3, -1, 0], [3, 1, 0]]]
    views = {
        "flat mesh, camera IN its plane (grazing: every ray with d.z = 0 lies in the box's zero-thickness slab)": ([-4, 0, 0], [0, 0, 0]),
        "flat mesh, axis-parallel central ray": ([0, 0, -3], [0, 0, 0]),
        "flat mesh, camera on the box's x = 3 face plane": ([3, 0, -3], [3, 0, 0]),
        "flat mesh, oblique": ([2, 1.5, -3], [0.5, 0, 0]),
    }
    for what, (eye, look) in views.items():
        s = scene(flat, eye, look)
        out, _ = gpu_render(ct, s)
        gref = oracle.ref_gpu_render(s)
        m = compare(out, gref, s.width, s.height)
        assert_parity(m, what)
        assert m["id_mismatch"] == 0, (what, m)
    # (2) t == min_dist: camera at z = -1 looking down +z at two parallel quads of ONE mesh at z = 0 (t = 1 on the central
    # axis-parallel ray only) and z = 1; fudge = 1.0
    quads = [[[-1, -1, 0], [1, -1, 0], [1, 1, 0]], [[-1, -1, 0], [1, 1, 0], [-1, 1, 0]],
             [[-2, -2, 1], [2, -2, 1], [2, 2, 1]], [[-2, -2, 1], [2, 2, 1], [-2, 2, 1]]]
    s = scene(quads, [0, 0, -1], [0, 0, 0], width=65, height=49)
    with ct.Renderer(s, fudge=1.0) as r:
        r.render()
        out = r.download()
    gref = oracle.ref_gpu_render(s, fudge=1.0)
    m = compare(out, gref, s.width, s.height)
    # pixels whose nearest triangle is at t <= 1 exactly may differ (reference: whole mesh missed; here: the far quad is hit)
    assert m["id_mismatch"] <= 4 and m["sentinel_mismatch"] <= 4, m
    same = out["hit_id"] == gref["hit_id"]
    assert np.array_equal(out["depth"][same].view(np.uint32), gref["depth"][same].view(np.uint32))


def test_closed_mirror_box_at_15_bounces_does_not_overflow_the_queues(ct, oracle):
    """queue capacity.  A closed box of 0.999 mirrors at the maximum bounce budget and an odd resolution keeps
    every ray alive for 16 levels, so the retired slot-block tails (holes) compound level after level; the kernels
    check every reservation against the queue capacity (FrameCounters.overflow -> CUTRACE_ERR_INTERNAL) and the
    capacity carries one slot block per producer warp AND level.  The frame must render, twice with the same bits."""
    from cutrace_b200.scene import FlatScene, LIGHT_POINT, OBJ_PLANE

    eye = np.asarray([0.1, 0.2, -0.3], np.float32)
    fwd, right, up = ct.look_at(eye, [0, 1, 0], np.asarray([0.3, 0.1, 1.0], np.float32))
    pts = np.asarray([[0, -1, 0], [0, 1, 0], [-1, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1]], np.float32)
    s = FlatScene(cam_pos=eye, cam_up=up, cam_forward=fwd, cam_right=right, ambient=0.1, width=1021, height=577,
                  pl_point=pts, pl_normal=-pts, pl_object=np.arange(6, dtype=np.uint32), obj_material=np.zeros(6, np.uint32),
                  obj_kind=np.full(6, OBJ_PLANE, np.uint32), mat_color=np.asarray([[0.9, 0.8, 0.7]], np.float32),
                  mat_specular=np.asarray([0.3], np.float32), mat_reflect=np.asarray([0.999], np.float32),
                  mat_phong=np.asarray([50], np.float32), mat_transparency=np.zeros(1, np.float32),
                  light_kind=np.asarray([LIGHT_POINT], np.uint32), light_vec=np.asarray([[0.2, 0.5, 0.1]], np.float32),
                  light_color=np.ones((1, 3), np.float32))
    with ct.Renderer(s, bounces=15) as r:
        st = r.render()
        a = r.download()
        st2 = r.render()
        b = r.download()
    n = s.width * s.height
    assert st["rays_primary"] == n and st["rays_reflect"] == 15 * n and st["rays_shadow"] == 16 * n
    assert st2["rays_total"] == st["rays_total"]
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    small = s.with_resolution(101, 57)
    out, _ = gpu_render(ct, small, bounces=15)
    ref = oracle.oracle_render(small, bounces=15)
    assert_parity(compare(out, ref, small.width, small.height), "mirror box, 15 bounces", oracle_is_host=True, chaotic=True)


def test_bench_ray_constants(ct):
    """bench.py's reference arm takes the unique-ray numerator from constants (it must not run this repo's renderer): they are what
    the renderer counts on the same frames."""
    import bench

    for name in ("triangle", "spheres1080", "mirror1080", "bunny4k"):
        scene, wl = bench.load_workload(name)
        for flags in (0, ct.FLAG_LAUNCHES, ct.FLAG_FRAME_KERNEL, ct.FLAG_PIXEL_KERNEL):
            with ct.Renderer(scene, flags=flags) as r:
                st = r.render()
            assert st["rays_total"] == bench.RAYS_PER_FRAME[name], (name, flags, st["rays_total"])


# ---- size-independent properties at full size ----------------------------------------------------------
def test_full_size_properties_bunny_4k(ct):
    s = load_golden_scene("bunny").with_resolution(3840, 2160)
    with ct.Renderer(s) as r:
        st1 = r.render()
        a = r.download()
        st2 = r.render()
        b = r.download()
    # idempotence: a non-branching scene has one ray per pixel and level -> bit-identical re-render
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert st1["rays_total"] == st2["rays_total"]
    assert np.all(a["hit_id"] != 0xFFFFFFFF)           # closed box: every primary ray hits
    assert np.all(np.isfinite(a["color"])) and a["color"].min() >= 0.0
    assert st1["max_depth"] == pytest.approx(float(a["depth"].max()), rel=0, abs=0)
    # horizontal mirror symmetry of the id image is NOT expected (camera is off-axis); instead check the
    # down-sampled frame against a directly rendered low-res frame (pixel-corner sampling: (2x,2y) of
    # the 4K grid is the same ray as (x,y) of the 1920x1080 grid)
    lo = s.with_resolution(1920, 1080)
    with ct.Renderer(lo) as r:
        r.render()
        c = r.download()
    a_ds = {k: a[k].reshape(2160, 3840, -1)[::2, ::2].reshape(1920 * 1080, -1).squeeze() for k in ("depth", "normal", "color", "hit_id")}
    m = compare(a_ds, c, 1920, 1080)
    assert m["id_mismatch"] == 0 and m["depth_max_rel"] == 0.0 and m["color_max_abs"] <= 1e-6, m


@pytest.mark.parametrize("sched", SCHEDULERS)
def test_tile_sharding_equals_single_ctx(ct, monkeypatch, sched):
    """world=3 interleaved tile shards rendered by three ctxs (one GPU) and stitched by cutrace_download
    are bit-identical to the unsharded frame — the multi-GPU path changes who renders a pixel, not what."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    s = load_golden_scene("mirror").with_resolution(333, 205)
    full, st = gpu_render(ct, s)
    acc = None
    rays = 0
    for rank in range(3):
        part, pst = gpu_render(ct, s, tile_rank=rank, tile_world=3)
        rays += pst["rays_total"]
        if acc is None:
            acc = {k: v.copy() for k, v in part.items() if k != "max_depth"}
        else:
            own = part["hit_id"] != 0xFFFFFFFF   # mirror.json is a closed box: foreign tiles read as misses
            for k in acc:
                acc[k][own] = part[k][own]
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(acc[k].view(np.uint32), full[k].view(np.uint32)), k
    assert rays == st["rays_total"]


@pytest.mark.parametrize("sched", SCHEDULERS)
def test_peer_frame_stores_equal_single_ctx(ct, monkeypatch, sched):
    """The gather-free multi-GPU path on one GPU: three sharded ctxs store their tiles straight into rank 0's row-major
    frame (cutrace_frame_attach = the in-process form of cutrace_frame_ipc_import); the assembled frame is bit-identical
    to the unsharded render."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", sched)
    for name, res in (("mirror", (333, 205)), ("sphere_plane", (200, 120))):
        s = load_golden_scene(name).with_resolution(*res)
        full, st = gpu_render(ct, s)
        rs = [ct.Renderer(s, tile_rank=r, tile_world=3) for r in range(3)]
        rs[0].frame_ipc_export()                 # gives rank 0 an exportable row-major frame of its own
        block = rs[0].frame_device()[0]
        for r in rs[1:]:
            r.frame_attach(block)
        rays = sum(r.render()["rays_total"] for r in rs)
        out = rs[0].download()
        with pytest.raises(ct.CutraceError):     # a ctx that renders into somebody else's frame has nothing to download
            rs[1].download()
        for k in ("depth", "normal", "hit_id"):
            assert np.array_equal(out[k].view(np.uint32), full[k].view(np.uint32)), (name, k)
        if name == "mirror":
            assert np.array_equal(out["color"].view(np.uint32), full["color"].view(np.uint32))
        else:
            assert np.abs(out["color"] - full["color"]).max() < 1e-5     # branching scene: float atomics
        assert rays == st["rays_total"]
        for r in rs:
            r.close()


def test_untile_of_gathered_rank_buffers(ct):
    """cutrace_untile_device on a rank-major concatenation of tile-major buffers (what the NCCL gather
    produces) against the host re-implementation in cutrace_b200.distributed.untile_host."""
    import torch

    from cutrace_b200.distributed import TileShardedRenderer, untile_host

    s = load_golden_scene("sphere_plane").with_resolution(200, 120)
    world = 2
    rs = [TileShardedRenderer(s, rank=r, world=world, device=0, exchange="gather") for r in range(world)]
    for r in rs:
        r.render()
    npx = rs[0].npx
    g_depth = torch.cat([r.depth for r in rs]); g_normal = torch.cat([r.normal for r in rs])
    g_color = torch.cat([r.color for r in rs]); g_id = torch.cat([r.hit_id for r in rs])
    n = s.width * s.height
    out_d = torch.empty(n, device="cuda"); out_c = torch.empty(3 * n, device="cuda")
    torch.cuda.synchronize()
    rs[0].r.untile_device(world, g_depth.data_ptr(), g_normal.data_ptr(), g_color.data_ptr(), g_id.data_ptr(), npx,
                          out_d.data_ptr(), None, out_c.data_ptr(), None)
    want_d = untile_host([r.depth.cpu().numpy().reshape(-1, 1) for r in rs], s.width, s.height, 1)
    want_c = untile_host([r.color.cpu().numpy().reshape(-1, 3) for r in rs], s.width, s.height, 3)
    assert np.array_equal(out_d.cpu().numpy().view(np.uint32), want_d.reshape(-1).view(np.uint32))
    assert np.array_equal(out_c.cpu().numpy().view(np.uint32), want_c.reshape(-1).view(np.uint32))
    full, _ = gpu_render(ct, s)
    assert np.array_equal(full["depth"].view(np.uint32), want_d.reshape(-1).view(np.uint32))
    for r in rs:
        r.close()


# ---- components ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 2048, 2049, 100_003, 1_500_000])
def test_radix_sort_is_a_stable_sort(ct, n):
    lib = ct._lib.load()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**63, n, dtype=np.uint64)
    if n > 100:
        keys[rng.integers(0, n, n // 3)] = keys[0]        # many duplicates: stability matters
        keys[rng.integers(0, n, n // 5)] &= np.uint64(0xFF)  # small keys: passes with a single live digit
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = keys.copy(), vals.copy()
    ct._lib.check(lib.cutrace_debug_radix_sort(k2.ctypes.data, v2.ctypes.data, n, 0))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])


def test_phong_pow_floor(ct):
    """The shading code skips powf(x, e) for x below 0.98 * 2^(-152/e) (the power underflows to +0 there): the device's own powf must
    agree bit for bit on both sides of the floor, for Phong exponents from 1 to FLT_MAX and for the exponents that get no floor."""
    lib = ct._lib.load()
    rng = np.random.default_rng(5)
    es = np.array([1.0, 1.0000001, 1.5, 2, 3, 10, 32, 50, 151, 152, 153, 200, 500, 1e3, 1e4, 1e6, 1e12, 1e30, 3e38, 3.4e38,
                   0.0, 0.5, 0.999, -1.0, -200.0, np.inf, -np.inf, np.nan], dtype=np.float32)
    xs, ee = [], []
    for e in es:
        with np.errstate(all="ignore"):
            fl = np.float32(0.98) * np.exp2(np.float32(-152.0) / np.float32(e)).astype(np.float32)
        fl = np.float32(fl) if np.isfinite(fl) else np.float32(1.0)
        around = [np.nextafter(fl, np.float32(0)), fl, np.nextafter(fl, np.float32(2)), fl * np.float32(0.999), fl * np.float32(1.001), fl * np.float32(0.5),
                  fl * np.float32(1.03), np.float32(0), np.float32(1), np.float32(1e-45), np.float32(1e-30), np.float32(0.5), np.float32(0.999999)]
        pts = np.concatenate([np.array(around, dtype=np.float32), rng.random(4000, dtype=np.float32),
                              (fl * (np.float32(0.9) + np.float32(0.15) * rng.random(4000, dtype=np.float32))).astype(np.float32)])
        xs.append(pts); ee.append(np.full(pts.shape, e, dtype=np.float32))
    x, e = np.concatenate(xs), np.concatenate(ee)
    a, b = np.empty_like(x), np.empty_like(x)
    ct._lib.check(lib.cutrace_debug_phong_pow(x.ctypes.data, e.ctypes.data, a.ctypes.data, b.ctypes.data, x.size, 0))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{int((a.view(np.uint32) != b.view(np.uint32)).sum())} of {x.size} differ"
    skipped = (x < (np.float32(0.98) * np.exp2(np.float32(-152.0) / e))) & (e >= 1) & (e <= np.float32(3e38))
    assert skipped.sum() > 10_000 and not a[skipped].any()       # the shortcut was exercised, and powf is +0 there


def test_bvh_validates_on_a_large_scene(ct):
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=24, width=256, height=144)   # 518,400 triangles, 579 objects
    for leaf in (1, 4, 8):
        with ct.Renderer(s, flags=ct.FLAG_VALIDATE_BVH, leaf_size=leaf) as r:
            r.validate_bvh()
            st = r.render()
            assert st["bvh_depth"] < 62 and st["bvh_nodes"] > 0


def test_fast_build_renders_the_same_frame_as_the_sah_treelet_build(ct):
    """CUTRACE_FLAG_FAST_BUILD keeps the LBVH topology, the default rebuilds every <=1024-primitive subtree with sweep SAH: two
    different trees over the same primitives — closest hits (ties broken by primitive id) and any-hit shadow results cannot differ,
    so the frames are bit-identical; the SAH tree must be the shallower-to-trace one (fewer nodes is not required, a valid tree is)."""
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    for s in (load_golden_scene("bunny").with_resolution(480, 270), synth.grid_scene(meshes, grid=6, width=320, height=180)):
        a, sa = gpu_render(ct, s)
        b, sb = gpu_render(ct, s, flags=ct.FLAG_FAST_BUILD)
        assert_same_frame(a, b, "sah treelets vs lbvh")
        assert sa["rays_total"] == sb["rays_total"]


def test_synthetic_grid_subset_vs_oracle(ct, oracle):
    """config 5 in small: instanced grid with mirror planes; a seeded pixel subset against the C oracle
    (brute force over every triangle is only affordable on a subset)."""
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=10, width=640, height=360)   # 90,000 triangles
    out, st = gpu_render(ct, s)
    px = np.random.default_rng(1).choice(s.width * s.height, 1500, replace=False).astype(np.uint64)
    ref = oracle.oracle_render(s, px=px)
    sub = {k: out[k][px.astype(np.int64)] for k in ("depth", "normal", "color", "hit_id")}
    assert_parity(compare(sub, ref), "synthetic grid subset vs oracle", oracle_is_host=True)


def test_output_stage_bytes_on_device(ct, oracle):
    """cutrace_encode_bytes_device == images.hpp byte mappings (oracle), on a real frame."""
    import torch

    s = load_golden_scene("sphere_plane").with_resolution(320, 180)
    with ct.Renderer(s) as r:
        r.render()
        out = r.download()
        n = s.width * s.height
        d = torch.from_numpy(out["depth"]).cuda(); nm = torch.from_numpy(out["normal"]).cuda(); c = torch.from_numpy(out["color"]).cuda()
        d8 = torch.empty(3 * n, dtype=torch.uint8, device="cuda"); n8 = torch.empty_like(d8); c8 = torch.empty_like(d8)
        r.encode_bytes_device(d.data_ptr(), nm.data_ptr(), c.data_ptr(), out["max_depth"], n, d8.data_ptr(), n8.data_ptr(), c8.data_ptr())
    od, on, oc = oracle.encode_bytes(out["depth"], out["normal"], out["color"], out["max_depth"])
    assert out["max_depth"] == oracle.max_depth(out["depth"])
    assert np.array_equal(d8.cpu().numpy().reshape(-1, 3), od)
    assert np.array_equal(n8.cpu().numpy().reshape(-1, 3), on)
    assert np.array_equal(c8.cpu().numpy().reshape(-1, 3), oc)


def test_errors_on_gpu(ct):
    s = load_golden_scene("triangle")
    lib = ct._lib.load()
    with ct.Renderer(s) as r:
        buf = np.empty(400, np.float32)
        assert lib.cutrace_download(r._ctx, buf.ctypes.data, None, None, None, None) == -5   # before any render
        assert b"before" in lib.cutrace_last_error()
    # primitive data is validated on the device, next to the bounds computation
    bad = load_golden_scene("bunny")
    bad.tri_p2 = bad.tri_p2.copy()
    bad.tri_p2[123, 1] = np.inf
    with pytest.raises(ct.CutraceError) as e:
        ct.Renderer(bad)
    assert e.value.code == -1 and "non-finite" in str(e.value)
    bad = load_golden_scene("bunny")
    bad.tri_object = bad.tri_object.copy()
    bad.tri_object[7] = 99
    with pytest.raises(ct.CutraceError) as e:
        ct.Renderer(bad)
    assert e.value.code == -1 and "out of range" in str(e.value)
    with pytest.raises(ct.CutraceError):
        ct.Renderer(s, device=99)
    with pytest.raises(ct.CutraceError):
        ct.Renderer(s, bounces=16)


def test_cli_end_to_end(ct, oracle, tmp_path):
    """`cutrace <scene>` (C++ host over the C-ABI): scene dump text, timing line, and the three JPEGs of main.cu:34-36
    decoded and compared with the oracle's byte images (JPEG q90 is lossy: PSNR, not equality)."""
    import subprocess

    from PIL import Image

    from conftest import ROOT
    from cutrace_b200 import host

    exe = os.path.join(ROOT, "bin", "cutrace")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "cli"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, os.path.join(ROOT, "scenes", "solids.json"), "--out-dir", str(tmp_path), "--dump-raw"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert " -> Have 5    objects:" in r.stdout and "  -> Object   #0    has type #1 " in r.stdout       # kernel.hpp:152-155
    assert " -> Have 2    lights:" in r.stdout and " -> Have 4    materials:" in r.stdout
    assert "Render time was" in r.stdout and "kernel time with setup/teardown was" in r.stdout          # main.cu:32
    s = host.load_scene(os.path.join(ROOT, "scenes", "solids.json"), base_dir=ROOT)
    ref = oracle.oracle_render(s)
    n = s.width * s.height
    raw = {"depth": np.fromfile(tmp_path / "depth.f32", np.float32), "normal": np.fromfile(tmp_path / "normal.f32", np.float32).reshape(n, 3),
           "color": np.fromfile(tmp_path / "color.f32", np.float32).reshape(n, 3), "hit_id": np.fromfile(tmp_path / "hit_id.u32", np.uint32)}
    assert_parity(compare(raw, ref, s.width, s.height), "CLI raw dump vs oracle", oracle_is_host=True)
    d8, n8, c8 = oracle.encode_bytes(ref["depth"], ref["normal"], ref["color"], oracle.max_depth(ref["depth"]))
    for fn, want in (("depth_map.jpg", d8), ("normal_map.jpg", n8), ("frame.jpg", c8)):
        im = Image.open(tmp_path / fn)
        assert im.size == (s.width, s.height)
        got = np.asarray(im.convert("RGB")).astype(np.float64).reshape(n, 3)
        mse = ((got - want) ** 2).mean()
        assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) > 30.0, fn


def test_cli_renders_the_synthetic_grid_from_json_and_stl_files(ct, tmp_path):
    """the G=3 grid written as reference-schema JSON + one binary STL per instance, rendered by
    the CLI from the files (cwd = their directory, like the reference resolves mesh paths) == the in-memory scene
    rendered through the C-ABI, bit for bit (both front-ends give the same arrays, the path is deterministic)."""
    import subprocess

    from conftest import ROOT
    from cutrace_b200 import synth

    exe = os.path.join(ROOT, "bin", "cutrace")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "cli"], check=True, stdout=subprocess.DEVNULL)
    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=3, width=192, height=108)
    synth.write_scene_files(s, str(tmp_path), *synth.grid_camera(3), name="grid3")
    r = subprocess.run([exe, "grid3.json", "--dump-raw"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert " -> Have 15   objects:" in r.stdout and " -> Have 3    lights:" in r.stdout
    out, st = gpu_render(ct, s)
    n = s.width * s.height
    assert np.array_equal(np.fromfile(tmp_path / "hit_id.u32", np.uint32), out["hit_id"].reshape(-1))
    assert np.array_equal(np.fromfile(tmp_path / "depth.f32", np.float32), out["depth"].reshape(-1))
    assert np.array_equal(np.fromfile(tmp_path / "normal.f32", np.float32), out["normal"].reshape(-1))
    assert np.array_equal(np.fromfile(tmp_path / "color.f32", np.float32), out["color"].reshape(-1))
    assert (out["hit_id"].reshape(-1) < 9).any()                      # the instances are on screen
    for fn in ("depth_map.jpg", "normal_map.jpg", "frame.jpg"):
        assert os.path.getsize(tmp_path / fn) > 500


def test_download_bytes_equals_host_output_stage(ct, oracle):
    s = load_golden_scene("mirror").with_resolution(240, 135)
    with ct.Renderer(s) as r:
        r.render()
        out = r.download()
        b = r.download_bytes()
    d8, n8, c8 = oracle.encode_bytes(out["depth"], out["normal"], out["color"], out["max_depth"])
    assert np.array_equal(b["depth_rgb"], d8) and np.array_equal(b["normal_rgb"], n8) and np.array_equal(b["color_rgb"], c8)


def test_overlapped_and_serialized_frames_are_bit_identical(ct, monkeypatch):
    """Multi-launch scheduler: shade kernels overlapping the trace chain on three streams vs everything on one stream
    (CUTRACE_FLAG_SERIALIZE) — the same kernels in a different order: same bits."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", "launches")
    for name, res in (("bunny", (640, 360)), ("sphere_plane", (640, 360))):
        s = load_golden_scene(name).with_resolution(*res)
        a, sa = gpu_render(ct, s)
        b, sb = gpu_render(ct, s, flags=ct.FLAG_SERIALIZE)
        for k in ("depth", "normal", "hit_id"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (name, k)
        if name == "bunny":      # one ray per pixel and level: deterministic sum
            assert np.array_equal(a["color"].view(np.uint32), b["color"].view(np.uint32))
        else:                    # material 1 reflects AND transmits: float atomics, order-dependent last bits
            assert np.abs(a["color"] - b["color"]).max() < 1e-5
        assert sa["rays_total"] == sb["rays_total"] and sb["trace_ms"] > 0 and sb["shade_ms"] > 0


def test_frame_kernel_equals_the_multi_launch_paths(ct, monkeypatch):
    """The persistent frame kernel (device-side level loop; CUTRACE_FLAG_FRAME_KERNEL) against the multi-launch scheduler that
    drives the same device functions (CUTRACE_FLAG_LAUNCHES: one launch per level and kind on three streams, captured as a
    CUDA graph from the second frame on).  Depth, normals, ids and every counter are identical; colours agree to the last
    bits (the two are separate compilations: tests/parity.py assert_same_frame); each scheduler is bit-reproducible.  The
    frame kernel is ONE launch and reports when each bounce level was complete."""
    for name, res in (("bunny", (640, 360)), ("mirror", (333, 205)), ("triangle", None)):
        s = load_golden_scene(name)
        if res:
            s = s.with_resolution(*res)
        with ct.Renderer(s, flags=ct.FLAG_FRAME_KERNEL) as r:
            sa = r.render()
            sa = r.render()
            a = r.download()
            ph = r.phase_ms()
            r.render()
            assert_same_frame(a, r.download(), name)          # two frame-kernel frames: same bits
        assert sa["kernel_launches"] == 1
        levels = 6 if name != "triangle" else 1
        assert len(ph) == levels + 1 and all(b >= a_ for a_, b in zip(ph, ph[1:])) and ph[-1] <= sa["render_ms"] * 1.5 + 0.05, ph
        with ct.Renderer(s, flags=ct.FLAG_LAUNCHES) as r:
            frames = []
            for _ in range(3):     # direct, graph capture, graph replay
                sb = r.render()
                frames.append(r.download())
            assert r.phase_ms() == []
        assert sb["kernel_launches"] == 2 * levels + 1
        for b in frames:
            assert_same_frame(a, b, name)                      # every rounding is pinned (common.cuh): same bits from every scheduler
        for k in ("rays_total", "rays_shadow", "rays_reflect", "shadow_casts", "max_depth"):
            assert sa[k] == sb[k], (name, k)
        # the per-pixel kernel: same G-buffer and counters, colours to the last bits
        with ct.Renderer(s, flags=ct.FLAG_PIXEL_KERNEL) as r:
            sp = r.render()
            p = r.download()
            r.render()
            assert_same_frame(p, r.download(), name)          # bit-reproducible
        assert sp["kernel_launches"] == 1 and sp["scheduler"] == 2
        assert_same_frame(p, a, name)
        for k in ("rays_total", "rays_shadow", "rays_reflect", "shadow_casts", "max_depth"):
            assert sa[k] == sp[k], (name, k)
        with ct.Renderer(s) as r:                              # the default is the per-pixel kernel
            assert r.render()["scheduler"] == 2


def test_direct_first_frame_and_graph_replays_are_bit_identical(ct, monkeypatch):
    """Multi-launch scheduler (CUTRACE_SCHEDULER=launches): the first frame of a ctx is enqueued stream by stream, the second
    captures the CUDA graph, later ones replay it: same bits and same counters every time (bunny.json: one ray per pixel
    and level, deterministic sum)."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", "launches")
    s = load_golden_scene("bunny").with_resolution(320, 180)
    with ct.Renderer(s) as r:
        frames = []
        for _ in range(4):
            st = r.render()
            frames.append((r.download(), st))
    a, sa = frames[0]
    for b, sb in frames[1:]:
        for k in ("depth", "normal", "color", "hit_id"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
        assert sa["rays_total"] == sb["rays_total"] and sa["kernel_launches"] == sb["kernel_launches"] == 13
        assert a["max_depth"] == b["max_depth"]


def test_render_download_fused_equals_render_then_download(ct):
    """cutrace_render_download (G-buffer copies under the bounce levels, colour at the end) must return exactly what
    cutrace_render + cutrace_download return — also right after a camera change, when the frame still holds the
    previous view until the primary rays have been traced (a copy that started too early would return the old view)."""
    s = load_golden_scene("bunny").with_resolution(1280, 720)
    eye2 = np.array([0.2, 0.5, 2.2], np.float32)
    f2, r2, u2 = ct.look_at(eye2, [0, 1, 0], [0, 0, 0])
    with ct.Renderer(s) as r:
        a, sta = r.render_download()
        r.render()
        b = r.download()
        for k in ("depth", "normal", "color", "hit_id"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
        assert a["max_depth"] == b["max_depth"] and sta["rays_total"] > 0
        for _ in range(3):    # replayed graph
            c, _ = r.render_download()
            assert np.array_equal(c["color"].view(np.uint32), b["color"].view(np.uint32))
        r.set_camera(eye2, u2, f2, r2, s.ambient, s.width, s.height)
        d, _ = r.render_download()
        r.render()
        e = r.download()
        for k in ("depth", "normal", "color", "hit_id"):
            assert np.array_equal(d[k].view(np.uint32), e[k].view(np.uint32)), k
        assert not np.array_equal(d["depth"], b["depth"])


def test_render_download_with_the_frame_kernel(ct, monkeypatch):
    """cutrace_render_download under the frame-kernel scheduler: trace(0) runs as its own kernel (the G-buffer copies start
    behind it), the frame kernel takes over at level 1.  G-buffer identical to render + download; colours to the last bits."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", "frame")
    s = load_golden_scene("bunny").with_resolution(1280, 720)
    with ct.Renderer(s) as r:
        a, sta = r.render_download()
        stb = r.render()
        b = r.download()
        assert sta["kernel_launches"] == 2 and stb["kernel_launches"] == 1
        assert_same_frame(a, b, "fused vs render + download")
        assert a["max_depth"] == b["max_depth"] and sta["rays_total"] == stb["rays_total"]
        c, _ = r.render_download()
        assert_same_frame(a, c, "two fused frames")


@pytest.mark.parametrize("width,height,skew", [(1000, 562, 0), (1001, 563, 0), (1000, 562, 4), (1920, 1080, 0)])
def test_render_download_straight_into_pinned_host_memory(ct, width, height, skew):
    """cutrace_render_download with pinned + mapped destinations (cutrace_host_alloc): the pixel kernel stores every finished pixel
    into the ctx's frame AND into the caller's images over PCIe — no copy afterwards.  Same bits as render + download, the
    device frame stays valid, NULL destinations are skipped, pageable destinations take the copy path.  Odd sizes (partial tiles
    at both edges), a width that is no multiple of 4, destinations that start 4 bytes into the pinned block, and BASELINE's 1080p."""
    import ctypes as C

    lib = ct._lib.load()
    s = load_golden_scene("bunny").with_resolution(width, height)     # odd sizes: partial tiles at both edges
    n = s.width * s.height
    ptrs, pinned = [], {}
    for k, m, dt in (("depth", 1, np.float32), ("normal", 3, np.float32), ("color", 3, np.float32), ("hit_id", 1, np.uint32)):
        p = lib.cutrace_host_alloc(n * m * 4 + 16)
        ptrs.append(p)
        pinned[k] = np.ctypeslib.as_array(C.cast(p + skew, C.POINTER(C.c_float if dt is np.float32 else C.c_uint32)), shape=(n * m,))
        pinned[k][:] = 0
    with ct.Renderer(s) as r:
        md, st = C.c_float(), ct.cutrace_stats()
        ct._lib.check(lib.cutrace_render_download(r._ctx, pinned["depth"].ctypes.data, pinned["normal"].ctypes.data, pinned["color"].ctypes.data,
                                                  pinned["hit_id"].ctypes.data, C.byref(md), C.byref(st)))
        assert st.kernel_launches == 1 and st.scheduler == 2
        dev = r.download()                                        # the device frame of the same render
        got = {k: pinned[k].reshape(dev[k].shape).copy() for k in dev if k != "max_depth"}
        assert_same_frame(got, dev, "pinned destinations vs device frame")
        assert md.value == dev["max_depth"]
        r.render()
        assert_same_frame(got, r.download(), "direct download vs plain render")
        pinned["color"][:] = 7.0                                  # only depth this time: colour must stay untouched
        ct._lib.check(lib.cutrace_render_download(r._ctx, pinned["depth"].ctypes.data, None, None, None, C.byref(md), None))
        assert np.all(pinned["color"] == 7.0) and np.array_equal(pinned["depth"], dev["depth"])
        pageable, _ = r.render_download()                         # numpy memory: copy path
        assert_same_frame(pageable, dev, "pageable destinations")
    del pinned, got
    for p in ptrs:
        lib.cutrace_host_free(p)


def test_set_camera_resizes_and_reuses_the_scene(ct):
    """cutrace_set_camera: new resolution / view on an uploaded scene (no rebuild) == a fresh ctx."""
    s = load_golden_scene("mirror").with_resolution(320, 180)
    big = s.with_resolution(500, 281)
    fresh, _ = gpu_render(ct, big)
    with ct.Renderer(s) as r:
        r.render()
        r.set_resolution(500, 281)
        r.render()
        out = r.download()
        for k in ("depth", "normal", "color", "hit_id"):
            assert np.array_equal(out[k].view(np.uint32), fresh[k].view(np.uint32)), k
        r.set_resolution(320, 180)
        r.render()
        small = r.download()
    again, _ = gpu_render(ct, s)
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(small[k].view(np.uint32), again[k].view(np.uint32)), k


def test_pixel_batches_equal_one_batch(ct, monkeypatch):
    """Wavefront schedulers: when the worst-case queues do not fit the memory budget the frame is rendered in pixel batches
    (CUTRACE_QUEUE_BUDGET_MB forces that here); the result must not change."""
    monkeypatch.setenv("CUTRACE_SCHEDULER", "launches")
    for name, res in (("bunny", (640, 360)), ("sphere_plane", (320, 180))):
        s = load_golden_scene(name).with_resolution(*res)
        one, st1 = gpu_render(ct, s)
        monkeypatch.setenv("CUTRACE_QUEUE_BUDGET_MB", "8" if name == "bunny" else "24")
        many, st2 = gpu_render(ct, s)
        monkeypatch.delenv("CUTRACE_QUEUE_BUDGET_MB")
        assert st2["kernel_launches"] > st1["kernel_launches"], "the budget did not force batches"
        for k in ("depth", "normal", "hit_id"):
            assert np.array_equal(one[k].view(np.uint32), many[k].view(np.uint32)), (name, k)
        if name == "bunny":
            assert np.array_equal(one["color"].view(np.uint32), many["color"].view(np.uint32))
        else:
            assert np.abs(one["color"] - many["color"]).max() < 1e-5
        assert st1["rays_total"] == st2["rays_total"] and st1["max_depth"] == st2["max_depth"]


def test_top_of_bvh_in_shared_memory_mode(ct, oracle, monkeypatch):
    """MODE 2 (breadth-first top of the tree staged in shared memory, rest global; off by default because it measured
    slower) stays correct: the BVH validates after the relabelling and a pixel subset matches the oracle."""
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=12, width=480, height=270)   # 129,600 triangles: does not fit in shared memory
    base, st0 = gpu_render(ct, s)
    monkeypatch.setenv("CUTRACE_SMEM_TOP_NODES", "700")
    with ct.Renderer(s, flags=ct.FLAG_VALIDATE_BVH) as r:
        r.validate_bvh()
        st = r.render()
        out = r.download()
    monkeypatch.delenv("CUTRACE_SMEM_TOP_NODES")
    assert st["smem_nodes"] == 700 and st0["smem_nodes"] == 0
    m = compare(out, base, s.width, s.height)
    assert m["id_mismatch"] <= 2 and m["depth_max_rel"] <= 1e-6, m
    px = np.random.default_rng(2).choice(s.width * s.height, 1000, replace=False).astype(np.uint64)
    ref = oracle.oracle_render(s, px=px)
    sub = {k: out[k][px.astype(np.int64)] for k in ("depth", "normal", "color", "hit_id")}
    assert_parity(compare(sub, ref), "MODE 2 subset vs oracle", oracle_is_host=True)


def test_cli_multi_gpu_in_one_process(ct, oracle, tmp_path):
    """`cutrace --gpus 2`: one process, one ctx and one host thread per GPU, GPU 1 stores its tiles into GPU 0's frame
    over NVLink (cutrace_enable_peer_access + cutrace_frame_attach).  Needs two GPUs."""
    import subprocess

    import torch

    from conftest import ROOT
    from cutrace_b200 import host

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "bin", "cutrace")
    outs = {}
    for gpus in (1, 2):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        r = subprocess.run([exe, os.path.join(ROOT, "scenes", "solids.json"), "--out-dir", str(d), "--dump-raw", "--gpus", str(gpus),
                            "--width", "640", "--height", "400"], capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 0, r.stderr
        outs[gpus] = {k: np.fromfile(d / f"{k}.{e}", t) for k, e, t in (("depth", "f32", np.float32), ("normal", "f32", np.float32),
                                                                       ("color", "f32", np.float32), ("hit_id", "u32", np.uint32))}
    for k in ("depth", "normal", "hit_id"):
        assert np.array_equal(outs[1][k].view(np.uint32), outs[2][k].view(np.uint32)), k
    assert np.abs(outs[1]["color"] - outs[2]["color"]).max() < 1e-5    # solids.json has a reflecting + transmitting sphere: float atomics


def test_integration_binding_against_the_reference_operator(ct):
    """oracle/_ref/integration_check: ONE cutrace::cpu::schema::default_cpu_scene (the reference's own host scene type)
    through (a) the reference's default_to_gpu + gpu::render<S,5,256> and (b) the flatten() + C-ABI binding that
    INTEGRATION.md shows, compared in the same process."""
    import json
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "oracle", "_ref", "integration_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/integration_check was not built (reference tree absent at build time)")
    r = subprocess.run([exe, "640", "360"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    m = json.loads(r.stdout.strip().splitlines()[-1])
    # depth differing on a pixel = another object won an edge tie there (normals then differ too): at most 2 of them
    assert m["sentinel_mismatch"] == 0, m
    assert m["depth_off_pixels"] <= 2, m
    if m["depth_off_pixels"] == 0:
        assert m["normal_max_abs"] <= 1e-5 and m["depth_max_rel"] <= 1e-6, m
    assert m["color_max_abs"] < 5e-2, m
    assert m["max_ref"] == m["max_new"]


def test_cta_work_cursors_render_the_same_frame_as_the_global_cursor(ct, monkeypatch):
    """The pixel kernel's per-CTA work cursors (render.cu: claim_segment — the default for scenes walked through L1 / L2 on frames of
    at least 2^20 pixels) against its single global cursor (CUTRACE_PIXEL_SEG=0), on a frame whose size is no multiple of a tile
    or a chunk: bit-identical, unsharded and as the three shards of a world-3 render stored into one frame."""
    s = load_golden_scene("mirror").with_resolution(1283, 1021)          # 1.31 M pixels, partial tiles at both edges
    flags = ct.FLAG_NO_SMEM_TOP | ct.FLAG_PIXEL_KERNEL                  # walk the BVH through L1 / L2: the mode that uses the cursors
    monkeypatch.setenv("CUTRACE_PIXEL_SEG", "0")
    want, st0 = gpu_render(ct, s, flags=flags)
    monkeypatch.delenv("CUTRACE_PIXEL_SEG")
    got, st1 = gpu_render(ct, s, flags=flags)
    assert st1["scheduler"] == 2 and st1["rays_total"] == st0["rays_total"]
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    monkeypatch.setenv("CUTRACE_PIXEL_SEG", "1")                          # force them on the shards too (0.44 M pixels each)
    rs = [ct.Renderer(s, tile_rank=r, tile_world=3, flags=flags) for r in range(3)]
    rs[0].frame_ipc_export()
    block = rs[0].frame_device()[0]
    for r in rs[1:]:
        r.frame_attach(block)
    rays = sum(r.render()["rays_total"] for r in rs)
    out = rs[0].download()
    for r in rs:
        r.close()
    assert rays == st0["rays_total"]
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(out[k].view(np.uint32), want[k].view(np.uint32)), ("world 3", k)
