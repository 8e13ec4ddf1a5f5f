"""Small CPU-side checks: synthetic scene generators are deterministic and well-formed, bench helpers, JPEG corner sizes."""
import io
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden_scene


def test_grid_scene_is_deterministic_and_counts_triangles():
    from cutrace_b200 import synth

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    assert [len(m) for m in meshes] == [1000, 800]          # bunny.stl, skull.stl
    a = synth.grid_scene(meshes, grid=5, width=64, height=36, seed=0)
    b = synth.grid_scene(meshes, grid=5, width=64, height=36, seed=0)
    assert a.n_triangles == 13 * 1000 + 12 * 800 and a.n_objects == 25 + 6 and a.n_planes == 6
    for k, v in a.to_npz_dict().items():
        assert np.array_equal(np.asarray(v), np.asarray(b.to_npz_dict()[k])), k
    assert a.max_children() == 1                             # mirrors but no transparency: one child per hit
    # G = 106 is the >= 10 M-triangle configuration of BASELINE.json
    assert (106 * 106 + 1) // 2 * 1000 + (106 * 106) // 2 * 800 == 10_112_400


def test_random_soup_is_well_formed(oracle):
    from cutrace_b200 import synth

    s = synth.random_soup(n_tri=60, n_sph=3, n_planes=2, n_lights=2, width=40, height=30, seed=4)
    assert s.n_objects == len(s.obj_material) == len(s.obj_kind)
    assert s.tri_object.max() < s.n_objects and s.obj_material.max() < len(s.mat_specular)
    assert s.max_children() == 2                             # material 2 reflects and transmits
    out = oracle.oracle_render(s)
    assert np.isfinite(out["color"]).all() and (out["hit_id"] != 0xFFFFFFFF).any()


def test_bench_helpers():
    import bench

    assert bench.survey_bytes_per_ray(1005) == 64 + 64 * 10 + 48 == 752          # SURVEY.md §8d, bunny.json
    assert bench.survey_bytes_per_ray(10_112_406) == 64 + 64 * 24 + 48 == 1648   # 10 M-triangle scene
    scene, wl = bench.load_workload("bunny4k")
    assert (scene.width, scene.height) == (3840, 2160) and bench.n_primitives(scene) == 1005
    # both arms print the same `config` object; the unique-ray numerator is a constant of (scene, resolution)
    cfg = bench.config_of(scene, wl)
    assert cfg["rays_per_frame"] == bench.RAYS_PER_FRAME["bunny4k"] == 248_825_266 and set(cfg) == {"workload", "rays_per_frame", "primitives", "bounces", "l2"}
    # compulsory DRAM bytes of a bunny.json frame: ~9 GB, i.e. well under the HBM roofline at 10 ms/frame
    n_px = 3840 * 2160
    comp = bench.compulsory_bytes_per_frame(n_px, 248_825_266, 4 * 6 * n_px, 4, 50_000, pixel_kernel=False)
    assert 8e9 < comp < 11e9                                                      # wavefront: queues dominate
    assert bench.compulsory_bytes_per_frame(n_px, 248_825_266, 4 * 6 * n_px, 4, 50_000) == 32 * n_px + 50_000   # pixel kernel: the frame
    cs = bench.ClockSampler(0, enabled=False)
    cs.lines = ["1965, 1965, Not Active, Not Active, Not Active, Active", "1950, 1965, Not Active, Not Active, Not Active, Not Active"]
    s = cs.summary()
    assert s["sm_mhz"] == 1957.5 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["samples"] == 2


@pytest.mark.parametrize("w,h", [(1, 1), (17, 9), (16, 16), (33, 47)])
def test_jpeg_corner_sizes(tmp_path, w, h):
    import subprocess

    from PIL import Image

    from cutrace_b200 import host

    if not os.path.exists(host.HOST_LIB_PATH):
        subprocess.run(["make", "-C", ROOT, "host"], check=True, stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(w * 100 + h)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), np.full((h, w), 128)], -1).astype(np.uint8)
    img[rng.integers(0, h), rng.integers(0, w)] = 255
    path = str(tmp_path / "t.jpg")
    host.write_jpeg(path, img, 90)
    im = Image.open(path)
    assert im.size == (w, h)
    got = np.asarray(im.convert("RGB")).astype(float)
    assert np.abs(got - img).mean() < 12.0


def test_bench_stdout_carries_only_the_result_line(tmp_path):
    """bench.py's contract is ONE JSON line on stdout; library banners on fd 1 (NCCL's version line) must land on stderr."""
    import json
    import subprocess
    import sys

    script = tmp_path / "emit_check.py"
    script.write_text(
        "import os, sys, subprocess\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "bench.claim_stdout()\n"
        "print('stray python print')\n"
        "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"
        "subprocess.run(['echo', 'stray child output'])\n"
        "bench.emit({'metric': 'x', 'value': 1})\n")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, check=True)
    assert json.loads(r.stdout) == {"metric": "x", "value": 1} and r.stdout.count("\n") == 1
    assert "NCCL version" in r.stderr and "stray python print" in r.stderr and "stray child output" in r.stderr
