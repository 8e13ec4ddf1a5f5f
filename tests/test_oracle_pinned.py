"""The oracle (oracle/cutrace_oracle.c) against the reference.

The reference has no golden vectors for the render path.  The pin is the
reference's OWN source compiled for the host (oracle/ref_host.cpp): tests/golden/*.npz were produced
by it (tools/make_golden.py) and the C restatement must reproduce them BIT-EXACTLY — both are
float32 code evaluated in the same order with contraction off.  When /root/reference is present
(build container) the shim itself is also run live on other resolutions / pixel subsets.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, REF, load_golden_scene


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_oracle_reproduces_reference_golden_bit_exact(oracle, name):
    g = np.load(os.path.join(GOLDEN, GOLDEN_CASES[name]))
    s = load_golden_scene(name).with_resolution(int(g["width"]), int(g["height"]))
    out = oracle.oracle_render(s, fudge=float(g["fudge"]), bounces=int(g["bounces"]))
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(_bits(out[k]), _bits(g[k])), f"{name}: {k} differs from the reference golden"


def test_reference_call_counts(oracle):
    """SURVEY.md §8a10 probe: bunny = 1 + 6*(1+4) ray_cast calls per pixel when every ray hits;
    triangle.json = 2.095 calls/px (400 px, 38 hit)."""
    s = load_golden_scene("triangle")
    c = oracle.oracle_render(s)["counters"]
    assert c["rays_primary"] == 400 and c["rays_shadow"] == 38 and c["casts"] == 2 * 400 + 38
    s = load_golden_scene("bunny").with_resolution(48, 27)
    c = oracle.oracle_render(s)["counters"]
    n = 48 * 27
    assert c["casts"] <= 31 * n and c["casts"] > 30.8 * n
    assert c["rays_shadow"] == 4 * (c["rays_primary"] + c["rays_reflect"]) or c["rays_shadow"] <= 4 * 6 * n


def test_pixel_subset_matches_full_frame(oracle):
    s = load_golden_scene("mirror").with_resolution(64, 36)
    full = oracle.oracle_render(s)
    px = np.array([0, 5, 64 * 36 - 1, 1000, 1001, 77], dtype=np.uint64)
    sub = oracle.oracle_render(s, px=px)
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(_bits(sub[k]), _bits(full[k][px.astype(np.int64)]))


def test_miss_sentinels(oracle):
    """depth +INF, zero normal, black colour on a miss (inc/kernel.hpp:47-56, shading.hpp:119)."""
    s = load_golden_scene("triangle")
    out = oracle.oracle_render(s)
    miss = out["hit_id"] == 0xFFFFFFFF
    assert miss.sum() == 362
    assert np.all(np.isposinf(out["depth"][miss]))
    assert np.all(out["normal"][miss] == 0) and np.all(out["color"][miss] == 0)
    assert abs(oracle.max_depth(out["depth"]) - 5.196) < 1e-3


def test_output_stage_bytes(oracle):
    """images.hpp:26-88 byte mapping: nearest = brightest, non-finite -> 0, colour truncates."""
    depth = np.array([1.0, 2.0, np.inf, 4.0], np.float32)
    normal = np.array([[0, 0, 0], [0, 0, -2], [1, 0, 0], [0, 1e-7, 0]], np.float32)
    color = np.array([[0.5, -1, 2], [1, 0.999, 0], [np.nan, 0.25, 0.75], [0, 0, 0]], np.float32)
    d8, n8, c8 = oracle.encode_bytes(depth, normal, color, 4.0)
    assert d8[:, 0].tolist() == [191, 127, 0, 0]
    assert n8.tolist() == [[0, 0, 0], [127, 127, 0], [255, 127, 127], [0, 0, 0]]
    assert c8.tolist() == [[127, 0, 255], [255, 254, 0], [0, 63, 191], [0, 0, 0]]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "inc")), reason="reference tree not present")
@pytest.mark.parametrize("name,res", [("sphere_plane", (97, 61)), ("mirror", (80, 45)), ("bunny", (40, 30))])
def test_oracle_vs_live_reference_host_build(oracle, name, res):
    from cutrace_b200.scene import load_scene_json

    oracle.build()
    s = load_scene_json(os.path.join(REF, "scene", f"{name}.json"), base_dir=REF).with_resolution(*res)
    a = oracle.oracle_render(s)
    b = oracle.ref_host_render(s)
    for k in ("depth", "normal", "color", "hit_id"):
        assert np.array_equal(_bits(a[k]), _bits(b[k])), f"{name} {res}: {k}"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scene")), reason="reference tree not present")
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_golden_scene_fixtures_match_reference_files(name):
    """tests/golden/scenes/*.npz are exactly what the JSON/STL front-end reads from the reference."""
    from cutrace_b200.scene import load_scene_json

    a = load_scene_json(os.path.join(REF, "scene", f"{name}.json"), base_dir=REF)
    b = load_golden_scene(name)
    da, db = a.to_npz_dict(), b.to_npz_dict()
    for k in da:
        assert np.array_equal(np.asarray(da[k]), np.asarray(db[k])), k
