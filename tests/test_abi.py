"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cutrace.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden_scene


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from cutrace_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from cutrace_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "cutrace.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cutrace_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed from include/cutrace.h"
    assert declared == set(_lib.SYMBOLS), f"binding list and header differ: {declared ^ set(_lib.SYMBOLS)}"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} is declared in include/cutrace.h but not exported"
    assert lib.cutrace_abi_version() == 1
    from cutrace_b200.scene import TILE

    assert lib.cutrace_tile_size() == TILE


def test_struct_layouts_match_header(lib):
    """ctypes mirrors vs the C structs: sizes are computed by compiling a tiny C file."""
    import subprocess
    import tempfile

    from cutrace_b200 import _lib
    from cutrace_b200.scene import cutrace_scene_desc

    src = '#include <stdio.h>\n#include "cutrace.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(cutrace_scene_desc), sizeof(cutrace_opts), sizeof(cutrace_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "t.c"), "w") as f:
            f.write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(cutrace_scene_desc), C.sizeof(_lib.cutrace_opts), C.sizeof(_lib.cutrace_stats)]


def test_default_opts_are_the_reference_call(lib):
    from cutrace_b200 import _lib

    o = _lib.cutrace_opts()
    lib.cutrace_default_opts(C.byref(o))
    assert o.bounces == 5 and abs(o.fudge - 1e-3) < 1e-9 and o.device == -1  # main.cu:30


def test_invalid_arguments_are_rejected_without_a_device(lib):
    import cutrace_b200 as ct

    s = load_golden_scene("triangle")
    d = s.as_desc()
    d.abi_version = 99
    ctx = C.c_void_p()
    assert lib.cutrace_upload_scene(C.byref(d), None, C.byref(ctx)) == -1
    assert b"abi_version" in lib.cutrace_last_error()
    bad = load_golden_scene("triangle")
    bad.obj_material = np.array([7], np.uint32)
    d = bad.as_desc()
    assert lib.cutrace_upload_scene(C.byref(d), None, C.byref(ctx)) == -1
    assert lib.cutrace_render(None, None) == -1
    assert lib.cutrace_download(None, None, None, None, None, None) == -1
    nan = load_golden_scene("triangle")
    nan.tri_p1 = nan.tri_p1.copy()
    nan.tri_p1[0, 0] = np.nan
    with pytest.raises(ct.CutraceError):
        ct.Renderer(nan)


def test_no_cpu_fallback(lib):
    """Without a GPU the product must fail loudly, never render on the CPU."""
    import torch

    import cutrace_b200 as ct

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ct.CutraceError) as e:
        ct.render(load_golden_scene("triangle"))
    assert e.value.code == -3


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under cutrace_b200/ may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "cutrace_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, fn), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|pyoracle|cutrace_oracle|libcutrace_ref", txt):
                    bad.append(fn)
    assert not bad, bad


@pytest.mark.parametrize("curve", [1, 0])
@pytest.mark.parametrize("dims", [(20, 20), (1920, 1080), (300, 2), (1021, 577), (16 * 9 * 3 + 5, 16 * 8 * 2 + 1)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_slot_mapping_is_a_bijection_and_balanced(lib, curve, dims, world):
    """Host-only hook over the kernels' own tile_of_slot / slot_of_tile (csrc/common.cuh): every screen tile has exactly one slot,
    the two functions invert each other, padding slots are reported, and (super-tile order) every rank's tiles cover the frame at
    tile granularity: every 8 x 8-tile window of an 8-rank frame holds every rank in nearly equal shares, and consecutive local tiles of a rank stay close."""
    from cutrace_b200 import distributed as dist_py
    from cutrace_b200.scene import TILE

    w, h = dims
    tiles_x, tiles_y = (w + TILE - 1) // TILE, (h + TILE - 1) // TILE
    n = tiles_x * tiles_y
    tx, ty = C.c_uint32(), C.c_uint32()
    owner = np.full((tiles_y, tiles_x), -1, np.int64)
    pos = []
    for slot in range(n):
        assert lib.cutrace_debug_tile_of_slot(w, h, world, curve, slot, C.byref(tx), C.byref(ty)) == 1
        assert tx.value < tiles_x and ty.value < tiles_y
        assert owner[ty.value, tx.value] == -1, "two slots show the same tile"
        owner[ty.value, tx.value] = slot % world
        assert lib.cutrace_debug_slot_of_tile(w, h, world, curve, tx.value, ty.value) == slot
        if curve:   # the Python mirror the gloo tests and untile_host use
            assert dist_py.tile_of_slot(slot, w, h) == (tx.value, ty.value) and dist_py.slot_of_tile(tx.value, ty.value, w, h) == slot
        pos.append((tx.value, ty.value))
    assert (owner >= 0).all()
    assert lib.cutrace_debug_tile_of_slot(w, h, world, curve, n, C.byref(tx), C.byref(ty)) == 0
    if curve and world == 8 and tiles_x >= 32 and tiles_y >= 32:
        for y0 in range(0, tiles_y - 8, 5):          # every 8 x 8-tile window holds every rank, in nearly equal shares
            for x0 in range(0, tiles_x - 8, 7):
                counts = np.bincount(owner[y0:y0 + 8, x0:x0 + 8].ravel(), minlength=8)
                assert counts.min() >= 5 and counts.max() <= 11, (x0, y0, counts)
        p = np.asarray(pos[0::8], np.int64)            # rank 0's tiles in its own order
        step = np.abs(np.diff(p, axis=0)).max(axis=1)
        assert np.percentile(step, 90) <= 9, "a rank's consecutive tiles are neighbours inside a super-tile"


@pytest.mark.parametrize("n_work,G", [(1024, 1), (4096, 148), (4096 * 148, 148), (4096 * 148 + 256, 148), (1_048_576 + 768, 148), (33_177_600 // 8 + 256, 148),
                                      (300_032, 7), (256, 320)])
def test_per_cta_work_segments_cover_the_frame_once(lib, n_work, G):
    """The pixel kernel's per-CTA cursors (csrc/common.cuh: seg_len / seg_to_work, render.cu: claim_segment): the CTAs' own index
    spaces map onto the frame's work items exactly once, 32-item warp blocks never straddle a chunk, and the segments are balanced
    to one 4096-item chunk."""
    seen = np.zeros(n_work, np.uint8)
    lens = []
    for k in range(G):
        ln = lib.cutrace_debug_segment_length(n_work, k, G)
        lens.append(ln)
        assert ln % 32 == 0 or n_work % 32
        for o in range(0, ln, 32):                      # a warp iteration: offsets o .. o + 31 of CTA k
            w0 = lib.cutrace_debug_segment_work(o, k, G)
            assert w0 + min(32, ln - o) <= n_work, (k, o, w0)
            assert w0 // 4096 == (w0 + min(32, ln - o) - 1) // 4096        # one chunk
            assert (w0 // 4096) % G == k                                     # ... of this CTA
            seen[w0:w0 + min(32, ln - o)] += 1
    assert sum(lens) == n_work and (seen == 1).all()
    assert max(lens) - min(lens) <= 4096
