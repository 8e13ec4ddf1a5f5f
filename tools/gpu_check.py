#!/usr/bin/env python
"""Verbose GPU diagnostic: builds nothing, renders the reference scenes through the C-ABI and prints
parity against the C oracle, the committed goldens and (if built) the reference's own sm_100a kernel.
Usage (on the GPU box):  python tools/gpu_check.py [--big]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cutrace_b200 as ct  # noqa: E402
from cutrace_b200.scene import FlatScene  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from parity import compare  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [("triangle", None, "triangle_20x20.npz"), ("sphere_plane", (160, 90), "sphere_plane_160x90.npz"),
         ("mirror", (160, 90), "mirror_160x90.npz"), ("bunny", (96, 54), "bunny_96x54.npz")]


def run(scene, **kw):
    with ct.Renderer(scene, flags=ct.FLAG_VALIDATE_BVH | kw.pop("flags", 0), **kw) as r:
        st = r.render()
        out = r.download()
    return out, st


def main():
    big = "--big" in sys.argv
    for name, res, gold in CASES:
        s = FlatScene.load(os.path.join(GOLD, "scenes", f"{name}.npz"))
        if res:
            s = s.with_resolution(*res)
        g = dict(np.load(os.path.join(GOLD, gold)))
        for label, kw in (("bvh+smem", {}), ("bvh global", dict(flags=ct.FLAG_NO_SMEM_TOP)), ("brute", dict(flags=ct.FLAG_BRUTE_FORCE))):
            try:
                out, st = run(s, **kw)
            except Exception as e:  # noqa: BLE001
                print(f"[{name}] {label}: FAILED {e}")
                continue
            m = compare(out, g)
            print(f"[{name} {s.width}x{s.height}] {label}: vs golden {json.dumps(m)}")
            print(f"    stats: { {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()} }")
        o = po.oracle_render(s)
        print(f"    oracle counters: {o['counters']}")
    if po.have_ref_gpu():
        for name, res in (("sphere_plane", (640, 360)), ("mirror", (480, 270)), ("bunny", (320, 180))):
            s = FlatScene.load(os.path.join(GOLD, "scenes", f"{name}.npz")).with_resolution(*res)
            t = time.time()
            ref = po.ref_gpu_render(s, iters=1, warmup=0)
            out, st = run(s)
            m = compare(out, ref)
            print(f"[{name} {res}] vs reference sm_100a kernel: {json.dumps(m)}  ref_ms={ref['render_ms']:.2f} new_ms={st['render_ms']:.3f}")
            cpu = po.oracle_render(s)
            print(f"        vs C oracle: {json.dumps(compare(out, cpu))}")
            print(f"        ref-gpu vs C oracle: {json.dumps(compare(ref, cpu))}")
    if big:
        s = FlatScene.load(os.path.join(GOLD, "scenes", "bunny.npz")).with_resolution(3840, 2160)
        with ct.Renderer(s) as r:
            for i in range(4):
                st = r.render()
                rays = st["rays_total"]
                print(f"bunny 4K run {i}: render_ms={st['render_ms']:.3f} trace={st['trace_ms']:.3f} shade={st['shade_ms']:.3f} "
                      f"rays={rays} Mrays/s={rays / st['render_ms'] / 1e3:.1f} build_ms={st['build_ms']:.3f} nodes={st['bvh_nodes']} depth={st['bvh_depth']}")
        with ct.Renderer(s, flags=ct.FLAG_NO_SMEM_TOP) as r:
            for i in range(3):
                st = r.render()
                print(f"bunny 4K (global mode) run {i}: render_ms={st['render_ms']:.3f} trace={st['trace_ms']:.3f} shade={st['shade_ms']:.3f}")


def ref4k():
    s = FlatScene.load(os.path.join(GOLD, "scenes", "bunny.npz")).with_resolution(3840, 2160)
    ref = po.ref_gpu_render(s, iters=2, warmup=1)
    with ct.Renderer(s) as r:
        st = r.render()
        out = r.download()
    m = compare(out, ref)
    print(f"bunny 4K vs reference sm_100a kernel: {json.dumps(m)}")
    print(f"bunny 4K: ref render_ms={ref['render_ms']:.2f}  new render_ms={st['render_ms']:.3f}  rays={st['rays_total']}")


if __name__ == "__main__":
    if "--ref4k" in sys.argv:
        ref4k()
    else:
        main()
