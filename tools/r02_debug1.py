#!/usr/bin/env python
"""debug: (C) frame kernel vs serialized colour differences, (A) depth differences of the grid scene vs the reference kernel"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cutrace_b200 as ct
from cutrace_b200 import synth
from cutrace_b200.scene import FlatScene
from oracle import pyoracle as po
G = os.path.join(ROOT, "tests", "golden", "scenes")
s = FlatScene.load(os.path.join(G, "bunny.npz")).with_resolution(640, 360)
def rd(flags=0, **kw):
    with ct.Renderer(s, flags=flags, **kw) as r:
        r.render(); st = r.render(); return r.download(), st
a, sa = rd(); b, sb = rd(ct.FLAG_SERIALIZE); a2, _ = rd()
for name, x, y in (("frame vs serialized", a, b), ("frame vs frame (repeat)", a, a2)):
    d = np.abs(x["color"] - y["color"]).max(axis=1)
    nz = np.nonzero(d)[0]
    print(name, "differing px", len(nz), "max abs", d.max(), "first", nz[:10], "ulp-ish rel", (d[nz] / np.maximum(1e-9, np.abs(y["color"][nz]).max(axis=1))).max() if len(nz) else 0)
print("rays", sa["rays_total"], sb["rays_total"], "casts", sa["shadow_casts"], sb["shadow_casts"])
# (A)
meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(G, "bunny.npz")), FlatScene.load(os.path.join(G, "mirror.npz")))[:2]
g = synth.grid_scene(meshes, grid=6, width=480, height=270)
ref = po.ref_gpu_render(g)
def rg(flags=0):
    with ct.Renderer(g, flags=flags) as r:
        r.render(); return r.download()
o = rg(); ob = rg(ct.FLAG_BRUTE_FORCE)
host = po.oracle_render(g)
for name, x in (("bvh", o), ("brute", ob), ("host oracle", host)):
    same = x["hit_id"] == ref["hit_id"]
    fin = np.isfinite(ref["depth"]) & same
    e = np.zeros_like(ref["depth"], dtype=np.float64)
    e[fin] = np.abs(x["depth"][fin].astype(np.float64) - ref["depth"][fin]) / np.maximum(1, np.abs(ref["depth"][fin]))
    bad = np.nonzero(e > 1e-6)[0]
    print(f"(A) {name}: id mismatch {(~same).sum()}  depth rel > 1e-6 on {len(bad)} px, max {e.max():.3g}; worst px {bad[np.argsort(-e[bad])][:6]}")
    for p in bad[np.argsort(-e[bad])][:4]:
        print(f"     px {p}: this {x['depth'][p]!r} ref {ref['depth'][p]!r} bvh {o['depth'][p]!r} brute {ob['depth'][p]!r} host {host['depth'][p]!r} id {ref['hit_id'][p]}")
