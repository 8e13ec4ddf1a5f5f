#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (one tool per gpurun call): LBVH build, all kernel variants, sharded +
attached frames, batches, output stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cutrace_b200 as ct
from cutrace_b200 import synth
from cutrace_b200.scene import FlatScene
G = os.path.join(ROOT, "tests", "golden", "scenes")
for name, res in (("triangle", None), ("sphere_plane", (64, 36)), ("mirror", (64, 36)), ("bunny", (48, 27))):
    s = FlatScene.load(os.path.join(G, name + ".npz"))
    if res: s = s.with_resolution(*res)
    for flags in (0, ct.FLAG_NO_SMEM_TOP, ct.FLAG_SERIALIZE, ct.FLAG_BRUTE_FORCE):
        with ct.Renderer(s, flags=flags | ct.FLAG_VALIDATE_BVH) as r:
            r.render(); r.download(); r.download_bytes(); r.render_download()
s = synth.random_soup(n_tri=200, n_sph=4, n_planes=2, n_lights=3, width=70, height=50, seed=3)
rs = [ct.Renderer(s, tile_rank=k, tile_world=3) for k in range(3)]
rs[0].frame_ipc_export(); blk = rs[0].frame_device()[0]
for r in rs[1:]: r.frame_attach(blk)
for r in rs: r.render()
rs[0].download()
for r in rs: r.close()
os.environ["CUTRACE_QUEUE_BUDGET_MB"] = "1"
with ct.Renderer(s) as r:
    r.render(); r.download()
os.environ.pop("CUTRACE_QUEUE_BUDGET_MB")
meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(G, "bunny.npz")), FlatScene.load(os.path.join(G, "mirror.npz")))[:2]
g = synth.grid_scene(meshes, grid=4, width=96, height=54)
os.environ["CUTRACE_SMEM_TOP_NODES"] = "300"
with ct.Renderer(g, flags=ct.FLAG_VALIDATE_BVH) as r:
    r.render(); r.set_resolution(40, 30); r.render(); r.download()
print("sanitize case done")
