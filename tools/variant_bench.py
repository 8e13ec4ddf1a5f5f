#!/usr/bin/env python
"""Times bunny 4K with the library given by $CUTRACE_B200_LIB in smem and global mode, leaf sizes 1..8."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cutrace_b200 as ct
from cutrace_b200.scene import FlatScene
wl = sys.argv[1] if len(sys.argv) > 1 else "bunny"
if wl.startswith("grid"):
    from cutrace_b200 import synth
    g = os.path.join(ROOT, "tests", "golden", "scenes")
    meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(g, "bunny.npz")), FlatScene.load(os.path.join(g, "mirror.npz")))[:2]
    s = synth.grid_scene(meshes, grid=int(wl[4:]), width=7680, height=4320)
else:
    s = FlatScene.load(os.path.join(ROOT, "tests", "golden", "scenes", wl + ".npz")).with_resolution(3840, 2160)
tag = os.path.basename(os.environ.get("CUTRACE_B200_LIB", "default"))
leafs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"])]
for leaf in leafs:
    for label, flags in (("smem", 0), ("global", ct.FLAG_NO_SMEM_TOP)):
        with ct.Renderer(s, flags=flags, leaf_size=leaf) as r:
            ms = []
            for i in range(5 if not wl.startswith("grid") else 3):
                st = r.render()
                ms.append((st["render_ms"], st["trace_ms"], st["shade_ms"]))
            m = np.median(np.array(ms[1:]), axis=0)
            print(f"{tag:34s} {wl} leaf={leaf} {label:6s} render={m[0]:7.3f} trace={m[1]:7.3f} shade={m[2]:7.3f} Mrays/s={st['rays_total']/m[0]/1e3:8.1f} nodes={st['bvh_nodes']} depth={st['bvh_depth']} build_ms={st['build_ms']:.2f}", flush=True)
