#!/usr/bin/env python
"""Generates tests/golden/ from the reference, run in the build container (needs /root/reference).

  tests/golden/scenes/<name>.npz   the four reference scenes flattened to cutrace_scene_desc arrays
                                   (JSON + STL read in place from $CUTRACE_REF/scene; derived data,
                                   no reference source is copied)
  tests/golden/<name>_<w>x<h>.npz  depth / normal / colour / hit_id produced by the REFERENCE's own
                                   source compiled for the host (oracle/_ref/libcutrace_ref_host.so,
                                   see oracle/ref_host.cpp) at a small resolution

The reference ships no golden vectors for the render path; these files are the pin.
Re-run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cutrace_b200.scene import load_scene_json  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

REF = os.environ.get("CUTRACE_REF", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {  # name -> golden resolution (None = as in the file)
    "triangle": None,
    "sphere_plane": (160, 90),
    "mirror": (160, 90),
    "bunny": (96, 54),
}


def main():
    po.build()
    os.makedirs(os.path.join(GOLD, "scenes"), exist_ok=True)
    for name, res in CASES.items():
        s = load_scene_json(os.path.join(REF, "scene", f"{name}.json"), base_dir=REF)
        s.save(os.path.join(GOLD, "scenes", f"{name}.npz"))
        g = s.with_resolution(*res) if res else s
        out = po.ref_host_render(g)
        path = os.path.join(GOLD, f"{name}_{g.width}x{g.height}.npz")
        np.savez_compressed(path, width=np.uint32(g.width), height=np.uint32(g.height), fudge=np.float32(1e-3),
                            bounces=np.uint32(5), depth=out["depth"], normal=out["normal"], color=out["color"],
                            hit_id=out["hit_id"])
        print(f"{name}: {s.n_triangles} tris, {s.n_spheres} spheres, {s.n_planes} planes -> {os.path.relpath(path, ROOT)}")


if __name__ == "__main__":
    main()
