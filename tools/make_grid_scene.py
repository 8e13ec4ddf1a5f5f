#!/usr/bin/env python
"""Writes the synthetic grid scene (BASELINE.json config 5) as files a stock cutrace reads:
<out>/<name>.json in the reference's schema + one binary STL per instance.

    python tools/make_grid_scene.py --grid 3 --out /tmp/grid3          # 9 instances, 8,200 triangles
    (cd /tmp/grid3 && /root/repo/bin/cutrace grid3.json)                # mesh paths are relative to the CWD, as in the reference

--grid 106 is the 10,112,400-triangle scene of the bench (11,236 STL files, ~0.5 GB).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cutrace_b200 import synth                     # noqa: E402
from cutrace_b200.scene import FlatScene           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=3)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--out", required=True)
    ap.add_argument("--name", default=None)
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    golden = os.path.join(root, "tests", "golden", "scenes")
    meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(golden, "bunny.npz")), FlatScene.load(os.path.join(golden, "mirror.npz")))[:2]
    s = synth.grid_scene(meshes, grid=a.grid, width=a.width, height=a.height, seed=0)
    path = synth.write_scene_files(s, a.out, *synth.grid_camera(a.grid), name=a.name or f"grid{a.grid}")
    print(f"{path}: {s.n_objects} objects, {s.n_triangles} triangles, {s.n_planes} planes, {s.n_lights} lights")


if __name__ == "__main__":
    main()
