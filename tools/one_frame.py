#!/usr/bin/env python
"""Renders a few frames of one bench.py workload through the C-ABI (for ncu): tools/one_frame.py <workload> [frames] [world] [flags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cutrace_b200 as ct  # noqa: E402

wl = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
scene, w = bench.load_workload(wl)
with ct.Renderer(scene, tile_rank=0, tile_world=world, flags=flags) as r:
    for _ in range(frames):
        st = r.render()
print(f"{w['label']} world={world} flags={flags}: render {st['render_ms']:.3f} ms, {st['kernel_launches']} launches, scheduler {st['scheduler']}, rays {st['rays_total']}")
