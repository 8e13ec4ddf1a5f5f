#!/usr/bin/env python
"""Renders one tile shard (rank 0 of --world) of a workload on one GPU; used under `ncu --metrics gpu__time_duration.sum`
to see how the per-level kernel times change with the shard size."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cutrace_b200 as ct
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="synthetic10m")
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--serial-only", action="store_true")
a = ap.parse_args()
scene, wl = bench.load_workload(a.workload)
for label, flags in ((("serial", ct.FLAG_SERIALIZE),) if a.serial_only else (("overlap", 0), ("serial", ct.FLAG_SERIALIZE))):
    with ct.Renderer(scene, tile_rank=a.rank, tile_world=a.world, flags=flags) as r:
        for i in range(a.frames):
            st = r.render()
        print(f"world={a.world} rank={a.rank} {label:8s} render={st['render_ms']:.3f} trace={st['trace_ms']:.3f} shade={st['shade_ms']:.3f} rays={st['rays_total']}")
