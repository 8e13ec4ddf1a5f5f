#!/usr/bin/env python
"""Work distribution of the pixel kernel: one cursor per CTA (CUTRACE_PIXEL_SEG) x super-tile slot order (CUTRACE_TILE_CURVE), both read
at upload, unsharded and as rank 0 of an 8-rank tile shard: median render_ms and the md5 of colour / depth (world = 1: every
combination must give the same frame).   usage: tools/seg_probe.py [workload ...]   env: PROBE_WORLDS=1,8 PROBE_COMBOS=00,01,10,11"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import cutrace_b200 as ct  # noqa: E402

print("library", ct._lib.LIB_PATH, flush=True)
for wl in sys.argv[1:] or ["synthetic10m"]:
    scene, _ = bench.load_workload(wl)
    for world in [int(x) for x in os.environ.get("PROBE_WORLDS", "1,8").split(",")]:
        for combo in os.environ.get("PROBE_COMBOS", "00,01,10,11").split(","):
            for key, v in (("CUTRACE_PIXEL_SEG", combo[0]), ("CUTRACE_TILE_CURVE", combo[1])):   # "d": the library's default
                if v == "d":
                    os.environ.pop(key, None)
                else:
                    os.environ[key] = v
            with ct.Renderer(scene, tile_rank=0, tile_world=world) as r:
                ms = [r.render()["render_ms"] for _ in range(5 if wl == "synthetic10m" else 9)]
                st = r.render()
                out = r.download(want=("color", "depth"))
            md5 = hashlib.md5(out["color"].tobytes()).hexdigest()[:10] + "/" + hashlib.md5(out["depth"].tobytes()).hexdigest()[:10]
            print(f"{wl:13s} world={world} seg={combo[0]} curve={combo[1]} render={np.median(ms[1:]):9.4f} ms (min {min(ms[1:]):.4f})  rays={st['rays_total']:>11d}  md5={md5}", flush=True)
