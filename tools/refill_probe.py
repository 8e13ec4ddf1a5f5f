#!/usr/bin/env python
"""Pixel kernel with and without lane refill (CUTRACE_PIXEL_REFILL, read at every upload), unsharded and as rank 0 of an 8-rank tile
shard, one process per library: median render_ms, rays and the md5 of the colour / depth images (refill must not change a bit).
needs a -DCTB_PIXEL_REFILL=1 build (tools/build_variant.sh refill -DCTB_PIXEL_REFILL=1; CUTRACE_B200_LIB=...): the product build compiles the refill loop out.
usage: tools/refill_probe.py [workload ...]     env: PROBE_REFILL=0,1  PROBE_WORLDS=1,8"""
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
        for refill in os.environ.get("PROBE_REFILL", "0,1").split(","):
            if refill == "default":
                os.environ.pop("CUTRACE_PIXEL_REFILL", None)
            else:
                os.environ["CUTRACE_PIXEL_REFILL"] = refill
            with ct.Renderer(scene, tile_rank=0, tile_world=world) as r:
                ms = [r.render()["render_ms"] for _ in range(5 if wl == "synthetic10m" else 9)]
                st = r.render()
                out = r.download(want=("color", "depth"))
            md5 = hashlib.md5(out["color"].tobytes()).hexdigest()[:10] + "/" + hashlib.md5(out["depth"].tobytes()).hexdigest()[:10]
            print(f"{wl:13s} world={world} refill={refill:7s} render={np.median(ms[1:]):9.4f} ms (min {min(ms[1:]):.4f})  rays={st['rays_total']:>11d}  md5={md5}", flush=True)
