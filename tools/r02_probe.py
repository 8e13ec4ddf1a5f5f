#!/usr/bin/env python
"""Round-2 scheduler probe on one GPU: every BASELINE workload, unsharded and as a 1/8 tile shard, with the persistent frame
kernel (default) and with the round-1 multi-launch graph (CUTRACE_SCHEDULER=launches).  Prints median render_ms, the frame
kernel's phase times and an md5 of the colour image (the schedulers must agree bit for bit on non-branching scenes).
usage: tools/r02_probe.py [workload ...]      (child mode: --child <mode> <workload> <world>)"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def child(workload, world, frames):
    import bench
    import cutrace_b200 as ct

    scene, wl = bench.load_workload(workload)
    tag = {"launches": "multi-launch", "frame": "frame-kernel", "pixel": "pixel-kernel"}[os.environ["CUTRACE_SCHEDULER"]]
    with ct.Renderer(scene, tile_rank=0, tile_world=world, flags=int(os.environ.get("PROBE_FLAGS", "0"))) as r:
        ms = []
        for _ in range(frames):
            st = r.render()
            ms.append(st["render_ms"])
        ph = r.phase_ms()
        out = r.download(want=("color",))
        digest = hashlib.md5(out["color"].tobytes()).hexdigest()[:10]
    print(f"{workload:13s} world={world} {tag:12s} render={np.median(ms[1:]):9.4f} ms (min {min(ms[1:]):.4f})  launches={st['kernel_launches']:2d} "
          f"rays={st['rays_total']:>11d}  md5={digest}  phases={' '.join(f'{x:.3f}' for x in ph)}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    wls = sys.argv[1:] or ["triangle", "spheres1080", "mirror1080", "bunny4k", "synthetic10m"]
    for wl in wls:
        for world in [int(x) for x in os.environ.get("PROBE_WORLDS", "1,8").split(",")]:
            if wl == "triangle" and world > 1:
                continue
            for sched in os.environ.get("PROBE_SCHEDS", "frame,launches,pixel").split(","):
                env = dict(os.environ)
                env["CUTRACE_SCHEDULER"] = sched
                frames = 5 if wl == "synthetic10m" else 9
                subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__), "--child", wl, str(world), str(frames)], env=env, check=False)


if __name__ == "__main__":
    main()
