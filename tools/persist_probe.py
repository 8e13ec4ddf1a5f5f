#!/usr/bin/env python
"""Times library variants on bunny.json@4K and the 10 M-triangle hall@8K in one call, and hashes the colour images so that
variants which must be bit-identical can be checked against the in-tree library.
usage: tools/persist_probe.py default variants/libcutrace_b200_pers8.so ...   (paths relative to cutrace_b200/lib)"""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def child(hall_path):
    import cutrace_b200 as ct
    from cutrace_b200.scene import FlatScene

    tag = os.path.basename(os.environ.get("CUTRACE_B200_LIB", "default"))
    g = os.path.join(ROOT, "tests", "golden", "scenes")
    scenes = [("bunny4k", FlatScene.load(os.path.join(g, "bunny.npz")).with_resolution(3840, 2160), 6)]
    if hall_path:
        z = np.load(hall_path)
        kw = {k: z[k] for k in z.files}
        kw["ambient"] = float(kw["ambient"]); kw["width"] = int(kw["width"]); kw["height"] = int(kw["height"])
        scenes.append(("hall8k", FlatScene(**kw), 4))
    for name, s, frames in scenes:
        with ct.Renderer(s) as r:
            ms = [r.render()["render_ms"] for _ in range(frames)]
            col = r.download(want=("color",))["color"]
            digest = hashlib.md5(col.tobytes()).hexdigest()[:12]
        with ct.Renderer(s, flags=ct.FLAG_SERIALIZE) as r:
            r.render()
            st = r.render()
        print(f"{tag:36s} {name:8s} frame={np.median(ms[1:]):8.3f} ms  serialized: trace={st['trace_ms']:7.3f} shade={st['shade_ms']:7.3f}  "
              f"colour md5={digest}", flush=True)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None)
    libs = sys.argv[1:]
    hall = "-"
    if not os.environ.get("PROBE_NO_HALL"):
        from cutrace_b200 import synth
        from cutrace_b200.scene import FlatScene

        t0 = time.time()
        g = os.path.join(ROOT, "tests", "golden", "scenes")
        meshes = synth.meshes_from_scenes(FlatScene.load(os.path.join(g, "bunny.npz")), FlatScene.load(os.path.join(g, "mirror.npz")))[:2]
        s = synth.grid_scene(meshes, grid=106, width=7680, height=4320)
        hall = os.path.join(tempfile.gettempdir(), "cutrace_hall.npz")
        np.savez(hall, **s.to_npz_dict())
        print(f"hall scene: {s.n_triangles} triangles, generated + saved in {time.time() - t0:.1f} s", flush=True)
    for lib in libs:
        env = dict(os.environ)
        if lib != "default":
            env["CUTRACE_B200_LIB"] = os.path.join(ROOT, "cutrace_b200", "lib", lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", hall], env=env, check=False)


if __name__ == "__main__":
    main()
