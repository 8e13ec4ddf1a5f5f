"""Differential fuzz of the two scene front-ends (C++ host library vs Python mirror): mutated files must be accepted or rejected
by both, with the same triangle count, and never crash.  usage: tools/fuzz_*_frontends.py [seed] [iterations]"""
import os, random, struct, sys, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from cutrace_b200 import host
from cutrace_b200.scene import SceneError, read_mesh
host.load()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
d = tempfile.mkdtemp()
doc = json.load(open(ROOT + "/scenes/solids.json"))
srcs = {e: open(ROOT + "/scenes/" + f, "rb").read() for e, f in (("stl_a", "tetra.stl"), ("stl_b", "octa.stl"), ("obj", "cube.obj"))}
# PLY (three encodings) and OFF seeds: the cube + cap of tests/test_host_frontend.py::test_ply_and_off_mesh_import
sys.path.insert(0, ROOT + "/tests")
from test_host_frontend import _ply_bytes  # noqa: E402
_V = [(-0.5, -0.5, -0.5), (0.5, -0.5, -0.5), (0.5, 0.5, -0.5), (-0.5, 0.5, -0.5), (-0.5, -0.5, 0.5), (0.5, -0.5, 0.5), (0.5, 0.5, 0.5), (-0.5, 0.5, 0.5), (0.1, 1.25, 0.3)]
_F = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3), (3, 8, 2), (7, 6, 2, 8, 3)]
for enc in ("ascii", "binary_little_endian", "binary_big_endian"):
    srcs["ply_" + enc] = _ply_bytes(_V, _F, enc)
srcs["off"] = ("OFF\n%d %d 0\n" % (len(_V), len(_F)) + "".join("%r %r %r\n" % v for v in _V) + "".join("%d %s\n" % (len(f), " ".join(map(str, f))) for f in _F)).encode()
ok = err = mism = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1500):
    kind = rng.choice(list(srcs))
    b = bytearray(srcs[kind])
    for _ in range(rng.randint(1, 3)):
        k = rng.randrange(4)
        pos = rng.randrange(len(b))
        if k == 0: del b[pos:pos + rng.randint(1, 60)]
        elif k == 1: b[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 8)))
        elif k == 2: b[pos] = rng.randrange(256)
        else:
            if len(b) >= 84: b[80:84] = struct.pack("<I", rng.choice([0, 1, 2**31, 2**32 - 1, rng.randrange(100)]))
    ext = ".obj" if kind == "obj" else ".off" if kind == "off" else ".ply" if kind.startswith("ply") else ".stl"
    mp = os.path.join(d, "m" + ext)
    open(mp, "wb").write(bytes(b))
    doc2 = {"objects": [{"type": "mesh", "file": mp, "material": 0}], "lights": doc["lights"], "materials": doc["materials"], "camera": doc["camera"]}
    p = os.path.join(d, "f.json")
    json.dump(doc2, open(p, "w"))
    a = bpy = None
    try:
        a = host.load_scene(p, base_dir="/")
        ok += 1
    except SceneError:
        err += 1
    try:
        from cutrace_b200.scene import load_scene_json
        bpy = load_scene_json(p, base_dir="/")
    except SceneError:
        bpy = None
    if (a is None) != (bpy is None) or (a is not None and a.n_triangles != bpy.n_triangles):
        mism += 1
        if mism <= 5:
            print("front-ends disagree:", kind, "cpp", None if a is None else a.n_triangles, "py", None if bpy is None else bpy.n_triangles, len(b))
            open(f"/tmp/cutrace_mism{mism}{ext}", "wb").write(bytes(b))
print("accepted", ok, "rejected", err, "front-end disagreements", mism)
