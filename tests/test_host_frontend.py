"""The callers either side of the path: the C++ scene front-end (JSON + STL), the JPEG writer and
the CLI's exit codes.  No GPU needed."""
import io
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CASES, REF, ROOT, load_golden_scene


@pytest.fixture(scope="module")
def host():
    from cutrace_b200 import host as h

    if not os.path.exists(h.HOST_LIB_PATH) or not os.path.exists(os.path.join(ROOT, "bin", "cutrace")):
        subprocess.run(["make", "-C", ROOT, "host", "cli"], check=True, stdout=subprocess.DEVNULL)
    h.load()
    return h


def _same(a, b):
    da, db = a.to_npz_dict(), b.to_npz_dict()
    return [k for k in da if not np.array_equal(np.asarray(da[k]), np.asarray(db[k]))]


def test_cpp_loader_equals_python_mirror_on_own_scene(host):
    from cutrace_b200.scene import load_scene_json

    path = os.path.join(ROOT, "scenes", "solids.json")
    a = host.load_scene(path, base_dir=ROOT)
    b = load_scene_json(path, base_dir=ROOT)
    assert _same(a, b) == []
    assert a.n_triangles == 4 + 8 + 1 and a.n_spheres == 1 and a.n_planes == 1 and a.n_objects == 5   # ASCII + binary STL + loose triangle
    assert a.obj_kind.tolist() == [1, 1, 2, 3, 0]
    assert a.mat_specular.tolist() == pytest.approx([0.5, 0.3, 0.2, 0.7])      # default 0.3 (default_schema.hpp:754-764)
    assert a.mat_phong.tolist() == [64, 32, 300, 500] and a.mat_transparency.tolist() == [0, 0, 0, 0.5]
    assert a.light_color[1].tolist() == [1, 1, 1]                                # default white
    # look_at: unit, orthogonal basis
    f, r, u = a.cam_forward, a.cam_right, a.cam_up
    assert abs(np.dot(f, r)) < 1e-6 and abs(np.dot(f, u)) < 1e-6 and abs(np.linalg.norm(u) - 1) < 1e-6


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scene")), reason="reference tree not present")
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_cpp_loader_reads_the_reference_scenes(host, name):
    a = host.load_scene(os.path.join(REF, "scene", f"{name}.json"), base_dir=REF)
    assert _same(a, load_golden_scene(name)) == []


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scene")), reason="reference tree not present")
def test_stale_schema_is_rejected_like_the_reference(host):
    """scene/bunny_small.json uses schema.md's stale spelling; default_schema.hpp rejects it."""
    from cutrace_b200.scene import SceneError

    p = os.path.join(REF, "scene", "bunny_small.json")
    with pytest.raises(SceneError) as e:
        host.load_scene(p, base_dir=REF)
    assert "material #0" in str(e.value) and "light #0" in str(e.value)
    s = host.load_scene(p, base_dir=REF, accept_aliases=True)      # documented superset
    assert s.n_triangles == 1000


BAD = {
    "not json": "{ this is not json",
    "missing camera": {"objects": [], "lights": [], "materials": []},
    "camera key missing": {"objects": [], "lights": [], "materials": [], "camera": {"eye": [0, 0, 0]}},
    "bad material index": {"objects": [{"type": "sphere", "center": [0, 0, 0], "radius": 1, "material": 3}], "lights": [],
                           "materials": [{"type": "solid", "color": [1, 1, 1]}],
                           "camera": {"eye": [0, 0, -5], "up": [0, 1, 0], "look": [0, 0, 0], "near_plane": 0.1, "far_plane": 10, "width": 8, "height": 8, "ambient": 0.1}},
    "unknown object type": {"objects": [{"type": "torus", "material": 0}], "lights": [], "materials": [{"type": "solid", "color": [1, 1, 1]}],
                            "camera": {"eye": [0, 0, -5], "up": [0, 1, 0], "look": [0, 0, 0], "near_plane": 0.1, "far_plane": 10, "width": 8, "height": 8, "ambient": 0.1}},
    "vector of two": {"objects": [{"type": "plane", "point": [0, 0], "normal": [0, 1, 0], "material": 0}], "lights": [],
                      "materials": [{"type": "solid", "color": [1, 1, 1]}],
                      "camera": {"eye": [0, 0, -5], "up": [0, 1, 0], "look": [0, 0, 0], "near_plane": 0.1, "far_plane": 10, "width": 8, "height": 8, "ambient": 0.1}},
    "missing mesh file": {"objects": [{"type": "mesh", "file": "does/not/exist.stl", "material": 0}], "lights": [],
                          "materials": [{"type": "solid", "color": [1, 1, 1]}],
                          "camera": {"eye": [0, 0, -5], "up": [0, 1, 0], "look": [0, 0, 0], "near_plane": 0.1, "far_plane": 10, "width": 8, "height": 8, "ambient": 0.1}},
}


@pytest.mark.parametrize("what", list(BAD))
def test_rejected_scenes_and_cli_exit_codes(host, tmp_path, what):
    from cutrace_b200.scene import SceneError, load_scene_json

    doc = BAD[what]
    p = tmp_path / "scene.json"
    p.write_text(doc if isinstance(doc, str) else json.dumps(doc))
    with pytest.raises(SceneError):
        host.load_scene(str(p))
    with pytest.raises(SceneError):
        load_scene_json(str(p))
    r = subprocess.run([os.path.join(ROOT, "bin", "cutrace"), str(p)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 254            # exit -2 like main.cu:16-19
    assert r.stderr.strip() and "Scene schema" in r.stdout


def test_cli_usage(host):
    r = subprocess.run([os.path.join(ROOT, "bin", "cutrace")], capture_output=True, text=True)
    assert r.returncode == 255 and "Usage:" in r.stderr and "<scene file>" in r.stderr     # main.cu:9-12


def test_jpeg_writer_decodes_like_a_q90_420_jpeg(host, tmp_path, oracle):
    from PIL import Image

    s = host.load_scene(os.path.join(ROOT, "scenes", "solids.json"), base_dir=ROOT).with_resolution(203, 117)   # not a multiple of 16
    o = oracle.oracle_render(s)
    d8, n8, c8 = oracle.encode_bytes(o["depth"], o["normal"], o["color"], oracle.max_depth(o["depth"]))
    for name, img in (("frame", c8), ("normal_map", n8), ("depth_map", d8)):
        rgb = img.reshape(s.height, s.width, 3)
        path = str(tmp_path / f"{name}.jpg")
        host.write_jpeg(path, rgb, 90)
        im = Image.open(path)
        assert im.size == (s.width, s.height) and im.mode == "RGB"
        got = np.asarray(im.convert("RGB")).astype(np.float64)
        buf = io.BytesIO()
        Image.fromarray(rgb).save(buf, "JPEG", quality=90, subsampling=2)
        pil = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.float64)
        psnr = lambda a: 10 * np.log10(255.0 ** 2 / max(((a - rgb) ** 2).mean(), 1e-12))   # noqa: E731
        assert psnr(got) > psnr(pil) - 1.5, (name, psnr(got), psnr(pil))        # as good as libjpeg at the same settings
        assert 0.6 < os.path.getsize(path) / len(buf.getvalue()) < 1.6


def test_look_at_matches_python(host):
    import ctypes as C

    from cutrace_b200.scene import look_at

    lib = host.load()
    rng = np.random.default_rng(3)
    for _ in range(20):
        pos, up, look = (rng.normal(size=3).astype(np.float32) for _ in range(3))
        out = [(C.c_float * 3)() for _ in range(3)]
        arr = [(C.c_float * 3)(*v) for v in (pos, up, look)]
        lib.cutrace_host_look_at(arr[0], arr[1], arr[2], out[0], out[1], out[2])
        f, r, u = look_at(pos, up, look)
        assert np.array_equal(np.array(out[0][:], np.float32), f) and np.array_equal(np.array(out[1][:], np.float32), r)
        assert np.array_equal(np.array(out[2][:], np.float32), u)


def test_obj_mesh_import(host, oracle):
    """Mesh import breadth: Wavefront OBJ with quads, a/b/c index forms and negative indices,
    fan-triangulated in face order — the C++ loader and the Python mirror agree, and the cube renders closed."""
    from cutrace_b200.scene import load_scene_json, read_obj

    path = os.path.join(ROOT, "scenes", "solids_obj.json")
    a = host.load_scene(path, base_dir=ROOT)
    b = load_scene_json(path, base_dir=ROOT)
    assert _same(a, b) == []
    cube = read_obj(os.path.join(ROOT, "scenes", "cube.obj"))
    assert cube.shape == (12, 3, 3)                                   # 6 quads -> 12 triangles
    assert a.n_triangles == 13 + 12 and a.obj_kind.tolist()[-1] == 1
    # every cube face normal of the fan triangulation points outward (consistent winding kept)
    n = np.cross(cube[:, 1] - cube[:, 0], cube[:, 2] - cube[:, 0])
    c = cube.mean(axis=1) - cube.reshape(-1, 3).mean(axis=0)
    assert np.all((n * c).sum(axis=1) > 0)
    out = oracle.oracle_render(a.with_resolution(96, 60))
    assert (out["hit_id"] == 5).sum() > 20                            # the cube (object #5) is visible


def test_synthetic_grid_as_json_and_stl_files(host, tmp_path):
    """SURVEY.md §8d config 5, "also emit a small G=3 version as JSON + STL files to prove schema drop-in": the files
    are in the reference's schema, and both front-ends read them back bit-identical to the in-memory scene."""
    from cutrace_b200 import synth
    from cutrace_b200.scene import load_scene_json

    meshes = synth.meshes_from_scenes(load_golden_scene("bunny"), load_golden_scene("mirror"))[:2]
    s = synth.grid_scene(meshes, grid=3, width=96, height=54)
    path = synth.write_scene_files(s, str(tmp_path), *synth.grid_camera(3), name="grid3")
    doc = json.load(open(path))
    assert set(doc) == {"objects", "lights", "materials", "camera"}
    assert set(doc["camera"]) == {"eye", "up", "look", "near_plane", "far_plane", "width", "height", "ambient"}   # all mandatory
    assert [o["type"] for o in doc["objects"]] == ["mesh"] * 9 + ["plane"] * 6
    assert all(m["type"] == "solid" for m in doc["materials"]) and all(li["type"] == "point" for li in doc["lights"])
    assert os.path.getsize(tmp_path / "grid3_mesh0000.stl") == 84 + 50 * 1000
    assert _same(s, load_scene_json(path, base_dir=str(tmp_path))) == []
    assert _same(s, host.load_scene(path, base_dir=str(tmp_path))) == []

    # every object kind and both light kinds
    soup = synth.random_soup(n_tri=80, n_sph=3, n_planes=2, n_lights=2, width=40, height=30, seed=2)
    path = synth.write_scene_files(soup, str(tmp_path), [0, 0.3, -5], [0, 1, 0], [0, 0, 0], name="soup")
    assert _same(soup, load_scene_json(path, base_dir=str(tmp_path))) == []
    assert _same(soup, host.load_scene(path, base_dir=str(tmp_path))) == []


def test_json_numbers_follow_picojson_rule(host, tmp_path):
    """The reference parses scene files with picojson: a number is the longest run of [0-9+-.eE] that strtod consumes
    entirely.  So "041" and "0." load (strict JSON would refuse them), "0x10" / "-inf" / a raw control character in a
    string do not."""
    from cutrace_b200.scene import SceneError

    src = open(os.path.join(ROOT, "scenes", "solids.json")).read()
    assert '"near_plane": 0.1' in src and '"ambient"' in src

    def load(text):
        p = tmp_path / "s.json"
        p.write_text(text)
        return host.load_scene(str(p), base_dir=ROOT)

    base = load(src)
    assert _same(base, load(src.replace('"near_plane": 0.1', '"near_plane": 041'))) == []      # near_plane is dead on the path
    assert _same(base, load(src.replace('"near_plane": 0.1', '"near_plane": 0.'))) == []
    assert _same(base, load(src.replace('"near_plane": 0.1', '"near_plane": 1.e-1'))) == []
    for bad in ('0x10', '-inf', 'nan', '1e', '--1', '.5'):
        with pytest.raises(SceneError):
            load(src.replace('"near_plane": 0.1', f'"near_plane": {bad}'))
    with pytest.raises(SceneError):
        load(src.replace('"near_plane"', '"near\tplane"'))


def test_mesh_tokens_must_be_whole_numbers(host, tmp_path):
    """Both front-ends read ASCII STL / OBJ as byte tokens split at ASCII white space, and a coordinate or index token has to
    be one number in full — "0.5x", a hex float or an embedded NUL is a malformed file, not a truncated value."""
    from cutrace_b200.scene import SceneError, read_mesh

    stl = open(os.path.join(ROOT, "scenes", "tetra.stl"), "rb").read()
    obj = open(os.path.join(ROOT, "scenes", "cube.obj"), "rb").read()
    assert b"vertex" in stl and b"\nf " in obj
    doc = json.load(open(os.path.join(ROOT, "scenes", "solids.json")))

    def both(name, data):
        path = tmp_path / name
        path.write_bytes(data)
        scene = {"objects": [{"type": "mesh", "file": str(path), "material": 0}], "lights": doc["lights"],
                 "materials": doc["materials"], "camera": doc["camera"]}
        sp = tmp_path / "s.json"
        sp.write_text(json.dumps(scene))
        out = []
        for fn in (lambda: host.load_scene(str(sp), base_dir="/").n_triangles, lambda: len(read_mesh(str(path)))):
            try:
                out.append(fn())
            except SceneError:
                out.append(None)
        return out

    assert both("ok.stl", stl) == [4, 4] and both("ok.obj", obj) == [12, 12]
    first = stl.index(b"vertex") + len(b"vertex ")
    num_end = stl.index(b" ", first)
    assert both("x.stl", stl[:num_end] + b"x" + stl[num_end:]) == [None, None]                 # "0.000000x"
    assert both("hex.stl", stl[:first] + b"0x1p0" + stl[num_end:]) == [None, None]
    assert both("nel.stl", stl[:num_end] + b"\x85" + stl[num_end:]) == [None, None]            # NEL is not white space
    assert both("split.stl", stl.replace(b"vertex ", b"vertex\n", 1)) == [4, 4]               # a token stream, not lines
    face = obj.index(b"\nf ") + 3
    assert both("idx.obj", obj[:face] + b"1x " + obj[face:]) == [None, None]
    assert both("nul.obj", obj[:face + 1] + b"\x00" + obj[face + 1:]) == [None, None]
    assert both("cr.obj", obj.replace(b"\n", b"\r\n")) == [12, 12]                             # CR is white space


def _ply_bytes(verts, faces, fmt):
    """A PLY file with extra vertex properties, an extra list property per face and a trailing extra element — everything a reader
    has to skip.  fmt: "ascii" | "binary_little_endian" | "binary_big_endian"."""
    import struct

    head = ["ply", f"format {fmt} 1.0", "comment written by the test", f"element vertex {len(verts)}", "property double x", "property float y",
            "property float z", "property uchar red", "property short flag", f"element face {len(faces)}", "property uchar kind",
            "property list uchar int vertex_indices", "property list ushort float weights", "element edge 2", "property int a", "property int b",
            "end_header"]
    out = ("\n".join(head) + "\n").encode()
    if fmt == "ascii":
        body = [f"{x!r} {y!r} {z!r} 200 -3" for x, y, z in verts]
        body += [f"7 {len(f)} " + " ".join(map(str, f)) + " 2 0.5 0.25" for f in faces]
        body += ["0 1", "1 2"]
        return out + ("\n".join(body) + "\n").encode()
    e = "<" if fmt == "binary_little_endian" else ">"
    for x, y, z in verts:
        out += struct.pack(e + "dffBh", x, y, z, 200, -3)
    for f in faces:
        out += struct.pack(e + "BB" + "i" * len(f), 7, len(f), *f) + struct.pack(e + "Hff", 2, 0.5, 0.25)
    return out + struct.pack(e + "iiii", 0, 1, 1, 2)


def test_ply_and_off_mesh_import(host, tmp_path):
    """Mesh import breadth (the reference reads whatever Assimp reads, inc/default_schema.hpp:516-545): Stanford PLY
    in its three encodings and OFF, polygons fan-triangulated in face order.  The C++ loader, the Python mirror and the expected
    triangles agree bit for bit; malformed files are errors in both."""
    from cutrace_b200.scene import SceneError, load_scene_json, read_mesh

    verts = [(-0.5, -0.5, -0.5), (0.5, -0.5, -0.5), (0.5, 0.5, -0.5), (-0.5, 0.5, -0.5), (-0.5, -0.5, 0.5), (0.5, -0.5, 0.5),
             (0.5, 0.5, 0.5), (-0.5, 0.5, 0.5), (0.1, 1.25, 0.3)]
    faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3), (3, 8, 2), (7, 6, 2, 8, 3)]
    v = np.asarray(verts, np.float64).astype(np.float32)
    want = np.asarray([[v[f[0]], v[f[k]], v[f[k + 1]]] for f in faces for k in range(1, len(f) - 1)], np.float32)
    files = {}
    for fmt in ("ascii", "binary_little_endian", "binary_big_endian"):
        files[fmt] = tmp_path / f"solid_{fmt}.ply"
        files[fmt].write_bytes(_ply_bytes(verts, faces, fmt))
    off = ["OFF  # a comment", f"{len(verts)} {len(faces)} 0", ""] + [f"{x!r} {y!r} {z!r}" for x, y, z in verts]
    off += ["# faces", *(f"{len(f)} " + " ".join(map(str, f)) + "  0.5 0.5 0.5" for f in faces)]
    files["off"] = tmp_path / "solid.off"
    files["off"].write_text("\n".join(off) + "\n")
    for name, path in files.items():
        got = read_mesh(str(path))
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    cam = {"near_plane": 0.1, "far_plane": 100.0, "eye": [0.5, 1.2, -4.5], "up": [0, 1, 0], "look": [0, 0, 0], "width": 64, "height": 40, "ambient": 0.03}
    doc = {"camera": cam, "materials": [{"type": "solid", "color": [0.8, 0.3, 0.2]}], "lights": [{"type": "sun", "direction": [0, -1, 0.5]}],
           "objects": [{"type": "mesh", "file": p.name, "material": 0} for p in files.values()]}
    scene = tmp_path / "solids_ply_off.json"
    scene.write_text(json.dumps(doc))
    a = host.load_scene(str(scene), base_dir=str(tmp_path))
    b = load_scene_json(str(scene), base_dir=str(tmp_path))
    assert _same(a, b) == []
    assert a.n_triangles == 4 * len(want) and a.obj_kind.tolist() == [1, 1, 1, 1]
    for k in range(4):
        sel = slice(k * len(want), (k + 1) * len(want))
        assert np.array_equal(np.stack([a.tri_p1[sel], a.tri_p2[sel], a.tri_p3[sel]], axis=1), want)
    # malformed inputs: errors in both front-ends
    bad = {
        "truncated.ply": _ply_bytes(verts, faces, "binary_little_endian")[:-40],
        "index.ply": _ply_bytes(verts, [(0, 1, 99)], "ascii"),
        "noface.ply": b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n0 0 0\n",
        "header.ply": b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\n",
        "index.off": b"OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 3\n",
        "short.off": b"OFF\n3 2 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n",
    }
    for name, blob in bad.items():
        (tmp_path / name).write_bytes(blob)
        doc["objects"] = [{"type": "mesh", "file": name, "material": 0}]
        scene.write_text(json.dumps(doc))
        with pytest.raises(SceneError):
            load_scene_json(str(scene), base_dir=str(tmp_path))
        with pytest.raises(SceneError):
            host.load_scene(str(scene), base_dir=str(tmp_path))


def test_mesh_front_ends_agree_on_mutated_files(host):
    """A short, seeded run of tools/fuzz_mesh_frontends.py (mutated STL / OBJ / PLY / OFF files): the C++ loader and the Python mirror
    accept or reject the same files with the same triangle count, and neither crashes."""
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_mesh_frontends.py"), "11", "400"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.endswith("front-end disagreements 0"), r.stdout[-2000:]
    assert int(last.split()[1]) > 10 and int(last.split()[3]) > 100      # both outcomes occur


def test_json_front_ends_agree_on_mutated_scenes(host):
    """A short, seeded run of tools/fuzz_json_frontends.py: mutated scene files are accepted or rejected by both front-ends, with
    identical arrays when accepted — including picojson's lax numbers ("041", "0.", "1.e5"), which the Python mirror rewrites
    before json.loads (scene.py: _picojson_numbers) and the C++ parser reads with the same rule (host/json.hpp)."""
    import sys

    from cutrace_b200.scene import SceneError, _picojson_numbers

    assert json.loads(_picojson_numbers('{"a": 0., "b": "0. \\\\\\" 041", "c": 041, "d": -1.e5, "e": [1, 2.5]}')) == {"a": 0.0, "b": '0. \\" 041', "c": 41.0, "d": -1e5, "e": [1, 2.5]}
    for bad in ('{"a": 1e}', '{"a": 1-2}', '{"a": -}', '{"a": 1.2.3}'):
        with pytest.raises(SceneError):
            _picojson_numbers(bad)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_json_frontends.py"), "3", "400"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "agree 400 disagree 0", r.stdout[-2000:]
