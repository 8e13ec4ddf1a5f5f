"""Flat structure-of-arrays scene: the host-side image of ``cutrace_scene_desc`` (include/cutrace.h).

This mirrors what the reference builds with ``default_schema::load_file`` + ``default_to_gpu``
(/root/reference/inc/loader.hpp:763-780, inc/default_schema.hpp:487-940, inc/cpu_to_gpu.hpp:188-198)
but stores it as numpy arrays that map 1:1 onto the C-ABI struct.

The JSON reader accepts exactly what ``default_schema.hpp`` accepts: objects
``triangle{p1,p2,p3,material}``, ``mesh{file,material}``, ``plane{point,normal,material}``,
``sphere{center,radius,material}``; lights ``sun{direction,color}``, ``point{point,color}``;
materials ``solid{color,specular=0.3,reflect=0,phong=32,transparency=0}``; a camera with all eight
keys mandatory.  Numbers are parsed as double and then cast (inc/json_helpers.hpp:88-93).
The native C++ loader used by the CLI lives in cutrace_b200/host/; this Python mirror exists so the
parity tests can build scenes without a compiler in the loop.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import struct
from dataclasses import dataclass, field

import numpy as np

ABI_VERSION = 1
LIGHT_SUN, LIGHT_POINT = 0, 1
NO_HIT = 0xFFFFFFFF
TILE = 16   # CUTRACE_TILE of include/cutrace.h (tests check it against cutrace_tile_size())

# object kinds = variant order of default_gpu_object (inc/default_schema.hpp:920)
OBJ_TRIANGLE, OBJ_MESH, OBJ_PLANE, OBJ_SPHERE = 0, 1, 2, 3


class SceneError(ValueError):
    """Raised where the reference loader reports an error and exits -2 (main.cu:16-19)."""


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _u32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint32))


def _normalized(v):
    """vector::normalized(): v * (1.0f / sqrtf(x*x+y*y+z*z)) in float32 (inc/vector.hpp:77-92)."""
    v = v.astype(np.float32)
    s = np.float32(v[0] * v[0]) + np.float32(v[1] * v[1])
    s = np.float32(s + np.float32(v[2] * v[2]))
    inv = np.float32(1.0) / np.sqrt(s, dtype=np.float32)
    return (v * inv).astype(np.float32)


def _cross(a, b):
    """vector::cross in float32 (inc/vector.hpp:65-71)."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    return np.array(
        [
            np.float32(a[1] * b[2]) - np.float32(a[2] * b[1]),
            np.float32(a[2] * b[0]) - np.float32(a[0] * b[2]),
            np.float32(a[0] * b[1]) - np.float32(a[1] * b[0]),
        ],
        dtype=np.float32,
    )


def look_at(pos, up, look):
    """cam::look_at (inc/default_schema.hpp:370-374). Returns (forward, right, up)."""
    pos, up, look = _f32(pos), _f32(up), _f32(look)
    forward = _normalized(look - pos)
    right = _normalized(_cross(forward, up))
    up2 = _normalized(_cross(right, forward))
    return forward, right, up2


_NUMBER = re.compile(rb"[+-]?((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?|inf|infinity|nan)\Z", re.IGNORECASE)
_INDEX = re.compile(rb"[+-]?\d+\Z")


def _mesh_number(tok):
    """One whole token (bytes) = one number, same grammar as parse_number() in cutrace_b200/host/scene_loader.cpp."""
    if not _NUMBER.match(tok):
        raise ValueError(tok)
    return float(tok)


def read_stl(path):
    """Binary or ASCII STL -> (n,3,3) float32 vertex array in file order.

    The reference imports meshes with Assimp (inc/default_schema.hpp:516-545) and reads only
    mVertices/mFaces; for STL that is the facet list in file order, stored normals ignored.
    """
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        # the reference silently loads an EMPTY mesh when Assimp cannot read the file
        # (inc/default_schema.hpp:522 `if(scene == nullptr) return;`); here that is an error
        raise SceneError(f"cannot open mesh file {path!r}") from e
    if len(data) >= 84:
        (n,) = struct.unpack_from("<I", data, 80)
        if 84 + 50 * n == len(data):
            rec = np.frombuffer(data, dtype=np.uint8, count=50 * n, offset=84).reshape(n, 50)
            return rec[:, 12:48].copy().view("<f4").reshape(n, 3, 3).astype(np.float32)
    # ASCII
    verts = []
    toks = data.split()      # bytes: ASCII white space only.  A token stream, like the C++ reader: "vertex" + three numbers
    k = 0
    while k < len(toks):
        if toks[k] == b"vertex":
            try:
                verts.append([_mesh_number(toks[k + 1]), _mesh_number(toks[k + 2]), _mesh_number(toks[k + 3])])
            except (IndexError, ValueError) as e:
                raise SceneError(f"bad vertex in ASCII STL {path!r}") from e
            k += 4
        else:
            k += 1
    if not verts or len(verts) % 3:
        raise SceneError(f"cannot read STL file {path!r}")
    return np.asarray(verts, dtype=np.float64).astype(np.float32).reshape(-1, 3, 3)


def read_obj(path):
    """Wavefront OBJ -> (n,3,3) float32: `v` and `f` records only, a/b/c index forms, negative indices, polygons
    fan-triangulated from their first vertex (mirror of cutrace_b200/host/scene_loader.cpp read_obj)."""
    try:
        with open(path, "rb") as f:
            lines = f.read().split(b"\n")      # bytes: lines end at LF only and tokens at ASCII white space, like std::getline / operator>>
    except OSError as e:
        raise SceneError(f"cannot open mesh file {path!r}") from e
    v, tris = [], []
    for line in lines:
        p = line.split()
        if not p:
            continue
        if p[0] == b"v":
            try:
                v.append([_mesh_number(p[1]), _mesh_number(p[2]), _mesh_number(p[3])])
            except (IndexError, ValueError) as e:
                raise SceneError(f"bad vertex in OBJ {path!r}") from e
        elif p[0] == b"f":
            idx = []
            for tok in p[1:]:
                head = tok.split(b"/")[0]
                if not _INDEX.match(head):
                    raise SceneError(f"bad face index in OBJ {path!r}")
                i = int(head)
                if i < 0:
                    i = len(v) + i + 1
                if not 1 <= i <= len(v):
                    raise SceneError(f"face index out of range in OBJ {path!r}")
                idx.append(i - 1)
            for k in range(1, len(idx) - 1):
                tris.append([v[idx[0]], v[idx[k]], v[idx[k + 1]]])
    if not tris:
        raise SceneError(f"no faces in OBJ file {path!r}")
    return np.asarray(tris, dtype=np.float64).astype(np.float32)


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply(path):
    """Stanford PLY (ascii / binary_little_endian / binary_big_endian) -> (n,3,3) float32: x, y, z of element `vertex`, the list
    property vertex_indices / vertex_index of element `face`, polygons fan-triangulated in face order; every other property and
    element is parsed and skipped (mirror of cutrace_b200/host/scene_loader.cpp read_ply)."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise SceneError(f"cannot open mesh file {path!r}") from e
    at = 0

    def next_line():
        nonlocal at
        if at >= len(data):
            return None
        e = data.find(b"\n", at)
        if e < 0:
            e = len(data)
        line = data[at:e]
        at = e + 1
        return line[:-1] if line.endswith(b"\r") else line

    if next_line() != b"ply":
        raise SceneError(f"not a PLY file: {path!r}")
    fmt, elems, ended = None, [], False
    while (line := next_line()) is not None:
        p = line.split()
        if not p or p[0] in (b"comment", b"obj_info"):
            continue
        if p[0] == b"end_header":
            ended = True
            break
        if p[0] == b"format":
            fmt = {b"ascii": 0, b"binary_little_endian": 1, b"binary_big_endian": 2}.get(p[1] if len(p) > 1 else b"")
        elif p[0] == b"element":
            if len(p) < 3 or not _INDEX.match(p[2]) or int(p[2]) < 0:
                raise SceneError(f"bad element line in PLY {path!r}")
            elems.append((p[1].decode("latin-1"), int(p[2]), []))
        elif p[0] == b"property":
            if not elems:
                raise SceneError(f"property before any element in PLY {path!r}")
            t = [x.decode("latin-1") for x in p[1:]]
            if t and t[0] == "list":
                if len(t) < 4 or t[1] not in _PLY_TYPES or _PLY_TYPES[t[1]][0] == "f" or t[2] not in _PLY_TYPES:
                    raise SceneError(f"bad list count type in PLY {path!r}" if len(t) >= 4 and t[2] in _PLY_TYPES else f"bad property line in PLY {path!r}")
                elems[-1][2].append((True, _PLY_TYPES[t[2]], _PLY_TYPES[t[1]], t[3]))
            else:
                if len(t) < 2 or t[0] not in _PLY_TYPES:
                    raise SceneError(f"bad property line in PLY {path!r}")
                elems[-1][2].append((False, _PLY_TYPES[t[0]], None, t[1]))
    if not ended or fmt is None:
        raise SceneError(f"bad PLY header in {path!r}")
    toks = data[at:].split() if fmt == 0 else None
    tk = 0

    def scalar(ty):
        nonlocal at, tk
        if fmt == 0:
            if tk >= len(toks):
                raise SceneError(f"truncated or malformed PLY body in {path!r}")
            try:
                x = _mesh_number(toks[tk])
            except ValueError as e:
                raise SceneError(f"truncated or malformed PLY body in {path!r}") from e
            tk += 1
            return x
        n = int(ty[1])
        if at + n > len(data):
            raise SceneError(f"truncated or malformed PLY body in {path!r}")
        x = np.frombuffer(data, dtype=(">" if fmt == 2 else "<") + ty, count=1, offset=at)[0]
        at += n
        return float(x)

    v, tris, saw_vertex = [], [], False
    for name, count, props in elems:
        is_vertex, is_face = name == "vertex", name == "face"
        names = [pr[3] for pr in props]
        if is_vertex and not all(any(n == c and not pr[0] for n, pr in zip(names, props)) for c in "xyz"):
            raise SceneError(f"PLY vertex element without x / y / z in {path!r}")
        il = next((k for k, pr in enumerate(props) if pr[0] and pr[3] in ("vertex_indices", "vertex_index")), -1) if is_face else -1
        if is_face and il < 0:
            raise SceneError(f"PLY face element without vertex_indices in {path!r}")
        if is_face and not saw_vertex:
            raise SceneError(f"PLY face element before the vertex element in {path!r}")
        saw_vertex = saw_vertex or is_vertex
        ix, iy, iz = (next(k for k, pr in enumerate(props) if pr[3] == c and not pr[0]) for c in "xyz") if is_vertex else (-1, -1, -1)
        if not props:
            continue      # rows without properties hold nothing (and a huge count must not spin here)
        for _ in range(count):
            xyz, idx = [0.0, 0.0, 0.0], []
            for k, (is_list, ty, cty, _n) in enumerate(props):
                if not is_list:
                    x = scalar(ty)
                    if k == ix:
                        xyz[0] = x
                    elif k == iy:
                        xyz[1] = x
                    elif k == iz:
                        xyz[2] = x
                    continue
                cnt = scalar(cty)
                if cnt < 0 or cnt > 1e6 or cnt != int(cnt):
                    raise SceneError(f"bad list length in PLY {path!r}")
                for _j in range(int(cnt)):
                    x = scalar(ty)
                    if k == il:
                        if x != int(x) or x < 0 or int(x) >= len(v):
                            raise SceneError(f"face index out of range in PLY {path!r}")
                        idx.append(int(x))
            if is_vertex:
                v.append(xyz)
            if is_face:
                for k in range(1, len(idx) - 1):
                    tris.append([v[idx[0]], v[idx[k]], v[idx[k + 1]]])
    if not tris:
        raise SceneError(f"no faces in PLY file {path!r}")
    return np.asarray(tris, dtype=np.float64).astype(np.float32)


def read_off(path):
    """Object File Format: `OFF`, `nv nf ne` (possibly on the OFF line), nv vertex lines, nf face lines `n i0 .. [colour]`, `#` comments;
    polygons fan-triangulated in face order (mirror of cutrace_b200/host/scene_loader.cpp read_off)."""
    try:
        with open(path, "rb") as f:
            raw = f.read().split(b"\n")
    except OSError as e:
        raise SceneError(f"cannot open mesh file {path!r}") from e
    lines = [t for t in (ln.split(b"#")[0].split() for ln in raw) if t]
    if not lines or lines[0][0] != b"OFF":
        raise SceneError(f"not an OFF file: {path!r}")
    head, row = lines[0][1:], 1
    if not head:
        if len(lines) < 2:
            raise SceneError(f"truncated OFF file {path!r}")
        head, row = lines[1], 2
    try:
        nv, nf = _mesh_number(head[0]), _mesh_number(head[1])
        if nv < 0 or nf < 0 or nv != int(nv) or nf != int(nf):
            raise ValueError(head)
    except (IndexError, ValueError) as e:
        raise SceneError(f"bad counts in OFF file {path!r}") from e
    nv, nf = int(nv), int(nf)
    if len(lines) < row + nv + nf:
        raise SceneError(f"truncated OFF file {path!r}")
    v, tris = [], []
    for ln in lines[row:row + nv]:
        try:
            v.append([_mesh_number(ln[0]), _mesh_number(ln[1]), _mesh_number(ln[2])])
        except (IndexError, ValueError) as e:
            raise SceneError(f"bad vertex in OFF {path!r}") from e
    for ln in lines[row + nv:row + nv + nf]:
        try:
            n = _mesh_number(ln[0])
            if n < 0 or n != int(n) or len(ln) < 1 + int(n):
                raise ValueError(ln)
        except ValueError as e:
            raise SceneError(f"bad face in OFF {path!r}") from e
        idx = []
        for tok in ln[1:1 + int(n)]:
            try:
                x = _mesh_number(tok)
            except ValueError:
                x = -1.0
            if x != int(x) or x < 0 or x >= nv:
                raise SceneError(f"face index out of range in OFF {path!r}")
            idx.append(int(x))
        for k in range(1, len(idx) - 1):
            tris.append([v[idx[0]], v[idx[k]], v[idx[k + 1]]])
    if not tris:
        raise SceneError(f"no faces in OFF file {path!r}")
    return np.asarray(tris, dtype=np.float64).astype(np.float32)


def read_mesh(path):
    """Mesh import by file extension: .obj, .ply, .off, everything else as STL (binary or ASCII)."""
    low = path.lower()
    if low.endswith(".obj"):
        return read_obj(path)
    if low.endswith(".ply"):
        return read_ply(path)
    if low.endswith(".off"):
        return read_off(path)
    return read_stl(path)


class cutrace_scene_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("cam_pos", C.c_float * 3),
        ("cam_up", C.c_float * 3),
        ("cam_forward", C.c_float * 3),
        ("cam_right", C.c_float * 3),
        ("ambient", C.c_float),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("n_triangles", C.c_uint64),
        ("tri_p1", C.c_void_p),
        ("tri_p2", C.c_void_p),
        ("tri_p3", C.c_void_p),
        ("tri_object", C.c_void_p),
        ("n_spheres", C.c_uint64),
        ("sph_center", C.c_void_p),
        ("sph_radius", C.c_void_p),
        ("sph_object", C.c_void_p),
        ("n_planes", C.c_uint64),
        ("pl_point", C.c_void_p),
        ("pl_normal", C.c_void_p),
        ("pl_object", C.c_void_p),
        ("n_objects", C.c_uint32),
        ("obj_material", C.c_void_p),
        ("obj_kind", C.c_void_p),
        ("n_materials", C.c_uint32),
        ("mat_color", C.c_void_p),
        ("mat_specular", C.c_void_p),
        ("mat_reflect", C.c_void_p),
        ("mat_phong", C.c_void_p),
        ("mat_transparency", C.c_void_p),
        ("n_lights", C.c_uint32),
        ("light_kind", C.c_void_p),
        ("light_vec", C.c_void_p),
        ("light_color", C.c_void_p),
    ]


_ARRAY_FIELDS = (
    ("tri_p1", np.float32, 3), ("tri_p2", np.float32, 3), ("tri_p3", np.float32, 3), ("tri_object", np.uint32, 0),
    ("sph_center", np.float32, 3), ("sph_radius", np.float32, 0), ("sph_object", np.uint32, 0),
    ("pl_point", np.float32, 3), ("pl_normal", np.float32, 3), ("pl_object", np.uint32, 0),
    ("obj_material", np.uint32, 0), ("obj_kind", np.uint32, 0),
    ("mat_color", np.float32, 3), ("mat_specular", np.float32, 0), ("mat_reflect", np.float32, 0),
    ("mat_phong", np.float32, 0), ("mat_transparency", np.float32, 0),
    ("light_kind", np.uint32, 0), ("light_vec", np.float32, 3), ("light_color", np.float32, 3),
)


@dataclass
class FlatScene:
    cam_pos: np.ndarray
    cam_up: np.ndarray
    cam_forward: np.ndarray
    cam_right: np.ndarray
    ambient: float
    width: int
    height: int
    tri_p1: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    tri_p2: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    tri_p3: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    tri_object: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    sph_center: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    sph_radius: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    sph_object: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    pl_point: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    pl_normal: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    pl_object: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    obj_material: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    obj_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))  # host-only (scene dump)
    mat_color: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    mat_specular: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    mat_reflect: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    mat_phong: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    mat_transparency: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    light_kind: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    light_vec: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    light_color: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))

    def __post_init__(self):
        for name in ("cam_pos", "cam_up", "cam_forward", "cam_right"):
            setattr(self, name, _f32(getattr(self, name), (3,)))
        for name, dt, k in _ARRAY_FIELDS:
            a = np.ascontiguousarray(np.asarray(getattr(self, name), dtype=dt))
            a = a.reshape(-1, 3) if k else a.reshape(-1)
            setattr(self, name, a)
        self.ambient = float(np.float32(self.ambient))
        self.width, self.height = int(self.width), int(self.height)

    # ---- sizes -----------------------------------------------------------------------------
    @property
    def n_triangles(self):
        return len(self.tri_object)

    @property
    def n_spheres(self):
        return len(self.sph_object)

    @property
    def n_planes(self):
        return len(self.pl_object)

    @property
    def n_objects(self):
        return len(self.obj_material)

    @property
    def n_lights(self):
        return len(self.light_kind)

    @property
    def n_primitives(self):
        return self.n_triangles + self.n_spheres + self.n_planes

    def with_resolution(self, width, height):
        import copy

        s = copy.copy(self)
        s.width, s.height = int(width), int(height)
        return s

    # ---- C view ----------------------------------------------------------------------------
    def as_desc(self):
        """Returns a ``cutrace_scene_desc`` that borrows this scene's arrays (keep self alive)."""
        d = cutrace_scene_desc()
        d.abi_version = ABI_VERSION
        for name in ("cam_pos", "cam_up", "cam_forward", "cam_right"):
            getattr(d, name)[:] = [float(x) for x in getattr(self, name)]
        d.ambient = self.ambient
        d.width, d.height = self.width, self.height
        d.n_triangles, d.n_spheres, d.n_planes = self.n_triangles, self.n_spheres, self.n_planes
        d.n_objects, d.n_materials, d.n_lights = self.n_objects, len(self.mat_specular), self.n_lights
        for name, _, _ in _ARRAY_FIELDS:
            a = getattr(self, name)
            setattr(d, name, a.ctypes.data if a.size else None)
        return d

    # ---- npz io (tests/golden fixtures) ------------------------------------------------------
    def to_npz_dict(self):
        out = {
            "cam_pos": self.cam_pos, "cam_up": self.cam_up, "cam_forward": self.cam_forward,
            "cam_right": self.cam_right, "ambient": np.float32(self.ambient),
            "width": np.uint32(self.width), "height": np.uint32(self.height),
        }
        for name, _, _ in _ARRAY_FIELDS:
            out[name] = getattr(self, name)
        return out

    def save(self, path):
        np.savez_compressed(path, **self.to_npz_dict())

    @classmethod
    def load(cls, path):
        z = np.load(path)
        kw = {k: z[k] for k in z.files}
        kw["ambient"] = float(kw["ambient"])
        kw["width"], kw["height"] = int(kw["width"]), int(kw["height"])
        return cls(**kw)

    # ---- unique-ray bookkeeping helpers --------------------------------------------------------
    def max_children(self, fudge_unused=None):
        """0/1/2: how many secondary rays one hit can spawn (inc/shading.hpp:130,141)."""
        r = self.mat_reflect.astype(np.float64) >= 1e-6
        t = self.mat_transparency.astype(np.float64) >= 1e-6
        if np.any(r & t):
            return 2
        return 1 if np.any(r | t) else 0


# ------------------------------------------------------------------------------------------------
# JSON front-end (mirror of default_schema.hpp's accepted input)
# ------------------------------------------------------------------------------------------------

def _num(o, key, default=None, what="value"):
    if key not in o:
        if default is None:
            raise SceneError(f"Cannot find key '{key}' in object.")
        return default
    v = o[key]
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        raise SceneError(f"Expected a value of type number for '{key}'.")
    return float(v)


def _vec(o, key, default=None):
    if key not in o:
        if default is None:
            raise SceneError(f"Cannot find key '{key}' in object.")
        return np.asarray(default, np.float32)
    v = o[key]
    if not isinstance(v, list) or len(v) != 3 or any(isinstance(x, bool) or not isinstance(x, (int, float)) for x in v):
        raise SceneError(f"Expected a 3-element array for '{key}'.")
    return np.asarray([float(x) for x in v], dtype=np.float64).astype(np.float32)


def _index(o, key):
    v = _num(o, key)
    return int(v)  # (size_t) cast of a double, inc/json_helpers.hpp:91


def scene_from_dict(doc, base_dir=".", accept_aliases=False):
    """Builds a FlatScene from a parsed scene JSON.

    ``accept_aliases`` additionally accepts the stale spellings of /root/reference/schema.md
    (``model``, ``position``, ``points``, untyped materials, missing camera keys) — a superset of
    the reference's behaviour, off by default.
    """
    for key in ("objects", "lights", "materials", "camera"):
        if key not in doc:
            raise SceneError(f"Cannot find key '{key}' in object.")
    cam = doc["camera"]
    if not isinstance(cam, dict):
        raise SceneError("Value is not a JSON object.")
    cam_defaults = dict(eye=[0, 0, 0], up=[0, 1, 0], look=[0, 0, 1], near_plane=0.1, far_plane=100.0,
                        width=1920, height=1080, ambient=0.1)  # inc/default_schema.hpp:835-842

    def cam_key(k, vec):
        if k not in cam and not accept_aliases:
            raise SceneError(f"Cannot find key '{k}' in object.")  # MK_MANDATORY, :888-897
        return _vec(cam, k, cam_defaults[k]) if vec else _num(cam, k, cam_defaults[k])

    eye, up, look = cam_key("eye", True), cam_key("up", True), cam_key("look", True)
    cam_key("near_plane", False), cam_key("far_plane", False)
    width, height = int(cam_key("width", False)), int(cam_key("height", False))
    ambient = np.float32(cam_key("ambient", False))
    forward, right, up2 = look_at(eye, up, look)

    mats = {"color": [], "specular": [], "reflect": [], "phong": [], "transparency": []}
    for m in doc["materials"]:
        if not isinstance(m, dict):
            raise SceneError("Value is not a JSON object.")
        ty = m.get("type", "solid" if accept_aliases else None)
        if ty != "solid":
            raise SceneError(f"Unknown material type {ty!r}.")
        mats["color"].append(_vec(m, "color"))
        mats["specular"].append(np.float32(_num(m, "specular", 0.3)))
        mats["reflect"].append(np.float32(_num(m, "reflect", 0.0)))
        mats["phong"].append(np.float32(_num(m, "phong", 32.0)))
        mats["transparency"].append(np.float32(_num(m, "transparency", 0.0)))
    n_mat = len(mats["specular"])

    lk, lv, lc = [], [], []
    for li in doc["lights"]:
        if not isinstance(li, dict):
            raise SceneError("Value is not a JSON object.")
        ty = li.get("type")
        if ty == "sun":
            lk.append(LIGHT_SUN)
            lv.append(_vec(li, "direction"))
        elif ty == "point":
            lk.append(LIGHT_POINT)
            key = "point" if ("point" in li or not accept_aliases) else "position"
            lv.append(_vec(li, key))
        else:
            raise SceneError(f"Unknown light type {ty!r}.")
        lc.append(_vec(li, "color", [1, 1, 1]))

    p1, p2, p3, tobj = [], [], [], []
    sc, sr, sobj = [], [], []
    pp, pn, pobj = [], [], []
    omat, okind = [], []
    for o in doc["objects"]:
        if not isinstance(o, dict):
            raise SceneError("Value is not a JSON object.")
        ty = o.get("type")
        oid = len(omat)
        if accept_aliases and ty == "model":
            ty = "mesh"
        if ty == "triangle":
            if accept_aliases and "points" in o and "p1" not in o:
                a, b, c = (np.asarray(x, np.float64).astype(np.float32) for x in o["points"])
            else:
                a, b, c = _vec(o, "p1"), _vec(o, "p2"), _vec(o, "p3")
            mat = _index(o, "material")
            p1.append(a[None]); p2.append(b[None]); p3.append(c[None]); tobj.append(np.full(1, oid, np.uint32))
            okind.append(OBJ_TRIANGLE)
        elif ty == "mesh":
            if "file" not in o or not isinstance(o["file"], str):
                raise SceneError("Cannot find key 'file' in object.")
            mat = _index(o, "material")
            path = o["file"]
            if not os.path.isabs(path):
                path = os.path.join(base_dir, path)
            v = read_mesh(path)
            p1.append(v[:, 0]); p2.append(v[:, 1]); p3.append(v[:, 2]); tobj.append(np.full(len(v), oid, np.uint32))
            okind.append(OBJ_MESH)
        elif ty == "plane":
            pp.append(_vec(o, "point")); pn.append(_vec(o, "normal")); pobj.append(oid)
            mat = _index(o, "material")
            okind.append(OBJ_PLANE)
        elif ty == "sphere":
            sc.append(_vec(o, "center")); sr.append(np.float32(_num(o, "radius"))); sobj.append(oid)
            mat = _index(o, "material")
            okind.append(OBJ_SPHERE)
        else:
            raise SceneError(f"Unknown object type {ty!r}.")
        if not (0 <= mat < n_mat):
            raise SceneError(f"material index {mat} out of range (have {n_mat} materials)")
        omat.append(mat)

    def cat(parts, k):
        return np.concatenate(parts).astype(np.float32) if parts else np.zeros((0, k), np.float32)

    return FlatScene(
        cam_pos=eye, cam_up=up2, cam_forward=forward, cam_right=right, ambient=float(ambient),
        width=width, height=height,
        tri_p1=cat(p1, 3), tri_p2=cat(p2, 3), tri_p3=cat(p3, 3),
        tri_object=np.concatenate(tobj) if tobj else np.zeros(0, np.uint32),
        sph_center=_f32(sc, (-1, 3)), sph_radius=_f32(sr), sph_object=_u32(sobj),
        pl_point=_f32(pp, (-1, 3)), pl_normal=_f32(pn, (-1, 3)), pl_object=_u32(pobj),
        obj_material=_u32(omat), obj_kind=_u32(okind),
        mat_color=_f32(mats["color"], (-1, 3)), mat_specular=_f32(mats["specular"]),
        mat_reflect=_f32(mats["reflect"]), mat_phong=_f32(mats["phong"]),
        mat_transparency=_f32(mats["transparency"]),
        light_kind=_u32(lk), light_vec=_f32(lv, (-1, 3)), light_color=_f32(lc, (-1, 3)),
    )


_JSON_STRICT_NUMBER = re.compile(r"-?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?\Z")
_STRTOD_NUMBER = re.compile(r"[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?\Z")


def _picojson_numbers(text):
    """The reference parses its scenes with picojson, whose number rule is laxer than strict JSON: a number starts at a digit or
    '-', is the longest run of [0-9+-.eE], and strtod must consume all of it — "041", "0." and "1.e5" are numbers, "1e" and "1-2"
    are errors (cutrace_b200/host/json.hpp restates the same rule).  This pass rewrites the runs strict JSON would reject into
    their value, outside string literals, so that json.loads accepts exactly what the C++ front-end accepts."""
    out, i, n, in_str = [], 0, len(text), False
    while i < n:
        c = text[i]
        if in_str:
            if c == "\\" and i + 1 < n:
                out.append(text[i:i + 2])
                i += 2
                continue
            in_str = c != '"'
        elif c == '"':
            in_str = True
        elif c == "-" or "0" <= c <= "9":
            q = i
            while q < n and text[q] in "0123456789+-.eE":
                q += 1
            run = text[i:q]
            if not _JSON_STRICT_NUMBER.match(run):
                if not _STRTOD_NUMBER.match(run):
                    raise SceneError("JSON parse error: bad number")
                v = float(run)
                run = repr(v) if v == v and abs(v) != float("inf") else ("-1e999" if v < 0 else "1e999")
            out.append(run)
            i = q
            continue
        out.append(c)
        i += 1
    return "".join(out)


def load_scene_json(path, base_dir=None, accept_aliases=False):
    """``default_schema::load_file`` equivalent. Mesh paths are relative to the CWD in the
    reference (schema.md:73-74); pass ``base_dir`` to resolve them elsewhere."""
    with open(path, "r") as f:
        try:
            doc = json.loads(_picojson_numbers(f.read()))
        except json.JSONDecodeError as e:
            raise SceneError(f"JSON parse error: {e}") from e
    if not isinstance(doc, dict):
        raise SceneError("Value is not a JSON object.")
    return scene_from_dict(doc, base_dir=base_dir if base_dir is not None else os.getcwd(),
                           accept_aliases=accept_aliases)
