"""Deterministic synthetic scenes (BASELINE.json config 5): a G x G grid of mesh instances with a
floor and two mirror planes.  The reference schema has no instancing or transforms
(/root/reference/inc/default_schema.hpp:603-606), so every instance is pre-transformed on the host
and becomes ONE mesh object (hit id = instance id, and the reference gets its per-mesh AABB culling).
G = 106 with the 1000-triangle bunny and the 800-triangle skull alternating gives 10,112,400 triangles.
"""
from __future__ import annotations

import numpy as np

from .scene import LIGHT_POINT, OBJ_MESH, OBJ_PLANE, FlatScene, look_at


def _normalise_mesh(v):
    """centre on the origin, scale the largest extent to 1"""
    v = np.asarray(v, np.float64)
    lo, hi = v.reshape(-1, 3).min(0), v.reshape(-1, 3).max(0)
    return (v - 0.5 * (lo + hi)) / (hi - lo).max()


def grid_scene(meshes, grid=106, width=7680, height=4320, seed=0, n_lights=3):
    """meshes: list of (n,3,3) vertex arrays used round-robin. Returns a FlatScene."""
    rng = np.random.default_rng(seed)
    base = [_normalise_mesh(m) for m in meshes]
    G = int(grid)
    p = [[], [], []]
    tobj, omat, okind = [], [], []
    for iz in range(G):
        for ix in range(G):
            k = iz * G + ix
            m = base[k % len(base)]
            scale = 0.55 + 0.3 * rng.random()
            ang = 2 * np.pi * rng.random()
            c, s = np.cos(ang), np.sin(ang)
            rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
            centre = np.array([ix - 0.5 * (G - 1), 0.5 * scale * (m[..., 1].max() - m[..., 1].min()), iz - 0.5 * (G - 1)])
            centre[1] -= 0.0
            v = (m * scale) @ rot.T + centre
            v = v.astype(np.float32)
            for j in range(3):
                p[j].append(v[:, j])
            tobj.append(np.full(len(v), k, np.uint32))
            omat.append(1 + (k % 6) if k % 7 else 7)   # every 7th instance is glossy-reflective
            okind.append(OBJ_MESH)
    n_inst = G * G
    half = 0.5 * G + 1.0
    # a closed hall like the reference's own box scenes (mirror.json / bunny.json): floor, ceiling, two mirror walls
    # behind the grid and two matte walls behind the camera
    top = 0.9 * G
    pl_point = np.array([[0, 0, 0], [0, 0, half], [-half, 0, 0], [half, 0, 0], [0, 0, -half], [0, top, 0]], np.float32)
    pl_normal = np.array([[0, 1, 0], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, -1, 0]], np.float32)
    pl_object = np.arange(n_inst, n_inst + 6, dtype=np.uint32)
    omat += [0, 8, 8, 0, 0, 0]
    okind += [OBJ_PLANE] * 6
    mat_color = np.array([[0.8, 0.8, 0.8], [0.894, 0.102, 0.110], [0.216, 0.494, 0.722], [0.302, 0.686, 0.290],
                          [1.0, 0.498, 0.0], [0.596, 0.306, 0.639], [0.9, 0.9, 0.3], [0.7, 0.7, 0.75], [0.1, 0.1, 0.1]], np.float32)
    mat_specular = np.array([0.2, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3, 0.6, 0.1], np.float32)
    mat_reflect = np.array([0.1, 0, 0, 0, 0, 0, 0, 0.3, 0.9], np.float32)
    mat_phong = np.array([100, 200, 200, 200, 200, 200, 200, 500, 1000], np.float32)
    mat_transparency = np.zeros(9, np.float32)
    lights = np.array([[0.0, 0.6 * G, -0.2 * G], [-0.3 * G, 0.5 * G, 0.3 * G], [0.3 * G, 0.4 * G, -0.4 * G], [0, 0.8 * G, 0]],
                      np.float32)[:n_lights]
    eye = np.array([0.30 * G, 0.34 * G, -0.46 * G], np.float32)   # inside the hall, looking down across the grid
    fwd, right, up = look_at(eye, [0, 1, 0], [0, 0, 0])
    return FlatScene(
        cam_pos=eye, cam_up=up, cam_forward=fwd, cam_right=right, ambient=0.02, width=width, height=height,
        tri_p1=np.concatenate(p[0]), tri_p2=np.concatenate(p[1]), tri_p3=np.concatenate(p[2]),
        tri_object=np.concatenate(tobj),
        pl_point=pl_point, pl_normal=pl_normal, pl_object=pl_object,
        obj_material=np.asarray(omat, np.uint32), obj_kind=np.asarray(okind, np.uint32),
        mat_color=mat_color, mat_specular=mat_specular, mat_reflect=mat_reflect, mat_phong=mat_phong,
        mat_transparency=mat_transparency,
        light_kind=np.full(len(lights), LIGHT_POINT, np.uint32), light_vec=lights,
        light_color=np.full((len(lights), 3), 0.45, np.float32),
    )


def meshes_from_scenes(*scenes):
    """Pulls the mesh objects out of FlatScenes (e.g. tests/golden/scenes/bunny.npz, mirror.npz) as
    (n,3,3) arrays, largest first."""
    out = []
    for s in scenes:
        for o in np.unique(s.tri_object):
            sel = s.tri_object == o
            if sel.sum() >= 100:
                out.append(np.stack([s.tri_p1[sel], s.tri_p2[sel], s.tri_p3[sel]], axis=1))
    out.sort(key=lambda a: -len(a))
    return out


def random_soup(n_tri=200, n_sph=5, n_planes=2, n_lights=2, width=96, height=64, seed=0, translucent=True,
                duplicates=False):
    """Random test scene: triangle soup meshes, spheres, planes, mixed materials (used by the parity tests)."""
    rng = np.random.default_rng(seed)
    n_mesh = max(1, n_tri // 40) if n_tri else 0
    c = rng.uniform(-1.5, 1.5, (n_tri, 1, 3))
    v = (c + rng.normal(0, 0.25, (n_tri, 3, 3))).astype(np.float32)
    tobj = np.sort(rng.integers(0, max(n_mesh, 1), n_tri)).astype(np.uint32) if n_tri else np.zeros(0, np.uint32)
    if duplicates and n_tri >= 4:   # coincident triangles in different objects + inside one object: tie-break rule
        v[1] = v[0]
        v[-1] = v[0]
        v[n_tri // 2] = v[n_tri // 2 - 1]
    n_obj = n_mesh + n_sph + n_planes
    n_mat = 6
    mat_reflect = np.array([0, 0.3, 0.05, 0, 0.5, 0.999], np.float32)
    mat_transparency = np.array([0, 0, 0.6, 0.3, 0, 0], np.float32) if translucent else np.zeros(6, np.float32)
    fwd, right, up = look_at([0, 0.3, -5], [0, 1, 0], [0, 0, 0])
    kinds = [1] * n_mesh + [3] * n_sph + [2] * n_planes
    pl_n = rng.normal(0, 1, (n_planes, 3)).astype(np.float32)
    pl_n[:, 1] = np.abs(pl_n[:, 1]) + 2.0   # roughly upward facing, not normalised on purpose
    return FlatScene(
        cam_pos=[0, 0.3, -5], cam_up=up, cam_forward=fwd, cam_right=right, ambient=0.05, width=width, height=height,
        tri_p1=v[:, 0], tri_p2=v[:, 1], tri_p3=v[:, 2], tri_object=tobj,
        sph_center=rng.uniform(-1.5, 1.5, (n_sph, 3)).astype(np.float32), sph_radius=rng.uniform(0.2, 0.7, n_sph).astype(np.float32),
        sph_object=np.arange(n_mesh, n_mesh + n_sph, dtype=np.uint32),
        pl_point=np.stack([np.zeros(n_planes), -2.0 - np.arange(n_planes), np.zeros(n_planes)], 1).astype(np.float32),
        pl_normal=pl_n, pl_object=np.arange(n_mesh + n_sph, n_obj, dtype=np.uint32),
        obj_material=rng.integers(0, n_mat, n_obj).astype(np.uint32), obj_kind=np.asarray(kinds, np.uint32),
        mat_color=rng.uniform(0.1, 1.0, (n_mat, 3)).astype(np.float32), mat_specular=rng.uniform(0.1, 0.8, n_mat).astype(np.float32),
        mat_reflect=mat_reflect, mat_phong=np.array([32, 200, 1000, 20, 500, 100], np.float32), mat_transparency=mat_transparency,
        light_kind=np.array([0, 1, 1, 0][:n_lights], np.uint32),
        light_vec=np.array([[-1, -1, 1], [-3, 5, -4], [3, 4, -2], [0.5, -1, 0.2]], np.float32)[:n_lights],
        light_color=np.full((n_lights, 3), 0.6, np.float32),
    )


def write_scene_files(scene, out_dir, eye, up, look, name="scene", near_plane=0.1, far_plane=1000.0):
    """Writes ``scene`` as files a stock cutrace reads: ``<name>.json`` in the reference's schema
    (/root/reference/inc/default_schema.hpp:487-897: objects/lights/materials/camera, every camera key present) and
    one binary STL per mesh object (80-byte header, uint32 count, 50-byte records; the stored facet normal is zero —
    the reference recomputes it, default_schema.hpp:72).  float32 values are written with their shortest exact decimal,
    so loading the files back gives the same arrays bit for bit.  ``eye/up/look`` are the camera's JSON parameters (a
    FlatScene only keeps the vectors look_at derived from them).  Mesh paths in the JSON are relative to ``out_dir``:
    the reference resolves them against the CWD (schema.md:73-74).  Returns the JSON path."""
    import json
    import os
    import struct

    from .scene import LIGHT_SUN, OBJ_MESH, OBJ_PLANE, OBJ_SPHERE, OBJ_TRIANGLE

    def f(x):
        return float(np.float32(x))          # json writes repr(double) — exact, and it rounds back to the same float32

    def v3(a):
        return [f(a[0]), f(a[1]), f(a[2])]

    os.makedirs(out_dir, exist_ok=True)
    objects = []
    order = np.argsort(scene.tri_object, kind="stable")
    first = np.searchsorted(scene.tri_object[order], np.arange(scene.n_objects + 1))
    sph = {int(o): i for i, o in enumerate(scene.sph_object)}
    pl = {int(o): i for i, o in enumerate(scene.pl_object)}
    for o in range(scene.n_objects):
        kind, mat = int(scene.obj_kind[o]), int(scene.obj_material[o])
        tri = order[first[o]:first[o + 1]]
        if kind == OBJ_TRIANGLE:
            t = int(tri[0])
            objects.append({"type": "triangle", "p1": v3(scene.tri_p1[t]), "p2": v3(scene.tri_p2[t]), "p3": v3(scene.tri_p3[t]),
                            "material": mat})
        elif kind == OBJ_MESH:
            rel = f"{name}_mesh{o:04d}.stl"
            rec = np.zeros((len(tri), 50), np.uint8)
            verts = np.stack([scene.tri_p1[tri], scene.tri_p2[tri], scene.tri_p3[tri]], axis=1).astype("<f4")
            rec[:, 12:48] = verts.reshape(len(tri), 9).view(np.uint8).reshape(len(tri), 36)
            with open(os.path.join(out_dir, rel), "wb") as fh:
                fh.write(b"cutrace-b200 synthetic instance".ljust(80, b"\0"))
                fh.write(struct.pack("<I", len(tri)))
                fh.write(rec.tobytes())
            objects.append({"type": "mesh", "file": rel, "material": mat})
        elif kind == OBJ_PLANE:
            i = pl[o]
            objects.append({"type": "plane", "point": v3(scene.pl_point[i]), "normal": v3(scene.pl_normal[i]), "material": mat})
        elif kind == OBJ_SPHERE:
            i = sph[o]
            objects.append({"type": "sphere", "center": v3(scene.sph_center[i]), "radius": f(scene.sph_radius[i]), "material": mat})
        else:
            raise ValueError(f"object {o} has unknown kind {kind}")
    lights = [{"type": "sun", "direction": v3(v), "color": v3(c)} if int(k) == LIGHT_SUN else
              {"type": "point", "point": v3(v), "color": v3(c)}
              for k, v, c in zip(scene.light_kind, scene.light_vec, scene.light_color)]
    materials = [{"type": "solid", "color": v3(scene.mat_color[i]), "specular": f(scene.mat_specular[i]),
                  "reflect": f(scene.mat_reflect[i]), "phong": f(scene.mat_phong[i]),
                  "transparency": f(scene.mat_transparency[i])} for i in range(len(scene.mat_specular))]
    camera = {"eye": v3(eye), "up": v3(up), "look": v3(look), "near_plane": near_plane, "far_plane": far_plane,
              "width": int(scene.width), "height": int(scene.height), "ambient": f(scene.ambient)}
    path = os.path.join(out_dir, f"{name}.json")
    with open(path, "w") as fh:
        json.dump({"objects": objects, "lights": lights, "materials": materials, "camera": camera}, fh, indent=1)
    return path


def grid_camera(grid):
    """(eye, up, look) that grid_scene() aims its camera with — the JSON camera of write_scene_files()."""
    G = int(grid)
    return np.array([0.30 * G, 0.34 * G, -0.46 * G], np.float32), [0, 1, 0], [0, 0, 0]
