"""Collider ingest and frustum culling: the host-side mirror of the reference's callers either side of the step
(SURVEY 8f ranks 2 and 4).

* `load_obj`            a neutral mesh format in place of the FBX SDK: `v` and `f` records of a Wavefront OBJ
                        (faces fan-triangulated, negative indices allowed), optionally grouped by `o` / `g`
* `mesh_triangles` ...  FbxLoader::_buildGeometryData (model/src/FbxLoader.cpp:1109-1216) through the C ABI
* `frustum_planes` ...  GeometryMath::getFrustumPlanes / OSP::getVisibleFrustumCulling through the C ABI
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ptr

IDENTITY_TRS = np.array([0, 0, 0, 0, 0, 0, 1, 1, 1], np.float32)


def load_obj(path_or_text: str, split_objects: bool = False):
    """Returns (xyz float32 [nv,3], idx uint32 [3*ntri]) or, with split_objects, {name: (xyz, idx)} with
    per-object index arrays into the shared vertex array (one collider mesh node per `o` / `g` record)."""
    text = path_or_text if "\n" in path_or_text else open(path_or_text).read()
    verts, groups, cur = [], {}, "default"
    for line in text.splitlines():
        t = line.split("#", 1)[0].split()
        if not t:
            continue
        if t[0] == "v":
            verts.append([float(t[1]), float(t[2]), float(t[3])])
        elif t[0] in ("o", "g"):
            cur = t[1] if len(t) > 1 else "default"
        elif t[0] == "f":
            ids = []
            for tok in t[1:]:
                i = int(tok.split("/")[0])
                ids.append(i - 1 if i > 0 else len(verts) + i)
            for k in range(1, len(ids) - 1):
                groups.setdefault(cur, []).extend([ids[0], ids[k], ids[k + 1]])
    xyz = np.asarray(verts, np.float32).reshape(-1, 3)
    if split_objects:
        return {k: (xyz, np.asarray(v, np.uint32)) for k, v in groups.items()}
    idx = np.asarray([i for v in groups.values() for i in v], np.uint32)
    return xyz, idx


def _ck(rc, what):
    if rc:
        raise RuntimeError(f"{what} failed (code {rc})")


def trs_matrix(trs9=IDENTITY_TRS) -> np.ndarray:
    out = np.zeros(16, np.float32)
    _ck(_lib.load().rb_ingest_trs_matrix(ptr(np.ascontiguousarray(trs9, np.float32)), ptr(out)), "rb_ingest_trs_matrix")
    return out.reshape(4, 4)


def mesh_triangles(xyz, idx, trs9=IDENTITY_TRS) -> np.ndarray:
    """Static model -> world-space triangles [ntri, 9] (feed to World.add_static_geometry)."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3); idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    out = np.zeros((idx.shape[0] // 3, 9), np.float32)
    _ck(_lib.load().rb_ingest_mesh_triangles(xyz.shape[0], ptr(xyz), idx.shape[0], ptr(idx), ptr(np.ascontiguousarray(trs9, np.float32)), ptr(out)), "rb_ingest_mesh_triangles")
    return out


def mesh_sphere(xyz, idx, trs9=IDENTITY_TRS) -> np.ndarray:
    """Collider mesh of an animated model -> its one object-space sphere (cx, cy, cz, r)."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3); idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    out = np.zeros(4, np.float32)
    _ck(_lib.load().rb_ingest_mesh_sphere(xyz.shape[0], ptr(xyz), idx.shape[0], ptr(idx), ptr(np.ascontiguousarray(trs9, np.float32)), ptr(out)), "rb_ingest_mesh_sphere")
    return out


def frustum_planes(inv_view_proj) -> np.ndarray:
    out = np.zeros(24, np.float32)
    _ck(_lib.load().rb_frustum_planes(ptr(np.ascontiguousarray(inv_view_proj, np.float32).reshape(-1)), ptr(out)), "rb_frustum_planes")
    return out.reshape(6, 4)


def cull_cubes(world, planes, cubes6) -> np.ndarray:
    """Indices of the cubes (centre, length, height, width) that meet the frustum, front to back."""
    c = np.ascontiguousarray(cubes6, np.float32).reshape(-1, 6)
    out = np.zeros(max(c.shape[0], 1), np.uint32); n = C.c_uint32()
    world._ck(world.L.rb_frustum_cull_cubes(world.h, ptr(np.ascontiguousarray(planes, np.float32).reshape(-1)), c.shape[0], ptr(c), ptr(out), out.shape[0], C.byref(n)))
    return out[: n.value]


def cull_static(world, inv_view_proj) -> np.ndarray:
    """Ids of the non-empty static cells that meet the frustum, front to back."""
    st = world.stats(); ncell = st["tri_cols_x"] * st["tri_cols_z"]
    out = np.zeros(max(ncell, 1), np.uint32); n = C.c_uint32()
    world._ck(world.L.rb_frustum_cull_static(world.h, ptr(np.ascontiguousarray(inv_view_proj, np.float32).reshape(-1)), ptr(out), out.shape[0], C.byref(n)))
    return out[: n.value]


def static_cells(world):
    """(boxes [ncell, 6] = min xyz, max xyz; ntri [ncell]) of the world's static triangle grid."""
    st = world.stats(); ncell = st["tri_cols_x"] * st["tri_cols_z"]
    boxes = np.zeros((max(ncell, 1), 6), np.float32); ntri = np.zeros(max(ncell, 1), np.uint32)
    world._ck(world.L.rb_get_static_cells(world.h, ptr(boxes), ptr(ntri)))
    return boxes[:ncell], ntri[:ncell]
