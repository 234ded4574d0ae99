"""ctypes binding for oracle/_ref/libref_physics.so (the reference's own physics/ + math/ sources,
compiled unmodified by oracle/Makefile). TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_physics.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_world_create.restype = C.c_void_p
        L.ref_world_destroy.argtypes = [C.c_void_p]
        L.ref_add_sphere_entity.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_float, C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
        L.ref_add_sphere_entity.restype = C.c_int
        L.ref_add_triangle_entity.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_add_triangle_entity.restype = C.c_int
        L.ref_build.argtypes = [C.c_void_p]
        L.ref_build.restype = C.c_double
        for n in ("ref_integrate", "ref_collide", "ref_step"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.ref_detect_frozen.argtypes = [C.c_void_p]
        L.ref_detect_frozen.restype = C.c_long
        L.ref_hit_count.argtypes = [C.c_void_p]
        L.ref_hit_count.restype = C.c_long
        L.ref_get_hits.argtypes = [C.c_void_p, _i32p, _f32p]
        L.ref_visualize.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_visualize.restype = C.c_long
        L.ref_get_counters.argtypes = [C.c_void_p, _i64p]
        L.ref_reset_counters.argtypes = [C.c_void_p]
        L.ref_entity_count.argtypes = [C.c_void_p]
        L.ref_get_state.argtypes = [C.c_void_p, _f32p, _f32p, _i32p, _f32p, _f32p]
        L.ref_get_model_matrices.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_set_force.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_set_torque.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_get_angular.argtypes = [C.c_void_p, _f32p, _f32p]
        L.ref_get_world_spheres.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_leaf_count.argtypes = [C.c_void_p]
        L.ref_leaf_count.restype = C.c_long
        L.ref_get_leaves.argtypes = [C.c_void_p, _f32p, _i32p, _i32p]
        L.ref_pred_sphere_triangle.argtypes = [_f32p, C.c_float, _f32p]
        L.ref_pred_sphere_sphere.argtypes = [_f32p, _f32p]
        L.ref_pred_sphere_cube.argtypes = [_f32p, _f32p]
        L.ref_pred_triangle_cube.argtypes = [_f32p, _f32p]
        L.ref_closest_point.argtypes = [_f32p, _f32p, _f32p]
        L.ref_state_update.argtypes = [_f32p, C.c_int, C.c_int, C.c_int]
        L.ref_bench.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_bench.restype = C.c_double
        L.ref_frustum_planes.argtypes = [_f32p, _f32p]
        L.ref_frustum_aabb.argtypes = [C.c_long, _f32p, _f32p, _f32p, _i32p]
        L.ref_inverse_view_projection.argtypes = [_f32p, _f32p, _f32p]
        L.ref_visible_frustum_culling.argtypes = [C.c_void_p, _f32p, _f32p, _i32p, C.c_long]
        L.ref_visible_frustum_culling.restype = C.c_long
        L.ref_trs_matrix.argtypes = [_f32p, _f32p]
        L.ref_ingest_triangles.argtypes = [C.c_long, _f32p, _i32p, _f32p, _f32p]
        L.ref_ingest_sphere.argtypes = [C.c_long, _f32p, _i32p, _f32p, _f32p]
        _lib = L
    return _lib


class RefWorld:
    """The compiled reference behind the same scene-level calls the product exposes.

    Entities are created in the canonical insertion order (static geometries, then bodies) so entity
    index = geom id for i < n_geoms and body id + n_geoms otherwise."""

    def __init__(self, scene=None):
        self.L = lib()
        self.h = self.L.ref_world_create()
        self.n_geoms = 0
        self.n_bodies = 0
        self.build_seconds = None
        if scene is not None:
            self.load(scene)
            self.build()

    def load(self, scene):
        for t in scene.tris:
            t = np.ascontiguousarray(t, dtype=np.float32)
            self.L.ref_add_triangle_entity(self.h, t.shape[0], t)
            self.n_geoms += 1
        ss = scene.sphere_start
        for b in range(scene.n_bodies):
            sp = np.ascontiguousarray(scene.obj_spheres[ss[b]:ss[b + 1]], dtype=np.float32)
            self.L.ref_add_sphere_entity(self.h, sp.shape[0], sp, float(scene.body_scale[b]), None,
                                         np.ascontiguousarray(scene.body_pos[b]), np.ascontiguousarray(scene.body_vel[b]),
                                         int(scene.body_active[b]), int(scene.body_gravity[b]))
            self.n_bodies += 1

    def build(self):
        self.build_seconds = self.L.ref_build(self.h)
        return self.build_seconds

    def step(self, ms=5):
        self.L.ref_step(self.h, ms)

    def integrate(self, ms=5):
        self.L.ref_integrate(self.h, ms)

    def collide(self, ms=5):
        self.L.ref_collide(self.h, ms)

    def bench(self, ms, nsteps) -> float:
        return self.L.ref_bench(self.h, ms, nsteps)

    def hits(self):
        """(ints[n,5] = kind,eA,sA,eB,pB ; floats[n,8] = depth,nx,ny,nz,px,py,pz,r) of the last collide/detect,
        in the order the reference produced them (duplicates across leaves kept)."""
        n = self.L.ref_hit_count(self.h)
        a = np.zeros((max(n, 1), 5), dtype=np.int32)
        f = np.zeros((max(n, 1), 8), dtype=np.float32)
        if n:
            self.L.ref_get_hits(self.h, a, f)
        return a[:n], f[:n]

    def detect_frozen(self):
        self.L.ref_detect_frozen(self.h)
        return self.hits()

    def contact_sets(self, hits=None):
        """De-duplicated sorted contact sets in body/geom ids:
        ss: rows (bodyA, sphereA, bodyB, sphereB) with (bodyA,sphereA) < (bodyB,sphereB);
        st: rows (body, sphere, geom, tri)."""
        a, _ = self.hits() if hits is None else hits
        g = self.n_geoms
        ss = a[a[:, 0] == 0][:, 1:].copy()
        st = a[a[:, 0] == 1][:, 1:].copy()
        if len(ss):
            ss[:, 0] -= g
            ss[:, 2] -= g
            swap = (ss[:, 0] > ss[:, 2]) | ((ss[:, 0] == ss[:, 2]) & (ss[:, 1] > ss[:, 3]))
            ss[swap] = ss[swap][:, [2, 3, 0, 1]]
            ss = np.unique(ss, axis=0)
        if len(st):
            st[:, 0] -= g
            st = np.unique(st, axis=0)
        return ss.reshape(-1, 4), st.reshape(-1, 4)

    def visualize(self):
        n = self.L.ref_visualize(self.h, None)
        out = np.zeros((max(n, 1), 3), dtype=np.int32)
        if n:
            self.L.ref_visualize(self.h, out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def counters(self):
        c = np.zeros(4, dtype=np.int64)
        self.L.ref_get_counters(self.h, c)
        return dict(tests_ss=int(c[0]), tests_st=int(c[1]), hits_ss=int(c[2]), hits_st=int(c[3]))

    def reset_counters(self):
        self.L.ref_reset_counters(self.h)

    def state(self):
        """Body-only state dict (static geometry entities stripped)."""
        n = self.L.ref_entity_count(self.h)
        pos = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32)
        fl = np.zeros(n, np.int32)
        mt = np.zeros((n, 3), np.float32)
        pt = np.zeros((n, 3), np.float32)
        self.L.ref_get_state(self.h, pos, vel, fl, mt, pt)
        g = self.n_geoms
        return dict(pos=pos[g:], vel=vel[g:], flags=fl[g:], model_t=mt[g:], prev_t=pt[g:])

    def model_matrices(self):
        n = self.L.ref_entity_count(self.h)
        cur = np.zeros((n, 16), np.float32)
        prev = np.zeros((n, 16), np.float32)
        self.L.ref_get_model_matrices(self.h, cur, prev)
        g = self.n_geoms
        return cur[g:].reshape(-1, 4, 4), prev[g:].reshape(-1, 4, 4)

    def set_state(self, body, pos=None, vel=None):
        p = None if pos is None else np.ascontiguousarray(pos, np.float32)
        v = None if vel is None else np.ascontiguousarray(vel, np.float32)
        self.L.ref_set_state(self.h, body + self.n_geoms,
                             None if p is None else p.ctypes.data_as(C.c_void_p),
                             None if v is None else v.ctypes.data_as(C.c_void_p))

    def set_flags(self, body, active=-1, gravity=-1, contact=-1):
        self.L.ref_set_flags(self.h, body + self.n_geoms, int(active), int(gravity), int(contact))

    def set_force(self, body, f):
        self.L.ref_set_force(self.h, body + self.n_geoms, np.ascontiguousarray(f, np.float32))

    def set_torque(self, body, t):
        self.L.ref_set_torque(self.h, body + self.n_geoms, np.ascontiguousarray(t, np.float32))

    def angular(self):
        n = self.L.ref_entity_count(self.h)
        ap = np.zeros((n, 3), np.float32)
        av = np.zeros((n, 3), np.float32)
        self.L.ref_get_angular(self.h, ap, av)
        return ap[self.n_geoms:], av[self.n_geoms:]

    def world_spheres(self, body, n):
        out = np.zeros((n, 4), np.float32)
        self.L.ref_get_world_spheres(self.h, body + self.n_geoms, out)
        return out

    def leaves(self):
        n = self.L.ref_leaf_count(self.h)
        cube = np.zeros((n, 4), np.float32)
        nt = np.zeros(n, np.int32)
        ns = np.zeros(n, np.int32)
        self.L.ref_get_leaves(self.h, cube, nt, ns)
        return cube, nt, ns


# ---- predicate micro entry points (vectorised loops over the scalar C calls) -------------------
def pred_sphere_triangle(c, r, t):
    L = lib()
    c = np.ascontiguousarray(c, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    r = np.broadcast_to(np.asarray(r, np.float32), (c.shape[0],))
    return np.array([L.ref_pred_sphere_triangle(c[i], float(r[i]), t[i]) for i in range(c.shape[0])], dtype=np.int32)


def pred_sphere_sphere(a, b):
    L = lib()
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 4)
    return np.array([L.ref_pred_sphere_sphere(a[i], b[i]) for i in range(a.shape[0])], dtype=np.int32)


def pred_sphere_cube(s, cube):
    L = lib()
    s = np.ascontiguousarray(s, np.float32).reshape(-1, 4)
    cube = np.ascontiguousarray(cube, np.float32).reshape(-1, 6)
    return np.array([L.ref_pred_sphere_cube(s[i], cube[i]) for i in range(s.shape[0])], dtype=np.int32)


def pred_triangle_cube(t, cube):
    L = lib()
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    cube = np.ascontiguousarray(cube, np.float32).reshape(-1, 6)
    return np.array([L.ref_pred_triangle_cube(t[i], cube[i]) for i in range(t.shape[0])], dtype=np.int32)


def closest_point(c, t):
    L = lib()
    c = np.ascontiguousarray(c, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    out = np.zeros((c.shape[0], 3), np.float32)
    for i in range(c.shape[0]):
        L.ref_closest_point(c[i], t[i], out[i])
    return out


def state_update(io18, flags, ms=5, nsteps=1):
    L = lib()
    io = np.ascontiguousarray(io18, np.float32).reshape(-1, 18).copy()
    flags = np.broadcast_to(np.asarray(flags, np.int32), (io.shape[0],))
    for i in range(io.shape[0]):
        L.ref_state_update(io[i], int(flags[i]), ms, nsteps)
    return io


# ---- frustum culling and collider ingest through the reference's own arithmetic ----
def frustum_planes(inv_vp):
    out = np.zeros(24, np.float32)
    lib().ref_frustum_planes(np.ascontiguousarray(inv_vp, np.float32).reshape(-1), out)
    return out.reshape(6, 4)


def frustum_aabb(planes, mins, maxs):
    mins = np.ascontiguousarray(mins, np.float32).reshape(-1, 3); maxs = np.ascontiguousarray(maxs, np.float32).reshape(-1, 3)
    out = np.zeros(mins.shape[0], np.int32)
    lib().ref_frustum_aabb(mins.shape[0], np.ascontiguousarray(planes, np.float32).reshape(-1), mins, maxs, out)
    return out


def inverse_view_projection(view, proj):
    out = np.zeros(16, np.float32)
    lib().ref_inverse_view_projection(np.ascontiguousarray(view, np.float32).reshape(-1), np.ascontiguousarray(proj, np.float32).reshape(-1), out)
    return out.reshape(4, 4)


def visible_frustum_culling(world, view, proj):
    """OSP::getVisibleFrustumCulling on a RefWorld's octree: 1-based visible leaf offsets, front to back."""
    n = world.L.ref_leaf_count(world.h)
    out = np.zeros(max(n, 1), np.int32)
    k = lib().ref_visible_frustum_culling(world.h, np.ascontiguousarray(view, np.float32).reshape(-1), np.ascontiguousarray(proj, np.float32).reshape(-1), out, n)
    return out[:k]


def trs_matrix(trs9):
    out = np.zeros(16, np.float32)
    lib().ref_trs_matrix(np.ascontiguousarray(trs9, np.float32), out)
    return out.reshape(4, 4)


def ingest_triangles(xyz3, idx, trs9):
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
    out = np.zeros((idx.shape[0] // 3, 9), np.float32)
    lib().ref_ingest_triangles(idx.shape[0], np.ascontiguousarray(xyz3, np.float32).reshape(-1), idx, np.ascontiguousarray(trs9, np.float32), out.reshape(-1))
    return out


def ingest_sphere(xyz3, idx, trs9):
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
    out = np.zeros(4, np.float32)
    lib().ref_ingest_sphere(idx.shape[0], np.ascontiguousarray(xyz3, np.float32).reshape(-1), idx, np.ascontiguousarray(trs9, np.float32), out)
    return out
