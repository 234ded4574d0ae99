"""ctypes binding for oracle/libreboot_oracle.so (plain-C restatement, oracle/reboot_oracle.c).
TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libreboot_oracle.so")
ORDER_REFERENCE = 0
ORDER_CANONICAL = 1

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no -march, no fast-math) -> libreboot_oracle.so"""
    src = os.path.join(_HERE, "reboot_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_add_geom.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_add_body.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_void_p, _f32p, _f32p, C.c_int, C.c_int]
        L.orc_build.argtypes = [C.c_void_p, C.c_int]
        L.orc_integrate.argtypes = [C.c_void_p, C.c_int]
        L.orc_collide.argtypes = [C.c_void_p, C.c_int]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_detect.argtypes = [C.c_void_p, C.c_int]
        L.orc_detect.restype = C.c_long
        L.orc_hit_count.argtypes = [C.c_void_p]
        L.orc_hit_count.restype = C.c_long
        L.orc_get_hits.argtypes = [C.c_void_p, _i32p, _f32p]
        L.orc_set_record.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_prune_slack.argtypes = [C.c_void_p, C.c_float]
        L.orc_get_counters.argtypes = [C.c_void_p, _i64p]
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_leaf_count.argtypes = [C.c_void_p]
        L.orc_get_state.argtypes = [C.c_void_p, _f32p, _f32p, _i32p, _f32p, _f32p]
        L.orc_get_model_matrices.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_get_angular.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_entity_set_position.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_set_force.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_set_torque.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_get_world_spheres.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_pred_sphere_triangle.argtypes = [C.c_long, _f32p, _f32p, _f32p, _i32p]
        L.orc_pred_sphere_sphere.argtypes = [C.c_long, _f32p, _f32p, _i32p]
        L.orc_pred_sphere_cube.argtypes = [C.c_long, _f32p, _f32p, _i32p]
        L.orc_pred_triangle_cube.argtypes = [C.c_long, _f32p, _f32p, _i32p]
        L.orc_closest_point.argtypes = [C.c_long, _f32p, _f32p, _f32p]
        L.orc_state_update.argtypes = [C.c_long, _f32p, _i32p, C.c_int, C.c_int]
        L.orc_frustum_planes.argtypes = [_f32p, _f32p]
        L.orc_frustum_aabb.argtypes = [C.c_long, _f32p, _f32p, _f32p, _i32p]
        L.orc_cube_box.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.orc_visible_frustum_culling.argtypes = [C.c_void_p, _f32p, _i32p, C.c_long]
        L.orc_visible_frustum_culling.restype = C.c_long
        L.orc_trs_matrix.argtypes = [_f32p, _f32p]
        L.orc_ingest_triangles.argtypes = [C.c_long, _f32p, _i32p, _f32p, _f32p]
        L.orc_ingest_sphere.argtypes = [C.c_long, _f32p, _i32p, _f32p, _f32p]
        _lib = L
    return _lib


def scale_matrix(s: float) -> np.ndarray:
    W = np.eye(4, dtype=np.float32)
    W[0, 0] = W[1, 1] = W[2, 2] = np.float32(s)
    return W.reshape(-1)


class OrcWorld:
    def __init__(self, scene=None, octree=True):
        self.L = lib()
        self.h = self.L.orc_create()
        self.n_geoms = 0
        self.n_bodies = 0
        if scene is not None:
            self.load(scene)
            self.build(octree)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def load(self, scene):
        for t in scene.tris:
            t = np.ascontiguousarray(t, dtype=np.float32)
            self.L.orc_add_geom(self.h, t.shape[0], t)
            self.n_geoms += 1
        ss = scene.sphere_start
        for b in range(scene.n_bodies):
            sp = np.ascontiguousarray(scene.obj_spheres[ss[b]:ss[b + 1]], dtype=np.float32)
            W = scale_matrix(scene.body_scale[b])
            self.L.orc_add_body(self.h, sp.shape[0], sp, W.ctypes.data_as(C.c_void_p),
                                np.ascontiguousarray(scene.body_pos[b]), np.ascontiguousarray(scene.body_vel[b]),
                                int(scene.body_active[b]), int(scene.body_gravity[b]))
            self.n_bodies += 1

    def build(self, octree=True):
        self.L.orc_build(self.h, 1 if octree else 0)

    def add_body_late(self, obj_spheres, pos, vel=(0, 0, 0), active=1, gravity=1, scale=1.0) -> int:
        """Physics::addEntity (physics/src/Physics.cpp:47-49): appended to the entity list after the build; it is
        integrated and collides from the next tick on. Only meaningful without the octree (canonical all-pairs order)."""
        sp = np.ascontiguousarray(obj_spheres, dtype=np.float32).reshape(-1, 4)
        W = scale_matrix(scale)
        self.L.orc_add_body(self.h, sp.shape[0], sp, W.ctypes.data_as(C.c_void_p), np.ascontiguousarray(pos, dtype=np.float32),
                            np.ascontiguousarray(vel, dtype=np.float32), int(active), int(gravity))
        self.n_bodies += 1
        return self.n_bodies - 1

    def entity_set_position(self, body, pos):
        """Entity::setPosition (model/src/Entity.cpp:373-391)"""
        self.L.orc_entity_set_position(self.h, int(body), np.ascontiguousarray(pos, dtype=np.float32))

    def step(self, ms=5, order=ORDER_REFERENCE):
        self.L.orc_step(self.h, ms, order)

    def integrate(self, ms=5):
        self.L.orc_integrate(self.h, ms)

    def collide(self, order=ORDER_REFERENCE):
        self.L.orc_collide(self.h, order)

    def detect(self, order=ORDER_REFERENCE):
        self.L.orc_detect(self.h, order)
        return self.hits()

    def set_record(self, on: bool):
        self.L.orc_set_record(self.h, 1 if on else 0)

    def hits(self):
        n = self.L.orc_hit_count(self.h)
        a = np.zeros((max(n, 1), 5), dtype=np.int32)
        f = np.zeros((max(n, 1), 8), dtype=np.float32)
        if n:
            self.L.orc_get_hits(self.h, a, f)
        return a[:n], f[:n]

    def contact_sets(self, hits=None):
        """Same convention as RefWorld.contact_sets (body ids / geom ids, sorted unique)."""
        a, _ = self.hits() if hits is None else hits
        ss = a[a[:, 0] == 0][:, 1:].copy()
        st = a[a[:, 0] == 1][:, 1:].copy()
        if len(ss):
            swap = (ss[:, 0] > ss[:, 2]) | ((ss[:, 0] == ss[:, 2]) & (ss[:, 1] > ss[:, 3]))
            ss[swap] = ss[swap][:, [2, 3, 0, 1]]
            ss = np.unique(ss, axis=0)
        if len(st):
            st = np.unique(st, axis=0)
        return ss.reshape(-1, 4), st.reshape(-1, 4)

    def counters(self):
        c = np.zeros(2, dtype=np.int64)
        self.L.orc_get_counters(self.h, c)
        return dict(tests_ss=int(c[0]), tests_st=int(c[1]))

    def state(self):
        n = self.n_bodies
        pos = np.zeros((n, 3), np.float32)
        vel = np.zeros((n, 3), np.float32)
        fl = np.zeros(n, np.int32)
        mt = np.zeros((n, 3), np.float32)
        pt = np.zeros((n, 3), np.float32)
        self.L.orc_get_state(self.h, pos, vel, fl, mt, pt)
        return dict(pos=pos, vel=vel, flags=fl, model_t=mt, prev_t=pt)

    def model_matrices(self):
        cur = np.zeros((self.n_bodies, 16), np.float32)
        prev = np.zeros((self.n_bodies, 16), np.float32)
        self.L.orc_get_model_matrices(self.h, cur, prev)
        return cur.reshape(-1, 4, 4), prev.reshape(-1, 4, 4)

    def angular(self):
        ap = np.zeros((self.n_bodies, 3), np.float32)
        av = np.zeros((self.n_bodies, 3), np.float32)
        self.L.orc_get_angular(self.h, ap, av)
        return ap, av

    def set_state(self, body, pos=None, vel=None):
        p = None if pos is None else np.ascontiguousarray(pos, np.float32)
        v = None if vel is None else np.ascontiguousarray(vel, np.float32)
        self.L.orc_set_state(self.h, body, None if p is None else p.ctypes.data_as(C.c_void_p),
                             None if v is None else v.ctypes.data_as(C.c_void_p))

    def set_flags(self, body, active=-1, gravity=-1, contact=-1):
        self.L.orc_set_flags(self.h, body, int(active), int(gravity), int(contact))

    def set_force(self, body, f):
        self.L.orc_set_force(self.h, body, np.ascontiguousarray(f, np.float32))

    def set_torque(self, body, t):
        self.L.orc_set_torque(self.h, body, np.ascontiguousarray(t, np.float32))

    def world_spheres(self, body, n):
        out = np.zeros((n, 4), np.float32)
        self.L.orc_get_world_spheres(self.h, body, out)
        return out

    def leaf_count(self):
        return self.L.orc_leaf_count(self.h)


def pred_sphere_triangle(c, r, t):
    c = np.ascontiguousarray(c, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, np.float32), (c.shape[0],)))
    out = np.zeros(c.shape[0], np.int32)
    lib().orc_pred_sphere_triangle(c.shape[0], c, r, t, out)
    return out


def pred_sphere_sphere(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 4)
    out = np.zeros(a.shape[0], np.int32)
    lib().orc_pred_sphere_sphere(a.shape[0], a, b, out)
    return out


def pred_sphere_cube(s, cube):
    s = np.ascontiguousarray(s, np.float32).reshape(-1, 4)
    cube = np.ascontiguousarray(cube, np.float32).reshape(-1, 6)
    out = np.zeros(s.shape[0], np.int32)
    lib().orc_pred_sphere_cube(s.shape[0], s, cube, out)
    return out


def pred_triangle_cube(t, cube):
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    cube = np.ascontiguousarray(cube, np.float32).reshape(-1, 6)
    out = np.zeros(t.shape[0], np.int32)
    lib().orc_pred_triangle_cube(t.shape[0], t, cube, out)
    return out


def closest_point(c, t):
    c = np.ascontiguousarray(c, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 9)
    out = np.zeros((c.shape[0], 3), np.float32)
    lib().orc_closest_point(c.shape[0], c, t, out)
    return out


def state_update(io18, flags, ms=5, nsteps=1):
    io = np.ascontiguousarray(io18, np.float32).reshape(-1, 18).copy()
    fl = np.ascontiguousarray(np.broadcast_to(np.asarray(flags, np.int32), (io.shape[0],)))
    lib().orc_state_update(io.shape[0], io, fl, ms, nsteps)
    return io


# ---- frustum culling and collider ingest ----
def frustum_planes(inv_vp):
    out = np.zeros(24, np.float32)
    lib().orc_frustum_planes(np.ascontiguousarray(inv_vp, np.float32).reshape(-1), out)
    return out.reshape(6, 4)


def frustum_aabb(planes, mins, maxs):
    mins = np.ascontiguousarray(mins, np.float32).reshape(-1, 3); maxs = np.ascontiguousarray(maxs, np.float32).reshape(-1, 3)
    out = np.zeros(mins.shape[0], np.int32)
    lib().orc_frustum_aabb(mins.shape[0], np.ascontiguousarray(planes, np.float32).reshape(-1), mins, maxs, out)
    return out


def cube_boxes(cubes6):
    """(centre, length, height, width) -> mins, maxs, |centre| exactly as OSP::getVisibleFrustumCulling forms them"""
    c = np.ascontiguousarray(cubes6, np.float32).reshape(-1, 6)
    mn = np.zeros((c.shape[0], 3), np.float32); mx = np.zeros((c.shape[0], 3), np.float32); key = np.zeros(c.shape[0], np.float32)
    k1 = np.zeros(1, np.float32)
    for i in range(c.shape[0]):
        lib().orc_cube_box(c[i], mn[i], mx[i], k1); key[i] = k1[0]
    return mn, mx, key


def visible_frustum_culling(world, inv_vp):
    n = world.leaf_count()
    out = np.zeros(max(n, 1), np.int32)
    k = lib().orc_visible_frustum_culling(world.h, np.ascontiguousarray(inv_vp, np.float32).reshape(-1), out, n)
    return out[:k]


def trs_matrix(trs9):
    out = np.zeros(16, np.float32)
    lib().orc_trs_matrix(np.ascontiguousarray(trs9, np.float32), out)
    return out.reshape(4, 4)


def ingest_triangles(xyz3, idx, trs9):
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
    out = np.zeros((idx.shape[0] // 3, 9), np.float32)
    lib().orc_ingest_triangles(idx.shape[0], np.ascontiguousarray(xyz3, np.float32).reshape(-1), idx, np.ascontiguousarray(trs9, np.float32), out.reshape(-1))
    return out


def ingest_sphere(xyz3, idx, trs9):
    idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
    out = np.zeros(4, np.float32)
    lib().orc_ingest_sphere(idx.shape[0], np.ascontiguousarray(xyz3, np.float32).reshape(-1), idx, np.ascontiguousarray(trs9, np.float32), out)
    return out
