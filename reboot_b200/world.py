"""Thin Python host mirror of the reference's physics interface over the C ABI.

`World` plays the role of the reference's `Physics` object (physics/include/Physics.h:29-52):
add static geometry / add models (`Physics::addEntities`), `build`, `step` (one MasterClock
kinematics tick, engine/src/MasterClock.cpp:36-45) and `contacts` (`Physics::visualize`'s payload).
Everything computes in libreboot_b200.so on the GPU; this file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RB_ACTIVE, RB_CONTACT, RB_GRAVITY, CONTACT_DTYPE, RbConfig, RbStats, ptr


class RebootError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"reboot_b200 error {code}: {msg}")
        self.code = code


class World:
    def __init__(self, scene=None, device: int = 0, stream=None, max_contacts: int = 0, margin_cap: float | None = None,
                 slab=None, rank: int = 0, nranks: int = 1, build: bool = True, ids=None, rm_max_hint: float = 0.0,
                 halo_ghost_cap: int = 0, sphere_stride: int = 0, body_reach_hint: float = 0.0, slab_forces: bool = False):
        self.L = _lib.load()
        cfg = RbConfig()
        self.L.rb_config_default(C.byref(cfg))
        cfg.device = device
        cfg.stream = stream
        cfg.max_contacts = max_contacts
        if margin_cap is not None:
            cfg.margin_cap = margin_cap
        if slab is not None:
            cfg.slab_lo, cfg.slab_hi = slab
        cfg.rank, cfg.nranks = rank, nranks
        cfg.rm_max_hint = rm_max_hint
        cfg.halo_ghost_cap = int(halo_ghost_cap)
        cfg.sphere_stride = int(sphere_stride); cfg.body_reach_hint = float(body_reach_hint); cfg.slab_forces = 1 if slab_forces else 0
        self.cfg = cfg
        self._ids = ids
        h = C.c_void_p()
        rc = self.L.rb_world_create(C.byref(cfg), C.byref(h))
        if rc:
            raise RebootError(rc, (self.L.rb_last_error(None) or b"").decode())
        self.h = h
        self.n_bodies = 0
        self.n_spheres = 0
        if scene is not None:
            self.load(scene)
            if build:
                self.build()

    # -- lifetime --
    def close(self):
        if getattr(self, "h", None):
            self.L.rb_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RebootError(rc, (self.L.rb_last_error(self.h) or b"").decode())

    # -- scene construction (Physics::addEntities) --
    def add_static_geometry(self, tris) -> int:
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
        gid = C.c_uint32()
        self._ck(self.L.rb_add_static_geometry(self.h, t.shape[0], ptr(t), C.byref(gid)))
        return gid.value

    def add_model(self, obj_spheres, pos, vel=(0, 0, 0), flags=RB_ACTIVE | RB_GRAVITY, world_xform=None) -> int:
        o = np.ascontiguousarray(obj_spheres, dtype=np.float32).reshape(-1, 4)
        init = _lib.RbBodyInit()
        init.pos[:] = [float(x) for x in pos]
        init.vel[:] = [float(x) for x in vel]
        init.flags = int(flags)
        X = None if world_xform is None else np.ascontiguousarray(world_xform, dtype=np.float32).reshape(16)
        bid = C.c_uint32()
        self._ck(self.L.rb_add_model(self.h, o.shape[0], ptr(o), ptr(X), C.byref(init), C.byref(bid)))
        self.n_bodies += 1
        self.n_spheres += o.shape[0]
        return bid.value

    def add_models(self, sphere_start, obj_spheres, pos, vel, flags, scale=None) -> int:
        ss = np.ascontiguousarray(sphere_start, dtype=np.uint32)
        o = np.ascontiguousarray(obj_spheres, dtype=np.float32).reshape(-1, 4)
        p = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(flags, dtype=np.uint32)
        s = None if scale is None else np.ascontiguousarray(scale, dtype=np.float32)
        n = p.shape[0]
        first = C.c_uint32()
        self._ck(self.L.rb_add_models(self.h, n, ptr(ss), ptr(o), ptr(s), ptr(p), ptr(v), ptr(f), C.byref(first)))
        self.n_bodies += n
        self.n_spheres += o.shape[0]
        return first.value

    def load(self, scene):
        for t in scene.tris:
            self.add_static_geometry(t)
        flags = (scene.body_active.astype(np.uint32) * RB_ACTIVE) | (scene.body_gravity.astype(np.uint32) * RB_GRAVITY)
        scale = None if np.all(scene.body_scale == 1.0) else scene.body_scale
        self.add_models(scene.sphere_start, scene.obj_spheres, scene.body_pos, scene.body_vel, flags, scale)

    def build(self):
        if self._ids is not None:
            ids = np.ascontiguousarray(self._ids, dtype=np.uint32)
            self._ck(self.L.rb_set_body_ids(self.h, ptr(ids), ids.shape[0]))
        self._ck(self.L.rb_build(self.h))

    def body_ids(self) -> np.ndarray:
        """Global ids of the bodies this world currently owns, row-aligned with state()."""
        n = self.stats()["n_bodies"]
        out = np.zeros(max(n, 1), np.uint32)
        self._ck(self.L.rb_get_body_ids(self.h, ptr(out)))
        return out[:n]

    # -- the tick --
    def step(self, ms: int = 5):
        self._ck(self.L.rb_step(self.h, ms))

    def integrate(self, ms: int = 5):
        self._ck(self.L.rb_integrate(self.h, ms))

    def collide(self):
        self._ck(self.L.rb_collide(self.h))

    def detect(self):
        self._ck(self.L.rb_detect(self.h))

    def sync(self):
        self._ck(self.L.rb_sync(self.h))

    def step_host(self, ms, force4=None, pos4=None) -> int:
        n = C.c_uint64()
        self._ck(self.L.rb_step_host(self.h, ms, ptr(force4) if isinstance(force4, np.ndarray) else force4,
                                     ptr(pos4) if isinstance(pos4, np.ndarray) else pos4, C.byref(n)))
        return n.value

    # -- results --
    def contact_count(self) -> int:
        n = C.c_uint64()
        self._ck(self.L.rb_contact_count(self.h, C.byref(n)))
        return n.value

    def contacts(self, sorted: bool = True) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self.L.rb_query_contacts(self.h, None, 0, C.byref(n), 0))
        out = np.zeros(max(n.value, 1), dtype=CONTACT_DTYPE)
        assert out.dtype.itemsize == C.sizeof(_lib.RbContact)
        self._ck(self.L.rb_query_contacts(self.h, ptr(out), out.shape[0], C.byref(n), 1 if sorted else 0))
        return out[: n.value]

    def contact_sets(self, contacts=None):
        """(ss[n,4] = bodyA,sphereA,bodyB,sphereB ; st[n,4] = body,sphere,geom,tri), sorted unique --
        body / geom ids as the C ABI reports them."""
        c = self.contacts() if contacts is None else contacts
        a = np.stack([c["body"], c["sphere"], c["other"], c["other_prim"]], axis=1).astype(np.int32)
        ss = a[c["kind"] == 0]
        st = a[c["kind"] == 1]
        if len(ss):
            swap = (ss[:, 0] > ss[:, 2]) | ((ss[:, 0] == ss[:, 2]) & (ss[:, 1] > ss[:, 3]))
            ss[swap] = ss[swap][:, [2, 3, 0, 1]]
            ss = np.unique(ss, axis=0)
        if len(st):
            st = np.unique(st, axis=0)
        return ss.reshape(-1, 4), st.reshape(-1, 4)

    def state(self):
        st = self.stats()
        n = st["n_bodies"]
        pos = np.zeros((n, 4), np.float32)
        vel = np.zeros((n, 4), np.float32)
        fl = np.zeros(n, np.uint32)
        mt = np.zeros((n, 4), np.float32)
        pt = np.zeros((n, 4), np.float32)
        self._ck(self.L.rb_get_state(self.h, ptr(pos), ptr(vel), ptr(fl)))
        self._ck(self.L.rb_get_model_translations(self.h, ptr(mt), ptr(pt)))
        return dict(pos=pos[:, :3].copy(), vel=vel[:, :3].copy(), flags=fl.astype(np.int32), model_t=mt[:, :3].copy(), prev_t=pt[:, :3].copy())

    def model_matrices(self):
        """(model[n,4,4], prev_model[n,4,4]) -- Entity::getMVP / getPrevMVP model buffers, row-major."""
        n = self.stats()["n_bodies"]
        cur = np.zeros((n, 16), np.float32)
        prev = np.zeros((n, 16), np.float32)
        self._ck(self.L.rb_get_model_matrices(self.h, ptr(cur), ptr(prev)))
        return cur.reshape(n, 4, 4), prev.reshape(n, 4, 4)

    def angular(self):
        n = self.stats()["n_bodies"]
        ap = np.zeros((n, 4), np.float32)
        av = np.zeros((n, 4), np.float32)
        self._ck(self.L.rb_get_angular(self.h, ptr(ap), ptr(av)))
        return ap[:, :3].copy(), av[:, :3].copy()

    def world_spheres(self):
        n = self.stats()["n_spheres"]
        out = np.zeros((n, 4), np.float32)
        self._ck(self.L.rb_get_world_spheres(self.h, ptr(out)))
        return out

    @staticmethod
    def _pad4(a):
        a = np.asarray(a, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((a.shape[0], 4), np.float32)
        out[:, :3] = a
        return out

    def set_state(self, first, pos=None, vel=None):
        p = None if pos is None else self._pad4(pos)
        v = None if vel is None else self._pad4(vel)
        count = (p if p is not None else v).shape[0]
        self._ck(self.L.rb_set_state(self.h, first, count, ptr(p), ptr(v)))

    def entity_set_position(self, body: int, pos):
        """Entity::setPosition (model/src/Entity.cpp:373-391): position, _prevMVP and _mvp move together."""
        p = np.ascontiguousarray(pos, dtype=np.float32).reshape(3)
        self._ck(self.L.rb_entity_set_position(self.h, int(body), ptr(p)))

    def set_flags(self, first, values, mask):
        v = np.ascontiguousarray(values, dtype=np.uint32).reshape(-1)
        self._ck(self.L.rb_set_flags(self.h, first, v.shape[0], ptr(v), mask))

    def set_force(self, first, f):
        f4 = self._pad4(f)
        self._ck(self.L.rb_set_force(self.h, first, f4.shape[0], ptr(f4)))
        self.sync()

    def set_torque(self, first, t):
        t4 = self._pad4(t)
        self._ck(self.L.rb_set_torque(self.h, first, t4.shape[0], ptr(t4)))

    def set_angular(self, first, apos=None, avel=None):
        a = None if apos is None else self._pad4(apos)
        b = None if avel is None else self._pad4(avel)
        count = (a if a is not None else b).shape[0]
        self._ck(self.L.rb_set_angular(self.h, first, count, ptr(a), ptr(b)))

    def stats(self) -> dict:
        s = RbStats()
        self._ck(self.L.rb_get_stats(self.h, C.byref(s)))
        return {k: (list(getattr(s, k)) if k == "list_rebuilds_by_reason" else getattr(s, k)) for k, _ in RbStats._fields_}

    def profile(self, on: bool):
        self._ck(self.L.rb_profile(self.h, 1 if on else 0))

    def kernel_ms(self):
        out = np.zeros(5, np.float32)
        n = C.c_uint64()
        self._ck(self.L.rb_get_kernel_ms(self.h, ptr(out), C.byref(n)))
        return dict(integrate_bin=float(out[0]), scan=float(out[1]), scatter=float(out[2]), collide=float(out[3]), halo=float(out[4])), n.value

    def launch_count(self) -> int:
        return int(self.L.rb_launch_count(self.h))
