"""Multi-GPU x-slabs: host-side partitioning and rendezvous.

One process per GPU: every rank owns the bodies whose centre starts in its slab of the x axis, static
geometry is replicated, and `rb_step` exchanges boundary spheres and migrants with the two neighbours
through NCCL on the world's own stream. The NCCL unique id is created by rank 0 and broadcast with
whatever the host already has -- here `torch.distributed`.

`SlabGroup` runs the same slab worlds inside ONE process in lockstep (`rb_step_group`, device-to-device
copies instead of NCCL): used to test the whole multi-rank path on a single GPU and usable by
single-process multi-GPU hosts.
"""
from __future__ import annotations

import copy
import ctypes as C
import glob
import os

import numpy as np

from . import _lib
from .world import World


def slab_bounds(rank: int, nranks: int, half: float = 1000.0):
    return (-half + 2.0 * half * rank / nranks, -half + 2.0 * half * (rank + 1) / nranks)


def owner_of(x: np.ndarray, nranks: int, half: float = 1000.0) -> np.ndarray:
    """Rank owning a body whose centre has this x (slab r covers [lo_r, hi_r); the outer slabs are open-ended)."""
    r = np.floor((np.asarray(x, np.float64) + half) * nranks / (2.0 * half)).astype(np.int64)
    return np.clip(r, 0, nranks - 1)


def partition_scene(scene, nranks: int, rank: int, half: float = 1000.0):
    """(sub-scene with the bodies this rank owns at t=0, global body ids, slab, r_max over all ranks).
    A body belongs to the slab its model translation lies in (single-sphere bodies: the sphere centre, as before)."""
    compound = scene.n_spheres != scene.n_bodies
    start = np.asarray(scene.sphere_start, np.int64)
    if compound:
        own = owner_of(scene.body_pos[:, 0], nranks, half) == rank
    else:
        own = owner_of(scene.body_pos[:, 0] + scene.obj_spheres[:, 0], nranks, half) == rank
    ids = np.nonzero(own)[0].astype(np.uint32)
    s = copy.copy(scene)
    s.body_pos = np.ascontiguousarray(scene.body_pos[own]); s.body_vel = np.ascontiguousarray(scene.body_vel[own])
    s.body_active = scene.body_active[own]; s.body_gravity = scene.body_gravity[own]; s.body_scale = scene.body_scale[own]
    if compound:
        cnt = (start[1:] - start[:-1])[own]
        idx = np.concatenate([np.arange(start[b], start[b + 1]) for b in ids]) if len(ids) else np.zeros(0, np.int64)
        s.obj_spheres = np.ascontiguousarray(scene.obj_spheres[idx]).reshape(-1, 4)
        s.sphere_start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    else:
        s.obj_spheres = np.ascontiguousarray(scene.obj_spheres[own]); s.sphere_start = np.arange(len(ids) + 1, dtype=np.int32)
    scale_per_sphere = np.repeat(scene.body_scale, start[1:] - start[:-1])
    rmax = float((scene.obj_spheres[:, 3] * scale_per_sphere).max()) if scene.n_bodies else 0.0
    return s, ids, slab_bounds(rank, nranks, half), rmax


def compound_layout(scene):
    """(sphere_stride, body_reach_hint) every rank of a compound-body scene must pass: the largest sphere count of any body and
    the farthest any body reaches from its model translation along x (|centre.x| + r, after the uniform scale)."""
    start = np.asarray(scene.sphere_start, np.int64)
    cnt = start[1:] - start[:-1]
    if scene.n_spheres == scene.n_bodies:
        return 0, 0.0
    sc = np.repeat(scene.body_scale, cnt).astype(np.float32)
    reach = np.abs(scene.obj_spheres[:, 0] * sc) + np.abs(scene.obj_spheres[:, 3] * sc)
    return int(cnt.max()), float(reach.max()) * 1.0001


def find_nccl() -> str:
    """The libnccl torch itself loaded (so the process holds ONE NCCL), else whatever the loader finds."""
    try:
        import nvidia.nccl  # noqa: F401
        for d in nvidia.nccl.__path__:
            hits = sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*")))
            if hits:
                return hits[0]
    except Exception:
        pass
    for pat in ("/usr/lib/x86_64-linux-gnu/libnccl.so.2", "/usr/local/graft/bin/libnccl.so.2*"):
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0]
    return "libnccl.so.2"


def broadcast_unique_id(make_id, rank: int, device=None) -> np.ndarray:
    """Rank 0 calls make_id() -> 128 bytes; everybody gets them via torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(_lib.RB_NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.from_numpy(np.frombuffer(bytes(make_id()), dtype=np.uint8).copy())
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def init_world(w: World, rank: int, nranks: int):
    """Join the slab communicator (call after w.build())."""
    import torch
    L = w.L
    path = find_nccl().encode()

    def make_id():
        buf = (C.c_uint8 * _lib.RB_NCCL_ID_BYTES)()
        rc = L.rb_dist_unique_id(path, buf)
        if rc:
            raise RuntimeError((L.rb_last_error(None) or b"").decode())
        return bytes(buf)

    uid = broadcast_unique_id(make_id, rank, torch.device("cuda", torch.cuda.current_device()))
    buf = (C.c_uint8 * _lib.RB_NCCL_ID_BYTES).from_buffer_copy(uid.tobytes())
    w._ck(L.rb_dist_init(w.h, path, buf))


def common_ghost_cap(n_local_bodies: int, r_max: float, slab_width: float, margin_cap: float = 0.0) -> int:
    """The halo capacity every rank must pass as World(halo_ghost_cap=...): the maximum over the ranks of the value
    each would pick by itself (torch.distributed all-reduce; call before building the worlds)."""
    import torch
    import torch.distributed as dist
    cap = int(_lib.load().rb_halo_ghost_cap_auto(int(n_local_bodies), float(r_max), float(margin_cap), float(slab_width)))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([cap], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def init_world_p2p(w: World, rank: int, nranks: int):
    """Join the slab ring with the fused compute+communication transport: every rank exports the CUDA-IPC
    handle of its halo blob, handles are all-gathered with torch.distributed, each rank opens its left /
    right neighbour's blob; from then on the integrate kernel stores ghost / migrant records straight into
    the neighbour's memory over NVLink (call after w.build())."""
    import torch
    import torch.distributed as dist
    L = w.L
    buf = (C.c_uint8 * _lib.RB_IPC_HANDLE_BYTES)()
    w._ck(L.rb_dist_p2p_export(w.h, buf))
    mine = torch.from_numpy(np.frombuffer(bytes(buf), dtype=np.uint8).copy()).to(torch.device("cuda", torch.cuda.current_device()))
    allh = [torch.zeros_like(mine) for _ in range(nranks)]
    dist.all_gather(allh, mine)
    hs = [t.cpu().numpy().tobytes() for t in allh]
    left = (C.c_uint8 * _lib.RB_IPC_HANDLE_BYTES).from_buffer_copy(hs[rank - 1]) if rank > 0 else None
    right = (C.c_uint8 * _lib.RB_IPC_HANDLE_BYTES).from_buffer_copy(hs[rank + 1]) if rank < nranks - 1 else None
    w._ck(L.rb_dist_p2p_connect(w.h, left, right))
    dist.barrier()


class SlabGroup:
    """N slab worlds of one scene in one process, stepped in lockstep."""

    def __init__(self, scene, nslabs: int, device: int = 0, **kw):
        self.L = _lib.load()
        self.worlds = []
        parts = [partition_scene(scene, nslabs, r) for r in range(nslabs)]
        # one halo buffer layout for everybody: the largest capacity any slab would pick by itself
        stride, reach = compound_layout(scene)
        cap = max(int(self.L.rb_halo_ghost_cap_auto(sub.n_bodies, float(max(rmax, reach)), 0.0, float(slab[1] - slab[0]))) for sub, _, slab, rmax in parts)
        for r, (sub, ids, slab, rmax) in enumerate(parts):
            w = World(sub, device=device, slab=slab, rank=r, nranks=nslabs, ids=ids, rm_max_hint=rmax, halo_ghost_cap=cap,
                      sphere_stride=stride, body_reach_hint=reach, **kw)
            self.worlds.append(w)
        for r, w in enumerate(self.worlds):
            left = self.worlds[r - 1].h if r > 0 else None
            right = self.worlds[r + 1].h if r < nslabs - 1 else None
            w._ck(self.L.rb_dist_attach_peers(w.h, left, right))
        self._arr = (C.c_void_p * nslabs)(*[w.h for w in self.worlds])

    def step(self, ms: int = 5):
        rc = self.L.rb_step_group(self._arr, len(self.worlds), ms)
        if rc:
            raise RuntimeError("rb_step_group failed: " + "; ".join((self.L.rb_last_error(w.h) or b"").decode() for w in self.worlds))

    def detect(self):
        rc = self.L.rb_detect_group(self._arr, len(self.worlds))
        if rc:
            raise RuntimeError("rb_detect_group failed: " + "; ".join((self.L.rb_last_error(w.h) or b"").decode() for w in self.worlds))

    def state(self, n_total: int):
        """Global state arrays indexed by global body id."""
        out = dict(pos=np.zeros((n_total, 3), np.float32), vel=np.zeros((n_total, 3), np.float32), flags=np.zeros(n_total, np.int32),
                   model_t=np.zeros((n_total, 3), np.float32), prev_t=np.zeros((n_total, 3), np.float32), owner=np.full(n_total, -1, np.int32))
        for r, w in enumerate(self.worlds):
            s = w.state()
            ids = w.body_ids()
            assert (out["owner"][ids] == -1).all(), "a body is owned twice"
            for k in ("pos", "vel", "flags", "model_t", "prev_t"):
                out[k][ids] = s[k]
            out["owner"][ids] = r
        return out

    def contact_sets(self):
        ss, st = [], []
        for w in self.worlds:
            a, b = w.contact_sets()
            ss.append(a); st.append(b)
        ss = np.unique(np.concatenate(ss), axis=0) if any(len(a) for a in ss) else np.zeros((0, 4), np.int32)
        st = np.unique(np.concatenate(st), axis=0) if any(len(a) for a in st) else np.zeros((0, 4), np.int32)
        return ss, st

    def close(self):
        for w in self.worlds:
            w.close()
