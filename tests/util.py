"""Shared helpers for the parity tests."""
import copy

import numpy as np

from reboot_b200 import scenes

F_ACTIVE, F_GRAVITY, F_CONTACT = 1, 2, 4


def golden_scenes():
    sc4 = scenes.config4(300, 24, 9)
    sc4.body_pos[:, 1] -= 1.5
    return {
        "c1": (scenes.config1(), [0, 9, 99, 299, 599]),
        "compound": (sc4, [0, 19, 59, 119]),
        "pile": (scenes.pile(150, 5, 0.5, 0), [0, 1, 5]),
        "c2": (scenes.config2(2000, 3), [0, 3]),
    }


def scene_at(sc, pos, vel):
    s = copy.copy(sc)
    s.body_pos = np.ascontiguousarray(pos, np.float32)
    s.body_vel = np.ascontiguousarray(vel, np.float32)
    return s


def restart(world_cls, sc, g, prefix, k, **kw):
    """World of any implementation restarted at the reference's state after k ticks: positions and
    velocities at build, then velocities again (the priming _updateKinematics(0) applies the friction
    factor to gravity-off bodies) and the contact flags."""
    pos, vel, fl = g[f"{prefix}_k{k}_s0_pos"], g[f"{prefix}_k{k}_s0_vel"], g[f"{prefix}_k{k}_s0_flags"]
    w = world_cls(scene_at(sc, pos, vel), **kw)
    n = pos.shape[0]
    if hasattr(w, "set_flags") and w.__class__.__name__ == "World":
        w.set_state(0, vel=vel)
        w.set_flags(0, fl.astype(np.uint32), F_CONTACT)
    else:
        for b in range(n):
            w.set_state(b, vel=vel[b])
            w.set_flags(b, contact=1 if fl[b] & F_CONTACT else 0)
    return w


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def order_independent_mask(hits_i, n_geoms, n_bodies, tri_base, frozen_ss=None, frozen_st=None):
    """Bodies whose reference result in this tick cannot depend on the reference's leaf / hash
    iteration order (SURVEY.md §8c level L2):
      * no sphere-sphere contact at the frozen state and no sphere-sphere hit in the log, and the
        sphere-triangle hits (if any) have strictly ascending global triangle ids in the reference's own
        hit order (== the canonical order)                                        -> compare bit-exactly
      * an ISOLATED sphere-sphere pair: each of the two bodies overlaps exactly one other body at the
        frozen state (each other), touches no triangle at the frozen state, and the log holds exactly
        one hit for the pair                                                       -> compare with the
        sphere-sphere tolerance (which body plays "modelA" follows pointer-hash order in the reference)
    Returns (mask_exact, mask_ss)."""
    ss_cnt = np.zeros(n_bodies, np.int32)
    st_cnt = np.zeros(n_bodies, np.int32)
    st_bad = np.zeros(n_bodies, bool)
    last_tri = np.full(n_bodies, -1, np.int64)
    partner = np.full(n_bodies, -1, np.int64)
    for kind, ea, sa, eb, pb in hits_i:
        a = ea - n_geoms
        if kind == 0:
            b = eb - n_geoms
            ss_cnt[a] += 1; ss_cnt[b] += 1
            partner[a] = b; partner[b] = a
        else:
            t = tri_base[eb] + pb
            st_cnt[a] += 1
            if t <= last_tri[a]:
                st_bad[a] = True
            last_tri[a] = t
    deg = np.zeros(n_bodies, np.int32)
    touches_tri = np.zeros(n_bodies, bool)
    if frozen_ss is not None and len(frozen_ss):
        np.add.at(deg, frozen_ss[:, 0], 1); np.add.at(deg, frozen_ss[:, 2], 1)
    if frozen_st is not None and len(frozen_st):
        touches_tri[frozen_st[:, 0]] = True
    exact = (ss_cnt == 0) & (deg == 0) & ~st_bad
    ss_ok = (ss_cnt == 1) & (st_cnt == 0) & (deg == 1) & ~touches_tri
    ss_ok &= np.where(partner >= 0, ss_ok[np.clip(partner, 0, n_bodies - 1)], False)
    return exact, ss_ok


def tri_bases(sc):
    return np.concatenate([[0], np.cumsum([t.shape[0] for t in sc.tris])]).astype(np.int64)
