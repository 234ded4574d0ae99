"""GPU suite (-m gpu): the boundary calls either side of the tick that change the world after the build --
Physics::addEntity (physics/src/Physics.cpp:47-49) = rb_add_model / rb_add_models after rb_build, and
Entity::setPosition (model/src/Entity.cpp:373-391) = rb_entity_set_position -- against the CPU oracle doing the same."""
import os
import subprocess

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from reboot_b200 import World, scenes
    from oracle import port_binding as pb
    return World, scenes, pb


def _same(w, o, tag, translations=True):
    a, b = w.state(), o.state()
    assert a["pos"].shape == b["pos"].shape, tag
    for key in ("pos", "vel") + (("model_t", "prev_t") if translations else ()):
        assert np.array_equal(bits(a[key]), bits(b[key])), (tag, key)
    assert np.array_equal(a["flags"], b["flags"]), tag


def _same_contacts(w, o, pb, tag):
    ss_g, st_g = w.contact_sets()
    ss_o, st_o = o.contact_sets(o.hits())
    assert np.array_equal(ss_g, ss_o) and np.array_equal(st_g, st_o), tag


@pytest.mark.parametrize("late_spheres", [1, 2])
def test_add_model_after_build_matches_oracle(mods, late_spheres):
    """A body added in the middle of a run (one sphere: the world stays on the list kernels; two spheres: it becomes a
    compound world) steps from the next tick on exactly like the oracle's late entity; the bodies that were there keep
    their state bit for bit across the re-layout, contact flags and both model translations included. Then a batch
    (rb_add_models after the build) with a larger radius, which changes every grid."""
    World, scenes, pb = mods
    sc = scenes.falling_spheres(400, 24, 2.0, 0.5, (0.3, 2.5), True, 17, "late")
    w = World(sc)
    o = pb.OrcWorld(sc, octree=False)
    for k in range(70):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "before")
    before = w.state()
    assert (before["flags"] & 4).any(), "some bodies must be in contact when the world is re-laid out"
    # one late body, dropped onto the pile next to body 0
    p0 = before["pos"][0]
    obj = [[0, 0, 0, 0.5]] if late_spheres == 1 else [[-0.3, 0, 0, 0.4], [0.3, 0.1, 0, 0.4]]
    pos = (float(p0[0]) + 0.6, float(p0[1]) + 1.2, float(p0[2]) + 0.1)
    bid = w.add_model(obj, pos, vel=(0.0, -1.0, 0.0))
    assert bid == 400 and w.stats()["n_bodies"] == 401
    oid = o.add_body_late(obj, pos, (0.0, -1.0, 0.0))
    assert oid == 400
    after = w.state()
    for key in ("pos", "vel", "model_t", "prev_t"):
        assert np.array_equal(bits(after[key][:400]), bits(before[key])), key      # untouched by the re-layout
    assert np.array_equal(after["flags"][:400], before["flags"])
    assert np.array_equal(bits(after["pos"][400]), bits(np.array(pos, np.float32)))
    for k in range(3):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "3 ticks after the add")           # (from its second tick on the late entity's _prevMVP is a real pose too)
    for k in range(120):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
        if k % 40 == 39:
            _same(w, o, ("late", k))
    _same_contacts(w, o, pb, "late")
    s = w.state()
    assert s["pos"][400, 1] < pos[1], "the late body has come down"
    # a batch: three more bodies, one with a larger radius (the grids are sized by it)
    p1 = s["pos"][5]
    start = np.array([0, 1, 2, 3], np.uint32)
    objs = np.array([[0, 0, 0, 0.5], [0, 0, 0, 0.8], [0, 0, 0, 0.3]], np.float32)
    poss = np.array([[p1[0] + 0.4, p1[1] + 1.5, p1[2]], [p1[0] - 3.0, p1[1] + 2.0, p1[2] + 1.0], [p1[0] + 5.0, p1[1] + 0.8, p1[2] - 2.0]], np.float32)
    vels = np.zeros((3, 3), np.float32)
    first = w.add_models(start, objs, poss, vels, np.full(3, 3, np.uint32))
    assert first == 401
    for i in range(3):
        o.add_body_late(objs[i:i + 1], poss[i], vels[i])
    for k in range(150):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
        if k % 50 == 49:
            _same(w, o, ("batch", k))
    _same_contacts(w, o, pb, "batch")
    assert w.stats()["n_bodies"] == 404 and (w.state()["pos"][401:, 1] < poss[:, 1]).all()
    w.close()


def test_add_model_keeps_forces_and_host_round_trip(mods):
    """Forces set before a post-build add still act on their bodies afterwards (the force array travels with the
    re-layout), and the per-tick host call keeps working with the new body count."""
    World, scenes, pb = mods
    sc = scenes.falling_spheres(300, 20, 2.0, 0.5, (0.3, 1.0), True, 5, "lateforce")
    w = World(sc)
    o = pb.OrcWorld(sc, octree=False)
    for k in range(80):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    f = np.zeros((300, 3), np.float32); f[:, 0] = 25.0
    w.set_force(0, f)
    for b in range(300):
        o.set_force(b, f[b])
    w.step(5); o.step(5, pb.ORDER_CANONICAL)
    w.add_model([[0, 0, 0, 0.5]], (0.0, 30.0, 0.0))
    o.add_body_late([[0, 0, 0, 0.5]], (0.0, 30.0, 0.0))
    for k in range(30):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "forces after the add")
    assert np.abs(w.state()["vel"][:300, 0]).max() > 0.01, "the forces must have moved somebody"
    pos4 = np.zeros((301, 4), np.float32)
    n = w.step_host(5, None, pos4)
    o.step(5, pb.ORDER_CANONICAL)
    assert np.array_equal(bits(pos4[:, :3].copy()), bits(o.state()["pos"])) and n == w.contact_count()
    w.close()


def test_entity_set_position_matches_oracle(mods):
    """Entity::setPosition from the host on a world whose bodies carry a translated W (Q1: the stored position is W.t + p):
    position, _mvp and _prevMVP move like the oracle's, and the run continues bit-identically."""
    World, scenes, pb = mods
    rng = np.random.default_rng(3)
    sc = scenes.falling_spheres(200, 16, 2.0, 0.5, (0.3, 2.0), True, 9, "setpos")
    w = World(sc); o = pb.OrcWorld(sc, octree=False)
    for k in range(60):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    ids = rng.choice(200, 20, replace=False)
    s = w.state()
    for b in ids:
        p = s["pos"][b] + np.array([0.3, 0.8, -0.2], np.float32)
        w.entity_set_position(int(b), p); o.entity_set_position(int(b), p)
    _same(w, o, "right after setPosition")
    for k in range(80):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "after setPosition")
    _same_contacts(w, o, pb, "after setPosition")
    w.close()
    # translated W: a single world built body by body with world transforms that carry a translation
    tris = sc.tris[0]
    w = World(); o = pb.OrcWorld()
    w.add_static_geometry(tris)
    o.L.orc_add_geom(o.h, tris.shape[0], np.ascontiguousarray(tris, np.float32)); o.n_geoms += 1
    import ctypes as C
    for b in range(40):      # (3 m apart: sphere-triangle contacts only)
        X = np.eye(4, dtype=np.float32); X[0, 3], X[1, 3], X[2, 3] = 0.25 * (b % 3), 0.5, -0.125 * (b % 5)
        p = np.array([-10.0 + 3.0 * (b % 7), 2.5 + 0.125 * (b % 4), -8.0 + 3.0 * (b // 7)], np.float32)
        w.add_model([[0, 0, 0, 0.5]], p, world_xform=X)
        o.L.orc_add_body(o.h, 1, np.array([[0, 0, 0, 0.5]], np.float32), X.reshape(-1).ctypes.data_as(C.c_void_p), np.ascontiguousarray(p), np.zeros(3, np.float32), 1, 1)
        o.n_bodies += 1
    w.build(); o.build(False)
    for k in range(120):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "translated W, before setPosition")
    assert (w.state()["flags"] & 4).any()
    for b in (3, 7, 11):
        p = np.array([-9.0 + b, 4.0, 9.0], np.float32)
        w.entity_set_position(b, p); o.entity_set_position(b, p)
    _same(w, o, "translated W, right after setPosition")
    for k in range(60):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
    _same(w, o, "translated W, after setPosition")
    w.close()


def test_facade_add_entity_set_position_and_mvp(mods, tmp_path):
    """Reference-style host code (include/reboot/Physics.hpp): Physics::addEntity in the middle of a run, Entity::setPosition
    and Entity::getMVP()/getPrevMVP() -- the same calls through the ctypes mirror give the same state and matrices."""
    World, scenes, pb = mods
    exe = str(tmp_path / "facade_add_test")
    lib = os.path.join(ROOT, "reboot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_add_test.cpp"),
                           "-L" + lib, "-lreboot_b200", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    rows = np.array([[float(x) for x in ln.split()] for ln in out])
    # the same through the ctypes mirror
    G, c, half = 16, 2.0, 16.0
    tris = []
    for z in range(G):
        for x in range(G):
            b = (x * c - half, 0, z * c - half); b1 = ((x + 1) * c - half, 0, z * c - half)
            bs = (x * c - half, 0, (z + 1) * c - half); bs1 = ((x + 1) * c - half, 0, (z + 1) * c - half)
            tris.append(bs + b1 + b); tris.append(bs + bs1 + b1)
    w = World()
    w.add_static_geometry(np.array(tris, np.float32))
    for i in range(5):
        for j in range(5):
            w.add_model([[0, 0, 0, 0.5]], (np.float32(-8.0) + np.float32(3.1) * np.float32(i), np.float32(1.0) + np.float32(0.25) * np.float32((i * 7 + j * 3) % 5),
                                           np.float32(-7.0) + np.float32(2.9) * np.float32(j)))
    w.build()
    for k in range(60):
        w.step(5)
    X = np.eye(4, dtype=np.float32) * np.float32(1.5); X[3, 3] = 1.0; X[0, 3] = 0.25
    w.add_model([[0, 0, 0, 0.4]], (0.5, 4.0, 0.25), vel=(0.0, -0.5, 0.0), world_xform=X)
    for k in range(40):
        w.step(5)
    w.entity_set_position(3, (2.0, 3.0, -1.0))
    w.entity_set_position(25, (-3.0, 5.0, 2.0))
    for k in range(30):
        w.step(5)
    s = w.state(); cur, prev = w.model_matrices()
    assert rows.shape == (26, 3 + 3 + 16 + 16)
    assert np.array_equal(bits(rows[:, 0:3].astype(np.float32)), bits(s["pos"])) and np.array_equal(bits(rows[:, 3:6].astype(np.float32)), bits(s["vel"]))
    assert np.array_equal(bits(rows[:, 6:22].astype(np.float32)), bits(cur.reshape(26, 16))) and np.array_equal(bits(rows[:, 22:38].astype(np.float32)), bits(prev.reshape(26, 16)))
    assert cur[25, 0, 0] == np.float32(1.5) and s["pos"][25, 1] < 5.0        # the late entity keeps its scaled W and has come down from where setPosition put it
    w.close()
