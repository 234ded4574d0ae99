"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against
  * the committed golden fixtures generated from the reference's own compiled sources,
  * the CPU oracle (oracle/reboot_oracle.c) on the same seeded inputs,
  * size-independent properties at BASELINE.json's full C3 size.
Bar: bit-exact contact sets and integration; resolved state bit-exact vs the canonical-order oracle
and vs the reference wherever the reference's result cannot depend on its hash/leaf order."""
import ctypes as C

import numpy as np
import pytest

from util import bits, golden_scenes, order_independent_mask, restart, tri_bases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def World():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from reboot_b200 import World as W
    return W


@pytest.mark.parametrize("name", ["c1", "compound", "pile", "c2"])
def test_L1_frozen_contact_sets_match_reference(World, golden, name):
    """Integration bit-exact, then frozen-state detection gives exactly the reference's contact sets."""
    sc, cps = golden_scenes()[name]
    g = golden[name]
    for cp in cps:
        w = restart(World, sc, g, name, cp)
        w.integrate(sc.dt_ms)
        s = w.state()
        assert np.array_equal(bits(s["pos"]), bits(g[f"{name}_k{cp}_sint_pos"]))
        assert np.array_equal(bits(s["vel"]), bits(g[f"{name}_k{cp}_sint_vel"]))
        w.detect()
        ss, st = w.contact_sets()
        assert np.array_equal(ss, g[f"{name}_k{cp}_frozen_ss"]), (name, cp)
        assert np.array_equal(st, g[f"{name}_k{cp}_frozen_st"]), (name, cp)
        s2 = w.state()   # detection must not touch state
        assert np.array_equal(bits(s2["pos"]), bits(s["pos"])) and np.array_equal(s2["flags"], s["flags"])
        w.close()


@pytest.mark.parametrize("name", ["c1", "compound", "pile", "c2"])
def test_L2_one_tick_matches_reference_where_order_free(World, golden, name):
    """One tick from the reference's state: positions, velocities and flags equal the reference's next
    state bit-for-bit on bodies whose result is order-free; sphere-sphere resolved bodies within 1e-5
    relative (the "modelA"/"modelB" role follows pointer-hash order in the reference). Depths and
    normals of the reported contacts match the reference's hit log within 1e-5 relative."""
    sc, cps = golden_scenes()[name]
    g = golden[name]
    nb, ng = sc.n_bodies, len(sc.tris)
    tb = tri_bases(sc)
    for cp in cps:
        w = restart(World, sc, g, name, cp)
        w.step(sc.dt_ms)
        s = w.state()
        hi, hf = g[f"{name}_k{cp}_hits_i"], g[f"{name}_k{cp}_hits_f"]
        exact, ssok = order_independent_mask(hi, ng, nb, tb, g[f"{name}_k{cp}_frozen_ss"], g[f"{name}_k{cp}_frozen_st"])
        rp, rv, rf = g[f"{name}_k{cp}_s1_pos"], g[f"{name}_k{cp}_s1_vel"], g[f"{name}_k{cp}_s1_flags"]
        bad = (bits(s["pos"]) != bits(rp)).any(1) | (bits(s["vel"]) != bits(rv)).any(1)
        assert int((bad & exact).sum()) <= max(1, int(0.002 * nb)), (name, cp)
        assert np.allclose(s["pos"][ssok], rp[ssok], rtol=1e-5, atol=1e-6)
        assert np.allclose(s["vel"][ssok], rv[ssok], rtol=1e-5, atol=1e-6)
        if name in ("c1", "c2"):
            # contact flag: the reference keeps the result of the LAST octree leaf that holds the sphere
            # (Physics.cpp:180). Without leaves the canonical rule is "any hit of the last
            # sphere": identical for every body that touches at most one triangle-bearing leaf; only
            # bodies straddling leaves may differ, and only in the contact bit.
            ok = exact & ~bad
            one_leaf = g[f"{name}_k{cp}_leaves_touched"] <= 1
            assert np.array_equal(s["flags"][ok & one_leaf], rf[ok & one_leaf])
            assert np.array_equal(s["flags"][ok] & 3, rf[ok] & 3)
            assert (s["flags"][ok] != rf[ok]).mean() < 0.10
        # contact payload: depth / normal of every reference sphere-triangle hit of an order-free body
        c = w.contacts()
        mine = {(int(r["body"]), int(r["sphere"]), int(r["other"]), int(r["other_prim"])): r for r in c[c["kind"] == 1]}
        n_cmp = 0
        for (kind, ea, sa, eb, pb), f in zip(hi, hf):
            if kind != 1 or not exact[ea - ng] or bad[ea - ng]:
                continue
            r = mine[(ea - ng, sa, eb, pb)]
            assert np.allclose(r["depth"], f[0], rtol=1e-5, atol=1e-6)
            assert np.allclose(r["normal"], f[1:4], rtol=1e-5, atol=1e-6, equal_nan=True)
            n_cmp += 1
        if name == "c1" and cp >= 299:   # spheres land after ~180 ticks
            assert n_cmp > 100
        w.close()


@pytest.mark.parametrize("name,nsteps", [("c1", 250), ("compound", 60), ("pile", 12), ("c2", 6)])
def test_trajectory_bit_exact_vs_canonical_oracle(World, port, name, nsteps):
    """Same seeded scene stepped on the GPU and by the CPU oracle in canonical order: positions,
    velocities, flags, model / previous-model translations and contact sets identical every tick."""
    sc, _ = golden_scenes()[name]
    w = World(sc)
    o = port.OrcWorld(sc, octree=False)
    for k in range(nsteps):
        w.step(sc.dt_ms)
        o.step(sc.dt_ms, port.ORDER_CANONICAL)
        if k % 10 == 9 or k == nsteps - 1 or k < 2:
            a, b = w.state(), o.state()
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(b[key])), (name, k, key)
            assert np.array_equal(a["flags"], b["flags"]), (name, k)
            ga, gb = w.contact_sets(), o.contact_sets()
            assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1]), (name, k)
    st = w.stats()
    oc = o.counters()
    assert st["contacts_dropped"] == 0
    assert st["hits_st"] > 0 or name in ("pile", "c2")
    assert oc["tests_ss"] >= 0
    w.close()


def test_voronoi_micro_vectors_through_the_kernels(World, golden):
    """Every hand-built sphere/triangle micro vector (face, edge and vertex regions; penetrating,
    exactly touching = miss (Q5), separated; degenerate triangles) as one world: pair (i, i) is
    reported iff the reference predicate says hit."""
    m = golden["micro"]
    n = 2253   # the structured cases; the random bulk is covered by the scene tests
    w = World(max_contacts=1_000_000)   # spheres also hit other cases' triangles; only pair (i, i) is checked
    w.add_static_geometry(m["st_tri"][:n])
    obj = np.zeros((n, 4), np.float32); obj[:, 3] = m["st_r"][:n]
    w.add_models(np.arange(n + 1), obj, m["st_c"][:n], np.zeros((n, 3), np.float32), np.full(n, 3, np.uint32))
    w.build()
    w.detect()
    _, st = w.contact_sets()
    got = np.zeros(n, np.int32)
    diag = st[st[:, 0] == st[:, 3]]
    got[diag[:, 0]] = 1
    assert np.array_equal(got, m["st_hit"][:n])
    w.close()


def test_sphere_sphere_micro_vectors(World, golden):
    """Touching spheres are a HIT (Q5). Pairs are placed as two-body worlds sharing one big world."""
    m = golden["micro"]
    n = 300
    a, b = m["ss_a"][:n].copy(), m["ss_b"][:n].copy()
    w = World(max_contacts=600 * 600)   # the pairs overlap each other too; only pair (2k, 2k+1) is checked
    pos = np.zeros((2 * n, 3), np.float32); obj = np.zeros((2 * n, 4), np.float32)
    pos[0::2], pos[1::2] = a[:, :3], b[:, :3]
    obj[0::2, 3], obj[1::2, 3] = a[:, 3], b[:, 3]
    w.add_models(np.arange(2 * n + 1), obj, pos, np.zeros_like(pos), np.full(2 * n, 1, np.uint32))
    w.build()
    w.detect()
    ss, _ = w.contact_sets()
    got = np.zeros(n, np.int32)
    for r in ss:
        if r[0] % 2 == 0 and r[2] == r[0] + 1:
            got[r[0] // 2] = 1
    assert np.array_equal(got, m["ss_hit"][:n])
    w.close()


def test_integration_bit_exact_with_forces_and_angular(World, golden, port):
    """StateVector::update incl. force, torque, angular state, gravity / contact friction, inactive
    bodies; setForce / setTorque gating (StateVector.cpp:147-167)."""
    m = golden["micro"]
    io, flags = m["sv_in"], m["sv_flags"]
    n = io.shape[0]
    w = World()
    obj = np.zeros((n, 4), np.float32); obj[:, 3] = 1e-3
    w.add_models(np.arange(n + 1), obj, io[:, 0:3], io[:, 3:6], (flags & 3).astype(np.uint32))
    w.build()
    # build primes with _updateKinematics(0): friction applies to gravity-off bodies -> restore velocity
    w.set_state(0, vel=io[:, 3:6])
    w.set_flags(0, flags.astype(np.uint32), 4)
    w.set_angular(0, io[:, 6:9], io[:, 9:12])
    w.set_force(0, io[:, 12:15])
    w.set_torque(0, io[:, 15:18])
    gate = ((flags & 4) != 0) | ((flags & 2) == 0)
    exp_in = io.copy()
    exp_in[~gate, 12:] = 0
    for k in range(7):
        w.integrate(5)
    s = w.state(); ap, av = w.angular()
    exp = port.state_update(exp_in, flags, 5, 7)
    assert np.array_equal(bits(s["pos"]), bits(exp[:, 0:3])) and np.array_equal(bits(s["vel"]), bits(exp[:, 3:6]))
    assert np.array_equal(bits(ap), bits(exp[:, 6:9])) and np.array_equal(bits(av), bits(exp[:, 9:12]))
    # rows whose force passes the gate also match the reference-generated fixture directly
    ref7 = m["sv_out_5ms_7"]
    assert np.array_equal(bits(s["pos"][gate]), bits(ref7[gate, 0:3]))
    w.close()


def test_edge_cases(World, port):
    from reboot_b200 import RebootError, scenes
    # empty world
    w = World(); w.build(); w.step(5); w.detect(); assert w.contact_count() == 0 and len(w.contacts()) == 0; w.close()
    # static geometry only
    tris, _ = scenes.terrain_triangles(8, 2.0, 1)
    w = World(); w.add_static_geometry(tris); w.build(); w.step(5); assert w.contact_count() == 0; w.close()
    # ragged compound bodies (1, 3 and 5 spheres), one inactive body, one sphere outside the world cube (Q7)
    w = World(); o = port.OrcWorld()
    w.add_static_geometry(tris); o.L.orc_add_geom(o.h, tris.shape[0], tris); o.n_geoms = 1
    rng = np.random.default_rng(4)
    specs = [(1, (0.0, 1.2, 0.0), 3), (3, (2.0, 1.3, 1.0), 3), (5, (-3.0, 1.4, 2.0), 3), (1, (1.0, 1.1, -3.0), 2), (1, (1000.6, 0.0, 0.0), 3), (1, (1000.4, 0.0, 0.0), 3)]
    for ns, p, fl in specs:
        obj = np.zeros((ns, 4), np.float32); obj[:, :3] = rng.uniform(-0.3, 0.3, (ns, 3)); obj[:, 3] = 0.5
        w.add_model(obj, p, (0, -1, 0), fl)
        o.L.orc_add_body(o.h, ns, obj, None, np.array(p, np.float32), np.array([0, -1, 0], np.float32), fl & 1, (fl >> 1) & 1); o.n_bodies += 1
    w.build(); o.build(False)
    for k in range(40):
        w.step(5); o.step(5, port.ORDER_CANONICAL)
    a, b = w.state(), o.state()
    assert np.array_equal(bits(a["pos"]), bits(b["pos"])) and np.array_equal(a["flags"], b["flags"])
    assert np.array_equal(bits(a["pos"][3]), bits(np.array(specs[3][1], np.float32)))   # inactive body never moves
    assert a["flags"][0] & 4 and a["flags"][2] & 4                                       # resting bodies are in contact
    # error paths: add after build, NaN input, capacity
    with pytest.raises(RebootError):
        w.add_static_geometry(tris)
    w.close()
    w = World()
    with pytest.raises(RebootError):
        w.add_model([[0, 0, 0, 0.5]], (float("nan"), 0, 0))
    with pytest.raises(RebootError):
        w.add_model(np.zeros((0, 4), np.float32), (0, 0, 0))
    w.close()
    sc = scenes.pile(64, 5, 0.5, 0)
    w = World(sc, max_contacts=4)
    w.detect()
    with pytest.raises(RebootError) as e:
        w.contacts()
    assert e.value.code == 3 and w.stats()["contacts_dropped"] > 0
    w.close()


def test_world_transform_scale_translation_and_big_sphere(World, port):
    """W = scale (Q2: radius = W[0]*r), W with a translation (Q1: setPosition accumulates W.t), and a
    sphere much larger than the triangle columns (selection walk instead of the 4-way merge)."""
    from reboot_b200 import scenes
    tris, _ = scenes.terrain_triangles(24, 2.0, 2)
    w = World(); o = port.OrcWorld()
    w.add_static_geometry(tris); o.L.orc_add_geom(o.h, tris.shape[0], tris); o.n_geoms = 1
    bodies = [([[0, 0, 0, 0.25]], (3.0, 1.5, 3.0), np.diag([2, 2, 2, 1]).astype(np.float32)),
              ([[0, 0, 0, 0.5], [0.4, 0, 0, 0.3]], (-6.0, 1.8, 2.0), np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.25], [0, 0, 1, -0.5], [0, 0, 0, 1]], np.float32)),
              ([[0, 0, 0, 3.5]], (10.0, 4.6, -8.0), None),
              ([[0, 0, 0, 0.5]], (-12.0, 1.9, -12.0), None)]
    for obj, p, X in bodies:
        obj = np.array(obj, np.float32)
        w.add_model(obj, p, (0.5, -2, 0.25), 3, X)
        o.L.orc_add_body(o.h, obj.shape[0], obj, None if X is None else np.ascontiguousarray(X).ctypes.data_as(C.c_void_p),
                         np.array(p, np.float32), np.array([0.5, -2, 0.25], np.float32), 1, 1); o.n_bodies += 1
    w.build(); o.build(False)
    o.L.orc_set_prune_slack(o.h, 8.0)
    assert np.array_equal(bits(w.world_spheres()[:, :3][0]), bits(o.world_spheres(0, 1)[0, :3]))
    assert w.world_spheres()[0, 3] == 0.5
    for k in range(80):
        w.step(5); o.step(5, port.ORDER_CANONICAL)
    a, b = w.state(), o.state()
    for key in ("pos", "vel", "model_t", "prev_t"):
        assert np.array_equal(bits(a[key]), bits(b[key])), key
    assert np.array_equal(a["flags"], b["flags"]) and (a["flags"] & 4).any()
    w.close()


def test_full_size_c3_properties(World, port):
    """BASELINE.json config 3 at full size (1M spheres vs 2M-triangle terrain): size-independent
    properties -- nothing dropped, margin never exceeded, detection idempotent and state-neutral,
    sampled contacts re-verified with the oracle predicate, sampled spheres re-checked for completeness
    against every nearby terrain triangle, step_host round trip consistent."""
    from reboot_b200 import scenes
    sc = scenes.config3()
    assert sc.n_bodies == 1_000_000 and sc.n_tris == 2_000_000
    w = World(sc)
    for k in range(260):    # long enough for the lowest spheres to land
        w.step(5)
    st = w.stats()
    assert st["contacts_dropped"] == 0 and st["margin_overflow"] == 0 and st["hits_st"] > 1000
    w.integrate(5)
    s0 = w.state()
    w.detect(); c1 = w.contacts()
    w.detect(); c2 = w.contacts()
    assert len(c1) > 1000 and np.array_equal(c1, c2)
    s1 = w.state()
    assert np.array_equal(bits(s0["pos"]), bits(s1["pos"])) and np.array_equal(s0["flags"], s1["flags"])
    tris = sc.tris[0]
    rng = np.random.default_rng(0)
    stc = c1[c1["kind"] == 1]
    pick = stc[rng.choice(len(stc), 3000, replace=False)]
    centers = s0["pos"][pick["body"]]
    assert port.pred_sphere_triangle(centers, 0.5, tris[pick["other_prim"]]).all()
    ssc = c1[c1["kind"] == 0]
    if len(ssc):
        d = np.linalg.norm(s0["pos"][ssc["body"]].astype(np.float64) - s0["pos"][ssc["other"]], axis=1)
        assert (d <= 1.0 + 1e-5).all() and (ssc["body"] < ssc["other"]).all()
    # completeness on a sample: every terrain triangle within 2 cells of the sphere, tested on the CPU
    G, c = 1000, 2.0
    have = set(zip(stc["body"].tolist(), stc["other_prim"].tolist()))
    sample = rng.choice(sc.n_bodies, 1500, replace=False)
    cells_x = np.clip(np.floor((s0["pos"][sample, 0] + G * c / 2) / c).astype(int), 0, G - 1)
    cells_z = np.clip(np.floor((s0["pos"][sample, 2] + G * c / 2) / c).astype(int), 0, G - 1)
    cs, ts, owner = [], [], []
    for b, cx, cz in zip(sample, cells_x, cells_z):
        for dz in (-1, 0, 1):
            for dx in (-1, 0, 1):
                x, z = cx + dx, cz + dz
                if 0 <= x < G and 0 <= z < G:
                    for t in (2 * (z * G + x), 2 * (z * G + x) + 1):
                        cs.append(s0["pos"][b]); ts.append(t); owner.append(b)
    hit = port.pred_sphere_triangle(np.array(cs), 0.5, tris[np.array(ts)])
    expect = set((int(o), int(t)) for o, t, h in zip(owner, ts, hit) if h)
    sample_set = set(sample.tolist())
    got = set((b, t) for (b, t) in have if b in sample_set)
    assert expect == got
    # host round trip: positions returned by the per-tick host call equal the device state
    pos4 = np.zeros((sc.n_bodies, 4), np.float32)
    n = w.step_host(5, None, pos4)
    s2 = w.state()
    assert np.array_equal(bits(pos4[:, :3]), bits(s2["pos"])) and n == w.contact_count()
    w.close()


def test_phased_kernel_bit_identical_to_first_generation_kernel(World, monkeypatch):
    """The phased single-sphere collide kernel (default) and the first-generation kernel
    (RB_COLLIDE_V1=1) implement the same canonical order: identical state and contacts on a dense scene."""
    from reboot_b200 import scenes
    sc = scenes.falling_spheres(60000, 200, 2.0, 0.5, (0.3, 4.0), True, 21, "ab")
    monkeypatch.setenv("RB_COLLIDE_V1", "1")
    w1 = World(sc)
    monkeypatch.delenv("RB_COLLIDE_V1")
    monkeypatch.setenv("RB_VERLET", "0")          # the phased kernel enumerating every tick
    w2 = World(sc)
    monkeypatch.delenv("RB_VERLET")
    w3 = World(sc)                                # default: the same kernel stepping from persistent lists
    for k in range(200):
        w1.step(5); w2.step(5); w3.step(5)
    a, b, c = w1.state(), w2.state(), w3.state()
    for key in ("pos", "vel", "model_t", "prev_t"):
        assert np.array_equal(bits(a[key]), bits(b[key])), key
        assert np.array_equal(bits(a[key]), bits(c[key])), key
    assert np.array_equal(a["flags"], b["flags"]) and np.array_equal(a["flags"], c["flags"])
    assert np.array_equal(w1.contacts(), w2.contacts()) and np.array_equal(w1.contacts(), w3.contacts())
    s1, s2, s3 = w1.stats(), w2.stats(), w3.stats()
    assert s1["tests_st"] == s2["tests_st"] and s1["hits_st"] == s2["hits_st"] and s1["hits_ss"] == s2["hits_ss"] and s1["tests_ss"] == s2["tests_ss"]
    assert s3["hits_st"] == s2["hits_st"] and s3["hits_ss"] == s2["hits_ss"]      # lists are supersets: more tests, the same hits
    assert s2["hits_ss"] > 1000 and s2["hits_st"] > 100000
    assert s2["list_ticks"] == 0 and 0 < s3["list_rebuilds"] <= s3["list_ticks"] <= 200   # (the host may sit out rebuild-heavy phases on the per-tick kernel)
    w1.close(); w2.close(); w3.close()


def test_persistent_lists_survive_host_edits(World, monkeypatch):
    """Persistent candidate lists (default) against per-tick enumeration (RB_VERLET=0) over a run that settles
    (lists get reused for many ticks) and is disturbed from the host: teleports, velocity kicks, bodies
    switched off and on, per-tick forces, interleaved rb_detect / rb_integrate + rb_collide calls.
    Bit-identical state, flags and contact sets at every checkpoint."""
    from reboot_b200 import scenes
    sc = scenes.falling_spheres(20000, 100, 2.0, 0.5, (0.3, 2.0), True, 33, "vl")
    monkeypatch.setenv("RB_VERLET", "0")
    wa = World(sc)
    monkeypatch.delenv("RB_VERLET")
    wb = World(sc)
    rng = np.random.default_rng(5)
    nb = sc.n_bodies

    def same(tag):
        a, b = wa.state(), wb.state()
        for key in ("pos", "vel", "model_t", "prev_t"):
            assert np.array_equal(bits(a[key]), bits(b[key])), (tag, key)
        assert np.array_equal(a["flags"], b["flags"]), tag
        assert np.array_equal(wa.contacts(), wb.contacts()), tag

    for k in range(400):
        wa.step(5); wb.step(5)
        if k % 50 == 49: same(("settle", k))
    st = wb.stats()
    assert 100 < st["list_ticks"] <= 400 and st["list_rebuilds"] < st["list_ticks"] - 50, st       # lists were reused once the pile settled
    settled_rebuilds = st["list_rebuilds"]
    # teleport a block of bodies on top of others, kick some velocities
    ids = rng.choice(nb, 500, replace=False)
    s = wa.state()
    pos = np.zeros((nb, 4), np.float32); pos[:, :3] = s["pos"]; pos[:, 3] = 1
    vel = np.zeros((nb, 4), np.float32); vel[:, :3] = s["vel"]
    pos[ids[:250], :3] = pos[ids[250:], :3] + np.float32(0.3)
    vel[ids[250:], :3] += rng.normal(0, 3, (250, 3)).astype(np.float32)
    for w in (wa, wb):
        w.set_state(0, pos[:, :3], vel[:, :3])
    for k in range(60):
        wa.step(5); wb.step(5)
    same("teleport")
    # switch bodies off and on
    off = np.zeros(nb, np.uint32); on = np.full(nb, 1, np.uint32)
    for w in (wa, wb):
        w.set_flags(0, off[:3000], 1)
    for k in range(20):
        wa.step(5); wb.step(5)
    same("off")
    for w in (wa, wb):
        w.set_flags(0, on[:3000], 1)
    for k in range(20):
        wa.step(5); wb.step(5)
    same("on")
    # per-tick forces through the host call, and the stand-alone phases in between
    f = np.zeros((nb, 4), np.float32); f[:, 0] = rng.normal(0, 40, nb); f[:, 2] = rng.normal(0, 40, nb); f[:, 3] = 1
    p4 = np.zeros((nb, 4), np.float32)
    for k in range(40):
        wa.step_host(5, f, p4); wb.step_host(5, f, p4)
        if k == 10:
            wa.detect(); wb.detect()
            assert np.array_equal(wa.contacts(), wb.contacts())
        if k == 20:
            for w in (wa, wb):
                w.integrate(5); w.collide()
    same("forces")
    assert wb.stats()["list_rebuilds"] > settled_rebuilds
    assert wa.stats()["margin_overflow"] == wb.stats()["margin_overflow"]
    wa.close(); wb.close()


def test_pipelined_host_round_trip_matches_blocking(World):
    """rb_step_host_async (4-deep ring, copies overlapping the tick) delivers, three calls later, exactly what
    the blocking rb_step_host delivers: positions in id order and the contact count."""
    from reboot_b200 import scenes
    from reboot_b200._lib import ptr
    sc = scenes.falling_spheres(30000, 120, 2.0, 0.5, (0.3, 3.0), True, 8, "pipe")
    w1, w2, w3 = World(sc), World(sc), World(sc)
    nb = sc.n_bodies
    rng = np.random.default_rng(1)
    force = np.zeros((nb, 4), np.float32); force[:, 0] = rng.uniform(-3, 3, nb); force[:, 2] = rng.uniform(-3, 3, nb)
    force3 = np.ascontiguousarray(force[:, :3])
    p_sync = np.zeros((nb, 4), np.float32)
    D = 4                                                             # RB_HOST_PIPE_DEPTH: results arrive D - 1 calls later
    p_async = [np.zeros((nb, 4), np.float32) for _ in range(D)]
    p_async3 = [np.zeros((nb, 3), np.float32) for _ in range(D)]      # rb_step_host_async3: packed xyz both ways
    hist = []
    n2, n3 = C.c_uint64(), C.c_uint64()
    for k in range(70):
        if k == 40:       # the forces change mid-run (they are taken from the uploaded array of each tick)
            force[:, 0] = rng.uniform(-30, 30, nb); force3 = np.ascontiguousarray(force[:, :3])
        n1 = w1.step_host(5, force, p_sync)
        hist.append((p_sync.copy(), n1))
        assert w2.L.rb_step_host_async(w2.h, 5, ptr(force), ptr(p_async[k % D]), C.byref(n2)) == 0
        assert w3.L.rb_step_host_async3(w3.h, 5, ptr(force3), ptr(p_async3[k % D]), C.byref(n3)) == 0
        if k >= D - 1:
            j = k - (D - 1)
            assert np.array_equal(bits(p_async[j % D]), bits(hist[j][0])) and n2.value == hist[j][1], k
            assert np.array_equal(bits(p_async3[j % D]), bits(hist[j][0][:, :3].copy())) and n3.value == hist[j][1], k
    w2.sync(); w3.sync()
    assert np.array_equal(bits(p_async[69 % D]), bits(p_sync)) and np.array_equal(bits(p_async3[69 % D]), bits(p_sync[:, :3].copy()))
    assert np.array_equal(bits(w1.state()["pos"]), bits(w2.state()["pos"])) and hist[-1][1] > 1000
    assert np.array_equal(bits(w1.state()["pos"]), bits(w3.state()["pos"])) and np.array_equal(bits(w1.state()["vel"]), bits(w3.state()["vel"]))
    w1.close(); w2.close(); w3.close()


def test_model_matrix_export_matches_oracle(World, port):
    """Entity::getMVP / getPrevMVP model buffers (the step's only product the renderer consumes): full 4x4
    row-major matrices, incl. W = scale and W with a translation, before the first tick and after ticks."""
    from reboot_b200 import scenes
    tris, _ = scenes.terrain_triangles(24, 2.0, 2)
    w = World(); o = port.OrcWorld()
    w.add_static_geometry(tris); o.L.orc_add_geom(o.h, tris.shape[0], tris); o.n_geoms = 1
    rng = np.random.default_rng(3)
    for k in range(40):
        X = np.eye(4, dtype=np.float32)
        if k % 3 == 1:
            X[0, 0] = X[1, 1] = X[2, 2] = np.float32(1.5)
        if k % 3 == 2:
            X[0, 3], X[1, 3], X[2, 3] = np.float32(0.25), np.float32(-0.125), np.float32(0.5)
        obj = np.array([[0, 0, 0, 0.4]], np.float32)
        # bodies whose W carries a translation stay airborne here: every resolution teleports them by W.t
        # (quirk Q1), further than any candidate margin -- that regime is counted in margin_overflow instead
        p = np.array([rng.uniform(-20, 20), rng.uniform(1.5, 4.0) + (30.0 if k % 3 == 2 else 0.0), rng.uniform(-20, 20)], np.float32)
        w.add_model(obj, p, (0, 0, 0), 3, X)
        o.L.orc_add_body(o.h, 1, obj, np.ascontiguousarray(X).ctypes.data_as(C.c_void_p), p, np.zeros(3, np.float32), 1, 1); o.n_bodies += 1
    w.build(); o.build(False)
    for nsteps in (0, 1, 60):
        for _ in range(nsteps):
            w.step(5); o.step(5, port.ORDER_CANONICAL)
        a, b = w.model_matrices(), o.model_matrices()
        assert np.array_equal(bits(a[0].reshape(-1, 16)), bits(b[0].reshape(-1, 16))), nsteps
        assert np.array_equal(bits(a[1].reshape(-1, 16)), bits(b[1].reshape(-1, 16))), nsteps
    assert (w.state()["flags"] & 4).any() and w.stats()["margin_overflow"] == 0
    w.close()


def test_compound_kernels_bit_identical(World, monkeypatch):
    """Compound bodies: the per-body kernel (RB_COMPOUND=legacy), the group kernel (group: a group of lanes per body, speculative
    canonical-order resolution) and the split kernels (split, the default: enumeration per sphere into lists, resolution per body
    from them) give identical state, contacts and test counts -- uniform 8-sphere bodies piled closely enough to collide with
    each other, and ragged bodies (1..5 spheres) next to an inactive one."""
    from reboot_b200 import scenes
    kinds = ("legacy", "group", "split")
    sc = scenes.config4(6000, 64)
    ws = []
    for kind in kinds:
        monkeypatch.setenv("RB_COMPOUND", kind)
        ws.append(World(sc))
    monkeypatch.delenv("RB_COMPOUND")
    for k in range(250):
        for w in ws: w.step(5)
    a = ws[0].state(); s1 = ws[0].stats(); c1 = ws[0].contacts()
    for w, kind in zip(ws[1:], kinds[1:]):
        b = w.state()
        for key in ("pos", "vel", "model_t", "prev_t"):
            assert np.array_equal(bits(a[key]), bits(b[key])), (kind, key)
        assert np.array_equal(a["flags"], b["flags"]), kind
        assert np.array_equal(c1, w.contacts()), kind
        s2 = w.stats()
        for key in ("hits_st", "tests_ss", "hits_ss"):
            assert s1[key] == s2[key], (kind, key)
        # (the split kernels keep their triangle lists for several ticks: supersets of the per-tick candidates, so more tests, the same hits)
        # (... and pruned with the exact predicate when they are built: fewer tests than the box-pruned per-tick candidates)
        assert s2["tests_st"] == s1["tests_st"] if kind == "group" else s2["hits_st"] <= s2["tests_st"] <= 1.5 * s1["tests_st"], kind
    assert s1["hits_ss"] > 100 and s1["hits_st"] > 10000
    for w in ws: w.close()
    tris, _ = scenes.terrain_triangles(8, 2.0, 1)
    ws = []
    for kind in kinds:
        monkeypatch.setenv("RB_COMPOUND", kind)
        w = World(); w.add_static_geometry(tris)
        rng = np.random.default_rng(4)
        for ns, p, fl in [(1, (0.0, 1.2, 0.0), 3), (3, (2.0, 1.3, 1.0), 3), (5, (-3.0, 1.4, 2.0), 3), (1, (1.0, 1.1, -3.0), 2), (2, (2.3, 1.6, 1.2), 3), (1, (1000.4, 0.0, 0.0), 3)]:
            obj = np.zeros((ns, 4), np.float32); obj[:, :3] = rng.uniform(-0.3, 0.3, (ns, 3)); obj[:, 3] = 0.5
            w.add_model(obj, p, (0, -1, 0), fl)
        w.build(); ws.append(w)
    monkeypatch.delenv("RB_COMPOUND")
    for k in range(120):
        # host edits behind the persistent triangle lists' back: a teleport, a kick, a body switched off and on again, Entity::setPosition
        for w in ws:
            if k == 30: w.set_state(1, pos=np.array([[2.6, 1.25, 0.4]], np.float32))
            if k == 45: w.set_state(2, vel=np.array([[1.5, 0.5, -2.0]], np.float32))
            if k == 55: w.set_flags(0, np.array([0], np.uint32), 1)
            if k == 75: w.set_flags(0, np.array([1], np.uint32), 1)
            if k == 90: w.entity_set_position(4, (2.2, 1.4, 1.1))
        for w in ws: w.step(5)
        a = ws[0].state()
        for w, kind in zip(ws[1:], kinds[1:]):
            b = w.state()
            assert np.array_equal(bits(a["pos"]), bits(b["pos"])) and np.array_equal(bits(a["vel"]), bits(b["vel"])) and np.array_equal(a["flags"], b["flags"]), (kind, k)
    for w in ws: w.close()


def test_c4_subscene_10k_compound_bodies_vs_canonical_oracle(World, port):
    """BASELINE config 4 at 1/25 of its size: 10,240 bodies x 8 collision spheres, dense enough that bodies overlap each other
    and the terrain from the first tick (about 0.7M sphere-sphere and 0.1M sphere-triangle tests per tick) -- state, flags,
    model translations bit-identical to the canonical-order oracle after every tick, contact sets of a frozen detection too."""
    from reboot_b200 import scenes
    sc = scenes.config4(10240, 96, 31)
    sc.body_pos[:, 1] -= 1.6          # start close to the ground: contacts within the first ticks
    w = World(sc)
    o = port.OrcWorld(sc, octree=False)
    for k in range(8):
        w.step(5); o.step(5, port.ORDER_CANONICAL)
        a, b = w.state(), o.state()
        for key in ("pos", "vel", "model_t", "prev_t"):
            assert np.array_equal(bits(a[key]), bits(b[key])), (k, key)
        assert np.array_equal(a["flags"], b["flags"]), k
    st = w.stats()
    assert st["hits_ss"] > 10000 and st["hits_st"] > 1000
    w.integrate(5); o.integrate(5)
    w.detect()
    ss_g, st_g = w.contact_sets()
    ss_o, st_o = o.contact_sets(o.detect(port.ORDER_CANONICAL))
    assert np.array_equal(ss_g, ss_o) and np.array_equal(st_g, st_o)
    assert len(ss_g) > 1000 and len(st_g) > 100
    w.close()


def test_full_size_c4_properties(World, port):
    """BASELINE.json config 4 at full size (262,144 bodies x 8 collision spheres on the 2M-triangle terrain): nothing dropped,
    detection idempotent and state-neutral, sampled sphere-triangle and every sphere-sphere contact re-verified with the oracle
    predicates at the reported sphere centres, and the whole run bit-identical between the default (split, persistent triangle
    lists) and the per-body kernels (state checksum after 300 ticks)."""
    import os
    from reboot_b200 import scenes
    sc = scenes.config4()
    assert sc.n_bodies == 262_144 and sc.n_spheres == 8 * 262_144 and sc.n_tris == 2_000_000
    w = World(sc, max_contacts=8 * sc.n_spheres)
    for k in range(300):
        w.step(5)
    st = w.stats()
    assert st["contacts_dropped"] == 0 and st["hits_st"] > 1000 and st["hits_ss"] > 1000
    w.integrate(5)
    s0 = w.state()
    w.detect(); c1 = w.contacts()
    w.detect(); c2 = w.contacts()
    assert len(c1) > 1000 and np.array_equal(c1, c2)
    s1 = w.state()
    assert np.array_equal(bits(s0["pos"]), bits(s1["pos"])) and np.array_equal(s0["flags"], s1["flags"])
    # world sphere centre = object centre + model translation (identity W), radius 0.4
    obj = sc.obj_spheres.reshape(sc.n_bodies, 8, 4)
    centre = lambda body, sph: (obj[body, sph, :3] + s0["model_t"][body]).astype(np.float32)
    tris = sc.tris[0]
    rng = np.random.default_rng(0)
    stc = c1[c1["kind"] == 1]
    pick = stc[rng.choice(len(stc), min(3000, len(stc)), replace=False)]
    assert port.pred_sphere_triangle(centre(pick["body"], pick["sphere"]), 0.4, tris[pick["other_prim"]]).all()
    ssc = c1[c1["kind"] == 0]
    assert len(ssc) > 100 and (ssc["body"] < ssc["other"]).all()
    d = np.linalg.norm(centre(ssc["body"], ssc["sphere"]).astype(np.float64) - centre(ssc["other"], ssc["other_prim"]), axis=1)
    assert (d <= 0.8 + 1e-5).all()
    w.close()
    # the same 300 ticks with the per-body kernel: identical state
    os.environ["RB_COMPOUND"] = "legacy"
    try:
        w2 = World(sc, max_contacts=8 * sc.n_spheres)
    finally:
        del os.environ["RB_COMPOUND"]
    for k in range(300):
        w2.step(5)
    w2.integrate(5)
    s2 = w2.state()
    for key in ("pos", "vel", "model_t", "prev_t"):
        assert np.array_equal(bits(s0[key]), bits(s2[key])), key
    assert np.array_equal(s0["flags"], s2["flags"])
    w2.close()
