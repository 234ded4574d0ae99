"""GPU suite: the multi-rank slab path (pack / exchange / migrate / ghost kernels, global-id canonical
order) emulated on ONE GPU by in-process slab worlds stepped in lockstep (rb_step_group; device-to-
device copies stand in for the NCCL send/recv). Result must be bit-identical to a single world."""
import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from reboot_b200 import World, dist, scenes
    return World, dist, scenes


def _crossing_scene(scenes, n=4000, seed=3):
    """Spheres on a wide terrain with lateral velocities so that bodies cross slab faces and pile up near them."""
    sc = scenes.falling_spheres(n, 400, 2.0, 0.5, (0.5, 3.0), True, seed, "cross")
    rng = np.random.default_rng(seed)
    sc.body_vel[:, 0] = rng.uniform(-12, 12, n).astype(np.float32)
    sc.body_vel[:, 2] = rng.uniform(-3, 3, n).astype(np.float32)
    # concentrate half of them around the slab faces of a 2- and 4-way split (x = 0, +-500... scaled: +-200)
    k = n // 2
    sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0], k) + rng.uniform(-3, 3, k)).astype(np.float32)
    return sc


@pytest.mark.parametrize("nslabs", [2, 4])
def test_slab_group_bit_identical_to_single_world(mods, nslabs):
    World, dist, scenes = mods
    sc = _crossing_scene(scenes)
    single = World(sc)
    group = dist.SlabGroup(sc, nslabs)
    # slab faces sit at multiples of 2000/nslabs; the scene straddles x = 0 (and +-500 only for 4 slabs is
    # outside the 800-wide terrain, so the x = 0 face carries the traffic)
    migrated = 0
    for k in range(150):
        single.step(5)
        group.step(5)
        if k % 15 == 14 or k < 3:
            a = single.state(); b = group.state(sc.n_bodies)
            assert (b["owner"] >= 0).all(), "a body was lost in migration"
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(b[key])), (k, key)
            assert np.array_equal(a["flags"], b["flags"]), k
            sa, sb = single.contact_sets(), group.contact_sets()
            assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]), k
        migrated += sum(w.stats()["migrated_out_last"] for w in group.worlds)
    assert migrated > 0, "the scene must exercise migration"
    assert sum(w.stats()["halo_sent_last"] for w in group.worlds) > 0
    # ownership follows position (list ticks keep a body that has only just crossed a face with its owner until the
    # next scheduled rebuild: up to margin_cap = half the largest radius beyond the face)
    b = group.state(sc.n_bodies)
    own = dist.owner_of(b["pos"][:, 0], nslabs)
    off = np.nonzero(b["owner"] != own)[0]
    faces = -1000.0 + 2000.0 * np.arange(1, nslabs) / nslabs
    assert all(np.abs(b["pos"][i, 0] - faces).min() <= 0.25 + 1e-4 and abs(int(b["owner"][i]) - int(own[i])) == 1 for i in off), off[:10]
    # frozen detection through the group path
    single.integrate(5)
    for w in group.worlds:
        w.integrate(5)
    single.detect(); group.detect()
    sa, sb = single.contact_sets(), group.contact_sets()
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    single.close(); group.close()


@pytest.mark.parametrize("nslabs", [2, 4])
def test_slab_group_steps_from_persistent_lists(mods, nslabs):
    """Slab worlds stepping from persistent candidate lists: a scene that comes to rest around the slab faces
    (so the lists get reused for many ticks, the neighbours' ghosts are looked up through the hot-body table,
    and a few bodies still migrate), against a single world -- bit-identical at every checkpoint."""
    World, dist, scenes = mods
    n = 6000
    sc = scenes.falling_spheres(n, 400, 2.0, 0.5, (0.3, 1.5), True, 11, "rest")
    rng = np.random.default_rng(11)
    sc.body_vel[:, 0] = rng.uniform(-1.5, 1.5, n).astype(np.float32)
    k = (3 * n) // 4
    faces = [0.0] if nslabs == 2 else [-500.0, 0.0, 500.0]
    sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0] if nslabs == 4 else faces, k) + rng.uniform(-40, 40, k)).astype(np.float32)
    single = World(sc)
    group = dist.SlabGroup(sc, nslabs)
    migrated = 0
    for t in range(500):
        single.step(5)
        group.step(5)
        if t % 25 == 24:
            a = single.state(); b = group.state(sc.n_bodies)
            assert (b["owner"] >= 0).all(), "a body was lost in migration"
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(b[key])), (t, key)
            assert np.array_equal(a["flags"], b["flags"]), t
            sa, sb = single.contact_sets(), group.contact_sets()
            assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]), t
            migrated += sum(w.stats()["migrated_out_last"] for w in group.worlds)
    st = [w.stats() for w in group.worlds]
    assert sum(s["halo_recv_last"] for s in st) > 0, "ghosts must be in play"
    busy = [s for s in st if s["n_bodies"] > 100]
    ss = single.stats()
    info = [(s["n_bodies"], s["list_ticks"], s["list_rebuilds"], s["list_hot_last"], s["list_rebuild_reason"]) for s in st + [ss]]
    assert busy and all(s["list_ticks"] > 200 and s["list_rebuilds"] < s["list_ticks"] // 2 for s in busy), info
    assert len(sa[0]) > 50, "sphere-sphere contacts (some across a face) must be in play"
    single.close(); group.close()


def test_mismatched_halo_layouts_are_refused(mods):
    """The neighbours store straight into a rank's halo buffers, so their layout (rb_config.halo_ghost_cap) must agree:
    attaching peers whose capacities differ fails loudly instead of corrupting memory."""
    World, dist, scenes = mods
    sc = _crossing_scene(scenes, n=600)
    parts = [dist.partition_scene(sc, 2, r) for r in range(2)]
    ws = []
    for r, (sub, ids, slab, rmax) in enumerate(parts):
        ws.append(World(sub, slab=slab, rank=r, nranks=2, ids=ids, rm_max_hint=rmax, halo_ghost_cap=8192 + 1024 * r))
    rc = ws[0].L.rb_dist_attach_peers(ws[0].h, None, ws[1].h)
    assert rc != 0 and b"halo buffer layouts differ" in ws[0].L.rb_last_error(ws[0].h)
    for w in ws:
        w.close()
    # and the helper picks one layout for uneven partitions (the bug this guards against: per-rank auto values)
    L = ws[0].L
    assert L.rb_halo_ghost_cap_auto(500_000, 0.5, 0.0, 1000.0) != L.rb_halo_ghost_cap_auto(505_000, 0.5, 0.0, 1000.0)


@pytest.mark.parametrize("nslabs", [2, 4])
def test_slab_lists_with_migration_every_few_ticks(mods, nslabs, monkeypatch):
    """Slab worlds kept on the persistent lists (RB_VL_NEVER_OFF=1) through a violent phase: bodies cross the faces every
    few ticks, so list ticks remove emigrants (swap-remove: the last row moves into the hole AFTER K1 linked the hot rows
    into the cell table by row), append immigrants and rebuild on the device's request. No host read in between -- the
    host's own decisions run on stale reports, like in a long bench run. Regression: a moved hot row missed its tick."""
    World, dist, scenes = mods
    sc = _crossing_scene(scenes)
    sc.body_vel[:, 0] *= np.float32(1.5)
    monkeypatch.setenv("RB_VL_NEVER_OFF", "1")
    single = World(sc)
    group = dist.SlabGroup(sc, nslabs)
    monkeypatch.delenv("RB_VL_NEVER_OFF")
    for k in range(140):
        single.step(5)
        group.step(5)
    a = single.state(); b = group.state(sc.n_bodies)
    assert (b["owner"] >= 0).all(), "a body was lost in migration"
    for key in ("pos", "vel", "model_t", "prev_t"):
        assert np.array_equal(bits(a[key]), bits(b[key])), key
    assert np.array_equal(a["flags"], b["flags"])
    sa, sb = single.contact_sets(), group.contact_sets()
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    st = [w.stats() for w in group.worlds]
    busy = [s for s in st if s["n_bodies"] > 100]
    assert busy and all(s["list_ticks"] >= 139 for s in busy), [(s["n_bodies"], s["list_ticks"]) for s in st]
    assert sum(s["list_rebuilds_by_reason"][0] + s["list_rebuilds_by_reason"][3] for s in st) > 10, "migrations must have forced rebuilds"
    single.close(); group.close()


def test_slab_worlds_export_model_matrices(mods):
    """Entity::getMVP / getPrevMVP model matrices from slab worlds (rows a rank owns, aligned with its body ids) equal the
    single world's, body by body, after bodies have migrated."""
    World, dist, scenes = mods
    sc = _crossing_scene(scenes)
    single = World(sc)
    group = dist.SlabGroup(sc, 2)
    for k in range(90):
        single.step(5); group.step(5)
    cur, prev = single.model_matrices()
    seen = np.zeros(sc.n_bodies, np.int32)
    for w in group.worlds:
        ids = w.body_ids()
        c, p = w.model_matrices()
        assert np.array_equal(bits(c), bits(cur[ids])) and np.array_equal(bits(p), bits(prev[ids]))
        seen[ids] += 1
    assert (seen == 1).all()
    single.close(); group.close()


def _compound_crossing_scene(scenes, n=3000, seed=5, ragged=False):
    """Compound bodies (8 spheres on a 2x2x2 lattice, or 1..8 when ragged) with lateral velocities, half of them started around
    the slab faces so that ghosts of whole bodies and migrations carry the traffic."""
    sc = scenes.config4(n, 400, seed)
    rng = np.random.default_rng(seed)
    sc.body_vel[:, 0] = rng.uniform(-12, 12, n).astype(np.float32)
    sc.body_vel[:, 2] = rng.uniform(-3, 3, n).astype(np.float32)
    k = n // 2
    sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0], k) + rng.uniform(-4, 4, k)).astype(np.float32)
    sc.body_pos[:, 1] -= 1.0
    if ragged:
        cnt = rng.integers(1, 9, n)
        keep = np.concatenate([np.arange(8 * b, 8 * b + cnt[b]) for b in range(n)])
        sc.obj_spheres = np.ascontiguousarray(sc.obj_spheres[keep])
        sc.sphere_start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    return sc


@pytest.mark.parametrize("nslabs,ragged", [(2, False), (4, True)])
def test_compound_slab_group_bit_identical_to_single_world(mods, nslabs, ragged):
    """BASELINE config 4 on several GPUs: compound bodies in x-slab worlds (fixed sphere stride per body row, ghost / migrant
    records of whole bodies) against one world -- state, flags, model translations and contact sets bit-identical while
    bodies collide with each other across the faces and migrate."""
    World, dist, scenes = mods
    sc = _compound_crossing_scene(scenes, ragged=ragged)
    single = World(sc)
    group = dist.SlabGroup(sc, nslabs)
    migrated = 0
    for k in range(160):
        single.step(5)
        group.step(5)
        if k % 16 == 15 or k < 3:
            a = single.state(); b = group.state(sc.n_bodies)
            assert (b["owner"] >= 0).all(), "a body was lost in migration"
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(b[key])), (k, key)
            assert np.array_equal(a["flags"], b["flags"]), k
            sa, sb = single.contact_sets(), group.contact_sets()
            assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]), k
        migrated += sum(w.stats()["migrated_out_last"] for w in group.worlds)
    assert migrated > 0, "the scene must exercise migration"
    assert sum(w.stats()["halo_sent_last"] for w in group.worlds) > 0
    st = single.stats()
    assert st["hits_ss"] > 0 and st["hits_st"] > 0
    b = group.state(sc.n_bodies)
    assert np.array_equal(b["owner"], dist.owner_of(b["model_t"][:, 0], nslabs)), "ownership follows the model translation"
    single.close(); group.close()


def test_slab_group_forces_torques_angular_travel_with_migrants(mods):
    """rb_config.slab_forces: bodies of slab worlds take StateVector::setForce / setTorque input and carry angular state; the
    arrays travel with migrating bodies (migrant records) and with rows moved by the swap-remove. Forces are re-applied every tick
    by GLOBAL body id (gated by the contact / gravity flags like the reference), torques once -- state, angular state, flags and
    contact sets bit-identical to a single world."""
    World, dist, scenes = mods
    sc = _crossing_scene(scenes, n=3000, seed=8)
    rng = np.random.default_rng(8)
    F = np.zeros((sc.n_bodies, 3), np.float32); F[:, 0] = rng.uniform(-30, 30, sc.n_bodies); F[:, 2] = rng.uniform(-10, 10, sc.n_bodies)
    Tq = rng.uniform(-2, 2, (sc.n_bodies, 3)).astype(np.float32)
    Av = rng.uniform(-1, 1, (sc.n_bodies, 3)).astype(np.float32)
    single = World(sc)
    single.set_torque(0, Tq); single.set_angular(0, avel=Av)
    group = dist.SlabGroup(sc, 2, slab_forces=True)
    for w in group.worlds:
        ids = w.body_ids()
        w.set_torque(0, Tq[ids]); w.set_angular(0, avel=Av[ids])
    migrated = 0
    for k in range(120):
        single.set_force(0, F)
        for w in group.worlds:
            w.set_force(0, F[w.body_ids()])
        single.step(5); group.step(5)
        migrated += sum(w.stats()["migrated_out_last"] for w in group.worlds)
        if k % 20 == 19 or k < 2:
            a = single.state(); b = group.state(sc.n_bodies)
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(b[key])), (k, key)
            assert np.array_equal(a["flags"], b["flags"]), k
            ap, av = single.angular()
            gp, gv = np.zeros_like(ap), np.zeros_like(av)
            for w in group.worlds:
                ids = w.body_ids(); p_, v_ = w.angular()
                gp[ids] = p_[: len(ids)]; gv[ids] = v_[: len(ids)]
            assert np.array_equal(bits(ap), bits(gp)) and np.array_equal(bits(av), bits(gv)), k
            sa, sb = single.contact_sets(), group.contact_sets()
            assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]), k
    assert migrated > 0, "the scene must exercise migration"
    # without the flag the setters are refused (migrants would lose the state)
    plain = dist.SlabGroup(sc, 2)
    from reboot_b200 import RebootError
    with pytest.raises(RebootError):
        plain.worlds[0].set_force(0, F[plain.worlds[0].body_ids()])
    plain.close(); single.close(); group.close()
