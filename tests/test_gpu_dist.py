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
    # ownership follows position
    b = group.state(sc.n_bodies)
    assert np.array_equal(b["owner"], dist.owner_of(b["pos"][:, 0], nslabs))
    # frozen detection through the group path
    single.integrate(5)
    for w in group.worlds:
        w.integrate(5)
    single.detect(); group.detect()
    sa, sb = single.contact_sets(), group.contact_sets()
    assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    single.close(); group.close()
