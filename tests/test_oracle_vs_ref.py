"""CPU suite (skipped where /root/reference was not available to build oracle/_ref): the plain-C
oracle against the LIVE compiled reference on fresh random inputs, beyond the committed fixtures."""
import numpy as np

from reboot_b200 import scenes
from util import bits


def test_random_predicates_live(ref, port):
    rng = np.random.default_rng(2024)
    N = 3000
    tri = rng.uniform(-2, 2, (N, 9)).astype(np.float32)
    c = rng.uniform(-2.5, 2.5, (N, 3)).astype(np.float32)
    r = rng.uniform(0.05, 1.5, N).astype(np.float32)
    assert np.array_equal(ref.pred_sphere_triangle(c, r, tri), port.pred_sphere_triangle(c, r, tri))
    assert np.array_equal(bits(ref.closest_point(c, tri)), bits(port.closest_point(c, tri)))
    a = rng.uniform(-1, 1, (N, 4)).astype(np.float32); b = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    a[:, 3] = np.abs(a[:, 3]); b[:, 3] = np.abs(b[:, 3])
    assert np.array_equal(ref.pred_sphere_sphere(a, b), port.pred_sphere_sphere(a, b))
    cube = np.concatenate([rng.uniform(-1, 1, (N, 3)), rng.uniform(0.2, 2, (N, 3))], 1).astype(np.float32)
    assert np.array_equal(ref.pred_sphere_cube(a, cube), port.pred_sphere_cube(a, cube))
    assert np.array_equal(ref.pred_triangle_cube(tri, cube), port.pred_triangle_cube(tri, cube))
    io = rng.uniform(-5, 5, (500, 18)).astype(np.float32); fl = rng.integers(0, 8, 500).astype(np.int32)
    assert np.array_equal(bits(ref.state_update(io, fl, 5, 5)), bits(port.state_update(io, fl, 5, 5)))


def test_scene_trajectory_live(ref, port):
    """Scenes not in the fixtures: same octree, same narrowphase test counts, same state for 150
    ticks. The dense scene has multi-partner sphere-sphere piles whose outcome follows the reference's
    pointer-hash entity order (not restated), so a few bodies may differ there."""
    for n, G, strict in ((120, 48, True), (200, 24, False)):
        sc = scenes.falling_spheres(n, G, 2.0, 0.5, (1.0, 5.0), True, 77, "live")
        w = ref.RefWorld(sc); o = port.OrcWorld(sc)
        assert len(w.leaves()[0]) == o.leaf_count()
        for k in range(150):
            w.step(5); o.step(5, port.ORDER_REFERENCE)
        s1, s2 = w.state(), o.state()
        far = np.abs(s1["pos"] - s2["pos"]).max(1) > 1e-5
        c1, c2 = w.counters(), o.counters()
        if strict:
            assert not far.any() and np.array_equal(s1["flags"], s2["flags"])
            assert np.array_equal(bits(s1["pos"]), bits(s2["pos"]))
            assert c1["tests_ss"] == c2["tests_ss"] and c1["tests_st"] == c2["tests_st"]
        else:
            assert far.mean() < 0.05


def test_world_transform_scale_and_visualize_channel(ref, port):
    """W = uniform scale: world radius = W[0] * r (Q2). The reference's own query channel
    (Physics::visualize payload) equals the sphere-triangle part of the hit log."""
    sc = scenes.falling_spheres(60, 16, 2.0, 0.25, (0.2, 1.0), True, 5, "scaled")
    sc.body_scale[:] = 2.0
    w = ref.RefWorld(sc); o = port.OrcWorld(sc)
    for k in range(60):
        w.step(5); o.step(5, port.ORDER_REFERENCE)
    assert np.allclose(w.world_spheres(3, 1), o.world_spheres(3, 1)) and abs(o.world_spheres(3, 1)[0, 3] - 0.5) < 1e-7
    a, _ = w.hits()
    viz = w.visualize()
    st = a[a[:, 0] == 1]
    assert set(map(tuple, viz)) >= set((int(r[1]), int(r[3]), int(r[4])) for r in st)
    s1, s2 = w.state(), o.state()
    assert np.allclose(s1["pos"], s2["pos"], rtol=1e-5, atol=1e-5)
    m1, m2 = w.model_matrices(), o.model_matrices()      # Entity::getMVP / getPrevMVP model buffers
    assert np.array_equal(bits(m1[0].reshape(-1, 16)), bits(m2[0].reshape(-1, 16))) and np.array_equal(bits(m1[1].reshape(-1, 16)), bits(m2[1].reshape(-1, 16)))
    assert m1[0][0, 0, 0] == 2.0 and m1[0][0, 3, 3] == 1.0
