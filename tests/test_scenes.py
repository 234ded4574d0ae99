"""CPU suite: synthetic scene generators are deterministic and well formed."""
import numpy as np

from reboot_b200 import scenes


def test_terrain_matches_reference_triangulation_invariant():
    tris, h = scenes.terrain_triangles(16, 2.0, 3)
    assert tris.shape == (2 * 16 * 16, 9) and tris.dtype == np.float32
    a, b, c = tris[:, 0:3], tris[:, 3:6], tris[:, 6:9]
    # procedural/src/ProcIsland.cpp:76-82: normal = (v0 - v1) x (v0 - v2) has y > 0
    n = np.cross(a - b, a - c)
    assert (n[:, 1] > 0).all()
    assert np.abs(tris[:, [0, 3, 6]]).max() <= 16.0 and h.max() <= 0.75 * 2.0 + 1e-6


def test_scenes_are_deterministic_and_in_world():
    a, b = scenes.config1(), scenes.config1()
    assert np.array_equal(a.body_pos, b.body_pos) and np.array_equal(a.tris[0], b.tris[0])
    assert a.n_bodies == 1000 and a.n_tris == 8192
    c2 = scenes.config2(1000)
    assert (c2.body_gravity == 0).all() and np.linalg.norm(c2.body_vel, axis=1).max() <= 5.0 + 1e-4
    c4 = scenes.config4(64, 16)
    assert c4.n_spheres == 64 * 8 and c4.sphere_start[-1] == 512
    c5 = scenes.config5(1000, 64, x_range=(-100, 0))
    assert len(c5.tris) == 2 and c5.tris[1].shape[0] == 12 * 100
    assert c5.body_pos[:, 0].min() >= -100 and c5.body_pos[:, 0].max() <= 0
    for s in (a, c2, c4, c5):
        assert np.abs(s.body_pos).max() < scenes.WORLD_HALF
