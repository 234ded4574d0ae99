"""GPU suite, SURVEY 8(f) ranks 2 and 4: the frustum-culling kernel against the oracle (flags bit-exact, front-to-back
order) on reference fixtures, on the C1 octree leaves and on the world's own static cells at full C3 size; a world fed
through the collider-ingest path against one fed the oracle's primitives."""
import os

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras.npz")


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import port_binding
    from reboot_b200 import World, ingest, scenes
    return World, ingest, scenes, port_binding


def _expected_order(flags, keys):
    idx = np.nonzero(flags)[0]
    return idx[np.lexsort((idx, keys[idx]))]          # by |centre|, equal distances in index order


def test_cull_cubes_matches_oracle_and_reference_fixtures(mods):
    World, ingest, scenes, port = mods
    gold = np.load(GOLD)
    w = World(scenes.config1())
    for k in range(gold["fr_inv"].shape[0]):
        pl = ingest.frustum_planes(gold["fr_inv"][k])
        mn, mx = gold["fr_mins"][k], gold["fr_maxs"][k]
        # cubes whose (centre -+ extent / 2) round back to the fixture's boxes exactly: power-of-two friendly values
        c = np.round((mn + mx) / 2); e = np.round(mx - mn) + 2
        cubes = np.concatenate([c, e], 1).astype(np.float32)
        bmn, bmx, keys = port.cube_boxes(cubes)
        want = port.frustum_aabb(pl, bmn, bmx)
        got = ingest.cull_cubes(w, pl, cubes)
        assert np.array_equal(got, _expected_order(want, keys)), k
    # the C1 octree leaves through the drop-in form of OSP::getVisibleFrustumCulling
    leaves = gold["c1_leaf_cubes"]; ntri = gold["c1_leaf_ntri"]
    cubes = np.concatenate([leaves[:, :3], np.repeat(leaves[:, 3:4], 3, 1)], 1).astype(np.float32)
    bearing = np.nonzero(ntri > 0)[0]
    keys1 = np.sqrt((leaves[:, :3].astype(np.float32) ** 2).sum(1, dtype=np.float32))
    for k in range(gold["c1_cull_inv"].shape[0]):
        pl = ingest.frustum_planes(gold["c1_cull_inv"][k])
        got = bearing[ingest.cull_cubes(w, pl, cubes[bearing])] + 1            # 1-based leaf offsets like the reference
        want = gold["c1_cull_visible"][k]; want = want[want >= 0]
        assert set(got.tolist()) == set(want.tolist()), k
        assert np.array_equal(keys1[got - 1], keys1[want - 1]), k             # same front-to-back distances
    w.close()


def test_cull_static_cells_full_size(mods):
    """The world's own static structure at C3 size (1M cells): visible cells = oracle predicate on the cell boxes."""
    World, ingest, scenes, port = mods
    import cams
    sc = scenes.config3()
    w = World(sc)
    boxes, ntri = ingest.static_cells(w)
    st = w.stats()
    assert boxes.shape[0] == st["tri_cols_x"] * st["tri_cols_z"] and int(ntri.sum()) >= sc.n_tris
    rng = np.random.default_rng(4)
    total = 0
    for k in range(4):
        V = cams.view_matrix(rng.uniform(-600, 600, 3) * [1, 0.05, 1] + [0, 30, 0], rng.uniform(0, 6.28), rng.uniform(-0.5, 0.0))
        P = cams.projection(70, 1.6, 0.1, rng.uniform(100, 900))
        inv = (np.linalg.inv(V.astype(np.float64)) @ np.linalg.inv(P.astype(np.float64))).astype(np.float32)    # view^-1 * projection^-1
        got = ingest.cull_static(w, inv)
        pl = port.frustum_planes(inv)
        live = ntri > 0
        flags = port.frustum_aabb(pl, boxes[:, :3], boxes[:, 3:]) * live
        ctr = ((boxes[:, :3] + boxes[:, 3:]) / np.float32(2.0)).astype(np.float32)
        keys = np.sqrt(((ctr[:, 0] * ctr[:, 0] + ctr[:, 1] * ctr[:, 1]).astype(np.float32) + (ctr[:, 2] * ctr[:, 2]).astype(np.float32)).astype(np.float32))
        assert np.array_equal(got, _expected_order(flags, keys)), k
        total += len(got)
    assert 1000 < total < 4 * boxes.shape[0]
    w.close()


def test_world_built_through_ingest_equals_world_built_from_oracle_primitives(mods):
    World, ingest, scenes, port = mods
    rng = np.random.default_rng(9)
    # a bumpy floor as an indexed mesh under a node transform, plus bodies whose spheres come from collider meshes
    G = 24
    gx, gz = np.meshgrid(np.arange(G + 1), np.arange(G + 1), indexing="xy")
    xyz = np.stack([gx.ravel() - G / 2, 0.3 * np.sin(gx.ravel() * 0.7) * np.cos(gz.ravel() * 0.5), gz.ravel() - G / 2], 1).astype(np.float32)
    idx = []
    for z in range(G):
        for x in range(G):
            b = z * (G + 1) + x
            idx += [b + G + 1, b + 1, b, b + G + 1, b + G + 2, b + 1]
    idx = np.asarray(idx, np.uint32)
    trs = np.array([1.5, -0.25, 2.0, 0, 30, 0, 2, 1, 2], np.float32)
    wa, wb = World(), World()
    import ctypes as C
    from reboot_b200._lib import ptr
    g = C.c_uint32()
    wa._ck(wa.L.rb_ingest_static_mesh(wa.h, xyz.shape[0], ptr(xyz), idx.shape[0], ptr(idx), ptr(trs), C.byref(g)))
    wb.add_static_geometry(port.ingest_triangles(xyz, idx.astype(np.int32), trs))
    for k in range(300):
        nv = 20
        mesh = (rng.normal(0, 0.3, (nv, 3)) + rng.uniform(-0.2, 0.2, 3)).astype(np.float32)
        midx = rng.integers(0, nv, 60).astype(np.uint32)
        mtrs = np.array([0, 0, 0, rng.uniform(0, 360), rng.uniform(0, 360), 0, 1, 1, 1], np.float32)
        pos = np.array([rng.uniform(-15, 15), rng.uniform(2, 6), rng.uniform(-15, 15)], np.float32)
        sa = ingest.mesh_sphere(mesh, midx, mtrs)
        sb = port.ingest_sphere(mesh, midx.astype(np.int32), mtrs)
        assert np.array_equal(bits(sa), bits(sb))
        wa.add_model(sa.reshape(1, 4), pos, (0, 0, 0), 3)
        wb.add_model(sb.reshape(1, 4), pos, (0, 0, 0), 3)
    wa.build(); wb.build()
    for k in range(250):
        wa.step(5); wb.step(5)
    a, b = wa.state(), wb.state()
    for key in ("pos", "vel", "model_t"):
        assert np.array_equal(bits(a[key]), bits(b[key])), key
    assert np.array_equal(wa.contacts(), wb.contacts()) and wa.stats()["hits_st"] > 1000
    wa.close(); wb.close()


def test_triangle_cube_predicate_on_the_device_matches_reference_fixture(golden):
    """GeometryMath::triangleCubeDetection as the device evaluates it (rb_triangle_cube_test; the same device function
    registers static triangles in the grid columns at rb_build) against the fixture generated from the reference's own
    compiled predicate, plus random pairs against the oracle restatement."""
    import ctypes as C
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from reboot_b200 import World
    from reboot_b200._lib import ptr
    from oracle import port_binding as pb
    m = golden["micro"]
    w = World()
    tri = np.ascontiguousarray(m["tc_tri"], np.float32).reshape(-1, 9); cube = np.ascontiguousarray(m["tc_cube"], np.float32).reshape(-1, 6)
    out = np.zeros(tri.shape[0], np.uint32)
    assert w.L.rb_triangle_cube_test(w.h, tri.shape[0], ptr(tri), ptr(cube), ptr(out)) == 0
    assert np.array_equal(out.astype(np.int32), np.asarray(m["tc_hit"]).astype(np.int32)) and 0 < out.sum() < len(out)
    rng = np.random.default_rng(12)
    n = 200000
    tri = rng.uniform(-3, 3, (n, 9)).astype(np.float32)
    cube = np.concatenate([rng.uniform(-2, 2, (n, 3)), rng.uniform(0.05, 4, (n, 3))], axis=1).astype(np.float32)
    tri[: n // 4] *= np.float32(0.2)                      # small triangles: inside, outside and grazing boxes
    tri[n // 4: n // 2, 1::3] = tri[n // 4: n // 2, 1:2]   # axis-aligned planes (y constant): the degenerate cross-product axes
    out = np.zeros(n, np.uint32)
    assert w.L.rb_triangle_cube_test(w.h, n, ptr(tri), ptr(cube), ptr(out)) == 0
    ref = pb.pred_triangle_cube(tri, cube)
    assert np.array_equal(out.astype(np.int32), np.asarray(ref).astype(np.int32)) and 0.05 < out.mean() < 0.95
    w.close()
