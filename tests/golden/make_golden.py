"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_physics.so = the reference's
unmodified physics/ + math/ sources compiled headless, see oracle/Makefile). Run in the build container
(needs /root/reference to have built oracle/_ref); the fixtures are committed so the GPU box, which has
no /root/reference, can check the oracle port and the CUDA path against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_binding as rb  # noqa: E402
from reboot_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def micro_sphere_triangle():
    """Sphere vs triangle in every Voronoi region (face, 3 edges, 3 vertices) at penetrating /
    touching (Q5: touching is a MISS) / separated distances, plus degenerate and random cases."""
    rng = np.random.default_rng(101)
    tris, cs, rs = [], [], []
    base = np.array([[0, 0, 0], [2, 0, 0], [0, 0, 2]], dtype=np.float64)
    for rep in range(40):
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0] if rep else np.eye(3)
        off = rng.uniform(-50, 50, 3) if rep else np.zeros(3)
        T = (base * rng.uniform(0.5, 3.0)) @ R.T + off
        A, B, C = T
        n = np.cross(B - A, C - A)
        n /= np.linalg.norm(n)
        feats = {
            "face": (A + B + C) / 3, "vA": A, "vB": B, "vC": C,
            "eAB": (A + B) / 2, "eBC": (B + C) / 2, "eCA": (C + A) / 2,
        }
        cen = (A + B + C) / 3
        for name, q in feats.items():
            d = n if name == "face" else (q - cen) / np.linalg.norm(q - cen) * 0.8 + n * 0.6
            d /= np.linalg.norm(d)
            r = rng.uniform(0.2, 1.0)
            for mult in (0.25, 0.9, 0.999, 1.0, 1.001, 1.1, 2.0):
                tris.append(T.reshape(-1)); cs.append(q + d * r * mult); rs.append(r)
            tris.append(T.reshape(-1)); cs.append(q - d * r * 0.5); rs.append(r)   # below the plane
    # axis-aligned exact-touching cases (all values exactly representable)
    T0 = np.array([0, 0, 0, 4, 0, 0, 0, 0, 4], dtype=np.float64)
    for c, r in (((1, 1, 1), 1.0), ((1, 0.5, 1), 0.5), ((1, 0.5, 1), 0.5000001), ((-1, 0, 0), 1.0), ((-1, 0, 0), 1.0000001),
                 ((2, 0, -1), 1.0), ((2, 0, -0.5), 1.0), ((5, 0, 0), 1.0), ((0, 0, 0), 0.5), ((1, 0, 1), 0.5)):
        tris.append(T0); cs.append(np.array(c, dtype=np.float64)); rs.append(r)
    # random bulk
    N = 6000
    tr = rng.uniform(-3, 3, (N, 9)); cc = rng.uniform(-3.5, 3.5, (N, 3)); rr = rng.uniform(0.05, 2.0, N)
    t = np.concatenate([np.array(tris), tr]).astype(f32)
    c = np.concatenate([np.array(cs), cc]).astype(f32)
    r = np.concatenate([np.array(rs), rr]).astype(f32)
    # degenerate triangles (zero area) and sphere centre exactly on the triangle
    t = np.concatenate([t, np.array([[0, 0, 0, 1, 0, 0, 2, 0, 0], [1, 1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 0, 2, 0, 0, 0, 0, 2]], dtype=f32)])
    c = np.concatenate([c, np.array([[0.5, 0.1, 0], [1, 1, 1.2], [0.5, 0, 0.5]], dtype=f32)])
    r = np.concatenate([r, np.array([0.5, 0.5, 0.5], dtype=f32)])
    hit = rb.pred_sphere_triangle(c, r, t)
    q = rb.closest_point(c, t)
    return dict(st_tri=t, st_c=c, st_r=r, st_hit=hit, st_closest=q)


def micro_sphere_sphere():
    rng = np.random.default_rng(102)
    a = rng.uniform(-5, 5, (4000, 4)); b = a + rng.normal(0, 0.7, (4000, 4))
    a[:, 3] = rng.uniform(0.1, 1, 4000); b[:, 3] = rng.uniform(0.1, 1, 4000)
    ex = np.array([[0, 0, 0, 0.5, 1, 0, 0, 0.5], [0, 0, 0, 0.5, 1.0000001, 0, 0, 0.5], [0, 0, 0, 0.5, 0.99999, 0, 0, 0.5],
                   [0, 0, 0, 1, 0, 0, 0, 1], [3, 4, 0, 2.5, 0, 0, 0, 2.5], [3, 4, 0, 2.5, 0, 0, 0, 2.4999998]], dtype=np.float64)
    a = np.concatenate([ex[:, :4], a]).astype(f32); b = np.concatenate([ex[:, 4:], b]).astype(f32)
    return dict(ss_a=a, ss_b=b, ss_hit=rb.pred_sphere_sphere(a, b))


def micro_cubes():
    rng = np.random.default_rng(103)
    N = 4000
    s = rng.uniform(-3, 3, (N, 4)); s[:, 3] = rng.uniform(0.05, 1.5, N)
    cube = np.concatenate([rng.uniform(-1, 1, (N, 3)), rng.uniform(0.2, 3, (N, 3))], 1)
    ex_s = np.array([[2, 0, 0, 1], [2.0000002, 0, 0, 1], [1001, 0, 0, 1], [1001.0001, 0, 0, 1], [1000.5, 1000.5, 0, 0.7071068], [0, 0, 0, 0.1]], dtype=np.float64)
    ex_c = np.array([[0, 0, 0, 2, 2, 2], [0, 0, 0, 2, 2, 2], [0, 0, 0, 2000, 2000, 2000], [0, 0, 0, 2000, 2000, 2000], [0, 0, 0, 2000, 2000, 2000], [0, 0, 0, 2000, 2000, 2000]], dtype=np.float64)
    s = np.concatenate([ex_s, s]).astype(f32); cube = np.concatenate([ex_c, cube]).astype(f32)
    t = rng.uniform(-2, 2, (N + 6, 9)).astype(f32)
    return dict(sc_s=s, sc_cube=cube, sc_hit=rb.pred_sphere_cube(s, cube), tc_tri=t, tc_cube=cube, tc_hit=rb.pred_triangle_cube(t, cube))


def micro_state():
    rng = np.random.default_rng(104)
    N = 1024
    io = rng.uniform(-20, 20, (N, 18)).astype(f32)
    io[: N // 4, 12:] = 0  # no force / torque
    flags = rng.integers(0, 8, N).astype(np.int32)
    out1 = rb.state_update(io, flags, 5, 1)
    out7 = rb.state_update(io, flags, 5, 7)
    out16 = rb.state_update(io, flags, 16, 3)
    return dict(sv_in=io, sv_flags=flags, sv_out_5ms_1=out1, sv_out_5ms_7=out7, sv_out_16ms_3=out16)


def scene_golden(sc, checkpoints, prefix):
    """For each k in checkpoints: reference state after k ticks, state after k+1 ticks, the ordered hit
    log of tick k+1, and the frozen-state contact sets right after the integration half of tick k+1."""
    out = {}
    w = rb.RefWorld(sc)
    cube, ntri, _ = w.leaves()
    out[f"{prefix}_leaves"] = np.array([len(cube)], dtype=np.int32)
    tl = cube[ntri > 0].astype(np.float64)          # leaves that hold triangles: (cx, cy, cz, edge)
    radii = (sc.obj_spheres[:, 3] * np.repeat(sc.body_scale, np.diff(sc.sphere_start))).astype(np.float64)
    rmax_body = np.maximum.reduceat(radii + np.linalg.norm(sc.obj_spheres[:, :3], axis=1), sc.sphere_start[:-1])

    def leaves_touched(pos):
        """per body: how many triangle-bearing octree leaves its (slightly grown) bounding sphere touches --
        bodies touching >= 2 are where the reference's per-leaf order / last-leaf-wins contact flag (Q6) applies"""
        cnt = np.zeros(pos.shape[0], np.int32)
        for c in tl:
            q = np.clip(pos.astype(np.float64), c[:3] - c[3] / 2, c[:3] + c[3] / 2)
            d2 = ((q - pos) ** 2).sum(1)
            cnt += d2 <= (rmax_body + 1e-3) ** 2
        return cnt
    k = 0
    for cp in checkpoints:
        while k < cp:
            w.step(sc.dt_ms); k += 1
        s0 = w.state()
        # frozen detection on a twin world advanced by the integration half only
        w.integrate(sc.dt_ms)
        a, f = w.detect_frozen()
        ss, st = w.contact_sets((a, f))
        out[f"{prefix}_k{cp}_frozen_ss"] = ss.astype(np.int32)
        out[f"{prefix}_k{cp}_frozen_st"] = st.astype(np.int32)
        s_int = w.state()
        out[f"{prefix}_k{cp}_leaves_touched"] = leaves_touched(s_int["pos"])
        w.collide(sc.dt_ms); k += 1
        a, f = w.hits()
        s1 = w.state()
        for nm, s in (("s0", s0), ("sint", s_int), ("s1", s1)):
            out[f"{prefix}_k{cp}_{nm}_pos"] = s["pos"]; out[f"{prefix}_k{cp}_{nm}_vel"] = s["vel"]; out[f"{prefix}_k{cp}_{nm}_flags"] = s["flags"].astype(np.int32)
        out[f"{prefix}_k{cp}_hits_i"] = a.astype(np.int32)
        out[f"{prefix}_k{cp}_hits_f"] = f
        viz = w.visualize()
        out[f"{prefix}_k{cp}_visualize"] = viz.astype(np.int32)
    c = w.counters()
    out[f"{prefix}_counters"] = np.array([c["tests_ss"], c["tests_st"], c["hits_ss"], c["hits_st"]], dtype=np.int64)
    return out


def extras_golden():
    """frustum planes / AABB culling / the full OSP::getVisibleFrustumCulling on the C1
    octree, and FbxLoader::_buildGeometryData's two branches, all through the reference's own arithmetic."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cams
    rng = np.random.default_rng(105)
    out = {}
    views, projs, invs, planes, mins, maxs, flags = [], [], [], [], [], [], []
    for k in range(40):
        V = cams.view_matrix(rng.uniform(-300, 300, 3), rng.uniform(0, 6.28), rng.uniform(-1, 1))
        P = cams.projection(rng.uniform(30, 100), rng.uniform(1, 2), 0.1, rng.uniform(50, 3000))
        inv = rb.inverse_view_projection(V, P)
        pl = rb.frustum_planes(inv)
        mn = rng.uniform(-1000, 900, (300, 3)).astype(f32); mx = (mn + rng.uniform(1, 400, (300, 3))).astype(f32)
        views.append(V); projs.append(P); invs.append(inv); planes.append(pl); mins.append(mn); maxs.append(mx); flags.append(rb.frustum_aabb(pl, mn, mx))
    out.update(fr_view=np.array(views), fr_proj=np.array(projs), fr_inv=np.array(invs), fr_planes=np.array(planes),
               fr_mins=np.array(mins), fr_maxs=np.array(maxs), fr_flags=np.array(flags))
    w = rb.RefWorld(scenes.config1())
    cube, ntri, _ = w.leaves()
    vis = []
    cv, cp = [], []
    for k in range(12):
        V = cams.view_matrix(rng.uniform(-60, 60, 3) * [1, 0.3, 1], rng.uniform(0, 6.28), rng.uniform(-0.6, 0.2))
        P = cams.projection(60, 1.6, 0.1, rng.uniform(30, 400))
        v = rb.visible_frustum_culling(w, V, P)
        cv.append(V); cp.append(P); vis.append(np.concatenate([v, np.full(len(cube) - len(v), -1, np.int32)]))
    out.update(c1_leaf_cubes=cube, c1_leaf_ntri=ntri.astype(np.int32), c1_cull_view=np.array(cv), c1_cull_proj=np.array(cp),
               c1_cull_inv=np.array([rb.inverse_view_projection(a, b) for a, b in zip(cv, cp)]), c1_cull_visible=np.array(vis))
    # ingest: meshes of 3..60 vertices under random node transforms; every fifth entirely at negative x (the FLT_MIN quirk)
    xs, ids, trss, tris, sphs, mats = [], [], [], [], [], []
    for k in range(60):
        nv = 64; xyz = rng.uniform(-5, 5, (nv, 3)).astype(f32); idx = rng.integers(0, nv, 96).astype(np.int32)
        trs = np.concatenate([rng.uniform(-100, 100, 3), rng.uniform(-180, 180, 3), rng.uniform(0.2, 4, 3)]).astype(f32)
        if k % 5 == 0:
            xyz[:, 0] -= 100; trs[:6] = 0
        xs.append(xyz); ids.append(idx); trss.append(trs)
        mats.append(rb.trs_matrix(trs)); tris.append(rb.ingest_triangles(xyz, idx, trs)); sphs.append(rb.ingest_sphere(xyz, idx, trs))
    out.update(in_xyz=np.array(xs), in_idx=np.array(ids), in_trs=np.array(trss), in_mat=np.array(mats), in_tris=np.array(tris), in_sphere=np.array(sphs))
    return out


def main():
    if not rb.available():
        raise SystemExit("oracle/_ref/libref_physics.so missing: run `make -C oracle ref` (needs /root/reference)")
    ex = extras_golden()
    np.savez_compressed(os.path.join(OUT, "extras.npz"), **ex)
    print("extras.npz", {k: v.shape for k, v in ex.items()})
    if "--extras-only" in sys.argv:
        return

    micro = {}
    micro.update(micro_sphere_triangle()); micro.update(micro_sphere_sphere()); micro.update(micro_cubes()); micro.update(micro_state())
    np.savez_compressed(os.path.join(OUT, "micro.npz"), **micro)
    print("micro.npz", {k: v.shape for k, v in micro.items()})

    g = scene_golden(scenes.config1(), [0, 9, 99, 299, 599], "c1")
    np.savez_compressed(os.path.join(OUT, "c1.npz"), **g)
    print("c1.npz counters", g["c1_counters"], "leaves", g["c1_leaves"])

    sc = scenes.config4(300, 24, 9)
    sc.body_pos[:, 1] -= 1.5
    g = scene_golden(sc, [0, 19, 59, 119], "compound")
    np.savez_compressed(os.path.join(OUT, "compound.npz"), **g)
    print("compound.npz counters", g["compound_counters"])

    g = scene_golden(scenes.pile(150, 5, 0.5, 0), [0, 1, 5], "pile")
    np.savez_compressed(os.path.join(OUT, "pile.npz"), **g)
    print("pile.npz counters", g["pile_counters"])

    g = scene_golden(scenes.config2(2000, 3), [0, 3], "c2")
    np.savez_compressed(os.path.join(OUT, "c2.npz"), **g)
    print("c2.npz counters", g["c2_counters"])


if __name__ == "__main__":
    main()
