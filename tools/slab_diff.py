"""Dev probe (GPU box): slab group vs single world (lists) vs single world (per-tick enumeration) on the scene of
test_slab_group_steps_from_persistent_lists, state compared every tick; prints the first divergence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reboot_b200 import World, scenes, dist

nslabs = int(os.environ.get("NSLAB", "2"))
n = 6000
sc = scenes.falling_spheres(n, 400, 2.0, 0.5, (0.3, 1.5), True, 11, "rest")
rng = np.random.default_rng(11)
sc.body_vel[:, 0] = rng.uniform(-1.5, 1.5, n).astype(np.float32)
k = (3 * n) // 4
faces = [0.0] if nslabs == 2 else [-500.0, 0.0, 500.0]
sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0] if nslabs == 4 else faces, k) + rng.uniform(-40, 40, k)).astype(np.float32)
os.environ["RB_VERLET"] = "0"; ref = World(sc)
del os.environ["RB_VERLET"]
single = World(sc)
group = dist.SlabGroup(sc, nslabs)
bits = lambda a: a.view(np.uint32)
prev = ref.state()
for t in range(int(os.environ.get("TICKS", "500"))):
    ref.step(5); single.step(5); group.step(5)
    a, s, g = ref.state(), single.state(), group.state(sc.n_bodies)
    bad_s = (bits(a["pos"]) != bits(s["pos"])).any(axis=1) | (bits(a["vel"]) != bits(s["vel"])).any(axis=1) | (a["flags"] != s["flags"])
    bad_g = (bits(a["pos"]) != bits(g["pos"])).any(axis=1) | (bits(a["vel"]) != bits(g["vel"])).any(axis=1) | (a["flags"] != g["flags"])
    bad_pt_s = (bits(a["prev_t"]) != bits(s["prev_t"])).any(axis=1) | (bits(a["model_t"]) != bits(s["model_t"])).any(axis=1)
    bad_pt_g = (bits(a["prev_t"]) != bits(g["prev_t"])).any(axis=1) | (bits(a["model_t"]) != bits(g["model_t"])).any(axis=1)
    if bad_s.any() or bad_g.any() or bad_pt_s.any() or bad_pt_g.any():
        print("tick", t, "single(lists) differs:", int(bad_s.sum()), "group differs:", int(bad_g.sum()), "mt/pt single:", int(bad_pt_s.sum()), "mt/pt group:", int(bad_pt_g.sum()))
        print("single stats", {k: v for k, v in single.stats().items() if k.startswith("list")})
        for w in group.worlds:
            print("slab stats", {k: v for k, v in w.stats().items() if k.startswith("list") or k.startswith("halo") or k.startswith("migr") or k == "n_bodies"})
        ca = ref.contacts()
        for name, bad, st, wc in (("single", bad_s | bad_pt_s, s, single.contacts()), ("group", bad_g | bad_pt_g, g, None)):
            rows = np.nonzero(bad)[0]
            for r in rows[:6]:
                d = np.linalg.norm(prev["pos"] - prev["pos"][r], axis=1)
                near = np.nonzero(d < 1.3)[0]
                print(name, "body", r, "owner", g["owner"][r] if "owner" in g else None, "pos_prev", prev["pos"][r], "vel_prev", prev["vel"][r], "flags_prev", prev["flags"][r],
                      "near", [(int(q), round(float(d[q]), 4)) for q in near if q != r])
                print("   ref  :", a["pos"][r], a["vel"][r], a["flags"][r], "mt", a["model_t"][r], "pt", a["prev_t"][r])
                print("   other:", st["pos"][r], st["vel"][r], st["flags"][r], "mt", st["model_t"][r], "pt", st["prev_t"][r])
                print("   ref contacts:", [tuple(c) for c in ca if c[0] == r or c[1] == r][:6])
                if wc is not None: print("   other contacts:", [tuple(c) for c in wc if c[0] == r or c[1] == r][:6])
        break
    prev = a
    if os.environ.get("STATS") and t % int(os.environ["STATS"]) == 0:
        keys = ("n_bodies", "list_ticks", "list_rebuilds", "list_hot_last", "list_rebuild_reason", "list_rebuilds_by_reason", "migrated_in_last", "halo_recv_last")
        print(t, "single", [single.stats().get(k) for k in keys[:6]], "slabs", [[w.stats().get(k) for k in keys] for w in group.worlds], flush=True)
else:
    print("no difference in", t + 1, "ticks")
