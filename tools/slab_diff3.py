"""Dev probe (1 GPU): like slab_diff2 but compared after every tick (use with RB_VL_NEVER_OFF=1: the forced decision does not depend on host reads)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from reboot_b200 import World, dist
import test_gpu_multirank as T
sc, _ = T._scene(os.environ.get("KIND", "cross"))
nslabs = int(os.environ.get("NSLAB", "2")); nticks = int(os.environ.get("TICKS", "30"))
single = World(sc); group = dist.SlabGroup(sc, nslabs)
bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
pa, pg = single.state(), group.state(sc.n_bodies)
for t in range(nticks):
    single.step(5); group.step(5)
    a, g = single.state(), group.state(sc.n_bodies)
    bad = np.nonzero((bits(a["pos"]) != bits(g["pos"])).any(axis=1) | (bits(a["vel"]) != bits(g["vel"])).any(axis=1) | (a["flags"] != g["flags"]))[0]
    if len(bad):
        print("tick", t, "differing", bad[:10])
        ca = single.contacts(); cg = group.contacts() if hasattr(group, "contacts") else None
        for b in bad[:3]:
            print(" body", b, "owner before", int(pg["owner"][b]), "after", int(g["owner"][b]))
            print("   before: pos", pa["pos"][b], "vel", pa["vel"][b], "flags", pa["flags"][b], "| group pos", pg["pos"][b], "vel", pg["vel"][b], "flags", pg["flags"][b], "pt", pa["prev_t"][b], pg["prev_t"][b])
            print("   after : pos", a["pos"][b], "vel", a["vel"][b], "flags", a["flags"][b], "| group pos", g["pos"][b], "vel", g["vel"][b], "flags", g["flags"][b])
            d = np.linalg.norm(pa["pos"] - pa["pos"][b], axis=1); near = np.nonzero(d < 1.6)[0]
            print("   neighbours before", [(int(n), int(pg["owner"][n]), round(float(d[n]), 4)) for n in near if n != b])
            print("   single contacts", [tuple(c) for c in ca if c[1] == b or (c[0] == 0 and c[3] == b)][:6])
            if cg is not None: print("   group contacts ", [tuple(c) for c in cg if c[1] == b or (c[0] == 0 and c[3] == b)][:6])
        print([{k: w.stats()[k] for k in ("list_ticks", "list_rebuilds", "list_hot_last", "list_rebuild_reason", "n_bodies", "migrated_in_last", "migrated_out_last")} for w in group.worlds])
        break
    pa, pg = a, g
else:
    print("no difference")
