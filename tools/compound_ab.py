"""A/B of the compound-body kernels: KIND=split|group (RB_COMPOUND) vs the per-body kernel (legacy), tick by tick."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reboot_b200 import World, scenes

def make(edge):
    if not edge:
        return World(scenes.config4(int(os.environ.get("NB", "4096")), 64))
    tris, _ = scenes.terrain_triangles(8, 2.0, 1)
    w = World(); w.add_static_geometry(tris)
    rng = np.random.default_rng(4)
    specs = [(1, (0.0, 1.2, 0.0), 3), (3, (2.0, 1.3, 1.0), 3), (5, (-3.0, 1.4, 2.0), 3), (1, (1.0, 1.1, -3.0), 2), (1, (1000.6, 0.0, 0.0), 3), (1, (1000.4, 0.0, 0.0), 3)]
    for ns, p, fl in specs:
        obj = np.zeros((ns, 4), np.float32); obj[:, :3] = rng.uniform(-0.3, 0.3, (ns, 3)); obj[:, 3] = 0.5
        w.add_model(obj, p, (0, -1, 0), fl)
    w.build(); return w

edge = "--edge" in sys.argv
os.environ["RB_COMPOUND"] = os.environ.get("KIND", "split"); a = make(edge)
os.environ["RB_COMPOUND"] = "legacy"; b = make(edge); os.environ.pop("RB_COMPOUND")
for k in range(int(os.environ.get("TICKS", "300"))):
    a.step(5); b.step(5)
    sa, sb = a.state(), b.state()
    bad = np.nonzero((sa["pos"].view(np.uint32) != sb["pos"].view(np.uint32)).any(axis=1) | (sa["vel"].view(np.uint32) != sb["vel"].view(np.uint32)).any(axis=1) | (sa["flags"] != sb["flags"]))[0]
    if len(bad):
        print("tick", k, "bodies", bad[:10], "of", len(bad))
        for i in bad[:3]: print(" ", i, sa["pos"][i], sb["pos"][i], sa["vel"][i], sb["vel"][i], sa["flags"][i], sb["flags"][i])
        break
else:
    print("identical for all ticks;", a.stats()["tests_st"], b.stats()["tests_st"], a.stats()["tests_ss"], b.stats()["tests_ss"], a.stats()["hits_st"], b.stats()["hits_st"])
