"""Dev probe (1 GPU): the 'cross' scene on NSLAB in-process slab worlds vs a single world, compared only at the END (no host reads
in between, so the host's list decisions run on stale reports like in a long bench run)."""
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
for t in range(nticks):
    single.step(5); group.step(5)
a, g = single.state(), group.state(sc.n_bodies)
bad = np.nonzero((bits(a["pos"]) != bits(g["pos"])).any(axis=1) | (bits(a["vel"]) != bits(g["vel"])).any(axis=1) | (a["flags"] != g["flags"]))[0]
print("ticks", nticks, "differing", bad[:10], [ (a["pos"][b], g["pos"][b], int(g["owner"][b])) for b in bad[:3]])
print([{k: w.stats()[k] for k in ("list_ticks", "list_rebuilds", "list_hot_last", "list_rebuild_reason", "n_bodies")} for w in group.worlds], {k: single.stats()[k] for k in ("list_ticks", "list_rebuilds", "list_hot_last")})
