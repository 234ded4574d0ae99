"""Dev probe (GPU box): per-window tick time and list statistics of in-process slab worlds splitting C3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from reboot_b200 import dist
N = int(os.environ.get("NSLAB", "2"))
from reboot_b200 import scenes
sc = scenes.config3()
g = dist.SlabGroup(sc, N, max_contacts=8 * sc.n_spheres // N)
win = int(os.environ.get("WIN", "25"))
prev = [w.stats() for w in g.worlds]
for k0 in range(0, int(os.environ.get("TICKS", "300")), win):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(win): g.step(5)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1000 / win
    st = [w.stats() for w in g.worlds]
    print("ticks %4d-%4d  %.4f ms/tick (both slabs, one GPU)" % (k0, k0 + win, ms),
          [(s["list_ticks"] - p["list_ticks"], s["list_rebuilds"] - p["list_rebuilds"], s["list_hot_last"], s["list_rebuild_reason"], s["halo_recv_last"], s["migrated_in_last"]) for s, p in zip(st, prev)], flush=True)
    prev = st
