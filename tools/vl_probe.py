"""Dev probe (GPU box, library built with -DRB_VL_DEBUG): who invalidates the persistent lists on C3?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import bench
from reboot_b200 import World
sc, _ = bench.c3_weak_scene(1, 0)
w = World(sc, max_contacts=8 * sc.n_spheres)
last = None
import time
for k in range(900):
    w.step(5)
    if k == 700:
        w.sync(); t0 = time.perf_counter()
        for _ in range(200): w.step(5)
        w.sync(); print("ms/step over 200 ticks: %.4f" % ((time.perf_counter() - t0) * 1000 / 200), flush=True)
    if k % 50 == 49 or k > 895:
        st = w.stats()
        s = w.state()
        sp = np.linalg.norm(s["vel"], axis=1)
        print(k, "ticks", st["list_ticks"], "rebuilds", st["list_rebuilds"], "hot", st["list_hot_last"], "why", st["list_rebuild_reason"],
              "speed>1: %d  >5: %d  max %.2f" % ((sp > 1).sum(), (sp > 5).sum(), sp.max()), "y<-900:", (s["pos"][:, 1] < -900).sum(),
              "contact flag:", ((s["flags"] & 4) != 0).sum(), flush=True)
