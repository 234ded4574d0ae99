"""Dev probe (GPU box): per-window tick time and list statistics from free fall to rest on C3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from reboot_b200 import World
sc, _ = bench.c3_weak_scene(1, 0)
w = World(sc, max_contacts=8 * sc.n_spheres)
win = int(os.environ.get("WIN", "25"))
prev = w.stats()
for k0 in range(0, int(os.environ.get("TICKS", "700")), win):
    w.sync(); t0 = time.perf_counter()
    for _ in range(win): w.step(5)
    w.sync(); ms = (time.perf_counter() - t0) * 1000 / win
    st = w.stats()
    print("ticks %4d-%4d  %.4f ms/tick  rebuilds %2d/%d  hot_last %6d why %d  tests/tick %.0f" % (k0, k0 + win, ms, st["list_rebuilds"] - prev["list_rebuilds"], win,
          st["list_hot_last"], st["list_rebuild_reason"], (st["tests_st"] + st["tests_ss"] - prev["tests_st"] - prev["tests_ss"]) / win), flush=True)
    prev = st
