"""Dev probe (GPU box): C5 (16M spheres r=0.25, 8M triangles + houses) on one GPU: per-window tick time, list statistics, stage times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reboot_b200 import World, scenes
n = int(os.environ.get("N", "16000000"))
sc = scenes.config5(n, 2000, 1234, True)
t0 = time.perf_counter(); w = World(sc, max_contacts=4 * sc.n_spheres + 1024); print("build %.2f s" % (time.perf_counter() - t0), flush=True)
win = int(os.environ.get("WIN", "50"))
prev = w.stats()
for k0 in range(0, int(os.environ.get("TICKS", "500")), win):
    w.sync(); t0 = time.perf_counter()
    for _ in range(win): w.step(5)
    w.sync(); ms = (time.perf_counter() - t0) * 1000 / win
    st = w.stats()
    print("ticks %4d-%4d  %.3f ms/tick  list ticks %2d rebuilds %2d  hot_last %7d  contacts %d  tests/tick %.0f (ss %.0f) overflow %d" % (k0, k0 + win, ms, st["list_ticks"] - prev["list_ticks"], st["list_rebuilds"] - prev["list_rebuilds"],
          st["list_hot_last"], st["contacts_last"], (st["tests_st"] + st["tests_ss"] - prev["tests_st"] - prev["tests_ss"]) / win, (st["tests_ss"] - prev["tests_ss"]) / win, st["margin_overflow"]), flush=True)
    prev = st
w.profile(True)
for _ in range(40): w.step(5)
w.sync()
kms, _ = w.kernel_ms()
print({k: round(v / 40, 4) for k, v in kms.items()})
