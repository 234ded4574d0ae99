"""Per-kernel times of one rank's share of the weak-scaled C3 scene as a plain single-GPU world."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reboot_b200 import World
for N in ([int(a) for a in sys.argv[1:]] or (1, 2, 4, 8)):
    sc, slab = bench.c3_weak_scene(N, 0)
    w = World(sc, max_contacts=8 * sc.n_spheres)
    for k in range(600): w.step(5)
    w.sync(); w.profile(True)
    for k in range(200): w.step(5)
    w.sync(); kms, _ = w.kernel_ms(); st = w.stats()
    print(N, "r=%.3f G=%d tris=%d" % (sc.meta["r"], sc.meta["G"], sc.n_tris), {k: round(v / 200 * 1000, 1) for k, v in kms.items()},
          "contacts", st["contacts_last"], "tests/step", (st["tests_st"] + st["tests_ss"]) // st["steps"], "cols", st["sphere_cols_x"], st["sphere_cols_z"], st["tri_cols_x"],
          {k: st[k] for k in ("list_ticks", "list_rebuilds", "list_hot_last", "list_skin", "list_rebuilds_by_reason", "margin_overflow")}, flush=True)
    w.close()
