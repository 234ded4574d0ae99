"""C4 (262,144 compound bodies x 8 spheres) per-kernel times: scene order as generated vs bodies pre-sorted spatially."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reboot_b200 import World, scenes

def run(name, sc, settle=600):
    t0 = time.time(); w = World(sc, max_contacts=8 * sc.n_spheres); tb = time.time() - t0
    for k in range(20): w.step(5)
    w.sync(); w.profile(True)
    for k in range(50): w.step(5)
    w.sync(); kms, _ = w.kernel_ms(); w.profile(False)
    print(name, "build %.2fs" % tb, "early:", {k: round(v / 50 * 1000, 1) for k, v in kms.items()}, flush=True)
    for k in range(settle): w.step(5)
    w.sync(); w.profile(True)
    for k in range(50): w.step(5)
    w.sync(); kms, _ = w.kernel_ms(); st = w.stats()
    print(name, "settled:", {k: round(v / 50 * 1000, 1) for k, v in kms.items()}, "sum %.1f" % (sum(kms.values()) / 50 * 1000), "contacts", st["contacts_last"], "tests_st", st["tests_st"], "tests_ss", st["tests_ss"], flush=True)
    w.profile(False)
    t0 = time.time()
    for k in range(200): w.step(5)
    w.sync(); print(name, "wall ms/tick %.4f" % ((time.time() - t0) / 200 * 1000), flush=True)
    w.close()

sc = scenes.config4()
if "--presort" in sys.argv:
    p = sc.body_pos
    key = (np.floor((p[:, 2] + 1000) / 8).astype(np.int64) * 4096 + np.floor((p[:, 0] + 1000) / 8).astype(np.int64)) * 64 + (np.floor((p[:, 2] + 1000)).astype(np.int64) % 8) * 8 + np.floor((p[:, 0] + 1000)).astype(np.int64) % 8
    o = np.argsort(key, kind="stable")
    sc.body_pos = sc.body_pos[o]; sc.body_vel = sc.body_vel[o]
run("C4" + (" presorted" if "--presort" in sys.argv else ""), sc, int(os.environ.get("SETTLE", "600")))
