"""Per-kernel times of BASELINE configs C2 and C4 on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reboot_b200 import World, scenes
for name, sc, settle in (("C2 100k box", scenes.config2(100_000), 50), ("C4 262144x8 compound", scenes.config4(), 600)):
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
    print(name, "settled:", {k: round(v / 50 * 1000, 1) for k, v in kms.items()}, "contacts", st["contacts_last"], "overflow", st["margin_overflow"], "dropped", st["contacts_dropped"], flush=True)
    w.close()
