"""ncu target: C4 settled ticks (compound bodies)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reboot_b200 import World, scenes
sc = scenes.config4()
w = World(sc, max_contacts=8 * sc.n_spheres)
for k in range(int(os.environ.get("TICKS", "640"))): w.step(5)
w.sync(); print("done", w.stats()["contacts_last"]); w.close()
