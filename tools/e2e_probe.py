import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from reboot_b200 import World
sc, _ = bench.c3_weak_scene(1, 0)
for re in ("0", "32"):
    os.environ["RB_REORDER_EVERY"] = re
    w = World(sc, max_contacts=8 * sc.n_spheres)
    for k in range(300): w.step(5)
    nb = sc.n_bodies
    force = torch.zeros((nb, 4), dtype=torch.float32).pin_memory(); pos = torch.zeros((nb, 4), dtype=torch.float32).pin_memory()
    def run(f, p, n=50):
        for _ in range(3): w.L.rb_step_host(w.h, 5, f, p, None)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): w.L.rb_step_host(w.h, 5, f, p, None)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1000 / n
    print("reorder", re, "none %.3f  pos %.3f  force %.3f  both %.3f ms" % (run(None, None), run(None, pos.data_ptr()), run(force.data_ptr(), None), run(force.data_ptr(), pos.data_ptr())), flush=True)
    w.close()
