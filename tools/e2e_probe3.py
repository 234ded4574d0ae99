"""Dev probe (GPU box): pipelined host round trip vs ring depth (library variants built with -DRB_PIPE=n)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import bench
from reboot_b200 import World
nbuf = int(os.environ.get("NBUF", "4"))
sc, _ = bench.c3_weak_scene(1, 0)
w = World(sc, max_contacts=8 * sc.n_spheres)
for k in range(600): w.step(5)
nb = sc.n_bodies
f3 = [torch.zeros((nb, 3), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
p3 = [torch.zeros((nb, 3), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
def run(use_f, use_p, n=300):
    def one(k):
        w.L.rb_step_host_async3(w.h, 5, f3[k % nbuf].data_ptr() if use_f else None, p3[k % nbuf].data_ptr() if use_p else None, None)
    for k in range(8): one(k)
    w.sync(); t0 = time.perf_counter()
    for k in range(n): one(k)
    w.sync(); return (time.perf_counter() - t0) * 1000 / n
print("nbuf", nbuf, "async3: none %.4f  pos %.4f  force %.4f  both %.4f ms" % (run(0, 0), run(0, 1), run(1, 0), run(1, 1)), flush=True)
w.close()
