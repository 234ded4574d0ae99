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

# does an unrelated transfer slow the tick? (none-mode ticks while a side stream keeps copying 12 MB device -> host)
w = World(sc, max_contacts=8 * sc.n_spheres)
for k in range(600): w.step(5)
side = torch.cuda.Stream(); d = torch.empty(12_000_000, dtype=torch.uint8, device="cuda"); h = torch.empty(12_000_000, dtype=torch.uint8).pin_memory()
def ticks(n, with_copy):
    w.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        w.step(5)
        if with_copy:
            with torch.cuda.stream(side): h.copy_(d, non_blocking=True)
    w.sync(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1000 / n
print("plain ticks %.4f ms   with an unrelated 12 MB D2H per tick on another stream %.4f ms" % (ticks(300, False), ticks(300, True)), flush=True)
def ticks_h2d(n):
    w.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        w.step(5)
        with torch.cuda.stream(side): d.copy_(h, non_blocking=True)
    w.sync(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1000 / n
print("with an unrelated 12 MB H2D per tick %.4f ms" % ticks_h2d(300), flush=True)
w.close()
