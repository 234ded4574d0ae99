"""Dev probe (GPU box): where does the pipelined host round trip spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from reboot_b200 import World
# raw PCIe: 16 MB each way, alone and concurrently
n = 16_000_000
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=50):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1000 / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
print("pcie 16MB: h2d %.3f ms  d2h %.3f ms  both concurrently %.3f ms" % (t(h2d), t(d2h), t(both)), flush=True)

sc, _ = bench.c3_weak_scene(1, 0)
w = World(sc, max_contacts=8 * sc.n_spheres)
for k in range(600): w.step(5)
nb = sc.n_bodies
force = [torch.zeros((nb, 4), dtype=torch.float32).pin_memory() for _ in range(3)]
pos = [torch.zeros((nb, 4), dtype=torch.float32).pin_memory() for _ in range(3)]
import ctypes as C
cnt = C.c_uint64(0)
def run(use_f, use_p, n=200):
    def one(k):
        f = force[k % 3].data_ptr() if use_f else None; p = pos[k % 3].data_ptr() if use_p else None
        w.L.rb_step_host_async(w.h, 5, f, p, C.byref(cnt))
    for k in range(6): one(k)
    w.sync(); t0 = time.perf_counter()
    for k in range(n): one(k)
    w.sync(); return (time.perf_counter() - t0) * 1000 / n
print("async: none %.4f  pos %.4f  force %.4f  both %.4f ms" % (run(0, 0), run(0, 1), run(1, 0), run(1, 1)), flush=True)
w.close()
