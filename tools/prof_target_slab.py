"""Dev probe (GPU box): the ncu target for the slab (multi-GPU) path on ONE GPU. The two ranks' worlds of the weak-scaled C3 bench
scene (bench.c3_weak_scene(2, r): 1M spheres per rank) live in this process as in-process neighbours (LOCAL transport: the same
kernels, K1 stores the halo records into the neighbour's buffers with peer stores), SETTLE untimed ticks, then TICKS ticks inside
cudaProfilerStart/Stop. Per-rank kernel durations and DRAM traffic of one rank = the launches of world 0."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from reboot_b200 import World, _lib
L = _lib.load()
N = 2
worlds = []
caps = []
scs = [bench.c3_weak_scene(N, r) for r in range(N)]
cap = max(int(L.rb_halo_ghost_cap_auto(sc.n_bodies, float(sc.meta["r"]), 0.0, float(slab[1] - slab[0]))) for sc, slab in scs)
for r, (sc, slab) in enumerate(scs):
    ids = np.arange(sc.n_bodies, dtype=np.uint32) + np.uint32(r * bench.C3_SPHERES)
    worlds.append(World(sc, slab=slab, rank=r, nranks=N, ids=ids, rm_max_hint=float(sc.meta["r"]), halo_ghost_cap=cap, max_contacts=8 * sc.n_spheres))
for r, w in enumerate(worlds):
    w._ck(L.rb_dist_attach_peers(w.h, worlds[r - 1].h if r > 0 else None, worlds[r + 1].h if r < N - 1 else None))
arr = (C.c_void_p * N)(*[w.h for w in worlds])
def step():
    assert L.rb_step_group(arr, N, 5) == 0
for _ in range(int(os.environ.get("SETTLE", "600"))): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("TICKS", "3"))): step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", [{k: w.stats()[k] for k in ("n_bodies", "list_ticks", "list_rebuilds", "list_hot_last", "halo_recv_last")} for w in worlds])
