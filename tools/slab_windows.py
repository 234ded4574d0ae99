"""Dev probe (torchrun on N GPUs): per-window tick time + list statistics of the weak-scaled C3 bench scene.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/slab_windows.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from reboot_b200 import World, dist as rbdist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); N = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream()
SN = int(os.environ.get("SCENE_N", str(N)))      # SCENE_N=8 on 2 GPUs: the first two slabs of the 8-way scene (small spheres, short skin)
sc, slab = bench.c3_weak_scene(SN, rank)
ids = np.arange(sc.n_bodies, dtype=np.uint32) + np.uint32(rank * bench.C3_SPHERES)
w = World(sc, device=lr, stream=stream.cuda_stream, slab=slab, rank=rank, nranks=N, max_contacts=8 * sc.n_spheres, build=False, ids=ids, rm_max_hint=float(sc.meta["r"]))
w.build()
if os.environ.get("RB_HALO", "p2p") == "nccl": rbdist.init_world(w, rank, N)
else: rbdist.init_world_p2p(w, rank, N)
win = int(os.environ.get("WIN", "25")); prev = w.stats()
for k0 in range(0, int(os.environ.get("TICKS", "300")), win):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(win): w.step(5)
    w.sync(); dist.barrier(); ms = (time.perf_counter() - t0) * 1000 / win
    s = w.stats()
    print("rank %d ticks %4d-%4d  %.4f ms/tick  list ticks %d rebuilds %d hot %d why %d ghosts_in %d migr_in %d" % (rank, k0, k0 + win, ms, s["list_ticks"] - prev["list_ticks"],
          s["list_rebuilds"] - prev["list_rebuilds"], s["list_hot_last"], s["list_rebuild_reason"], s["halo_recv_last"], s["migrated_in_last"]), flush=True)
    prev = s
w.close(); dist.barrier(); dist.destroy_process_group()
