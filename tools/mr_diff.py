"""Dev probe (2-GPU box): the 'cross' scene of tests/test_gpu_multirank.py on two ranks (P2P transport) against a single world,
compared after EVERY tick; prints the first difference with the ranks' list statistics. Run it a few times: the transports
let the ranks drift by a tick, so a race shows only in some runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

def worker(rank, nranks, port, halo, q, nticks):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(nranks), LOCAL_RANK=str(rank))
    import torch, torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=nranks, device_id=torch.device("cuda", rank))
    from reboot_b200 import World, dist as rbdist
    import test_gpu_multirank as T
    sc, cps = T._scene("cross")
    sub, ids, slab, rmax = rbdist.partition_scene(sc, nranks, rank)
    cap = rbdist.common_ghost_cap(sub.n_bodies, rmax, slab[1] - slab[0])
    w = World(sub, device=rank, slab=slab, rank=rank, nranks=nranks, ids=ids, rm_max_hint=rmax, halo_ghost_cap=cap)
    (rbdist.init_world if halo == "nccl" else rbdist.init_world_p2p)(w, rank, nranks)
    out = []
    sync_every = int(os.environ.get("SYNC_EVERY", "1"))
    for t in range(nticks):
        w.step(5)
        if t % sync_every == sync_every - 1 or t == nticks - 1:
            s = w.state(); st = w.stats()
            out.append((t, w.body_ids().copy(), s["pos"], s["vel"], s["flags"], {k: st[k] for k in ("list_ticks", "list_rebuilds", "list_hot_last", "list_rebuild_reason", "migrated_out_last", "migrated_in_last", "halo_recv_last", "n_bodies")}))
    dist.barrier()
    q.put((rank, out))
    w.close(); dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    halo = os.environ.get("RB_HALO", "p2p"); nticks = int(os.environ.get("TICKS", "60"))
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, 29900 + os.getpid() % 900, halo, q, nticks)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    [p.join() for p in procs]
    from reboot_b200 import World
    import test_gpu_multirank as T
    sc, _ = T._scene("cross")
    single = World(sc, device=0)
    bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
    k = 0
    for t in range(nticks):
        single.step(5)
        if k < len(res[0][1]) and res[0][1][k][0] == t:
            a = single.state()
            pos = np.zeros_like(a["pos"]); vel = np.zeros_like(a["vel"]); fl = np.zeros_like(a["flags"]); owner = np.full(len(fl), -1)
            for rank, out in res:
                _, ids, p, v, f, st = out[k]
                pos[ids] = p; vel[ids] = v; fl[ids] = f; owner[ids] = rank
            bad = np.nonzero((bits(a["pos"]) != bits(pos)).any(axis=1) | (bits(a["vel"]) != bits(vel)).any(axis=1) | (a["flags"] != fl))[0]
            if len(bad):
                print("tick", t, "differing bodies", bad[:12], "owners", owner[bad[:12]])
                for b in bad[:4]:
                    print("  body", b, "single pos", a["pos"][b], "vel", a["vel"][b], "flags", a["flags"][b], "| ranks pos", pos[b], "vel", vel[b], "flags", fl[b])
                    d = np.linalg.norm(a["pos"] - a["pos"][b], axis=1); near = np.nonzero(d < 1.5)[0]
                    print("    neighbours", [(int(n), int(owner[n]), round(float(d[n]), 3)) for n in near if n != b])
                for rank, out in res:
                    print("  rank", rank, "stats now", out[k][5], "prev", out[k - 1][5] if k else None)
                break
            k += 1
    else:
        print("no difference in", nticks, "ticks;", [out[-1][5] for _, out in res])
