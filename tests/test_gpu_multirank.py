"""GPU suite, REAL multi-rank transports: two processes, one GPU each, stepping the two x-slabs of one scene
through (a) the fused CUDA-IPC peer-store transport with its arrival-flag handshake (RB_HALO=p2p, the default of
bench.py) and (b) ncclSend / ncclRecv -- state, flags, model translations and contact sets must equal a single
world's bit for bit at every checkpoint. tests/test_gpu_dist.py covers the same kernels with the in-process LOCAL
transport on one GPU; this file is the parity evidence for the transports the scaling numbers are measured with.

Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`); on a 1-GPU box
every test here SKIPS visibly."""
import os
import sys

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(kind):
    from reboot_b200 import scenes
    if kind == "cross":      # lateral velocities: bodies cross the x = 0 face and pile up around it (migrants + ghosts)
        n = 4000
        sc = scenes.falling_spheres(n, 400, 2.0, 0.5, (0.5, 3.0), True, 3, "cross")
        rng = np.random.default_rng(3)
        sc.body_vel[:, 0] = rng.uniform(-12, 12, n).astype(np.float32)
        sc.body_vel[:, 2] = rng.uniform(-3, 3, n).astype(np.float32)
        k = n // 2
        sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0], k) + rng.uniform(-3, 3, k)).astype(np.float32)
        return sc, [0, 2, 14, 59, 104, 149]
    if kind == "compound":   # BASELINE config 4 across ranks: 8-sphere bodies crossing the face (ghosts and migrants are whole bodies)
        n = 3000
        sc = scenes.config4(n, 400, 5)
        rng = np.random.default_rng(5)
        sc.body_vel[:, 0] = rng.uniform(-12, 12, n).astype(np.float32)
        sc.body_vel[:, 2] = rng.uniform(-3, 3, n).astype(np.float32)
        k = n // 2
        sc.body_pos[:k, 0] = (rng.choice([-200.0, 0.0, 200.0], k) + rng.uniform(-4, 4, k)).astype(np.float32)
        sc.body_pos[:, 1] -= 1.0
        return sc, [0, 2, 15, 63, 111, 159]
    n = 6000                 # comes to rest around the face: persistent lists reused, ghosts looked up every tick
    sc = scenes.falling_spheres(n, 400, 2.0, 0.5, (0.3, 1.5), True, 11, "rest")
    rng = np.random.default_rng(11)
    sc.body_vel[:, 0] = rng.uniform(-1.5, 1.5, n).astype(np.float32)
    k = (3 * n) // 4
    sc.body_pos[:k, 0] = (rng.uniform(-40, 40, k)).astype(np.float32)
    return sc, [0, 24, 99, 199, 299]


def _snapshot(w, ids=None):
    s = w.state()
    ss, st = w.contact_sets()
    return dict(ids=(np.arange(len(s["flags"]), dtype=np.uint32) if ids is None else ids), pos=s["pos"], vel=s["vel"], flags=s["flags"],
                model_t=s["model_t"], prev_t=s["prev_t"], ss=ss, st=st)


def _worker(rank, nranks, port, halo, kind, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(nranks), LOCAL_RANK=str(rank))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=nranks, device_id=torch.device("cuda", rank))
        from reboot_b200 import World, dist as rbdist
        sc, cps = _scene(kind)
        sub, ids, slab, rmax = rbdist.partition_scene(sc, nranks, rank)
        stride, reach = rbdist.compound_layout(sc)
        cap = rbdist.common_ghost_cap(sub.n_bodies, max(rmax, reach), slab[1] - slab[0])
        w = World(sub, device=rank, slab=slab, rank=rank, nranks=nranks, ids=ids, rm_max_hint=rmax, halo_ghost_cap=cap,
                  sphere_stride=stride, body_reach_hint=reach)
        if halo == "nccl":
            rbdist.init_world(w, rank, nranks)
        else:
            rbdist.init_world_p2p(w, rank, nranks)
        out = []
        for t in range(cps[-1] + 1):
            w.step(5)
            if t in cps:
                out.append((t, _snapshot(w, w.body_ids())))
        st = w.stats()
        dist.barrier()
        q.put((rank, None, out, {k: st[k] for k in ("halo_sent_last", "halo_recv_last", "migrated_out_last", "list_ticks", "list_rebuilds", "n_bodies")}))
        w.close()
        dist.destroy_process_group()
    except Exception as e:      # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, traceback.format_exc() + repr(e), None, None))


@pytest.mark.parametrize("kind", ["cross", "rest", "compound"])
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_two_ranks_bit_identical_to_single_world(halo, kind):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("SKIPPED, not passed: the P2P / NCCL transport parity test needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500) + (7 if halo == "nccl" else 0) + (13 if kind == "rest" else 0) + (29 if kind == "compound" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, halo, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, err, _, _ in res:
        assert err is None, f"rank {rank}: {err}"
    res.sort(key=lambda r: r[0])
    # the single world, same scene, same ticks
    from reboot_b200 import World
    sc, cps = _scene(kind)
    single = World(sc, device=0)
    nb = sc.n_bodies
    k = 0
    migrated = 0
    for t in range(cps[-1] + 1):
        single.step(5)
        if t in cps:
            a = _snapshot(single)
            merged = {key: np.zeros_like(a[key]) for key in ("pos", "vel", "flags", "model_t", "prev_t")}
            owned = np.zeros(nb, np.int32)
            ss, st = [], []
            for rank, _, out, _ in res:
                tt, s = out[k]
                assert tt == t
                owned[s["ids"]] += 1
                for key in merged:
                    merged[key][s["ids"]] = s[key]
                ss.append(s["ss"]); st.append(s["st"])
            assert (owned == 1).all(), (halo, kind, t, "every body is owned by exactly one rank")
            for key in ("pos", "vel", "model_t", "prev_t"):
                assert np.array_equal(bits(a[key]), bits(merged[key])), (halo, kind, t, key)
            assert np.array_equal(a["flags"], merged["flags"]), (halo, kind, t)
            ss = np.unique(np.concatenate(ss), axis=0) if any(len(x) for x in ss) else np.zeros((0, 4), np.int32)
            st = np.unique(np.concatenate(st), axis=0) if any(len(x) for x in st) else np.zeros((0, 4), np.int32)
            assert np.array_equal(a["ss"], ss) and np.array_equal(a["st"], st), (halo, kind, t)
            k += 1
    stats = [r[3] for r in res]
    assert sum(s["halo_recv_last"] for s in stats) > 0, "ghosts must be in play"
    assert sum(s["n_bodies"] for s in stats) == nb
    if kind == "rest":
        assert all(s["list_ticks"] > 100 for s in stats), stats      # the ranks really stepped from persistent lists
    single.close()
