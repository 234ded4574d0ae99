"""CPU suite, world_size 2 over gloo: the host-side logic of the multi-GPU path -- slab ownership,
scene partitioning, the unique-id rendezvous and the agreement on the halo buffer layout -- without any GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from reboot_b200 import dist as rbdist, scenes
    sc = scenes.falling_spheres(5000, 200, 2.0, 0.5, (1.0, 4.0), True, 9, "gloo")
    sub, ids, slab, rmax = rbdist.partition_scene(sc, world_size, rank)
    # every body is owned exactly once, by the slab that contains it
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(counts, torch.tensor([len(ids)], dtype=torch.int64))
    owned = torch.zeros(sc.n_bodies, dtype=torch.int32)
    owned[torch.from_numpy(ids.astype(np.int64))] = 1
    dist.all_reduce(owned)
    ok = bool((owned == 1).all()) and sum(int(c) for c in counts) == sc.n_bodies
    ok &= bool(((sub.body_pos[:, 0] >= slab[0]) | (rank == 0)).all() and ((sub.body_pos[:, 0] < slab[1]) | (rank == world_size - 1)).all())
    ok &= abs(rmax - 0.5) < 1e-7 and sub.n_tris == sc.n_tris        # static geometry replicated, same reach everywhere
    # unique-id rendezvous: rank 0's bytes arrive everywhere
    uid = rbdist.broadcast_unique_id(lambda: bytes(range(128)), rank)
    ok &= uid.tobytes() == bytes(range(128))
    # one halo buffer layout for every rank: the maximum of what each would pick for its own (uneven) body count
    from reboot_b200 import _lib
    L = _lib.load()
    n_fake = 400_000 + 150_000 * rank                      # uneven partition: the per-rank values differ ...
    mine = int(L.rb_halo_ghost_cap_auto(n_fake, 0.5, 0.0, 1000.0))
    cap = rbdist.common_ghost_cap(n_fake, 0.5, 1000.0)
    caps = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(caps, torch.tensor([cap], dtype=torch.int64))
    ok &= all(int(c) == cap for c in caps) and cap >= mine   # ... the agreed one is the same everywhere and covers everybody
    ok &= cap == int(L.rb_halo_ghost_cap_auto(400_000 + 150_000 * (world_size - 1), 0.5, 0.0, 1000.0))
    q.put((rank, ok, len(ids)))
    dist.destroy_process_group()


def test_partition_and_rendezvous_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == 5000


def test_owner_of_and_slab_bounds():
    sys.path.insert(0, ROOT)
    from reboot_b200 import dist as rbdist
    assert rbdist.slab_bounds(0, 4) == (-1000.0, -500.0) and rbdist.slab_bounds(3, 4) == (500.0, 1000.0)
    x = np.array([-1500.0, -1000.0, -500.0001, -500.0, 0.0, 499.9, 500.0, 999.0, 1200.0])
    assert rbdist.owner_of(x, 4).tolist() == [0, 0, 0, 1, 2, 2, 3, 3, 3]
    assert (rbdist.owner_of(x, 1) == 0).all()


def _compound_worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from reboot_b200 import dist as rbdist, scenes
    # ragged compound bodies (1..8 spheres): ownership by the body's model translation, spheres travel with their body
    sc = scenes.config4(3000, 200, 7)
    rng = np.random.default_rng(7)
    cnt = rng.integers(1, 9, sc.n_bodies)
    keep = np.concatenate([np.arange(8 * b, 8 * b + cnt[b]) for b in range(sc.n_bodies)])
    sc.obj_spheres = np.ascontiguousarray(sc.obj_spheres[keep]); sc.sphere_start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    sub, ids, slab, rmax = rbdist.partition_scene(sc, world_size, rank)
    stride, reach = rbdist.compound_layout(sc)
    owned = torch.zeros(sc.n_bodies, dtype=torch.int32)
    owned[torch.from_numpy(ids.astype(np.int64))] = 1
    dist.all_reduce(owned)
    ok = bool((owned == 1).all())
    ok &= bool((rbdist.owner_of(sub.body_pos[:, 0], world_size) == rank).all())
    ok &= sub.n_spheres == int(cnt[ids].sum()) and sub.sphere_start[-1] == sub.n_spheres and len(sub.sphere_start) == len(ids) + 1
    # every sub-scene sphere is the global scene's sphere of the same body and index
    b0 = int(ids[0]); s0 = int(sc.sphere_start[b0])
    ok &= np.array_equal(sub.obj_spheres[: cnt[b0]], sc.obj_spheres[s0: s0 + cnt[b0]])
    # the layout every rank must agree on: the same on all of them, and it covers every body
    lay = [None] * world_size
    dist.all_gather_object(lay, (stride, round(reach, 6)))
    ok &= all(l == lay[0] for l in lay) and stride == int(cnt.max()) and reach >= float((np.abs(sc.obj_spheres[:, 0]) + sc.obj_spheres[:, 3]).max())
    ok &= abs(rmax - 0.4) < 1e-6
    q.put((rank, ok, len(ids)))
    dist.destroy_process_group()


def test_compound_partition_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_compound_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == 3000
