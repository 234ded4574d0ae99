#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 physics step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--settle S] [--impl reference]

A "step" is one physics tick (engine/src/MasterClock.cpp:36-45: every Entity::_updateKinematics,
then Physics::_physicsProcess) of config C3: 1,000,000 spheres (r=0.5) vs a 2,000,000-triangle
procedural heightfield, measured in the SETTLED regime (S untimed ticks first) -- the steady state
and the heavier of the two regimes; the free-fall figure is reported beside it.

N > 1 (one process per GPU under torchrun) is the weak-scaled C3: N x 1M spheres of radius 0.5/sqrt(N)
on a (1000*sqrt(N))^2-cell terrain inside the same 2000^3 world, x-slab partitioned, static geometry
replicated. Boundary spheres (ghosts) and migrating bodies are exchanged with the two x-neighbours every
tick: by default the integrate kernel stores the records straight into the neighbour's memory over NVLink
(CUDA-IPC peer stores + an arrival-flag handshake, RB_HALO=p2p); RB_HALO=nccl uses ncclSend / ncclRecv.
value = N * ticks/s (each tick processes N C3-sized units), so weak-scaling efficiency is value_N / (N * value_1).

Every line also carries `c5`: BASELINE config 5 at FIXED size -- 16,000,000 spheres (r=0.25) on an
8,000,000-triangle terrain plus static box "houses", split into N x-slabs (N = 1: one world) -- with its own
steps/s, so that c5.steps_per_s at N=8 over N=1 is the strong-scaling figure of the north star, and a state
checksum (xor over bodies of a hash of global id and position bits) that must be the same at every N.
`c4` is the same for BASELINE config 4: 262,144 compound bodies x 8 collision spheres, split into N x-slabs.

--impl reference times the reference's own CPU implementation (oracle/_ref = its unmodified physics/ + math/
sources; falls back to the C port): ONE real full-size C3 tick (octree build ~2.5 min + ~2 min per tick on one
core -- the reference step is single-threaded), with the bounded-tile estimate beside it.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT_MS = 5
C3_SPHERES = 1_000_000
C3_G = 1000


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (or None)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return float(json.load(open(p))["traffic_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def c3_weak_scene(nranks: int, rank: int, seed: int = 1234):
    """C3 weak-scaled to `nranks` slabs (nranks == 1 is exactly C3). Returns (scene of this rank, slab)."""
    from reboot_b200 import scenes
    s = math.sqrt(nranks)
    G = int(round(C3_G * s))
    c = 2000.0 / G
    r = 0.5 / s
    lo = -1000.0 + 2000.0 * rank / nranks
    hi = -1000.0 + 2000.0 * (rank + 1) / nranks
    tris, h = scenes.terrain_triangles(G, c, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1 + 1000 * rank * (nranks > 1)))
    half = G * c / 2.0 - 2.0
    n = C3_SPHERES
    x = rng.uniform(max(lo, -half), min(hi, half), n)
    z = rng.uniform(-half, half, n)
    y = rng.uniform(2.0 / s, 12.0 / s, n) + scenes._height_upper(h, G, c, x, z)
    pos = np.stack([x, y, z], axis=1).astype(np.float32)
    sc = scenes.Scene(name=f"C3x{nranks}", tris=[tris], meta=dict(G=G, c=c, r=r, seed=seed),
                      **scenes._single_sphere_bodies(pos, np.zeros_like(pos), r))
    return sc, (lo, hi)


def cpu_sample_scene(frac: int, seed: int = 1234):
    """A spatial 1/frac tile of C3 at the same sphere and triangle density."""
    from reboot_b200 import scenes
    G = max(8, int(round(C3_G / math.sqrt(frac))))
    n = int(round(C3_SPHERES * (G * G) / (C3_G * C3_G)))
    return scenes.falling_spheres(n, G, 2.0, 0.5, (2.0, 12.0), True, seed, f"C3/{frac}"), (G * G) / (C3_G * C3_G)


def time_cpu_reference(frac: int, settle: int, steps: int):
    """Times the reference CPU path on a bounded sample. Returns dict for cpu_baseline / --impl reference."""
    from oracle import ref_binding as rb
    sc, fraction = cpu_sample_scene(frac)
    if rb.available():
        kind = "reference"
        w = rb.RefWorld(sc)
        build_s = w.build_seconds
        t0 = time.perf_counter(); w.bench(DT_MS, settle); t_settle = time.perf_counter() - t0
        w.reset_counters()
        secs = w.bench(DT_MS, steps)
        c = w.counters()
        tests = c["tests_ss"] + c["tests_st"]
    else:
        from oracle import port_binding as pb
        kind = "port"
        t0 = time.perf_counter(); w = pb.OrcWorld(sc); build_s = time.perf_counter() - t0
        w.set_record(False)
        t0 = time.perf_counter()
        for _ in range(settle):
            w.step(DT_MS, pb.ORDER_REFERENCE)
        t_settle = time.perf_counter() - t0
        w.L.orc_reset_counters(w.h)
        t0 = time.perf_counter()
        for _ in range(steps):
            w.step(DT_MS, pb.ORDER_REFERENCE)
        secs = time.perf_counter() - t0
        c = w.counters()
        tests = c["tests_ss"] + c["tests_st"]
    sample_steps_s = steps / secs
    return dict(value=sample_steps_s * fraction, unit="steps/s", cores=1, kind=kind,
                sample=(f"{sc.n_bodies} spheres on a {sc.meta['G']}x{sc.meta['G']}-cell terrain ({sc.n_tris} tris) = 1/{1/fraction:.0f} spatial tile of C3 "
                        f"at the same density; {settle} settle ticks then {steps} timed ticks at {sample_steps_s:.2f} sample-steps/s, scaled by the tile fraction; "
                        f"octree build {build_s:.2f}s and settle {t_settle:.1f}s not timed; single-threaded like the reference's physics thread"),
                sample_steps_per_s=sample_steps_s, sample_tests_per_s=tests / secs, fraction=fraction)


# ------------------------------------------------------------------------------------------
def time_cpu_reference_full(ticks: int):
    """ONE (or a few) REAL full-size C3 tick(s) of the compiled reference: 1M sphere entities + the 2M-triangle
    terrain through Physics::addEntities (octree build) and the MasterClock tick driver, single-threaded."""
    from oracle import ref_binding as rb
    from reboot_b200 import scenes
    sc = scenes.config3()
    t0 = time.perf_counter(); w = rb.RefWorld(); w.load(sc); t_load = time.perf_counter() - t0
    build_s = w.build()
    w.reset_counters()
    secs = w.bench(DT_MS, ticks)
    c = w.counters()
    return dict(value=ticks / secs, unit="steps/s", cores=1, kind="reference",
                sample=(f"the FULL C3 scene (1,000,000 sphere entities, 2,000,000 triangles): {ticks} real tick(s) of the unmodified reference sources "
                        f"from the scene's initial (free-fall) state at {secs / ticks:.1f} s per tick; entity creation {t_load:.0f} s and the octree build "
                        f"(Physics::addEntities) {build_s:.0f} s are not timed; single-threaded like the reference's physics thread. The settled regime "
                        f"costs the reference more per tick (leaf membership only grows, SURVEY Q4), so this figure flatters it"),
                tests_per_s=(c["tests_ss"] + c["tests_st"]) / secs, build_s=build_s, s_per_tick=secs / ticks)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_binding as rb
    t0 = time.time()
    tile = time_cpu_reference(args.ref_frac, args.ref_settle, max(3, min(args.steps, 10)))
    if rb.available() and not args.ref_tile_only:
        cb = time_cpu_reference_full(args.ref_full_ticks)
        same = True
    else:
        cb = dict(tile); cb["tests_per_s"] = tile["sample_tests_per_s"]
        same = False
    line = {
        "impl": "reference", "metric": "physics steps/s @1M spheres vs 2M-tri terrain (C3, settled)", "value": cb["value"], "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: 1,000,000 spheres r=0.5 vs 2,000,000-triangle procedural heightfield (sphere-triangle + sphere-sphere + integration)",
                   "same_config": same, "timed_ticks": args.ref_full_ticks if same else None},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contact_tests_per_s": cb["tests_per_s"], "wall_s": time.time() - t0,
        "tile_estimate": {"value": tile["value"], "unit": "steps/s", "sample": tile["sample"],
                          "note": "linear tile scaling of an octree over-estimates the full-size rate (depth and leaf duplication grow with the scene)"},
    }
    emit(line)


def state_checksum(w, ids=None) -> int:
    """xor over the bodies this world owns of a 32-bit hash of (global id, position bits): order-free, so the ranks'
    values xor into the single-world value when -- and only when -- every body has the same position bits."""
    import ctypes as C
    n = w.stats()["n_bodies"]
    pos = np.zeros((max(n, 1), 4), np.float32)
    w._ck(w.L.rb_get_state(w.h, pos.ctypes.data_as(C.c_void_p), None, None))
    gid = (np.arange(n, dtype=np.uint64) if ids is None else w.body_ids().astype(np.uint64))
    b = pos[:n, :3].copy().view(np.uint32).astype(np.uint64)
    M = np.uint64(0xFFFFFFFF)
    h = (gid * np.uint64(0x9E3779B1)) & M
    for k, mul in enumerate((0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F)):
        h = ((h ^ b[:, k]) * np.uint64(mul)) & M
        h ^= h >> np.uint64(15)
    return int(np.bitwise_xor.reduce(h)) if n else 0


def c5_partition(nranks: int, rank: int, n_total: int, seed: int = 1234):
    """BASELINE config 5 at fixed size, the part one rank owns: the SAME 16M positions on every N (generated in full,
    then cut by x-slab), static geometry replicated. Returns (scene, global ids or None, slab)."""
    from reboot_b200 import scenes, dist as rbdist
    sc = scenes.config5(n_total, 2000, seed, True)
    if nranks == 1:
        return sc, None, (-1000.0, 1000.0)
    sub, ids, slab, _ = rbdist.partition_scene(sc, nranks, rank)
    return sub, ids, slab


def tris_touching_slab(sc, slab, pad):
    """static triangles whose x range meets [slab_lo - pad, slab_hi + pad]: what this rank's step can ever read"""
    n = 0
    for t in sc.tris:
        x = t[:, 0::3]
        n += int(((x.max(axis=1) >= slab[0] - pad) & (x.min(axis=1) <= slab[1] + pad)).sum())
    return n


def run_ours(args):
    import torch
    import torch.distributed as dist
    from reboot_b200 import World

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus
    if world_size != N:
        if world_size == 1 and N > 1:
            raise SystemExit(f"--gpus {N} needs torchrun --nproc-per-node {N}")
        N = world_size
    torch.cuda.set_device(local_rank)
    if N > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    halo = os.environ.get("RB_HALO", "p2p")

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if N == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def xor_over_ranks(x):
        if N == 1:
            return x
        parts = [None] * N
        dist.all_gather_object(parts, int(x))
        out = 0
        for v in parts:
            out ^= v
        return out

    stream = torch.cuda.Stream()

    def make_world(sc, slab, ids, r_hint, max_contacts, stride=0, reach=0.0):
        cap = 0
        if N > 1:
            from reboot_b200 import dist as rbdist
            cap = rbdist.common_ghost_cap(sc.n_bodies, max(r_hint, reach), slab[1] - slab[0])      # one halo buffer layout on every rank
        t0 = time.perf_counter()
        w = World(sc, device=local_rank, stream=stream.cuda_stream, slab=slab, rank=rank, nranks=N,
                  max_contacts=max_contacts, build=False, ids=ids, rm_max_hint=r_hint, halo_ghost_cap=cap,
                  sphere_stride=stride if N > 1 else 0, body_reach_hint=reach if N > 1 else 0.0)
        w.build()
        build_s = time.perf_counter() - t0
        if N > 1:
            from reboot_b200 import dist as rbdist
            if halo == "nccl":
                rbdist.init_world(w, rank, N)            # baseline transport: ncclSend / ncclRecv per tick
            else:
                rbdist.init_world_p2p(w, rank, N)        # K1 stores into the neighbours' memory over NVLink + flag handshake
        return w, build_s

    def timed(w, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(nsteps):
                w.step(DT_MS)
            e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / nsteps

    # =============================== C3 (weak-scaled at N > 1): the headline ===============================
    sc, slab = c3_weak_scene(N, rank)
    ids = (np.arange(sc.n_bodies, dtype=np.uint32) + np.uint32(rank * C3_SPHERES)) if N > 1 else None
    w, build_s = make_world(sc, slab, ids, float(sc.meta["r"]), 8 * sc.n_spheres)
    my_tris = tris_touching_slab(sc, slab, 4.0 * float(sc.meta["r"])) if N > 1 else sc.n_tris
    # ---- free-fall regime (reported beside the headline) ----
    for _ in range(max(3, args.warmup)):
        w.step(DT_MS)
    st_a = w.stats()
    ff_ms = timed(w, min(args.steps, 100))
    st_b = w.stats()
    ff_tests = (st_b["tests_ss"] + st_b["tests_st"] - st_a["tests_ss"] - st_a["tests_st"]) / min(args.steps, 100)
    # ---- settle, warm up, timed region ----
    for _ in range(args.settle):
        w.step(DT_MS)
    for _ in range(max(3, args.warmup)):
        w.step(DT_MS)
    w.sync()
    st0 = w.stats(); l0 = w.launch_count()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_step = timed(w, args.steps)
    clk = clocks.stop() if rank == 0 else None
    st1 = w.stats(); l1 = w.launch_count()
    tests_step = (st1["tests_ss"] + st1["tests_st"] - st0["tests_ss"] - st0["tests_st"]) / args.steps
    contacts = st1["contacts_last"]
    checksum = xor_over_ranks(state_checksum(w, ids))
    # ---- per-kernel durations (CUDA events on the same stream), same number of ticks, right after ----
    w.profile(True)
    for _ in range(args.steps):
        w.step(DT_MS)
    w.sync()
    kms, _ = w.kernel_ms()
    w.profile(False)
    collide_ms = kms["collide"] / args.steps
    # algorithmic bytes of one collide launch (DESIGN.md "roofline"), PER RANK: every in-grid sphere once (16 B), every
    # triangle this rank's slab (+ halo) can touch once (36 B), every contact record (24 B), every resolved body's state
    # read + write (64 B). At N > 1 the replicated triangles outside the slab are NOT counted.
    resolved = min(contacts, st1["n_bodies"])
    alg_bytes = st1["n_spheres"] * 16 + my_tris * 36 + contacts * 24 + resolved * 64
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (collide_ms * 1e-3) / 1e9
    # whole-step algorithmic bytes: N_body*64 + N_tri*36 + N_contact*24, per rank
    step_bytes = st1["n_bodies"] * 64 + my_tris * 36 + contacts * 24
    # ---- e2e: the per-tick host round trip through the C ABI with pinned HOST buffers ----
    nb = st1["n_bodies"]
    rows = int(w.L.rb_host_row_capacity(w.h))      # slab worlds hand out rows; their count moves with migration
    RING = 4          # RB_HOST_PIPE_DEPTH of include/reboot_physics.h: results land three calls later
    slab_world = N > 1
    force = [torch.zeros((rows, 4), dtype=torch.float32).pin_memory() for _ in range(RING)]
    pos = [torch.zeros((rows, 4), dtype=torch.float32).pin_memory() for _ in range(RING)]
    force3 = [torch.zeros((rows, 3), dtype=torch.float32).pin_memory() for _ in range(RING)]
    pos3 = [torch.zeros((rows, 3), dtype=torch.float32).pin_memory() for _ in range(RING)]
    e2e_steps = max(20, min(args.steps, 200))

    def host_tick(k, packed=True):
        # one reference-style tick through the C ABI with HOST buffers: this tick's forces go up from pinned
        # memory, the tick runs, its positions + contact count come back (4-deep ring, see rb_step_host_async).
        # packed: xyz arrays, 12 B per body each way (rb_step_host_async3); else float4 arrays.
        # Slab worlds take no force upload (migrating bodies do not carry forces): their round trip is download only.
        f3 = None if slab_world else force3[k % RING].data_ptr()
        f4 = None if slab_world else force[k % RING].data_ptr()
        if packed:
            rc = w.L.rb_step_host_async3(w.h, DT_MS, f3, pos3[k % RING].data_ptr(), None)
        else:
            rc = w.L.rb_step_host_async(w.h, DT_MS, f4, pos[k % RING].data_ptr(), None)
        assert rc == 0, w.L.rb_last_error(w.h)

    def time_host_ticks(packed):
        for k in range(6):
            host_tick(k, packed)
        w.sync()
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            host_tick(k, packed)
        w.sync()          # the last tick's positions have landed
        barrier()
        return (time.perf_counter() - t0) * 1000 / e2e_steps

    e2e_xyzw_ms = time_host_ticks(False)
    e2e_ms = time_host_ticks(True)
    # blocking variant (upload, tick, download, sync -- nothing overlapped), for the record
    fb = None if slab_world else force[0].data_ptr()
    for _ in range(3):
        w.L.rb_step_host(w.h, DT_MS, fb, pos[0].data_ptr(), None)
    t0 = time.perf_counter()
    for _ in range(20):
        w.L.rb_step_host(w.h, DT_MS, fb, pos[0].data_ptr(), None)
    e2e_blocking_ms = (time.perf_counter() - t0) * 1000 / 20
    e2e_ms = max_over_ranks(e2e_ms)
    final = w.stats()
    w.close()
    del w, force, pos, force3, pos3

    # =============================== C5 at fixed size on N GPUs ===============================
    c5 = None
    if not args.no_c5:
        sc5, ids5, slab5 = c5_partition(N, rank, args.c5_spheres)
        w5, c5_build_s = make_world(sc5, slab5, ids5, 0.25, 4 * sc5.n_spheres + 1024)
        for _ in range(args.c5_settle):
            w5.step(DT_MS)
        for _ in range(max(3, args.warmup)):
            w5.step(DT_MS)
        w5.sync()
        a5 = w5.stats()
        c5_steps = min(args.steps, 100)
        c5_ms = timed(w5, c5_steps)
        b5 = w5.stats()
        c5_sum = xor_over_ranks(state_checksum(w5, ids5))
        c5 = {"workload": f"C5 fixed size: {args.c5_spheres:,} spheres r=0.25 on an 8,000,000-triangle terrain + {sc5.n_tris - 8_000_000} house triangles, {N} x-slab(s)",
              "steps_per_s": 1000.0 / c5_ms, "ms_per_step": c5_ms, "steps": c5_steps, "settle": args.c5_settle,
              "ticks_before_checksum": args.c5_settle + max(3, args.warmup) + c5_steps, "checksum": f"{c5_sum:08x}",
              "rank0": {"n_bodies": b5["n_bodies"], "contacts_last": b5["contacts_last"], "contacts_dropped": b5["contacts_dropped"], "margin_overflow": b5["margin_overflow"],
                        "tests_per_step": (b5["tests_ss"] + b5["tests_st"] - a5["tests_ss"] - a5["tests_st"]) / c5_steps,
                        "halo_recv_last": b5["halo_recv_last"], "list_ticks": b5["list_ticks"], "list_rebuilds": b5["list_rebuilds"], "device_bytes": b5["device_bytes"]},
              "build_s": c5_build_s,
              "note": "strong scaling of the north star = c5.steps_per_s at N=8 over N=1; the checksum must be identical at every N (same scene, same ticks)"}
        w5.close()

    # =============================== C4 at fixed size on N GPUs (compound bodies) ===============================
    c4 = None
    if not args.no_c4:
        from reboot_b200 import scenes, dist as rbdist
        sc4 = scenes.config4(args.c4_bodies)
        stride4, reach4 = rbdist.compound_layout(sc4)
        ids4, slab4 = None, (-1000.0, 1000.0)
        if N > 1:
            sc4, ids4, slab4, _ = rbdist.partition_scene(sc4, N, rank)
        w4, c4_build_s = make_world(sc4, slab4, ids4, float(sc4.meta["r"]), 8 * sc4.n_spheres + 1024, stride4, reach4)
        for _ in range(args.c4_settle + max(3, args.warmup)):
            w4.step(DT_MS)
        w4.sync()
        a4 = w4.stats()
        c4_steps = min(args.steps, 100)
        c4_ms = timed(w4, c4_steps)
        b4 = w4.stats()
        c4_sum = xor_over_ranks(state_checksum(w4, ids4))
        c4 = {"workload": f"C4 fixed size: {args.c4_bodies:,} compound bodies x 8 collision spheres on the C3 terrain (2,000,000 triangles), {N} x-slab(s)",
              "steps_per_s": 1000.0 / c4_ms, "ms_per_step": c4_ms, "steps": c4_steps, "settle": args.c4_settle,
              "ticks_before_checksum": args.c4_settle + max(3, args.warmup) + c4_steps, "checksum": f"{c4_sum:08x}",
              "rank0": {"n_bodies": b4["n_bodies"], "contacts_last": b4["contacts_last"], "contacts_dropped": b4["contacts_dropped"], "margin_overflow": b4["margin_overflow"],
                        "tests_per_step": (b4["tests_ss"] + b4["tests_st"] - a4["tests_ss"] - a4["tests_st"]) / c4_steps,
                        "halo_recv_last": b4["halo_recv_last"], "migrated_out_last": b4["migrated_out_last"]},
              "build_s": c4_build_s,
              "note": "BASELINE config 4 (256k mobile models x 8 collision spheres, 1/2/4/8 GPUs): per-tick enumeration (compound bodies have no persistent lists yet); the checksum must be identical at every N"}
        w4.close()

    if rank == 0:
        cpu = None
        if N == 1 and not args.no_cpu:
            cpu = time_cpu_reference(args.cpu_frac, args.cpu_settle, args.cpu_steps)
        line = {
            "metric": "physics steps/s @1M spheres vs 2M-tri terrain (C3, settled)", "value": N * 1000.0 / ms_step, "unit": "steps/s",
            "n_gpus": N, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: 1,000,000 spheres r=0.5 vs 2,000,000-triangle procedural heightfield (sphere-triangle + sphere-sphere + integration)"
                                    if N == 1 else f"C3 weak-scaled x{N}: {N}x1M spheres r={sc.meta['r']:.4f} on a {sc.meta['G']}^2-cell terrain ({sc.n_tris} tris replicated), x-slabs, halo = " + halo),
                       "regime": f"settled ({args.settle} untimed ticks first)", "dt_ms": DT_MS,
                       "l2": "working set > L2: state + sorted spheres + grids + triangles ~ 0.3 GB per GPU, no flush needed",
                       "unit_note": "value = n_gpus x ticks/s (each tick advances n_gpus C3-sized units)"},
            "e2e": {"value": N * 1000.0 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": 0 if slab_world else nb * 12, "d2h_bytes_per_step": nb * 12 + 8, "ms_per_step": e2e_ms,
                    "xyzw_ms_per_step": e2e_xyzw_ms, "blocking_ms_per_step": e2e_blocking_ms,
                    "note": "rb_step_host_async3: every tick uploads that tick's forces (xyz per body) from pinned host memory, runs the tick and downloads its positions (xyz) + contact count; copies go through a 4-deep ring on their own streams so they overlap neighbouring ticks. xyzw_ms_per_step = the same with float4 arrays (rb_step_host_async), blocking_ms_per_step = float4 with nothing overlapped (rb_step_host). Slab worlds (N > 1) take no force upload: download only"},
            "gpu_launches": int(l1 - l0),
            "clocks": clk,
            "state_checksum": f"{checksum:08x}",
            "build_s": build_s,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic() if N == 1 else None,
                         "kernel": "rb_list_step_kernel (+ the amortised list rebuilds)" if final.get("list_ticks") else "rb_collide_single_kernel<resolve>", "kernel_ms": collide_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "per_rank": True, "tris_counted": my_tris,
                         "step_algorithmic_bytes": step_bytes, "step_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         "kernel_ms_all": {k: v / args.steps for k, v in kms.items()}},
            "contact_tests_per_s": N * tests_step * 1000.0 / ms_step, "contact_tests_per_step": tests_step, "contacts_per_step": contacts,
            "free_fall": {"ms_per_step": ff_ms, "steps_per_s": N * 1000.0 / ff_ms, "contact_tests_per_step": ff_tests},
            "audit": {"contacts_dropped": final["contacts_dropped"], "margin_overflow": final["margin_overflow"],
                      "halo_sent_last": final["halo_sent_last"], "halo_recv_last": final["halo_recv_last"], "device_bytes": final["device_bytes"],
                      "list_ticks": final["list_ticks"], "list_rebuilds": final["list_rebuilds"], "timed_list_rebuilds": st1["list_rebuilds"] - st0["list_rebuilds"], "list_hot_last": final["list_hot_last"], "list_skin": final["list_skin"],
                      "list_rebuilds_by_reason": dict(zip(("grid_or_emigrant", "hot_table_full", "host", "immigrant"), final["list_rebuilds_by_reason"]))},
        }
        if c5 is not None:
            line["c5"] = c5
        if c4 is not None:
            line["c4"] = c4
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, torchrun notices) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--settle", type=int, default=600)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-frac", type=int, default=1024)
    ap.add_argument("--cpu-settle", type=int, default=400)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--ref-frac", type=int, default=512)
    ap.add_argument("--ref-settle", type=int, default=200)
    ap.add_argument("--ref-full-ticks", type=int, default=1, help="real full-size C3 ticks the reference arm times (about 2 min each + 3 min of build)")
    ap.add_argument("--ref-tile-only", action="store_true", help="reference arm: only the bounded tile estimate")
    ap.add_argument("--no-c5", action="store_true", help="skip the fixed-size C5 sub-record")
    ap.add_argument("--no-c4", action="store_true", help="skip the fixed-size C4 (compound bodies) sub-record")
    ap.add_argument("--c4-bodies", type=int, default=262_144)
    ap.add_argument("--c4-settle", type=int, default=600)
    ap.add_argument("--c5-spheres", type=int, default=16_000_000)
    ap.add_argument("--c5-settle", type=int, default=600)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
