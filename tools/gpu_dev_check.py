"""Developer GPU check: prints parity of the CUDA path vs the canonical-order oracle and quick timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reboot_b200 import World, scenes
from oracle import port_binding as pb

def bitdiff(a, b):
    return int((a.view(np.uint32) != b.view(np.uint32)).any(axis=1).sum())

def parity(sc, nsteps, every=10, label=""):
    w = World(sc); o = pb.OrcWorld(sc, octree=False)
    t0 = time.time()
    for k in range(nsteps):
        w.step(5); o.step(5, pb.ORDER_CANONICAL)
        if k % every == every - 1 or k < 2:
            sg, so = w.state(), o.state()
            cg = w.contact_sets(); co = o.contact_sets()
            print(f"[{label}] step {k}: pos bitdiff {bitdiff(sg['pos'], so['pos'])} vel {bitdiff(sg['vel'], so['vel'])} "
                  f"flags {(sg['flags'] != so['flags']).sum()} mt {bitdiff(sg['model_t'], so['model_t'])} pt {bitdiff(sg['prev_t'], so['prev_t'])} "
                  f"| contacts gpu ss/st {len(cg[0])}/{len(cg[1])} oracle {len(co[0])}/{len(co[1])} "
                  f"same_ss {np.array_equal(cg[0], co[0])} same_st {np.array_equal(cg[1], co[1])}", flush=True)
    print(f"[{label}] stats {w.stats()} ({time.time()-t0:.1f}s)")
    w.close()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "parity"):
    parity(scenes.falling_spheres(300, 32, 2.0, 0.5, (0.3, 1.5), True, 11, "mini"), 120, 20, "mini")
    parity(scenes.pile(200, 5, 0.5, 0), 20, 5, "pile")
    parity(scenes.config2(3000, 3), 20, 5, "c2-3k")
    sc4 = scenes.config4(400, 24, 9)
    sc4.body_pos[:, 1] -= 1.5
    parity(sc4, 120, 20, "compound")
if which in ("all", "time"):
    import ctypes
    for name, sc in (("C2-100k", scenes.config2(100_000)), ("C3-1M", scenes.config3())):
        t0 = time.time(); w = World(sc); print(f"[{name}] build {time.time()-t0:.2f}s", w.stats(), flush=True)
        for k in range(5): w.step(5)
        w.sync()
        w.profile(True)
        n = 50
        for k in range(n): w.step(5)
        w.sync()
        kms, _ = w.kernel_ms(); print(f"[{name}] free-fall per-step ms:", {k: v / n for k, v in kms.items()}, flush=True)
        w.profile(False)
        t0 = time.time()
        for k in range(200): w.step(5)
        w.sync(); dt = time.time() - t0
        print(f"[{name}] 200 steps wall {dt*1000/200:.3f} ms/step; stats {w.stats()}", flush=True)
        # settle
        for k in range(600): w.step(5)
        w.sync()
        w.profile(True)
        for k in range(n): w.step(5)
        w.sync()
        kms, _ = w.kernel_ms(); print(f"[{name}] settled per-step ms:", {k: v / n for k, v in kms.items()}, "contacts", w.contact_count(), flush=True)
        w.profile(False)
        t0 = time.time()
        for k in range(200): w.step(5)
        w.sync(); dt = time.time() - t0
        print(f"[{name}] settled 200 steps wall {dt*1000/200:.3f} ms/step; stats {w.stats()}", flush=True)
        w.close()
