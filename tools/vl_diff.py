"""Dev probe (GPU box): list stepping vs per-tick enumeration on C3, state compared every tick."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reboot_b200 import World, scenes
sc = scenes.config3()
os.environ["RB_VERLET"] = "0"; wa = World(sc)
del os.environ["RB_VERLET"]; wb = World(sc)
bits = lambda a: a.view(np.uint32)
prev = wa.state(); pst = wb.stats()
for k in range(int(os.environ.get("TICKS", "300"))):
    wa.step(5); wb.step(5)
    a, b = wa.state(), wb.state()
    bad = (bits(a["pos"]) != bits(b["pos"])).any(axis=1) | (bits(a["vel"]) != bits(b["vel"])).any(axis=1)
    sb = wb.stats()
    if bad.any():
        rows = np.nonzero(bad)[0]
        print("tick", k, "differing", len(rows), "rebuild this tick:", sb["list_rebuilds"] - pst["list_rebuilds"], "hot", sb["list_hot_last"], "reason", sb["list_rebuild_reason"])
        ca, cb = wa.contacts(), wb.contacts()
        print("contacts", len(ca), len(cb))
        for r in rows[:4]:
            d = np.linalg.norm(prev["pos"] - prev["pos"][r], axis=1)
            near = np.nonzero(d < 1.3)[0]
            print(" body", r, "pos_prev", prev["pos"][r], "vel_prev", prev["vel"][r], "near", [(int(n), float(d[n])) for n in near if n != r])
            print("   fused:", a["pos"][r], a["vel"][r], " lists:", b["pos"][r], b["vel"][r])
            print("   fused contacts:", [tuple(c) for c in ca if c[0] == r or c[1] == r][:4], " lists contacts:", [tuple(c) for c in cb if c[0] == r or c[1] == r][:4])
        break
    prev = a; pst = sb
else:
    print("no difference")
