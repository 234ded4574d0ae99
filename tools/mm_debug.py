import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from reboot_b200 import World, scenes
from oracle import port_binding as port
tris, _ = scenes.terrain_triangles(24, 2.0, 2)
w = World(); o = port.OrcWorld()
w.add_static_geometry(tris); o.L.orc_add_geom(o.h, tris.shape[0], tris); o.n_geoms = 1
rng = np.random.default_rng(3)
for k in range(40):
    X = np.eye(4, dtype=np.float32)
    if k % 3 == 1: X[0, 0] = X[1, 1] = X[2, 2] = np.float32(1.5)
    if k % 3 == 2: X[0, 3], X[1, 3], X[2, 3] = np.float32(0.25), np.float32(-0.125), np.float32(0.5)
    obj = np.array([[0, 0, 0, 0.4]], np.float32)
    p = np.array([rng.uniform(-20, 20), rng.uniform(1.5, 4.0), rng.uniform(-20, 20)], np.float32)
    w.add_model(obj, p, (0, 0, 0), 3, X)
    o.L.orc_add_body(o.h, 1, obj, np.ascontiguousarray(X).ctypes.data_as(C.c_void_p), p, np.zeros(3, np.float32), 1, 1); o.n_bodies += 1
w.build(); o.build(False)
for k in range(61):
    w.step(5); o.step(5, port.ORDER_CANONICAL)
    a, b = w.model_matrices(), o.model_matrices(); sa, sb = w.state(), o.state()
    d = np.nonzero((a[0].reshape(-1,16).view(np.uint32) != b[0].reshape(-1,16).view(np.uint32)).any(1))[0]
    dp = np.nonzero((sa["pos"].view(np.uint32) != sb["pos"].view(np.uint32)).any(1))[0]
    dm = np.nonzero((sa["model_t"].view(np.uint32) != sb["model_t"].view(np.uint32)).any(1))[0]
    if len(d) or len(dp) or len(dm):
        print("step", k, "matrix diff bodies", d, "pos diff", dp, "mt diff", dm)
        for i in d[:2]: print(i, i % 3, a[0][i], b[0][i], sa["model_t"][i], sb["model_t"][i], sa["flags"][i], sb["flags"][i])
        break
print("overflow", w.stats()["margin_overflow"])
