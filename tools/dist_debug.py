import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from reboot_b200 import World, dist, scenes
from test_gpu_dist import _crossing_scene
sc = _crossing_scene(scenes)
g = dist.SlabGroup(sc, 2)
for k in range(40):
    g.step(5)
    st = [w.stats() for w in g.worlds]
    print(k, [(s["n_bodies"], s["halo_sent_last"], s["halo_recv_last"], s["migrated_out_last"], s["migrated_in_last"]) for s in st])
