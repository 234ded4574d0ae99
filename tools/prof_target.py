"""Dev probe (GPU box): the ncu target. C3, SETTLE untimed ticks, then TICKS ticks inside cudaProfilerStart/Stop
(run ncu with --profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reboot_b200 import World, scenes
cfg = os.environ.get("CFG", "c3")
sc = {"c3": scenes.config3, "c2": scenes.config2, "c4": scenes.config4}[cfg]()
w = World(sc)
for _ in range(int(os.environ.get("SETTLE", "600"))):
    w.step(5)
w.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("TICKS", "4"))):
    w.step(5)
w.sync()
torch.cuda.profiler.stop()
print("ok", {k: v for k, v in w.stats().items() if k.startswith("list")})
w.close()
