"""Dev helper: build an A/B variant of the library with extra -D flags into gpurun_variants/<name>.so
(select it at run time with REBOOT_B200_LIB=<path>).   python tools/build_variant.py name -DRB_X=1 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reboot_b200 import build as B

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "variants"); os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"lib_{name}.so")
srcs = [os.path.join(B.CSRC, s) for s in B.SOURCES]
B.link_library(B.NVCC_FLAGS + extra, srcs, out, tag="_" + name)
print(out)
