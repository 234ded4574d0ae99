"""Compiles libreboot_b200.so in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["rb_kernels.cu", "rb_world.cu", "rb_dist.cu", "rb_extras.cu"]
OUT = os.path.join(CSRC, "libreboot_b200.so")

NVCC_FLAGS = [
    "-shared", "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",                      # parity: no FMA contraction (reference is x86 SSE without FMA)
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-DRB_HAVE_DIST",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, "rb_device.cuh"), os.path.join(os.path.dirname(_HERE), "include", "reboot_physics.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    flags = list(NVCC_FLAGS)
    if not os.path.exists(os.path.join(CSRC, "rb_dist.cu")):
        flags.remove("-DRB_HAVE_DIST")
    cmd = [nvcc_path()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs + ["-ldl"]
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT
