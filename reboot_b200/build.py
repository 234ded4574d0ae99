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
    "-DRB_HAVE_CDP",                    # a second, relocatable pass over rb_kernels.cu provides the device-launched list rebuild
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
    try:
        link_library(flags, srcs, OUT, verbose)
    except (subprocess.CalledProcessError, OSError) as e:
        if "-DRB_HAVE_CDP" not in flags:
            raise
        # the relocatable pass (device-side launches, needs libcudadevrt) is an optimisation, not a requirement:
        # without it the host enqueues the list-rebuild chain itself
        print(f"[build] relocatable pass failed ({e}); building without device-side launches")
        link_library([f for f in flags if f != "-DRB_HAVE_CDP"], srcs, OUT, verbose)
    return OUT


def link_library(flags, srcs, out, verbose=False, tag=""):
    """normal pass over every source + the relocatable chain pass over rb_kernels.cu (gate kernel and the kernels it
    tail-launches), device-linked against cudadevrt, all into one shared library"""
    nv = nvcc_path()
    cflags = [f for f in flags if f != "-shared"]
    objdir = os.path.join(CSRC, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    v = ["-Xptxas", "-v"] if verbose else []
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([nv] + cflags + v + ["-c", s, "-o", o], cwd=CSRC))
    extra = []
    if "-DRB_HAVE_CDP" in flags:
        chain_o = os.path.join(objdir, "rb_kernels.chain.o"); dlink_o = os.path.join(objdir, "rb_kernels.chain.dlink.o")
        p = subprocess.Popen([nv] + cflags + v + ["-rdc=true", "-DRB_CHAIN_TU", "-c", os.path.join(CSRC, "rb_kernels.cu"), "-o", chain_o], cwd=CSRC)
        if p.wait():
            raise subprocess.CalledProcessError(p.returncode, "nvcc (chain pass)")
        subprocess.check_call([nv, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-dlink", chain_o, "-o", dlink_o, "-lcudadevrt"], cwd=CSRC)
        extra = [chain_o, dlink_o]
    for p in procs:
        if p.wait():
            raise subprocess.CalledProcessError(p.returncode, "nvcc")
    subprocess.check_call([nv, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", out] + objs + extra + ["-ldl", "-lcudadevrt"], cwd=CSRC)
