"""The header-only C++ facade (include/reboot/Physics.hpp: reference class/method names over the C ABI).
CPU: it compiles and links against libreboot_b200.so. GPU: reference-style host code produces the
same state as the same scene driven through the ctypes mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "facade_test")
    lib = os.path.join(ROOT, "reboot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
                           "-L" + lib, "-lreboot_b200", "-Wl,-rpath," + lib, "-o", exe])
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU path" in r.stderr


def test_master_clock_tick_driver(tmp_path):
    """MasterClock facade (5 ms kinematics tick, subscription order, overrun accounting) -- no GPU involved."""
    exe = str(tmp_path / "clock_test")
    lib = os.path.join(ROOT, "reboot_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "clock_test.cpp"),
                           "-L" + lib, "-lreboot_b200", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    ticks, calls, ms, alternating = int(out[1]), int(out[3]), int(out[5]), int(out[7])
    assert 25 <= ticks <= 45 and calls == ticks and ms == 5 and alternating == 1      # 210 ms / 5 ms, scheduler slack allowed
    slow_ticks, slow_over = int(out[12]), int(out[14])
    assert 4 <= slow_ticks <= 9 and slow_over == slow_ticks


@pytest.mark.gpu
def test_facade_matches_c_abi(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from reboot_b200 import World
    exe = _build(tmp_path)
    run = subprocess.run([exe, "6", "120"], capture_output=True, text=True, check=True)
    # culling over the static cells + collider ingest through the facade: x-extent of the scaled mesh is 8 -> r = 4,
    # centre = min + r on every axis (min = (8, -2, -4)); the NDC cube meets the four cells around the origin
    assert "extras: visible 4 sphere r 4 cx 12 cy 2 tris 2" in run.stderr, run.stderr
    out = run.stdout.strip().splitlines()
    rows = np.array([[float(x) for x in ln.split()] for ln in out[:-1]])
    assert int(out[-1].split()[1]) > 0
    # same scene through the ctypes mirror
    G, c, half = 16, 2.0, 16.0
    tris = []
    for z in range(G):
        for x in range(G):
            b = (x * c - half, 0, z * c - half); b1 = ((x + 1) * c - half, 0, z * c - half)
            bs = (x * c - half, 0, (z + 1) * c - half); bs1 = ((x + 1) * c - half, 0, (z + 1) * c - half)
            tris.append(bs + b1 + b); tris.append(bs + bs1 + b1)
    w = World()
    w.add_static_geometry(np.array(tris, np.float32))
    for i in range(6):
        for j in range(6):
            w.add_model([[0, 0, 0, 0.5]], (np.float32(-10.0) + np.float32(3.1) * np.float32(i), np.float32(1.0) + np.float32(0.25) * np.float32((i * 7 + j * 3) % 5),
                                           np.float32(-9.0) + np.float32(2.9) * np.float32(j)))
    w.build()
    for k in range(120):
        w.step(5)
    w.set_force(0, [[3.0, 0, 0]])
    for k in range(10):
        w.step(5)
    s = w.state()
    assert np.allclose(rows[:, 1:4], s["pos"], rtol=0, atol=1e-6) and np.allclose(rows[:, 4:7], s["vel"], rtol=0, atol=1e-6)
    assert np.array_equal(rows[:, 7].astype(int), (s["flags"] >> 2) & 1)
    assert s["vel"][0, 0] > 0.0 and (s["flags"][0] & 4)   # the force stuck because the body is in contact
    w.close()
