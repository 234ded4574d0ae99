"""CPU suite: the C-ABI shared library loads, exports every symbol include/reboot_physics.h declares,
its structs match the Python mirrors, and the product fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from reboot_b200 import _lib


def _declared_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/reboot_physics.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_struct_layouts():
    assert C.sizeof(_lib.RbContact) == 36 and _lib.CONTACT_DTYPE.itemsize == 36
    assert C.sizeof(_lib.RbBodyInit) == 28
    cfg = _lib.RbConfig()
    _lib.load().rb_config_default(C.byref(cfg))
    # the reference's compile-time constants (Physics.cpp:7-9, StateVector.h:28-30, GeometryMath.cpp:509,537)
    assert cfg.world_half == 1000.0 and abs(cfg.gravity + 9.8) < 1e-6 and abs(cfg.friction - 0.95) < 1e-7
    assert abs(cfg.pushout - 1.0001) < 1e-7 and abs(cfg.ss_damp - 0.1) < 1e-8
    assert cfg.nranks == 1 and cfg.slab_lo == -1000.0 and cfg.slab_hi == 1000.0 and cfg.margin_cap == 0.0


def test_no_cpu_fallback_without_device():
    """On a box without a GPU world creation must fail with RB_ERR_NO_DEVICE, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    L = _lib.load()
    h = C.c_void_p()
    rc = L.rb_world_create(None, C.byref(h))
    assert rc == _lib.RB_ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in L.rb_last_error(None)
    from reboot_b200 import World, RebootError
    with pytest.raises(RebootError):
        World()


def test_product_never_touches_oracle():
    """The product package may not import, link, load or execute anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = re.compile(r"import\s+oracle|from\s+oracle|oracle/|libreboot_oracle|libref_physics|reboot_oracle|ref_binding|port_binding")
    for dirpath, _, files in os.walk(os.path.join(root, "reboot_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f"{f} references the oracle"
