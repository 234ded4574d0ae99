import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    out = {}
    for name in ("micro", "c1", "compound", "pile", "c2"):
        with np.load(os.path.join(d, name + ".npz")) as z:
            out[name] = {k: z[k] for k in z.files}
    return out


@pytest.fixture(scope="session")
def port():
    from oracle import port_binding as pb
    pb.lib()
    return pb


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libref_physics.so not built (needs /root/reference)")
    rb.lib()
    return rb
