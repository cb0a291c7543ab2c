import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle and the CUDA library are built (cheap no-op when fresh)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def tz2():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "tz2.npz"))
    d = dict(crd=g["crd"].reshape(101, -1), names=g["names"], resnum=g["resnum"], mass=g["mass"])
    d["res"] = lambda lo, hi: np.nonzero((d["resnum"] >= lo) & (d["resnum"] <= hi))[0].astype(np.int32)
    return d


@pytest.fixture(scope="session")
def saves():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_saves.npz"))


@pytest.fixture(scope="session")
def live():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_live.npz"))


@pytest.fixture(scope="session")
def b200(built):
    """Initialised CUDA library on device 0 (gpu tests only)."""
    import cpptraj_b200 as b
    b.init(1)
    yield b
    b.shutdown()
