import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (hyper-ray-tracer_b200/)."""
    return graft.load_package()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle binding (test infrastructure)."""
    return graft.load_oracle()


@pytest.fixture(scope="session")
def has_gpu(pkg):
    return pkg.native.device_count() > 0


def make_rays(orc_mod, origins, directions, time=0.0, tmin=0.001, tmax=np.inf):
    o = np.asarray(origins, dtype=np.float32).reshape(-1, 3)
    d = np.asarray(directions, dtype=np.float32).reshape(-1, 3)
    rays = np.zeros(o.shape[0], dtype=orc_mod.RAY_DTYPE)
    rays["o"] = o
    rays["d"] = d
    rays["time"] = time
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


def build_both(pkg, orc_mod, world):
    """Emit one description onto libhrt (host-side flattener only) and onto the oracle."""
    gb = pkg.HrtBackend()
    ob = orc_mod.OracleBackend()
    e1 = pkg.scene.emit(world, gb)
    e2 = pkg.scene.emit(world, ob)
    assert e1.root == e2.root
    return gb, ob, e1, e2
