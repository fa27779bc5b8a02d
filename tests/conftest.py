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


def nested_tree_world(pkg):
    """Trees behind ray-space pushes, beside other objects and inside a list: every way a tree's span can end."""
    S = pkg.scene
    rng = np.random.default_rng(9)
    white = S.Lambertian(S.SolidColor((0.7, 0.7, 0.7)))

    def cloud(n, lo, hi):
        return S.BvhNode([S.Sphere(tuple(rng.uniform(lo, hi, 3)), float(rng.uniform(0.2, 0.6)), white) for _ in range(n)], 0.0, 1.0)

    a = S.Translation(S.Rotation(1, cloud(24, -3, 3), 20.0), (4.0, 0.5, -2.0))          # pushes x 2 around one tree
    b = S.Translation(S.List([cloud(16, -2, 2), S.Sphere((0.0, 3.0, 0.0), 0.8, white)]), (-5.0, 0.0, 1.0))  # a sibling behind the tree
    c = cloud(12, -8, -4)                                                                  # a bare tree
    d = S.Rotation(0, S.Translation(cloud(10, 1, 3), (0.0, -2.0, 0.0)), -35.0)
    return S.BvhNode([a, b, c, d, S.Sphere((0.0, -1002.0, 0.0), 1000.0, white)], 0.0, 1.0)
