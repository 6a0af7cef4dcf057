import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok() -> bool:
    try:
        from amcl3d_test_b200 import api

        return api.device_info(0)["device_count"] > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: only auto-skip when the user
    # did not ask for gpu tests explicitly.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (checker only)."""
    from oracle import loader

    loader.lib()
    return loader


@pytest.fixture(scope="session")
def world_c1(O):
    """BASELINE config 1 shape: 20x20x5 m @0.1 m, 600 particles, 2k-point cloud; grid by the CPU oracle."""
    from amcl3d_test_b200 import synth

    sm = synth.make_map((20.0, 20.0, 5.0), 0.1)
    mn, mx = O.compute_bounds(sm.points)
    m = O.Map(mn, mx, 0.1, 0.05)
    d2 = O.edt_exact(m, sm.points, sub=2)
    prob = O.grid_from_edt(m, d2, 2, O.GRID_QUIRK)
    m.set_prob(prob)
    cloud = synth.make_cloud(sm, 2000, 12.0)
    return dict(synth=sm, map=m, d2=d2, prob=prob, cloud=cloud,
                T=synth.make_particles(sm, 600, "T"), G=synth.make_particles(sm, 600, "G"))


@pytest.fixture(scope="session")
def gpu_c1(world_c1):
    """Grid3d on cuda:0 holding the oracle-built C1 grid (identical grid on both sides)."""
    from amcl3d_test_b200 import Grid3d

    w = world_c1
    g = Grid3d(0)
    g.setMap(w["map"].mn, w["map"].mx, w["map"].resol)
    g.uploadGrid(w["prob"], w["map"].size, w["map"].sensor_dev)
    return g


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
