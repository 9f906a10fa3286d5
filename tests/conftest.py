import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FILL = -999.0
BLINDS = ["mask", "nearest_gate", "lowest_sweep", "hybrid", "nearest_valid"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from pycwr_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "gridc_reference.npz"))


def golden_volume(g, prefix, radar=None):
    """Rebuild (vol_azimuth, vol_range, fix_elevation, vol_value) from the fixture arrays."""
    pre = prefix if radar is None else "%sr%d_" % (prefix, radar)
    elev = g[pre + "elev"]
    ne = elev.shape[0]
    if (pre + "az0") in g.files:
        va = [g[pre + "az%d" % k] for k in range(ne)]
        vr = [g[pre + "rng%d" % k] for k in range(ne)]
    else:
        va = [g[pre + "az"]] * ne
        vr = [g[pre + "rng"]] * ne
    vv = [g[pre + "val%d" % k] for k in range(ne)]
    return va, vr, elev, vv


@pytest.fixture(scope="session")
def ref_module():
    """The compiled unmodified reference (only where oracle/_ref was built: this container)."""
    from oracle import build_ref
    return build_ref.load()
