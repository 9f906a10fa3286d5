"""The volume packer pinned to the reference's own ``PRD.get_vol_data`` /
``_extract_sweep_volume`` (``pycwr/core/NRadar.py:655-712,1931-1961``): ``tests/golden/packer_reference.npz`` holds
their outputs for aligned / native range modes and ``max_range_km`` = None, a clip inside the table and the
keep-one-gate clip, written by ``tests/golden/make_packer_golden.py`` from /root/reference.

CPU: the host packer (``products.pack_volume`` / ``ArrayPRD.get_vol_data``) reproduces every array bit for bit.
GPU: the device packer (``DeviceVolume.from_sweeps`` / ``pcg_volume_create_raw``) builds a volume whose gridding
products equal, bit for bit, those of a volume uploaded from the reference's own tuple."""
import os

import numpy as np
import pytest

from conftest import FILL, ROOT
from pycwr_b200.products import ArrayPRD, pack_volume

CASES = [("aligned", None), ("aligned", 21.0), ("aligned", 0.1), ("native", None), ("native", 33.0), ("native", 0.05)]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "packer_reference.npz"))


def raw(ref):
    n = int(ref["nsweep"])
    aligned = {f: [(ref["raw%d_az" % i], ref["raw%d_rg" % i], ref["raw%d_%s" % (i, f)]) for i in range(n)]
               for f in ("dBZ", "ZDR")}
    native = {"dBZ": [(ref["raw%d_az" % i], ref["raw%d_rgn" % i], ref["raw%d_dBZn" % i]) for i in range(n)]}
    return aligned, native


def want(ref, ci, field):
    pre = "c%d_%s_" % (ci, field)
    n = int(ref["nsweep"])
    return ([ref[pre + "az%d" % k] for k in range(n)], [ref[pre + "rg%d" % k] for k in range(n)], ref[pre + "fix"],
            [ref[pre + "val%d" % k] for k in range(n)], ref[pre + "site"])


def same_volume(got, w):
    va, vr, fe, vv, site = w
    assert np.array_equal(got[2], fe) and got[2].dtype == np.float64
    assert (got[4], got[5], got[6]) == tuple(site)
    for k in range(len(va)):
        for g, r in ((got[0][k], va[k]), (got[1][k], vr[k]), (got[3][k], vv[k])):
            assert g.dtype == np.float64 and g.shape == r.shape and np.array_equal(g, r), k


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_host_packer_reproduces_the_reference_bit_for_bit(ref, ci):
    mode, max_range = CASES[ci]
    aligned, native = raw(ref)
    prd = ArrayPRD(aligned, ref["fix"], 231.5, lon=118.25, lat=31.75, native_fields=native)
    for field in ("dBZ", "ZDR"):
        got = prd.get_vol_data(field_name=field, fillvalue=FILL, range_mode=mode, max_range_km=max_range)
        same_volume(got, want(ref, ci, field))
        assert not any(np.isnan(v).any() for v in got[3])
    # the plain function on the sweeps the mode selects
    sweeps = native["dBZ"] if mode == "native" else aligned["dBZ"]
    same_volume(pack_volume(sweeps, ref["fix"], 231.5, 118.25, 31.75, FILL, max_range), want(ref, ci, "dBZ"))


def test_default_range_mode_and_keep_one_gate(ref):
    aligned, native = raw(ref)
    prd = ArrayPRD(aligned, ref["fix"], 231.5, native_fields=native)
    assert [r.size for r in prd.get_vol_data("dBZ")[1]] == list(ref["default_dBZ_ng"])      # native by default
    assert [r.size for r in prd.get_vol_data("ZDR")[1]] == list(ref["default_ZDR_ng"])      # aligned by default
    assert all(r.size == 1 for r in prd.get_vol_data("dBZ", max_range_km=0.05)[1])           # NRadar.py:660-661
    with pytest.raises(KeyError):
        prd.get_vol_data("KDP")
    with pytest.raises(ValueError):
        prd.get_vol_data("dBZ", range_mode="full")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_device_packer_against_the_reference_volume(ref, ci, dtype):
    from pycwr_b200 import RadarGridC as G
    from pycwr_b200.runtime import DeviceVolume
    mode, max_range = CASES[ci]
    aligned, native = raw(ref)
    rng = np.random.default_rng(40 + ci)
    ext = 30e3 if max_range is None else max(2.0 * max_range * 1e3, 600.0)
    gx, gy = rng.uniform(-ext, ext, (29, 23)), rng.uniform(-ext, ext, (29, 23))
    lv = np.array([250.0, 400.0, 1200.0, 3000.0, 6500.0])
    for field in ("dBZ", "ZDR"):
        va, vr, fe, vv, site = want(ref, ci, field)
        sweeps = native[field] if (mode == "native" and field in native) else aligned[field]
        host = DeviceVolume(va, vr, fe, vv, fillvalue=FILL, value_dtype=dtype)
        dev = DeviceVolume.from_sweeps(sweeps, ref["fix"], fillvalue=FILL, value_dtype=dtype, max_range_km=max_range)
        try:
            assert dev.value_dtype == host.value_dtype
            assert list(dev.naz) == list(host.naz) and list(dev.ng) == list(host.ng)
            assert np.array_equal(dev.fix_elevation, fe) and np.array_equal(dev.last_gate, host.last_gate)
            for blind in ("mask", "hybrid"):
                a, ai = G.get_CAPPI_3d(None, None, None, host, site[0], gx, gy, lv, FILL, None, blind, return_index=True)
                b, bi = G.get_CAPPI_3d(None, None, None, dev, site[0], gx, gy, lv, FILL, None, blind, return_index=True)
                assert np.array_equal(a, b) and np.array_equal(ai, bi)
                assert max_range is not None or (a != FILL).any()
            assert np.array_equal(G.get_CR_xy(None, None, None, host, site[0], gx, gy, FILL),
                                  G.get_CR_xy(None, None, None, dev, site[0], gx, gy, FILL))
        finally:
            host.close()
            dev.close()
