"""Local moving-window VVP: oracle pinning on the CPU, kernel parity on the GPU.

Golden values: the reference's own ``_run_local_vvp`` (tests/golden/make_vvp_golden.py).  Bars: window
statistics (valid_count, valid_fraction, azimuth_coverage_deg, azimuth_sector_count) and every NaN mask
bit-exact; u / v / offset / speed / rmse within 1e-5 relative (north_star's bar for velocities — the
kernel sums the normal equations in a different order and its medians are exact, so the observed
difference is ~1e-12); direction within 1e-5 deg modulo 360; condition number within 1e-6 relative.
"""
import os
import sys

import numpy as np
import pytest

from oracle import vvp_oracle as vo

HERE = os.path.dirname(__file__)
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_vvp_golden import CASES, synthetic_sweep  # noqa: E402

GOLD = os.path.join(HERE, "golden", "vvp_reference.npz")
STATS = ("valid_count", "azimuth_sector_count", "valid_fraction", "azimuth_coverage_deg")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def check(out, ref, rtol, cond_rtol, max_cond):
    for k in STATS:
        assert out[k].dtype == ref[k].dtype
        assert np.array_equal(out[k], ref[k], equal_nan=True), k
    for k in ("u", "v", "radial_offset", "wind_speed", "fit_rmse", "wind_direction", "condition_number"):
        assert np.array_equal(np.isnan(out[k]), np.isnan(ref[k])), k
    m = np.isfinite(ref["u"])
    for k in ("u", "v", "radial_offset", "wind_speed", "fit_rmse"):
        np.testing.assert_allclose(out[k][m], ref[k][m], rtol=rtol, atol=rtol, err_msg=k)
    d = np.abs(out["wind_direction"][m] - ref["wind_direction"][m])
    assert np.all(np.minimum(d, 360.0 - d) <= 1e-5)
    c = np.isfinite(ref["condition_number"])
    ok = ref["condition_number"][c] <= max_cond
    np.testing.assert_allclose(out["condition_number"][c][ok], ref["condition_number"][c][ok], rtol=cond_rtol)
    assert np.all(out["condition_number"][c][~ok] > max_cond * (1 - cond_rtol))


def test_oracle_matches_reference(gold):
    for name, seed, naz, nrange, elev, az_num, bin_num, kw in CASES:
        ref = {k: gold["%s_%s" % (name, k)] for k in vo.F64_KEYS + vo.I32_KEYS}
        out = vo.run_local_vvp(gold[name + "_az"], gold[name + "_el"], gold[name + "_vr"], az_num, bin_num, **kw)
        check(out, ref, 1e-9, 1e-9, kw.get("max_condition_number", 1e9))
        assert np.isfinite(ref["u"]).sum() > 100


def test_golden_inputs_are_reproducible(gold):
    from make_vvp_golden import SECTORS
    for name, seed, naz, nrange, elev, az_num, bin_num, kw in CASES:
        az, el, vr = synthetic_sweep(seed, naz, nrange, elev, sector=SECTORS.get(name))
        assert np.array_equal(az, gold[name + "_az"]) and np.array_equal(vr, gold[name + "_vr"], equal_nan=True)


@pytest.mark.gpu
def test_kernel_matches_reference(gold):
    from pycwr_b200 import wind
    for name, seed, naz, nrange, elev, az_num, bin_num, kw in CASES:
        ref = {k: gold["%s_%s" % (name, k)] for k in vo.F64_KEYS + vo.I32_KEYS}
        out = wind._run_local_vvp(gold[name + "_az"], gold[name + "_el"], gold[name + "_vr"], az_num, bin_num, **kw)
        check(out, ref, 1e-5, 1e-6, kw.get("max_condition_number", 1e9))
        m = np.isfinite(ref["u"])
        assert np.max(np.abs(out["u"][m] - ref["u"][m])) < 1e-8      # observed ~1e-12: report drift early


@pytest.mark.gpu
def test_kernel_default_window_against_oracle():
    """The reference's default 91 x 9 window on a 360-ray sweep (a slice of gates: the oracle is a Python loop)."""
    from pycwr_b200 import wind
    az, el, vr = synthetic_sweep(21, 360, 26, 1.45, outliers=0.06, missing=0.2)
    out = wind._run_local_vvp(az, el, vr, 91, 9)
    ref = vo.run_local_vvp(az[:], el, vr, 91, 9)
    check(out, ref, 1e-5, 1e-6, 1e9)
    assert np.isfinite(out["u"]).sum() > 3000


@pytest.mark.gpu
def test_kernel_edge_cases():
    from pycwr_b200 import wind
    az, el, vr = synthetic_sweep(8, 36, 12, 2.0)
    with pytest.raises(ValueError, match="az_num must be odd"):
        wind._run_local_vvp(az, el, vr, 4, 3)
    with pytest.raises(ValueError, match="bin_num must be odd"):
        wind._run_local_vvp(az, el, vr, 5, 2)
    with pytest.raises(ValueError, match="azimuth length"):
        wind._run_local_vvp(az[:-1], el[:-1], vr, 5, 3)
    # bin window wider than the sweep: nothing is analysed, statistics stay NaN / 0 like the reference
    out = wind._run_local_vvp(az, el, vr, 5, 13)
    assert np.isnan(out["valid_fraction"]).all() and not out["valid_count"].any()
    # all missing
    out = wind._run_local_vvp(az, el, np.full_like(vr, -999.0), 5, 3, min_valid_count=0, min_valid_fraction=0.0,
                              min_azimuth_sector_count=0)
    ref = vo.run_local_vvp(az, el, np.full_like(vr, -999.0), 5, 3, min_valid_count=0, min_valid_fraction=0.0,
                           min_azimuth_sector_count=0)
    check(out, ref, 1e-5, 1e-6, 1e9)
    # noise-free uniform wind (MAD ~ 1e-15: the Huber loop stops at once), non-finite azimuth on one ray,
    # scalar elevation
    az2 = az.copy(); az2[7] = np.nan
    vr2 = np.repeat(((5.0 * np.sin(np.deg2rad(az)) - 3.0 * np.cos(np.deg2rad(az))) * np.cos(np.deg2rad(2.0)))[:, None], 12, 1)
    out = wind._run_local_vvp(az2, 2.0, vr2, 9, 3, min_valid_count=3)
    ref = vo.run_local_vvp(az2, 2.0, vr2, 9, 3, min_valid_count=3)
    check(out, ref, 1e-5, 1e-6, 1e9)
