"""Vertical sections / pseudo-RHI: oracle pinning on the CPU, kernel parity on the GPU.

Golden values come from the reference's own ``PRD._interpolate_ppi_strip_arrays``,
``PRD._build_section_line`` and ``cartesian_to_antenna_cwr`` (tests/golden/make_section_golden.py).
Tolerances: interpolation with explicit targets is bit-exact; with the fused geometry the targets differ
by the last ulps of cos / sin / atan2 (CUDA vs NumPy), so azimuth / range / height are compared to
1e-12 relative and the interpolated moment to 1e-6 dB (far inside the 1e-3 dB bar of the gridding path).
"""
import os

import numpy as np
import pytest

from oracle import section_oracle as so

GOLD = os.path.join(os.path.dirname(__file__), "golden", "section_reference.npz")
NLINE, NRHI = 4, 6


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    ne = len(g["elev"])
    sweeps = [(g["az_%d" % s], g["rg_%d" % s], g["v_%d" % s]) for s in range(ne)]
    return g, sweeps


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


# ---- CPU: the oracle against the reference's outputs --------------------------------------------

def test_oracle_strip_interp_matches_reference_bit_for_bit(gold):
    g, sweeps = gold
    with np.errstate(invalid="ignore"):
        for s, sw in enumerate(sweeps):
            out = so.strip_interp(*sw, g["t_az_%d" % s], g["t_rg_%d" % s])
            assert same(out, g["t_val_%d" % s])
            assert np.isfinite(out).sum() > 300


def test_oracle_section_and_rhi_match_reference(gold):
    g, sweeps = gold
    for li in range(NLINE):
        ln = g["line_%d" % li]
        x, y, d = so.section_line(ln[:2], ln[2:4], ln[4])
        assert same(x, g["x_%d" % li]) and same(y, g["y_%d" % li]) and same(d, g["d_%d" % li])
        for name, re in (("def", None), ("alt", 7.1e6)):
            v, a, r, z = so.section_rows(sweeps, g["elev"], x, y, float(g["h"]), re)
            assert same(v, g["sec_%s_val_%d" % (name, li)])
            assert same(a, g["sec_%s_az_%d" % (name, li)])
            assert same(r, g["sec_%s_rg_%d" % (name, li)])
            assert same(z, g["sec_%s_z_%d" % (name, li)])
    for ai in range(NRHI):
        assert same(so.rhi_rows(sweeps, float(g["rhi_az_%d" % ai])), g["rhi_val_%d" % ai])


def test_host_line_builder_and_errors(gold):
    from pycwr_b200 import section
    g, _ = gold
    for li in range(NLINE):
        ln = g["line_%d" % li]
        x, y, d = section._build_section_line(ln[:2], ln[2:4], ln[4])
        assert same(x, g["x_%d" % li]) and same(y, g["y_%d" % li]) and same(d, g["d_%d" % li])
    with pytest.raises(ValueError, match="must not be identical"):
        section._build_section_line((1.0, 2.0), (1.0, 2.0), 100.0)
    with pytest.raises(ValueError, match="2-element"):
        section._build_section_line((1.0, 2.0, 3.0), (1.0, 2.0), 100.0)
    assert same(section._pad_section_rows([[1.0], [1.0, 2.0, 3.0]]), np.array([[1.0, np.nan, np.nan], [1.0, 2.0, 3.0]]))
    az, val = section._section_table(np.array([-2.0, 10.0, 10.0, 200.0]), np.arange(8.0).reshape(4, 2))
    assert same(az, np.array([10.0, 200.0, 358.0])) and same(val, np.array([[2.0, 3.0], [6.0, 7.0], [0.0, 1.0]]))


# ---- GPU: the kernel through the C ABI ----------------------------------------------------------

@pytest.mark.gpu
def test_strip_interp_explicit_targets_bit_exact(gold):
    from pycwr_b200 import section
    g, sweeps = gold
    for s, sw in enumerate(sweeps):
        out = section._interpolate_ppi_strip_arrays(*sw, g["t_az_%d" % s], g["t_rg_%d" % s])
        assert same(out, g["t_val_%d" % s])


@pytest.mark.gpu
def test_section_lines_match_reference(gold):
    from pycwr_b200 import section
    g, sweeps = gold
    sv = section.SectionVolume(sweeps, g["elev"])
    for li in range(NLINE):
        for name, re in (("def", None), ("alt", 7.1e6)):
            v, a, r, z = sv.sample_line(g["x_%d" % li], g["y_%d" % li], float(g["h"]), re)
            ref_v = g["sec_%s_val_%d" % (name, li)]
            np.testing.assert_allclose(a, g["sec_%s_az_%d" % (name, li)], rtol=1e-12, atol=1e-10)
            np.testing.assert_allclose(r, g["sec_%s_rg_%d" % (name, li)], rtol=1e-12, atol=1e-6)
            np.testing.assert_allclose(z, g["sec_%s_z_%d" % (name, li)], rtol=1e-12, atol=1e-6)
            assert same(np.isnan(v), np.isnan(ref_v))
            np.testing.assert_allclose(v, ref_v, rtol=0, atol=1e-6, equal_nan=True)
    sv.close()


@pytest.mark.gpu
def test_section_with_reference_targets_bit_exact(gold):
    """Feeding the reference's own (azimuth, range) targets removes the libm difference: bit-exact."""
    from pycwr_b200 import section
    g, sweeps = gold
    sv = section.SectionVolume(sweeps, g["elev"])
    for li in range(NLINE):
        out = sv.sample_targets(g["sec_def_az_%d" % li], g["sec_def_rg_%d" % li])
        assert same(out, g["sec_def_val_%d" % li])
    sv.close()


@pytest.mark.gpu
def test_rhi_cut_bit_exact(gold):
    from pycwr_b200 import section
    g, sweeps = gold
    sv = section.SectionVolume(sweeps, g["elev"])
    for ai in range(NRHI):
        assert same(sv.sample_rhi(float(g["rhi_az_%d" % ai])), g["rhi_val_%d" % ai])
    sv.close()


@pytest.mark.gpu
def test_prd_level_section_and_rhi(gold):
    from pycwr_b200 import section
    from pycwr_b200.products import ArrayPRD
    g, sweeps = gold
    rng = np.random.default_rng(3)
    order = rng.permutation(len(sweeps))                  # scan order != elevation order, rays shuffled
    scan = []
    for s in order:
        az, rg, v = sweeps[int(s)]
        if s == 0:                                        # the duplicated ray is order dependent: drop it
            keep = np.ones(az.size, bool); keep[10] = False
            az, v = az[keep], v[keep]
        p = rng.permutation(az.size)
        scan.append((az[p], rg, v[p]))
    prd = ArrayPRD({"dBZ": scan}, g["elev"][order], float(g["h"]), lon=118.0, lat=32.0)
    ds = section.extract_section(prd, (-31.0, -12.0), (35.5, 27.1), sample_spacing=300.0)
    ref = g["sec_def_val_0"]
    got = ds["dBZ"].values
    assert got.shape == ref.shape and list(ds["source_sweep"].values) == list(np.argsort(g["elev"][order]))
    rows = [r for r in range(len(sweeps)) if r != 0]
    np.testing.assert_allclose(got[rows], ref[rows], rtol=0, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(ds["z"].values, g["sec_def_z_0"], rtol=1e-12, atol=1e-6)
    assert same(ds.coords["distance"].values, g["d_0"])
    assert ds.attrs["section_type"] == "section" and ds.attrs["point_units"] == "meters"
    rhi = section.extract_rhi(prd, 47.3)
    assert same(rhi["dBZ"].values[rows], g["rhi_val_1"][rows])
    assert rhi.attrs["target_azimuth"] == 47.3 and rhi["z"].values.shape == g["rhi_val_1"].shape
    with pytest.raises(ValueError, match="linear"):
        section.extract_section(prd, (0, 0), (1, 1), interpolation="cubic")
    with pytest.raises(ValueError, match="point_units"):
        section.extract_section(prd, (0, 0), (1, 1), point_units="mi")
    with pytest.raises(KeyError):
        section.extract_section(prd, (0, 0), (1, 1), field_name="V")
    ll = section.extract_section_lonlat(prd, (117.8, 31.9), (118.3, 32.2), sample_spacing=500.0)
    assert ll["dBZ"].values.shape[0] == len(sweeps) and np.isfinite(ll["dBZ"].values).any()


@pytest.mark.gpu
def test_section_degenerate_inputs():
    from pycwr_b200 import section, _lib
    az = np.array([10.0, 100.0, 250.0]); rg = np.array([1000.0, 2000.0]); v = np.arange(6.0).reshape(3, 2)
    nan = section._interpolate_ppi_strip_arrays(az[:1], rg, v[:1], [10.0], [1500.0])      # one ray
    assert np.isnan(nan).all()
    nan = section._interpolate_ppi_strip_arrays(az, rg[:1], v[:, :1], [10.0], [1000.0])  # one gate
    assert np.isnan(nan).all()
    assert section._interpolate_ppi_strip_arrays(az, rg, v, [], []).shape == (0,)
    out = section._interpolate_ppi_strip_arrays(az, rg, v, [[10.0, 300.0]], [[1000.0, 2000.0]])
    assert out.shape == (1, 2) and out[0, 0] == 0.0
    # wrap-around: 300 deg lies between the last ray (250) and the first + 360 (370)
    w = (370.0 - 300.0) / 120.0
    assert out[0, 1] == pytest.approx(w * 5.0 + (1 - w) * 1.0, abs=1e-12)
    sv = section.SectionVolume([(az, rg, v)], [1.0])
    with pytest.raises(ValueError):
        sv.sample_line([0.0], [0.0, 1.0], 0.0)
    with pytest.raises(ValueError, match="positive"):
        sv.sample_line([0.0], [0.0], 0.0, effective_earth_radius=-1.0)
    sv.close()
    # the packed float32 layout has already merged its ray pairs: refused, never silently resampled
    from pycwr_b200.runtime import DeviceVolume
    az2 = [np.linspace(0, 359, 360)] * 2; rg2 = [np.arange(100) * 100.0 + 100.0] * 2
    dv = DeviceVolume(az2, rg2, [0.5, 1.5], [np.zeros((360, 100))] * 2)
    if dv.value_dtype == "f32":
        import ctypes
        o = np.zeros(2)
        rc = _lib.load().pcg_section_sample(dv.handle, 0, 1, None, None, 0.0, 1.0, o.ctypes.data_as(ctypes.c_void_p),
                                            o.ctypes.data_as(ctypes.c_void_p), 0, o.ctypes.data_as(ctypes.c_void_p),
                                            None, None, None)
        assert rc == -1 and b"float64" in _lib.load().pcg_last_error()
    dv.close()
