"""Pin the CPU oracle (oracle/grid_oracle.c): against the reference's own known answers, the
committed fixtures generated from the compiled reference, and — where oracle/_ref exists — the
compiled reference itself on fresh random inputs.  Bit-exact (integer/index bar) on values."""
import numpy as np
import pytest

from conftest import BLINDS, FILL, golden_volume
from oracle import oracle as O
from pycwr_b200 import synth


def _two_sweep(values):
    az = np.array([0.0, 90.0, 180.0, 270.0])
    rg = np.array([1000.0, 2000.0])
    return ([az.copy(), az.copy()], [rg.copy(), rg.copy()], np.array([1.0, 2.0]),
            [np.full((4, 2), values[0]), np.full((4, 2), values[1])])


def _target(h, el=1.5):
    x, y, z = O.antenna_to_cartesian(1500.0, 0.0, el, h)
    return np.array([[x]]), np.array([[y]]), np.array([z])


def test_known_answer_cappi3d_15():
    # reference test/test_regression_grid.py:55-75
    va, vr, fe, vv = _two_sweep((10.0, 20.0))
    gx, gy, lv = _target(100.0)
    out = O.get_CAPPI_3d(va, vr, fe, vv, 100.0, gx, gy, lv, FILL)
    assert np.isclose(out[0, 0, 0], 15.0, rtol=0.0, atol=1e-6)


def test_known_answer_mosaic_35():
    # reference test/test_regression_grid.py:77-109
    gx, gy, lv = _target(100.0)
    vols = [_two_sweep((10.0, 20.0)), _two_sweep((30.0, 40.0))]
    out = O.get_mosaic_CAPPI_3d([v[0] for v in vols], [v[1] for v in vols], [v[2] for v in vols],
                                [v[3] for v in vols], np.zeros(2), np.zeros(2), np.full(2, 100.0),
                                gx, gy, lv, FILL)
    assert np.isclose(out[0, 0, 0], 35.0, rtol=0.0, atol=1e-6)


def test_known_answer_single_sweep_10():
    # reference test/test_regression_grid.py:111-160
    az = np.array([0.0, 90.0, 180.0, 270.0])
    rg = np.array([1000.0, 2000.0])
    x, y, _ = O.antenna_to_cartesian(1500.0, 0.0, 1.0, 100.0)
    out = O.get_CAPPI_xy([az], [rg], np.array([1.0]), [np.full((4, 2), 10.0)], 100.0,
                         np.array([[x]]), np.array([[y]]), 126.0, FILL)
    assert np.isclose(out[0, 0], 10.0, rtol=0.0, atol=1e-6)


def test_known_answer_blind_zone():
    # reference test/test_examples_interp.py:537-583: x=500 m inside the first gate, level 0
    az = np.array([0.0, 90.0, 180.0, 270.0])
    rg = np.array([1000.0, 2000.0])
    va, vr, fe = [az, az], [rg, rg], np.array([1.0, 2.0])
    vv = [np.full((4, 2), 10.0), np.full((4, 2), 20.0)]
    gx, gy = np.array([[500.0]]), np.array([[0.0]])
    m = O.get_CAPPI_3d(va, vr, fe, vv, 0.0, gx, gy, np.array([0.0]), FILL, None, "mask")
    h = O.get_CAPPI_3d(va, vr, fe, vv, 0.0, gx, gy, np.array([0.0]), FILL, None, "hybrid")
    assert m[0, 0, 0] == FILL
    assert np.isclose(h[0, 0, 0], 10.0, atol=1e-6)


def test_known_answer_geometry():
    # reference test/test_regression_geometry.py:205-220 and :152-203 (survey probe values)
    for (x, y), want in zip([(0, 1), (1, 0), (0, -1), (-1, 0), (-1, -1)], [0.0, 90.0, 180.0, 270.0, 225.0]):
        assert abs(O.xy_to_azimuth(float(x), float(y)) - want) < 1e-12
    az, r, z = O.xye_to_antenna(12345.0, -5432.0, 1.2, 150.0)
    assert (az, r, z) == (113.75027217143396, 13490.897084932862, 443.2396382074803)
    rng = np.random.default_rng(20260319)           # :11-25 forward/inverse round trip
    for _ in range(256):
        r0, a0, e0 = rng.uniform(500, 2e5), rng.uniform(0, 360), rng.uniform(0.3, 20)
        x, y, z = O.antenna_to_cartesian(r0, a0, e0, 150.0)
        a1, r1, e1 = O.cartesian_to_antenna(x, y, z, 150.0)
        assert np.allclose([a1, r1, e1], [a0, r0, e0], rtol=1e-7)


@pytest.mark.parametrize("tag", ["uniform", "ragged"])
def test_golden_case_a(golden, tag):
    pre = "A_%s_" % tag
    va, vr, fe, vv = golden_volume(golden, pre)
    h = float(golden[pre + "h"])
    gx, gy, lv = golden[pre + "gx"], golden[pre + "gy"], golden[pre + "levels"]
    for b in BLINDS:
        got = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b)
        assert np.array_equal(got, golden[pre + "cappi3d_" + b]), b
    got = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, 8.1e6, "hybrid")
    assert np.array_equal(got, golden[pre + "cappi3d_re8p1e6"])
    assert np.array_equal(O.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL), golden[pre + "cr"])
    assert np.array_equal(O.ppi_to_grid(va[2], vr[2], fe[2], vv[2], h, gx, gy, FILL, None, "hybrid"),
                          golden[pre + "ppi2_hybrid"])
    assert np.array_equal(O.ppi_to_grid(va[0], vr[0], fe[0], vv[0], h, gx, gy, FILL, None, "mask"),
                          golden[pre + "ppi0_mask"])
    assert np.array_equal(O.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL, None,
                                         "nearest_gate", 3.0), golden[pre + "single_bw3"])
    assert np.array_equal(O.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL, None,
                                         "mask", 0.0), golden[pre + "single_bw0"])


def test_golden_case_b_mosaic(golden):
    vols = [golden_volume(golden, "B_", radar=i) for i in range(3)]
    args = ([v[0] for v in vols], [v[1] for v in vols], [v[2] for v in vols], [v[3] for v in vols],
            golden["B_rx"], golden["B_ry"], golden["B_rh"], golden["B_gx"], golden["B_gy"], golden["B_levels"], FILL)
    assert np.array_equal(O.get_mosaic_CAPPI_3d(*args), golden["B_mosaic"])
    assert np.array_equal(O.get_mosaic_CAPPI_3d(*args, [8.0e6, None, 8.6e6]), golden["B_mosaic_re"])


def test_golden_case_c_axis_ties(golden):
    va, vr, fe, vv = golden_volume(golden, "C_")
    gx, gy, lv = golden["C_gx"], golden["C_gy"], golden["C_levels"]
    assert np.array_equal(O.get_CAPPI_3d(va, vr, fe, vv, 100.0, gx, gy, lv, FILL, None, "hybrid"),
                          golden["C_cappi3d_hybrid"])
    assert np.array_equal(O.get_CR_xy(va, vr, fe, vv, 100.0, gx, gy, FILL), golden["C_cr"])


def test_golden_case_d_scalars(golden):
    pts = golden["D_pts"]
    assert np.array_equal([O.xy_to_azimuth(x, y) for x, y in pts], golden["D_xy_to_azimuth"])
    assert np.array_equal([O.xye_to_antenna(x, y, 1.2, 150.0) for x, y in pts], golden["D_xye"])
    assert np.array_equal([O.xye_to_antenna(x, y, 1.2, 150.0, 8.1e6) for x, y in pts], golden["D_xye_re"])
    assert np.array_equal([O.cartesian_to_antenna(x, y, 3000.0, 150.0) for x, y in pts], golden["D_c2a"])


def test_index_records_consistent_with_values():
    """The index records describe the value path: state/flags imply fill exactly where they must."""
    vol = synth.make_volume(5, naz=60, ngate=120, gate_m=1000.0)
    va, vr, fe, vv, h = synth.vol_tuple(vol)
    rng = np.random.default_rng(3)
    gx, gy = rng.uniform(-130e3, 130e3, (30, 30)), rng.uniform(-130e3, 130e3, (30, 30))
    lv = np.array([500.0, 2000.0, 7000.0])
    out, idx = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "mask", return_index=True)
    state = idx[..., 0]
    assert np.all(out[state <= 2] == FILL)
    pair = state == 4
    assert np.all(idx[..., 2][pair] == idx[..., 1][pair] + 1)
    ok = pair & (idx[..., 3] >= 0)
    assert np.all((idx[..., 4][ok] >= 1) & (idx[..., 4][ok] < 120))
    assert np.all((idx[..., 3][ok] >= 0) & (idx[..., 3][ok] < 60))


def test_oracle_equals_compiled_reference_random(ref_module):
    if ref_module is None:
        pytest.skip("compiled reference (oracle/_ref) not present on this machine")
    R = ref_module
    for seed in range(3):
        vol = synth.make_volume(100 + seed, naz=72 + seed, ngate=150, gate_m=900.0)
        va, vr, fe, vv, h = synth.vol_tuple(vol)
        rng = np.random.default_rng(seed)
        gx, gy = rng.uniform(-140e3, 140e3, (40, 33)), rng.uniform(-140e3, 140e3, (40, 33))
        lv = np.array([300.0, 1500.0, 4000.0, 11000.0])
        for b in BLINDS:
            assert np.array_equal(R.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b),
                                  O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b))
        assert np.array_equal(R.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL),
                              O.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL))
        assert np.array_equal(R.ppi_to_grid(va[1], vr[1], fe[1], vv[1], h, gx, gy, FILL, None, "nearest_gate"),
                              O.ppi_to_grid(va[1], vr[1], fe[1], vv[1], h, gx, gy, FILL, None, "nearest_gate"))


def test_oracle_equals_compiled_reference_many_sweeps(ref_module):
    """The many-sweep regime (phased-array scans, BASELINE config C5: sweeps 0.9 degrees apart up to 57.6 degrees,
    30 m gates, levels every 250 m, targets beyond the last gate): the oracle is the checker of the GPU kernels that
    serve volumes with more than 16 sweeps, so it is pinned to the compiled reference there as well."""
    if ref_module is None:
        pytest.skip("compiled reference (oracle/_ref) not present on this machine")
    R = ref_module
    for seed, ne in ((7, 24), (8, 64)):
        vol = synth.make_volume(300 + seed, elevations=np.linspace(0.9, 0.9 * ne, ne), naz=90 + seed, ngate=400,
                                gate_m=30.0)
        va, vr, fe, vv, h = synth.vol_tuple(vol)
        rng = np.random.default_rng(seed)
        gx, gy = rng.uniform(-13e3, 13e3, (31, 29)), rng.uniform(-13e3, 13e3, (31, 29))   # some beyond 12 km
        lv = np.arange(1, 41) * 250.0
        for b in BLINDS:
            assert np.array_equal(R.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b),
                                  O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b)), (ne, b)
        assert np.array_equal(R.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL),
                              O.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL)), ne
