"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the
committed fixtures generated from the compiled reference.

Bars (BASELINE.json north_star): interpolation indices and valid/fill masks BIT-EXACT; moment
values within 1e-3 dB absolute (dBZ/ZDR) and 1e-5 relative (velocity).  With float64 moment
storage the value path is the reference's own unfused arithmetic, so the tests additionally
report (and bound) the bit-identical fraction.
"""
import numpy as np
import pytest

from conftest import BLINDS, FILL, golden_volume
from oracle import oracle as O
from pycwr_b200 import RadarGridC as G
from pycwr_b200 import synth
from pycwr_b200.runtime import DeviceVolume

pytestmark = pytest.mark.gpu

ATOL_DB = 1e-3
RTOL_V = 1e-5


@pytest.fixture(params=["f32", "f64"], autouse=True)
def storage_mode(request):
    """Every parity test runs on both device layouts: the production float32-pair layout (exact
    indices, FP32 blending) and the float64 layout (the reference's arithmetic)."""
    old = G.VALUE_DTYPE
    G.VALUE_DTYPE = request.param
    yield request.param
    G.VALUE_DTYPE = old


def assert_parity(got, want, idx_got=None, idx_want=None, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    assert np.array_equal(got == FILL, want == FILL), "fill mask differs: " + what
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN mask differs: " + what
    ok = (want != FILL) & ~np.isnan(want)
    if ok.any():
        err = np.abs(got[ok] - want[ok])
        assert err.max() <= ATOL_DB, "%s: max |err| %.3e" % (what, err.max())
        if G.VALUE_DTYPE == "f64":
            # reference arithmetic: relative bar element by element
            assert np.all(err <= RTOL_V * np.abs(want[ok]) + 1e-9), "%s: relative bar" % what
        else:
            # FP32 blending: 1e-5 relative to the magnitude of the field (an interpolated value can
            # be arbitrarily close to zero while its neighbours are not)
            assert err.max() <= RTOL_V * max(np.abs(want[ok]).max(), 1.0), "%s: relative bar" % what
    if idx_got is not None:
        assert np.array_equal(idx_got, idx_want), "interpolation indices differ: " + what


def test_known_answers_on_gpu():
    # reference test/test_regression_grid.py:55-160, test/test_examples_interp.py:537-583
    az = np.array([0.0, 90.0, 180.0, 270.0])
    rg = np.array([1000.0, 2000.0])
    x, y, z = G.antenna_to_cartesian(1500.0, 0.0, 1.5, 100.0)
    gx, gy, lv = np.array([[x]]), np.array([[y]]), np.array([z])
    vol = ([az, az], [rg, rg], np.array([1.0, 2.0]), [np.full((4, 2), 10.0), np.full((4, 2), 20.0)])
    out = G.get_CAPPI_3d(*vol, 100.0, gx, gy, lv, FILL)
    assert np.isclose(out[0, 0, 0], 15.0, rtol=0.0, atol=1e-6)
    vol2 = ([az, az], [rg, rg], np.array([1.0, 2.0]), [np.full((4, 2), 30.0), np.full((4, 2), 40.0)])
    mos = G.get_mosaic_CAPPI_3d([vol[0], vol2[0]], [vol[1], vol2[1]], [vol[2], vol2[2]], [vol[3], vol2[3]],
                                np.zeros(2), np.zeros(2), np.full(2, 100.0), gx, gy, lv, FILL)
    assert np.isclose(mos[0, 0, 0], 35.0, rtol=0.0, atol=1e-6)
    x1, y1, _ = G.antenna_to_cartesian(1500.0, 0.0, 1.0, 100.0)
    one = G.get_CAPPI_xy([az], [rg], np.array([1.0]), [np.full((4, 2), 10.0)], 100.0,
                         np.array([[x1]]), np.array([[y1]]), 126.0, FILL)
    assert np.isclose(one[0, 0], 10.0, rtol=0.0, atol=1e-6)
    g500, g0 = np.array([[500.0]]), np.array([[0.0]])
    m = G.get_CAPPI_3d(*vol, 0.0, g500, g0, np.array([0.0]), FILL, None, "mask")
    h = G.get_CAPPI_3d(*vol, 0.0, g500, g0, np.array([0.0]), FILL, None, "hybrid")
    assert m[0, 0, 0] == FILL and np.isclose(h[0, 0, 0], 10.0, atol=1e-6)


@pytest.mark.parametrize("tag", ["uniform", "ragged"])
def test_golden_case_a(golden, tag):
    pre = "A_%s_" % tag
    va, vr, fe, vv = golden_volume(golden, pre)
    h = float(golden[pre + "h"])
    gx, gy, lv = golden[pre + "gx"], golden[pre + "gy"], golden[pre + "levels"]
    for b in BLINDS:
        got, gi = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b, return_index=True)
        _, wi = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b, return_index=True)
        assert_parity(got, golden[pre + "cappi3d_" + b], gi, wi, "cappi3d " + b)
    assert_parity(G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, 8.1e6, "hybrid"),
                  golden[pre + "cappi3d_re8p1e6"], what="custom Re")
    assert_parity(G.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL), golden[pre + "cr"], what="CR")
    got, gi = G.ppi_to_grid(va[2], vr[2], fe[2], vv[2], h, gx, gy, FILL, None, "hybrid", return_index=True)
    _, wi = O.ppi_to_grid(va[2], vr[2], fe[2], vv[2], h, gx, gy, FILL, None, "hybrid", return_index=True)
    assert_parity(got, golden[pre + "ppi2_hybrid"], gi, wi, "ppi hybrid")
    assert_parity(G.ppi_to_grid(va[0], vr[0], fe[0], vv[0], h, gx, gy, FILL, None, "mask"),
                  golden[pre + "ppi0_mask"], what="ppi mask")
    assert_parity(G.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL, None,
                                 "nearest_gate", 3.0), golden[pre + "single_bw3"], what="single bw3")
    assert_parity(G.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL, None, "mask", 0.0),
                  golden[pre + "single_bw0"], what="single bw0")


def test_golden_case_b_mosaic(golden):
    vols = [golden_volume(golden, "B_", radar=i) for i in range(3)]
    args = ([v[0] for v in vols], [v[1] for v in vols], [v[2] for v in vols], [v[3] for v in vols],
            golden["B_rx"], golden["B_ry"], golden["B_rh"], golden["B_gx"], golden["B_gy"], golden["B_levels"], FILL)
    assert_parity(G.get_mosaic_CAPPI_3d(*args), golden["B_mosaic"], what="mosaic")
    assert_parity(G.get_mosaic_CAPPI_3d(*args, [8.0e6, None, 8.6e6]), golden["B_mosaic_re"], what="mosaic Re")


def test_golden_case_c_axis_and_diagonal_ties(golden):
    """Regular grid through the radar with round-number azimuth/range tables: whole rows sit
    exactly on table entries, so any last-bit difference in az flips an index."""
    va, vr, fe, vv = golden_volume(golden, "C_")
    gx, gy, lv = golden["C_gx"], golden["C_gy"], golden["C_levels"]
    got, gi = G.get_CAPPI_3d(va, vr, fe, vv, 100.0, gx, gy, lv, FILL, None, "hybrid", return_index=True)
    _, wi = O.get_CAPPI_3d(va, vr, fe, vv, 100.0, gx, gy, lv, FILL, None, "hybrid", return_index=True)
    assert_parity(got, golden["C_cappi3d_hybrid"], gi, wi, "axis ties")
    assert_parity(G.get_CR_xy(va, vr, fe, vv, 100.0, gx, gy, FILL), golden["C_cr"], what="axis ties CR")


@pytest.mark.parametrize("cid,scale", [(1, 0.5), (2, 0.12), (3, 0.06), (5, 0.08)])
def test_configs_against_oracle(cid, scale):
    """BASELINE configs (grid thinned so the oracle finishes in seconds), stress moments."""
    cfg = synth.config(cid, style="stress", scale=scale)
    vol = cfg["volume"]
    gx, gy = cfg["GridX"], cfg["GridY"]
    for field in cfg["fields"]:
        va, vr, fe, vv, h = synth.vol_tuple(vol, field)
        if cfg["kind"] == "cr":
            assert_parity(G.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL),
                          O.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL), what=cfg["name"])
            continue
        for blind in ("mask", "hybrid"):
            got, gi = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, cfg["levels"], FILL, None, blind, return_index=True)
            want, wi = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, cfg["levels"], FILL, None, blind, return_index=True)
            assert_parity(got, want, gi, wi, "%s %s %s" % (cfg["name"], field, blind))
            same = np.mean(got == want)
            print("%s %s %s [%s]: bit-identical value fraction %.6f, max |err| %.3e" % (
                cfg["name"], field, blind, G.VALUE_DTYPE, same, np.abs(got - want).max()))
            if G.VALUE_DTYPE == "f64":
                assert same > 0.5, "bit-identical value fraction %.6f" % same


def test_multifield_volume_shares_indices():
    cfg = synth.config(2, style="stress", scale=0.08)
    vol = cfg["volume"]
    va, vr, fe = vol["vol_azimuth"], vol["vol_range"], vol["fix_elevation"]
    fields = [vol["vol_value"][f] for f in cfg["fields"]]
    dv = DeviceVolume(va, vr, fe, fields, value_dtype=G.VALUE_DTYPE)
    out = G.get_CAPPI_3d(None, None, None, dv, vol["radar_height"], cfg["GridX"], cfg["GridY"], cfg["levels"], FILL)
    assert out.shape == (3,) + (cfg["levels"].size,) + cfg["GridX"].shape
    for k, f in enumerate(cfg["fields"]):
        want = O.get_CAPPI_3d(va, vr, fe, vol["vol_value"][f], vol["radar_height"], cfg["GridX"], cfg["GridY"],
                              cfg["levels"], FILL)
        assert_parity(out[k], want, what="multi-field " + f)
    dv.close()


def test_packed_layout_falls_back_when_it_cannot_serve():
    az = synth.azimuth_table(np.random.default_rng(1), 12)
    rg = np.arange(1, 9) * 1000.0
    val = np.random.default_rng(2).uniform(0, 50, (12, 8))
    one = DeviceVolume([az], [rg], np.array([1.0]), [val], value_dtype="f32")
    assert one.value_dtype == "f64"                       # single sweep
    dup = DeviceVolume([az, az], [np.array([1e3, 1e3, 2e3, 3e3, 4e3, 5e3, 6e3, 7e3]), rg], np.array([1.0, 2.0]),
                       [val, val], value_dtype="f32")
    assert dup.value_dtype == "f64"                       # duplicate range entries
    ok = DeviceVolume([az, az], [rg, rg], np.array([1.0, 2.0]), [val, val], value_dtype="f32")
    assert ok.value_dtype == "f32"
    g = np.random.default_rng(3).uniform(-9e3, 9e3, (5, 7))
    with pytest.raises(ValueError, match="fillvalue"):
        G.get_CAPPI_3d(None, None, None, ok, 10.0, g, g, np.array([100.0]), -9999.0)
    for v in (one, dup, ok):
        v.close()


def test_edge_cases():
    az = synth.azimuth_table(np.random.default_rng(1), 12)
    rg = np.arange(1, 9) * 1000.0
    val = np.random.default_rng(2).uniform(0, 50, (12, 8))
    g = np.random.default_rng(3).uniform(-9e3, 9e3, (5, 7))
    lv = np.array([100.0, 900.0])
    # empty volume -> all fill (RadarGridC.pyx:398-399)
    out = G.get_CAPPI_3d([], [], np.zeros(0), [], 10.0, g, g.T.copy().T, lv, FILL)
    assert out.shape == (2, 5, 7) and np.all(out == FILL)
    assert np.all(G.get_CR_xy([], [], np.zeros(0), [], 10.0, g, g, FILL) == FILL)
    # degenerate sweeps (naz < 2 or ng < 2) -> fill (RadarGridC.pyx:273-274,337-339)
    assert np.all(G.ppi_to_grid(az[:1], rg, 1.0, val[:1], 10.0, g, g, FILL) == FILL)
    assert np.all(G.ppi_to_grid(az, rg[:1], 1.0, val[:, :1], 10.0, g, g, FILL) == FILL)
    got = G.get_CAPPI_3d([az[:1], az], [rg, rg], np.array([1.0, 8.0]), [val[:1], val], 10.0, g, g, lv, FILL)
    want = O.get_CAPPI_3d([az[:1], az], [rg, rg], np.array([1.0, 8.0]), [val[:1], val], 10.0, g, g, lv, FILL)
    assert_parity(got, want, what="degenerate lower sweep")
    # empty grid
    e = np.zeros((0, 4))
    assert G.get_CAPPI_3d([az], [rg], np.array([1.0]), [val], 10.0, e, e, lv, FILL).shape == (2, 0, 4)
    # no levels
    assert G.get_CAPPI_3d([az], [rg], np.array([1.0]), [val], 10.0, g, g, np.zeros(0), FILL).shape == (0, 5, 7)
    # all-fill moments and a grid point exactly at the radar
    g0 = np.zeros((1, 1))
    allfill = np.full((12, 8), FILL)
    assert G.get_CAPPI_3d([az, az], [rg, rg], np.array([1.0, 8.0]), [allfill, allfill], 10.0, g, g, lv, FILL).max() == FILL
    got = G.get_CAPPI_3d([az, az], [rg, rg], np.array([1.0, 8.0]), [val, val], 10.0, g0, g0, np.array([10.0, 50.0]), FILL, None, "nearest_valid")
    want = O.get_CAPPI_3d([az, az], [rg, rg], np.array([1.0, 8.0]), [val, val], 10.0, g0, g0, np.array([10.0, 50.0]), FILL, None, "nearest_valid")
    assert_parity(got, want, what="origin")
    # more than 64 levels goes through several launches
    lv80 = np.linspace(50.0, 4000.0, 80)
    got = G.get_CAPPI_3d([az, az], [rg, rg], np.array([1.0, 8.0]), [val, val], 10.0, g, g, lv80, FILL, None, "hybrid")
    want = O.get_CAPPI_3d([az, az], [rg, rg], np.array([1.0, 8.0]), [val, val], 10.0, g, g, lv80, FILL, None, "hybrid")
    assert_parity(got, want, what="80 levels")


def test_packed_volume_with_degenerate_and_ragged_sweeps():
    """Sweeps with a single gate / a single ray and per-sweep range tables inside a packed volume:
    the r2 path must route them to the reference chain (no table is indexed for a degenerate sweep)."""
    rng = np.random.default_rng(5)
    az = [synth.azimuth_table(rng, n) for n in (24, 1, 30, 24, 28)]
    rg = [np.arange(1, 41) * 500.0, np.arange(1, 31) * 500.0, np.array([700.0]), np.arange(1, 36) * 600.0,
          np.arange(1, 41) * 500.0]
    fe = np.array([0.5, 1.5, 3.0, 6.0, 12.0])
    vv = [synth.stress_moment(rng, a.size, r.size, -5.0, 70.0) for a, r in zip(az, rg)]
    g = rng.uniform(-22e3, 22e3, (23, 2, 19)).transpose(1, 0, 2)
    gx, gy = np.ascontiguousarray(g[0]), np.ascontiguousarray(g[1])
    lv = np.linspace(100.0, 6000.0, 13)
    for blind in ("mask", "hybrid", "nearest_valid"):
        got, gi = G.get_CAPPI_3d(az, rg, fe, vv, 35.0, gx, gy, lv, FILL, None, blind, return_index=True)
        want, wi = O.get_CAPPI_3d(az, rg, fe, vv, 35.0, gx, gy, lv, FILL, None, blind, return_index=True)
        assert_parity(got, want, gi, wi, "degenerate/ragged " + blind)
    assert_parity(G.get_CR_xy(az, rg, fe, vv, 35.0, gx, gy, FILL), O.get_CR_xy(az, rg, fe, vv, 35.0, gx, gy, FILL),
                  what="degenerate/ragged CR")


def test_levels_below_the_antenna_and_negative_elevations():
    """Levels at / below the radar height and a negative lowest elevation (mountain radar): the r2
    thresholds are monotone only above the antenna, everything else must take the reference chain."""
    rng = np.random.default_rng(9)
    ne = 4
    az = [synth.azimuth_table(rng, 36) for _ in range(ne)]
    rg = [np.arange(1, 61) * 1000.0] * ne
    vv = [synth.stress_moment(rng, 36, 60, -5.0, 70.0) for _ in range(ne)]
    gx, gy = np.meshgrid(np.linspace(-50e3, 50e3, 37), np.linspace(-48e3, 48e3, 33), indexing="ij")
    lv = np.array([-200.0, 0.0, 1500.0, 2499.0, 2500.0, 2501.0, 3000.0, 4000.0, 7000.0])
    for fe in (np.array([-0.8, 0.0, 1.0, 3.0]), np.array([0.0, 0.7, 2.0, 5.0]), np.array([0.5, 1.5, 2.5, 4.0])):
        for blind in ("mask", "hybrid"):
            got, gi = G.get_CAPPI_3d(az, rg, fe, vv, 2500.0, gx, gy, lv, FILL, None, blind, return_index=True)
            want, wi = O.get_CAPPI_3d(az, rg, fe, vv, 2500.0, gx, gy, lv, FILL, None, blind, return_index=True)
            assert_parity(got, want, gi, wi, "antenna-level %s %s" % (fe[0], blind))


@pytest.mark.parametrize("shape", [(1, 50), (50, 1), (3, 7), (5, 3), (4, 8), (9, 33)])
def test_odd_grid_shapes(shape):
    """Every warp-footprint variant (4 x 8 patches, 1 x 32 rows) and partially filled tiles."""
    cfg = synth.config(3, style="stress", scale=0.02)
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    gx, gy = rng.uniform(-140e3, 140e3, shape), rng.uniform(-140e3, 140e3, shape)
    lv = cfg["levels"][::5]
    got, gi = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid", return_index=True)
    want, wi = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid", return_index=True)
    assert_parity(got, want, gi, wi, "grid %dx%d" % shape)
    assert_parity(G.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL), O.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL),
                  what="CR grid %dx%d" % shape)


def test_non_finite_grid_coordinates():
    """NaN / inf grid coordinates follow the reference's comparisons (every test on NaN is false)."""
    cfg = synth.config(1, style="stress", scale=0.05)
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
    gx, gy = cfg["GridX"].copy(), cfg["GridY"].copy()
    gx[2, 3] = np.nan
    gy[5, 1] = np.inf
    gx[7, 7] = -np.inf
    gy[9, 2] = np.nan
    lv = np.array([800.0, 2500.0, 6000.0])
    got = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid")
    want = O.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid")
    assert_parity(got, want, what="non-finite grid")


def _fuzz_geometry(seed, ne_lo, ne_hi, tag):
    rng = np.random.default_rng(1000 + seed)
    ne = int(rng.integers(ne_lo, ne_hi))
    top = rng.choice([6.0, 20.0, 60.0])
    fe = np.sort(rng.uniform(rng.choice([-0.5, 0.2, 0.5]), top, ne))
    fe = fe + np.arange(ne) * 1e-3                       # strictly increasing
    h = float(rng.choice([0.0, 35.0, 1800.0, 4500.0]))
    Re = float(rng.choice([8494666.6666666661, 6371000.0, 1.2e7, 5.0e7]))
    step = float(rng.choice([30.0, 250.0, 1000.0]))
    ng0 = int(rng.integers(40, 400))
    regular = bool(rng.integers(0, 2))
    shared = bool(rng.integers(0, 2))
    az, rg, vv = [], [], []
    base = (np.arange(ng0) + 1.0) * step
    if not regular:
        base = np.sort(base + rng.uniform(-0.3, 0.3, ng0) * step)
    for k in range(ne):
        naz = int(rng.integers(8, 200))
        az.append(synth.azimuth_table(rng, naz))
        r = base if shared else base[: max(4, ng0 - 7 * k)]
        rg.append(np.ascontiguousarray(r))
        vv.append(synth.stress_moment(rng, naz, r.size, -5.0, 70.0))
    ext = float(base[-1]) * 1.05
    gx = rng.uniform(-ext, ext, (33, 29))
    gy = rng.uniform(-ext, ext, (33, 29))
    gx[0, :4] = [0.0, step, -step, 0.0]                  # over the radar and on the first gate
    gy[0, :4] = [0.0, 0.0, 0.0, step]
    zmax = h + max(500.0, ext * np.tan(np.deg2rad(min(top, 45.0))))
    lv = np.unique(np.concatenate([np.sort(rng.uniform(h - 300.0, zmax, 11)), [h, h + 1e-3, h + step]]))
    for blind in ("mask", "hybrid", "nearest_valid"):
        got, gi = G.get_CAPPI_3d(az, rg, fe, vv, h, gx, gy, lv, FILL, Re, blind, return_index=True)
        want, wi = O.get_CAPPI_3d(az, rg, fe, vv, h, gx, gy, lv, FILL, Re, blind, return_index=True)
        assert_parity(got, want, gi, wi, "%s %d %s (ne=%d h=%g Re=%g step=%g regular=%s shared=%s)" % (
            tag, seed, blind, ne, h, Re, step, regular, shared))
        # the production instantiation (no index records) returns the same values bit for bit
        plain = G.get_CAPPI_3d(az, rg, fe, vv, h, gx, gy, lv, FILL, Re, blind)
        assert np.array_equal(plain, got), "%s %d %s: values differ between the kernels with and without index records" % (
            tag, seed, blind)
    assert_parity(G.get_CR_xy(az, rg, fe, vv, h, gx, gy, FILL, Re), O.get_CR_xy(az, rg, fe, vv, h, gx, gy, FILL, Re),
                  what="%s CR %d" % (tag, seed))


@pytest.mark.parametrize("seed", range(24))
def test_randomised_geometry_parity(seed):
    """Random radar heights, earth radii, elevation sets (dense and sparse, up to 60 degrees), gate
    spacings (regular and irregular, per-sweep lengths), level sets and blind modes: interpolation
    indices and masks must stay bit-exact over the whole parameter space of the r2 thresholds."""
    _fuzz_geometry(seed, 2, 14, "fuzz")


@pytest.mark.parametrize("seed", range(100, 110))
def test_many_sweep_geometry_parity(seed):
    """The same fuzz over volumes with 17 .. 47 sweeps (phased-array scans, BASELINE config C5): these take the kernel
    that keeps the azimuth brackets of the column's current sweep pair in registers instead of a [sweep] table per
    column (pcg_r3.cuh, LAZY) — regular and irregular range tables, shared and per-sweep lengths, every blind mode."""
    _fuzz_geometry(seed, 17, 48, "many-sweep")


def _many_sweep_volume(rng, ne, naz, ng, step):
    fe = np.linspace(0.9, 0.9 * ne, ne)
    rg = (np.arange(ng) + 1.0) * step
    az = [synth.azimuth_table(rng, naz) for _ in range(ne)]
    return az, [rg] * ne, fe


def test_many_sweep_mosaic_and_multifield():
    """Volumes with more than 16 sweeps through the mosaic merge (fill-aware max over radars, RadarGridC.pyx:568-592)
    and as a multi-field volume sharing one set of indices: the LAZY instantiations of cappi3d_r3_kernel with MERGE
    and with NF = 3."""
    rng = np.random.default_rng(77)
    vols, vals = [], []
    for ne in (24, 33):
        az, rg, fe = _many_sweep_volume(rng, ne, 96, 180, 150.0)
        vols.append((az, rg, fe))
        vals.append([synth.stress_moment(rng, 96, 180, -5.0, 70.0) for _ in range(ne)])
    gx, gy = np.meshgrid(np.linspace(-26e3, 26e3, 41), np.linspace(-26e3, 26e3, 37), indexing="ij")
    lv = np.arange(1, 25) * 500.0
    rx, ry, rh = np.array([0.0, 9000.0]), np.array([0.0, -4000.0]), np.array([120.0, 640.0])
    args = ([v[0] for v in vols], [v[1] for v in vols], [v[2] for v in vols], vals, rx, ry, rh, gx, gy, lv, FILL)
    assert_parity(G.get_mosaic_CAPPI_3d(*args), O.get_mosaic_CAPPI_3d(*args), what="many-sweep mosaic")
    az, rg, fe = vols[0]
    fields = [vals[0], [synth.stress_moment(rng, 96, 180, -2.0, 5.0) for _ in range(fe.size)],
              [synth.stress_moment(rng, 96, 180, -27.0, 27.0) for _ in range(fe.size)]]
    dv = DeviceVolume(az, rg, fe, fields, value_dtype=G.VALUE_DTYPE)
    out = G.get_CAPPI_3d(None, None, None, dv, 120.0, gx, gy, lv, FILL, None, "hybrid")
    for k, f in enumerate(fields):
        assert_parity(out[k], O.get_CAPPI_3d(az, rg, fe, f, 120.0, gx, gy, lv, FILL, None, "hybrid"),
                      what="many-sweep multi-field %d" % k)
    dv.close()


def test_full_size_c3_properties():
    """BASELINE config C3 at full size (1201 x 1201 x 40 from 9 x 360 x 1840): size-independent
    properties plus a scattered sample against the oracle."""
    if G.VALUE_DTYPE != "f32":
        pytest.skip("full-size run on the production layout only")
    cfg = synth.config(3, style="stress")
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
    gx, gy, lv = cfg["GridX"], cfg["GridY"], cfg["levels"]
    dv = DeviceVolume(va, vr, fe, vv, fillvalue=FILL, value_dtype="f32")
    a = np.array(G.get_CAPPI_3d(None, None, None, dv, h, gx, gy, lv, FILL, None, "hybrid"))
    b = np.array(G.get_CAPPI_3d(None, None, None, dv, h, gx, gy, lv, FILL, None, "hybrid"))
    dv.close()
    assert a.shape == (40, 1201, 1201) and np.array_equal(a, b)                  # deterministic
    valid = a != FILL
    assert 0.3 < valid.mean() < 0.99
    assert a[valid].min() >= -5.0 - 1e-3 and a[valid].max() <= 70.0 + 1e-3       # convex combinations of the moments
    # a scattered sample of columns, every level, against the oracle (indices bit-exact)
    rng = np.random.default_rng(12)
    ix, iy = rng.integers(0, 1201, 1500), rng.integers(0, 1201, 1500)
    sx, sy = np.ascontiguousarray(gx[ix, iy][:, None]), np.ascontiguousarray(gy[ix, iy][:, None])
    want, wi = O.get_CAPPI_3d(va, vr, fe, vv, h, sx, sy, lv, FILL, None, "hybrid", return_index=True)
    assert_parity(a[:, ix, iy][:, :, None], want, what="C3 full-size sample")
    got_s, gi = G.get_CAPPI_3d(va, vr, fe, vv, h, sx, sy, lv, FILL, None, "hybrid", return_index=True)
    assert np.array_equal(gi, wi) and np.array_equal(np.array(got_s), a[:, ix, iy][:, :, None])   # grid-shape independent
    # a constant field without gaps comes back exactly, wherever the geometry allows a value
    const = [np.full_like(v, 23.5) for v in vv]
    c = np.array(G.get_CAPPI_3d(va, vr, fe, const, h, gx, gy, lv, FILL, None, "hybrid"))
    geo = c != FILL
    assert np.all(c[geo] == 23.5) and np.array_equal(geo | valid, geo)           # data gaps only remove voxels
    # linearity in the moments: grid(2 v + 3) == 2 grid(v) + 3 where v has a value (same fill pattern)
    lin = [np.where(v == FILL, FILL, 2.0 * v + 3.0) for v in vv]
    d = np.array(G.get_CAPPI_3d(va, vr, fe, lin, h, gx, gy, lv, FILL, None, "hybrid"))
    assert np.array_equal(d != FILL, valid)
    assert np.max(np.abs(d[valid] - (2.0 * a[valid] + 3.0))) <= 1e-3


def test_inputs_not_mutated_and_output_owned():
    cfg = synth.config(1, scale=0.1)
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
    copies = [a.copy() for a in vv]
    gx0 = cfg["GridX"].copy()
    a = G.get_CR_xy(va, vr, fe, vv, h, cfg["GridX"], cfg["GridY"], FILL)
    b = G.get_CR_xy(va, vr, fe, vv, h, cfg["GridX"], cfg["GridY"], FILL)
    assert all(np.array_equal(x, y) for x, y in zip(vv, copies)) and np.array_equal(gx0, cfg["GridX"])
    assert a is not b and np.array_equal(a, b)
    a[...] = 0.0
    assert not np.array_equal(a, b)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]


def test_device_geometry_agreement_with_glibc():
    """How often the device's az / r / el are bit-identical to the host libm chain (reported),
    and that the differences never exceed a couple of ulps."""
    import ctypes
    from pycwr_b200 import _lib
    rng = np.random.default_rng(20260319)
    n = 400000
    x, y = rng.uniform(-3e5, 3e5, n), rng.uniform(-3e5, 3e5, n)
    az, r, el = (np.empty(n) for _ in range(3))
    _lib.check(_lib.load().pcg_debug_cartesian_to_antenna(_lib.as_ptr(x), _lib.as_ptr(y), n, 3000.0, 321.0,
                                                         O.DEFAULT_RE, _lib.as_ptr(az), _lib.as_ptr(r), _lib.as_ptr(el)))
    waz, wr, wel = O.bulk_cartesian_to_antenna(x, y, 3000.0, 321.0)
    frac = {k: float(np.mean(a == b)) for k, a, b in (("az", az, waz), ("r", r, wr), ("el", el, wel))}
    print("device==glibc bit-identical fractions:", frac)
    assert frac["r"] > 0.99 and frac["az"] > 0.5 and frac["el"] > 0.5
    assert np.max(np.abs(az - waz)) < 2e-13 and np.max(np.abs(el - wel)) < 1e-9
    assert np.max(np.abs(r - wr) / wr) < 1e-9
    az2, r2 = np.empty(n), np.empty(n)
    _lib.check(_lib.load().pcg_debug_xye_to_antenna(_lib.as_ptr(x), _lib.as_ptr(y), n, 2.4, 321.0, O.DEFAULT_RE,
                                                   _lib.as_ptr(az2), _lib.as_ptr(r2)))
    _, wr2, _ = O.bulk_xye_to_antenna(x, y, 2.4, 321.0)
    print("CR geometry r bit-identical fraction:", float(np.mean(r2 == wr2)))
    assert np.mean(r2 == wr2) > 0.99


def test_fused_sqrt_is_the_ieee_sqrt():
    """The production kernel's sqrt (kept with its reciprocal-root by-product) must equal the
    correctly rounded square root for every r2 the path can see."""
    import ctypes
    from pycwr_b200 import _lib
    rng = np.random.default_rng(7)
    n = 4_000_000
    a = np.concatenate([rng.uniform(1e-2, 1e3, n // 4), rng.uniform(1e3, 1e8, n // 4),
                        rng.uniform(1e8, 1e13, n // 4), 10.0 ** rng.uniform(-2, 13, n // 4)])
    bad = ctypes.c_uint64(123)
    worst = ctypes.c_double(0.0)
    _lib.check(_lib.load().pcg_debug_sqrt_check(_lib.as_ptr(a), a.size, ctypes.byref(bad), ctypes.byref(worst)))
    print("fused sqrt mismatches: %d of %d, worst |y*r-1| = %.3e" % (bad.value, a.size, worst.value))
    assert bad.value == 0
    assert worst.value < 1e-15


def test_repeated_calls_do_not_grow_device_memory():
    """Per-call volume arenas and threshold tables come from the library's device pool: a service that grids
    volume after volume (different shapes included) must not accumulate device memory or change its answers."""
    torch = pytest.importorskip("torch")
    cfg = synth.config(3, style="stress", scale=0.05)
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
    gx, gy, lv = cfg["GridX"], cfg["GridY"], cfg["levels"]
    first = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid")

    def one_round(i):
        n = 5 + (i % 4)                                   # sweep count and grid shape vary from call to call
        out = G.get_CAPPI_3d(va[:n], vr[:n], fe[:n], vv[:n], h, gx[: 20 + i % 7], gy[: 20 + i % 7], lv[:: 1 + i % 3],
                             FILL, None, BLINDS[i % len(BLINDS)])
        G.get_CR_xy(va[:n], vr[:n], fe[:n], vv[:n], h, gx, gy, FILL)
        return out

    for i in range(8):
        one_round(i)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for i in range(60):
        one_round(i)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, "device memory grew by %.1f MB over 60 calls" % ((free0 - free1) / 2 ** 20)
    again = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, "hybrid")
    assert np.array_equal(first, again)


def test_concurrent_callers_get_their_own_answers():
    """ctypes releases the GIL: several host threads may be inside the library at once (a product server
    gridding different radars).  Entry points serialise on the library's lock; results must not mix."""
    import threading
    cfgs = [synth.config(3, style="stress", scale=0.03), synth.config(2, style="stress", scale=0.05)]
    jobs = []
    for i, cfg in enumerate(cfgs * 2):
        va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"])
        jobs.append((va, vr, fe, vv, h, cfg["GridX"], cfg["GridY"], cfg["levels"], BLINDS[(i + 3) % len(BLINDS)]))
    want = [G.get_CAPPI_3d(*j[:8], FILL, None, j[8]) for j in jobs]
    got = [[None] * 4 for _ in jobs]
    errors = []

    def worker(t):
        try:
            for rep in range(4):
                got[t][rep] = G.get_CAPPI_3d(*jobs[t][:8], FILL, None, jobs[t][8])
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for t in range(len(jobs)):
        for rep in range(4):
            assert np.array_equal(got[t][rep], want[t]), "thread %d call %d" % (t, rep)
