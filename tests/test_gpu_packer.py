"""the volume packer on the device (`pcg_volume_create_raw`, `DeviceVolume.from_sweeps`).

The reference packs on the host (`PRD.get_vol_data` / `_extract_sweep_volume`, NRadar.py:665-712,1931-1961):
sweeps ordered by fixed angle, rays by azimuth, optional max_range_km clip, NaN -> fill.  The device packer must
give the same volume: every product of a raw-packed volume equals, bit for bit, the product of the volume packed
by the host restatement (`products.pack_volume`, itself checked against the oracle elsewhere)."""
import numpy as np
import pytest

from conftest import FILL
from pycwr_b200 import RadarGridC as G
from pycwr_b200 import synth
from pycwr_b200.products import ArrayPRD, pack_volume
from pycwr_b200.runtime import DeviceVolume

pytestmark = pytest.mark.gpu


def raw_sweeps(seed, ne=6, fields=1, naz=(40, 90), ng=(60, 140), shared_range=True):
    """Sweeps as a reader leaves them: scan order != elevation order, rays rotated / shuffled, NaN missing."""
    rng = np.random.default_rng(seed)
    fix = rng.permutation(np.sort(rng.uniform(0.4, 15.0, ne)) + np.arange(ne) * 1e-3)
    n_g = int(rng.integers(*ng))
    sweeps = []
    for s in range(ne):
        n_a = int(rng.integers(*naz))
        az = synth.azimuth_table(rng, n_a)
        gates = (np.arange(n_g if shared_range else n_g - 3 * s) + 1.0) * 300.0
        vals = []
        for _ in range(fields):
            v = synth.stress_moment(rng, n_a, gates.size, -5.0, 70.0)
            vals.append(np.where(v == FILL, np.nan, v))
        p = np.roll(np.arange(n_a), int(rng.integers(0, n_a))) if s % 2 else rng.permutation(n_a)
        sweeps.append((az[p], gates, [v[p] for v in vals] if fields > 1 else vals[0][p]))
    return sweeps, fix


def grids(seed, ext=40e3):
    rng = np.random.default_rng(seed)
    gx, gy = rng.uniform(-ext, ext, (31, 27)), rng.uniform(-ext, ext, (31, 27))
    return gx, gy, np.array([300.0, 1200.0, 2500.0, 4000.0, 7000.0])


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("max_range_km", [None, 25.0, 0.1])
def test_raw_packer_equals_host_packer(dtype, max_range_km):
    sweeps, fix = raw_sweeps(3)
    gx, gy, lv = grids(4)
    va, vr, fe, vv, h, _, _ = pack_volume(sweeps, fix, 150.0, fillvalue=FILL, max_range_km=max_range_km)
    host = DeviceVolume(va, vr, fe, vv, fillvalue=FILL, value_dtype=dtype)
    raw = DeviceVolume.from_sweeps(sweeps, fix, fillvalue=FILL, value_dtype=dtype, max_range_km=max_range_km)
    assert raw.value_dtype == host.value_dtype and list(raw.ng) == list(host.ng) and list(raw.naz) == list(host.naz)
    assert np.array_equal(raw.fix_elevation, fe) and np.array_equal(raw.last_gate, host.last_gate)
    for blind in ("mask", "hybrid", "nearest_valid"):
        a, ai = G.get_CAPPI_3d(None, None, None, host, 150.0, gx, gy, lv, FILL, None, blind, return_index=True)
        b, bi = G.get_CAPPI_3d(None, None, None, raw, 150.0, gx, gy, lv, FILL, None, blind, return_index=True)
        assert np.array_equal(a, b) and np.array_equal(ai, bi)
    assert np.array_equal(G.get_CR_xy(None, None, None, host, 150.0, gx, gy, FILL),
                          G.get_CR_xy(None, None, None, raw, 150.0, gx, gy, FILL))
    if max_range_km is None:
        assert np.isfinite(a[a != FILL]).all() and (a != FILL).mean() > 0.2
    host.close()
    raw.close()


def test_raw_packer_multi_field_ragged_and_non_prefix_clip():
    sweeps, fix = raw_sweeps(8, ne=5, fields=3, shared_range=False)
    gx, gy, lv = grids(9)
    va, vr, fe, vv, h, _, _ = pack_volume([(a, r, v[0]) for a, r, v in sweeps], fix, 20.0, fillvalue=FILL)
    host_fields = [pack_volume([(a, r, v[f]) for a, r, v in sweeps], fix, 20.0, fillvalue=FILL)[3] for f in range(3)]
    host = DeviceVolume(va, vr, fe, host_fields, fillvalue=FILL)
    raw = DeviceVolume.from_sweeps(sweeps, fix, fillvalue=FILL)
    assert raw.nfield == 3
    a = G.get_CAPPI_3d(None, None, None, host, 20.0, gx, gy, lv, FILL, None, "hybrid")
    b = G.get_CAPPI_3d(None, None, None, raw, 20.0, gx, gy, lv, FILL, None, "hybrid")
    assert np.array_equal(a, b)
    host.close(); raw.close()
    # a range table whose far gates are not at the end: the clip is a mask, not a prefix (host compaction)
    one, fix1 = raw_sweeps(12, ne=3)
    odd = []
    for az, rg, v in one:
        rg = rg.copy(); rg[5] = 90e3                       # an out-of-place gate beyond the clip
        odd.append((az, rg, v))
    va, vr, fe, vv, h, _, _ = pack_volume(odd, fix1, 0.0, fillvalue=FILL, max_range_km=30.0)
    host = DeviceVolume(va, vr, fe, vv, fillvalue=FILL)
    raw = DeviceVolume.from_sweeps(odd, fix1, fillvalue=FILL, max_range_km=30.0)
    assert list(raw.ng) == list(host.ng)
    assert np.array_equal(G.get_CAPPI_3d(None, None, None, host, 0.0, gx, gy, lv, FILL),
                          G.get_CAPPI_3d(None, None, None, raw, 0.0, gx, gy, lv, FILL))
    host.close(); raw.close()


def test_prd_products_use_the_device_packer_and_match_the_host_path():
    from pycwr_b200 import _lib, products
    sweeps, fix = raw_sweeps(21, ne=7)
    prd = ArrayPRD({"dBZ": sweeps}, fix, 80.0, lon=118.0, lat=32.0)
    xr_, yr_ = np.linspace(-35e3, 35e3, 41), np.linspace(-30e3, 30e3, 37)
    lv = np.array([500.0, 1500.0, 3000.0])
    products.add_product_CAPPI_3d_xy(prd, xr_, yr_, lv)
    products.add_product_CR_xy(prd, xr_, yr_)
    assert len(prd._dev_cache) == 1 and not prd._vol_cache            # packed once, on the device
    got3d = [v for k, v in prd.product.items() if k.startswith("CAPPI_3D")][0].values
    gotcr = [v for k, v in prd.product.items() if k.startswith("CR")][0].values
    va, vr, fe, vv, h, _, _ = pack_volume(sweeps, fix, 80.0, fillvalue=FILL)
    gx, gy = np.meshgrid(xr_, yr_, indexing="ij")
    want = G.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL)
    wcr = G.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL)
    assert np.array_equal(np.where(want == FILL, np.nan, want), got3d, equal_nan=True)
    assert np.array_equal(np.where(wcr == FILL, np.nan, wcr), gotcr, equal_nan=True)
    with pytest.raises(KeyError):
        prd.get_device_volume("ZDR")
    # C ABI argument checks
    import ctypes
    lib = _lib.load()
    h_ = ctypes.c_void_p()
    assert lib.pcg_volume_create_raw(1, None, None, None, None, 1, None, None, None, None, FILL, 1, ctypes.byref(h_)) == -1


def test_missing_marker_is_applied_on_the_device():
    """`missing=np.nan` == the wrappers' `np.where(GridV == fillvalue, np.nan, GridV)` (NRadar.py:1339,1392,1437)."""
    sweeps, fix = raw_sweeps(31, ne=6, fields=2)
    gx, gy, lv = grids(32, ext=60e3)
    dv = DeviceVolume.from_sweeps(sweeps, fix, fillvalue=FILL)
    plain = G.get_CAPPI_3d(None, None, None, dv, 40.0, gx, gy, lv, FILL, None, "hybrid")
    nan = G.get_CAPPI_3d(None, None, None, dv, 40.0, gx, gy, lv, FILL, None, "hybrid", missing=np.nan)
    assert (plain == FILL).any() and np.array_equal(np.where(plain == FILL, np.nan, plain), nan, equal_nan=True)
    one = G.get_CAPPI_xy(None, None, None, dv, 40.0, gx, gy, 1500.0, FILL, missing=-1.0)
    ref = G.get_CAPPI_xy(None, None, None, dv, 40.0, gx, gy, 1500.0, FILL)
    assert (ref == FILL).any() and np.array_equal(np.where(ref == FILL, -1.0, ref), one)
    single = DeviceVolume.from_sweeps([(a, r, v[0]) for a, r, v in sweeps], fix, fillvalue=FILL)
    cr = G.get_CR_xy(None, None, None, single, 40.0, gx, gy, FILL)
    crn = G.get_CR_xy(None, None, None, single, 40.0, gx, gy, FILL, missing=np.nan)
    assert np.array_equal(np.where(cr == FILL, np.nan, cr), crn, equal_nan=True)
    dv.close(); single.close()
