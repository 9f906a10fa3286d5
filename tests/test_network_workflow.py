"""``run_radar_network_3d`` / the PRD product wrappers with the reference's signatures.

CPU part: parameter names, order and defaults equal the reference's (``tests/golden/reference_signatures.json``,
written from /root/reference by ``tests/golden/make_signature_golden.py``), three-level option resolution
(argument > config file > ``cfg.network``, ``pycwr/interp/RadarInterp.py:205-215``), file selection.
GPU part: the workflow driven the way the reference's own tests drive it — temporary radar directories,
``read_auto`` / ``get_radar_info`` patched, a ``FakePRD`` (``test/test_examples_interp.py:143-426``).
"""
import inspect
import json
import os
from datetime import datetime
from pathlib import Path
from unittest import mock

import numpy as np
import pytest

from conftest import FILL, ROOT
from pycwr_b200 import interp as I
from pycwr_b200 import products as P


def _golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json"), encoding="utf-8") as fh:
        return json.load(fh)


def _params(fn, drop_self=False):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.kind not in (p.POSITIONAL_OR_KEYWORD,):
            continue
        out.append((p.name, p.default is not inspect.Parameter.empty, None if p.default is inspect.Parameter.empty else p.default))
    return out


def _same(ours, ref, skip_first=False, allow_extra_tail=False):
    """Same names in the same order; defaults equal (compared through ``ast.literal_eval``-able reprs)."""
    import ast
    ref = ref[1:] if skip_first else ref
    ours = ours[1:] if skip_first else ours
    assert [o[0] for o in ours[:len(ref)]] == [r["name"] for r in ref]
    if not allow_extra_tail:
        assert len(ours) == len(ref)
    for o, r in zip(ours, ref):
        assert o[1] == r["has_default"], o[0]
        if r["has_default"]:
            want = ast.literal_eval(r["default"])
            assert (o[2] == want) or (o[2] is want), (o[0], o[2], want)


def test_network_signatures_match_the_reference():
    ref = _golden()["pycwr/interp/RadarInterp.py"]
    for name in ("run_radar_network_3d", "compose_network_volume", "grid_single_radar_to_latlon_3d",
                 "grid_single_radar_cr_to_latlon", "build_latlon_grid", "_compute_observability_mask",
                 "_compute_quality_weight_volume", "_build_composite_weight_cache", "select_radar_files",
                 "discover_radar_files", "load_network_config", "radar_network_3d_to_netcdf"):
        _same(_params(getattr(I, name)), ref[name])
    assert len(_params(I.run_radar_network_3d)) == 54


def test_product_wrapper_signatures_match_the_reference():
    ref = _golden()["pycwr/core/NRadar.py"]
    mix = P.B200ProductMixin
    for name in ("add_product_CR_xy", "add_product_CAPPI_xy", "add_product_CAPPI_3d_xy", "add_product_CR_lonlat",
                 "add_product_CAPPI_lonlat", "add_product_VIL_xy", "add_product_ET_xy", "add_product_VIL_lonlat",
                 "add_product_ET_lonlat", "_grid_reflectivity_volume_xy", "_grid_reflectivity_volume_lonlat"):
        _same(_params(getattr(mix, name)), ref[name], skip_first=True)
    _same(_params(P.grid_3d_network_xy), ref["grid_3d_network_xy"])
    _same(_params(P.ArrayPRD.get_vol_data), ref["get_vol_data"], skip_first=True)


def test_option_resolution_argument_then_file_then_cfg(tmp_path):
    assert I._resolve_network_value("max_range_km", None) == 460.0                  # cfg.network
    assert I._resolve_network_value("max_range_km", 230.0) == 230.0                  # explicit wins
    assert I._resolve_network_value("max_range_km", None, network_cfg={"max_range_km": 150.0}) == 150.0
    assert I._resolve_network_value("nope", None, network_cfg={}) is None
    cfgfile = tmp_path / "net.json"
    cfgfile.write_text(json.dumps({"network": {"composite_method": "max", "lon_res_deg": 0.05}}), encoding="utf-8")
    loaded = I.load_network_config(cfgfile)
    assert loaded == {"composite_method": "max", "lon_res_deg": 0.05}
    (tmp_path / "net.txt").write_text("x", encoding="utf-8")
    with pytest.raises(ValueError):
        I.load_network_config(tmp_path / "net.txt")
    yml = tmp_path / "net.yaml"
    yml.write_text("network:\n  blind_method: mask\n", encoding="utf-8")
    assert I.load_network_config(yml) == {"blind_method": "mask"}
    saved = dict(I.cfg.network)
    try:
        I.load_network_config(cfgfile, inplace=True)
        assert I.cfg.network.composite_method == "max" and I.cfg.network["lon_res_deg"] == 0.05
    finally:
        I.cfg.network.clear()
        I.cfg.network.update(saved)
    # helper coercions (RadarInterp.py:191-202,259-278,396-458)
    assert I._coerce_field_names(None) == ["dBZ"] and I._coerce_field_names("ZDR") == ["ZDR"]
    assert I._coerce_field_range_mode(["dBZ", "ZDR"], None) == {"dBZ": "native", "ZDR": "aligned"}
    assert I._coerce_field_range_mode(["dBZ", "ZDR"], "aligned") == {"dBZ": "native", "ZDR": "aligned"}
    with pytest.raises(ValueError):
        I._coerce_field_range_mode(["ZDR"], {"ZDR": "weird"})
    assert I._coerce_output_products(["cr", "CR", "et"]) == ["CR", "ET"]
    with pytest.raises(ValueError):
        I._coerce_output_products(["KDP"])
    assert I._coerce_quality_weight_terms(None) == ["range", "beam", "vertical", "qc", "blockage"]
    with pytest.raises(ValueError):
        I._coerce_quality_weight_terms(["speed"])
    lv = np.array([1000.0, 3000.0])
    assert np.array_equal(I._coerce_product_level_heights(lv, None), np.arange(1000.0, 3250.0, 500.0))
    assert np.array_equal(I._coerce_product_level_heights(lv, [500.0]), [500.0])
    assert I._normalize_blind_method(None) == "mask" and I._normalize_blind_method("HYBRID") == "hybrid"
    with pytest.raises(ValueError):
        I._normalize_blind_method("closest")
    with pytest.raises(ValueError):
        I._validate_network_field_names(["dBZ", "V"])


def _radar_dirs(tmp_path, names=("Z9001", "Z9002"), stamp="20260317070000"):
    dirs = []
    for i, name in enumerate(names):
        d = tmp_path / name
        d.mkdir()
        (d / ("%s_%s_test.bin" % ("AB"[i % 2], stamp))).write_text("", encoding="ascii")
        dirs.append(str(d))
    return dirs


def test_select_radar_files_picks_nearest_within_tolerance(tmp_path):
    dirs = _radar_dirs(tmp_path)
    (Path(dirs[0]) / "A_20260317065400_test.bin").write_text("", encoding="ascii")
    (Path(dirs[1]) / "notes.txt").write_text("", encoding="ascii")
    got = I.select_radar_files(dirs, "2026-03-17T07:01:00", tolerance_minutes=10)
    assert [g["radar_id"] for g in got] == ["Z9001", "Z9002"]
    assert got[0]["path"].endswith("A_20260317070000_test.bin") and got[0]["delta_seconds"] == 60.0
    assert got[0]["scan_time"] == datetime(2026, 3, 17, 7, 0, 0)
    with pytest.raises(ValueError):
        I.select_radar_files(dirs, datetime(2026, 3, 17, 9, 0, 0), tolerance_minutes=10)
    with pytest.raises(ValueError):
        I.select_radar_files([], "2026-03-17T07:00:00")
    with pytest.raises(ValueError):
        I.select_radar_files(dirs, None)
    with pytest.raises(ValueError):
        I.parse_radar_time_from_filename("no_time_here.bin")


def test_run_radar_network_3d_argument_errors_come_before_any_gpu_work(tmp_path):
    with pytest.raises(ValueError, match="target_time is required"):
        I.run_radar_network_3d(None)
    with pytest.raises(ValueError, match="lon/lat bounds"):
        I.run_radar_network_3d("2026-03-17T07:00:00", radar_dirs=["x"], level_heights=[1000.0])
    kw = dict(radar_dirs=_radar_dirs(tmp_path), lon_min=120.0, lon_max=120.1, lat_min=30.0, lat_max=30.1,
              lon_res_deg=0.01, lat_res_deg=0.01, level_heights=[1000.0])
    with pytest.raises(ValueError, match="coverage_policy"):
        I.run_radar_network_3d("2026-03-17T07:00:00", coverage_policy="all", **kw)
    with pytest.raises(ValueError, match="plot_style"):
        I.run_radar_network_3d("2026-03-17T07:00:00", plot_style="fancy", **kw)
    with pytest.raises(ValueError, match="velocity"):
        I.run_radar_network_3d("2026-03-17T07:00:00", field_names=["V"], **kw)
    with pytest.raises(ValueError, match="blind_method"):
        I.run_radar_network_3d("2026-03-17T07:00:00", blind_method="closest", **kw)
    with pytest.raises(ValueError, match="no coarser than 0.01"):
        I.run_radar_network_3d("2026-03-17T07:00:00", **dict(kw, lon_res_deg=0.1))
    with pytest.raises(ValueError, match="No radar files matched"):
        I.run_radar_network_3d("2026-03-17T09:00:00", **kw)
    with pytest.raises(ImportError):         # readers are hooks: nothing to read the (empty) files with here
        I.run_radar_network_3d("2026-03-17T07:00:00", **kw)


# ------------------------------------------------------------------------------------------------
# GPU: the workflow itself
# ------------------------------------------------------------------------------------------------

class _Value(object):
    def __init__(self, v):
        self.values = v


class FakePRD(object):
    """The reference tests' ``FakePRD`` (test/test_examples_interp.py:146-175) with a volume that really covers
    the target cell: two sweeps (1 and 2 degrees), four rays, constant moments."""

    def __init__(self, radar_id, lon=120.0, lat=30.0, values=None, alt=None, sweeps=(1.0, 2.0)):
        self.radar_id = radar_id
        self.effective_earth_radius = None
        self.scan_info = {"longitude": _Value(np.float64(lon)), "latitude": _Value(np.float64(lat))}
        if alt is not None:
            self.scan_info["altitude"] = _Value(np.float64(alt))
        self.values = values or {}
        self.sweeps = sweeps
        self.calls = []

    def get_vol_data(self, field_name="dBZ", fillvalue=-999.0, range_mode="aligned", max_range_km=None):
        self.calls.append((field_name, range_mode, max_range_km))
        if field_name not in self.values:
            raise KeyError(field_name)
        ne = len(self.sweeps)
        az = np.array([0.0, 90.0, 180.0, 270.0])
        rng = np.arange(1000.0, 6000.0, 1000.0)
        self.vol = ([az.copy() for _ in range(ne)], [rng.copy() for _ in range(ne)],
                    np.array(self.sweeps, dtype=np.float64),
                    [np.full((4, 5), float(self.values[field_name])) for _ in range(ne)], 100.0,
                    float(self.scan_info["longitude"].values), float(self.scan_info["latitude"].values))
        return self.vol


def _sub(tmp_path, name):
    d = tmp_path / name
    d.mkdir()
    return d


def _run(tmp_path, prds, info=None, names=None, **kw):
    """run_radar_network_3d with read_auto / get_radar_info patched like the reference's tests patch them."""
    names = names or sorted(prds)
    dirs = _radar_dirs(tmp_path, names)
    reads, lookups = [], []

    def fake_read_auto(path, station_lon=None, station_lat=None, station_alt=None, effective_earth_radius=None):
        rid = Path(path).parent.name
        reads.append((rid, station_lon, station_lat, station_alt))
        return prds[rid]

    def fake_get_radar_info(path):
        rid = Path(path).parent.name
        lookups.append(rid)
        if info is None:
            return 30.0, 120.0, 100.0, 2.8
        if rid not in info:
            raise KeyError(rid)
        return info[rid]

    args = dict(target_time="2026-03-17T07:00:00", radar_dirs=dirs, lon_min=120.02, lon_max=120.02, lat_min=30.0,
                lat_max=30.0, lon_res_deg=0.01, lat_res_deg=0.01, level_heights=np.array([150.0]), parallel=False)
    args.update(kw)
    with mock.patch("pycwr_b200.interp.read_auto", side_effect=fake_read_auto):
        with mock.patch("pycwr_b200.interp.get_radar_info", side_effect=fake_get_radar_info):
            ds = I.run_radar_network_3d(**args)
    return ds, reads, lookups


@pytest.mark.gpu
def test_run_radar_network_3d_supports_multiple_fields(tmp_path):
    """test/test_examples_interp.py:143-245 without the gridding patched out: two co-located radars, 10 / 30 dBZ
    and 1 / 3 ZDR, exp_weighted -> 20 and 2."""
    prds = {"Z9001": FakePRD("Z9001", values={"dBZ": 10.0, "ZDR": 1.0}),
            "Z9002": FakePRD("Z9002", values={"dBZ": 30.0, "ZDR": 3.0})}
    ds, reads, lookups = _run(tmp_path, prds, field_names=["dBZ", "ZDR"], composite_method="exp_weighted",
                              level_heights=np.array([150.0, 160.0]))
    assert ds["dBZ"].dims == ("z", "lat", "lon") and np.asarray(ds["dBZ"].values).shape == (2, 1, 1)
    assert np.allclose(ds["dBZ"].values, 20.0, atol=1e-3) and np.allclose(ds["ZDR"].values, 2.0, atol=1e-3)
    assert ds.attrs["composite_method"] == "exp_weighted"
    assert json.loads(ds.attrs["field_range_mode"]) == {"dBZ": "native", "ZDR": "aligned"}
    assert json.loads(ds["dBZ"].attrs["actual_max_range_km_by_radar"]) == {"Z9001": 5.0, "Z9002": 5.0}
    assert json.loads(ds["dBZ"].attrs["actual_max_range_km_by_radar_and_sweep"])["Z9001"] == [5.0, 5.0]
    assert json.loads(ds["dBZ"].attrs["source_field_name_by_radar"]) == {"Z9001": "dBZ", "Z9002": "dBZ"}
    assert ds.attrs["radar_count"] == 2 and ds.attrs["use_qc"] == "false" and ds.attrs["qc_method"] == "none"
    assert ds.attrs["target_time"] == "2026-03-17T07:00:00" and ds.attrs["network_config_path"] == ""
    assert len(ds.attrs["radar_files"].split("|")) == 2
    assert json.loads(ds.attrs["output_products"]) == ["CR", "VIL", "ET"]
    assert np.array_equal(np.asarray(ds["coverage_count"].values), np.full((2, 1, 1), 2, dtype=np.int16))
    assert np.array_equal(np.asarray(ds["blind_mask"].values), np.zeros((2, 1, 1), dtype=np.int8))
    assert np.allclose(ds["CR"].values, 30.0, atol=1e-3)
    assert ("dBZ", "native", 460.0) in prds["Z9001"].calls and ("ZDR", "aligned", 460.0) in prds["Z9001"].calls
    # stations come from the lookup, and are handed to the reader (test/test_examples_interp.py:331-426)
    assert lookups == ["Z9001", "Z9002"]
    assert reads == [("Z9001", 120.0, 30.0, 100.0), ("Z9002", 120.0, 30.0, 100.0)]


@pytest.mark.gpu
def test_run_radar_network_3d_skips_missing_field_on_one_radar(tmp_path):
    """test/test_examples_interp.py:247-329."""
    prds = {"Z9001": FakePRD("Z9001", values={"dBZ": 20.0, "ZDR": 2.0}),
            "Z9002": FakePRD("Z9002", values={"dBZ": 20.0})}
    ds, _, _ = _run(tmp_path, prds, field_names=["dBZ", "ZDR"], composite_method="exp_weighted")
    assert np.allclose(ds["dBZ"].values, 20.0, atol=1e-3) and np.allclose(ds["ZDR"].values, 2.0, atol=1e-3)
    assert json.loads(ds["ZDR"].attrs["actual_max_range_km_by_radar"])["Z9002"] is None


@pytest.mark.gpu
def test_run_radar_network_3d_station_fallback_and_nearest(tmp_path):
    """No station table entry -> the file itself is read for its site (RadarInterp.py:1784-1795); `nearest`
    picks the closer radar per cell."""
    prds = {"Z9001": FakePRD("Z9001", lon=120.0, lat=30.0, values={"dBZ": 10.0}, alt=100.0),
            "Z9002": FakePRD("Z9002", lon=120.06, lat=30.0, values={"dBZ": 30.0}, alt=100.0)}
    ds, reads, lookups = _run(tmp_path, prds, info={}, lon_min=120.02, lon_max=120.04, composite_method="nearest",
                              output_products=[], plot_overview=False, plot_height_levels=[])
    assert lookups == ["Z9001", "Z9002"]
    assert reads[:2] == [("Z9001", None, None, None), ("Z9002", None, None, None)]      # the pre-reads
    assert reads[2:] == [("Z9001", 120.0, 30.0, 100.0), ("Z9002", 120.06, 30.0, 100.0)]
    got = np.asarray(ds["dBZ"].values)[0, 0]
    assert got.shape == (3,) and np.allclose(got[[0, 2]], [10.0, 30.0], atol=1e-3)
    assert "CR" not in ds and "VIL" not in ds
    assert json.loads(ds.attrs["output_products"]) == []


@pytest.mark.gpu
def test_run_radar_network_3d_config_file_and_product_levels(tmp_path):
    prds = {"Z9001": FakePRD("Z9001", values={"dBZ": 25.0}), "Z9002": FakePRD("Z9002", values={"dBZ": 35.0})}
    cfgfile = tmp_path / "network.json"
    cfgfile.write_text(json.dumps({"network": {"composite_method": "max", "blind_method": "mask",
                                               "product_level_heights": [140.0, 150.0, 160.0],
                                               "et_threshold_dbz": 30.0}}), encoding="utf-8")
    ds, _, _ = _run(tmp_path, prds, config_path=str(cfgfile))
    assert ds.attrs["composite_method"] == "max" and ds.attrs["blind_method"] == "mask"
    assert ds.attrs["network_config_path"] == str(cfgfile)
    assert np.allclose(ds["dBZ"].values, 35.0, atol=1e-3)
    assert json.loads(ds["VIL"].attrs["source_levels"]) == [140.0, 150.0, 160.0]
    assert ds["ET"].attrs["threshold_dbz"] == 30.0 and ds["ET"].dims == ("lat", "lon")
    assert np.asarray(ds["ET_TOPPED"].values).dtype == np.uint8
    # an explicit argument beats the file
    ds2, _, _ = _run(_sub(tmp_path, "b"), prds, config_path=str(cfgfile), composite_method="exp_weighted")
    assert ds2.attrs["composite_method"] == "exp_weighted" and np.allclose(ds2["dBZ"].values, 30.0, atol=1e-3)


@pytest.mark.gpu
def test_network_with_irregular_radar_takes_the_per_radar_flow(tmp_path):
    """A single-sweep radar has no packed layout: the network product must still come out, through the
    reference's per-radar flow, and equal what the fused kernels give for the regular radar alone where the
    single-sweep radar sees nothing."""
    prds = {"Z9001": FakePRD("Z9001", values={"dBZ": 10.0}),
            "Z9002": FakePRD("Z9002", values={"dBZ": 30.0}, sweeps=(1.5,))}
    ds, _, _ = _run(tmp_path, prds, composite_method="exp_weighted", level_heights=np.array([150.0, 400.0]))
    got = np.asarray(ds["dBZ"].values)[:, 0, 0]
    assert np.allclose(got[0], 20.0, atol=1e-3)      # both radars see level 150 m (el ~ 1.5 deg)
    assert np.isnan(got[1]) or np.allclose(got[1], 10.0, atol=1e-3)
    assert np.asarray(ds["coverage_count"].values)[0, 0, 0] == 2


@pytest.mark.gpu
def test_fused_and_per_radar_flows_agree(tmp_path):
    from pycwr_b200 import synth
    rng = np.random.default_rng(5)
    prds = {}
    for i, rid in enumerate(("Z9001", "Z9002", "Z9003")):
        vol = synth.make_volume(100 + i, naz=72, ngate=200, gate_m=1000.0, fields=("dBZ",), style="stress")
        sw = []
        for k in range(vol["fix_elevation"].size):
            val = vol["vol_value"]["dBZ"][k].copy()
            val[val == FILL] = np.nan
            sw.append((vol["vol_azimuth"][k], vol["vol_range"][k], val))
        prds[rid] = P.ArrayPRD({"dBZ": sw}, vol["fix_elevation"], vol["radar_height"], lon=120.0 + 0.8 * i,
                               lat=30.0 + rng.uniform(-0.5, 0.5), beam_width=np.full(len(sw), 1.0))
    info = {rid: (p.lat, p.lon, 50.0, 2.8) for rid, p in prds.items()}
    kw = dict(lon_min=119.0, lon_max=123.0, lat_min=29.0, lat_max=31.0, lon_res_deg=0.1, lat_res_deg=0.1,
              level_heights=np.array([1000.0, 2500.0, 6000.0]), plot_style="simple", info=info)
    fused, _, _ = _run(_sub(tmp_path, "a"), prds, **kw)
    try:
        I.FUSED_NETWORK = False
        plain, _, _ = _run(_sub(tmp_path, "b"), prds, **kw)
    finally:
        I.FUSED_NETWORK = True
    for name in ("dBZ", "CR", "VIL", "ET"):
        a, b = np.asarray(fused[name].values), np.asarray(plain[name].values)
        assert np.array_equal(np.isnan(a), np.isnan(b)), name
        ok = ~np.isnan(a)
        assert ok.any() and np.max(np.abs(a[ok] - b[ok])) <= (1e-3 if name in ("dBZ", "CR") else 1e-3 * max(1.0, np.abs(b[ok]).max())), name
    for name in ("coverage_count", "blind_mask", "ET_TOPPED"):
        assert np.array_equal(np.asarray(fused[name].values), np.asarray(plain[name].values)), name
    # a scalar blockage weight other than 1 changes the quality-weighted composite (per-radar flow only)
    blocked, _, _ = _run(_sub(tmp_path, "c"), prds, blockage_config={"Z9001": 0.25}, **kw)
    a, b = np.asarray(fused["dBZ"].values), np.asarray(blocked["dBZ"].values)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) > 1e-3


@pytest.mark.gpu
def test_qc_hook_and_attrs(tmp_path):
    class QcPRD(FakePRD):
        nsweeps = 1
        fields = [{"dBZ": 1, "KDP": 1}]

        def apply_dualpol_qc(self, inplace=True, **kwargs):
            self.qc_kwargs = kwargs
            self.values = dict(self.values, Zc=self.values["dBZ"] - 5.0)
            return self

    prds = {"Z9001": QcPRD("Z9001", values={"dBZ": 30.0}), "Z9002": FakePRD("Z9002", values={"dBZ": 30.0})}
    prds["Z9002"].nsweeps = 1
    prds["Z9002"].fields = [{"dBZ": 1}]                      # no KDP / PhiDP: QC falls back to the original
    ds, _, _ = _run(tmp_path, prds, use_qc=True, composite_method="max", quality_weight_terms=["range"])
    assert prds["Z9001"].qc_kwargs["band"] == "C" and prds["Z9001"].qc_kwargs["clear_air_max_ref"] == 15.0
    assert json.loads(ds["dBZ"].attrs["source_field_name_by_radar"]) == {"Z9001": "Zc", "Z9002": "dBZ"}
    assert ds["dBZ"].attrs["qc_applied"] == "partial" and ds.attrs["use_qc"] == "true"
    qc = json.loads(ds.attrs["qc_applied_by_radar"])
    assert qc["Z9001"] == {"applied": True, "error": None}
    assert qc["Z9002"] == {"applied": False, "error": "missing_dualpol_fields"}
    assert np.allclose(ds["dBZ"].values, 30.0, atol=1e-3)       # max(25 corrected, 30 original)
    with pytest.raises(ValueError, match="moments required"):
        _run(_sub(tmp_path, "x"), prds, use_qc=True, qc_fallback="raise")


@pytest.mark.gpu
def test_run_radar_network_3d_fans_out_over_devices(tmp_path, monkeypatch):
    """``parallel`` / ``max_workers`` one level down: the composite of every field goes through
    ``network_compose(devices=...)``.  On a one-GPU box the device list is forced to ``0,0`` (two slabs, two host
    threads on the same GPU): the dataset must equal the single-device run variable by variable."""
    def make():
        return {"Z9001": FakePRD("Z9001", lon=120.0, lat=30.0, values={"dBZ": 10.0, "ZDR": 1.0}),
                "Z9002": FakePRD("Z9002", lon=120.05, lat=30.02, values={"dBZ": 30.0, "ZDR": 3.0})}
    kw = dict(field_names=["dBZ", "ZDR"], lon_min=119.98, lon_max=120.08, lat_min=29.97, lat_max=30.05,
              level_heights=np.array([150.0, 400.0, 900.0]), composite_method="quality_weighted")
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    monkeypatch.delenv("PYCWR_B200_DEVICES", raising=False)
    one, _, _ = _run(tmp_path / "a", make(), **kw)
    monkeypatch.setenv("PYCWR_B200_DEVICES", "0,0")
    many, _, _ = _run(tmp_path / "b", make(), **kw)
    assert sorted(one.data_vars) == sorted(many.data_vars)
    for name in one.data_vars:
        a, b = np.asarray(one[name].values), np.asarray(many[name].values)
        assert a.shape == b.shape and a.dtype == b.dtype, name
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), name
    assert np.isfinite(np.asarray(one["dBZ"].values)).any()
    assert I._resolve_devices(False, None) == [0, 0]                       # the override wins
    monkeypatch.delenv("PYCWR_B200_DEVICES")
    assert I._resolve_devices(False, 8) is None and I._resolve_devices(True, 1) is None
