"""Host-side mirror of the reference's network-compositing interface
(``pycwr/interp/RadarInterp.py``) over the B200 kernels.

Same names, argument meaning and errors as the reference for the functions on the gridding path:

* ``build_latlon_grid`` (:382-393), ``_build_composite_weight_cache`` (:1379-1396) — host NumPy,
  2-D, exactly the reference's expressions;
* ``compose_network_volume`` (:1399-1487) — elementwise GPU kernel on already gridded volumes;
* ``_compute_observability_mask`` (:540-600), ``_compute_quality_weight_volume`` (:603-696) —
  stand-alone GPU kernel (any sweep count);
* ``grid_single_radar_to_latlon_3d`` (:963-1023), ``grid_single_radar_cr_to_latlon`` (:1026-1068)
  on duck-typed PRD objects (``get_vol_data()`` / ``.vol`` / ``.scan_info``, like the reference's
  own test doubles);
* ``network_compose`` — what the reference cannot express: all radars of one field merged in ONE
  fused kernel (gridding + sanity filter + quality + observability + composite + coverage + CR),
  optionally on a row slab of the grid (multi-GPU sharding);
* ``run_radar_network_3d`` — the array-level core of :1610-2193 with the reference's option names
  and defaults, output variable names, dims ``(z, lat, lon)`` and attrs; radars are passed as
  PRD-like objects (file discovery and readers are out of scope).

There is no CPU fallback: every function that grids or merges calls the CUDA library.
"""
import ctypes
import json

import numpy as np

from . import RadarGridC as G
from . import _lib
from .runtime import DeviceVolume, _f64_1d, _f64_2d, pinned_empty

DEFAULTS = {                     # pycwr/configure/config.py:12-62 as resolved in RadarInterp.py:1672-1774
    "composite_method": "quality_weighted",
    "influence_radius_m": 300000.0,
    "fillvalue": -999.0,
    "blind_method": "hybrid",
    "max_range_km": 460.0,
    "quality_weight_terms": ("range", "beam", "vertical"),
    "range_decay_m": 230000.0,
    "beam_radius_ref_m": 3000.0,
    "vertical_gap_ref_deg": 0.5,
    "vil_min_dbz": 18.0,
    "vil_max_dbz_cap": 56.0,
    "et_threshold_dbz": 18.0,
}
_METHODS = ("exp_weighted", "nearest", "max", "quality_weighted")


def build_latlon_grid(lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg):
    """Regular lon/lat target grid, meshgrid ``xy`` -> (nlat, nlon) (RadarInterp.py:382-393)."""
    lon_res_deg = float(lon_res_deg)
    lat_res_deg = float(lat_res_deg)
    if lon_res_deg <= 0 or lat_res_deg <= 0:
        raise ValueError("lon_res_deg and lat_res_deg must be positive.")
    if lon_max < lon_min or lat_max < lat_min:
        raise ValueError("Grid bounds must satisfy min <= max.")
    lon = np.arange(float(lon_min), float(lon_max) + lon_res_deg * 0.5, lon_res_deg, dtype=np.float64)
    lat = np.arange(float(lat_min), float(lat_max) + lat_res_deg * 0.5, lat_res_deg, dtype=np.float64)
    grid_lon, grid_lat = np.meshgrid(lon, lat, indexing="xy")
    return lon, lat, grid_lon, grid_lat


def _check_method(method):
    method = str(method)
    if method not in _METHODS:
        raise ValueError("Unsupported composite_method: %s." % method)
    return method


def _build_composite_weight_cache(radar_sites_xy, grid_x, grid_y, method="exp_weighted", influence_radius_m=300000.0):
    """Per-radar 2-D weight (or squared-distance) fields (RadarInterp.py:1379-1396); host NumPy
    like the reference: nradar small 2-D arrays, evaluated with the same expressions."""
    method = str(method)
    sites = np.asarray(radar_sites_xy, dtype=np.float64)
    gx = np.asarray(grid_x, dtype=np.float64)
    gy = np.asarray(grid_y, dtype=np.float64)
    cache = []
    for sx, sy in sites:
        d2 = (gx - sx) ** 2 + (gy - sy) ** 2
        if method in ("exp_weighted", "quality_weighted"):
            cache.append(np.exp(-4.0 * d2 / (float(influence_radius_m) ** 2)))
        elif method == "nearest":
            cache.append(d2)
        elif method == "max":
            return None
        else:
            raise ValueError("Unsupported composite_method: %s." % method)
    return cache


def compose_network_volume(radar_volumes, radar_sites_xy, grid_x, grid_y, fillvalue=-999.0, method="exp_weighted",
                           influence_radius_m=300000.0, weight_cache=None, quality_weights=None):
    """Composite already gridded single-radar volumes (RadarInterp.py:1399-1487).
    Returns ``(composite, valid_count int16)``."""
    if not radar_volumes:
        raise ValueError("radar_volumes must contain at least one volume.")
    method = _check_method(method)
    influence_radius_m = float(influence_radius_m)
    if influence_radius_m <= 0:
        raise ValueError("influence_radius_m must be positive.")
    sites = np.asarray(radar_sites_xy, dtype=np.float64)
    if sites.shape != (len(radar_volumes), 2):
        raise ValueError("radar_sites_xy must have shape (nradar, 2).")
    vols = [np.ascontiguousarray(v, dtype=np.float64) for v in radar_volumes]
    shape = vols[0].shape
    for v in vols[1:]:
        if v.shape != shape:
            raise ValueError("All radar volumes must share the same shape.")
    if len(shape) != 3:
        raise ValueError("radar volumes must have shape (nz, ny, nx).")
    n = len(vols)
    w2d = None
    if method != "max":
        if weight_cache is None:
            weight_cache = _build_composite_weight_cache(sites, grid_x, grid_y, method, influence_radius_m)
        w2d = np.ascontiguousarray(np.stack([np.asarray(w, dtype=np.float64) for w in weight_cache], axis=0))
        if w2d.shape != (n,) + shape[1:]:
            raise ValueError("weight_cache must hold one (ny, nx) field per radar.")
    quals = None
    if method == "quality_weighted" and quality_weights is not None:
        quals = [np.ascontiguousarray(q, dtype=np.float64) for q in quality_weights]
        if len(quals) != n or any(q.shape != shape for q in quals):
            raise ValueError("quality_weights must match radar_volumes in shape.")
    out = np.empty(shape, dtype=np.float64)
    count = np.empty(shape, dtype=np.int16)
    if out.size:
        vp = (_lib.c_double_p * n)(*[v.ctypes.data_as(_lib.c_double_p) for v in vols])
        qp = (_lib.c_double_p * n)(*[q.ctypes.data_as(_lib.c_double_p) for q in quals]) if quals else None
        _lib.check(_lib.load().pcg_compose_volumes(
            n, vp, qp, _lib.as_ptr(w2d) if w2d is not None else None, _lib.NET_METHODS[method],
            shape[0], shape[1], shape[2], float(fillvalue), _lib.as_ptr(out), _lib.as_ptr(count)))
    return out, count


def _gates(vol_range, fix_elevation, radar_height, grid_x, grid_y, level_heights, radar_x, radar_y, beam_widths,
           effective_earth_radius, terms, range_decay_m, beam_radius_ref_m, vertical_gap_ref_deg, want_mask,
           want_quality):
    gx = np.ascontiguousarray(grid_x, dtype=np.float64)
    gy = np.ascontiguousarray(grid_y, dtype=np.float64)
    if gx.ndim != 2 or gy.shape != gx.shape:
        raise ValueError("grid_x and grid_y must be 2-D arrays of the same shape.")
    fix = np.ascontiguousarray(fix_elevation, dtype=np.float64).ravel()
    levels = np.ascontiguousarray(level_heights, dtype=np.float64).ravel()
    ne = int(fix.size)
    shape = (levels.size,) + gx.shape
    mask = np.zeros(shape, dtype=np.uint8) if want_mask else None
    qual = np.zeros(shape, dtype=np.float64) if want_quality else None
    if ne and levels.size and gx.size:
        first = np.array([float(vol_range[k][0]) for k in range(ne)], dtype=np.float64)
        last = np.array([float(vol_range[k][-1]) for k in range(ne)], dtype=np.float64)
        bw = None
        if beam_widths is not None:                      # RadarInterp.py:558-562
            bw = np.asarray(beam_widths, dtype=np.float64).ravel()
            if bw.size < ne:
                bw = np.concatenate([bw, np.full(ne - bw.size, float(bw[-1]))])
            bw = np.ascontiguousarray(bw[:ne])
        radius = G._resolve_effective_earth_radius(effective_earth_radius)
        code = sum(_lib.QUALITY_TERMS[t] for t in set(terms or []) if t in _lib.QUALITY_TERMS)
        _lib.check(_lib.load().pcg_network_gates(
            ne, fix.ctypes.data_as(_lib.c_double_p), first.ctypes.data_as(_lib.c_double_p),
            last.ctypes.data_as(_lib.c_double_p), bw.ctypes.data_as(_lib.c_double_p) if bw is not None else None,
            float(radar_height), radius, float(radar_x), float(radar_y), _lib.as_ptr(gx), _lib.as_ptr(gy),
            gx.shape[0], gx.shape[1], levels.ctypes.data_as(_lib.c_double_p), int(levels.size), code,
            float(range_decay_m), float(beam_radius_ref_m), float(vertical_gap_ref_deg),
            _lib.as_ptr(mask) if mask is not None else None, _lib.as_ptr(qual) if qual is not None else None))
    return mask, qual


def _compute_observability_mask(vol_range, fix_elevation, radar_height, grid_x, grid_y, level_heights, radar_x,
                                radar_y, beam_widths=None, effective_earth_radius=None):
    """Bool (nz, ny, nx): voxels really observed by the radar (RadarInterp.py:540-600)."""
    mask, _ = _gates(vol_range, fix_elevation, radar_height, grid_x, grid_y, level_heights, radar_x, radar_y,
                     beam_widths, effective_earth_radius, (), 1.0, 1.0, 1.0, True, False)
    return mask.astype(bool)


def _compute_quality_weight_volume(vol_range, fix_elevation, radar_height, grid_x, grid_y, level_heights, radar_x,
                                   radar_y, beam_widths=None, effective_earth_radius=None, quality_weight_terms=None,
                                   range_decay_m=230000.0, beam_radius_ref_m=3000.0, vertical_gap_ref_deg=0.5):
    """Geometric quality weights (nz, ny, nx) float64 (RadarInterp.py:603-696)."""
    _, qual = _gates(vol_range, fix_elevation, radar_height, grid_x, grid_y, level_heights, radar_x, radar_y,
                     beam_widths, effective_earth_radius, quality_weight_terms, range_decay_m, beam_radius_ref_m,
                     vertical_gap_ref_deg, False, True)
    return qual


def _get_sweep_beam_widths(prd, sweep_count):
    """RadarInterp.py:527-537."""
    n = int(sweep_count)
    try:
        bw = np.asarray(prd.scan_info["beam_width"].values, dtype=np.float64)
    except Exception:
        bw = np.full(n, 1.0, dtype=np.float64)
    if bw.size == 0:
        return np.full(n, 1.0, dtype=np.float64)
    if bw.size < n:
        bw = np.pad(bw, (0, n - bw.size), constant_values=float(bw[-1]))
    return np.where(np.isfinite(bw[:n]), bw[:n], 1.0)


def _prd_volume(prd, field_name, fillvalue, range_mode, max_range_km):
    if hasattr(prd, "_resolve_field_range_mode"):
        range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
    prd.get_vol_data(field_name=field_name, fillvalue=fillvalue, range_mode=range_mode, max_range_km=max_range_km)
    vol_azimuth, vol_range, fix_elevation, vol_value, radar_height = prd.vol[:5]
    return vol_azimuth, vol_range, np.asarray(fix_elevation, dtype=np.float64), vol_value, float(radar_height), range_mode


def _prd_device_volume(prd, field_name, fillvalue, range_mode, max_range_km):
    """(DeviceVolume, radar height, range mode used, owned): the PRD's cached device volume when it offers one
    (``get_device_volume``: packed by the GPU from the raw sweeps), else a fresh upload of ``prd.vol``."""
    if hasattr(prd, "get_device_volume"):
        if hasattr(prd, "_resolve_field_range_mode"):
            range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
        dv = prd.get_device_volume(field_name=field_name, fillvalue=fillvalue, range_mode=range_mode,
                                   max_range_km=max_range_km)
        return dv, float(prd.radar_height), range_mode, False
    va, vr, fe, vv, h, used = _prd_volume(prd, field_name, fillvalue, range_mode, max_range_km)
    return DeviceVolume(va, vr, fe, vv, fillvalue=fillvalue, value_dtype="f32"), h, used, True


def _sanity_filter(grid_value, vol_value, fillvalue):
    """RadarInterp.py:1003-1014 on the host, for volumes the fused kernel cannot take."""
    lo, hi = np.inf, -np.inf
    for v in vol_value:
        ok = np.isfinite(v) & (v != fillvalue)
        if ok.any():
            lo, hi = min(lo, float(v[ok].min())), max(hi, float(v[ok].max()))
    if lo > hi:
        return np.full_like(grid_value, float(fillvalue), dtype=np.float64)
    tol = max(1e-6, 1e-6 * max(abs(lo), abs(hi), 1.0))
    bad = np.isfinite(grid_value) & ((grid_value < lo - tol) | (grid_value > hi + tol))
    return np.where(bad, float(fillvalue), grid_value) if bad.any() else grid_value


def grid_single_radar_to_latlon_3d(prd, field_name, grid_x, grid_y, level_heights, radar_x, radar_y, fillvalue=-999.0,
                                   effective_earth_radius=None, range_mode=None, max_range_km=None,
                                   blind_method="mask", return_metadata=False):
    """Grid one radar volume onto the shared Cartesian grid (RadarInterp.py:963-1023)."""
    va, vr, fe, vv, h, range_mode = _prd_volume(prd, field_name, fillvalue, range_mode, max_range_km)
    gx = np.asarray(grid_x, dtype=np.float64)
    gy = np.asarray(grid_y, dtype=np.float64)
    levels = np.asarray(level_heights, dtype=np.float64)
    dv = DeviceVolume(va, vr, fe, vv, fillvalue=float(fillvalue), value_dtype="f32")
    try:
        if dv.value_dtype == "f32":
            res = network_compose([dv], [(float(radar_x), float(radar_y))], [h], gx, gy, levels, fillvalue=fillvalue,
                                  method="max", blind_method=blind_method,
                                  effective_earth_radius=effective_earth_radius, outputs=())
            grid_value = res["composite"]
        else:
            local_x = gx - float(radar_x)
            local_y = gy - float(radar_y)
            grid_value = G.get_CAPPI_3d(None, None, None, dv, h, local_x, local_y, levels, float(fillvalue),
                                        effective_earth_radius=effective_earth_radius, blind_method=blind_method)
            grid_value = _sanity_filter(np.asarray(grid_value), vv, fillvalue)
    finally:
        dv.close()
    if return_metadata:
        return grid_value, {
            "range_mode": range_mode,
            "actual_max_range_km": max(float(r[-1]) for r in vr) / 1000.0,
            "actual_max_range_km_by_sweep": [float(r[-1]) / 1000.0 for r in vr],
        }
    return grid_value


def grid_single_radar_cr_to_latlon(prd, field_name, grid_x, grid_y, radar_x, radar_y, fillvalue=-999.0,
                                   effective_earth_radius=None, range_mode=None, max_range_km=None):
    """Grid one radar's composite reflectivity (RadarInterp.py:1026-1068)."""
    va, vr, fe, vv, h, _ = _prd_volume(prd, field_name, fillvalue, range_mode, max_range_km)
    local_x = np.asarray(grid_x, dtype=np.float64) - float(radar_x)
    local_y = np.asarray(grid_y, dtype=np.float64) - float(radar_y)
    return np.asarray(G.get_CR_xy(va, vr, fe, vv, h, local_x, local_y, float(fillvalue),
                                  effective_earth_radius=effective_earth_radius), dtype=np.float64)


def network_compose(volumes, radar_sites_xy, radar_heights, grid_x, grid_y, level_heights, fillvalue=-999.0,
                    method="quality_weighted", influence_radius_m=300000.0, blind_method="hybrid",
                    quality_weight_terms=("range", "beam", "vertical"), range_decay_m=230000.0,
                    beam_radius_ref_m=3000.0, vertical_gap_ref_deg=0.5, beam_widths=None, effective_earth_radius=None,
                    sanity_filter=True, outputs=("valid_count", "coverage_count", "blind_mask", "CR"), rows=None,
                    allow_empty=False, missing=None, column_products=None, composite=True):
    """All radars of one field merged in one fused kernel launch.

    ``volumes``: packed :class:`DeviceVolume` per radar (field 0 is used).  ``rows=(r0, r1)``
    restricts the work to a row slab of the grid (multi-GPU sharding); the returned arrays then
    have ``r1 - r0`` rows.  Returns a dict with ``composite`` and the requested ``outputs``
    (``valid_count`` int16, ``coverage_count`` int16, ``blind_mask`` int8, ``CR`` with NaN where
    no echo, ``quality`` = weight of the last radar) plus ``tie_voxels``.  ``allow_empty``: an empty
    radar list (a slab no radar reaches) yields the all-fill result instead of the reference's error.
    ``missing``: value stored in empty composite cells instead of ``fillvalue`` (``np.nan`` = the dataset layout
    of ``build_network_dataset``, applied on the device).  ``column_products=(vil_min_dbz, vil_max_dbz_cap,
    et_threshold_dbz)`` adds ``VIL`` / ``ET`` / ``ET_TOPPED`` of the composite, computed before it leaves the
    device; ``composite=False`` then skips returning the volume itself.
    """
    method = _check_method(method)
    gx = _f64_2d(grid_x, "grid_x")
    gy = _f64_2d(grid_y, "grid_y")
    if gy.shape != gx.shape:
        raise ValueError("grid_x and grid_y must have the same shape.")
    if rows is not None:
        r0, r1 = int(rows[0]), int(rows[1])
        gx = np.ascontiguousarray(gx[r0:r1])
        gy = np.ascontiguousarray(gy[r0:r1])
    levels = np.ascontiguousarray(np.asarray(level_heights, dtype=np.float64).ravel())
    n = len(volumes)
    if n == 0:
        if not allow_empty:
            raise ValueError("radar_volumes must contain at least one volume.")
        shape = (int(levels.size),) + gx.shape
        want = set(outputs or ())
        res = {"composite": np.full(shape, float(fillvalue if missing is None else missing)), "tie_voxels": 0}
        if column_products is not None:
            res["VIL"] = np.full(gx.shape, np.nan)
            res["ET"] = np.full(gx.shape, np.nan)
            res["ET_TOPPED"] = np.zeros(gx.shape, dtype=np.uint8)
        if "valid_count" in want:
            res["valid_count"] = np.zeros(shape, dtype=np.int16)
        if "coverage_count" in want or "blind_mask" in want:
            res["coverage_count"] = np.zeros(shape, dtype=np.int16)
            res["blind_mask"] = np.ones(shape, dtype=np.int8)
        if "CR" in want:
            res["CR"] = np.full(gx.shape, np.nan)
        if "quality" in want:
            res["quality"] = np.zeros(shape)
        return res
    sites = np.asarray(radar_sites_xy, dtype=np.float64).reshape(n, 2)
    rx = np.ascontiguousarray(sites[:, 0])
    ry = np.ascontiguousarray(sites[:, 1])
    rh = np.ascontiguousarray(np.asarray(radar_heights, dtype=np.float64).ravel())
    if rh.size != n:
        raise ValueError("radar_heights must have one entry per radar.")
    if effective_earth_radius is None or np.isscalar(effective_earth_radius):
        radii = [effective_earth_radius] * n
    else:
        radii = list(effective_earth_radius)
        if len(radii) != n:
            raise ValueError("effective_earth_radius must be None, a scalar, or a sequence matching the radar count.")
    re = np.array([G._resolve_effective_earth_radius(r) for r in radii], dtype=np.float64)
    prm = _lib.NetworkParams(
        _lib.NET_METHODS[method], G._resolve_blind_code(blind_method),
        sum(_lib.QUALITY_TERMS[t] for t in set(quality_weight_terms or []) if t in _lib.QUALITY_TERMS),
        1 if sanity_filter else 0, float(influence_radius_m), float(range_decay_m), float(beam_radius_ref_m),
        float(vertical_gap_ref_deg))
    keep = []
    bwp = None
    if beam_widths is not None:
        ptrs = []
        for i, bw in enumerate(beam_widths):
            if bw is None:
                ptrs.append(None)
                continue
            b = np.asarray(bw, dtype=np.float64).ravel()
            ne = volumes[i].ne
            if b.size < ne:
                b = np.concatenate([b, np.full(ne - b.size, float(b[-1]))])
            b = np.ascontiguousarray(b[:ne])
            keep.append(b)
            ptrs.append(b.ctypes.data_as(_lib.c_double_p))
        bwp = (_lib.c_double_p * n)(*ptrs)
    nz, (nx, ny) = int(levels.size), gx.shape
    shape = (nz, nx, ny)
    want = set(outputs or ())
    res = {}
    if composite or column_products is None:
        res["composite"] = pinned_empty(shape, np.float64)
    cpar = None
    if column_products is not None:
        cpar = np.ascontiguousarray(column_products, dtype=np.float64).ravel()
        if cpar.size != 3:
            raise ValueError("column_products must be (vil_min_dbz, vil_max_dbz_cap, et_threshold_dbz)")
        res["VIL"] = np.full((nx, ny), np.nan)
        res["ET"] = np.full((nx, ny), np.nan)
        res["ET_TOPPED"] = np.zeros((nx, ny), dtype=np.uint8)
    if "valid_count" in want:
        res["valid_count"] = pinned_empty(shape, np.int16)
    if "coverage_count" in want or "blind_mask" in want:
        res["coverage_count"] = pinned_empty(shape, np.int16)
        res["blind_mask"] = pinned_empty(shape, np.int8)
    if "CR" in want:
        res["CR"] = np.empty((nx, ny), dtype=np.float64)
    if "quality" in want:
        res["quality"] = pinned_empty(shape, np.float64)
    ties = ctypes.c_uint64(0)
    handles = (ctypes.c_void_p * n)(*[v.handle for v in volumes])

    def ptr(key):
        return _lib.as_ptr(res[key]) if key in res else None

    if gx.size and (nz or "CR" in want):
        _lib.check(_lib.load().pcg_network_compose_ex(
            n, handles, rx.ctypes.data_as(_lib.c_double_p), ry.ctypes.data_as(_lib.c_double_p),
            rh.ctypes.data_as(_lib.c_double_p), re.ctypes.data_as(_lib.c_double_p), bwp, ctypes.byref(prm),
            _lib.as_ptr(gx), _lib.as_ptr(gy), nx, ny, levels.ctypes.data_as(_lib.c_double_p), nz, float(fillvalue),
            float(fillvalue if missing is None else missing),
            ptr("composite"), ptr("valid_count"), ptr("coverage_count"), ptr("blind_mask"), ptr("CR"), ptr("quality"),
            ctypes.byref(ties), None if cpar is None else cpar.ctypes.data_as(_lib.c_double_p),
            ptr("VIL"), ptr("ET"), ptr("ET_TOPPED")))
    res["tie_voxels"] = int(ties.value)
    return res


# ------------------------------------------------------------------------------------------------
# run_radar_network_3d: array-level core with the reference's option names, defaults and layout
# ------------------------------------------------------------------------------------------------

class SimpleDataset(dict):
    """Stand-in with xarray's Dataset layout when xarray is not installed: ``ds[name]`` is a
    :class:`SimpleVariable` with ``dims``, ``values`` and ``attrs``; ``coords`` and ``attrs`` dicts."""

    def __init__(self, coords):
        super().__init__()
        self.coords = _Coords()
        for k, v in coords.items():
            self.coords[k] = v
        self.attrs = {}

    def __setitem__(self, name, value):
        if isinstance(value, tuple):
            value = SimpleVariable(value[0], np.asarray(value[1]))
        super().__setitem__(name, value)


class _Coords(dict):
    def __setitem__(self, name, value):
        if not isinstance(value, SimpleVariable):
            value = SimpleVariable((name,), np.asarray(value))
        super().__setitem__(name, value)


class SimpleVariable(object):
    def __init__(self, dims, values):
        self.dims = tuple(dims)
        self.values = values
        self.attrs = {}


def _new_dataset(coords):
    try:
        import xarray as xr
        return xr.Dataset(coords=coords)
    except ImportError:
        return SimpleDataset(coords)


def _project(lon, lat, lon_0, lat_0):
    """aeqd about (lon_0, lat_0): pyproj (WGS84) like the reference (RadarInterp.py:1811-1812) when it is
    installed, else the spherical formula of pycwr/core/transforms.py:660-732; returns (x, y, which)."""
    try:
        import pyproj
        proj = pyproj.Proj({"proj": "aeqd", "lon_0": lon_0, "lat_0": lat_0})
        x, y = proj(lon, lat, inverse=False)
        return np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), "pyproj"
    except ImportError:
        from .synth import aeqd_forward
        x, y = aeqd_forward(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64), lon_0, lat_0)
        return x, y, "spherical"


def _coerce_product_level_heights(level_heights, product_level_heights):
    """RadarInterp.py:421-438."""
    level_heights = np.asarray(level_heights, dtype=np.float64)
    if level_heights.size == 0:
        return level_heights
    if product_level_heights is None or len(product_level_heights) == 0:
        if level_heights.size == 1:
            return level_heights.copy()
        return np.arange(float(level_heights.min()), float(level_heights.max()) + 250.0, 500.0, dtype=np.float64)
    return np.asarray(product_level_heights, dtype=np.float64)


def run_radar_network_3d(radars, radar_lonlat, lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg,
                         level_heights, product_level_heights=None, field_names=("dBZ",),
                         output_products=("CR", "VIL", "ET"), composite_method=None, influence_radius_m=None,
                         fillvalue=None, effective_earth_radius=None, field_range_mode=None, max_range_km=None,
                         blind_method=None, quality_weight_terms=None, range_decay_m=None, beam_radius_ref_m=None,
                         vertical_gap_ref_deg=None, vil_min_dbz=None, vil_max_dbz_cap=None, et_threshold_dbz=None,
                         lon_0=None, lat_0=None, rows=None):
    """Array-level core of ``run_radar_network_3d`` (RadarInterp.py:1610-2193).

    ``radars``: PRD-like objects (``get_vol_data`` / ``.vol`` / ``.scan_info``), ``radar_lonlat``:
    their (lon, lat).  Every other option has the reference's name and default.  Returns the
    product dataset: per field ``(z, lat, lon)`` with fill -> NaN, ``coverage_count`` int16,
    ``blind_mask`` int8, ``CR`` / ``VIL`` / ``ET`` / ``ET_topped`` ``(lat, lon)``.
    """
    def opt(value, key):
        return DEFAULTS[key] if value is None else value

    composite_method = _check_method(opt(composite_method, "composite_method"))
    influence_radius_m = float(opt(influence_radius_m, "influence_radius_m"))
    fillvalue = float(opt(fillvalue, "fillvalue"))
    blind_method = opt(blind_method, "blind_method")
    max_range_km = opt(max_range_km, "max_range_km")
    terms = list(opt(quality_weight_terms, "quality_weight_terms"))
    field_names = [str(f) for f in field_names]
    for f in field_names:                    # RadarInterp.py:461-468
        if f.upper() in ("V", "VR", "VELOCITY"):
            raise ValueError("3D network mosaic does not currently support velocity fields; "
                             "remove %s from field_names." % f)
    output_products = [str(p).upper() for p in output_products]
    for p in output_products:
        if p not in ("CR", "VIL", "ET"):
            raise ValueError("Unsupported output product %s." % p)
    if not radars:
        raise ValueError("No radar files matched the target time within tolerance.")
    level_heights = np.asarray(level_heights, dtype=np.float64)
    product_levels = _coerce_product_level_heights(level_heights, product_level_heights)
    lon, lat, grid_lon, grid_lat = build_latlon_grid(lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg)
    lonlat = np.asarray(radar_lonlat, dtype=np.float64).reshape(len(radars), 2)
    if lon_0 is None:
        lon_0 = float(np.mean(lonlat[:, 0]))
    if lat_0 is None:
        lat_0 = float(np.mean(lonlat[:, 1]))
    grid_x, grid_y, which = _project(grid_lon, grid_lat, lon_0, lat_0)
    sx, sy, _ = _project(lonlat[:, 0], lonlat[:, 1], lon_0, lat_0)
    sites = np.stack([np.atleast_1d(sx), np.atleast_1d(sy)], axis=1)
    range_modes = dict(field_range_mode or {})
    ds = _new_dataset({"z": level_heights, "lat": lat, "lon": lon})
    kw = dict(fillvalue=fillvalue, method=composite_method, influence_radius_m=influence_radius_m,
              blind_method=blind_method, quality_weight_terms=terms,
              range_decay_m=float(opt(range_decay_m, "range_decay_m")),
              beam_radius_ref_m=float(opt(beam_radius_ref_m, "beam_radius_ref_m")),
              vertical_gap_ref_deg=float(opt(vertical_gap_ref_deg, "vertical_gap_ref_deg")),
              effective_earth_radius=effective_earth_radius, rows=rows)
    prod = None
    for field in field_names:
        mode = range_modes.get(field, "native" if field == "dBZ" else None)   # RadarInterp.py:271-272
        vols, heights, bws, actual, owned = [], [], [], {}, []
        for i, prd in enumerate(radars):
            dv, h, used, mine = _prd_device_volume(prd, field, fillvalue, mode, max_range_km)
            vols.append(dv)
            owned.append(mine)
            heights.append(h)
            bws.append(_get_sweep_beam_widths(prd, dv.ne))
            actual[str(i)] = float(np.nanmax(dv.last_gate)) / 1000.0
            range_modes[field] = used
        try:
            is_ref = field == "dBZ"
            want_cols = is_ref and ("VIL" in output_products or "ET" in output_products) and product_levels.size
            same_levels = bool(want_cols) and np.array_equal(product_levels, level_heights)
            cpar = (float(opt(vil_min_dbz, "vil_min_dbz")), float(opt(vil_max_dbz_cap, "vil_max_dbz_cap")),
                    float(opt(et_threshold_dbz, "et_threshold_dbz")))
            # the composite leaves the device with NaN in its empty cells (build_network_dataset's np.where,
            # RadarInterp.py:1539) and with its column products already derived (no host pass, no re-upload)
            res = network_compose(vols, sites, heights, grid_x, grid_y, level_heights, beam_widths=bws,
                                  outputs=("coverage_count", "blind_mask", "CR") if is_ref else (), missing=np.nan,
                                  column_products=cpar if same_levels else None, **kw)
            ds[field] = (("z", "lat", "lon"), np.asarray(res["composite"]))
            ds[field].attrs = {
                "standard_name": "radar_volume_mosaic", "long_name": "%s 3D radar network mosaic" % field,
                "coordinates": "z lat lon", "range_mode": range_modes[field],
                "requested_max_range_km": "full" if max_range_km is None else float(max_range_km),
                "actual_max_range_km_by_radar": json.dumps(actual, ensure_ascii=False), "qc_applied": "false",
            }
            if is_ref:
                ds["coverage_count"] = (("z", "lat", "lon"), res["coverage_count"])
                ds["coverage_count"].attrs = {"long_name": "number_of_radars_with_real_observability",
                                              "coordinates": "z lat lon"}
                ds["blind_mask"] = (("z", "lat", "lon"), res["blind_mask"])
                ds["blind_mask"].attrs = {"long_name": "blind_zone_mask", "flag_values": [0, 1],
                                          "flag_meanings": "covered blind", "coordinates": "z lat lon"}
                if "CR" in output_products:
                    ds["CR"] = (("lat", "lon"), res["CR"])
                    ds["CR"].attrs = {"units": "dBZ", "long_name": "network_composite_reflectivity",
                                      "composite_method": "max"}
                if same_levels:
                    prod = res
                elif want_cols:                                                  # RadarInterp.py:1971-2000
                    prod = network_compose(vols, sites, heights, grid_x, grid_y, product_levels, beam_widths=bws,
                                           outputs=(), column_products=cpar, composite=False, **kw)
        finally:
            for v, mine in zip(vols, owned):
                if mine:
                    v.close()
    if prod is not None and ("VIL" in output_products or "ET" in output_products):
        levels_json = json.dumps(product_levels.tolist())
        if "VIL" in output_products:             # RadarInterp.py:2064-2084
            ds["VIL"] = (("lat", "lon"), prod["VIL"])
            ds["VIL"].attrs = {"units": "kg m-2", "long_name": "vertically_integrated_liquid",
                               "min_dbz": float(opt(vil_min_dbz, "vil_min_dbz")),
                               "max_dbz_cap": float(opt(vil_max_dbz_cap, "vil_max_dbz_cap")),
                               "source_levels": levels_json}
        if "ET" in output_products:              # RadarInterp.py:2085-2119
            thr = float(opt(et_threshold_dbz, "et_threshold_dbz"))
            ds["ET"] = (("lat", "lon"), prod["ET"])
            ds["ET"].attrs = {"units": "m", "long_name": "echo_top_height", "threshold_dbz": thr,
                              "source_levels": levels_json}
            ds["ET_TOPPED"] = (("lat", "lon"), prod["ET_TOPPED"])
            ds["ET_TOPPED"].attrs = {"units": "1", "long_name": "echo_top_topped_flag",
                                     "flag_values": np.array([0, 1], dtype=np.uint8),
                                     "flag_meanings": "resolved topped_at_highest_sampled_level",
                                     "threshold_dbz": thr, "source_levels": levels_json}
    ds.attrs = {
        "product_type": "radar_network_3d", "radar_count": len(radars), "composite_method": composite_method,
        "influence_radius_m": influence_radius_m, "fillvalue": fillvalue, "projection": "aeqd",
        "projection_lon_0": float(lon_0), "projection_lat_0": float(lat_0), "projection_backend": which,
        "field_range_mode": json.dumps(range_modes, ensure_ascii=False), "blind_method": blind_method,
        "requested_max_range_km": "full" if max_range_km is None else float(max_range_km), "use_qc": "false",
        "quality_weight_terms": json.dumps(terms),
        "effective_earth_radius": "default" if effective_earth_radius is None else float(effective_earth_radius),
    }
    return ds
