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

import threading

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


def _prd_device_volume(prd, field_name, fillvalue, range_mode, max_range_km, device=None):
    """(DeviceVolume, radar height, range mode used, owned): the PRD's cached device volume when it offers one
    (``get_device_volume``: packed by the GPU from the raw sweeps), else a fresh upload of ``prd.vol`` — on the
    calling thread's current device (``device``: the same index, for the PRD's per-device cache)."""
    if hasattr(prd, "get_device_volume"):
        if hasattr(prd, "_resolve_field_range_mode"):
            range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
        kw = {} if device is None else {"device": device}
        dv = prd.get_device_volume(field_name=field_name, fillvalue=fillvalue, range_mode=range_mode,
                                   max_range_km=max_range_km, **kw)
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
                    allow_empty=False, missing=None, column_products=None, composite=True, devices=None):
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

    ``devices=[d0, d1, ...]`` (with :class:`~pycwr_b200.runtime.HostVolume` volumes): ONE call over several GPUs of
    the box — the grid's rows are split into slabs balanced by the (column, radar) pairs they hold, one host thread
    per GPU packs the radars reaching its slab, composites it and returns its rows over that GPU's own PCIe link into
    the shared pinned result arrays (no device-to-device exchange: the consumer of the product is the host).  The
    device analogue of the reference's fork pool (RadarInterp.py:1816-1861); results equal the single-GPU call's
    bit for bit.
    """
    method = _check_method(method)
    if devices is not None:
        if rows is not None:
            raise ValueError("rows= and devices= are two ways of sharding: give one")
        return _network_compose_devices(
            list(devices), volumes, radar_sites_xy, radar_heights, grid_x, grid_y, level_heights, fillvalue, method,
            influence_radius_m, blind_method, quality_weight_terms, range_decay_m, beam_radius_ref_m,
            vertical_gap_ref_deg, beam_widths, effective_earth_radius, sanity_filter, outputs, missing,
            column_products, composite)
    volumes = [v.on() if hasattr(v, "on") else v for v in volumes]
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
        # (page-locked like the volumes: the library overwrites every cell when the grid is not empty)
        res["VIL"] = pinned_empty((nx, ny), np.float64)
        res["ET"] = pinned_empty((nx, ny), np.float64)
        res["ET_TOPPED"] = pinned_empty((nx, ny), np.uint8)
        if nx * ny == 0 or nz == 0:
            res["VIL"][...] = np.nan
            res["ET"][...] = np.nan
            res["ET_TOPPED"][...] = 0
    if "valid_count" in want:
        res["valid_count"] = pinned_empty(shape, np.int16)
    if "coverage_count" in want or "blind_mask" in want:
        res["coverage_count"] = pinned_empty(shape, np.int16)
        res["blind_mask"] = pinned_empty(shape, np.int8)
    if "CR" in want:
        res["CR"] = pinned_empty((nx, ny), np.float64)
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


_DEVICE_SHARES = {}      # device list -> share of the grid's cost each device takes (measured, see below)


def _network_compose_devices(devices, volumes, radar_sites_xy, radar_heights, grid_x, grid_y, level_heights, fillvalue,
                             method, influence_radius_m, blind_method, quality_weight_terms, range_decay_m,
                             beam_radius_ref_m, vertical_gap_ref_deg, beam_widths, effective_earth_radius,
                             sanity_filter, outputs, missing, column_products, composite):
    """``network_compose`` over several GPUs of one box: row slabs, one host thread per GPU, every slab written by
    its own GPU into the shared pinned result (``pcg_network_compose_rows``)."""
    import threading

    from . import sharding
    if not devices:
        raise ValueError("devices must name at least one GPU")
    n = len(volumes)
    if n == 0:
        raise ValueError("radar_volumes must contain at least one volume.")
    if any(not hasattr(v, "on") for v in volumes):
        raise ValueError("devices= needs HostVolume objects (each GPU packs the radars its slab needs)")
    gx = _f64_2d(grid_x, "grid_x")
    gy = _f64_2d(grid_y, "grid_y")
    if gy.shape != gx.shape:
        raise ValueError("grid_x and grid_y must have the same shape.")
    levels = np.ascontiguousarray(np.asarray(level_heights, dtype=np.float64).ravel())
    sites = np.ascontiguousarray(np.asarray(radar_sites_xy, dtype=np.float64).reshape(n, 2))
    rh = np.ascontiguousarray(np.asarray(radar_heights, dtype=np.float64).ravel())
    if rh.size != n:
        raise ValueError("radar_heights must have one entry per radar.")
    if effective_earth_radius is None or np.isscalar(effective_earth_radius):
        radii = [effective_earth_radius] * n
    else:
        radii = list(effective_earth_radius)
        if len(radii) != n:
            raise ValueError("effective_earth_radius must be None, a scalar, or a sequence matching the radar count.")
    re_all = np.array([G._resolve_effective_earth_radius(r) for r in radii], dtype=np.float64)
    prm = _lib.NetworkParams(
        _lib.NET_METHODS[method], G._resolve_blind_code(blind_method),
        sum(_lib.QUALITY_TERMS[t] for t in set(quality_weight_terms or []) if t in _lib.QUALITY_TERMS),
        1 if sanity_filter else 0, float(influence_radius_m), float(range_decay_m), float(beam_radius_ref_m),
        float(vertical_gap_ref_deg))
    nz, (nx, ny) = int(levels.size), gx.shape
    shape = (nz, nx, ny)
    want = set(outputs or ())
    miss = float(fillvalue if missing is None else missing)
    res = {}
    if composite or column_products is None:
        res["composite"] = pinned_empty(shape, np.float64)
    cpar = None
    if column_products is not None:
        cpar = np.ascontiguousarray(column_products, dtype=np.float64).ravel()
        if cpar.size != 3:
            raise ValueError("column_products must be (vil_min_dbz, vil_max_dbz_cap, et_threshold_dbz)")
        res["VIL"] = pinned_empty((nx, ny), np.float64)
        res["ET"] = pinned_empty((nx, ny), np.float64)
        res["ET_TOPPED"] = pinned_empty((nx, ny), np.uint8)
    if "valid_count" in want:
        res["valid_count"] = pinned_empty(shape, np.int16)
    if "coverage_count" in want or "blind_mask" in want:
        res["coverage_count"] = pinned_empty(shape, np.int16)
        res["blind_mask"] = pinned_empty(shape, np.int8)
    if "CR" in want:
        res["CR"] = pinned_empty((nx, ny), np.float64)
    reach = [sharding.radar_reach_m(v.vol_range) for v in volumes]
    costs = None
    if len(devices) == 1:
        bounds = [(0, nx)]          # (the kernel culls radars per tile: nothing to balance or to drop here)
    else:
        # The cost estimate only balances the slabs: 17 columns of the grid are enough for it (a contiguous copy —
        # reductions over a strided view of the 72 MB arrays cost 18 ms).  The shares of the devices start equal and
        # follow the rates measured on the previous call with the same device list: on the GPU boxes of this pool the
        # device-to-host path of GPUs 0-3 is about half as fast as that of GPUs 4-7 when all eight return results
        # (profiles/r02_host_link.md), and a slab is done when its rows are back in host memory.
        cols = np.unique(np.linspace(0, ny - 1, min(ny, 17)).astype(np.int64)) if ny else np.zeros(0, np.int64)
        costs = sharding.row_costs(np.take(gx, cols, axis=1), np.take(gy, cols, axis=1), sites, reach)
        shares = _DEVICE_SHARES.get(tuple(devices))
        if shares is None or len(shares) != len(devices):
            shares = np.full(len(devices), 1.0 / len(devices))
        bounds = sharding.slab_bounds_shares(costs, shares)
    elapsed = [0.0] * len(devices)
    ties = [0] * len(devices)
    errors = []
    import os
    import time
    trace = [None] * len(devices) if os.environ.get("PCG_PYTRACE") else None      # per-thread timeline (diagnostic)
    t_call = time.perf_counter()

    def ptr(key):
        return _lib.as_ptr(res[key]) if key in res else None

    def slab(di):
        try:
            t_in = time.perf_counter()
            r0, r1 = bounds[di]
            if r1 <= r0:
                return
            idx = list(range(n)) if len(devices) == 1 else sharding.radars_reaching(gx[r0:r1], gy[r0:r1], sites, reach)
            if not idx or nx * ny == 0:
                for key, val in (("composite", miss), ("valid_count", 0), ("coverage_count", 0), ("blind_mask", 1),
                                 ("CR", np.nan), ("VIL", np.nan), ("ET", np.nan), ("ET_TOPPED", 0)):
                    if key in res:
                        res[key][..., r0:r1, :] = val
                return
            _lib.set_device(devices[di])
            dvs = [volumes[i].on(devices[di]) for i in idx]
            m = len(idx)
            keep = []
            bwp = None
            if beam_widths is not None:
                ptrs = []
                for i, dv in zip(idx, dvs):
                    bw = beam_widths[i]
                    if bw is None:
                        ptrs.append(None)
                        continue
                    b = np.asarray(bw, dtype=np.float64).ravel()
                    if b.size < dv.ne:
                        b = np.concatenate([b, np.full(dv.ne - b.size, float(b[-1]))])
                    b = np.ascontiguousarray(b[:dv.ne])
                    keep.append(b)
                    ptrs.append(b.ctypes.data_as(_lib.c_double_p))
                bwp = (_lib.c_double_p * m)(*ptrs)
            rx = np.ascontiguousarray(sites[idx, 0])
            ry = np.ascontiguousarray(sites[idx, 1])
            hh = np.ascontiguousarray(rh[idx])
            rr = np.ascontiguousarray(re_all[idx])
            handles = (ctypes.c_void_p * m)(*[v.handle for v in dvs])
            t = ctypes.c_uint64(0)
            dp = _lib.c_double_p
            t_prep = time.perf_counter()
            _lib.check(_lib.load().pcg_network_compose_rows(
                m, handles, rx.ctypes.data_as(dp), ry.ctypes.data_as(dp), hh.ctypes.data_as(dp), rr.ctypes.data_as(dp),
                bwp, ctypes.byref(prm), _lib.as_ptr(gx), _lib.as_ptr(gy), nx, r0, r1 - r0, ny,
                levels.ctypes.data_as(dp), nz, float(fillvalue), miss, ptr("composite"), ptr("valid_count"),
                ptr("coverage_count"), ptr("blind_mask"), ptr("CR"), ctypes.byref(t),
                None if cpar is None else cpar.ctypes.data_as(dp), ptr("VIL"), ptr("ET"), ptr("ET_TOPPED")))
            ties[di] = int(t.value)
            t_out = time.perf_counter()
            elapsed[di] = t_out - t_in
            if trace is not None:
                trace[di] = (1e3 * (t_in - t_call), 1e3 * (t_prep - t_in), 1e3 * (t_out - t_prep), r1 - r0, m)
        except BaseException as exc:                # noqa: BLE001 - re-raised by the caller's thread below
            errors.append(exc)

    prev = _lib.current_device()
    if len(devices) == 1:
        slab(0)
    else:
        threads = [threading.Thread(target=slab, args=(di,), name="pycwr-b200-gpu%d" % devices[di])
                   for di in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    _lib.set_device(prev)
    if errors:
        raise errors[0]
    if costs is not None and len(set(devices)) == len(devices):
        # next call's shares: the cost each device got through per second, damped (the rates depend on who else is
        # copying at the same time) and floored so that no device is starved by one slow call
        done = np.array([float(costs[b[0]:b[1]].sum()) for b in bounds])
        secs = np.array(elapsed)
        if np.all(secs > 0.0) and np.all(done > 0.0):
            rate = done / secs
            new = 0.5 * shares / shares.sum() + 0.5 * rate / rate.sum()
            new = np.maximum(new, 0.25 / len(devices))
            _DEVICE_SHARES[tuple(devices)] = new / new.sum()
    if trace is not None:
        print("[pcg pytrace] alloc+bounds %.2f ms; per device (start, prep, call ms, rows, radars): %s; total %.2f ms" % (
            1e3 * 0, ["%.1f/%.1f/%.1f/%d/%d" % t if t else "-" for t in trace], 1e3 * (time.perf_counter() - t_call)),
            flush=True)
    res["tie_voxels"] = int(sum(ties))
    res["slab_rows"] = [b[1] - b[0] for b in bounds]
    return res


# ------------------------------------------------------------------------------------------------
# dataset stand-ins and projection
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

    @property
    def data_vars(self):
        """The variable names (xarray: a mapping whose iteration yields them)."""
        return list(self.keys())


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


# ------------------------------------------------------------------------------------------------
# run_radar_network_3d: the reference's signature, option resolution and dataset layout
# ------------------------------------------------------------------------------------------------

class _NetworkConfig(dict):
    """``cfg.network`` of pycwr/configure/config.py:12-62: a dict with attribute access."""
    __getattr__ = dict.get

    def __setattr__(self, key, value):
        self[key] = value


NETWORK_DEFAULTS = {             # pycwr/configure/config.py:12-62
    "radar_dirs": [], "field_names": ["dBZ"], "field_range_mode": {"dBZ": "native"},
    "lon_min": None, "lon_max": None, "lat_min": None, "lat_max": None, "lon_res_deg": 0.01, "lat_res_deg": 0.01,
    "level_heights": [], "product_level_heights": [], "max_range_km": 460.0, "time_tolerance_minutes": 10,
    "composite_method": "quality_weighted", "influence_radius_m": 300000.0,
    "quality_weight_terms": ["range", "beam", "vertical", "qc", "blockage"], "range_decay_m": 230000.0,
    "beam_radius_ref_m": 3000.0, "vertical_gap_ref_deg": 0.5, "qc_weight_field": "QC_MASK", "blockage_config": None,
    "fillvalue": -999.0, "output_dir": None, "effective_earth_radius": None, "file_pattern": "*.bin*",
    "parallel": True, "max_workers": None, "blind_method": "hybrid", "output_products": ["CR", "VIL", "ET"],
    "plot_overview": True, "plot_style": "reference", "plot_canvas_px": [920, 790], "plot_height_levels": [],
    "plot_output_dir": None, "plot_format": "png", "plot_product_for_levels": "CAPPI",
    "coverage_policy": "observable_only", "vil_min_dbz": 18.0, "vil_max_dbz_cap": 56.0, "et_threshold_dbz": 18.0,
    "use_qc": False, "qc_method": "dualpol", "qc_band": "C", "qc_use_existing_kdp": True, "qc_fallback": "original",
    "qc_clear_air_mode": "label", "qc_clear_air_max_ref": 15.0, "qc_clear_air_max_rhohv": 0.97,
    "qc_clear_air_max_phidp_texture": 10.0, "qc_clear_air_max_snr": 20.0,
}


class _Cfg(object):
    """Stand-in for ``pycwr.configure.config.cfg``: only the ``network`` section matters on this path."""

    def __init__(self):
        import copy
        self.network = _NetworkConfig(copy.deepcopy(NETWORK_DEFAULTS))


cfg = _Cfg()

# The reference takes these three from pycwr.io / pyproj / the PRD (RadarInterp.py:20-27); readers are out
# of scope here, so they are module-level hooks: the reference's own when pycwr is installed,
# patched in tests exactly like test/test_examples_interp.py:198-203 patches them.
try:                                                # pragma: no cover - pycwr is not installed on the GPU box
    from pycwr.io import read_auto
    from pycwr.io.util import get_radar_info
except Exception:                                   # noqa: BLE001
    def read_auto(path, station_lon=None, station_lat=None, station_alt=None, effective_earth_radius=None):
        raise ImportError("pycwr.io.read_auto is not available: install pycwr's readers or patch "
                          "pycwr_b200.interp.read_auto with a loader returning PRD objects.")

    def get_radar_info(path):
        raise ImportError("pycwr.io.util.get_radar_info is not available.")

FUSED_NETWORK = True             # False: always take the reference's per-radar flow (tests compare the two)
_PRODUCT_REFERENCE_NOTES = {
    "VIL": [
        "Greene and Clark (1972), Monthly Weather Review, doi:10.1175/1520-0493(1972)100<0548:VILWNA>2.3.CO;2",
        "NOAA WDTD VIL product guidance: https://vlab.noaa.gov/web/wdtd/-/vertically-integrated-liquid-vil-",
    ],
    "ET": [
        "Lakshmanan et al. (2013), Weather and Forecasting, doi:10.1175/WAF-D-12-00084.1",
        "NOAA WDTD xx dBZ Echo Top guidance: https://vlab.noaa.gov/web/wdtd/-/xx-dbz-echo-top-et-",
        "WSR-88D ROC ICD 2620003Y (Enhanced Echo Tops): https://www.roc.noaa.gov/public-documents/icds/2620003Y.pdf",
    ],
    "MOSAIC": [
        "吴翀;双偏振雷达的资料质量分析,相态识別及组网应用[D];"
        "南京信息工程大学;2018年。相关组网讨论见第4章。",
        "Lakshmanan et al. (2006), Weather and Forecasting, doi:10.1175/WAF942.1",
    ],
}


def _resolve_network_value(name, value, network_cfg=None):
    """Explicit argument, else the loaded config file, else ``cfg.network`` (RadarInterp.py:205-215)."""
    if value is not None:
        return value
    if network_cfg is None:
        network_cfg = getattr(cfg, "network", None)
    if network_cfg is None:
        return None
    if isinstance(network_cfg, dict):
        return network_cfg.get(name)
    return getattr(network_cfg, name, None)


def _read_network_config_file(config_path):
    """JSON / TOML / YAML network options; a top-level ``network`` table is unwrapped (RadarInterp.py:218-248)."""
    from pathlib import Path
    config_path = Path(config_path)
    suffix = config_path.suffix.lower()
    if suffix == ".json":
        with config_path.open("r", encoding="utf-8") as handle:
            data = json.load(handle)
    elif suffix == ".toml":
        try:
            import tomllib
        except ImportError as exc:              # pragma: no cover
            raise ValueError("TOML configuration requires Python 3.11+ or tomllib.") from exc
        with config_path.open("rb") as handle:
            data = tomllib.load(handle)
    elif suffix in (".yml", ".yaml"):
        try:
            import yaml
        except ImportError as exc:              # pragma: no cover
            raise ValueError("YAML configuration requires PyYAML to be installed.") from exc
        with config_path.open("r", encoding="utf-8") as handle:
            data = yaml.safe_load(handle)
    else:
        raise ValueError("Unsupported config file suffix %s; use .json, .toml, .yml, or .yaml." % suffix)
    if data is None:
        return {}
    if not isinstance(data, dict):
        raise ValueError("Network config file must contain a mapping/object.")
    if "network" in data:
        data = data["network"]
    if not isinstance(data, dict):
        raise ValueError("The 'network' section must contain a mapping/object.")
    return data


def load_network_config(config_path, inplace=False):
    """RadarInterp.py:251-256."""
    network_config = _read_network_config_file(config_path)
    if inplace:
        cfg.network.update(network_config)
    return network_config


def _coerce_field_names(field_names, network_cfg=None):
    """RadarInterp.py:191-202."""
    if field_names is None:
        field_names = _resolve_network_value("field_names", None, network_cfg=network_cfg)
    if field_names is None:
        return ["dBZ"]
    if isinstance(field_names, str):
        field_names = [field_names]
    field_names = [str(name) for name in field_names if str(name)]
    if not field_names:
        raise ValueError("field_names must contain at least one field.")
    return field_names


def _coerce_field_range_mode(field_names, field_range_mode, network_cfg=None):
    """Per-field range mode; dBZ is never left on ``aligned`` (RadarInterp.py:259-278)."""
    field_names = _coerce_field_names(field_names, network_cfg=network_cfg)
    if field_range_mode is None:
        field_range_mode = _resolve_network_value("field_range_mode", None, network_cfg=network_cfg)
    modes = dict.fromkeys(field_names, "aligned")
    if isinstance(field_range_mode, str):
        modes = dict.fromkeys(field_names, field_range_mode)
    elif field_range_mode:
        for name in field_names:
            if name in field_range_mode:
                modes[name] = str(field_range_mode[name])
    if modes.get("dBZ") == "aligned":
        modes["dBZ"] = "native"
    for name, mode in modes.items():
        if mode not in ("aligned", "native"):
            raise ValueError("range_mode for %s must be 'aligned' or 'native'." % name)
    return modes


def _resolve_qc_field_name(field_name, use_qc):
    """RadarInterp.py:281-290."""
    if not use_qc:
        return field_name
    return {"dBZ": "Zc", "ZDR": "ZDRc", "KDP": "KDPc"}.get(field_name, field_name)


def _normalize_blind_method(blind_method):
    """RadarInterp.py:293-301."""
    blind_method = "mask" if blind_method is None else str(blind_method).lower()
    if blind_method not in ("mask", "nearest_gate", "lowest_sweep", "hybrid", "nearest_valid"):
        raise ValueError("blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', "
                         "'hybrid', or 'nearest_valid'.")
    return blind_method


def _radar_supports_dualpol_qc(prd):
    """RadarInterp.py:304-312."""
    for sweep in range(prd.nsweeps):
        dataset = prd.fields[sweep]
        if "dBZ" not in dataset:
            return False
        if ("KDP" not in dataset) and ("PhiDP" not in dataset):
            return False
    return True


def _get_prd_station_info(prd):
    """(lat, lon, altitude) from PRD metadata (RadarInterp.py:315-325)."""
    scan_info = getattr(prd, "scan_info", None)
    if scan_info is None:
        raise ValueError("Radar object does not provide scan_info.")
    altitude = float(scan_info["altitude"].values) if "altitude" in scan_info else 0.0
    return float(scan_info["latitude"].values), float(scan_info["longitude"].values), altitude


def parse_radar_time_from_filename(path):
    """The first 14-digit group of the file name is the scan time (RadarInterp.py:326-331)."""
    import re
    from datetime import datetime
    from pathlib import Path
    match = re.search(r"(\d{14})", Path(path).name)
    if match is None:
        raise ValueError("Unable to parse scan time from filename: %s" % path)
    return datetime.strptime(match.group(1), "%Y%m%d%H%M%S")


def discover_radar_files(radar_dirs, pattern="*.bin*"):
    """RadarInterp.py:334-346."""
    from pathlib import Path
    if not radar_dirs:
        raise ValueError("radar_dirs must contain at least one directory.")
    radar_files = {}
    for radar_dir in radar_dirs:
        directory = Path(radar_dir)
        if not directory.is_dir():
            continue
        matches = sorted(path for path in directory.glob(pattern) if path.is_file())
        if matches:
            radar_files[directory.name] = matches
    return radar_files


def select_radar_files(radar_dirs, target_time, tolerance_minutes=10, pattern="*.bin*"):
    """Nearest-in-time file of every radar directory (RadarInterp.py:349-379)."""
    from datetime import datetime
    if target_time is None:
        raise ValueError("target_time is required.")
    if isinstance(target_time, str):
        target_time = datetime.fromisoformat(target_time)
    limit = float(tolerance_minutes) * 60.0
    selected = []
    for radar_id, files in discover_radar_files(radar_dirs, pattern=pattern).items():
        scored = [(abs((parse_radar_time_from_filename(p) - target_time).total_seconds()), i) for i, p in enumerate(files)]
        if not scored:
            continue
        delta, i = min(scored)                   # first of equal deltas, like the reference's strict '<'
        if delta <= limit:
            selected.append({"radar_id": radar_id, "path": str(files[i]),
                             "scan_time": parse_radar_time_from_filename(files[i]), "delta_seconds": delta})
    if not selected:
        raise ValueError("No radar files matched the target time within tolerance.")
    return selected


def _coerce_output_products(output_products, network_cfg=None):
    """RadarInterp.py:396-409."""
    output_products = _resolve_network_value("output_products", output_products, network_cfg=network_cfg)
    if output_products is None:
        return ["CR", "VIL", "ET"]
    if isinstance(output_products, str):
        output_products = [output_products]
    valid = []
    for name in output_products:
        name = str(name).upper()
        if name not in ("CR", "VIL", "ET"):
            raise ValueError("Unsupported output product %s." % name)
        if name not in valid:
            valid.append(name)
    return valid


def _coerce_plot_height_levels(level_heights, plot_height_levels, network_cfg=None):
    """RadarInterp.py:412-418."""
    plot_height_levels = _resolve_network_value("plot_height_levels", plot_height_levels, network_cfg=network_cfg)
    if level_heights is None:
        return np.asarray([], dtype=np.float64)
    if plot_height_levels is None or len(plot_height_levels) == 0:
        return np.asarray(level_heights, dtype=np.float64)
    return np.asarray(plot_height_levels, dtype=np.float64)


def _coerce_product_level_heights(level_heights, product_level_heights, network_cfg=None):
    """RadarInterp.py:421-438."""
    product_level_heights = _resolve_network_value("product_level_heights", product_level_heights,
                                                   network_cfg=network_cfg)
    if level_heights is None:
        return np.asarray([], dtype=np.float64)
    level_heights = np.asarray(level_heights, dtype=np.float64)
    if level_heights.size == 0:
        return level_heights
    if product_level_heights is None or len(product_level_heights) == 0:
        if level_heights.size == 1:
            return level_heights.copy()
        return np.arange(float(level_heights.min()), float(level_heights.max()) + 250.0, 500.0, dtype=np.float64)
    return np.asarray(product_level_heights, dtype=np.float64)


def _coerce_quality_weight_terms(quality_weight_terms, network_cfg=None):
    """RadarInterp.py:441-458."""
    quality_weight_terms = _resolve_network_value("quality_weight_terms", quality_weight_terms, network_cfg=network_cfg)
    if quality_weight_terms is None:
        quality_weight_terms = ["range", "beam", "vertical", "qc", "blockage"]
    if isinstance(quality_weight_terms, str):
        quality_weight_terms = [quality_weight_terms]
    valid = []
    for term in quality_weight_terms:
        normalized = str(term).strip().lower()
        if normalized not in ("range", "beam", "vertical", "qc", "blockage"):
            raise ValueError("Unsupported quality weight term %s." % term)
        if normalized not in valid:
            valid.append(normalized)
    return valid


def _validate_network_field_names(field_names):
    """RadarInterp.py:461-468."""
    for field_name in field_names:
        if str(field_name).upper() in ("V", "VR", "VELOCITY"):
            raise ValueError("3D network mosaic does not currently support velocity fields; "
                             "remove %s from field_names." % field_name)


def radar_network_3d_to_netcdf(dataset, output_path):
    """RadarInterp.py:1593-1598 (needs an xarray dataset; file output is outside the gridding path)."""
    from pathlib import Path
    output_path = Path(output_path)
    if not hasattr(dataset, "to_netcdf"):
        raise ImportError("writing netCDF needs xarray (the dataset is the in-memory stand-in).")
    output_path.parent.mkdir(parents=True, exist_ok=True)
    dataset.to_netcdf(str(output_path))
    return str(output_path)


def _default_network_output_path(output_dir, target_time, field_names):
    """RadarInterp.py:1601-1607."""
    from pathlib import Path
    return str(Path(output_dir) / ("radar_network_3d_%s_%s.nc" % (target_time.strftime("%Y%m%d%H%M%S"),
                                                                    "_".join(field_names))))


class _Radar(object):
    """One selected radar of a network run: the PRD, where it sits on the shared grid, its QC state."""

    def __init__(self, radar_id, prd, lon, lat, x, y, path=None, scan_time=None):
        self.radar_id, self.prd, self.lon, self.lat, self.x, self.y = radar_id, prd, lon, lat, x, y
        self.path, self.scan_time = path, scan_time
        self.qc_applied, self.qc_error = False, None


def _apply_qc(radar, qc_method, qc_band, qc_use_existing_kdp, qc_fallback, clear_air):
    """The QC hook of the per-radar worker (RadarInterp.py:1117-1142): the QC itself is the PRD's method."""
    if qc_method != "dualpol":
        raise ValueError("Unsupported qc_method: %s" % qc_method)
    if _radar_supports_dualpol_qc(radar.prd):
        try:
            radar.prd = radar.prd.apply_dualpol_qc(inplace=True, band=qc_band, use_existing_kdp=qc_use_existing_kdp,
                                                   **clear_air)
            radar.qc_applied = True
        except Exception as exc:                    # noqa: BLE001 - the reference swallows every QC failure here
            if qc_fallback != "original":
                raise
            radar.qc_error = str(exc)
    else:
        if qc_fallback != "original":
            raise ValueError("Radar %s does not have the moments required for dual-pol QC." % radar.radar_id)
        radar.qc_error = "missing_dualpol_fields"


def _blockage_of(radar, blockage_config, grid_shape):
    """Per-radar blockage weight: scalar or a grid-shaped array clipped to [0, 1] (RadarInterp.py:1257-1269)."""
    value = 1.0
    if blockage_config:
        value = blockage_config.get(radar.radar_id, 1.0)
    arr = np.asarray(value, dtype=np.float64)
    if arr.ndim == 0:
        return float(arr), None
    if arr.shape != tuple(grid_shape):
        raise ValueError("blockage_config for %s must match the target grid shape." % radar.radar_id)
    return None, np.clip(arr, 0.0, 1.0)


def _qc_weight_grid(radar, qc_weight_field, grid_x, grid_y, levels, fillvalue, eff, max_range_km):
    """Gridded QC mask clipped to [0, 1] (RadarInterp.py:1232-1256)."""
    try:
        qc_grid = grid_single_radar_to_latlon_3d(radar.prd, qc_weight_field, grid_x, grid_y, levels, radar.x, radar.y,
                                                 fillvalue=fillvalue, effective_earth_radius=eff, range_mode="aligned",
                                                 max_range_km=max_range_km, blind_method="mask")
    except KeyError:
        return None
    return np.clip(np.where(np.isfinite(qc_grid) & (qc_grid != fillvalue), qc_grid, 0.0), 0.0, 1.0)


def _network_field_unfused(radars, field, source_names, mode, o, levels, want_obs, want_cr):
    """One field the way the reference's per-radar worker grids it (RadarInterp.py:1158-1356) — every step is a
    GPU call, but every radar's full volume crosses PCIe.  Used when the fused kernel cannot express the request:
    QC-mask or array blockage weights, a scalar blockage other than 1, or a volume outside the packed layout
    (one sweep, duplicate / unsorted table entries, a fill value float32 cannot hold)."""
    nrad = len(radars)
    shape = (int(levels.size),) + o["grid_x"].shape
    vols, quals, actual, by_sweep, obs, crs = [], [], {}, {}, [], []
    for radar in radars:
        prd, src = radar.prd, source_names[radar.radar_id]
        eff = getattr(prd, "effective_earth_radius", o["effective_earth_radius"])
        try:
            grid, meta = grid_single_radar_to_latlon_3d(
                prd, src, o["grid_x"], o["grid_y"], levels, radar.x, radar.y, fillvalue=o["fillvalue"],
                effective_earth_radius=eff, range_mode=mode, max_range_km=o["max_range_km"],
                blind_method=o["blind_method"], return_metadata=True)
            actual[radar.radar_id] = meta["actual_max_range_km"]
            by_sweep[radar.radar_id] = meta.get("actual_max_range_km_by_sweep")
            prd.get_vol_data(field_name=src, fillvalue=o["fillvalue"], range_mode=mode, max_range_km=o["max_range_km"])
            _, vol_range, fix_elevation, _, radar_height = prd.vol[:5]
            bw = _get_sweep_beam_widths(prd, len(fix_elevation))
            qual = _compute_quality_weight_volume(
                vol_range, fix_elevation, radar_height, o["grid_x"], o["grid_y"], levels, radar.x, radar.y,
                beam_widths=bw, effective_earth_radius=eff, quality_weight_terms=o["terms"],
                range_decay_m=o["range_decay_m"], beam_radius_ref_m=o["beam_radius_ref_m"],
                vertical_gap_ref_deg=o["vertical_gap_ref_deg"])
            if want_cr:
                crs.append(grid_single_radar_cr_to_latlon(
                    prd, src, o["grid_x"], o["grid_y"], radar.x, radar.y, fillvalue=o["fillvalue"],
                    effective_earth_radius=eff, range_mode=mode, max_range_km=o["max_range_km"]))
            if want_obs:
                obs.append(_compute_observability_mask(vol_range, fix_elevation, radar_height, o["grid_x"], o["grid_y"],
                                                       levels, radar.x, radar.y, beam_widths=bw,
                                                       effective_earth_radius=eff))
            if radar.qc_applied and ("qc" in o["terms"]) and o["qc_weight_field"]:
                qcw = _qc_weight_grid(radar, o["qc_weight_field"], o["grid_x"], o["grid_y"], levels, o["fillvalue"], eff,
                                      o["max_range_km"])
                if qcw is not None:
                    qual = qual * qcw
            if "blockage" in o["terms"]:
                scalar, arr = _blockage_of(radar, o["blockage_config"], o["grid_x"].shape)
                qual = qual * (scalar if arr is None else arr[np.newaxis, :, :])
        except KeyError:                            # the radar does not carry this field (RadarInterp.py:1338-1347)
            grid = np.full(shape, o["fillvalue"], dtype=np.float64)
            qual = np.zeros(shape, dtype=np.float64)
            actual[radar.radar_id] = None
            by_sweep[radar.radar_id] = None
            if want_cr:
                crs.append(np.full(o["grid_x"].shape, np.nan))
            if want_obs:
                obs.append(np.zeros(shape, dtype=bool))
        vols.append(np.asarray(grid, dtype=np.float64))
        quals.append(np.asarray(qual, dtype=np.float64))
    sites = np.array([[r.x, r.y] for r in radars], dtype=np.float64).reshape(nrad, 2)
    composite, _ = compose_network_volume(vols, sites, o["grid_x"], o["grid_y"], fillvalue=o["fillvalue"],
                                          method=o["method"], influence_radius_m=o["influence_radius_m"],
                                          quality_weights=quals)
    res = {"composite": composite, "actual": actual, "by_sweep": by_sweep}
    if want_obs:                                    # RadarInterp.py:1938-1945
        count = np.sum(np.stack(obs, axis=0), axis=0, dtype=np.int16)
        res["coverage_count"] = count.astype(np.int16)
        res["blind_mask"] = (count == 0).astype(np.int8)
    if want_cr:                                     # RadarInterp.py:2047-2063
        stack = np.stack([np.asarray(c, dtype=np.float64) for c in crs], axis=0)
        ok = np.isfinite(stack) & (stack != o["fillvalue"])
        res["CR"] = np.where(ok.any(axis=0), np.max(np.where(ok, stack, -np.inf), axis=0), np.nan)
    return res


def _resolve_devices(parallel, max_workers):
    """The reference's ``parallel`` / ``max_workers`` (RadarInterp.py:1816-1824: a fork pool over radars) mapped to the
    GPUs of the box: a device list for ``network_compose(devices=...)`` or None for the current device.
    ``PYCWR_B200_DEVICES="0,1,..."`` overrides the choice (tests, pinning a job to a subset of the GPUs)."""
    import os
    env = os.environ.get("PYCWR_B200_DEVICES")
    if env:
        return [int(tok) for tok in env.split(",") if tok.strip() != ""] or None
    if not parallel:
        return None
    count = _lib.device_count()
    if max_workers is not None:
        count = min(count, max(1, int(max_workers)))
    return list(range(count)) if count > 1 else None


class _PrdVolumes(object):
    """One PRD's field as ``network_compose(devices=...)`` wants its volumes: packed on whichever GPU asks for it
    (``on(device)``; the PRD's own per-device cache when it has one), closed afterwards if it was uploaded here."""

    def __init__(self, prd, field_name, fillvalue, mode, max_range_km, first_device):
        self._args = (prd, field_name, fillvalue, mode, max_range_km)
        self._dev, self._owned = {}, []
        self._lock = threading.Lock()
        first = self.on(first_device)
        self.ne, self.network_ready = first.ne, first.network_ready
        self.last_gate = np.asarray(first.last_gate, dtype=np.float64)
        self.vol_range = [np.array([g]) for g in self.last_gate]          # what the reach of the radar is read from

    def on(self, device=None):
        device = _lib.current_device() if device is None else int(device)
        with self._lock:                                   # (two slabs of one device may ask at the same time)
            dv = self._dev.get(device)
            if dv is None:
                prev = _lib.current_device()
                _lib.set_device(device)
                try:
                    prd, field_name, fillvalue, mode, max_range_km = self._args
                    dv, self.height, _, mine = _prd_device_volume(prd, field_name, fillvalue, mode, max_range_km,
                                                                  device=device)
                finally:
                    _lib.set_device(prev)
                self._dev[device] = dv
                if mine:
                    self._owned.append(dv)
        return dv

    def close(self):
        for dv in self._owned:
            dv.close()
        self._dev, self._owned = {}, []


def _network_field(radars, field, source_names, mode, o, levels, is_ref, cpar):
    """Composite one field on ``levels``; ``cpar`` = VIL / ET parameters when the column products are wanted too.
    Returns the result dict (``composite`` with NaN in its empty cells, coverage / blind / CR for the reflectivity
    field, ``VIL`` / ``ET`` / ``ET_TOPPED``) plus the per-radar range metadata."""
    want_cr = is_ref and o["want_cr"]
    fused = FUSED_NETWORK
    if "blockage" in o["terms"] and o["blockage_config"]:
        for radar in radars:
            scalar, arr = _blockage_of(radar, o["blockage_config"], o["grid_x"].shape)
            fused = fused and arr is None and scalar == 1.0
    if "qc" in o["terms"] and o["qc_weight_field"] and any(r.qc_applied for r in radars):
        fused = False
    present, vols, owned, heights, bws, radii, actual, by_sweep = [], [], [], [], [], [], {}, {}
    devices = o.get("devices") if o.get("rows") is None else None
    if fused:
        for radar in radars:
            try:
                if devices:
                    dv = _PrdVolumes(radar.prd, source_names[radar.radar_id], o["fillvalue"], mode, o["max_range_km"],
                                     devices[0])
                    h, mine = dv.height, True
                else:
                    dv, h, _, mine = _prd_device_volume(radar.prd, source_names[radar.radar_id], o["fillvalue"], mode,
                                                        o["max_range_km"])
            except KeyError:
                actual[radar.radar_id] = None
                by_sweep[radar.radar_id] = None
                continue
            vols.append(dv); owned.append(mine); heights.append(h); present.append(radar)
            bws.append(_get_sweep_beam_widths(radar.prd, dv.ne))
            radii.append(getattr(radar.prd, "effective_earth_radius", o["effective_earth_radius"]))
            last = np.asarray(dv.last_gate, dtype=np.float64)
            actual[radar.radar_id] = float(np.nanmax(last)) / 1000.0
            by_sweep[radar.radar_id] = [float(v) / 1000.0 for v in last]
            if not dv.network_ready:
                fused = False
    try:
        if fused and present:
            sites = np.array([[r.x, r.y] for r in present], dtype=np.float64)
            res = network_compose(
                vols, sites, heights, o["grid_x"], o["grid_y"], levels, fillvalue=o["fillvalue"], method=o["method"],
                influence_radius_m=o["influence_radius_m"], blind_method=o["blind_method"],
                quality_weight_terms=o["terms"], range_decay_m=o["range_decay_m"],
                beam_radius_ref_m=o["beam_radius_ref_m"], vertical_gap_ref_deg=o["vertical_gap_ref_deg"],
                beam_widths=bws, effective_earth_radius=radii, rows=o.get("rows"),
                outputs=(("coverage_count", "blind_mask") + (("CR",) if want_cr else ())) if is_ref else (),
                missing=np.nan, column_products=cpar, devices=devices or None)
            res["actual"], res["by_sweep"] = actual, by_sweep
            return res
    finally:
        for v, mine in zip(vols, owned):
            if mine:
                v.close()
    res = _network_field_unfused(radars, field, source_names, mode, o, levels, is_ref, want_cr)
    comp = res["composite"]
    if cpar is not None:
        from .products import column_products
        res.update({k: v for k, v in column_products(comp, levels, o["fillvalue"], cpar[0], cpar[1], cpar[2]).items()
                    if k != "CR"})
    res["composite"] = np.where(comp == o["fillvalue"], np.nan, comp)       # RadarInterp.py:1539
    return res


def _network_dataset(radars, o, field_names, field_range_mode, level_heights, product_levels, output_products, lon, lat):
    """Everything of run_radar_network_3d after the radars are read: per-field composites, coverage, CR / VIL / ET,
    assembled with the reference's names, dims and attrs (RadarInterp.py:1490-1590,1938-2119)."""
    level_heights = np.asarray(level_heights, dtype=np.float64)
    ds = _new_dataset({"z": level_heights, "lat": lat, "lon": lon})
    for name, attrs in (("z", {"units": "m", "long_name": "height_above_mean_sea_level"}),
                        ("lat", {"units": "degree_north", "long_name": "latitude"}),
                        ("lon", {"units": "degree_east", "long_name": "longitude"})):
        ds.coords[name].attrs = attrs
    o = dict(o, want_cr="CR" in output_products)
    cpar = (float(o["vil_min_dbz"]), float(o["vil_max_dbz_cap"]), float(o["et_threshold_dbz"]))
    want_cols = ("dBZ" in field_names) and product_levels.size > 0 and ("VIL" in output_products or "ET" in output_products)
    same_levels = want_cols and np.array_equal(product_levels, level_heights)
    actual_by_field, by_sweep_by_field, source_by_field = {}, {}, {}
    prod = None
    corrected = {"dBZ": "Zc", "ZDR": "ZDRc", "KDP": "KDPc"}
    for field in field_names:
        is_ref = field == "dBZ"
        source_names = {r.radar_id: _resolve_qc_field_name(field, r.qc_applied) for r in radars}
        source_by_field[field] = source_names
        res = _network_field(radars, field, source_names, field_range_mode[field], o, level_heights, is_ref,
                             cpar if (is_ref and same_levels) else None)
        actual_by_field[field], by_sweep_by_field[field] = res["actual"], res["by_sweep"]
        flags = [bool(r.qc_applied) and corrected.get(field) == source_names[r.radar_id] for r in radars]
        ds[field] = (("z", "lat", "lon"), np.asarray(res["composite"]))
        ds[field].attrs = {
            "standard_name": "radar_volume_mosaic", "long_name": "%s 3D radar network mosaic" % field,
            "coordinates": "z lat lon", "range_mode": field_range_mode[field],
            "requested_max_range_km": "full" if o["max_range_km"] is None else float(o["max_range_km"]),
            "actual_max_range_km_by_radar": json.dumps(res["actual"], ensure_ascii=False),
            "actual_max_range_km_by_radar_and_sweep": json.dumps(res["by_sweep"], ensure_ascii=False),
            "source_field_name_by_radar": json.dumps(source_names, ensure_ascii=False),
            "qc_applied": "true" if (flags and all(flags)) else ("partial" if any(flags) else "false"),
        }
        if not is_ref:
            continue
        ds["coverage_count"] = (("z", "lat", "lon"), np.asarray(res["coverage_count"]))
        ds["coverage_count"].attrs = {"long_name": "number_of_radars_with_real_observability",
                                      "coverage_policy": o["coverage_policy"], "coordinates": "z lat lon"}
        ds["blind_mask"] = (("z", "lat", "lon"), np.asarray(res["blind_mask"]))
        ds["blind_mask"].attrs = {"long_name": "blind_zone_mask", "flag_values": [0, 1],
                                  "flag_meanings": "covered blind", "coverage_policy": o["coverage_policy"],
                                  "coordinates": "z lat lon"}
        if "CR" in output_products:
            ds["CR"] = (("lat", "lon"), res["CR"])
            ds["CR"].attrs = {"units": "dBZ", "long_name": "network_composite_reflectivity", "composite_method": "max"}
        if same_levels:
            prod = res
        elif want_cols:                              # RadarInterp.py:1971-2000: the products' own level set
            prod = _network_field(radars, field, source_names, field_range_mode[field], dict(o, want_cr=False),
                                  product_levels, False, cpar)
    if "dBZ" not in field_names:                     # RadarInterp.py:1938-1945 without a reflectivity field
        shape = (level_heights.size,) + o["grid_x"].shape
        ds["coverage_count"] = (("z", "lat", "lon"), np.zeros(shape, dtype=np.int16))
        ds["blind_mask"] = (("z", "lat", "lon"), np.ones(shape, dtype=np.int8))
    if prod is not None:
        levels_json = json.dumps(product_levels.tolist())
        if "VIL" in output_products:                 # RadarInterp.py:2064-2084
            ds["VIL"] = (("lat", "lon"), prod["VIL"])
            ds["VIL"].attrs = {
                "units": "kg m-2", "long_name": "vertically_integrated_liquid", "min_dbz": cpar[0],
                "max_dbz_cap": cpar[1], "source_levels": levels_json,
                "references": json.dumps(_PRODUCT_REFERENCE_NOTES["VIL"], ensure_ascii=False),
                "implementation_note": ("Uses Greene and Clark (1972) liquid-water relation and a 56 dBZ cap; "
                                        "the low-reflectivity cutoff is a configurable pycwr workflow parameter."),
            }
        if "ET" in output_products:                  # RadarInterp.py:2085-2119
            refs = json.dumps(_PRODUCT_REFERENCE_NOTES["ET"], ensure_ascii=False)
            ds["ET"] = (("lat", "lon"), prod["ET"])
            ds["ET"].attrs = {
                "units": "m", "long_name": "echo_top_height", "threshold_dbz": cpar[2], "source_levels": levels_json,
                "references": refs,
                "implementation_note": ("Highest threshold-exceeding layer plus linear interpolation to the crossing "
                                        "height; see ET_TOPPED for columns that still exceed the threshold at the "
                                        "highest sampled level."),
            }
            ds["ET_TOPPED"] = (("lat", "lon"), np.asarray(prod["ET_TOPPED"], dtype=np.uint8))
            ds["ET_TOPPED"].attrs = {"units": "1", "long_name": "echo_top_topped_flag",
                                     "flag_values": np.array([0, 1], dtype=np.uint8),
                                     "flag_meanings": "resolved topped_at_highest_sampled_level",
                                     "threshold_dbz": cpar[2], "source_levels": levels_json, "references": refs}
    return ds, actual_by_field, source_by_field


def run_network_arrays(radars, radar_lonlat, lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg,
                       level_heights, product_level_heights=None, field_names=("dBZ",),
                       output_products=("CR", "VIL", "ET"), composite_method=None, influence_radius_m=None,
                       fillvalue=None, effective_earth_radius=None, field_range_mode=None, max_range_km=None,
                       blind_method=None, quality_weight_terms=None, range_decay_m=None, beam_radius_ref_m=None,
                       vertical_gap_ref_deg=None, vil_min_dbz=None, vil_max_dbz_cap=None, et_threshold_dbz=None,
                       lon_0=None, lat_0=None, rows=None, radar_ids=None, blockage_config=None):
    """``run_radar_network_3d`` from the point where the radars are in memory: ``radars`` are PRD-like objects
    (``get_vol_data`` / ``.vol`` / ``.scan_info``, optionally ``get_device_volume``), ``radar_lonlat`` their
    (lon, lat).  Every option has the reference's name and default; ``rows=(r0, r1)`` restricts the work to a row
    slab of the grid (multi-GPU sharding).  Returns the product dataset."""
    def opt(value, key):
        return DEFAULTS[key] if value is None else value

    if not radars:
        raise ValueError("No radar files matched the target time within tolerance.")
    field_names = _coerce_field_names(list(field_names) if not isinstance(field_names, str) else field_names, {})
    _validate_network_field_names(field_names)
    output_products = _coerce_output_products(list(output_products), {})
    modes = _coerce_field_range_mode(field_names, field_range_mode, {})
    level_heights = np.asarray(level_heights, dtype=np.float64)
    product_levels = _coerce_product_level_heights(level_heights, product_level_heights, {})
    lon, lat, grid_lon, grid_lat = build_latlon_grid(lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg)
    lonlat = np.asarray(radar_lonlat, dtype=np.float64).reshape(len(radars), 2)
    lon_0 = float(np.mean(lonlat[:, 0])) if lon_0 is None else lon_0
    lat_0 = float(np.mean(lonlat[:, 1])) if lat_0 is None else lat_0
    grid_x, grid_y, which = _project(grid_lon, grid_lat, lon_0, lat_0)
    sx, sy, _ = _project(lonlat[:, 0], lonlat[:, 1], lon_0, lat_0)
    sx, sy = np.atleast_1d(sx), np.atleast_1d(sy)
    ids = [str(i) for i in range(len(radars))] if radar_ids is None else [str(i) for i in radar_ids]
    items = [_Radar(ids[i], prd, lonlat[i, 0], lonlat[i, 1], float(sx[i]), float(sy[i])) for i, prd in enumerate(radars)]
    o = {
        "grid_x": grid_x, "grid_y": grid_y, "fillvalue": float(opt(fillvalue, "fillvalue")),
        "method": _check_method(opt(composite_method, "composite_method")),
        "influence_radius_m": float(opt(influence_radius_m, "influence_radius_m")),
        "blind_method": _normalize_blind_method(opt(blind_method, "blind_method")),
        "max_range_km": opt(max_range_km, "max_range_km"),
        "terms": _coerce_quality_weight_terms(list(opt(quality_weight_terms, "quality_weight_terms")), {}),
        "range_decay_m": float(opt(range_decay_m, "range_decay_m")),
        "beam_radius_ref_m": float(opt(beam_radius_ref_m, "beam_radius_ref_m")),
        "vertical_gap_ref_deg": float(opt(vertical_gap_ref_deg, "vertical_gap_ref_deg")),
        "effective_earth_radius": effective_earth_radius, "qc_weight_field": "QC_MASK",
        "blockage_config": blockage_config, "coverage_policy": "observable_only",
        "vil_min_dbz": opt(vil_min_dbz, "vil_min_dbz"), "vil_max_dbz_cap": opt(vil_max_dbz_cap, "vil_max_dbz_cap"),
        "et_threshold_dbz": opt(et_threshold_dbz, "et_threshold_dbz"), "rows": rows,
    }
    if rows is not None:
        lat = lat[int(rows[0]):int(rows[1])]
    ds, _, _ = _network_dataset(items, o, field_names, modes, level_heights, product_levels, output_products, lon, lat)
    ds.attrs = {
        "product_type": "radar_network_3d", "radar_count": len(radars), "composite_method": o["method"],
        "influence_radius_m": o["influence_radius_m"], "fillvalue": o["fillvalue"], "projection": "aeqd",
        "projection_lon_0": float(lon_0), "projection_lat_0": float(lat_0), "projection_backend": which,
        "field_range_mode": json.dumps(modes, ensure_ascii=False), "blind_method": o["blind_method"],
        "requested_max_range_km": "full" if o["max_range_km"] is None else float(o["max_range_km"]), "use_qc": "false",
        "quality_weight_terms": json.dumps(o["terms"]),
        "effective_earth_radius": "default" if effective_earth_radius is None else float(effective_earth_radius),
    }
    return ds


def run_radar_network_3d(
    target_time,
    config_path=None,
    radar_dirs=None,
    lon_min=None,
    lon_max=None,
    lat_min=None,
    lat_max=None,
    lon_res_deg=None,
    lat_res_deg=None,
    level_heights=None,
    product_level_heights=None,
    field_names=None,
    output_products=None,
    output_path=None,
    time_tolerance_minutes=None,
    composite_method=None,
    influence_radius_m=None,
    fillvalue=None,
    effective_earth_radius=None,
    field_range_mode=None,
    max_range_km=None,
    parallel=None,
    max_workers=None,
    blind_method=None,
    use_qc=None,
    qc_method=None,
    qc_band=None,
    qc_use_existing_kdp=None,
    qc_fallback=None,
    qc_clear_air_mode=None,
    qc_clear_air_max_ref=None,
    qc_clear_air_max_rhohv=None,
    qc_clear_air_max_phidp_texture=None,
    qc_clear_air_max_snr=None,
    plot_overview=None,
    plot_style=None,
    plot_canvas_px=None,
    plot_height_levels=None,
    plot_output_dir=None,
    plot_format=None,
    plot_product_for_levels=None,
    coverage_policy=None,
    quality_weight_terms=None,
    range_decay_m=None,
    beam_radius_ref_m=None,
    vertical_gap_ref_deg=None,
    qc_weight_field=None,
    blockage_config=None,
    vil_min_dbz=None,
    vil_max_dbz_cap=None,
    et_threshold_dbz=None,
    pattern=None,
    lon_0=None,
    lat_0=None,
):
    """The reference's 3-D radar-network workflow with its signature, three-level option resolution (argument,
    config file, ``cfg.network``) and dataset layout (pycwr/interp/RadarInterp.py:1610-2193).

    What differs is where the work runs: the selected volumes are read on the host (``read_auto``), packed to the
    device, and every field is composited by the fused network kernels in one call (``network_compose``) instead
    of the fork pool of per-radar NumPy workers (:1816-1861).  ``parallel`` / ``max_workers`` keep their meaning one
    level down: with ``parallel`` true and more than one GPU in the box the composite fans out over
    ``min(max_workers, GPUs)`` devices (``network_compose(devices=...)``: row slabs, one host thread per GPU, every
    slab returned by its own GPU), otherwise it runs on the current device.  Plot options are accepted and recorded but no figure is drawn (presentation is out of
    scope): ``plot_files`` is ``{}``.
    """
    from datetime import datetime
    if isinstance(target_time, str):
        target_time = datetime.fromisoformat(target_time)
    if target_time is None:
        raise ValueError("target_time is required.")
    ncfg = load_network_config(config_path) if config_path is not None else getattr(cfg, "network", None)

    def rv(name, value):
        return _resolve_network_value(name, value, network_cfg=ncfg)

    radar_dirs = rv("radar_dirs", radar_dirs)
    lon_min, lon_max = rv("lon_min", lon_min), rv("lon_max", lon_max)
    lat_min, lat_max = rv("lat_min", lat_min), rv("lat_max", lat_max)
    lon_res_deg, lat_res_deg = rv("lon_res_deg", lon_res_deg), rv("lat_res_deg", lat_res_deg)
    level_heights = rv("level_heights", level_heights)
    time_tolerance_minutes = rv("time_tolerance_minutes", time_tolerance_minutes)
    composite_method = rv("composite_method", composite_method) or "quality_weighted"
    influence_radius_m = rv("influence_radius_m", influence_radius_m) or 300000.0
    fillvalue = rv("fillvalue", fillvalue)
    effective_earth_radius = rv("effective_earth_radius", effective_earth_radius)
    max_range_km = rv("max_range_km", max_range_km)
    pattern = rv("file_pattern", pattern) or "*.bin*"
    parallel, max_workers = rv("parallel", parallel), rv("max_workers", max_workers)
    blind_method = rv("blind_method", blind_method)
    use_qc = rv("use_qc", use_qc)
    qc_method = rv("qc_method", qc_method) or "dualpol"
    qc_band = rv("qc_band", qc_band) or "C"
    qc_use_existing_kdp = rv("qc_use_existing_kdp", qc_use_existing_kdp)
    qc_fallback = rv("qc_fallback", qc_fallback) or "original"
    qc_clear_air_mode = rv("qc_clear_air_mode", qc_clear_air_mode) or "label"
    qc_clear_air_max_ref = rv("qc_clear_air_max_ref", qc_clear_air_max_ref)
    qc_clear_air_max_rhohv = rv("qc_clear_air_max_rhohv", qc_clear_air_max_rhohv)
    qc_clear_air_max_phidp_texture = rv("qc_clear_air_max_phidp_texture", qc_clear_air_max_phidp_texture)
    qc_clear_air_max_snr = rv("qc_clear_air_max_snr", qc_clear_air_max_snr)
    output_products = _coerce_output_products(output_products, network_cfg=ncfg)
    plot_overview = rv("plot_overview", plot_overview)
    plot_style = rv("plot_style", plot_style) or "reference"
    plot_canvas_px = rv("plot_canvas_px", plot_canvas_px)
    plot_height_levels = _coerce_plot_height_levels(level_heights, plot_height_levels, network_cfg=ncfg)
    product_level_heights = _coerce_product_level_heights(level_heights, product_level_heights, network_cfg=ncfg)
    coverage_policy = rv("coverage_policy", coverage_policy) or "observable_only"
    quality_weight_terms = _coerce_quality_weight_terms(quality_weight_terms, network_cfg=ncfg)
    range_decay_m = rv("range_decay_m", range_decay_m)
    beam_radius_ref_m = rv("beam_radius_ref_m", beam_radius_ref_m)
    vertical_gap_ref_deg = rv("vertical_gap_ref_deg", vertical_gap_ref_deg)
    qc_weight_field = rv("qc_weight_field", qc_weight_field)
    blockage_config = rv("blockage_config", blockage_config)
    vil_min_dbz = rv("vil_min_dbz", vil_min_dbz)
    vil_max_dbz_cap = rv("vil_max_dbz_cap", vil_max_dbz_cap)
    et_threshold_dbz = rv("et_threshold_dbz", et_threshold_dbz)
    field_names = _coerce_field_names(field_names, network_cfg=ncfg)
    _validate_network_field_names(field_names)
    field_range_mode = _coerce_field_range_mode(field_names, field_range_mode, network_cfg=ncfg)
    output_dir = rv("output_dir", None)

    if radar_dirs is None:
        raise ValueError("radar_dirs are required.")
    if level_heights is None:
        raise ValueError("level_heights are required.")
    if None in (lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg):
        raise ValueError("lon/lat bounds and resolution are required.")
    fillvalue = -999.0 if fillvalue is None else fillvalue
    qc_use_existing_kdp = True if qc_use_existing_kdp is None else qc_use_existing_kdp
    qc_clear_air_max_ref = 15.0 if qc_clear_air_max_ref is None else qc_clear_air_max_ref
    qc_clear_air_max_rhohv = 0.97 if qc_clear_air_max_rhohv is None else qc_clear_air_max_rhohv
    qc_clear_air_max_phidp_texture = 10.0 if qc_clear_air_max_phidp_texture is None else qc_clear_air_max_phidp_texture
    qc_clear_air_max_snr = 20.0 if qc_clear_air_max_snr is None else qc_clear_air_max_snr
    use_qc = False if use_qc is None else use_qc
    if coverage_policy != "observable_only":
        raise ValueError("Only coverage_policy='observable_only' is currently supported.")
    vil_min_dbz = 18.0 if vil_min_dbz is None else vil_min_dbz
    vil_max_dbz_cap = 56.0 if vil_max_dbz_cap is None else vil_max_dbz_cap
    et_threshold_dbz = 18.0 if et_threshold_dbz is None else et_threshold_dbz
    blind_method = _normalize_blind_method(blind_method if blind_method is not None else "hybrid")
    plot_style = str(plot_style).lower()
    if plot_style not in ("reference", "simple"):
        raise ValueError("plot_style must be 'reference' or 'simple'.")
    if (output_products or plot_height_levels.size > 0) and "dBZ" not in field_names:      # RadarInterp.py:1763-1765
        field_names = list(field_names) + ["dBZ"]
        field_range_mode = _coerce_field_range_mode(field_names, field_range_mode, network_cfg=ncfg)
    if plot_style == "reference" and "CR" in output_products:                              # RadarInterp.py:742-744
        if float(lon_res_deg) > 0.015 or float(lat_res_deg) > 0.015:
            raise ValueError("reference plot style requires lon/lat resolution no coarser than 0.01 degree.")
    range_decay_m = 230000.0 if range_decay_m is None else range_decay_m
    beam_radius_ref_m = 3000.0 if beam_radius_ref_m is None else beam_radius_ref_m
    vertical_gap_ref_deg = 0.5 if vertical_gap_ref_deg is None else vertical_gap_ref_deg
    qc_weight_field = "QC_MASK" if qc_weight_field is None else qc_weight_field
    _check_method(composite_method)

    selected = select_radar_files(radar_dirs, target_time,
                                  tolerance_minutes=10 if time_tolerance_minutes is None else time_tolerance_minutes,
                                  pattern=pattern)
    stations = []
    for item in selected:                           # RadarInterp.py:1784-1795
        try:
            s_lat, s_lon, s_alt, _ = get_radar_info(item["path"])
            stations.append((float(s_lat), float(s_lon), float(s_alt)))
        except Exception:                           # noqa: BLE001 - like the reference: fall back to the file itself
            stations.append(_get_prd_station_info(read_auto(item["path"], effective_earth_radius=effective_earth_radius)))
    radar_lats = np.array([s[0] for s in stations], dtype=np.float64)
    radar_lons = np.array([s[1] for s in stations], dtype=np.float64)
    lon_0 = float(np.mean(radar_lons)) if lon_0 is None else lon_0
    lat_0 = float(np.mean(radar_lats)) if lat_0 is None else lat_0
    lon, lat, grid_lon, grid_lat = build_latlon_grid(lon_min, lon_max, lat_min, lat_max, lon_res_deg, lat_res_deg)
    grid_x, grid_y, which = _project(grid_lon, grid_lat, lon_0, lat_0)
    level_heights = np.asarray(level_heights, dtype=np.float64)

    clear_air = dict(clear_air_mode=qc_clear_air_mode, clear_air_max_ref=qc_clear_air_max_ref,
                     clear_air_max_rhohv=qc_clear_air_max_rhohv,
                     clear_air_max_phidp_texture=qc_clear_air_max_phidp_texture,
                     clear_air_max_snr=qc_clear_air_max_snr)
    radars = []
    for item, (s_lat, s_lon, s_alt) in zip(selected, stations):      # the reading half of the worker, :1104-1142
        prd = read_auto(item["path"], station_lon=s_lon, station_lat=s_lat, station_alt=s_alt,
                        effective_earth_radius=effective_earth_radius)
        x, y, _ = _project(np.float64(s_lon), np.float64(s_lat), lon_0, lat_0)
        radar = _Radar(item["radar_id"], prd, float(s_lon), float(s_lat), float(x), float(y), item["path"],
                       item["scan_time"])
        if use_qc:
            _apply_qc(radar, qc_method, qc_band, qc_use_existing_kdp, qc_fallback, clear_air)
        radars.append(radar)

    o = {
        "grid_x": grid_x, "grid_y": grid_y, "fillvalue": float(fillvalue), "method": str(composite_method),
        "influence_radius_m": float(influence_radius_m), "blind_method": blind_method, "max_range_km": max_range_km,
        "terms": quality_weight_terms, "range_decay_m": float(range_decay_m),
        "beam_radius_ref_m": float(beam_radius_ref_m), "vertical_gap_ref_deg": float(vertical_gap_ref_deg),
        "effective_earth_radius": effective_earth_radius, "qc_weight_field": qc_weight_field,
        "blockage_config": blockage_config, "coverage_policy": coverage_policy, "vil_min_dbz": vil_min_dbz,
        "vil_max_dbz_cap": vil_max_dbz_cap, "et_threshold_dbz": et_threshold_dbz,
        "devices": _resolve_devices(parallel, max_workers),
    }
    dataset, actual_by_field, source_by_field = _network_dataset(
        radars, o, field_names, field_range_mode, level_heights, product_level_heights, output_products, lon, lat)
    qc_by_radar = {r.radar_id: {"applied": bool(r.qc_applied), "error": r.qc_error} for r in radars}
    attrs = {                                       # RadarInterp.py:1558-1589
        "product_type": "radar_network_3d", "target_time": target_time.isoformat(), "radar_count": len(selected),
        "radar_files": "|".join(item["path"] for item in selected), "composite_method": composite_method,
        "influence_radius_m": float(influence_radius_m), "fillvalue": float(fillvalue), "projection": "aeqd",
        "projection_lon_0": float(lon_0), "projection_lat_0": float(lat_0),
        "field_range_mode": json.dumps(field_range_mode, ensure_ascii=False), "blind_method": blind_method,
        "requested_max_range_km": "full" if max_range_km is None else float(max_range_km),
        "actual_max_range_km_by_field": json.dumps(actual_by_field, ensure_ascii=False),
        "use_qc": "true" if use_qc else "false", "qc_method": qc_method if use_qc else "none",
        "qc_clear_air_mode": qc_clear_air_mode if use_qc else "ignore",
        "qc_clear_air_max_ref": float(qc_clear_air_max_ref), "qc_clear_air_max_rhohv": float(qc_clear_air_max_rhohv),
        "qc_clear_air_max_phidp_texture": float(qc_clear_air_max_phidp_texture),
        "qc_clear_air_max_snr": float(qc_clear_air_max_snr),
        "qc_applied_by_radar": json.dumps(qc_by_radar, ensure_ascii=False),
        "quality_weight_terms": json.dumps(quality_weight_terms),
        "algorithm_references": json.dumps(_PRODUCT_REFERENCE_NOTES["MOSAIC"], ensure_ascii=False),
        "effective_earth_radius": "default" if effective_earth_radius is None else float(effective_earth_radius),
        "network_config_path": "" if config_path is None else str(config_path),
        # RadarInterp.py:2184-2189
        "output_products": json.dumps(output_products), "coverage_policy": coverage_policy,
        "product_level_heights": json.dumps(np.asarray(product_level_heights).tolist()), "plot_style": plot_style,
        "plot_canvas_px": json.dumps(list(plot_canvas_px) if plot_canvas_px is not None else [920, 790]),
        "plot_files": json.dumps({}, ensure_ascii=False),
        "projection_backend": which,
    }
    dataset.attrs = attrs
    if output_path is None and output_dir:
        output_path = _default_network_output_path(output_dir, target_time, field_names)
    if output_path is not None:
        radar_network_3d_to_netcdf(dataset, output_path)
    return dataset
