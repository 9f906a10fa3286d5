"""Drop-in replacement for the reference's compiled module ``pycwr.core.RadarGridC``.

Exports the same eleven names with the same positional signatures, defaults and errors as
``pycwr/core/RadarGridC.pyx`` (listed in ``pycwr/core/RadarGridC.py:25-37``); the four gridding
functions run on the GPU through ``libpycwr_b200.so`` (hand-written sm_100a kernels), the
scalar helpers are a few IEEE-double operations evaluated with the host libm exactly like the
reference (they are called with Python scalars, not on the hot path).

Extensions that keep the reference call shape: wherever the reference takes the four volume
arguments ``(vol_azimuth, vol_range, fix_elevation, vol_value)``, ``vol_value`` may be a
:class:`pycwr_b200.runtime.DeviceVolume` (already resident in HBM — the other three are then
ignored), and ``return_index=True`` additionally returns the per-voxel interpolation records.
"""
import ctypes
import math

import numpy as np

from . import _lib
from .runtime import DeviceVolume, _f64_1d, _f64_2d, pinned_empty

__all__ = [
    "antenna_to_cartesian", "cartesian_to_antenna", "get_CAPPI_3d", "get_CAPPI_xy", "get_CR_xy",
    "get_mosaic_CAPPI_3d", "interp_azimuth", "interp_ppi", "ppi_to_grid", "xy_to_azimuth",
    "xye_to_antenna",
]

# RadarGridC.pyx:6-9
PI = 3.141592653589793
DEG2RAD = PI / 180.0
RAD2DEG = 180.0 / PI
DEFAULT_EFFECTIVE_EARTH_RADIUS = 8494666.6666666661

_BLIND_CODES = {"mask": 0, "nearest_gate": 1, "lowest_sweep": 2, "hybrid": 3, "nearest_valid": 4}

# results are returned in recycled page-locked memory when True (fast D2H); plain numpy otherwise
USE_PINNED_OUTPUT = True


def _resolve_effective_earth_radius(effective_earth_radius):
    # RadarGridC.pyx:12-19
    if effective_earth_radius is None:
        return DEFAULT_EFFECTIVE_EARTH_RADIUS
    radius = float(effective_earth_radius)
    if radius <= 0.0:
        raise ValueError("effective_earth_radius must be positive")
    return radius


def _resolve_blind_code(blind_method):
    # RadarGridC.pyx:54-69
    method = "mask" if blind_method is None else str(blind_method).lower()
    if method not in _BLIND_CODES:
        raise ValueError(
            "blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', "
            "'hybrid', or 'nearest_valid'."
        )
    return _BLIND_CODES[method]


def _clip_unit(value):
    if value > 1.0:
        return 1.0
    if value < -1.0:
        return -1.0
    return value


# ------------------------------------------------------------------------------------------
# scalar helpers (RadarGridC.pyx:161-233, 479-481)
# ------------------------------------------------------------------------------------------

def xy_to_azimuth(x, y):
    """RadarGridC.pyx:34-38,189-198."""
    az = PI / 2.0 - math.atan2(float(y), float(x))
    if az < 0.0:
        az += 2.0 * PI
    return az * RAD2DEG


def antenna_to_cartesian(ranges, azimuth, elevation, h, effective_earth_radius=None):
    """RadarGridC.pyx:72-95,161-168."""
    ranges, azimuth, elevation, h = float(ranges), float(azimuth), float(elevation), float(h)
    R = _resolve_effective_earth_radius(effective_earth_radius)
    theta_a = azimuth * DEG2RAD
    theta_e = elevation * DEG2RAD
    sin_e = math.sin(theta_e)
    cos_e = math.cos(theta_e)
    radar_radius = R + h
    z = math.sqrt(ranges * ranges + radar_radius * radar_radius + 2.0 * ranges * radar_radius * sin_e) - R
    arc_arg = _clip_unit(ranges * cos_e / (R + z))
    s = R * math.asin(arc_arg)
    return s * math.sin(theta_a), s * math.cos(theta_a), z


def xye_to_antenna(x, y, elevation, h, effective_earth_radius=None):
    """RadarGridC.pyx:98-129,170-177."""
    x, y, elevation, h = float(x), float(y), float(elevation), float(h)
    R = _resolve_effective_earth_radius(effective_earth_radius)
    s = math.sqrt(x * x + y * y)
    phi = s / R
    theta_e = elevation * DEG2RAD
    sin_phi = math.sin(phi)
    cos_phi = math.cos(phi)
    denom = cos_phi - sin_phi * math.tan(theta_e)
    radar_radius = R + h
    azimuth = xy_to_azimuth(x, y)
    if denom == 0.0:
        return azimuth, math.nan, math.nan
    gate_radius = radar_radius / denom
    z = gate_radius - R
    ranges_sq = radar_radius * radar_radius + gate_radius * gate_radius - 2.0 * radar_radius * gate_radius * cos_phi
    if ranges_sq < 0.0:
        ranges_sq = 0.0
    return azimuth, math.sqrt(ranges_sq), z


def cartesian_to_antenna(x, y, z, h, effective_earth_radius=None):
    """RadarGridC.pyx:132-158,179-186."""
    x, y, z, h = float(x), float(y), float(z), float(h)
    R = _resolve_effective_earth_radius(effective_earth_radius)
    s = math.sqrt(x * x + y * y)
    phi = s / R
    radar_radius = R + h
    gate_radius = R + z
    ranges_sq = radar_radius * radar_radius + gate_radius * gate_radius - 2.0 * radar_radius * gate_radius * math.cos(phi)
    if ranges_sq < 0.0:
        ranges_sq = 0.0
    ranges = math.sqrt(ranges_sq)
    if ranges == 0.0:
        elevation = 0.0
    else:
        cos_arg = _clip_unit((radar_radius * radar_radius + ranges_sq - gate_radius * gate_radius)
                             / (2.0 * radar_radius * ranges))
        elevation = (math.acos(cos_arg) - PI / 2.0) * RAD2DEG
    return xy_to_azimuth(x, y), ranges, elevation


def interp_azimuth(az, az_0, az_1, dat_0, dat_1, fillvalue):
    """RadarGridC.pyx:236-252,479-481."""
    az, az_0, az_1, dat_0, dat_1, fillvalue = map(float, (az, az_0, az_1, dat_0, dat_1, fillvalue))
    if az_1 == az_0:
        return fillvalue
    if dat_0 == fillvalue and dat_1 == fillvalue:
        return fillvalue
    if dat_0 == fillvalue:
        return dat_1
    if dat_1 == fillvalue:
        return dat_0
    return ((az_1 - az) * dat_0 + (az - az_0) * dat_1) / (az_1 - az_0)


def interp_ppi(az, r, az_0, az_1, r_0, r_1, mat_00, mat_01, mat_10, mat_11, fillvalue):
    """RadarGridC.pyx:200-233."""
    az, r, az_0, az_1, r_0, r_1, mat_00, mat_01, mat_10, mat_11, fillvalue = map(
        float, (az, r, az_0, az_1, r_0, r_1, mat_00, mat_01, mat_10, mat_11, fillvalue))
    az_span = az_1 - az_0
    r_span = r_1 - r_0
    if az_span == 0.0 or r_span == 0.0:
        return fillvalue
    f = fillvalue
    if mat_00 != f and mat_01 != f and mat_10 != f and mat_11 != f:
        return (mat_00 * (az_1 - az) * (r_1 - r) + mat_10 * (az - az_0) * (r_1 - r)
                + mat_01 * (az_1 - az) * (r - r_0) + mat_11 * (az - az_0) * (r - r_0)) / r_span / az_span
    if mat_00 != f and mat_01 != f:
        return (mat_00 * (r_1 - r) + mat_01 * (r - r_0)) / r_span
    if mat_10 != f and mat_11 != f:
        return (mat_10 * (r_1 - r) + mat_11 * (r - r_0)) / r_span
    if mat_00 != f and mat_10 != f:
        return (mat_00 * (az_1 - az) + mat_10 * (az - az_0)) / az_span
    if mat_01 != f and mat_11 != f:
        return (mat_01 * (az_1 - az) + mat_11 * (az - az_0)) / az_span
    return fillvalue


# ------------------------------------------------------------------------------------------
# gridding entry points
# ------------------------------------------------------------------------------------------

# moment storage used when a call has to pack its own volume: "f32" = production layout
# (exact indices/masks, values within a few float32 ulps), "f64" = reference arithmetic
VALUE_DTYPE = "f32"


def _as_volume(vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue, dtype=None):
    if isinstance(vol_value, DeviceVolume):
        return vol_value, False
    return DeviceVolume(vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue=float(fillvalue),
                        value_dtype=dtype or VALUE_DTYPE), True


def _redo_in_f64(rc, vol, owned, args):
    """NaN / inf target coordinates (flagged by the packed kernels): the reference's result then hinges
    on how every comparison and lerp treats NaN operand by operand; only the float64 layout restates it
    operation by operation, so a volume this call packed itself is packed again that way."""
    if rc != _lib.PCG_WARN_NONFINITE or not owned or vol.value_dtype != "f32":
        return None
    vol.close()
    return _as_volume(*args, dtype="f64")[0]


def _new_output(shape):
    if USE_PINNED_OUTPUT:
        return pinned_empty(shape, np.float64)
    return np.empty(shape, dtype=np.float64)


def ppi_to_grid(azimuth, ranges, elevation, mat_ppi, radar_height, GridX, GridY, fillvalue,
                effective_earth_radius=None, blind_method="mask", return_index=False):
    """Grid one PPI sweep onto a Cartesian plane (RadarGridC.pyx:303-352)."""
    azimuth = _f64_1d(azimuth, "azimuth")
    ranges = _f64_1d(ranges, "ranges")
    mat_ppi = _f64_2d(mat_ppi, "mat_ppi")
    gx = _f64_2d(GridX, "GridX")
    gy = _f64_2d(GridY, "GridY")
    radius = _resolve_effective_earth_radius(effective_earth_radius)
    blind_code = _resolve_blind_code(blind_method)
    nx, ny = gx.shape[0], gy.shape[1]
    out = _new_output((nx, ny))
    idx = np.empty((nx, ny, _lib.IDX_STRIDE), dtype=np.int32) if return_index else None
    if azimuth.shape[0] < 2 or ranges.shape[0] < 2:
        out.fill(fillvalue)          # RadarGridC.pyx:337-339
        if idx is not None:
            idx[...] = -1
            idx[..., 0] = 3
            idx[..., 1:3] = 0
            idx[..., 5] = 16
        return (out, idx) if return_index else out
    vol = DeviceVolume([azimuth], [ranges], np.array([float(elevation)]), [mat_ppi], fillvalue=float(fillvalue),
                       value_dtype="f64")
    try:
        _lib.check(_lib.load().pcg_ppi_to_grid(
            vol.handle, 0, float(elevation), float(radar_height), radius, _lib.as_ptr(gx), _lib.as_ptr(gy),
            nx, ny, float(fillvalue), blind_code, _lib.as_ptr(out),
            _lib.as_ptr(idx) if idx is not None else None))
    finally:
        vol.close()
    return (out, idx) if return_index else out


def get_CR_xy(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY, fillvalue,
              effective_earth_radius=None, *, missing=None):
    """Composite reflectivity: max over sweeps, fill ignored (RadarGridC.pyx:596-624).
    ``missing`` (extension, keyword-only): value stored in empty cells instead of ``fillvalue`` —
    ``np.nan`` fuses the wrappers' ``np.where(GridV == fillvalue, np.nan, GridV)`` into the device call."""
    gx = _f64_2d(GridX, "GridX")
    gy = _f64_2d(GridY, "GridY")
    radius = _resolve_effective_earth_radius(effective_earth_radius)
    pack_args = (vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue)
    vol, owned = _as_volume(*pack_args)
    nx, ny = gx.shape[0], gy.shape[1]
    out = _new_output((nx, ny))
    try:
        if vol.ne == 0:
            out.fill(fillvalue if missing is None else missing)
        else:
            for _ in range(2):
                rc = _lib.check(_lib.load().pcg_get_cr_ex(
                    vol.handle, float(radar_height), radius, _lib.as_ptr(gx), _lib.as_ptr(gy), nx, ny,
                    float(fillvalue), float(fillvalue if missing is None else missing), _lib.as_ptr(out)))
                redo = _redo_in_f64(rc, vol, owned, pack_args)
                if redo is None:
                    break
                vol = redo
    finally:
        if owned:
            vol.close()
    return out


def _cappi(vol, radar_height, gx, gy, levels, fillvalue, radius, blind_code, beam_width_deg,
           return_index, missing=None):
    nx, ny = gx.shape[0], gx.shape[1]
    nz = int(levels.shape[0])
    shape = (nz, nx, ny) if vol.nfield == 1 else (vol.nfield, nz, nx, ny)
    out = _new_output(shape)
    idx = np.empty((nz, nx, ny, _lib.IDX_STRIDE), dtype=np.int32) if return_index else None
    rc = _lib.PCG_OK
    if out.size:
        rc = _lib.check(_lib.load().pcg_get_cappi3d_ex(
            vol.handle, float(radar_height), radius, _lib.as_ptr(gx), _lib.as_ptr(gy), nx, ny,
            levels.ctypes.data_as(_lib.c_double_p), nz, float(fillvalue), blind_code,
            float(beam_width_deg), float(fillvalue if missing is None else missing), _lib.as_ptr(out),
            _lib.as_ptr(idx) if idx is not None else None))
    return out, idx, rc


def get_CAPPI_xy(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
                 level_height, fillvalue, effective_earth_radius=None, blind_method="mask",
                 beam_width_deg=1.0, return_index=False, *, missing=None):
    """One CAPPI level (RadarGridC.pyx:356-477); ``missing``: see ``get_CR_xy``."""
    gx = _f64_2d(GridX, "GridX")
    gy = _f64_2d(GridY, "GridY")
    radius = _resolve_effective_earth_radius(effective_earth_radius)
    blind_code = _resolve_blind_code(blind_method)
    pack_args = (vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue)
    vol, owned = _as_volume(*pack_args)
    try:
        levels = np.array([float(level_height)], dtype=np.float64)
        for _ in range(2):
            out, idx, rc = _cappi(vol, radar_height, gx, gy, levels, fillvalue, radius, blind_code,
                                  float(beam_width_deg), return_index, missing)
            redo = _redo_in_f64(rc, vol, owned, pack_args)
            if redo is None:
                break
            vol = redo
    finally:
        if owned:
            vol.close()
    out = out[0] if vol.nfield == 1 else out[:, 0]
    return (out, idx[0]) if return_index else out


def get_CAPPI_3d(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
                 level_heights, fillvalue, effective_earth_radius=None, blind_method="mask",
                 return_index=False, *, missing=None):
    """3-D CAPPI volume, shape (nz, nx, ny) (RadarGridC.pyx:485-521; beam width is not forwarded
    by the reference, so the single-sweep gate always uses 1.0 degree here too); ``missing``: see ``get_CR_xy``."""
    gx = _f64_2d(GridX, "GridX")
    gy = _f64_2d(GridY, "GridY")
    levels = _f64_1d(level_heights, "level_heights")
    radius = _resolve_effective_earth_radius(effective_earth_radius)
    blind_code = _resolve_blind_code(blind_method)
    pack_args = (vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue)
    vol, owned = _as_volume(*pack_args)
    try:
        for _ in range(2):
            out, idx, rc = _cappi(vol, radar_height, gx, gy, levels, fillvalue, radius, blind_code, 1.0,
                                  return_index, missing)
            redo = _redo_in_f64(rc, vol, owned, pack_args)
            if redo is None:
                break
            vol = redo
    finally:
        if owned:
            vol.close()
    return (out, idx) if return_index else out


def get_mosaic_CAPPI_3d(vol_azimuth, vol_range, fix_elevation, vol_value, radar_x, radar_y,
                        radar_height, GridX, GridY, level_heights, fillvalue,
                        effective_earth_radius=None):
    """Multi-radar maximum-value 3-D CAPPI mosaic (RadarGridC.pyx:525-593)."""
    radar_x = _f64_1d(radar_x, "radar_x")
    radar_y = _f64_1d(radar_y, "radar_y")
    radar_height = _f64_1d(radar_height, "radar_height")
    gx = _f64_2d(GridX, "GridX")
    gy = _f64_2d(GridY, "GridY")
    levels = _f64_1d(level_heights, "level_heights")
    nradar = int(radar_height.shape[0])
    if not (len(vol_azimuth) == len(vol_range) == len(fix_elevation) == len(vol_value) == nradar):
        raise ValueError("Radar volume lists and radar position arrays must have the same length.")
    radar_radius = effective_earth_radius
    if radar_radius is None:
        radar_radius = [None] * nradar
    elif not isinstance(radar_radius, (list, tuple)):
        radar_radius = [radar_radius] * nradar
    elif len(radar_radius) != nradar:
        raise ValueError("effective_earth_radius must be None, a scalar, or a sequence matching the radar count.")
    radii = np.array([_resolve_effective_earth_radius(r) for r in radar_radius], dtype=np.float64)
    nx, ny, nz = gx.shape[0], gx.shape[1], int(levels.shape[0])
    out = _new_output((nz, nx, ny))
    vols, owned = [], []
    try:
        for i in range(nradar):
            v, o = _as_volume(vol_azimuth[i], vol_range[i], fix_elevation[i], vol_value[i], fillvalue)
            vols.append(v)
            owned.append(o)
        if nradar == 0:
            out.fill(fillvalue)
        elif out.size:
            handles = (ctypes.c_void_p * nradar)(*[v.handle for v in vols])
            _lib.check(_lib.load().pcg_get_mosaic_cappi3d(
                nradar, handles, radar_x.ctypes.data_as(_lib.c_double_p),
                radar_y.ctypes.data_as(_lib.c_double_p), radar_height.ctypes.data_as(_lib.c_double_p),
                radii.ctypes.data_as(_lib.c_double_p), _lib.as_ptr(gx), _lib.as_ptr(gy), nx, ny,
                levels.ctypes.data_as(_lib.c_double_p), nz, float(fillvalue), _lib.as_ptr(out)))
    finally:
        for v, o in zip(vols, owned):
            if o:
                v.close()
    return out
