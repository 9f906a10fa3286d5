"""Vertical sections and pseudo-RHI cuts on the GPU.

Host-side mirror of the section methods of the reference's ``PRD`` (``pycwr/core/NRadar.py``):

==============================  ==============================================================
``_build_section_line``         ``NRadar.py:796-810``
``_deduplicate_sorted_azimuth``  ``NRadar.py:813-821``
``_pad_section_rows``           ``NRadar.py:823-831``
``_interpolate_ppi_strip_arrays``  ``NRadar.py:833-913`` (one sweep, explicit targets)
``extract_section``             ``NRadar.py:1008-1096``
``extract_section_lonlat``      ``NRadar.py:1098-1121``
``extract_rhi``                 ``NRadar.py:1123-1200`` (the ppi branch of ``:1262-1274``)
==============================  ==============================================================

All sampling runs in ``section_kernel`` through ``pcg_section_sample`` (one launch per section: every
sweep × every line point); there is no CPU path — a missing library raises at import of ``_lib``.
Sweeps are packed once per ``(prd, field, range_mode)`` as a float64 NaN-marked device volume
(``SectionVolume``) and reused by later sections.
"""
import ctypes

import numpy as np

from . import _lib
from .interp import _new_dataset
from .runtime import DeviceVolume

DEFAULT_EFFECTIVE_EARTH_RADIUS = 6371.0 * 1000.0 * 4.0 / 3.0      # pycwr/core/transforms.py:48
_AEQD_R = 6370997.0


def _build_section_line(start_point, end_point, sample_spacing):
    start_point = np.asarray(start_point, dtype=np.float64)
    end_point = np.asarray(end_point, dtype=np.float64)
    if start_point.shape != (2,) or end_point.shape != (2,):
        raise ValueError("start_point and end_point must be 2-element coordinates.")
    delta = end_point - start_point
    line_length = float(np.hypot(delta[0], delta[1]))
    if line_length == 0.0:
        raise ValueError("start_point and end_point must not be identical.")
    npoints = max(2, int(np.ceil(line_length / float(sample_spacing))) + 1)
    distance = np.linspace(0.0, line_length, npoints, dtype=np.float64)
    fraction = distance / line_length
    return start_point[0] + fraction * delta[0], start_point[1] + fraction * delta[1], distance


def _deduplicate_sorted_azimuth(azimuth, values):
    if azimuth.size == 0:
        return azimuth, values
    unique_azimuth, unique_index = np.unique(azimuth, return_index=True)
    if unique_azimuth.size == azimuth.size:
        return azimuth, values
    return unique_azimuth, values[unique_index, :]


def _pad_section_rows(rows, fill_value=np.nan):
    rows = [np.asarray(r, dtype=np.float64).reshape(-1) for r in rows]
    out = np.full((len(rows), max(r.size for r in rows)), fill_value, dtype=np.float64)
    for i, r in enumerate(rows):
        out[i, : r.size] = r
    return out


def _section_table(azimuth, values):
    """The azimuth table the strip interpolation searches (``NRadar.py:845-850``): duplicates dropped, then
    reduced mod 360.  On a table that is no longer ascending after the reduction the reference's
    ``np.searchsorted`` result depends on the previous key (numpy's binary search carries its bounds from
    one key to the next), i.e. it is undefined; such a table is rotated back into ascending order here,
    which is what the wrap-around logic of ``:851-852`` assumes."""
    azimuth = np.asarray(azimuth, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    azimuth, values = _deduplicate_sorted_azimuth(azimuth, values)
    azimuth = np.mod(azimuth, 360.0)
    if azimuth.size > 1:
        drops = np.nonzero(np.diff(azimuth) < 0.0)[0]
        if drops.size == 1 and azimuth[-1] <= azimuth[0]:
            k = int(drops[0]) + 1
            azimuth = np.roll(azimuth, -k)
            values = np.roll(values, -k, axis=0)
    return np.ascontiguousarray(azimuth), np.ascontiguousarray(values)


class SectionVolume(object):
    """Sweeps of one moment resident on the device for section sampling: float64, NaN = missing, azimuth
    tables prepared by ``_section_table``.  ``sweeps`` = ``[(azimuth sorted, ranges, values), ...]`` in the
    order the rows of the section should have (the reference orders by fixed angle)."""

    def __init__(self, sweeps, fix_elevation, fillvalue=-999.0):
        self.fix_elevation = np.ascontiguousarray(fix_elevation, dtype=np.float64)
        if len(sweeps) != self.fix_elevation.size:
            raise ValueError("one fixed elevation per sweep is required")
        az, rg, val, self.usable = [], [], [], []
        for a, r, v in sweeps:
            a = np.asarray(a, dtype=np.float64).reshape(-1)
            r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1)
            v = np.asarray(v, dtype=np.float64)
            ok = v.ndim == 2 and a.size >= 2 and r.size >= 2          # NRadar.py:842-843
            if ok:
                a, v = _section_table(a, v)
                ok = a.size >= 2                                        # :846-847
            if not ok:                                                  # an all-NaN row: one ray, never usable
                a, r, v = np.zeros(1), np.zeros(1) if r.size == 0 else r[:1].copy(), np.full((1, 1), np.nan)
            self.usable.append(ok)
            az.append(a); rg.append(r); val.append(np.ascontiguousarray(v[:, : r.size]))
        self.ranges = rg
        self.volume = DeviceVolume(az, rg, self.fix_elevation, val, fillvalue=fillvalue, value_dtype="f64")
        self.ne = len(sweeps)

    def close(self):
        self.volume.close()

    def _call(self, npts, x, y, h, Re, taz, trg, per_sweep):
        lib = _lib.load()
        n = self.ne * npts
        out = np.empty((4 if x is not None else 1, self.ne, npts), dtype=np.float64)
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
        planes = [out[i] for i in range(out.shape[0])] + [None] * 3
        if n:
            _lib.check(lib.pcg_section_sample(self.volume.handle, 0, int(npts), p(x), p(y), float(h), float(Re),
                                              p(taz), p(trg), int(per_sweep), p(planes[0]), p(planes[1]),
                                              p(planes[2]), p(planes[3])))
        return out

    def sample_line(self, x_line, y_line, radar_height, effective_earth_radius=None):
        """(values, azimuth, range, z), each ``[ne][npts]``: the per-sweep loop of ``extract_section``."""
        x = np.ascontiguousarray(x_line, dtype=np.float64).reshape(-1)
        y = np.ascontiguousarray(y_line, dtype=np.float64).reshape(-1)
        if x.shape != y.shape:
            raise ValueError("x_line and y_line must have the same length")
        Re = DEFAULT_EFFECTIVE_EARTH_RADIUS if effective_earth_radius is None else float(effective_earth_radius)
        if Re <= 0:
            raise ValueError("effective_earth_radius must be positive")
        out = self._call(x.size, x, y, radar_height, Re, None, None, 0)
        return out[0], out[1], out[2], out[3]

    def sample_targets(self, target_azimuth, target_range):
        """Explicit targets: ``[npts]`` (shared by every sweep) or ``[ne][npts]``; returns ``[ne][npts]``."""
        taz = np.ascontiguousarray(target_azimuth, dtype=np.float64)
        trg = np.ascontiguousarray(target_range, dtype=np.float64)
        if taz.shape != trg.shape or taz.ndim not in (1, 2) or (taz.ndim == 2 and taz.shape[0] != self.ne):
            raise ValueError("targets must be [npts] or [ne][npts] and of equal shape")
        return self._call(taz.shape[-1], None, None, 0.0, 1.0, taz, trg, 1 if taz.ndim == 2 else 0)[0]

    def sample_rhi(self, azimuth):
        """Every sweep at its own gates along one azimuth (``_extract_ppi_rhi``), NaN-padded rows."""
        width = max(r.size for r in self.ranges)
        trg = _pad_section_rows(self.ranges)
        for i, ok in enumerate(self.usable):
            if not ok:
                trg[i, :] = np.nan
        taz = np.full((self.ne, width), float(azimuth), dtype=np.float64)
        return self.sample_targets(taz, trg)


def _interpolate_ppi_strip_arrays(azimuth, ranges, values, target_azimuth, target_range):
    """``PRD._interpolate_ppi_strip_arrays`` for one sweep (``NRadar.py:833-913``)."""
    target_range = np.asarray(target_range, dtype=np.float64)
    target_azimuth = np.broadcast_to(np.asarray(target_azimuth, dtype=np.float64), target_range.shape)
    values = np.asarray(values, dtype=np.float64)
    if target_range.size == 0 or values.ndim != 2:
        return np.full(target_range.shape, np.nan, dtype=np.float64)
    sv = SectionVolume([(azimuth, ranges, values)], [0.0])
    try:
        out = sv.sample_targets(target_azimuth.reshape(-1), target_range.reshape(-1))[0]
    finally:
        sv.close()
    return out.reshape(target_range.shape)


# ------------------------------------------------------------------------------------------------
# PRD-level entry points (plain-array PRD: pycwr_b200.products.ArrayPRD or anything with the same
# ``fields`` / ``fix_elevation`` / ``radar_height`` / ``lon`` / ``lat`` attributes)
# ------------------------------------------------------------------------------------------------

def _xy_to_lonlat(x, y, lon_0, lat_0, R=_AEQD_R):
    """Spherical aeqd inverse, the closed form of ``cartesian_to_geographic_aeqd``
    (``pycwr/core/transforms.py:825-897``); coordinates only, not on the sampling path."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    lat0, lon0 = np.deg2rad(lat_0), np.deg2rad(lon_0)
    rho = np.hypot(x, y)
    c = rho / R
    with np.errstate(invalid="ignore", divide="ignore"):
        arg = np.cos(c) * np.sin(lat0) + np.divide(y * np.sin(c) * np.cos(lat0), rho)
    lat = np.where(rho == 0, lat0, np.arcsin(np.clip(arg, -1.0, 1.0)))
    lon = lon0 + np.arctan2(x * np.sin(c), rho * np.cos(lat0) * np.cos(c) - y * np.sin(lat0) * np.sin(c))
    return ((np.rad2deg(lon) + 180.0) % 360.0) - 180.0, np.rad2deg(lat)


def _lonlat_to_xy(lon, lat, lon_0, lat_0):
    from .synth import aeqd_forward
    return aeqd_forward(lon, lat, lon_0, lat_0)


def _section_volume(prd, field_name, range_mode):
    """Device volume of ``field_name`` in fixed-angle order, cached on the PRD like ``_vol_cache``."""
    cache = prd.__dict__.setdefault("_section_cache", {})
    key = (field_name, range_mode)
    if key not in cache:
        if field_name not in prd.fields:
            raise KeyError(field_name)
        order = np.argsort(np.asarray(prd.fix_elevation, dtype=np.float64))       # _fixed_angle_sort_index
        sweeps = []
        for s in order:
            az, rg, val = prd.fields[field_name][int(s)]
            az = np.asarray(az, dtype=np.float64)
            rays = np.argsort(az)                                                  # _extract_sweep_volume :705-707
            sweeps.append((az[rays], np.asarray(rg, dtype=np.float64), np.asarray(val, dtype=np.float64)[rays, :]))
        cache[key] = (SectionVolume(sweeps, np.asarray(prd.fix_elevation, dtype=np.float64)[order]), order)
    return cache[key]


def _dataset(field_name, distance, rows, attrs):
    ds = _new_dataset({"sweep": np.arange(rows[field_name].shape[0], dtype=np.int32),
                       "distance": np.asarray(distance, dtype=np.float64)})
    for name, values in rows.items():
        dims = ("sweep",) if name == "source_sweep" else ("sweep", "distance")
        ds[name] = (dims, values)
    ds.attrs.update(attrs)
    return ds


def extract_section(prd, start, end, field_name="dBZ", point_units="km", interpolation="linear", range_mode=None,
                    sample_spacing=None):
    """``PRD.extract_section``: moment, beam height and antenna coordinates of every sweep along the
    straight line ``start`` -> ``end`` (radar-relative Cartesian, km or m)."""
    range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
    if interpolation != "linear":
        raise ValueError("Only linear interpolation is currently supported.")
    if point_units not in ("km", "m"):
        raise ValueError("point_units must be 'km' or 'm'.")
    scale = 1000.0 if point_units == "km" else 1.0
    start_xy = np.asarray(start, dtype=np.float64) * scale
    end_xy = np.asarray(end, dtype=np.float64) * scale
    sv, order = _section_volume(prd, field_name, range_mode)
    if sample_spacing is not None:                                                 # NRadar.py:779-793
        spacing = float(sample_spacing)
        if spacing <= 0.0:
            raise ValueError("sample_spacing must be positive")
    else:
        ref_range = np.asarray(prd.fields[field_name][0][1], dtype=np.float64)
        if ref_range.size < 2:
            raise ValueError("section extraction requires at least two range gates.")
        spacing = float(ref_range[1] - ref_range[0])
        if spacing <= 0.0:
            raise ValueError("range spacing must be positive to extract sections.")
    x_line, y_line, distance = _build_section_line(start_xy, end_xy, spacing)
    values, azimuth, ranges, z = sv.sample_line(x_line, y_line, prd.radar_height,
                                                getattr(prd, "effective_earth_radius", None))
    lon_line, lat_line = _xy_to_lonlat(x_line, y_line, prd.lon, prd.lat)
    ne = sv.ne
    tile = lambda a: np.tile(np.asarray(a, dtype=np.float64), (ne, 1))            # noqa: E731
    rows = {field_name: values, "x": tile(x_line), "y": tile(y_line), "z": z, "lon": tile(lon_line),
            "lat": tile(lat_line), "azimuth": azimuth, "range": ranges,
            "elevation": np.repeat(sv.fix_elevation[:, None], distance.size, axis=1),
            "distance_ground": tile(np.hypot(x_line, y_line)), "source_sweep": np.asarray(order, dtype=np.int32)}
    return _dataset(field_name, distance, rows, {
        "field_name": field_name, "section_type": "section", "scan_type": "ppi", "range_mode": range_mode,
        "interpolation": interpolation, "start_point": tuple(float(v) for v in start_xy),
        "end_point": tuple(float(v) for v in end_xy), "point_units": "meters"})


def extract_section_lonlat(prd, start_lonlat, end_lonlat, field_name="dBZ", interpolation="linear", range_mode=None,
                           sample_spacing=None):
    """``PRD.extract_section_lonlat``: the same section between two geographic end points."""
    sx, sy = _lonlat_to_xy(start_lonlat[0], start_lonlat[1], prd.lon, prd.lat)
    ex, ey = _lonlat_to_xy(end_lonlat[0], end_lonlat[1], prd.lon, prd.lat)
    ds = extract_section(prd, (float(np.asarray(sx).reshape(-1)[0]), float(np.asarray(sy).reshape(-1)[0])),
                         (float(np.asarray(ex).reshape(-1)[0]), float(np.asarray(ey).reshape(-1)[0])),
                         field_name=field_name, point_units="m", interpolation=interpolation,
                         range_mode=range_mode, sample_spacing=sample_spacing)
    ds.attrs["start_lonlat"] = tuple(float(v) for v in start_lonlat)
    ds.attrs["end_lonlat"] = tuple(float(v) for v in end_lonlat)
    return ds


def extract_rhi(prd, azimuth=None, field_name="dBZ", range_mode=None):
    """The ppi branch of ``PRD.extract_rhi``: an RHI-style cut along one azimuth."""
    if azimuth is None:
        raise ValueError("azimuth is required when extracting an RHI from a ppi volume.")
    range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
    sv, order = _section_volume(prd, field_name, range_mode)
    target = float(azimuth)
    values = sv.sample_rhi(target)
    Re = getattr(prd, "effective_earth_radius", None)
    Re = DEFAULT_EFFECTIVE_EARTH_RADIUS if Re is None else float(Re)
    xs, ys, zs, lons, lats, rgs, els, azs, dist = [], [], [], [], [], [], [], [], []
    for i in range(sv.ne):                    # antenna_to_cartesian_cwr (transforms.py:187-199): coordinates only
        r = np.asarray(prd.fields[field_name][int(order[i])][1], dtype=np.float64)
        e, a = np.deg2rad(sv.fix_elevation[i]), np.deg2rad(target)
        radial = r * np.cos(e)
        gate_radius = np.sqrt(radial ** 2 + (Re + prd.radar_height + r * np.sin(e)) ** 2)
        s = Re * np.arcsin(np.clip(radial / gate_radius, -1.0, 1.0))
        x, y = s * np.sin(a), s * np.cos(a)
        lon, lat = _xy_to_lonlat(x, y, prd.lon, prd.lat)
        xs.append(x); ys.append(y); zs.append(gate_radius - Re); lons.append(lon); lats.append(lat)
        rgs.append(r); els.append(np.full(r.shape, sv.fix_elevation[i])); azs.append(np.full(r.shape, target))
        dist.append(np.hypot(x, y))
    width = values.shape[1]
    rows = {field_name: values, "x": _pad_section_rows(xs), "y": _pad_section_rows(ys), "z": _pad_section_rows(zs),
            "lon": _pad_section_rows(lons), "lat": _pad_section_rows(lats), "azimuth": _pad_section_rows(azs),
            "range": _pad_section_rows(rgs), "elevation": _pad_section_rows(els),
            "distance_ground": _pad_section_rows(dist), "source_sweep": np.asarray(order, dtype=np.int32)}
    distance_axis = np.asarray(max(dist, key=lambda row: row.size), dtype=np.float64)
    assert distance_axis.size == width
    return _dataset(field_name, distance_axis, rows, {
        "field_name": field_name, "section_type": "rhi", "scan_type": "ppi", "range_mode": range_mode,
        "interpolation": "linear", "target_azimuth": target})
