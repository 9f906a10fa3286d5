"""CPU oracle for vertical sections / pseudo-RHI cuts.

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and the golden generator); the product path never
touches it.  Restates, point by point with scalar IEEE-double arithmetic,

* ``cartesian_xy_elevation_to_range_z`` (``pycwr/core/transforms.py:224-250``): ``section_geometry``;
* ``PRD._interpolate_ppi_strip_arrays`` (``pycwr/core/NRadar.py:833-913``) with its helper
  ``_deduplicate_sorted_azimuth`` (``:813-821``): ``strip_interp``;
* ``PRD._build_section_line`` (``:796-810``): ``section_line``;
* the per-sweep loops of ``PRD.extract_section`` (``:1048-1065``) and ``PRD._extract_ppi_rhi``
  (``:1144-1155``): ``section_rows`` / ``rhi_rows``.

PINNED: ``tests/golden/section_reference.npz`` holds outputs of the reference's own functions
(``tests/golden/make_section_golden.py``); ``tests/test_section.py`` checks this file against them
bit for bit (interpolation) and to 1 ulp-level tolerances (geometry).
"""
import bisect
import math

import numpy as np

DEFAULT_RE = 6371.0 * 1000.0 * 4.0 / 3.0          # transforms.py:48


def section_line(start_point, end_point, sample_spacing):
    """NRadar.py:796-810: sample points every ``sample_spacing`` metres, both ends included."""
    p0 = np.asarray(start_point, dtype=np.float64)
    p1 = np.asarray(end_point, dtype=np.float64)
    if p0.shape != (2,) or p1.shape != (2,):
        raise ValueError("start_point and end_point must be 2-element coordinates.")
    d = p1 - p0
    length = float(np.hypot(d[0], d[1]))
    if length == 0.0:
        raise ValueError("start_point and end_point must not be identical.")
    n = max(2, int(np.ceil(length / float(sample_spacing))) + 1)
    distance = np.linspace(0.0, length, n, dtype=np.float64)
    frac = distance / length
    return p0[0] + frac * d[0], p0[1] + frac * d[1], distance


def section_geometry(x, y, elevation_deg, h, effective_earth_radius=None):
    """transforms.py:224-250: (azimuth deg, slant range m, beam height m) of ground points (x, y) seen on
    the cone of elevation ``elevation_deg`` from a radar at altitude ``h``."""
    Re = DEFAULT_RE if effective_earth_radius is None else float(effective_earth_radius)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ground = np.hypot(x, y)
    phi = ground / Re
    tan_e = np.tan(np.deg2rad(elevation_deg))
    cphi = np.cos(phi)
    a = Re + h
    with np.errstate(invalid="ignore", divide="ignore"):
        z = a / (cphi - np.sin(phi) * tan_e) - Re
    g = Re + z
    r2 = a ** 2 + g * g - 2 * a * g * cphi
    rng = np.sqrt(np.maximum(r2, 0.0))
    az = np.rad2deg(np.mod(np.pi / 2 - np.arctan2(y, x), 2 * np.pi))
    return az, rng, z


def dedupe_sorted_azimuth(azimuth, values):
    """NRadar.py:813-821: first ray of every run of equal azimuths."""
    azimuth = np.asarray(azimuth, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    if azimuth.size == 0:
        return azimuth, values
    uniq, first = np.unique(azimuth, return_index=True)
    if uniq.size == azimuth.size:
        return azimuth, values
    return uniq, values[first, :]


def _pymod(a, b):
    return float(np.mod(np.float64(a), np.float64(b)))


def strip_interp(azimuth, ranges, values, target_azimuth, target_range):
    """NRadar.py:833-913, one target at a time."""
    azimuth = np.asarray(azimuth, dtype=np.float64)
    ranges = np.asarray(ranges, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    taz = np.asarray(target_azimuth, dtype=np.float64)
    trg = np.asarray(target_range, dtype=np.float64)
    out = np.full(trg.shape, np.nan, dtype=np.float64)
    if values.ndim != 2 or azimuth.size < 2 or ranges.size < 2:
        return out
    azimuth, values = dedupe_sorted_azimuth(azimuth, values)
    if azimuth.size < 2:
        return out
    az = [_pymod(v, 360.0) for v in azimuth]
    n = len(az)
    ext = [az[-1] - 360.0] + az + [az[0] + 360.0]
    rows = [n - 1] + list(range(n)) + [0]
    rg = [float(v) for v in ranges]
    ng = len(rg)
    flat_t = taz.reshape(-1)
    flat_r = trg.reshape(-1)
    flat_o = out.reshape(-1)
    fin = math.isfinite
    for i in range(flat_r.size):
        t = _pymod(flat_t[i], 360.0)
        r = float(flat_r[i])
        if not (fin(t) and fin(r)) or r < rg[0] or r > rg[-1]:
            continue
        ah = min(max(bisect.bisect_right(ext, t), 1), n + 1)
        gh = min(max(bisect.bisect_right(rg, r), 1), ng - 1)
        al, gl = ah - 1, gh - 1
        a0, a1, r0, r1 = ext[al], ext[ah], rg[gl], rg[gh]
        if not (a1 > a0 and r1 > r0):
            continue
        m00 = float(values[rows[al], gl]); m01 = float(values[rows[al], gh])
        m10 = float(values[rows[ah], gl]); m11 = float(values[rows[ah], gh])
        if fin(m00) and fin(m01) and fin(m10) and fin(m11):
            v = (m00 * (a1 - t) * (r1 - r) + m10 * (t - a0) * (r1 - r) + m01 * (a1 - t) * (r - r0)
                 + m11 * (t - a0) * (r - r0)) / ((a1 - a0) * (r1 - r0))
        elif fin(m00) and fin(m01):
            v = (m00 * (r1 - r) + m01 * (r - r0)) / (r1 - r0)
        elif fin(m10) and fin(m11):
            v = (m10 * (r1 - r) + m11 * (r - r0)) / (r1 - r0)
        elif fin(m00) and fin(m10):
            v = (m00 * (a1 - t) + m10 * (t - a0)) / (a1 - a0)
        elif fin(m01) and fin(m11):
            v = (m01 * (a1 - t) + m11 * (t - a0)) / (a1 - a0)
        else:
            continue
        flat_o[i] = v
    return out


def section_rows(sweeps, fix_elevation, x_line, y_line, h, effective_earth_radius=None):
    """Per-sweep loop of extract_section (NRadar.py:1048-1065) over sweeps already ordered by elevation:
    returns (values, azimuth, range, z), each [ne][npts]."""
    vals, azs, rgs, zs = [], [], [], []
    for (az, rg, v), e in zip(sweeps, fix_elevation):
        taz, trg, z = section_geometry(x_line, y_line, float(e), h, effective_earth_radius)
        vals.append(strip_interp(az, rg, v, taz, trg))
        azs.append(taz); rgs.append(trg); zs.append(z)
    return np.stack(vals), np.stack(azs), np.stack(rgs), np.stack(zs)


def rhi_rows(sweeps, azimuth):
    """Per-sweep loop of _extract_ppi_rhi (NRadar.py:1144-1155): every sweep sampled at its own gates along
    one azimuth; rows padded with NaN to the longest sweep (``_pad_section_rows`` :823-831)."""
    rows = [strip_interp(az, rg, v, np.full(np.shape(rg), float(azimuth)), rg) for az, rg, v in sweeps]
    width = max(r.size for r in rows)
    out = np.full((len(rows), width), np.nan, dtype=np.float64)
    for i, r in enumerate(rows):
        out[i, : r.size] = r
    return out
