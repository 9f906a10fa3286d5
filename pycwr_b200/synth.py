"""Deterministic synthetic radar volumes and target grids for the five BASELINE configs.

Shapes and dtypes follow what the reference's readers emit and what
``PRD.get_vol_data`` hands to the gridding kernels (reference
``pycwr/core/NRadar.py:1931-1961``): per sweep a sorted float64 azimuth table,
a float64 range table ``(i+1)*dr`` (``pycwr/io/WSR98DFile.py:596-604``) and a
float64 ``(naz, ngate)`` moment array with missing gates set to the fill value.
Seeds follow the style of the reference's ``test/test_regression_geometry.py:14``
(20260319 + config id).
"""
import numpy as np

FILL = -999.0
BASE_SEED = 20260319
VCP21 = np.array([0.5, 1.45, 2.4, 3.35, 4.3, 6.0, 9.9, 14.6, 19.5], dtype=np.float64)

FIELD_RANGES = {"dBZ": (-5.0, 70.0), "ZDR": (-2.0, 5.0), "V": (-27.0, 27.0)}


def azimuth_table(rng, naz):
    """Jittered, sorted azimuth table in [0, 360) (one per sweep)."""
    step = 360.0 / naz
    a0 = rng.uniform(0.0, step)
    az = a0 + np.arange(naz) * step + rng.uniform(-0.1, 0.1, naz) * step
    az = np.mod(az, 360.0)
    return np.sort(az).astype(np.float64)


def stress_moment(rng, naz, ng, lo, hi, fill=FILL):
    """i.i.d. values with 30 % fill gates, 5 % fill rays and 5 % fill range blocks."""
    val = rng.uniform(lo, hi, (naz, ng))
    val[rng.random((naz, ng)) < 0.30] = fill
    val[rng.random(naz) < 0.05, :] = fill
    nblk = max(1, ng // 64)
    blk = rng.random(nblk) < 0.05
    gate_blk = np.minimum(np.arange(ng) // 64, nblk - 1)
    val[:, blk[gate_blk]] = fill
    # the reference's readers emit float32 moments (pycwr/io/PAFile.py:434-444) which
    # PRD.get_vol_data widens to float64: keep the synthetic moments float32-representable too
    return np.ascontiguousarray(val.astype(np.float32), dtype=np.float64)


def storm_moment(rng, az, rg, elev, lo, hi, cells, fill=FILL):
    """Sum of 3-D Gaussian cells sampled at gate centres, 0.5-unit quantised, fill below lo+5."""
    a = np.deg2rad(az)[:, None]
    e = np.deg2rad(elev)
    s = rg[None, :] * np.cos(e)
    x = s * np.sin(a)
    y = s * np.cos(a)
    z = rg[None, :] * np.sin(e) + (rg[None, :] ** 2) / (2.0 * 8494666.6666666661)
    acc = np.zeros((az.size, rg.size))
    for cx, cy, cz, sx, sz, amp in cells:
        acc += amp * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sx * sx) - ((z - cz) ** 2) / (2 * sz * sz))
    val = lo + (hi - lo) * np.clip(acc, 0.0, 1.0)
    val = np.round(val * 2.0) / 2.0
    val[acc < 0.07] = fill
    return np.ascontiguousarray(val.astype(np.float32), dtype=np.float64)


def make_volume(seed, elevations=VCP21, naz=366, ngate=1840, gate_m=250.0, fields=("dBZ",),
                style="stress", radar_height=None, max_extent_m=None):
    """Return a dict with the reference's 7-tuple layout per field plus metadata."""
    rng = np.random.default_rng(seed)
    elevations = np.asarray(elevations, dtype=np.float64)
    ne = elevations.size
    vol_azimuth = [azimuth_table(rng, naz) for _ in range(ne)]
    ranges = (np.arange(ngate, dtype=np.float64) + 1.0) * float(gate_m)
    vol_range = [ranges.copy() for _ in range(ne)]
    h = float(rng.uniform(50.0, 2000.0)) if radar_height is None else float(radar_height)
    extent = ranges[-1] if max_extent_m is None else max_extent_m
    cells = [
        (rng.uniform(-0.8, 0.8) * extent, rng.uniform(-0.8, 0.8) * extent, rng.uniform(1e3, 9e3),
         rng.uniform(0.03, 0.12) * extent, rng.uniform(1.5e3, 5e3), rng.uniform(0.4, 1.0))
        for _ in range(20)
    ]
    values = {}
    for name in fields:
        lo, hi = FIELD_RANGES.get(name, (-5.0, 70.0))
        per_sweep = []
        for k in range(ne):
            if style == "stress":
                per_sweep.append(stress_moment(rng, naz, ngate, lo, hi))
            elif style == "storm":
                per_sweep.append(storm_moment(rng, vol_azimuth[k], ranges, elevations[k], lo, hi, cells))
            else:
                raise ValueError("style must be 'stress' or 'storm'")
        values[name] = per_sweep
    return {
        "vol_azimuth": vol_azimuth,
        "vol_range": vol_range,
        "fix_elevation": elevations.copy(),
        "vol_value": values,
        "radar_height": h,
        "beam_widths": np.full(ne, 360.0 / naz),
    }


def vol_tuple(volume, field="dBZ"):
    """(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height) for one field."""
    return (volume["vol_azimuth"], volume["vol_range"], volume["fix_elevation"],
            volume["vol_value"][field], volume["radar_height"])


def xy_grid(half_extent_m, step_m):
    """``np.meshgrid(XRange, YRange, indexing="ij")`` as the product methods build it
    (reference ``pycwr/core/NRadar.py:1329``)."""
    n = int(round(2 * half_extent_m / step_m)) + 1
    axis = np.linspace(-half_extent_m, half_extent_m, n)
    gx, gy = np.meshgrid(axis, axis, indexing="ij")
    return axis, axis.copy(), np.ascontiguousarray(gx), np.ascontiguousarray(gy)


_AEQD_R = 6370997.0


def aeqd_forward(lon, lat, lon_0, lat_0, R=_AEQD_R):
    """Spherical azimuthal-equidistant forward projection (degrees -> metres).

    The reference's lon/lat products project with pyproj's WGS84 ``aeqd``
    (``pycwr/core/NRadar.py:553-558``); pyproj is not installed in this image, so the
    synthetic lon/lat configs use the spherical formula (same closed form as the reference's
    ``geographic_to_cartesian_aeqd``, ``pycwr/core/transforms.py:660-732``).  The projection is
    outside the kernel boundary: kernels and oracle receive the same projected metres.
    """
    lon = np.deg2rad(np.asarray(lon, dtype=np.float64))
    lat = np.deg2rad(np.asarray(lat, dtype=np.float64))
    lon0 = np.deg2rad(lon_0)
    lat0 = np.deg2rad(lat_0)
    cos_c = np.sin(lat0) * np.sin(lat) + np.cos(lat0) * np.cos(lat) * np.cos(lon - lon0)
    c = np.arccos(np.clip(cos_c, -1.0, 1.0))
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(c == 0.0, 1.0, c / np.sin(c))
    x = R * k * np.cos(lat) * np.sin(lon - lon0)
    y = R * k * (np.cos(lat0) * np.sin(lat) - np.sin(lat0) * np.cos(lat) * np.cos(lon - lon0))
    return x, y


def lonlat_grid(lon_0, lat_0, half_extent_m, n, indexing="ij"):
    """Regular lon/lat axes spanning about +-half_extent_m around the site, projected to metres."""
    dlat = np.rad2deg(half_extent_m / _AEQD_R)
    dlon = dlat / np.cos(np.deg2rad(lat_0))
    lon = np.linspace(lon_0 - dlon, lon_0 + dlon, n)
    lat = np.linspace(lat_0 - dlat, lat_0 + dlat, n)
    glon, glat = np.meshgrid(lon, lat, indexing=indexing)
    gx, gy = aeqd_forward(glon, glat, lon_0, lat_0)
    return lon, lat, np.ascontiguousarray(gx, dtype=np.float64), np.ascontiguousarray(gy, dtype=np.float64)


# --------------------------------------------------------------------------------------
# The five BASELINE.json configs.  ``scale`` < 1 shrinks the *grid* (not the volume) so the
# CPU oracle finishes in seconds; scale == 1 is the full benchmark size.
# --------------------------------------------------------------------------------------

def config(cid, style="stress", scale=1.0, nradar=None):
    seed = BASE_SEED + cid
    if cid == 1:  # CR, 9x366x1840, 301x301 @ 1 km
        vol = make_volume(seed, fields=("dBZ",), style=style)
        n = max(3, int(round(300 * scale)) + 1)
        ax = np.linspace(-150e3, 150e3, n)
        gx, gy = np.meshgrid(ax, ax, indexing="ij")
        return {"id": 1, "name": "C1_CR_301x301", "kind": "cr", "volume": vol, "fields": ["dBZ"],
                "GridX": np.ascontiguousarray(gx), "GridY": np.ascontiguousarray(gy),
                "levels": None, "blind_method": "mask"}
    if cid == 2:  # CAPPI 601x601x20, dBZ + ZDR + V
        vol = make_volume(seed, fields=("dBZ", "ZDR", "V"), style=style)
        n = max(3, int(round(600 * scale)) + 1)
        ax = np.linspace(-150e3, 150e3, n)
        gx, gy = np.meshgrid(ax, ax, indexing="ij")
        levels = np.arange(1, 21, dtype=np.float64) * 500.0
        return {"id": 2, "name": "C2_CAPPI_601x601x20_3f", "kind": "cappi3d", "volume": vol,
                "fields": ["dBZ", "ZDR", "V"], "GridX": np.ascontiguousarray(gx),
                "GridY": np.ascontiguousarray(gy), "levels": levels, "blind_method": "mask"}
    if cid == 3:  # full 3-D volume, lon/lat variant, 1201x1201x40 @ 250 m
        vol = make_volume(seed, naz=360, fields=("dBZ",), style=style)
        n = max(3, int(round(1200 * scale)) + 1)
        _, _, gx, gy = lonlat_grid(116.47, 39.81, 150e3, n, indexing="ij")
        levels = np.arange(1, 41, dtype=np.float64) * 500.0
        return {"id": 3, "name": "C3_3D_lonlat_1201x1201x40", "kind": "cappi3d", "volume": vol,
                "fields": ["dBZ"], "GridX": gx, "GridY": gy, "levels": levels, "blind_method": "mask"}
    if cid == 5:  # X-band phased array 64x720x2000 @ 30 m -> 1001x1001x60 @ 100 m
        vol = make_volume(seed, elevations=np.linspace(0.9, 57.6, 64), naz=720, ngate=2000,
                          gate_m=30.0, fields=("dBZ",), style=style)
        n = max(3, int(round(1000 * scale)) + 1)
        ax = np.linspace(-50e3, 50e3, n)
        gx, gy = np.meshgrid(ax, ax, indexing="ij")
        levels = np.arange(1, 61, dtype=np.float64) * 250.0
        return {"id": 5, "name": "C5_PA_1001x1001x60", "kind": "cappi3d", "volume": vol,
                "fields": ["dBZ"], "GridX": np.ascontiguousarray(gx), "GridY": np.ascontiguousarray(gy),
                "levels": levels, "blind_method": "mask"}
    if cid == 4:  # 32-radar network on a lon/lat grid
        return network_config(style=style, scale=scale, nradar=32 if nradar is None else nradar)
    raise ValueError("config id must be 1..5")


def network_config(style="storm", scale=1.0, nradar=32, nlon=3000, nlat=3000, nz=30,
                   share_volume=False, with_volumes=True):
    """C4: ``nradar`` S-band sites on a jittered 4x8 lattice in 100-130E / 20-50N, 0.01 deg grid.

    ``scale`` shrinks the number of grid points (coarser resolution over the same domain).
    ``share_volume`` reuses one synthetic volume for every site (cuts generation time for
    throughput runs; parity runs use distinct volumes).
    """
    seed = BASE_SEED + 4
    rng = np.random.default_rng(seed)
    nlon = max(4, int(round(nlon * scale)))
    nlat = max(4, int(round(nlat * scale)))
    lon = 100.0 + np.arange(nlon) * (30.0 / nlon)
    lat = 20.0 + np.arange(nlat) * (30.0 / nlat)
    rows, cols = 4, 8
    sites = []
    for k in range(nradar):
        i, j = divmod(k % (rows * cols), cols)
        slon = 100.0 + (j + 0.5) * 30.0 / cols + rng.uniform(-0.8, 0.8)
        slat = 20.0 + (i + 0.5) * 30.0 / rows + rng.uniform(-1.5, 1.5)
        sites.append((slon, slat, float(rng.uniform(50.0, 2000.0))))
    lon_0 = float(np.mean([s[0] for s in sites]))
    lat_0 = float(np.mean([s[1] for s in sites]))
    glon, glat = np.meshgrid(lon, lat, indexing="xy")  # (nlat, nlon) — RadarInterp.py:392
    gx, gy = aeqd_forward(glon, glat, lon_0, lat_0)
    volumes = []
    base = None
    for k, (slon, slat, alt) in enumerate(sites if with_volumes else []):
        if share_volume and base is not None:
            vol = dict(base)
            vol["radar_height"] = alt
        else:
            vol = make_volume(seed + 100 + k, fields=("dBZ",), style=style, radar_height=alt,
                              max_extent_m=460e3)
            base = vol
        volumes.append(vol)
    sx, sy = aeqd_forward(np.array([s[0] for s in sites]), np.array([s[1] for s in sites]), lon_0, lat_0)
    levels = np.arange(1, nz + 1, dtype=np.float64) * 500.0
    return {"id": 4, "name": "C4_network_%dr_%dx%dx%d" % (nradar, nlat, nlon, nz), "kind": "network",
            "volumes": volumes, "fields": ["dBZ"], "lon": lon, "lat": lat,
            "GridX": np.ascontiguousarray(gx), "GridY": np.ascontiguousarray(gy),
            "radar_x": np.asarray(sx, dtype=np.float64), "radar_y": np.asarray(sy, dtype=np.float64),
            "radar_alt": np.array([s[2] for s in sites], dtype=np.float64),
            "levels": levels, "blind_method": "hybrid", "composite_method": "quality_weighted",
            "max_range_km": 460.0, "influence_radius_m": 300000.0,
            "quality_weight_terms": ["range", "beam", "vertical", "qc", "blockage"]}
