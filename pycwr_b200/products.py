"""Host product wrappers over the B200 kernels — the reference's ``PRD.add_product_*`` methods
(``pycwr/core/NRadar.py:1276-1457,1746-1818``), ``grid_3d_network_xy`` (``:2042-2122``), the volume
packer ``get_vol_data`` (``:1931-1961``) and the column products ``derive_cr / derive_vil /
derive_et`` (``pycwr/core/RadarProduct.py:26-85``).

The wrappers are plain functions taking the PRD as first argument plus :class:`B200ProductMixin`,
which a ``PRD`` subclass can mix in to route its product methods through the GPU with unchanged
signatures and ``self.product`` layout (coordinate / variable names, dims, fill -> NaN).  They work
on duck-typed PRD objects (``get_vol_data``, ``.vol``, ``.product``, ``.effective_earth_radius``);
``xarray`` is only needed if the PRD's ``.product`` is an ``xarray.Dataset``.
"""
import numpy as np

from . import RadarGridC as G
from . import _lib
from .interp import SimpleDataset, _new_dataset, _project

FILL = -999.0


# ------------------------------------------------------------------------------------------------
# packer (A10): PRD.get_vol_data / _extract_sweep_volume on plain arrays
# ------------------------------------------------------------------------------------------------

def pack_volume(sweeps, fix_elevation, radar_height, radar_lon=0.0, radar_lat=0.0, fillvalue=FILL,
                max_range_km=None):
    """Build the reference's ``PRD.vol`` 7-tuple from per-sweep ``(azimuth, ranges, values)`` arrays in
    scan order (``pycwr/core/NRadar.py:665-712,1931-1961``): sweeps ordered by fixed angle, rays by
    azimuth, float64, optional ``max_range_km`` clip (``:655-663``: at least one gate is kept), NaN ->
    ``fillvalue``."""
    fix = np.asarray(fix_elevation, dtype=np.float64)
    order = np.argsort(fix)
    vol_azimuth, vol_range, vol_value = [], [], []
    for s in order:
        az, rng, val = sweeps[int(s)]
        az = np.asarray(az, dtype=np.float64)
        rng = np.asarray(rng, dtype=np.float64)
        val = np.asarray(val, dtype=np.float64)
        rays = np.argsort(az)
        az, val = az[rays], val[rays]
        if max_range_km is not None:
            keep = rng <= float(max_range_km) * 1000.0
            if not keep.any():
                rng, val = rng[:1], val[:, :1]
            else:
                rng, val = rng[keep], val[:, keep]
        vol_azimuth.append(np.ascontiguousarray(az))
        vol_range.append(np.ascontiguousarray(rng))
        vol_value.append(np.ascontiguousarray(np.where(np.isnan(val), float(fillvalue), val)))
    return vol_azimuth, vol_range, fix[order], vol_value, float(radar_height), float(radar_lon), float(radar_lat)


class ArrayPRD(object):
    """Minimal PRD stand-in over plain arrays (the role of the reference tests' ``FakePRD``,
    ``test/test_examples_interp.py:86-96``): one set of sweeps per field, ``get_vol_data`` fills
    ``.vol`` like ``NRadar.PRD`` and caches per ``(field, fill, mode, max_range)`` (``:1938-1941``)."""

    def __init__(self, fields, fix_elevation, radar_height, lon=0.0, lat=0.0, beam_width=None,
                 effective_earth_radius=None, native_fields=None):
        self.fields = fields                     # {name: [(azimuth, ranges, values), ...] in scan order}
        # native-range copies of a moment (the reference's ``extended_fields``, NRadar.py:669-680): used by
        # ``range_mode="native"`` when present, else the aligned sweeps serve both modes
        self.native_fields = dict(native_fields or {})
        self.fix_elevation = np.asarray(fix_elevation, dtype=np.float64)
        self.radar_height = float(radar_height)
        self.lon, self.lat = float(lon), float(lat)
        self.effective_earth_radius = effective_earth_radius
        self.scan_info = {}
        if beam_width is not None:
            self.scan_info["beam_width"] = _Values(np.asarray(beam_width, dtype=np.float64))
        self.scan_info["longitude"] = _Values(np.float64(lon))
        self.scan_info["latitude"] = _Values(np.float64(lat))
        self.product = SimpleDataset({})
        self._vol_cache = {}
        self._dev_cache = {}
        self.vol = None

    @staticmethod
    def _product_name(base_name, range_mode):    # NRadar.py:650-652
        return base_name if range_mode == "aligned" else "%s_native" % base_name

    def _resolve_field_range_mode(self, field_name, range_mode=None):   # NRadar.py:233-235,256-261
        if range_mode is not None:
            if range_mode not in ("aligned", "native"):
                raise ValueError("range_mode must be 'aligned' or 'native'.")
            return range_mode
        return "native" if str(field_name) in ("dBZ", "Zc") else "aligned"

    def _sweeps_of(self, field_name, range_mode):
        if field_name not in self.fields:
            raise KeyError(field_name)
        if range_mode == "native" and field_name in self.native_fields:
            return self.native_fields[field_name]
        return self.fields[field_name]

    def get_vol_data(self, field_name="dBZ", fillvalue=FILL, range_mode=None, max_range_km=None):
        range_mode = self._resolve_field_range_mode(field_name, range_mode=range_mode)
        max_range_km = None if max_range_km is None else float(max_range_km)
        sweeps = self._sweeps_of(field_name, range_mode)
        key = (field_name, float(fillvalue), range_mode, max_range_km)
        if key not in self._vol_cache:
            self._vol_cache[key] = pack_volume(sweeps, self.fix_elevation, self.radar_height,
                                               self.lon, self.lat, fillvalue, max_range_km)
        self.vol = self._vol_cache[key]
        return self.vol

    def get_device_volume(self, field_name="dBZ", fillvalue=FILL, range_mode=None, max_range_km=None,
                          value_dtype="f32", device=None):
        """The device form of ``get_vol_data``: the sweeps go to the GPU as stored (scan order,
        NaN) and are sorted / clipped / fill-marked by the pack kernels; cached per key like ``_vol_cache``.
        ``device``: the GPU the calling thread has selected (``_lib.set_device``), part of the cache key — the
        multi-GPU network product asks for the same field on several devices."""
        from . import _lib
        range_mode = self._resolve_field_range_mode(field_name, range_mode=range_mode)
        max_range_km = None if max_range_km is None else float(max_range_km)
        sweeps = self._sweeps_of(field_name, range_mode)
        device = _lib.current_device() if device is None else int(device)
        key = (field_name, float(fillvalue), range_mode, max_range_km, value_dtype, device)
        if key not in self._dev_cache:
            from .runtime import DeviceVolume
            self._dev_cache[key] = DeviceVolume.from_sweeps(sweeps, self.fix_elevation,
                                                            fillvalue=fillvalue, value_dtype=value_dtype,
                                                            max_range_km=max_range_km)
        return self._dev_cache[key]

    # vertical sections: NRadar.py:1008-1121,1123-1200 on the GPU
    def extract_section(self, start, end, field_name="dBZ", point_units="km", interpolation="linear",
                        range_mode=None, sample_spacing=None):
        from . import section
        return section.extract_section(self, start, end, field_name, point_units, interpolation, range_mode,
                                       sample_spacing)

    def extract_section_lonlat(self, start_lonlat, end_lonlat, field_name="dBZ", interpolation="linear",
                               range_mode=None, sample_spacing=None):
        from . import section
        return section.extract_section_lonlat(self, start_lonlat, end_lonlat, field_name, interpolation,
                                              range_mode, sample_spacing)

    def extract_rhi(self, azimuth=None, field_name="dBZ", range_mode=None):
        from . import section
        return section.extract_rhi(self, azimuth, field_name, range_mode)


class _Values(object):
    def __init__(self, values):
        self.values = values


# ------------------------------------------------------------------------------------------------
# column products (f1)
# ------------------------------------------------------------------------------------------------

def column_products(volume, level_heights, fillvalue=FILL, min_dbz=18.0, max_dbz_cap=56.0, threshold_dbz=18.0):
    """CR, VIL, ET and the topped flag of a gridded volume in one pass over its columns."""
    vol = np.ascontiguousarray(volume, dtype=np.float64)
    if vol.ndim != 3:
        raise ValueError("volume must have shape (nz, nx, ny).")
    levels = np.ascontiguousarray(np.asarray(level_heights, dtype=np.float64).ravel())
    nz, nx, ny = vol.shape
    out = {"CR": np.full((nx, ny), np.nan), "VIL": np.full((nx, ny), np.nan), "ET": np.full((nx, ny), np.nan),
           "ET_TOPPED": np.zeros((nx, ny), dtype=np.uint8)}
    if nx * ny == 0:
        return out
    if levels.size != nz:
        raise ValueError("level_heights must have one entry per level of the volume.")
    _lib.check(_lib.load().pcg_column_products(
        _lib.as_ptr(vol), levels.ctypes.data_as(_lib.c_double_p), nz, nx, ny,
        float(fillvalue), float(min_dbz), float(max_dbz_cap), float(threshold_dbz), _lib.as_ptr(out["CR"]),
        _lib.as_ptr(out["VIL"]), _lib.as_ptr(out["ET"]), _lib.as_ptr(out["ET_TOPPED"])))
    return out


def derive_cr(volume, fillvalue=FILL):
    """RadarProduct.py:26-30."""
    return column_products(volume, np.zeros(np.shape(volume)[0]), fillvalue)["CR"]


def derive_vil(volume, level_heights, fillvalue=FILL, min_dbz=18.0, max_dbz_cap=56.0):
    """RadarProduct.py:33-46."""
    level_heights = np.asarray(level_heights, dtype=np.float64)
    if level_heights.size < 2:
        return np.full(np.shape(volume)[1:], np.nan, dtype=np.float64)
    return column_products(volume, level_heights, fillvalue, min_dbz, max_dbz_cap)["VIL"]


def derive_et(volume, level_heights, fillvalue=FILL, threshold_dbz=18.0, return_topped=False):
    """RadarProduct.py:49-85."""
    level_heights = np.asarray(level_heights, dtype=np.float64)
    if level_heights.size == 0:
        et = np.full(np.shape(volume)[1:], np.nan, dtype=np.float64)
        topped = np.zeros(np.shape(volume)[1:], dtype=np.uint8)
        return (et, topped) if return_topped else et
    res = column_products(volume, level_heights, fillvalue, threshold_dbz=threshold_dbz)
    return (res["ET"], res["ET_TOPPED"]) if return_topped else res["ET"]


# ------------------------------------------------------------------------------------------------
# PRD product methods (A16)
# ------------------------------------------------------------------------------------------------

_X_ATTRS = {"units": "meters", "long_name": "east_distance_from_radar", "axis": "xy_coordinate",
            "comment": "Distance from radar in east"}
_Y_ATTRS = {"units": "meters", "long_name": "north_distance_from_radar", "axis": "xy_coordinate",
            "comment": "Distance from radar in north"}


def _volume_of(prd, range_mode, field_name="dBZ", fillvalue=FILL):
    range_mode = prd._resolve_field_range_mode(field_name, range_mode=range_mode)
    if hasattr(prd, "get_device_volume"):       # packed on the device, resident across products
        dv = prd.get_device_volume(field_name=field_name, fillvalue=fillvalue, range_mode=range_mode)
        return range_mode, None, None, None, dv, float(prd.radar_height)
    prd.get_vol_data(field_name=field_name, fillvalue=fillvalue, range_mode=range_mode)
    return (range_mode,) + tuple(prd.vol[:5])


def _set_coord(prd, name, values, attrs):
    prd.product.coords[name] = values
    prd.product.coords[name].attrs = dict(attrs)


def _put(prd, name, dims, values, attrs):
    prd.product[name] = (dims, values)
    prd.product[name].attrs = dict(attrs)


# The reference's wrappers end with ``np.where(GridV == fillvalue, np.nan, GridV)`` (NRadar.py:1339,1392,1437): a
# host pass that costs more than the gridding call at C3 size.  Here the gridding call itself stores NaN in the
# empty cells (``missing=np.nan`` -> pcg_get_*_ex) before the result leaves the device.


def _local_grid(prd, XLon, YLat):
    GridLon, GridLat = np.meshgrid(XLon, YLat, indexing="ij")
    if hasattr(prd, "_get_local_projection"):     # NRadar.py:553-558
        gx, gy = prd._get_local_projection()(GridLon, GridLat, inverse=False)
        return np.asarray(gx, dtype=np.float64), np.asarray(gy, dtype=np.float64)
    lon0 = float(prd.scan_info["longitude"].values)
    lat0 = float(prd.scan_info["latitude"].values)
    gx, gy, _ = _project(GridLon, GridLat, lon0, lat0)
    return gx, gy


def add_product_CR_xy(prd, XRange, YRange, range_mode=None):
    """NRadar.py:1322-1356."""
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CR_xy(va, vr, fe, vv, h, np.asarray(GridX, dtype=np.float64), np.asarray(GridY, dtype=np.float64), FILL,
                        prd.effective_earth_radius, missing=np.nan)
    name = prd._product_name("CR", mode)
    _set_coord(prd, "x_cr", XRange, dict(_X_ATTRS, standard_name="CR_product_x_axis "))
    _set_coord(prd, "y_cr", YRange, dict(_Y_ATTRS, standard_name="CR_product_y_axis "))
    _put(prd, name, ("x_cr", "y_cr"), GridV,
         {"units": "dBZ", "standard_name": "Composite_reflectivity_factor",
          "long_name": "Composite_reflectivity_factor", "axis": "xy_coordinate",
          "comment": "Maximum reflectance of all level"})


def add_product_CAPPI_xy(prd, XRange, YRange, level_height, range_mode=None):
    """NRadar.py:1358-1394."""
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CAPPI_xy(va, vr, fe, vv, h, np.asarray(GridX, dtype=np.float64), np.asarray(GridY, dtype=np.float64), level_height, FILL,
                           prd.effective_earth_radius, missing=np.nan)
    name = prd._product_name("CAPPI_%d" % level_height, mode)
    xn, yn = "x_cappi_%d" % level_height, "y_cappi_%d" % level_height
    _set_coord(prd, xn, XRange, dict(_X_ATTRS, standard_name="CAPPI_product_x_axis "))
    _set_coord(prd, yn, YRange, dict(_Y_ATTRS, standard_name="CAPPI_product_y_axis "))
    _put(prd, name, (xn, yn), GridV,
         {"units": "dBZ", "standard_name": "Constant_altitude_plan_position_indicator",
          "long_name": "Constant_altitude_plan_position_indicator", "axis": "xy_coordinate",
          "comment": "CAPPI of level %d m." % level_height})


def add_product_CAPPI_3d_xy(prd, XRange, YRange, level_heights, range_mode=None):
    """NRadar.py:1396-1457."""
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    level_heights = np.asarray(level_heights, dtype=np.float64)
    GridV = G.get_CAPPI_3d(va, vr, fe, vv, h, np.asarray(GridX, dtype=np.float64), np.asarray(GridY, dtype=np.float64), level_heights, FILL,
                           prd.effective_earth_radius, missing=np.nan)
    _set_coord(prd, "z_cappi", level_heights, {"units": "meters", "standard_name": "CAPPI_product_z_axis ",
                                                "long_name": "height_above_mean_sea_level", "axis": "z_coordinate"})
    _set_coord(prd, "x_cappi_3d", XRange, dict(_X_ATTRS, standard_name="CAPPI_product_x_axis "))
    _set_coord(prd, "y_cappi_3d", YRange, dict(_Y_ATTRS, standard_name="CAPPI_product_y_axis "))
    _put(prd, prd._product_name("CAPPI_3D", mode), ("z_cappi", "x_cappi_3d", "y_cappi_3d"), GridV,
         {"units": "dBZ", "standard_name": "Constant_altitude_plan_position_indicator",
          "long_name": "3D_constant_altitude_plan_position_indicator", "axis": "xyz_coordinate",
          "comment": "3D CAPPI volume for the current radar"})


def add_product_CR_lonlat(prd, XLon, YLat, range_mode=None):
    """NRadar.py:1746-1780."""
    GridX, GridY = _local_grid(prd, XLon, YLat)
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CR_xy(va, vr, fe, vv, h, GridX, GridY, FILL, prd.effective_earth_radius, missing=np.nan)
    _set_coord(prd, "lon_cr", XLon, {"units": "degrees", "long_name": "longitude"})
    _set_coord(prd, "lat_cr", YLat, {"units": "degrees", "long_name": "latitude"})
    _put(prd, prd._product_name("CR_geo", mode), ("lon_cr", "lat_cr"), GridV,
         {"units": "dBZ", "standard_name": "Composite_reflectivity_factor_lonlat_grid",
          "long_name": "Composite_reflectivity_factor", "axis": "lonlat_coordinate"})


def add_product_CAPPI_lonlat(prd, XLon, YLat, level_height, range_mode=None):
    """NRadar.py:1782-1818."""
    GridX, GridY = _local_grid(prd, XLon, YLat)
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CAPPI_xy(va, vr, fe, vv, h, GridX, GridY, level_height, FILL, prd.effective_earth_radius,
                           missing=np.nan)
    xn, yn = "lon_cappi_%d" % level_height, "lat_cappi_%d" % level_height
    _set_coord(prd, xn, XLon, {"units": "degrees", "long_name": "longitude"})
    _set_coord(prd, yn, YLat, {"units": "degrees", "long_name": "latitude"})
    _put(prd, prd._product_name("CAPPI_geo_%d" % level_height, mode), (xn, yn), GridV,
         {"units": "dBZ", "standard_name": "Constant_altitude_plan_position_indicator",
          "long_name": "Constant_altitude_plan_position_indicator", "axis": "lonlat_coordinate",
          "comment": "CAPPI of level %d m" % level_height})


def grid_reflectivity_volume_xy(prd, XRange, YRange, level_heights, range_mode=None, blind_method="mask"):
    """NRadar.py:1276-1297 (the volume behind add_product_VIL_xy / ET_xy)."""
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    level_heights = np.asarray(level_heights, dtype=np.float64)
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CAPPI_3d(va, vr, fe, vv, h, np.asarray(GridX, dtype=np.float64), np.asarray(GridY, dtype=np.float64), level_heights, FILL,
                           prd.effective_earth_radius, blind_method)
    return mode, FILL, level_heights, GridV


def _column_products_on(prd, GridX, GridY, level_heights, range_mode, blind_method, min_dbz, max_dbz_cap,
                        threshold_dbz):
    from .runtime import DeviceVolume
    gx = np.ascontiguousarray(GridX, dtype=np.float64)
    gy = np.ascontiguousarray(GridY, dtype=np.float64)
    levels = np.ascontiguousarray(np.asarray(level_heights, dtype=np.float64).ravel())
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    owned = not isinstance(vv, DeviceVolume)
    dv = DeviceVolume(va, vr, fe, vv, fillvalue=FILL) if owned else vv
    nx, ny = gx.shape
    out = {"CR": np.full((nx, ny), np.nan), "VIL": np.full((nx, ny), np.nan), "ET": np.full((nx, ny), np.nan),
           "ET_TOPPED": np.zeros((nx, ny), dtype=np.uint8)}
    try:
        if nx * ny:
            rc = _lib.check(_lib.load().pcg_get_column_products(
                dv.handle, float(h), G._resolve_effective_earth_radius(prd.effective_earth_radius), _lib.as_ptr(gx),
                _lib.as_ptr(gy), nx, ny, levels.ctypes.data_as(_lib.c_double_p), int(levels.size), FILL,
                G._resolve_blind_code(blind_method), float(min_dbz), float(max_dbz_cap), float(threshold_dbz),
                _lib.as_ptr(out["CR"]), _lib.as_ptr(out["VIL"]), _lib.as_ptr(out["ET"]), _lib.as_ptr(out["ET_TOPPED"])))
            if rc == _lib.PCG_WARN_NONFINITE:      # NaN / inf targets: the operation-by-operation path decides
                GridV = G.get_CAPPI_3d(None, None, None, dv, h, gx, gy, levels, FILL, prd.effective_earth_radius,
                                       blind_method)
                out = column_products(GridV, levels, FILL, min_dbz, max_dbz_cap, threshold_dbz)
    finally:
        if owned:
            dv.close()
    return mode, out


def volume_column_products(prd, XRange, YRange, level_heights, range_mode=None, blind_method="mask", min_dbz=18.0,
                           max_dbz_cap=56.0, threshold_dbz=18.0):
    """CR / VIL / ET / ET_TOPPED planes of one radar: ``_grid_reflectivity_volume_xy`` + ``derive_*`` (NRadar.py:
    1276-1297,1459-1572) in one device call — the gridded volume never leaves the GPU (``pcg_get_column_products``)."""
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    return _column_products_on(prd, GridX, GridY, level_heights, range_mode, blind_method, min_dbz, max_dbz_cap,
                               threshold_dbz)


def grid_reflectivity_volume_lonlat(prd, XLon, YLat, level_heights, range_mode=None, blind_method="mask"):
    """NRadar.py:1298-1320 (the volume behind add_product_VIL_lonlat / ET_lonlat)."""
    GridX, GridY = _local_grid(prd, XLon, YLat)
    level_heights = np.asarray(level_heights, dtype=np.float64)
    mode, va, vr, fe, vv, h = _volume_of(prd, range_mode)
    GridV = G.get_CAPPI_3d(va, vr, fe, vv, h, GridX, GridY, level_heights, FILL, prd.effective_earth_radius,
                           blind_method)
    return mode, FILL, level_heights, GridV


def volume_column_products_lonlat(prd, XLon, YLat, level_heights, range_mode=None, blind_method="mask", min_dbz=18.0,
                                  max_dbz_cap=56.0, threshold_dbz=18.0):
    """The lon/lat twin of :func:`volume_column_products` (NRadar.py:1298-1320 + derive_*)."""
    GridX, GridY = _local_grid(prd, XLon, YLat)
    return _column_products_on(prd, GridX, GridY, level_heights, range_mode, blind_method, min_dbz, max_dbz_cap,
                               threshold_dbz)


# PRODUCT_REFERENCE_NOTES of pycwr/core/RadarProduct.py:4-13 (attribute payload of the VIL / ET products)
PRODUCT_REFERENCE_NOTES = {
    "VIL": [
        "Greene and Clark (1972), Monthly Weather Review, doi:10.1175/1520-0493(1972)100<0548:VILWNA>2.3.CO;2",
        "NOAA WDTD VIL product guidance: https://vlab.noaa.gov/web/wdtd/-/vertically-integrated-liquid-vil-",
    ],
    "ET": [
        "Lakshmanan et al. (2013), Weather and Forecasting, doi:10.1175/WAF-D-12-00084.1",
        "NOAA WDTD xx dBZ Echo Top guidance: https://vlab.noaa.gov/web/wdtd/-/xx-dbz-echo-top-et-",
        "WSR-88D ROC ICD 2620003Y (Enhanced Echo Tops): https://www.roc.noaa.gov/public-documents/icds/2620003Y.pdf",
    ],
}

# what differs between the Cartesian and the lon/lat flavour of a column product: coordinate names and attrs,
# product-name suffix, standard_name suffix, axis label (NRadar.py:1459-1572 vs :1820-1929)
_FLAVOUR = {
    False: {"xc": "x_%s", "yc": "y_%s", "suffix": "", "std": "", "axis": "xy_coordinate",
            "x": lambda tag: {"units": "meters", "standard_name": "%s_product_x_axis " % tag,
                              "long_name": "east_distance_from_radar", "axis": "xy_coordinate",
                              "comment": "Distance from radar in east"},
            "y": lambda tag: {"units": "meters", "standard_name": "%s_product_y_axis " % tag,
                              "long_name": "north_distance_from_radar", "axis": "xy_coordinate",
                              "comment": "Distance from radar in north"}},
    True: {"xc": "lon_%s", "yc": "lat_%s", "suffix": "_geo", "std": "_lonlat_grid", "axis": "lonlat_coordinate",
           "x": lambda tag: {"units": "degrees", "standard_name": "%s_product_lon_axis " % tag,
                             "long_name": "longitude_%s" % tag.lower(), "axis": "lonlat_coordinate"},
           "y": lambda tag: {"units": "degrees", "standard_name": "%s_product_lat_axis " % tag,
                             "long_name": "latitude_%s" % tag.lower(), "axis": "lonlat_coordinate"}},
}


def _add_column_product(prd, geo, tag, X, Y, level_heights, range_mode, blind_method, min_dbz=18.0, max_dbz_cap=56.0,
                        threshold_dbz=18.0):
    """One of VIL / ET on either grid flavour: the volume is gridded and reduced in one device call
    (``pcg_get_column_products``) and the planes are written with the reference's names and attrs."""
    import json
    fl = _FLAVOUR[bool(geo)]
    levels = np.asarray(level_heights, dtype=np.float64)
    call = volume_column_products_lonlat if geo else volume_column_products
    mode, prod = call(prd, X, Y, levels, range_mode, blind_method, min_dbz, max_dbz_cap, threshold_dbz)
    xc, yc = fl["xc"] % tag.lower(), fl["yc"] % tag.lower()
    _set_coord(prd, xc, X, fl["x"](tag))
    _set_coord(prd, yc, Y, fl["y"](tag))
    src = json.dumps(levels.tolist())
    refs = json.dumps(PRODUCT_REFERENCE_NOTES[tag], ensure_ascii=False)
    if tag == "VIL":
        _put(prd, prd._product_name("VIL" + fl["suffix"], mode), (xc, yc), prod["VIL"],
             {"units": "kg m-2", "standard_name": "Vertically_integrated_liquid_water" + fl["std"],
              "long_name": "Vertically_integrated_liquid_water", "axis": fl["axis"], "source_levels": src,
              "min_dbz": float(min_dbz), "max_dbz_cap": float(max_dbz_cap), "references": refs})
        return
    _put(prd, prd._product_name("ET" + fl["suffix"], mode), (xc, yc), prod["ET"],
         {"units": "meters", "standard_name": "Echo_top_height" + fl["std"], "long_name": "Echo_top_height",
          "axis": fl["axis"], "source_levels": src, "threshold_dbz": float(threshold_dbz), "references": refs})
    _put(prd, prd._product_name("ET_TOPPED" + fl["suffix"], mode), (xc, yc),
         np.asarray(prod["ET_TOPPED"], dtype=np.uint8),
         {"units": "1", "standard_name": "Echo_top_topped_flag" + fl["std"], "long_name": "Echo_top_topped_flag",
          "axis": fl["axis"], "flag_values": np.array([0, 1], dtype=np.uint8),
          "flag_meanings": "resolved topped_at_highest_sampled_level", "source_levels": src,
          "threshold_dbz": float(threshold_dbz), "references": refs})


def add_product_VIL_xy(prd, XRange, YRange, level_heights, range_mode=None, blind_method="mask", min_dbz=18.0,
                       max_dbz_cap=56.0):
    """NRadar.py:1459-1506."""
    _add_column_product(prd, False, "VIL", XRange, YRange, level_heights, range_mode, blind_method, min_dbz=min_dbz,
                        max_dbz_cap=max_dbz_cap)


def add_product_ET_xy(prd, XRange, YRange, level_heights, range_mode=None, blind_method="mask", threshold_dbz=18.0):
    """NRadar.py:1508-1572."""
    _add_column_product(prd, False, "ET", XRange, YRange, level_heights, range_mode, blind_method,
                        threshold_dbz=threshold_dbz)


def add_product_VIL_lonlat(prd, XLon, YLat, level_heights, range_mode=None, blind_method="mask", min_dbz=18.0,
                           max_dbz_cap=56.0):
    """NRadar.py:1820-1866."""
    _add_column_product(prd, True, "VIL", XLon, YLat, level_heights, range_mode, blind_method, min_dbz=min_dbz,
                        max_dbz_cap=max_dbz_cap)


def add_product_ET_lonlat(prd, XLon, YLat, level_heights, range_mode=None, blind_method="mask", threshold_dbz=18.0):
    """NRadar.py:1868-1929."""
    _add_column_product(prd, True, "ET", XLon, YLat, level_heights, range_mode, blind_method,
                        threshold_dbz=threshold_dbz)


def add_product_VIL_ET_xy(prd, XRange, YRange, level_heights, range_mode=None, blind_method="mask", min_dbz=18.0,
                          max_dbz_cap=56.0, threshold_dbz=18.0):
    """VIL, ET and ET_TOPPED of one radar from ONE gridded volume (the reference grids it twice, once per
    ``add_product_VIL_xy`` / ``add_product_ET_xy`` call, NRadar.py:1459-1572); same names and attrs."""
    add_product_VIL_xy(prd, XRange, YRange, level_heights, range_mode, blind_method, min_dbz, max_dbz_cap)
    add_product_ET_xy(prd, XRange, YRange, level_heights, range_mode, blind_method, threshold_dbz)


def grid_3d_network_xy(radars, XRange, YRange, level_heights, field_name="dBZ", fillvalue=FILL, lon_0=None,
                       lat_0=None, range_mode=None):
    """Maximum-value multi-radar CAPPI network on a shared Cartesian grid (NRadar.py:2042-2122)."""
    if not radars:
        raise ValueError("radars must contain at least one PRD instance.")
    XRange = np.asarray(XRange, dtype=np.float64)
    YRange = np.asarray(YRange, dtype=np.float64)
    level_heights = np.asarray(level_heights, dtype=np.float64)
    lons = np.array([float(r.scan_info["longitude"].values) for r in radars], dtype=np.float64)
    lats = np.array([float(r.scan_info["latitude"].values) for r in radars], dtype=np.float64)
    lon_0 = float(np.mean(lons)) if lon_0 is None else lon_0
    lat_0 = float(np.mean(lats)) if lat_0 is None else lat_0
    GridX, GridY = np.meshgrid(XRange, YRange, indexing="ij")
    va, vr, fe, vv, radii = [], [], [], [], []
    n = len(radars)
    rx, ry, rh = np.empty(n), np.empty(n), np.empty(n)
    for i, radar in enumerate(radars):
        mode = radar._resolve_field_range_mode(field_name, range_mode=range_mode)
        radar.get_vol_data(field_name=field_name, fillvalue=fillvalue, range_mode=mode)
        a, r, e, v, h, lon, lat = radar.vol
        x, y, _ = _project(np.float64(lon), np.float64(lat), lon_0, lat_0)
        rx[i], ry[i], rh[i] = float(x), float(y), h
        radii.append(getattr(radar, "effective_earth_radius", None))
        va.append(a); vr.append(r); fe.append(e); vv.append(v)
    GridV = G.get_mosaic_CAPPI_3d(va, vr, fe, vv, rx, ry, rh, np.asarray(GridX, dtype=np.float64), np.asarray(GridY, dtype=np.float64),
                                  level_heights, fillvalue, radii)
    product = _new_dataset({"z": level_heights, "x": XRange, "y": YRange})
    product["network_3d"] = (("z", "x", "y"), np.where(GridV == fillvalue, np.nan, GridV))
    product["network_3d"].attrs = {
        "units": "dBZ", "long_name": "3D_multi_radar_CAPPI_network",
        "comment": "Maximum-value multi-radar CAPPI mosaic on a shared Cartesian grid",
        "grid_origin_lon": lon_0, "grid_origin_lat": lat_0, "field_name": field_name,
        "range_mode": radars[0]._resolve_field_range_mode(field_name, range_mode=range_mode),
    }
    return product


class B200ProductMixin(object):
    """Mix into ``pycwr.core.NRadar.PRD`` (before it in the MRO) to run its gridding products on the GPU."""
    add_product_CR_xy = add_product_CR_xy
    add_product_CAPPI_xy = add_product_CAPPI_xy
    add_product_CAPPI_3d_xy = add_product_CAPPI_3d_xy
    add_product_CR_lonlat = add_product_CR_lonlat
    add_product_CAPPI_lonlat = add_product_CAPPI_lonlat
    _grid_reflectivity_volume_xy = grid_reflectivity_volume_xy
    _grid_reflectivity_volume_lonlat = grid_reflectivity_volume_lonlat
    add_product_VIL_xy = add_product_VIL_xy
    add_product_ET_xy = add_product_ET_xy
    add_product_VIL_lonlat = add_product_VIL_lonlat
    add_product_ET_lonlat = add_product_ET_lonlat

    @staticmethod
    def _interpolate_ppi_strip_arrays(azimuth, ranges, values, target_azimuth, target_range):
        """``extract_section`` / ``_extract_ppi_rhi`` of the reference keep their loops and call this per
        sweep (``NRadar.py:1061-1063,1153-1155``); ``pycwr_b200.section`` has the one-launch versions."""
        from . import section
        return section._interpolate_ppi_strip_arrays(azimuth, ranges, values, target_azimuth, target_range)
