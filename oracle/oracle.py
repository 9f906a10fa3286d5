"""ctypes front end of ``oracle/grid_oracle.c`` — TEST INFRASTRUCTURE ONLY.

The functions mirror the reference module ``pycwr/core/RadarGridC.pyx`` (same
names, same positional order) so that parity tests read like the reference's
own tests, and add ``return_index=True`` to expose the per-voxel interpolation
records (state, lower/upper sweep, iaz, ir, flags) that the reference never
returns.  Parity status: PINNED — see tests/test_oracle_pinning.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgrid_oracle.so")
_SRC = os.path.join(_HERE, "grid_oracle.c")

IDX_STRIDE = 10
DEFAULT_RE = 8494666.6666666661  # RadarGridC.pyx:9
BLIND_CODES = {"mask": 0, "nearest_gate": 1, "lowest_sweep": 2, "hybrid": 3, "nearest_valid": 4}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_i32p = ctypes.POINTER(ctypes.c_int32)


class _Volume(ctypes.Structure):
    _fields_ = [
        ("ne", ctypes.c_int),
        ("naz", _ip),
        ("ng", _ip),
        ("az", ctypes.POINTER(_dp)),
        ("rng", ctypes.POINTER(_dp)),
        ("val", ctypes.POINTER(_dp)),
        ("fix_elev", _dp),
    ]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-shared", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.og_xy_to_azimuth.restype = ctypes.c_double
        L.og_xy_to_azimuth.argtypes = [ctypes.c_double] * 2
        L.og_interp_linear.restype = ctypes.c_double
        L.og_interp_linear.argtypes = [ctypes.c_double] * 6
        L.og_interp_ppi.restype = ctypes.c_double
        L.og_interp_ppi.argtypes = [ctypes.c_double] * 11
        for name in ("og_antenna_to_cartesian", "og_xye_to_antenna", "og_cartesian_to_antenna"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [ctypes.c_double] * 5 + [_dp] * 3
        _lib = L
    return _lib


def resolve_re(effective_earth_radius):
    if effective_earth_radius is None:
        return DEFAULT_RE
    radius = float(effective_earth_radius)
    if radius <= 0.0:
        raise ValueError("effective_earth_radius must be positive")
    return radius


def resolve_blind(blind_method):
    method = "mask" if blind_method is None else str(blind_method).lower()
    if method not in BLIND_CODES:
        raise ValueError(
            "blind_method must be one of 'mask', 'nearest_gate', 'lowest_sweep', "
            "'hybrid', or 'nearest_valid'."
        )
    return BLIND_CODES[method]


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PackedVolume:
    """Keeps the numpy arrays alive and exposes an ``og_volume`` struct."""

    def __init__(self, vol_azimuth, vol_range, fix_elevation, vol_value):
        self.az = [_c(a) for a in vol_azimuth]
        self.rng = [_c(a) for a in vol_range]
        self.val = [_c(a) for a in vol_value]
        self.fix = _c(fix_elevation)
        ne = int(self.fix.shape[0])
        if not (len(self.az) >= ne and len(self.rng) >= ne and len(self.val) >= ne):
            raise ValueError("sweep lists shorter than fix_elevation")
        self.ne = ne
        self.naz = np.array([a.shape[0] for a in self.az[:ne]], dtype=np.int32)
        self.ng = np.array([a.shape[0] for a in self.rng[:ne]], dtype=np.int32)
        for k in range(ne):
            if self.val[k].shape != (self.naz[k], self.ng[k]):
                raise ValueError("sweep %d: value shape does not match azimuth/range" % k)
        self._azp = (_dp * max(ne, 1))(*[a.ctypes.data_as(_dp) for a in self.az[:ne]])
        self._rgp = (_dp * max(ne, 1))(*[a.ctypes.data_as(_dp) for a in self.rng[:ne]])
        self._vlp = (_dp * max(ne, 1))(*[a.ctypes.data_as(_dp) for a in self.val[:ne]])
        self.struct = _Volume(
            ne,
            self.naz.ctypes.data_as(_ip),
            self.ng.ctypes.data_as(_ip),
            self._azp,
            self._rgp,
            self._vlp,
            self.fix.ctypes.data_as(_dp),
        )


def _scalar3(fn, *args):
    a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    fn(*args, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def antenna_to_cartesian(ranges, azimuth, elevation, h, effective_earth_radius=None):
    return _scalar3(lib().og_antenna_to_cartesian, ranges, azimuth, elevation, h,
                    resolve_re(effective_earth_radius))


def xye_to_antenna(x, y, elevation, h, effective_earth_radius=None):
    return _scalar3(lib().og_xye_to_antenna, x, y, elevation, h, resolve_re(effective_earth_radius))


def cartesian_to_antenna(x, y, z, h, effective_earth_radius=None):
    return _scalar3(lib().og_cartesian_to_antenna, x, y, z, h, resolve_re(effective_earth_radius))


def xy_to_azimuth(x, y):
    return lib().og_xy_to_azimuth(x, y)


def interp_azimuth(az, az_0, az_1, dat_0, dat_1, fillvalue):
    return lib().og_interp_linear(az, az_0, az_1, dat_0, dat_1, fillvalue)


def interp_ppi(az, r, az_0, az_1, r_0, r_1, m00, m01, m10, m11, fillvalue):
    return lib().og_interp_ppi(az, r, az_0, az_1, r_0, r_1, m00, m01, m10, m11, fillvalue)


def ppi_to_grid(azimuth, ranges, elevation, mat_ppi, radar_height, GridX, GridY, fillvalue,
                effective_earth_radius=None, blind_method="mask", return_index=False):
    az, rg, val = _c(azimuth), _c(ranges), _c(mat_ppi)
    gx, gy = _c(GridX), _c(GridY)
    shape = (gx.shape[0], gy.shape[1])
    out = np.empty(shape, dtype=np.float64)
    idx = np.empty(shape + (IDX_STRIDE,), dtype=np.int32) if return_index else None
    lib().og_ppi_to_grid(
        az.ctypes.data_as(_dp), ctypes.c_int(az.shape[0]), rg.ctypes.data_as(_dp),
        ctypes.c_int(rg.shape[0]), ctypes.c_double(elevation), val.ctypes.data_as(_dp),
        ctypes.c_double(radar_height), ctypes.c_double(resolve_re(effective_earth_radius)),
        gx.ctypes.data_as(_dp), gy.ctypes.data_as(_dp), ctypes.c_long(out.size),
        ctypes.c_double(fillvalue), ctypes.c_int(resolve_blind(blind_method)),
        out.ctypes.data_as(_dp), idx.ctypes.data_as(_i32p) if idx is not None else None,
    )
    return (out, idx) if return_index else out


def get_CR_xy(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
              fillvalue, effective_earth_radius=None):
    pv = PackedVolume(vol_azimuth, vol_range, fix_elevation, vol_value)
    gx, gy = _c(GridX), _c(GridY)
    out = np.empty((gx.shape[0], gy.shape[1]), dtype=np.float64)
    lib().og_get_cr(
        ctypes.byref(pv.struct), ctypes.c_double(radar_height),
        ctypes.c_double(resolve_re(effective_earth_radius)), gx.ctypes.data_as(_dp),
        gy.ctypes.data_as(_dp), ctypes.c_long(out.size), ctypes.c_double(fillvalue),
        out.ctypes.data_as(_dp),
    )
    return out


def get_CAPPI_3d(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
                 level_heights, fillvalue, effective_earth_radius=None, blind_method="mask",
                 return_index=False, _beam_width_deg=1.0):
    pv = PackedVolume(vol_azimuth, vol_range, fix_elevation, vol_value)
    gx, gy, lv = _c(GridX), _c(GridY), _c(level_heights)
    shape = (lv.shape[0], gx.shape[0], gx.shape[1])
    out = np.empty(shape, dtype=np.float64)
    idx = np.empty(shape + (IDX_STRIDE,), dtype=np.int32) if return_index else None
    lib().og_get_cappi3d(
        ctypes.byref(pv.struct), ctypes.c_double(radar_height),
        ctypes.c_double(resolve_re(effective_earth_radius)), gx.ctypes.data_as(_dp),
        gy.ctypes.data_as(_dp), ctypes.c_long(gx.size), lv.ctypes.data_as(_dp),
        ctypes.c_int(lv.shape[0]), ctypes.c_double(fillvalue),
        ctypes.c_int(resolve_blind(blind_method)), ctypes.c_double(_beam_width_deg),
        out.ctypes.data_as(_dp), idx.ctypes.data_as(_i32p) if idx is not None else None,
    )
    return (out, idx) if return_index else out


def get_CAPPI_xy(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
                 level_height, fillvalue, effective_earth_radius=None, blind_method="mask",
                 beam_width_deg=1.0, return_index=False):
    res = get_CAPPI_3d(vol_azimuth, vol_range, fix_elevation, vol_value, radar_height, GridX, GridY,
                       np.array([level_height], dtype=np.float64), fillvalue, effective_earth_radius,
                       blind_method, return_index=return_index, _beam_width_deg=float(beam_width_deg))
    if return_index:
        return res[0][0], res[1][0]
    return res[0]


def get_mosaic_CAPPI_3d(vol_azimuth, vol_range, fix_elevation, vol_value, radar_x, radar_y,
                        radar_height, GridX, GridY, level_heights, fillvalue,
                        effective_earth_radius=None):
    radar_x, radar_y, radar_height = _c(radar_x), _c(radar_y), _c(radar_height)
    nradar = radar_height.shape[0]
    if not (len(vol_azimuth) == len(vol_range) == len(fix_elevation) == len(vol_value) == nradar):
        raise ValueError("Radar volume lists and radar position arrays must have the same length.")
    if effective_earth_radius is None:
        radii = [None] * nradar
    elif not isinstance(effective_earth_radius, (list, tuple)):
        radii = [effective_earth_radius] * nradar
    elif len(effective_earth_radius) != nradar:
        raise ValueError("effective_earth_radius must be None, a scalar, or a sequence matching the radar count.")
    else:
        radii = list(effective_earth_radius)
    re = np.array([resolve_re(r) for r in radii], dtype=np.float64)
    pvs = [PackedVolume(vol_azimuth[i], vol_range[i], fix_elevation[i], vol_value[i]) for i in range(nradar)]
    structs = (_Volume * max(nradar, 1))(*[p.struct for p in pvs])
    gx, gy, lv = _c(GridX), _c(GridY), _c(level_heights)
    out = np.empty((lv.shape[0], gx.shape[0], gx.shape[1]), dtype=np.float64)
    scratch = np.empty(2 * gx.size + out.size, dtype=np.float64)
    lib().og_get_mosaic_cappi3d(
        ctypes.c_int(nradar), structs, radar_x.ctypes.data_as(_dp), radar_y.ctypes.data_as(_dp),
        radar_height.ctypes.data_as(_dp), re.ctypes.data_as(_dp), gx.ctypes.data_as(_dp),
        gy.ctypes.data_as(_dp), ctypes.c_long(gx.size), lv.ctypes.data_as(_dp),
        ctypes.c_int(lv.shape[0]), ctypes.c_double(fillvalue), out.ctypes.data_as(_dp),
        scratch.ctypes.data_as(_dp),
    )
    return out


def bulk_cartesian_to_antenna(x, y, z, h, effective_earth_radius=None):
    x, y = _c(x).ravel(), _c(y).ravel()
    az, r, el = (np.empty_like(x) for _ in range(3))
    lib().og_bulk_cartesian_to_antenna(
        x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), ctypes.c_long(x.size), ctypes.c_double(z),
        ctypes.c_double(h), ctypes.c_double(resolve_re(effective_earth_radius)),
        az.ctypes.data_as(_dp), r.ctypes.data_as(_dp), el.ctypes.data_as(_dp))
    return az, r, el


def bulk_xye_to_antenna(x, y, elevation, h, effective_earth_radius=None):
    x, y = _c(x).ravel(), _c(y).ravel()
    az, r, z = (np.empty_like(x) for _ in range(3))
    lib().og_bulk_xye_to_antenna(
        x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), ctypes.c_long(x.size),
        ctypes.c_double(elevation), ctypes.c_double(h),
        ctypes.c_double(resolve_re(effective_earth_radius)),
        az.ctypes.data_as(_dp), r.ctypes.data_as(_dp), z.ctypes.data_as(_dp))
    return az, r, z
