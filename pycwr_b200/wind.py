"""Host-side mirror of the gridding step of ``retrieve_wind_volume_xy``
(``pycwr/retrieve/WindField.py``): per-sweep VVP samples gathered onto the target horizontal grid
(``_grid_vvp_retrieval_xy`` :629-689) and the per-column vertical reconstruction
(``_reconstruct_wind_volume_chunk`` :1183-1229 over ``_vertical_profile_from_sweeps`` :692-810).
Same names, arguments, defaults, output keys / dtypes and errors; the neighbour search and the sums
run on the GPU (no cKDTree, no per-target Python loop).  Upstream of it, ``_run_local_vvp`` (:1843-1941,
SURVEY.md §8 row f4) — the moving-window least-squares retrieval that produces those samples — runs as one
kernel launch per sweep (one warp per window) instead of a Python loop over every ray and gate."""
import numpy as np

from . import _lib

_SAMPLE_KEYS = ("x", "y", "z", "u", "v", "fit_rmse", "valid_count", "valid_fraction", "condition_number",
                "azimuth_sector_count")


def _result_array(retrieval, key):
    value = retrieval[key]
    if hasattr(value, "values"):
        value = value.values
    return np.asarray(value, dtype=np.float64).reshape(-1)


def _flatten_valid_vvp_samples(retrieval):
    """WindField.py:561-594 (host logic: a mask over a few thousand samples)."""
    cols = {k: _result_array(retrieval, k) for k in _SAMPLE_KEYS}
    mask = np.ones(cols["x"].shape, dtype=bool)
    for k in ("x", "y", "z", "u", "v", "fit_rmse", "valid_count", "valid_fraction"):
        mask &= np.isfinite(cols[k])
    return {k: v[mask] for k, v in cols.items()}


def _grid_vvp_retrieval_xy(retrieval, target_x, target_y, horizontal_radius_m, horizontal_min_neighbors=3,
                           max_horizontal_radius_m=None, distance_power=2.0):
    """Grid one sweep VVP analysis onto the target horizontal grid (WindField.py:629-689)."""
    target_x = np.asarray(target_x, dtype=np.float64)
    target_y = np.asarray(target_y, dtype=np.float64)
    if target_x.shape != target_y.shape:
        raise ValueError("target_x and target_y must have the same shape")
    shape = target_x.shape
    outputs = {
        "u": np.full(shape, np.nan, dtype=np.float64),
        "v": np.full(shape, np.nan, dtype=np.float64),
        "z": np.full(shape, np.nan, dtype=np.float64),
        "fit_rmse": np.full(shape, np.nan, dtype=np.float64),
        "valid_count": np.zeros(shape, dtype=np.int32),
        "neighbor_count": np.zeros(shape, dtype=np.int32),
        "condition_number": np.full(shape, np.nan, dtype=np.float64),
        "azimuth_sector_count": np.zeros(shape, dtype=np.int32),
        "expanded_radius": np.zeros(shape, dtype=np.uint8),
    }
    samples = _flatten_valid_vvp_samples(retrieval)
    ns = int(samples["x"].size)
    nt = int(target_x.size)
    if ns == 0 or nt == 0:
        return outputs
    initial = float(horizontal_radius_m)
    if max_horizontal_radius_m is None:
        max_horizontal_radius_m = max(initial, 2.0 * initial)
    soa = np.ascontiguousarray(np.stack([samples[k] for k in _SAMPLE_KEYS], axis=0))
    tx = np.ascontiguousarray(target_x.reshape(-1))
    ty = np.ascontiguousarray(target_y.reshape(-1))
    f64 = np.empty((5, nt), dtype=np.float64)
    i32 = np.empty((3, nt), dtype=np.int32)
    exp = np.empty(nt, dtype=np.uint8)
    _lib.check(_lib.load().pcg_wind_gather(
        _lib.as_ptr(soa), ns, _lib.as_ptr(tx), _lib.as_ptr(ty), nt, initial, float(max_horizontal_radius_m),
        int(horizontal_min_neighbors), float(distance_power), _lib.as_ptr(f64), _lib.as_ptr(i32), _lib.as_ptr(exp)))
    for k, name in enumerate(("u", "v", "z", "fit_rmse", "condition_number")):
        outputs[name] = f64[k].reshape(shape)
    for k, name in enumerate(("valid_count", "neighbor_count", "azimuth_sector_count")):
        outputs[name] = i32[k].reshape(shape)
    outputs["expanded_radius"] = exp.reshape(shape)
    return outputs


def _normalize_level_heights(level_heights):
    """WindField.py:448-457."""
    levels = np.asarray(level_heights, dtype=np.float64).reshape(-1)
    if levels.size == 0:
        raise ValueError("level_heights must contain at least one value")
    if not np.all(np.isfinite(levels)):
        raise ValueError("level_heights must be finite")
    if levels.size > 1 and np.any(np.diff(levels) <= 0.0):
        raise ValueError("level_heights must be strictly increasing")
    return levels


def reconstruct_wind_volume(sweep_z, sweep_u, sweep_v, sweep_fit_rmse, sweep_valid_count, sweep_neighbor_count,
                            sweep_expanded_radius, level_heights, vertical_tolerance_m, max_vertical_gap_m):
    """Every column of the target volume from the per-sweep gridded layers ``(nsweep, nx, ny)``
    (WindField.py:1183-1229 for the whole grid instead of one x-slab per pool worker).  Returns the
    reference's ten ``(nz, nx, ny)`` arrays."""
    levels = _normalize_level_heights(level_heights)
    layers = [np.asarray(a, dtype=np.float64) for a in (sweep_z, sweep_u, sweep_v, sweep_fit_rmse, sweep_valid_count,
                                                        sweep_neighbor_count, sweep_expanded_radius)]
    shape = layers[0].shape
    if len(shape) != 3 or any(a.shape != shape for a in layers):
        raise ValueError("sweep layers must share the shape (nsweep, nx, ny)")
    nsweep, nx, ny = shape
    ncol, nz = nx * ny, int(levels.size)
    stack = np.ascontiguousarray(np.stack(layers, axis=0))
    f64 = np.empty((5, nz, nx, ny), dtype=np.float64)
    i32 = np.empty((4, nz, nx, ny), dtype=np.int32)
    flag = np.empty((nz, nx, ny), dtype=np.uint8)
    if ncol:
        _lib.check(_lib.load().pcg_wind_vertical(
            _lib.as_ptr(stack), nsweep, ncol, levels.ctypes.data_as(_lib.c_double_p), nz, float(vertical_tolerance_m),
            float(max_vertical_gap_m), _lib.as_ptr(f64), _lib.as_ptr(i32), _lib.as_ptr(flag)))
    return {
        "u": f64[0], "v": f64[1], "fit_rmse": f64[2], "valid_count": i32[0], "source_sweep_count": i32[1],
        "neighbor_count": i32[2], "effective_vertical_support": i32[3], "sweep_spread_m": f64[3],
        "quality_score": f64[4], "quality_flag": flag,
    }


def grid_wind_volume(retrievals, target_x, target_y, level_heights, horizontal_radius_m, vertical_tolerance_m,
                     max_vertical_gap_m, horizontal_min_neighbors=3, max_horizontal_radius_m=None, distance_power=2.0):
    """The two stages back to back for a list of per-sweep VVP retrievals: what
    ``_retrieve_wind_volume_to_target_grid`` (WindField.py:1232-1492) does between the VVP solver and the
    dataset assembly."""
    layers = [_grid_vvp_retrieval_xy(r, target_x, target_y, horizontal_radius_m, horizontal_min_neighbors,
                                     max_horizontal_radius_m, distance_power) for r in retrievals]
    stack = lambda key: np.stack([np.asarray(l[key], dtype=np.float64) for l in layers], axis=0)
    return reconstruct_wind_volume(stack("z"), stack("u"), stack("v"), stack("fit_rmse"), stack("valid_count"),
                                   stack("neighbor_count"), stack("expanded_radius"), level_heights,
                                   vertical_tolerance_m, max_vertical_gap_m)


# ------------------------------------------------------------------------------------------------
# f4: local moving-window VVP on one sweep
# ------------------------------------------------------------------------------------------------

_VVP_F64 = ("u", "v", "wind_speed", "wind_direction", "radial_offset", "fit_rmse", "condition_number",
            "valid_fraction", "azimuth_coverage_deg")
_VVP_I32 = ("azimuth_sector_count", "valid_count")


def _run_local_vvp(azimuth, elevation, radial_velocity, az_num, bin_num, fillvalue=-999.0, min_valid_fraction=0.7,
                   min_valid_count=20, min_azimuth_coverage_deg=0.0, min_azimuth_sector_count=2,
                   max_condition_number=1.0e9):
    """``_run_local_vvp`` (``WindField.py:1843-1941``): dict of ``(naz, nrange)`` arrays with the reference's
    keys and dtypes."""
    if az_num % 2 != 1:
        raise ValueError("az_num must be odd")
    if bin_num % 2 != 1:
        raise ValueError("bin_num must be odd")
    vel = np.asarray(radial_velocity, dtype=np.float64)
    if fillvalue is not None:                                              # _as_float_array :50-55
        vel = np.where(vel == float(fillvalue), np.nan, vel)
    vel = np.ascontiguousarray(vel)
    azimuth = np.ascontiguousarray(azimuth, dtype=np.float64).reshape(-1)
    if vel.ndim != 2 or azimuth.size != vel.shape[0]:
        raise ValueError("azimuth length must match the first radial_velocity dimension")
    if np.asarray(elevation).ndim == 0:
        elevation = np.full(azimuth.shape, float(elevation), dtype=np.float64)
    else:
        elevation = np.ascontiguousarray(elevation, dtype=np.float64).reshape(-1)
    if elevation.shape != azimuth.shape:
        raise ValueError("elevation must be scalar or match the azimuth shape")
    naz, nrange = vel.shape
    f64 = np.empty((len(_VVP_F64), naz, nrange), dtype=np.float64)
    i32 = np.empty((len(_VVP_I32), naz, nrange), dtype=np.int32)
    _lib.check(_lib.load().pcg_vvp_local(
        azimuth.ctypes.data, elevation.ctypes.data, vel.ctypes.data, int(naz), int(nrange), int(az_num), int(bin_num),
        float(min_valid_fraction), int(min_valid_count), float(min_azimuth_coverage_deg),
        int(min_azimuth_sector_count), float(max_condition_number), f64.ctypes.data, i32.ctypes.data))
    out = {k: f64[i] for i, k in enumerate(_VVP_F64)}
    out.update({k: i32[i] for i, k in enumerate(_VVP_I32)})
    return out
