"""NumPy restatement of the local moving-window VVP retrieval — TEST INFRASTRUCTURE ONLY.

Follows ``pycwr/retrieve/WindField.py``: ``_run_local_vvp`` (:1843-1941) and its helpers
``_azimuth_coverage_deg`` (:58-68), ``_azimuth_sector_count`` (:71-81), ``_effective_min_sector_count``
(:84-90), ``_design_matrix`` (:93-106), ``_wind_direction_from_uv`` (:109-112),
``_solve_weighted_least_squares`` (:115-200) and ``_vvp_window_weights`` (:1832-1840).

Written window by window like the reference, but with the window statistics taken per RAY (a ray either
holds a finite sample or it does not) and the 3x3 normal equations accumulated explicitly, which is
how the CUDA kernel is organised; LAPACK is used only for the condition number and the 3x3 solve.
Parity status: PINNED against ``tests/golden/vvp_reference.npz`` (outputs of the reference's own
``_run_local_vvp``, ``tests/golden/make_vvp_golden.py``): counts, valid fraction, coverage, sector count
and NaN masks bit-exact; fitted quantities to 1e-9 (summation order of the normal equations).
"""
import numpy as np

F64_KEYS = ("u", "v", "wind_speed", "wind_direction", "radial_offset", "fit_rmse", "condition_number",
            "valid_fraction", "azimuth_coverage_deg")
I32_KEYS = ("azimuth_sector_count", "valid_count")


def window_weights(az_num, bin_num):
    da = np.arange(az_num, dtype=np.float64) - (az_num - 1) / 2.0
    db = np.arange(bin_num, dtype=np.float64) - (bin_num - 1) / 2.0
    wa = np.exp(-0.5 * (da / max(1.0, (az_num - 1) / 3.0)) ** 2)
    wb = np.exp(-0.5 * (db / max(1.0, (bin_num - 1) / 3.0)) ** 2)
    return wa, wb


def _coverage(azm_rows):
    """360 minus the widest gap between consecutive distinct azimuths (mod 360) of the occupied rays."""
    a = np.unique(azm_rows[np.isfinite(azm_rows)])
    if a.size < 2:
        return 0.0
    gaps = np.diff(np.concatenate((a, [a[0] + 360.0])))
    return float(max(0.0, 360.0 - gaps.max()))


def _sectors(azm_rows):
    a = azm_rows[np.isfinite(azm_rows)]
    return int(np.unique(np.floor(a / 45.0).astype(np.int32)).size)


def _solve_once(X, y, w, max_cond):
    sq = np.sqrt(w)
    Xw = X * sq[:, None]
    yw = y * sq
    N = Xw.T @ Xw
    cond = float(np.linalg.cond(N))
    if not np.isfinite(cond) or cond > max_cond:
        return None, cond
    try:
        return np.linalg.solve(N, Xw.T @ yw), cond
    except np.linalg.LinAlgError:
        return None, cond


def _fit(X, y, w, max_cond):
    """(coefficients or None, condition number, rmse) of the Huber-reweighted fit (:115-200)."""
    c, cond = _solve_once(X, y, w, max_cond)
    if c is None:
        return None, cond, np.nan
    for _ in range(5):
        res = y - X @ c
        scale = 1.4826 * np.median(np.abs(res - np.median(res)))
        if not np.isfinite(scale) or scale < 1.0e-6:
            break
        nrm = np.abs(res) / (1.5 * scale)
        with np.errstate(divide="ignore"):
            huber = np.where(nrm <= 1.0, 1.0, 1.0 / nrm)
        nc, ncond = _solve_once(X, y, w * huber, max_cond)
        if nc is None:
            break
        done = bool(np.all(np.abs(nc - c) <= 1.0e-5 + 1.0e-5 * np.abs(c)))
        c, cond = nc, ncond
        if done:
            break
    res = y - X @ c
    return c, cond, float(np.sqrt(np.mean(res ** 2)))


def run_local_vvp(azimuth, elevation, radial_velocity, az_num, bin_num, fillvalue=-999.0, min_valid_fraction=0.7,
                  min_valid_count=20, min_azimuth_coverage_deg=0.0, min_azimuth_sector_count=2,
                  max_condition_number=1.0e9):
    if az_num % 2 != 1:
        raise ValueError("az_num must be odd")
    if bin_num % 2 != 1:
        raise ValueError("bin_num must be odd")
    vel = np.asarray(radial_velocity, dtype=np.float64)
    if fillvalue is not None:
        vel = np.where(vel == float(fillvalue), np.nan, vel)
    az = np.asarray(azimuth, dtype=np.float64).reshape(-1)
    el = np.asarray(elevation, dtype=np.float64)
    el = np.full(az.shape, float(el)) if el.ndim == 0 else el.reshape(-1)
    naz, nrange = vel.shape
    ha, hb = az_num // 2, bin_num // 2
    d0 = np.sin(np.deg2rad(az)) * np.cos(np.deg2rad(el))
    d1 = np.cos(np.deg2rad(az)) * np.cos(np.deg2rad(el))
    with np.errstate(invalid="ignore"):
        azm = np.where(np.isfinite(az), np.mod(az, 360.0), np.nan)
    wa, wb = window_weights(az_num, bin_num)
    out = {k: np.full((naz, nrange), np.nan) for k in F64_KEYS}
    out.update({k: np.zeros((naz, nrange), dtype=np.int32) for k in I32_KEYS})
    for i in range(naz):
        rows = np.arange(i - ha, i + ha + 1) % naz                   # np.pad(mode="wrap")
        row_azm = azm[rows]
        requested = int(min_azimuth_sector_count)
        required = 0 if requested <= 0 else min(requested, max(_sectors(row_azm), 1))
        X_rows = np.stack((d0[rows], d1[rows], np.ones(az_num)), axis=1)
        for g in range(hb, nrange - hb):
            win = vel[rows, g - hb: g + hb + 1]
            fin = np.isfinite(win)
            count = int(fin.sum())
            frac = float(count / win.size)
            occupied = row_azm[fin.any(axis=1)]
            cov = _coverage(occupied)
            sec = _sectors(occupied)
            out["valid_count"][i, g] = count
            out["valid_fraction"][i, g] = frac
            out["azimuth_coverage_deg"][i, g] = cov
            out["azimuth_sector_count"][i, g] = sec
            if count < int(min_valid_count) or frac < float(min_valid_fraction) or \
                    cov < float(min_azimuth_coverage_deg) or sec < required:
                continue
            rr, cc = np.nonzero(fin & np.isfinite(X_rows[:, 0])[:, None] & np.isfinite(X_rows[:, 1])[:, None])
            if rr.size == 0:
                continue
            c, cond, rmse = _fit(X_rows[rr], win[rr, cc], wa[rr] * wb[cc], float(max_condition_number))
            out["condition_number"][i, g] = cond
            if c is None:
                continue
            u, v = float(c[0]), float(c[1])
            out["u"][i, g], out["v"][i, g], out["radial_offset"][i, g] = u, v, float(c[2])
            out["wind_speed"][i, g] = float(np.hypot(u, v))
            out["wind_direction"][i, g] = float((270.0 - np.rad2deg(np.arctan2(v, u))) % 360.0)
            out["fit_rmse"][i, g] = rmse
    return out
