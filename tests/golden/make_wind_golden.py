#!/usr/bin/env python
"""Generate tests/golden/wind_reference.npz from the UNMODIFIED reference wind-gridding functions
(``pycwr/retrieve/WindField.py``: ``_grid_vvp_retrieval_xy``, ``_reconstruct_wind_volume_chunk``),
imported from /root/reference with the absent third-party modules stubbed (scipy's cKDTree, which
the reference uses for the neighbour search, is installed here).  Build container only.

    python tests/golden/make_wind_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_network_golden import import_reference  # noqa: E402  (installs the stub finder)


def synthetic_retrievals(seed=41, nsweep=5):
    """Per-sweep VVP analyses on a coarse polar lattice (every third ray / gate like the reference's
    local VVP windows), with missing samples, tilted heights and noisy diagnostics."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(nsweep):
        elev = np.deg2rad(0.5 + 1.3 * s)
        az = np.deg2rad(np.arange(0.0, 360.0, 9.0) + rng.uniform(0, 3))
        rg = np.arange(3000.0, 60000.0, 2500.0)
        A, R = np.meshgrid(az, rg, indexing="ij")
        x, y = R * np.cos(elev) * np.sin(A), R * np.cos(elev) * np.cos(A)
        z = 120.0 + R * np.sin(elev) + R * R / (2 * 8.5e6)
        u = 12.0 + 0.0004 * z + rng.normal(0, 0.8, x.shape)
        v = -5.0 + 0.0002 * z + rng.normal(0, 0.8, x.shape)
        rmse = np.abs(rng.normal(1.5, 1.0, x.shape))
        cnt = rng.integers(0, 90, x.shape).astype(np.float64)
        frac = rng.uniform(-0.1, 1.2, x.shape)
        cond = rng.uniform(1.0, 40.0, x.shape)
        sec = rng.integers(1, 9, x.shape).astype(np.float64)
        for a in (u, v, rmse):
            a[rng.random(x.shape) < 0.12] = np.nan
        frac[rng.random(x.shape) < 0.03] = np.nan
        cond[rng.random(x.shape) < 0.05] = np.nan          # not part of the validity mask
        out.append({"x": x, "y": y, "z": z, "u": u, "v": v, "fit_rmse": rmse, "valid_count": cnt,
                    "valid_fraction": frac, "condition_number": cond, "azimuth_sector_count": sec})
    return out


def main():
    import_reference()
    import pycwr.retrieve.WindField as W
    out = {}
    rets = synthetic_retrievals()
    tx, ty = np.meshgrid(np.linspace(-52e3, 52e3, 27), np.linspace(-50e3, 50e3, 23), indexing="ij")
    out["tx"], out["ty"] = tx, ty
    out["nsweep"] = np.array(len(rets))
    layers = []
    for s, r in enumerate(rets):
        for k, a in r.items():
            out["s%d_%s" % (s, k)] = a
        g = W._grid_vvp_retrieval_xy(r, tx, ty, horizontal_radius_m=4000.0, horizontal_min_neighbors=3)
        layers.append(g)
        for k, a in g.items():
            out["g%d_%s" % (s, k)] = a
    # other parameter sets on sweep 0: no expansion, explicit max radius, non-default power
    for tag, kw in (("noexp", dict(horizontal_radius_m=5000.0, max_horizontal_radius_m=5000.0)),
                    ("wide", dict(horizontal_radius_m=2500.0, max_horizontal_radius_m=9000.0, horizontal_min_neighbors=5)),
                    ("p15", dict(horizontal_radius_m=6000.0, distance_power=1.5))):
        g = W._grid_vvp_retrieval_xy(rets[0], tx, ty, **kw)
        for k, a in g.items():
            out["g_%s_%s" % (tag, k)] = a
    stack = lambda key: np.stack([np.asarray(l[key], dtype=np.float64) for l in layers], axis=0)
    levels = np.arange(250.0, 6001.0, 250.0)
    out["levels"] = levels
    for tag, tol, gap in (("a", 250.0, 1500.0), ("b", 120.0, 700.0)):
        _, _, vol = W._reconstruct_wind_volume_chunk(
            0, tx.shape[0], stack("z"), stack("u"), stack("v"), stack("fit_rmse"), stack("valid_count"),
            stack("neighbor_count"), stack("expanded_radius"), levels, tol, gap)
        for k, a in vol.items():
            out["v%s_%s" % (tag, k)] = a
    path = os.path.join(HERE, "wind_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
