#!/usr/bin/env python
"""Generate tests/golden/vvp_reference.npz from the UNMODIFIED reference ``_run_local_vvp``
(``pycwr/retrieve/WindField.py:1843-1941``), imported from /root/reference with the absent third-party
modules stubbed.  Build container only.

    python tests/golden/make_vvp_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_network_golden import import_reference  # noqa: E402


def synthetic_sweep(seed, naz, nrange, elev, fill=-999.0, outliers=0.04, missing=0.15, sector=None):
    """Radial velocity of a sheared wind seen on a cone, with noise, gross outliers (so the Huber loop has
    work to do), missing patches and optionally an empty azimuth sector."""
    rng = np.random.default_rng(seed)
    az = np.sort(np.mod(np.arange(naz) * (360.0 / naz) + rng.uniform(-0.3, 0.3, naz), 360.0))
    el = np.full(naz, elev) + rng.normal(0.0, 0.02, naz)
    rg = np.arange(nrange)[None, :]
    u = 9.0 + 0.05 * rg + 2.0 * np.sin(np.deg2rad(az))[:, None]
    v = -4.0 + 0.08 * rg
    a, e = np.deg2rad(az)[:, None], np.deg2rad(el)[:, None]
    vr = (u * np.sin(a) + v * np.cos(a)) * np.cos(e) + 0.3 + rng.normal(0.0, 0.6, (naz, nrange))
    bad = rng.random(vr.shape) < outliers
    vr[bad] += rng.choice([-1.0, 1.0], bad.sum()) * rng.uniform(8.0, 25.0, bad.sum())
    vr[rng.random(vr.shape) < missing] = fill
    vr[:, nrange // 2] = fill                                   # an empty gate ring
    if sector is not None:
        vr[(az >= sector[0]) & (az < sector[1]), :] = np.nan    # NaN is missing as well
    return az, el, vr


CASES = [
    # name, seed, naz, nrange, elev, az_num, bin_num, kwargs
    ("a", 1, 72, 30, 1.5, 19, 5, {}),
    ("b", 2, 90, 24, 3.4, 31, 3, {"min_valid_fraction": 0.5, "min_valid_count": 10, "min_azimuth_coverage_deg": 60.0}),
    ("c", 3, 48, 20, 0.5, 9, 9, {"min_azimuth_sector_count": 1, "max_condition_number": 2200.0}),
    ("d", 4, 40, 18, 6.0, 41, 1, {"min_valid_fraction": 0.3, "min_azimuth_sector_count": 3}),   # window wraps fully
    ("e", 5, 60, 16, 2.4, 15, 5, {"min_valid_fraction": 0.4, "min_valid_count": 5}),
]
SECTORS = {"e": (100.0, 220.0)}


def main():
    import_reference()
    import pycwr.retrieve.WindField as W
    out = {}
    for name, seed, naz, nrange, elev, az_num, bin_num, kw in CASES:
        az, el, vr = synthetic_sweep(seed, naz, nrange, elev, sector=SECTORS.get(name))
        res = W._run_local_vvp(az, el if name != "c" else elev, vr, az_num, bin_num, **kw)
        out["%s_az" % name], out["%s_el" % name], out["%s_vr" % name] = az, el if name != "c" else np.float64(elev), vr
        for k, v in res.items():
            out["%s_%s" % (name, k)] = v
        print(name, "fits:", int(np.isfinite(res["u"]).sum()), "of", res["u"].size,
              "cond-only:", int((np.isfinite(res["condition_number"]) & ~np.isfinite(res["u"])).sum()))
    path = os.path.join(HERE, "vvp_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
