#!/usr/bin/env python
"""Generate tests/golden/packer_reference.npz from the UNMODIFIED reference packer.

Build-container only: imports ``pycwr.core.NRadar`` from ``/root/reference`` (absent third-party modules replaced by
the inert stubs of ``make_network_golden.py``) and calls ``PRD.get_vol_data`` / ``PRD._extract_sweep_volume``
(``pycwr/core/NRadar.py:655-712,1931-1961``) as the reference wrote them, on a duck-typed volume that carries
exactly what those two methods read: ``fields`` (aligned sweeps), ``extended_fields`` (native-range reflectivity),
``scan_info`` and ``_fixed_angle_sort_index``.  Cases: aligned / native range mode, ``max_range_km`` = None, a
clip inside the table and the keep-one-gate clip (:655-663), sweeps stored out of elevation order, rays in scan
order with a random start, NaN for missing, ragged native tables.

    python tests/golden/make_packer_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_network_golden as mng  # noqa: E402

CASES = [("aligned", None), ("aligned", 21.0), ("aligned", 0.1), ("native", None), ("native", 33.0), ("native", 0.05)]


class _V(object):
    def __init__(self, values):
        self.values = values

    def argsort(self):
        return _V(np.argsort(self.values))


def raw_volume(seed=23):
    """Per sweep (scan order): azimuth, aligned range + dBZ / ZDR, native range + dBZ (longer, finer)."""
    rng = np.random.default_rng(seed)
    fix = np.array([2.4, 0.5, 9.9, 1.45, 4.3, 0.5004])        # not in elevation order, two nearly equal
    sweeps = []
    for s in range(fix.size):
        naz = int(rng.integers(50, 80))
        az = np.mod(rng.uniform(0, 360) + np.arange(naz) * (360.0 / naz) + rng.uniform(-0.3, 0.3, naz), 360.0)
        ng = 90 + 5 * s
        rg = 300.0 * (np.arange(ng) + 1.0)
        ngn = 150 - 7 * s
        rgn = 250.0 * (np.arange(ngn) + 1.0)

        def moment(n, lo, hi):
            v = rng.uniform(lo, hi, (naz, n))
            v[rng.random((naz, n)) < 0.25] = np.nan
            v[rng.integers(0, naz)] = np.nan
            return v.astype(np.float32).astype(np.float64)       # reader data is float32
        sweeps.append({"az": az, "rg": rg, "dBZ": moment(ng, -5, 70), "ZDR": moment(ng, -2, 5),
                       "rgn": rgn, "dBZn": moment(ngn, -5, 70)})
    return fix, sweeps


def main():
    mng.import_reference()
    import pycwr.core.NRadar as NR
    PRD = NR.PRD
    fix, sweeps = raw_volume()
    fake = types.SimpleNamespace()
    fake.fields = [{"azimuth": _V(s["az"]), "range": _V(s["rg"]), "dBZ": _V(s["dBZ"]), "ZDR": _V(s["ZDR"])}
                   for s in sweeps]
    fake.extended_fields = {"dBZ": {i: {"azimuth": s["az"], "range": s["rgn"], "data": s["dBZn"]}
                                    for i, s in enumerate(sweeps)}}
    fake.scan_info = {"fixed_angle": _V(fix), "altitude": _V(np.float64(231.5)),
                      "longitude": _V(np.float64(118.25)), "latitude": _V(np.float64(31.75))}
    fake._fixed_angle_sort_index = fake.scan_info["fixed_angle"].argsort().values      # NRadar.py:218
    fake._vol_cache = {}
    fake._field_prefers_native_range = PRD._field_prefers_native_range
    fake._validate_range_mode = PRD._validate_range_mode
    fake._clip_range_window = PRD._clip_range_window
    fake._resolve_field_range_mode = types.MethodType(PRD._resolve_field_range_mode, fake)
    fake._extract_sweep_volume = types.MethodType(PRD._extract_sweep_volume, fake)
    get_vol_data = types.MethodType(PRD.get_vol_data, fake)

    out = {"fix": fix, "nsweep": np.int64(fix.size)}
    for i, s in enumerate(sweeps):
        for k, v in s.items():
            out["raw%d_%s" % (i, k)] = v
    for ci, (mode, mr) in enumerate(CASES):
        for field in ("dBZ", "ZDR"):
            get_vol_data(field_name=field, fillvalue=-999.0, range_mode=mode, max_range_km=mr)
            va, vr, fe, vv, h, lon, lat = fake.vol
            pre = "c%d_%s_" % (ci, field)
            out[pre + "fix"] = np.asarray(fe)
            out[pre + "site"] = np.array([h, lon, lat])
            for k in range(len(va)):
                out[pre + "az%d" % k], out[pre + "rg%d" % k], out[pre + "val%d" % k] = va[k], vr[k], vv[k]
    # the default range mode: native for dBZ, aligned for the rest (NRadar.py:256-261)
    get_vol_data(field_name="dBZ")
    out["default_dBZ_ng"] = np.array([r.size for r in fake.vol[1]])
    get_vol_data(field_name="ZDR")
    out["default_ZDR_ng"] = np.array([r.size for r in fake.vol[1]])
    path = os.path.join(HERE, "packer_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
