#!/usr/bin/env python
"""Generate tests/golden/network_reference.npz from the UNMODIFIED reference network code.

Build-container only: imports ``pycwr.interp.RadarInterp`` and ``pycwr.core.RadarProduct`` from
``/root/reference`` with the third-party modules that are absent here (xarray, pyproj, easydict,
netCDF4, matplotlib, cartopy ...) replaced by inert stubs — none of them is touched by the array
functions exercised below — and routes the reference's ``get_CAPPI_3d`` / ``get_CR_xy`` to the
compiled reference kernels in ``oracle/_ref`` (what a real installation would import).

    python tests/golden/make_network_golden.py
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from pycwr_b200 import synth  # noqa: E402

FILL = -999.0
MISSING = ["xarray", "pyproj", "easydict", "netCDF4", "matplotlib", "cartopy", "shapefile", "pyart", "xradar",
           "flask", "h5py", "PIL"]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=self.__name__ + "." + name)
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference():
    """(RadarInterp, RadarProduct) of the reference, gridding routed to the compiled kernels."""
    sys.meta_path.insert(0, _Finder())

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            dict.__setitem__(self, k, v)
            object.__setattr__(self, k, v)

        __setitem__ = __setattr__

    ed = _Stub("easydict")
    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed
    sys.path.insert(0, "/root/reference")
    import pycwr.core.RadarProduct as RP
    import pycwr.interp.RadarInterp as RI
    build_ref.build()
    ref = build_ref.load()
    if ref is None:
        raise SystemExit("compiled reference kernels not available")
    RI.get_CAPPI_3d = ref.get_CAPPI_3d
    sys.modules["pycwr.core.RadarGridC"].get_CR_xy = ref.get_CR_xy     # imported inside the function
    return RI, RP


class FakePRD(object):
    """Duck-typed PRD in the style of the reference's own tests (test/test_examples_interp.py:86-96)."""

    def __init__(self, vol):
        self._vol = vol
        self.scan_info = {}

    def get_vol_data(self, field_name="dBZ", fillvalue=FILL, range_mode="aligned", max_range_km=None):
        va, vr, fe, vv, h = self._vol
        self.vol = (va, vr, fe, vv, h, 120.0, 30.0)


def small_network(seed=31, nradar=3):
    """Three radars with ragged native-range sweeps on a common 34 x 30 grid, 6 levels."""
    rng = np.random.default_rng(seed)
    vols, sites, bws = [], [], []
    for i in range(nradar):
        ne = 4 + i
        elev = np.sort(rng.uniform(0.4, 12.0, ne))
        va, vr, vv = [], [], []
        for k in range(ne):
            naz = 48 + (k % 3)
            ng = 70 - 6 * k
            va.append(synth.azimuth_table(rng, naz))
            vr.append((np.arange(ng) + 1.0) * 1500.0)
            vv.append(synth.stress_moment(rng, naz, ng, -5.0, 70.0))
        vols.append((va, vr, elev, vv, float(rng.uniform(50.0, 1500.0))))
        sites.append((float(rng.uniform(-60e3, 60e3)), float(rng.uniform(-60e3, 60e3))))
        bws.append(rng.uniform(0.8, 1.2, ne))
    x = np.linspace(-120e3, 120e3, 34) + 37.0
    y = np.linspace(-110e3, 110e3, 30) - 11.0
    gx, gy = np.meshgrid(x, y, indexing="xy")           # (nlat, nlon) like build_latlon_grid
    levels = np.array([300.0, 1000.0, 2000.0, 3500.0, 6000.0, 9000.0])
    return vols, np.array(sites), bws, np.ascontiguousarray(gx), np.ascontiguousarray(gy), levels


def main():
    RI, RP = import_reference()
    out = {}
    vols, sites, bws, gx, gy, levels = small_network()
    out["N_sites"], out["N_gx"], out["N_gy"], out["N_levels"] = sites, gx, gy, levels
    out["N_nradar"] = np.array(len(vols))
    terms = ["range", "beam", "vertical"]
    grids, quals, masks, crs = [], [], [], []
    for i, (va, vr, fe, vv, h) in enumerate(vols):
        pre = "N_r%d_" % i
        out[pre + "elev"], out[pre + "h"], out[pre + "bw"] = fe, np.array(h), bws[i]
        for k in range(fe.size):
            out[pre + "az%d" % k], out[pre + "rng%d" % k], out[pre + "val%d" % k] = va[k], vr[k], vv[k]
        prd = FakePRD(vols[i])
        for blind in ("mask", "hybrid"):
            g = RI.grid_single_radar_to_latlon_3d(prd, "dBZ", gx, gy, levels, sites[i, 0], sites[i, 1], fillvalue=FILL,
                                                  range_mode="native", max_range_km=460.0, blind_method=blind)
            out[pre + "grid_" + blind] = g
        grids.append(out[pre + "grid_hybrid"])
        q = RI._compute_quality_weight_volume(vr, fe, h, gx, gy, levels, sites[i, 0], sites[i, 1], beam_widths=bws[i],
                                              quality_weight_terms=terms)
        out[pre + "quality"] = q
        out[pre + "quality_range_only"] = RI._compute_quality_weight_volume(
            vr, fe, h, gx, gy, levels, sites[i, 0], sites[i, 1], beam_widths=None, quality_weight_terms=["range"],
            range_decay_m=150000.0)
        quals.append(q)
        m = RI._compute_observability_mask(vr, fe, h, gx, gy, levels, sites[i, 0], sites[i, 1], beam_widths=bws[i])
        out[pre + "mask"] = m
        masks.append(m)
        c = RI.grid_single_radar_cr_to_latlon(prd, "dBZ", gx, gy, sites[i, 0], sites[i, 1], fillvalue=FILL,
                                              range_mode="native", max_range_km=460.0)
        out[pre + "cr"] = c
        crs.append(c)
    coverage = np.sum(np.stack(masks, axis=0), axis=0, dtype=np.int16)      # RadarInterp.py:1938-1945
    out["N_coverage"] = coverage
    out["N_blind"] = (coverage == 0).astype(np.int8)
    for method in ("max", "nearest", "exp_weighted", "quality_weighted"):
        cache = RI._build_composite_weight_cache(sites, gx, gy, method=method, influence_radius_m=300000.0)
        comp, cnt = RI.compose_network_volume(grids, sites, gx, gy, fillvalue=FILL, method=method,
                                              influence_radius_m=300000.0, weight_cache=cache, quality_weights=quals)
        out["N_comp_" + method], out["N_count_" + method] = comp, cnt
    comp, cnt = RI.compose_network_volume(grids, sites, gx, gy, fillvalue=FILL, method="exp_weighted",
                                          influence_radius_m=80000.0)
    out["N_comp_exp_r80"] = comp
    # CR max-merge (RadarInterp.py:2047-2063 restated: those lines live inside run_radar_network_3d)
    stack = np.stack(crs, axis=0)
    ok = np.isfinite(stack) & (stack != FILL)
    out["N_CR"] = np.where(np.any(ok, axis=0), np.max(np.where(ok, stack, -np.inf), axis=0), np.nan)
    # column products on the default composite
    comp = out["N_comp_quality_weighted"]
    out["N_derive_cr"] = RP.derive_cr(comp, fillvalue=FILL)
    out["N_derive_vil"] = RP.derive_vil(comp, levels, fillvalue=FILL)
    et, topped = RP.derive_et(comp, levels, fillvalue=FILL, return_topped=True)
    out["N_derive_et"], out["N_derive_topped"] = et, topped
    out["N_derive_et_35"] = RP.derive_et(comp, levels, fillvalue=FILL, threshold_dbz=35.0)
    # the NumPy-twin geometry on scattered points (pycwr/core/transforms.py:201-222)
    from pycwr.core.transforms import cartesian_xyz_to_antenna
    rng = np.random.default_rng(5)
    px, py = rng.uniform(-3e5, 3e5, 4096), rng.uniform(-3e5, 3e5, 4096)
    az, r, el = cartesian_xyz_to_antenna(px, py, 2500.0, 321.0)
    out["T_x"], out["T_y"], out["T_az"], out["T_r"], out["T_el"] = px, py, az, r, el
    # single-sweep radar: observability + quality (RadarInterp.py:566-578, 653-659)
    va, vr, fe, vv, h = vols[0]
    out["S_mask"] = RI._compute_observability_mask(vr[:1], fe[:1], h, gx, gy, levels, 0.0, 0.0, beam_widths=np.array([2.5]))
    out["S_quality"] = RI._compute_quality_weight_volume(vr[:1], fe[:1], h, gx, gy, levels, 0.0, 0.0,
                                                         beam_widths=np.array([2.5]), quality_weight_terms=terms)
    path = os.path.join(HERE, "network_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
