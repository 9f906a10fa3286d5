#!/usr/bin/env python
"""Generate tests/golden/section_reference.npz from the UNMODIFIED reference section code.

Build-container only: imports ``pycwr.core.NRadar`` and ``pycwr.core.transforms`` from ``/root/reference``
(absent third-party modules replaced by the inert stubs of ``make_network_golden.py`` — none is touched by
the array functions used here) and records the outputs of ``PRD._build_section_line``,
``cartesian_to_antenna_cwr`` and ``PRD._interpolate_ppi_strip_arrays`` on a small synthetic volume with
missing patches, a duplicated ray, ragged sweeps, rays at 0 and 359.999 deg and targets outside [0, 360).

    python tests/golden/make_section_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_network_golden as mng  # noqa: E402


def synthetic_sweeps(seed=11):
    rng = np.random.default_rng(seed)
    elev = np.array([0.5, 1.5, 2.4, 4.3, 9.9])
    shapes = [(92, 140), (90, 140), (88, 120), (61, 90), (45, 60)]
    sweeps = []
    for s, (naz, ng) in enumerate(shapes):
        # ascending within [0, 360): on a table that is unsorted after the reference's mod 360 its
        # np.searchsorted result depends on the previous key (numpy's binary search keeps state), i.e.
        # the reference itself is undefined there
        az = np.sort(np.mod(np.arange(naz) * (360.0 / naz) + rng.uniform(-0.8, 0.8, naz) + 0.37 * s + 1.0, 360.0))
        if s == 0:
            az[10] = az[9]                      # duplicated ray: the reference keeps the first
        if s == 1:
            az[-1] = 359.999                    # wrap gap of 1e-3 deg + first ray
        if s == 2:
            az[0] = 0.0                         # a ray exactly at north
        gates = 500.0 + 300.0 * np.arange(ng) + (5.0 * s)
        a, g = np.meshgrid(np.deg2rad(az), gates, indexing="ij")
        v = 25.0 + 20.0 * np.sin(3 * a + s) * np.cos(g / 9000.0) + rng.normal(0.0, 1.0, a.shape)
        v[rng.random(a.shape) < 0.12] = np.nan
        v[(g > 20000.0) & (g < 24000.0) & (np.sin(a) > 0.3)] = np.nan      # a missing block
        v[5:8, :] = np.nan                                                   # missing rays
        v[:, 30] = np.nan                                                    # a missing gate ring
        sweeps.append((az, gates, v))
    return elev, sweeps


def main():
    mng.import_reference()
    import pycwr.core.NRadar as NR
    from pycwr.core.transforms import cartesian_to_antenna_cwr
    PRD = NR.PRD
    elev, sweeps = synthetic_sweeps()
    h = 312.5
    out = {"elev": elev, "h": np.float64(h)}
    for s, (az, rg, v) in enumerate(sweeps):
        out["az_%d" % s] = az
        out["rg_%d" % s] = rg
        out["v_%d" % s] = v
    lines = [((-31000.0, -12000.0), (35500.0, 27100.0), 300.0),
             ((0.0, 0.0), (0.0, 41000.0), 250.0),              # starts over the radar, due north
             ((-20000.0, 3.0), (20000.0, 3.0), 173.0),          # passes 3 m from the radar
             ((60000.0, 60000.0), (10000.0, -5000.0), 1000.0)]  # mostly beyond the last gate
    for li, (p0, p1, sp) in enumerate(lines):
        x, y, d = PRD._build_section_line(p0, p1, sp)
        out["line_%d" % li] = np.array([p0[0], p0[1], p1[0], p1[1], sp])
        out["x_%d" % li], out["y_%d" % li], out["d_%d" % li] = x, y, d
        for re_name, re in (("def", None), ("alt", 7.1e6)):
            vals, azs, rgs, zs = [], [], [], []
            for s, (az, rg, v) in enumerate(sweeps):
                taz, trg, z = cartesian_to_antenna_cwr(x, y, float(elev[s]), h, effective_earth_radius=re)
                vals.append(PRD._interpolate_ppi_strip_arrays(az, rg, v, taz, trg))
                azs.append(taz); rgs.append(trg); zs.append(z)
            out["sec_%s_val_%d" % (re_name, li)] = np.stack(vals)
            out["sec_%s_az_%d" % (re_name, li)] = np.stack(azs)
            out["sec_%s_rg_%d" % (re_name, li)] = np.stack(rgs)
            out["sec_%s_z_%d" % (re_name, li)] = np.stack(zs)
    for ai, azimuth in enumerate([0.0, 47.3, 181.25, 359.9, 361.0, -12.5]):
        rows = [PRD._interpolate_ppi_strip_arrays(az, rg, v, np.full(rg.shape, azimuth), rg) for az, rg, v in sweeps]
        out["rhi_az_%d" % ai] = np.float64(azimuth)
        out["rhi_val_%d" % ai] = PRD._pad_section_rows(rows)
    # explicit random targets incl. NaN / inf / out-of-range / exact table hits
    rng = np.random.default_rng(5)
    for s, (az, rg, v) in enumerate(sweeps):
        taz = rng.uniform(-30.0, 400.0, 600)
        trg = rng.uniform(0.0, rg[-1] * 1.05, 600)
        taz[:40] = np.mod(az[rng.integers(0, az.size, 40)], 360.0)
        trg[40:80] = rg[rng.integers(0, rg.size, 40)]
        taz[80] = np.nan; trg[81] = np.nan; taz[82] = np.inf; trg[83] = np.inf
        trg[84] = rg[0]; trg[85] = rg[-1]
        out["t_az_%d" % s], out["t_rg_%d" % s] = taz, trg
        out["t_val_%d" % s] = PRD._interpolate_ppi_strip_arrays(az, rg, v, taz, trg)
    path = os.path.join(HERE, "section_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
