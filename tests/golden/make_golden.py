#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference kernels (oracle/_ref).

Run in the build container only (needs /root/reference to have been compiled by
``oracle/build_ref.py``).  The fixtures are small, committed, and travel to the GPU box where
the reference does not exist; both the CPU oracle and the CUDA path are tested against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from pycwr_b200 import synth  # noqa: E402

FILL = -999.0
BLINDS = ["mask", "nearest_gate", "lowest_sweep", "hybrid", "nearest_valid"]


def small_volume(seed, ne=5, naz=40, ng=60, gate_m=1500.0, ragged=False):
    rng = np.random.default_rng(seed)
    elev = np.sort(rng.uniform(0.4, 18.0, ne))
    va, vr, vv = [], [], []
    for k in range(ne):
        n_az = naz + (k % 3 if ragged else 0)
        n_g = ng - (5 * k if ragged else 0)
        va.append(synth.azimuth_table(rng, n_az))
        vr.append((np.arange(n_g) + 1.0) * gate_m * (1.0 + (0.1 * k if ragged else 0.0)))
        vv.append(synth.stress_moment(rng, n_az, n_g, -5.0, 70.0))
    return va, vr, elev, vv, float(rng.uniform(50.0, 2000.0))


def main():
    build_ref.build()
    R = build_ref.load()
    if R is None:
        raise SystemExit("reference module not available (no /root/reference?)")
    out = {}

    # ---- case A: random scattered targets, every blind mode, uniform + ragged sweeps
    for tag, ragged in (("uniform", False), ("ragged", True)):
        va, vr, fe, vv, h = small_volume(7 if not ragged else 8, ragged=ragged)
        rng = np.random.default_rng(11)
        gx = rng.uniform(-100e3, 100e3, (24, 20))
        gy = rng.uniform(-100e3, 100e3, (24, 20))
        gx[0, :5] = [0.0, 0.0, 300.0, -40e3, 40e3]          # origin, blind zone, diagonals
        gy[0, :5] = [0.0, 500.0, 300.0, -40e3, -40e3]
        lv = np.array([200.0, 1000.0, 2500.0, 6000.0, 12000.0, 20000.0])
        pre = "A_%s_" % tag
        for k in range(len(fe)):
            out[pre + "az%d" % k] = va[k]
            out[pre + "rng%d" % k] = vr[k]
            out[pre + "val%d" % k] = vv[k]
        out[pre + "elev"] = fe
        out[pre + "h"] = np.array(h)
        out[pre + "gx"], out[pre + "gy"], out[pre + "levels"] = gx, gy, lv
        for b in BLINDS:
            out[pre + "cappi3d_" + b] = R.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, None, b)
        out[pre + "cappi3d_re8p1e6"] = R.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, lv, FILL, 8.1e6, "hybrid")
        out[pre + "cr"] = R.get_CR_xy(va, vr, fe, vv, h, gx, gy, FILL)
        out[pre + "ppi2_hybrid"] = R.ppi_to_grid(va[2], vr[2], fe[2], vv[2], h, gx, gy, FILL, None, "hybrid")
        out[pre + "ppi0_mask"] = R.ppi_to_grid(va[0], vr[0], fe[0], vv[0], h, gx, gy, FILL, None, "mask")
        out[pre + "single_bw3"] = R.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL,
                                                 None, "nearest_gate", 3.0)
        out[pre + "single_bw0"] = R.get_CAPPI_xy(va[:1], vr[:1], fe[:1], vv[:1], h, gx, gy, 1500.0, FILL,
                                                 None, "mask", 0.0)

    # ---- case B: three-radar mosaic (RadarGridC.pyx:525-593)
    vols = [small_volume(20 + i) for i in range(3)]
    rng = np.random.default_rng(12)
    gx = rng.uniform(-150e3, 150e3, (16, 18))
    gy = rng.uniform(-150e3, 150e3, (16, 18))
    lv = np.array([500.0, 3000.0, 9000.0])
    rx = np.array([-60e3, 10e3, 70e3])
    ry = np.array([20e3, -50e3, 40e3])
    rh = np.array([v[4] for v in vols])
    for i, v in enumerate(vols):
        for k in range(len(v[2])):
            out["B_r%d_az%d" % (i, k)] = v[0][k]
            out["B_r%d_rng%d" % (i, k)] = v[1][k]
            out["B_r%d_val%d" % (i, k)] = v[3][k]
        out["B_r%d_elev" % i] = v[2]
    out["B_gx"], out["B_gy"], out["B_levels"] = gx, gy, lv
    out["B_rx"], out["B_ry"], out["B_rh"] = rx, ry, rh
    out["B_mosaic"] = R.get_mosaic_CAPPI_3d([v[0] for v in vols], [v[1] for v in vols], [v[2] for v in vols],
                                            [v[3] for v in vols], rx, ry, rh, gx, gy, lv, FILL)
    out["B_mosaic_re"] = R.get_mosaic_CAPPI_3d([v[0] for v in vols], [v[1] for v in vols],
                                               [v[2] for v in vols], [v[3] for v in vols], rx, ry, rh, gx,
                                               gy, lv, FILL, [8.0e6, None, 8.6e6])

    # ---- case C: regular grid through the origin with round-number tables (axis / diagonal ties)
    az = np.arange(0.0, 360.0, 5.0)
    rg = np.arange(1.0, 41.0) * 1000.0
    fe = np.array([0.5, 1.5, 2.5, 4.0])
    rng = np.random.default_rng(13)
    vv = [synth.stress_moment(rng, az.size, rg.size, -5.0, 70.0) for _ in fe]
    ax = np.arange(-40e3, 40e3 + 1, 2000.0)
    gx, gy = np.meshgrid(ax, ax, indexing="ij")
    lv = np.array([500.0, 1500.0, 3000.0])
    out["C_az"], out["C_rng"], out["C_elev"] = az, rg, fe
    for k in range(fe.size):
        out["C_val%d" % k] = vv[k]
    out["C_gx"], out["C_gy"], out["C_levels"] = gx, gy, lv
    out["C_cappi3d_hybrid"] = R.get_CAPPI_3d([az] * 4, [rg] * 4, fe, vv, 100.0, gx, gy, lv, FILL, None, "hybrid")
    out["C_cr"] = R.get_CR_xy([az] * 4, [rg] * 4, fe, vv, 100.0, gx, gy, FILL)

    # ---- case D: scalar helpers (reference test/test_regression_geometry.py:152-243)
    pts = np.array([[12345.0, -5432.0], [1.0, 1.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0], [0.0, 10.0],
                    [10.0, 0.0], [0.0, -10.0], [-10.0, 0.0], [-7.0, -7.0], [83211.5, 40022.25]])
    out["D_pts"] = pts
    out["D_xy_to_azimuth"] = np.array([R.xy_to_azimuth(x, y) for x, y in pts])
    out["D_xye"] = np.array([R.xye_to_antenna(x, y, 1.2, 150.0) for x, y in pts])
    out["D_xye_re"] = np.array([R.xye_to_antenna(x, y, 1.2, 150.0, 8.1e6) for x, y in pts])
    out["D_c2a"] = np.array([R.cartesian_to_antenna(x, y, 3000.0, 150.0) for x, y in pts])
    out["D_a2c"] = np.array([R.antenna_to_cartesian(r, a, e, 150.0)
                             for r, a, e in [(1500.0, 0.0, 1.5), (75000.0, 123.4, 0.5), (230000.0, 359.9, 19.5)]])
    out["D_interp_azimuth"] = np.array([
        R.interp_azimuth(1.5, 1.0, 2.0, 10.0, 20.0, FILL), R.interp_azimuth(1.5, 1.0, 1.0, 10.0, 20.0, FILL),
        R.interp_azimuth(1.5, 1.0, 2.0, FILL, 20.0, FILL), R.interp_azimuth(1.5, 1.0, 2.0, 10.0, FILL, FILL),
        R.interp_azimuth(1.5, 1.0, 2.0, FILL, FILL, FILL)])
    out["D_interp_ppi"] = np.array([
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, 1.0, 2.0, 3.0, 4.0, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, 1.0, 2.0, FILL, 4.0, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, FILL, 2.0, 3.0, 4.0, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, 1.0, FILL, 3.0, FILL, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, FILL, 2.0, FILL, 4.0, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0, FILL, 2.0, 3.0, FILL, FILL),
        R.interp_ppi(1.25, 1400.0, 1.0, 1.0, 1000.0, 2000.0, 1.0, 2.0, 3.0, 4.0, FILL)])

    path = os.path.join(HERE, "gridc_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
