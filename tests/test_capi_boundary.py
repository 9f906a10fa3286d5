"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the Python module exports the reference's eleven names with its signatures, the
scalar helpers agree with the oracle / fixtures, and argument errors match the reference's."""
import inspect
import os
import re

import numpy as np
import pytest

from conftest import FILL, ROOT
from oracle import oracle as O


def test_library_exports_every_declared_symbol():
    from pycwr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pycwr_b200.h")).read()
    declared = set(re.findall(r"\b(pcg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.pcg_version()


def test_module_exports_reference_names_and_signatures():
    from pycwr_b200 import RadarGridC as G
    names = ["antenna_to_cartesian", "cartesian_to_antenna", "get_CAPPI_3d", "get_CAPPI_xy", "get_CR_xy",
             "get_mosaic_CAPPI_3d", "interp_azimuth", "interp_ppi", "ppi_to_grid", "xy_to_azimuth",
             "xye_to_antenna"]        # pycwr/core/RadarGridC.py:25-37
    assert sorted(G.__all__) == sorted(names)
    def params(fn):
        # the reference's positional parameters; extensions are `return_index` and keyword-only options
        return [n for n, p in inspect.signature(fn).parameters.items()
                if n != "return_index" and p.kind != inspect.Parameter.KEYWORD_ONLY]
    assert params(G.get_CR_xy) == ["vol_azimuth", "vol_range", "fix_elevation", "vol_value", "radar_height",
                                   "GridX", "GridY", "fillvalue", "effective_earth_radius"]
    assert params(G.get_CAPPI_xy) == ["vol_azimuth", "vol_range", "fix_elevation", "vol_value", "radar_height",
                                      "GridX", "GridY", "level_height", "fillvalue", "effective_earth_radius",
                                      "blind_method", "beam_width_deg"]
    assert params(G.get_CAPPI_3d) == ["vol_azimuth", "vol_range", "fix_elevation", "vol_value", "radar_height",
                                      "GridX", "GridY", "level_heights", "fillvalue", "effective_earth_radius",
                                      "blind_method"]
    assert params(G.get_mosaic_CAPPI_3d) == ["vol_azimuth", "vol_range", "fix_elevation", "vol_value", "radar_x",
                                             "radar_y", "radar_height", "GridX", "GridY", "level_heights",
                                             "fillvalue", "effective_earth_radius"]
    assert params(G.ppi_to_grid) == ["azimuth", "ranges", "elevation", "mat_ppi", "radar_height", "GridX",
                                     "GridY", "fillvalue", "effective_earth_radius", "blind_method"]
    sig = inspect.signature(G.get_CAPPI_xy).parameters
    assert sig["blind_method"].default == "mask" and sig["beam_width_deg"].default == 1.0


def test_scalar_helpers_match_fixtures_bit_exact(golden):
    from pycwr_b200 import RadarGridC as G
    pts = golden["D_pts"]
    assert np.array_equal([G.xy_to_azimuth(x, y) for x, y in pts], golden["D_xy_to_azimuth"])
    assert np.array_equal([G.xye_to_antenna(x, y, 1.2, 150.0) for x, y in pts], golden["D_xye"])
    assert np.array_equal([G.xye_to_antenna(x, y, 1.2, 150.0, 8.1e6) for x, y in pts], golden["D_xye_re"])
    assert np.array_equal([G.cartesian_to_antenna(x, y, 3000.0, 150.0) for x, y in pts], golden["D_c2a"])
    got = [G.antenna_to_cartesian(r, a, e, 150.0)
           for r, a, e in [(1500.0, 0.0, 1.5), (75000.0, 123.4, 0.5), (230000.0, 359.9, 19.5)]]
    assert np.array_equal(got, golden["D_a2c"])
    got = [G.interp_azimuth(1.5, 1.0, 2.0, 10.0, 20.0, FILL), G.interp_azimuth(1.5, 1.0, 1.0, 10.0, 20.0, FILL),
           G.interp_azimuth(1.5, 1.0, 2.0, FILL, 20.0, FILL), G.interp_azimuth(1.5, 1.0, 2.0, 10.0, FILL, FILL),
           G.interp_azimuth(1.5, 1.0, 2.0, FILL, FILL, FILL)]
    assert np.array_equal(got, golden["D_interp_azimuth"])
    a = (1.25, 1400.0, 1.0, 2.0, 1000.0, 2000.0)
    got = [G.interp_ppi(*a, 1.0, 2.0, 3.0, 4.0, FILL), G.interp_ppi(*a, 1.0, 2.0, FILL, 4.0, FILL),
           G.interp_ppi(*a, FILL, 2.0, 3.0, 4.0, FILL), G.interp_ppi(*a, 1.0, FILL, 3.0, FILL, FILL),
           G.interp_ppi(*a, FILL, 2.0, FILL, 4.0, FILL), G.interp_ppi(*a, FILL, 2.0, 3.0, FILL, FILL),
           G.interp_ppi(1.25, 1400.0, 1.0, 1.0, 1000.0, 2000.0, 1.0, 2.0, 3.0, 4.0, FILL)]
    assert np.array_equal(got, golden["D_interp_ppi"])


def test_scalar_helpers_match_oracle_random():
    from pycwr_b200 import RadarGridC as G
    rng = np.random.default_rng(20260319)
    for _ in range(500):
        x, y = rng.uniform(-3e5, 3e5, 2)
        z, h, e = rng.uniform(0, 2e4), rng.uniform(0, 3e3), rng.uniform(0.2, 30)
        assert G.cartesian_to_antenna(x, y, z, h) == O.cartesian_to_antenna(x, y, z, h)
        assert G.xye_to_antenna(x, y, e, h) == O.xye_to_antenna(x, y, e, h)
        assert G.xy_to_azimuth(x, y) == O.xy_to_azimuth(x, y)
        r, a = rng.uniform(100, 4e5), rng.uniform(0, 360)
        assert G.antenna_to_cartesian(r, a, e, h) == O.antenna_to_cartesian(r, a, e, h)


def test_argument_errors_match_reference():
    from pycwr_b200 import RadarGridC as G
    az = np.array([0.0, 90.0, 180.0, 270.0])
    rg = np.array([1000.0, 2000.0])
    val = np.full((4, 2), 10.0)
    g = np.zeros((2, 2))
    with pytest.raises(ValueError, match="effective_earth_radius must be positive"):
        G.get_CR_xy([az], [rg], np.array([1.0]), [val], 0.0, g, g, FILL, -1.0)
    with pytest.raises(ValueError, match="blind_method must be one of"):
        G.get_CAPPI_3d([az], [rg], np.array([1.0]), [val], 0.0, g, g, np.array([0.0]), FILL, None, "bogus")
    with pytest.raises(ValueError, match="dtype mismatch"):
        G.get_CAPPI_3d([az], [rg], np.array([1.0]), [val], 0.0, g.astype(np.float32), g, np.array([0.0]), FILL)
    with pytest.raises(ValueError, match="must have the same length"):
        G.get_mosaic_CAPPI_3d([[az]], [[rg]], [np.array([1.0])], [[val]], np.zeros(2), np.zeros(2),
                              np.zeros(2), g, g, np.array([0.0]), FILL)
    with pytest.raises(ValueError, match="sequence matching the radar count"):
        G.get_mosaic_CAPPI_3d([[az]], [[rg]], [np.array([1.0])], [[val]], np.zeros(1), np.zeros(1),
                              np.zeros(1), g, g, np.array([0.0]), FILL, [1.0, 2.0])
    with pytest.raises(ValueError, match="effective_earth_radius must be positive"):
        G.xye_to_antenna(1.0, 1.0, 1.0, 0.0, 0.0)


def test_product_path_has_no_oracle_or_cpu_fallback():
    """The shipped package never touches oracle/ (grep the sources)."""
    pkg = os.path.join(ROOT, "pycwr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle/|oracle\.|libgrid_oracle|_ref/|build_ref", text), f
                assert "/root/reference" not in text, f


def test_committed_traffic_captures_name_real_workloads():
    """`roofline.traffic` of the bench line is read from profiles/traffic.json by workload name: every key must be a
    workload `synth` can build (or a kernel name kept from round 1), and the headline workloads must have a capture
    whose DRAM bytes are read + write."""
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    table = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    src = open(os.path.join(root, "pycwr_b200", "synth.py")).read()
    names = set(re.findall(r'"name": "(C[0-9][A-Za-z0-9_]+)"', src)) | {"C4_network_32r_3000x3000x30"}
    for key, rec in table.items():
        assert key in names or key.startswith("cappi3d_"), key
        assert rec["bytes"] == rec["read"] + rec["write"]
        assert os.path.basename(rec["report"]) == rec["report"] and rec["report"].endswith(".ncu-rep")
    for must in ("C3_3D_lonlat_1201x1201x40", "C4_network_32r_3000x3000x30", "C5_PA_1001x1001x60"):
        assert must in table, must
