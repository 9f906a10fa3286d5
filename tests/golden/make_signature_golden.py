"""Record the reference's public signatures on the gridding path (parameter names, order and defaults) so the
host mirror can be checked against them without /root/reference (which does not exist on the GPU box).

    python tests/golden/make_signature_golden.py      # writes tests/golden/reference_signatures.json
"""
import ast
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

WANTED = {
    "pycwr/interp/RadarInterp.py": [
        "run_radar_network_3d", "compose_network_volume", "grid_single_radar_to_latlon_3d",
        "grid_single_radar_cr_to_latlon", "build_latlon_grid", "_compute_observability_mask",
        "_compute_quality_weight_volume", "_build_composite_weight_cache", "select_radar_files",
        "discover_radar_files", "load_network_config", "radar_network_3d_to_netcdf",
    ],
    "pycwr/core/NRadar.py": [
        "add_product_CR_xy", "add_product_CAPPI_xy", "add_product_CAPPI_3d_xy", "add_product_CR_lonlat",
        "add_product_CAPPI_lonlat", "add_product_VIL_xy", "add_product_ET_xy", "add_product_VIL_lonlat",
        "add_product_ET_lonlat", "_grid_reflectivity_volume_xy", "_grid_reflectivity_volume_lonlat",
        "grid_3d_network_xy", "get_vol_data",
    ],
}


def signature(fn):
    args = fn.args
    names = [a.arg for a in args.args]
    defaults = [None] * (len(names) - len(args.defaults)) + [ast.unparse(d) for d in args.defaults]
    has_default = [False] * (len(names) - len(args.defaults)) + [True] * len(args.defaults)
    return [{"name": n, "default": d if h else None, "has_default": h} for n, d, h in zip(names, defaults, has_default)]


def main():
    out = {}
    for rel, wanted in WANTED.items():
        tree = ast.parse(open(os.path.join(REF, rel), encoding="utf-8").read())
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name in wanted:
                out.setdefault(rel, {})[node.name] = signature(node)
    with open(os.path.join(HERE, "reference_signatures.json"), "w", encoding="utf-8") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print({k: sorted(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
