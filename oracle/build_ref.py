#!/usr/bin/env python
"""Build the UNMODIFIED reference gridding kernels into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  This compiles ``/root/reference/pycwr/core/RadarGridC.pyx``
from where it lies (the same recipe as the reference's ``setup.py:85-90``: a plain
``Extension`` + ``cythonize(language_level=3)``), writing every generated file
(the ``.c``, the objects and the ``.so``) below ``oracle/_ref/`` which is
git-ignored.  No reference source is copied into the tracked tree.

The resulting top-level module ``RadarGridC`` (imported from ``oracle/_ref``)
is used (a) to pin the C restatement ``oracle/grid_oracle.c`` bit-for-bit,
(b) to generate the fixtures under ``tests/golden/`` and (c) as the
``cpu_baseline.kind == "reference"`` arm of ``bench.py``.  The product path
(``pycwr_b200``) never imports it.

``/root/reference`` only exists in the build container; on the GPU box the
prebuilt ``.so`` travels with the snapshot and this script is a no-op.
"""
import os
import shutil
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/pycwr/core/RadarGridC.pyx"
OUT = os.path.join(HERE, "_ref")


def ref_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(OUT, "RadarGridC" + suffix)


def build(force=False, verbose=False):
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(REF_PYX):
        return so
    os.makedirs(OUT, exist_ok=True)
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution

    build_dir = os.path.join(OUT, "build")
    ext = Extension("RadarGridC", sources=[REF_PYX], include_dirs=[numpy.get_include()])
    exts = cythonize([ext], language_level=3, build_dir=build_dir, quiet=not verbose)
    dist = Distribution({"ext_modules": exts, "script_name": "build_ref"})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = OUT
    cmd.build_temp = build_dir
    cmd.inplace = 0
    cmd.ensure_finalized()
    old = os.dup(1)
    try:
        if not verbose:
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
        cmd.run()
    finally:
        os.dup2(old, 1)
    shutil.rmtree(build_dir, ignore_errors=True)
    return so if os.path.exists(so) else None


def load():
    """Import the compiled reference module (None if it is not available)."""
    so = ref_so_path()
    if not os.path.exists(so):
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location("RadarGridC", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print("reference oracle:", path)
