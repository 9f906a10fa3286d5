"""ctypes binding of ``libpycwr_b200.so`` (C ABI declared in ``include/pycwr_b200.h``).

The library is the product: if it is missing or no CUDA device is usable, the compute entry
points fail loudly — there is no CPU fallback.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYCWR_B200_LIB selects another build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("PYCWR_B200_LIB") or os.path.join(_HERE, "libpycwr_b200.so")

PCG_OK, PCG_ERR_ARG, PCG_ERR_CUDA, PCG_ERR_NOMEM = 0, -1, -2, -3
PCG_WARN_NONFINITE = 1
IDX_STRIDE = 10
VALUE_F64, VALUE_F32 = 0, 1

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_pp = ctypes.POINTER(c_double_p)
c_double_ppp = ctypes.POINTER(c_double_pp)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
_vp, _d, _i = ctypes.c_void_p, ctypes.c_double, ctypes.c_int

# symbol -> (restype, argtypes); mirrors include/pycwr_b200.h one to one
SIGNATURES = {
    "pcg_last_error": (ctypes.c_char_p, []),
    "pcg_version": (ctypes.c_char_p, []),
    "pcg_device_count": (_i, [c_int_p]),
    "pcg_set_device": (_i, [_i]),
    "pcg_host_alloc": (_i, [ctypes.c_size_t, c_void_pp]),
    "pcg_host_free": (_i, [_vp]),
    "pcg_launch_count": (ctypes.c_int64, []),
    "pcg_volume_create": (_i, [_i, c_int_p, c_int_p, c_double_pp, c_double_pp, _i, c_double_ppp,
                               c_double_p, _d, _i, c_void_pp]),
    "pcg_volume_destroy": (_i, [_vp]),
    "pcg_volume_create_raw": (_i, [_i, c_int_p, c_int_p, c_double_pp, c_double_pp, _i, c_double_ppp, _vp,
                                   ctypes.POINTER(c_int_p), c_double_p, _d, _i, ctypes.POINTER(ctypes.c_void_p)]),
    "pcg_volume_info": (_i, [_vp, c_int_p, c_int_p, c_int_p, ctypes.POINTER(ctypes.c_size_t)]),
    "pcg_get_cappi3d": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, c_double_p, _i, _d, _i, _d, _vp, _vp]),
    "pcg_get_cr": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, _d, _vp]),
    "pcg_get_cappi3d_ex": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, c_double_p, _i, _d, _i, _d, _d, _vp, _vp]),
    "pcg_get_cr_ex": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, _d, _d, _vp]),
    "pcg_ppi_to_grid": (_i, [_vp, _i, _d, _d, _d, _vp, _vp, _i, _i, _d, _i, _vp, _vp]),
    "pcg_get_mosaic_cappi3d": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p,
                                    _vp, _vp, _i, _i, c_double_p, _i, _d, _vp]),
    "pcg_get_cappi3d_dev": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, c_double_p, _i, _d, _i, _d, _d, _d,
                                 _i, _vp, _vp, _vp]),
    "pcg_get_cr_dev": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, _d, _d, _d, _vp, _vp]),
    "pcg_volume_value_range": (_i, [_vp, _i, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_uint64)]),
    "pcg_network_compose": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_pp, _vp,
                                 _vp, _vp, _i, _i, c_double_p, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp,
                                 ctypes.POINTER(ctypes.c_uint64)]),
    "pcg_network_compose_ex": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_pp, _vp,
                                    _vp, _vp, _i, _i, c_double_p, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp,
                                    ctypes.POINTER(ctypes.c_uint64), c_double_p, _vp, _vp, _vp]),
    "pcg_network_compose_rows": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_pp, _vp,
                                      _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _i, _i, c_double_p, _i, _d, _d,
                                      _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_uint64), c_double_p, _vp, _vp,
                                      _vp]),
    "pcg_peer_alloc": (_i, [ctypes.c_size_t, c_void_pp, _vp]),
    "pcg_peer_open": (_i, [_vp, c_void_pp]),
    "pcg_peer_close": (_i, [_vp]),
    "pcg_peer_free": (_i, [_vp]),
    "pcg_network_compose_gather_dev": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p,
                                            c_double_pp, _vp, _vp, _vp, _i, _i, c_double_p, _i, _d, _vp, _vp, _vp,
                                            _vp, _vp, _i, c_void_pp, ctypes.c_longlong, ctypes.c_longlong, _vp]),
    "pcg_network_compose_dev": (_i, [_i, c_void_pp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_pp,
                                     _vp, _vp, _vp, _i, _i, c_double_p, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcg_network_gates": (_i, [_i, c_double_p, c_double_p, c_double_p, c_double_p, _d, _d, _d, _d, _vp, _vp, _i, _i,
                               c_double_p, _i, _i, _d, _d, _d, _vp, _vp]),
    "pcg_compose_volumes": (_i, [_i, c_double_pp, c_double_pp, _vp, _i, _i, _i, _i, _d, _vp, _vp]),
    "pcg_column_products": (_i, [_vp, c_double_p, _i, _i, _i, _d, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "pcg_get_column_products": (_i, [_vp, _d, _d, _vp, _vp, _i, _i, c_double_p, _i, _d, _i, _d, _d, _d, _vp, _vp, _vp,
                                     _vp]),
    "pcg_column_products_dev": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _d, _d, _vp, _vp, _vp, _vp, _vp]),
    "pcg_wind_gather": (_i, [_vp, ctypes.c_longlong, _vp, _vp, ctypes.c_longlong, _d, _d, _i, _d, _vp, _vp, _vp]),
    "pcg_wind_vertical": (_i, [_vp, _i, ctypes.c_longlong, c_double_p, _i, _d, _d, _vp, _vp, _vp]),
    "pcg_section_sample": (_i, [_vp, _i, _i, _vp, _vp, _d, _d, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "pcg_vvp_local": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _i, _d, _i, _d, _vp, _vp]),
    "pcg_debug_cartesian_to_antenna": (_i, [_vp, _vp, ctypes.c_size_t, _d, _d, _d, _vp, _vp, _vp]),
    "pcg_debug_sqrt_check": (_i, [_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64), c_double_p]),
    "pcg_debug_xye_to_antenna": (_i, [_vp, _vp, ctypes.c_size_t, _d, _d, _d, _vp, _vp]),
}


class NetworkParams(ctypes.Structure):
    """pcg_network_params (include/pycwr_b200.h)."""
    _fields_ = [("method", _i), ("blind_code", _i), ("quality_terms", _i), ("sanity_filter", _i),
                ("influence_radius_m", _d), ("range_decay_m", _d), ("beam_radius_ref_m", _d),
                ("vertical_gap_ref_deg", _d)]


NET_METHODS = {"max": 0, "nearest": 1, "exp_weighted": 2, "quality_weighted": 3}
QUALITY_TERMS = {"range": 1, "beam": 2, "vertical": 4}


class PcgError(RuntimeError):
    """CUDA / allocation failure reported by the native library."""


_lib = None


def load():
    """Load the native library (ImportError with build instructions when it is absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pycwr_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C pycwr_b200/csrc`). There is no CPU fallback." % LIB_PATH
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library diverge
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    """Map a pcg_status to the exception the reference would have raised."""
    if rc == PCG_OK or rc == PCG_WARN_NONFINITE:
        return rc
    msg = load().pcg_last_error().decode("utf-8", "replace")
    if rc == PCG_ERR_ARG:
        raise ValueError(msg)
    if rc == PCG_ERR_NOMEM:
        raise MemoryError(msg)
    raise PcgError(msg)


def device_count():
    n = ctypes.c_int(0)
    if load().pcg_device_count(ctypes.byref(n)) != PCG_OK:
        return 0
    return n.value


def launch_count():
    return int(load().pcg_launch_count())


_tls = threading.local()


def set_device(device):
    """``pcg_set_device`` for the calling thread (CUDA's current device is per host thread; all library state
    is per device, so one thread per GPU can drive the GPUs of a box side by side)."""
    check(load().pcg_set_device(int(device)))
    _tls.device = int(device)


def current_device():
    """The device this thread last selected through :func:`set_device` (0 before any call)."""
    return getattr(_tls, "device", 0)


def as_ptr(arr):
    """Raw data pointer of a C-contiguous numpy array as c_void_p."""
    return ctypes.c_void_p(arr.ctypes.data)
