"""Host-side runtime objects: device-resident packed volumes and a pinned-host array pool.

``DeviceVolume`` is the B200 form of the reference's ``PRD.vol`` tuple
(``pycwr/core/NRadar.py:1931-1961``): the per-sweep azimuth / range / value lists packed into
arenas that stay in HBM, so several products (CR, CAPPI levels, VIL/ET volumes, network
passes) can be gridded from one upload.
"""
import ctypes
import threading
import weakref

import numpy as np

from . import _lib


def _f64_1d(a, what):
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got %r (%s)" % (a.dtype.name, what))
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d) (%s)" % (a.ndim, what))
    return np.ascontiguousarray(a)


def _f64_2d(a, what):
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got %r (%s)" % (a.dtype.name, what))
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d) (%s)" % (a.ndim, what))
    return np.ascontiguousarray(a)


class DeviceVolume:
    """One radar volume (1..4 moments sharing azimuth/range tables) resident on the GPU.

    ``value_dtype="f32"`` (default) asks for the production layout (float32 radial pairs, exact
    indices, values within a few float32 ulps); ``"f64"`` keeps float64 moments and runs the
    kernels that restate the reference operation by operation.  The library falls back to
    ``"f64"`` for volumes the packed layout cannot serve (a single sweep, unsorted or duplicate
    table entries); ``self.value_dtype`` reports the layout in use.  ``fillvalue`` must be the
    value later passed to the gridding calls.
    """

    def __init__(self, vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue=-999.0,
                 value_dtype="f32"):
        lib = _lib.load()
        fix = _f64_1d(fix_elevation, "fix_elevation")
        ne = int(fix.shape[0])
        # one moment: list of (naz, ng) arrays; several: list of such lists
        if ne > 0 and len(vol_value) > 0 and isinstance(vol_value[0], (list, tuple)):
            fields = [list(f) for f in vol_value]
        else:
            fields = [list(vol_value)]
        if len(vol_azimuth) < ne or len(vol_range) < ne or any(len(f) < ne for f in fields):
            raise IndexError("list index out of range")   # what the reference raises on use
        az = [_f64_1d(vol_azimuth[s], "vol_azimuth[%d]" % s) for s in range(ne)]
        rg = [_f64_1d(vol_range[s], "vol_range[%d]" % s) for s in range(ne)]
        vals = [[_f64_2d(f[s], "vol_value[%d]" % s) for s in range(ne)] for f in fields]
        naz = np.array([a.shape[0] for a in az], dtype=np.int32)
        ng = np.array([a.shape[0] for a in rg], dtype=np.int32)
        for f in vals:
            for s in range(ne):
                if f[s].shape[0] < naz[s] or f[s].shape[1] < ng[s]:
                    raise IndexError("sweep %d: value array smaller than its azimuth/range tables" % s)
                if f[s].shape != (naz[s], ng[s]):
                    f[s] = np.ascontiguousarray(f[s][: naz[s], : ng[s]])
        n1 = max(ne, 1)
        azp = (_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in az])
        rgp = (_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in rg])
        fld = []
        for f in vals:
            fld.append((_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in f]))
        fpp = (_lib.c_double_pp * len(fld))(*[ctypes.cast(x, _lib.c_double_pp) for x in fld])
        handle = ctypes.c_void_p()
        dtype_code = {"f64": _lib.VALUE_F64, "f32": _lib.VALUE_F32}[value_dtype]
        _lib.check(lib.pcg_volume_create(
            ne, naz.ctypes.data_as(_lib.c_int_p), ng.ctypes.data_as(_lib.c_int_p),
            ctypes.cast(azp, _lib.c_double_pp), ctypes.cast(rgp, _lib.c_double_pp), len(fld),
            ctypes.cast(fpp, _lib.c_double_ppp), fix.ctypes.data_as(_lib.c_double_p), float(fillvalue),
            dtype_code, ctypes.byref(handle)))
        self._adopt(handle, naz, ng, fix, rg, fillvalue, len(fld))

    def _adopt(self, handle, naz, ng, fix, rg, fillvalue, nfield):
        lib = _lib.load()
        self.handle = handle
        self.ne = int(fix.shape[0])
        self.nfield = int(nfield)
        self.fillvalue = float(fillvalue)
        used = ctypes.c_int(0)
        _lib.check(lib.pcg_volume_info(handle, None, None, ctypes.byref(used), None))
        self.value_dtype = "f32" if used.value == _lib.VALUE_F32 else "f64"
        # the fused network kernels take packed volumes of at most 32 sweeps (kNetMaxSweeps); anything else is
        # composited through the reference's per-radar flow (interp._network_field_unfused)
        self.network_ready = self.value_dtype == "f32" and 2 <= int(fix.shape[0]) <= 32
        self.naz = naz
        self.ng = ng
        self.fix_elevation = fix.copy()
        self.first_gate = np.array([r[0] if r.size else np.nan for r in rg], dtype=np.float64)
        self.last_gate = np.array([r[-1] if r.size else np.nan for r in rg], dtype=np.float64)
        self.nvalues = int(sum(int(a) * int(b) for a, b in zip(naz, ng)))
        self.table_bytes = int(8 * (naz.sum() + ng.sum()))
        self._finalizer = weakref.finalize(self, lib.pcg_volume_destroy, handle)

    @classmethod
    def from_sweeps(cls, sweeps, fix_elevation, fillvalue=-999.0, value_dtype="f32", max_range_km=None):
        """Pack sweeps as the file reader left them — ``[(azimuth, ranges, values), ...]`` in scan order, rays
        unsorted, NaN for missing; ``values`` one ``(naz, ng)`` array or a list of them (several moments) —
        the work of ``PRD.get_vol_data`` / ``_extract_sweep_volume`` (``pycwr/core/NRadar.py:665-712,1931-1961``)
        without its host passes: the host only argsorts the azimuths; the row gather, the ``max_range_km`` clip
        (``:655-663``) and NaN -> ``fillvalue`` run inside the device pack kernels (``pcg_volume_create_raw``)."""
        lib = _lib.load()
        fix_all = _f64_1d(fix_elevation, "fix_elevation")
        if len(sweeps) != fix_all.shape[0]:
            raise ValueError("one fixed elevation per sweep is required")
        order = np.argsort(fix_all)                                   # _fixed_angle_sort_index
        fix = np.ascontiguousarray(fix_all[order])
        ne = int(fix.shape[0])
        az, rg, rays, lds, keep_alive = [], [], [], [], []
        nfield, vals = None, None
        for s in order:
            a, r, v = sweeps[int(s)]
            a = _f64_1d(a, "azimuth")
            r = _f64_1d(r, "ranges")
            moments = list(v) if isinstance(v, (list, tuple)) else [v]
            if nfield is None:
                nfield, vals = len(moments), [[] for _ in moments]
            if len(moments) != nfield:
                raise ValueError("every sweep must carry the same moments")
            n_used = r.shape[0]
            mask = None
            if max_range_km is not None:                              # NRadar.py:655-663
                keep = r <= float(max_range_km) * 1000.0
                k = int(keep.sum())
                if k == 0:
                    n_used = min(1, r.shape[0])
                elif keep[:k].all():
                    n_used = k                                        # a prefix of every ray: no copy
                else:
                    mask = keep
            if mask is not None:
                r = np.ascontiguousarray(r[mask])
                n_used = r.shape[0]
            ld = None
            for f, m in enumerate(moments):
                m = np.asarray(m)
                if m.dtype != np.float64 or m.ndim != 2 or not m.flags.c_contiguous:
                    m = np.ascontiguousarray(m, dtype=np.float64)
                    if m.ndim != 2:
                        raise ValueError("Sweep field must be 2-D for Cartesian gridding.")
                if mask is not None:
                    m = np.ascontiguousarray(m[:, mask])
                if m.shape[0] < a.shape[0] or m.shape[1] < n_used:
                    raise IndexError("sweep %d: value array smaller than its azimuth/range tables" % int(s))
                ld = m.shape[1] if ld is None else ld
                if m.shape[1] != ld:
                    raise ValueError("moments of one sweep must share their shape")
                vals[f].append(m)
            az.append(a)
            rg.append(np.ascontiguousarray(r[:n_used]))
            rays.append(np.ascontiguousarray(np.argsort(a), dtype=np.int32))
            lds.append(ld if ld is not None else n_used)
        naz = np.array([a.shape[0] for a in az], dtype=np.int32)
        ng = np.array([r.shape[0] for r in rg], dtype=np.int32)
        n1 = max(ne, 1)
        azp = (_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in az])
        rgp = (_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in rg])
        orp = (_lib.c_int_p * n1)(*[a.ctypes.data_as(_lib.c_int_p) for a in rays])
        ldv = np.array(lds if lds else [0], dtype=np.int64)
        fld = [(_lib.c_double_p * n1)(*[a.ctypes.data_as(_lib.c_double_p) for a in f]) for f in (vals or [[]])]
        fpp = (_lib.c_double_pp * len(fld))(*[ctypes.cast(x, _lib.c_double_pp) for x in fld])
        handle = ctypes.c_void_p()
        dtype_code = {"f64": _lib.VALUE_F64, "f32": _lib.VALUE_F32}[value_dtype]
        _lib.check(lib.pcg_volume_create_raw(
            ne, naz.ctypes.data_as(_lib.c_int_p), ng.ctypes.data_as(_lib.c_int_p),
            ctypes.cast(azp, _lib.c_double_pp), ctypes.cast(rgp, _lib.c_double_pp), len(fld),
            ctypes.cast(fpp, _lib.c_double_ppp), ldv.ctypes.data, ctypes.cast(orp, ctypes.POINTER(_lib.c_int_p)),
            fix.ctypes.data_as(_lib.c_double_p), float(fillvalue), dtype_code, ctypes.byref(handle)))
        self = cls.__new__(cls)
        self._adopt(handle, naz, ng, fix, rg, fillvalue, len(fld))
        self.sweep_order = order
        return self

    def close(self):
        self._finalizer()

    @property
    def device_bytes(self):
        n = ctypes.c_size_t(0)
        _lib.check(_lib.load().pcg_volume_info(self.handle, None, None, None, ctypes.byref(n)))
        return n.value


class HostVolume:
    """One radar volume held on the HOST in the reference's ``PRD.vol`` form, uploaded on demand to whichever GPU a
    product needs it on (``on(device)``, one :class:`DeviceVolume` per device, kept until ``close``).  What the
    multi-GPU network product takes: each GPU packs only the radars that reach its slab of the grid."""

    def __init__(self, vol_azimuth, vol_range, fix_elevation, vol_value, fillvalue=-999.0, value_dtype="f32"):
        self.vol_azimuth, self.vol_range, self.vol_value = list(vol_azimuth), list(vol_range), vol_value
        self.fix_elevation = np.ascontiguousarray(fix_elevation, dtype=np.float64)
        self.fillvalue, self.value_dtype = float(fillvalue), value_dtype
        self.ne = int(self.fix_elevation.shape[0])
        self.last_gate = np.array([r[-1] if len(r) else np.nan for r in self.vol_range], dtype=np.float64)
        self._dev = {}
        self._lock = threading.Lock()

    def on(self, device=None):
        """The packed volume on ``device`` (default: the calling thread's current device)."""
        device = _lib.current_device() if device is None else int(device)
        with self._lock:
            dv = self._dev.get(device)
        if dv is None:
            _lib.set_device(device)
            dv = DeviceVolume(self.vol_azimuth, self.vol_range, self.fix_elevation, self.vol_value,
                              fillvalue=self.fillvalue, value_dtype=self.value_dtype)
            with self._lock:
                self._dev[device] = dv
        return dv

    def close(self):
        with self._lock:
            vols, self._dev = list(self._dev.values()), {}
        for dv in vols:
            dv.close()


class _PinnedPool:
    """Recycles page-locked host blocks: pinning is slow (~ms per 10 MB), reuse is free."""

    def __init__(self, max_cached_bytes=8 << 30):
        self._free = {}
        self._cached = 0
        self._max = max_cached_bytes
        self._lock = threading.Lock()

    @staticmethod
    def _bucket(nbytes):
        b = 1 << 16
        while b < nbytes:
            b <<= 1
        return b

    def empty(self, shape, dtype=np.float64):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        bucket = self._bucket(max(nbytes, 1))
        with self._lock:
            lst = self._free.get(bucket)
            ptr = lst.pop() if lst else None
            if ptr is not None:
                self._cached -= bucket
        if ptr is None:
            p = ctypes.c_void_p()
            _lib.check(_lib.load().pcg_host_alloc(bucket, ctypes.byref(p)))
            ptr = p.value
        raw = (ctypes.c_char * bucket).from_address(ptr)
        arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
        weakref.finalize(raw, self._release, ptr, bucket)
        return arr

    def _release(self, ptr, bucket):
        with self._lock:
            if self._cached + bucket <= self._max:
                self._free.setdefault(bucket, []).append(ptr)
                self._cached += bucket
                return
        try:
            _lib.load().pcg_host_free(ctypes.c_void_p(ptr))
        except Exception:
            pass


pinned = _PinnedPool()


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (fast H2D/D2H), recycled on collection."""
    return pinned.empty(shape, dtype)


def pinned_copy(a):
    out = pinned_empty(np.shape(a), np.asarray(a).dtype)
    out[...] = a
    return out
