#!/usr/bin/env python
"""bench.py — grid voxels/s of the polar-volume -> Cartesian gridding hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c5] [--style stress|storm]
    python bench.py --impl reference ...      # the reference's own Cython path on the host cores

One "step" = one pass of the hot path over one batch of synthetic input: a full single-radar
3-D gridding (BASELINE config C3: 9x360x1840 VCP21-like volume -> 1201x1201x40 lon/lat grid at
250 m).  `value` is timed with CUDA events around the kernel with every input resident in HBM;
`e2e` is the same metric through the reference-facing API (pycwr_b200.RadarGridC.get_CAPPI_3d)
with HOST buffers — volume upload, grid H2D, kernel, D2H — inside the timed region.
With N > 1 (torchrun, one rank per GPU) every rank grids its own radar volume (radars are
independent objects: weak scaling, no data-path collective); `value` is the aggregate.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FILL = -999.0
METRIC = "grid voxels/sec"
UNIT = "voxels/s"
WORKLOADS = {"c1": 1, "c2": 2, "c3": 3, "c5": 5}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(cfg, nfield):
    """every input read once, every output written once, reference dtypes."""
    vol = cfg["volume"]
    ncol = cfg["GridX"].size
    nval = sum(a.size for a in vol["vol_value"][cfg["fields"][0]])
    ntab = sum(a.size for a in vol["vol_azimuth"]) + sum(a.size for a in vol["vol_range"])
    ne = vol["fix_elevation"].size
    if cfg["kind"] == "cr":
        return 8 * ncol + 16 * ncol + 8 * nval + 8 * ntab + 8 * ne
    nz = cfg["levels"].size
    return 8 * nfield * nz * ncol + 16 * ncol + 8 * nfield * nval + 8 * ntab + 8 * (ne + nz)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own Cython kernels (oracle/_ref) when they were
# built, else the C restatement (oracle/grid_oracle.c) — the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------------

def _cpu_impl():
    from oracle import build_ref
    ref = build_ref.load()
    if ref is not None:
        return ref, "reference"
    from oracle import oracle as port
    return port, "port"


_SHARED = {}      # inherited by forked workers: no pickling of the 48 MB volume per task


def _cpu_chunk_shared(levels):
    a = _SHARED["args"]
    return _cpu_chunk((a[0], a[1], a[2], a[3], a[4], levels, a[5]))


def _cpu_chunk(args):
    mod_kind, vol, h, gx, gy, levels, blind = args
    mod, _ = _cpu_impl()
    va, vr, fe, vv = vol
    t0 = time.perf_counter()
    out = mod.get_CAPPI_3d(va, vr, fe, vv, h, gx, gy, levels, FILL, None, blind)
    return out.size, time.perf_counter() - t0


def cpu_sample(cfg, nlevels, workers):
    """Time the CPU implementation on a bounded sample: the first `nlevels` levels per worker of
    the full grid.  Returns (voxels/s, kind, cores, description)."""
    from pycwr_b200 import synth
    mod, kind = _cpu_impl()
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"], cfg["fields"][0])
    gx, gy = cfg["GridX"], cfg["GridY"]
    levels = cfg["levels"] if cfg["levels"] is not None else np.array([1500.0])
    blind = cfg["blind_method"]
    if workers <= 1:
        lv = np.ascontiguousarray(levels[:nlevels])
        n, dt = _cpu_chunk((kind, (va, vr, fe, vv), h, gx, gy, lv, blind))
        return n / dt, kind, 1, "%s x %d of %d levels, %dx%d grid" % (cfg["name"], lv.size, levels.size, gx.shape[0], gx.shape[1])
    import multiprocessing as mp
    ctx = mp.get_context("fork")     # the reference's own strategy (RadarInterp.py:1816-1824)
    chunks = []
    for w in range(workers):
        lv = np.ascontiguousarray(levels[(w * nlevels) % levels.size:][:nlevels])
        if lv.size == 0:
            lv = np.ascontiguousarray(levels[:nlevels])
        chunks.append(lv)
    _SHARED["args"] = (kind, (va, vr, fe, vv), h, gx, gy, blind)
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_chunk_shared, [np.ascontiguousarray(levels[:1])] * workers)   # start the workers
        t0 = time.perf_counter()
        res = pool.map(_cpu_chunk_shared, chunks, chunksize=1)
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return total / wall, kind, workers, "%s: %d workers x %d levels, %dx%d grid" % (
        cfg["name"], workers, nlevels, gx.shape[0], gx.shape[1])


def cpu_full_call(cfg, workers):
    """The reference's implementation on the WHOLE call of the workload: its levels are split into contiguous
    chunks over `workers` forked processes (the reference's own fan-out strategy, RadarInterp.py:1816-1824), so
    one step does exactly the work of one product step.  Returns (voxels/s, kind, workers used, description)."""
    from pycwr_b200 import synth
    mod, kind = _cpu_impl()
    va, vr, fe, vv, h = synth.vol_tuple(cfg["volume"], cfg["fields"][0])
    gx, gy = cfg["GridX"], cfg["GridY"]
    levels = cfg["levels"] if cfg["levels"] is not None else np.array([1500.0])
    blind = cfg["blind_method"]
    workers = max(1, min(workers, int(levels.size)))
    chunks = [np.ascontiguousarray(c) for c in np.array_split(levels, workers) if c.size]
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    _SHARED["args"] = (kind, (va, vr, fe, vv), h, gx, gy, blind)
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_cpu_chunk_shared, [np.ascontiguousarray(levels[:1])[:0]] * len(chunks))   # start the workers
        t0 = time.perf_counter()
        res = pool.map(_cpu_chunk_shared, chunks, chunksize=1)
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return total / wall, kind, len(chunks), "%s: the whole call, %d levels split over %d workers, %dx%d grid" % (
        cfg["name"], levels.size, len(chunks), gx.shape[0], gx.shape[1]), wall


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host threads; a step is the workload's whole
    call (the same voxels a product step produces), so ms_per_step compares with the product arm's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pycwr_b200 import synth
    cfg = synth.config(WORKLOADS[args.workload], style=args.style)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    vals, walls = [], []
    desc = kind = None
    for _ in range(args.steps):
        v, kind, workers_used, desc, w = cpu_full_call(cfg, workers)
        vals.append(v)
        walls.append(w)
    workers = workers_used
    wall = float(np.sum(walls))
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": cfg["name"], "style": args.style},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



# --------------------------------------------------------------------------------------------
# network composite (BASELINE config C4): 32 radars -> 3000 x 3000 x 30, row slabs over the ranks
# --------------------------------------------------------------------------------------------

class _DevArray(object):
    """A float64 device buffer owned by the library, viewable as a torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _fused_gather(torch, dist, dev, rank, world, lib, launch_args, nlat, r0, steps, flush, reference):
    """Network composite with the slab gather fused into the tile kernel (pcg_network_compose_gather_dev):
    timed like the NCCL variant (CUDA events around the launch, max over ranks, barrier outside the events) and
    checked bit for bit against the NCCL all_gather of the same slabs."""
    import ctypes

    from pycwr_b200 import _lib
    a = launch_args
    nz, nlon = a["nz"], a["nlon"]
    nbytes = nz * nlat * nlon * 8
    full = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    peers, opened, err = [], [], None

    def agree(ok):
        """Every rank must take the same branch: a failure anywhere cancels the fused run everywhere."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    try:
        _lib.check(lib.pcg_peer_alloc(nbytes, ctypes.byref(full), handle))
    except Exception as exc:  # noqa: BLE001
        err = exc
    if not agree(err is None):
        if full.value:
            lib.pcg_peer_free(full)
        raise RuntimeError("peer buffer allocation failed on a rank: %r" % (err,))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle))
    try:
        for r in range(world):
            if r == rank:
                peers.append(full.value)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            _lib.check(lib.pcg_peer_open(buf, ctypes.byref(p)))
            peers.append(p.value)
            opened.append(p)
    except Exception as exc:  # noqa: BLE001
        err = exc
    if not agree(err is None):
        for p in opened:
            lib.pcg_peer_close(p)
        dist.barrier()
        lib.pcg_peer_free(full)
        raise RuntimeError("peer mapping failed on a rank: %r" % (err,))
    parr = (ctypes.c_void_p * world)(*peers)
    dp = _lib.c_double_p

    def launch():
        if a["rows"] == 0:
            return
        if a["n"] == 0:
            # no radar reaches this slab: its rows are empty in every rank's grid (written through the same mappings)
            for q in range(world):
                torch.as_tensor(_DevArray(peers[q], (nz, nlat, nlon)), device=dev)[:, r0:r0 + a["rows"]] = FILL
            return
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.pcg_network_compose_gather_dev(
            a["n"], a["handles"], a["rx"].ctypes.data_as(dp), a["ry"].ctypes.data_as(dp), a["rh"].ctypes.data_as(dp),
            a["re"].ctypes.data_as(dp), a["bwp"], ctypes.byref(a["prm"]), a["gx"].data_ptr(), a["gy"].data_ptr(),
            a["rows"], nlon, a["levels"].ctypes.data_as(dp), nz, FILL, None, None, a["cov"].data_ptr(),
            a["blind"].data_ptr(), a["cr"].data_ptr(), world, parr, nlat, r0, st))

    try:
        for _ in range(2):
            launch()
            torch.cuda.synchronize()
            dist.barrier()
        # every rank now holds the full grid: compare with the NCCL-gathered slabs
        parts, bounds = reference
        mine = torch.as_tensor(_DevArray(full.value, (nz, nlat, nlon)), device=dev)
        same = True
        for r, (b0, b1) in enumerate(bounds):
            same = same and bool(torch.equal(mine[:, b0:b1], parts[r][:, : b1 - b0]))
        ok = torch.tensor([1 if same else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            flush.zero_()
            dist.barrier()
            e0.record()
            launch()
            e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in ev)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del mine
        return {"ms_per_step": float(t[0].item()) / steps, "check": "bit-identical to the NCCL all_gather" if int(ok.item())
                else "MISMATCH against the NCCL all_gather",
                "how": "net_tile_kernel stores each finished tile of the slab into every peer's grid (NVLink P2P, CUDA IPC mappings)"}
    finally:
        torch.cuda.synchronize()
        dist.barrier()
        for p in opened:
            lib.pcg_peer_close(p)
        dist.barrier()
        lib.pcg_peer_free(full)


def _static_traffic(key):
    """DRAM bytes per launch of the dominant kernel of workload `key`, from the committed ncu capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key, {}).get("bytes")
    except Exception:                                   # noqa: BLE001
        return None


def gpu_numa_cpus(torch, index):
    """CPUs local to GPU `index` (sysfs local_cpulist of its PCI function), or None when the box does not say."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        text = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        return cpus or None
    except Exception:                                   # noqa: BLE001 - no sysfs entry / no permission: run unbound
        return None


def network_e2e_devices(args, devices):
    """The network product end to end as ONE host call over `devices` (interp.network_compose(devices=...)): host
    grids in; composite (NaN in its empty cells), coverage, blind mask, CR, VIL, ET and the topped flag out, every
    slab returned by its own GPU into the shared pinned arrays.  Volumes are packed per device on the first call."""
    from pycwr_b200 import interp, synth
    from pycwr_b200.runtime import HostVolume
    cfg = synth.network_config(style="stress", scale=args.network_scale, nradar=args.network_radars)
    hvs = [HostVolume(v["vol_azimuth"], v["vol_range"], v["fix_elevation"], v["vol_value"]["dBZ"], fillvalue=FILL)
           for v in cfg["volumes"]]
    sites = np.stack([cfg["radar_x"], cfg["radar_y"]], axis=1)
    heights = [v["radar_height"] for v in cfg["volumes"]]
    bws = [v["beam_widths"] for v in cfg["volumes"]]
    from pycwr_b200.runtime import pinned_copy
    gx, gy, lv = pinned_copy(cfg["GridX"]), pinned_copy(cfg["GridY"]), cfg["levels"]      # inputs in pinned host memory
    call = lambda: interp.network_compose(                                                      # noqa: E731
        hvs, sites, heights, gx, gy, lv, fillvalue=FILL, beam_widths=bws, missing=float("nan"),
        outputs=("coverage_count", "blind_mask", "CR"), column_products=(18.0, 56.0, 18.0), devices=devices)
    try:
        for _ in range(2 if len(devices) == 1 else 5):      # (the slab shares follow the measured device rates)
            res = call()
        reps, ts = 3, []
        for _ in range(reps):
            t0 = time.perf_counter()
            res = call()
            ts.append(time.perf_counter() - t0)
        dt = float(np.mean(ts))
        out_bytes = sum(int(np.asarray(a).nbytes) for k, a in res.items() if hasattr(a, "nbytes"))
        slab_rows = list(res["slab_rows"])
        del res
        one_gpu_ms = None
        if len(devices) > 1:
            # the SAME composite unsharded on one GPU of this box, in this run (inputs resident, device-timed like
            # the sharded `ms_per_step`): what the strong-scaling numbers of this line are measured against
            one_gpu_ms = _network_one_gpu_ms(devices[0], hvs, sites, heights, bws, cfg)
    finally:
        for h in hvs:
            h.close()
    voxels = int(lv.size) * gx.size
    return {"value": voxels / dt, "unit": UNIT, "ms_per_call": dt * 1e3, "ms_best": float(min(ts)) * 1e3,
            "n_gpus": len(devices), "slab_rows": slab_rows, "one_gpu_kernel_ms_same_run": one_gpu_ms,
            "h2d_bytes_per_step": int(gx.nbytes + gy.nbytes), "d2h_bytes_per_step": out_bytes,
            "api": "pycwr_b200.interp.network_compose(devices=[0..N-1]): one host call, one thread per GPU, every "
                   "slab copied to the shared pinned result by its own GPU (no device-to-device exchange)"}


def _network_one_gpu_ms(device, hvs, sites, heights, bws, cfg, steps=3):
    """Device time of the unsharded C4 composite on `device` (pcg_network_compose_dev, full grid, L2 flushed)."""
    import ctypes

    import torch

    from pycwr_b200 import _lib
    _lib.set_device(device)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        dvs = [h.on(device) for h in hvs]
        gx = torch.from_numpy(np.ascontiguousarray(cfg["GridX"])).to(dev)
        gy = torch.from_numpy(np.ascontiguousarray(cfg["GridY"])).to(dev)
        levels = cfg["levels"]
        nz, (nlat, nlon) = int(levels.size), cfg["GridX"].shape
        comp = torch.empty((nz, nlat, nlon), dtype=torch.float64, device=dev)
        cov = torch.empty((nz, nlat, nlon), dtype=torch.int16, device=dev)
        blind = torch.empty((nz, nlat, nlon), dtype=torch.int8, device=dev)
        cr = torch.empty((nlat, nlon), dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        n = len(dvs)
        lib = _lib.load()
        prm = _lib.NetworkParams(_lib.NET_METHODS["quality_weighted"], 3, 7, 1, 300000.0, 230000.0, 3000.0, 0.5)
        handles = (ctypes.c_void_p * n)(*[v.handle for v in dvs])
        dp = _lib.c_double_p
        rx, ry = np.ascontiguousarray(sites[:, 0]), np.ascontiguousarray(sites[:, 1])
        rh = np.ascontiguousarray(np.array(heights, dtype=np.float64))
        re = np.full(n, 8494666.6666666661)
        keep = [np.ascontiguousarray(b) for b in bws]
        bwp = (_lib.c_double_p * n)(*[b.ctypes.data_as(dp) for b in keep])

        def launch():
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.pcg_network_compose_dev(
                n, handles, rx.ctypes.data_as(dp), ry.ctypes.data_as(dp), rh.ctypes.data_as(dp), re.ctypes.data_as(dp),
                bwp, ctypes.byref(prm), gx.data_ptr(), gy.data_ptr(), nlat, nlon, levels.ctypes.data_as(dp), nz, FILL,
                comp.data_ptr(), None, cov.data_ptr(), blind.data_ptr(), cr.data_ptr(), None, st))

        for _ in range(3):
            launch()
            flush.zero_()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            launch()
            b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        del comp, cov, blind, cr, flush, gx, gy
        torch.cuda.empty_cache()
    return ms


def bench_network(args, torch, dist, dev, rank, world):
    """Fused network kernel on this rank's row slab (inputs resident in HBM), then the NCCL gather
    of the finished composite slabs.  Returns the `network` object of the JSON line (rank 0)."""
    import ctypes

    from pycwr_b200 import _lib, sharding, synth
    from pycwr_b200.runtime import DeviceVolume

    import numpy as _np
    # sites / grid first (cheap); volumes only for the radars that reach this rank's slab
    meta = synth.network_config(style="stress", scale=args.network_scale, nradar=args.network_radars,
                                with_volumes=False)
    gx_h, gy_h, levels = meta["GridX"], meta["GridY"], meta["levels"]
    sites = _np.stack([meta["radar_x"], meta["radar_y"]], axis=1)
    sites_cfg = {"alt": meta["radar_alt"]}
    nlat, nlon = gx_h.shape
    nz = int(levels.size)
    reach = [460e3 * 1.002 / 0.999 + 1000.0] * len(sites)
    # row slabs balanced by the (column, radar) pairs each row holds: the middle of the network is denser
    bounds = sharding.slab_bounds_weighted(sharding.row_costs(gx_h, gy_h, sites, reach), world)
    r0, r1 = bounds[rank]
    mine = sharding.radars_reaching(gx_h[r0:r1], gy_h[r0:r1], sites, reach)
    vols, heights, bws = [], [], []
    for k in mine:
        v = synth.make_volume(synth.BASE_SEED + 104 + k, fields=("dBZ",), style="stress",
                              radar_height=sites_cfg["alt"][k])
        vols.append(DeviceVolume(v["vol_azimuth"], v["vol_range"], v["fix_elevation"], v["vol_value"]["dBZ"],
                                 fillvalue=FILL, value_dtype="f32"))
        heights.append(v["radar_height"])
        bws.append(_np.ascontiguousarray(v["beam_widths"]))
        del v
    n = len(mine)
    rows = r1 - r0
    gx = torch.from_numpy(gx_h[r0:r1]).to(dev)
    gy = torch.from_numpy(gy_h[r0:r1]).to(dev)
    comp = torch.empty((nz, rows, nlon), dtype=torch.float64, device=dev)
    cov = torch.empty((nz, rows, nlon), dtype=torch.int16, device=dev)
    blind = torch.empty((nz, rows, nlon), dtype=torch.int8, device=dev)
    cr = torch.empty((rows, nlon), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    prm = _lib.NetworkParams(_lib.NET_METHODS["quality_weighted"], 3, 7, 1, 300000.0, 230000.0, 3000.0, 0.5)
    handles = (ctypes.c_void_p * max(n, 1))(*[v.handle for v in vols])
    rx = _np.ascontiguousarray(sites[mine, 0]) if n else _np.zeros(1)
    ry = _np.ascontiguousarray(sites[mine, 1]) if n else _np.zeros(1)
    rh = _np.ascontiguousarray(_np.array(heights, dtype=_np.float64)) if n else _np.zeros(1)
    re = _np.full(max(n, 1), 8494666.6666666661)
    bwp = (_lib.c_double_p * max(n, 1))(*[b.ctypes.data_as(_lib.c_double_p) for b in bws]) if n else None
    dp = _lib.c_double_p

    def launch():
        if n == 0 or rows == 0:
            return
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.pcg_network_compose_dev(
            n, handles, rx.ctypes.data_as(dp), ry.ctypes.data_as(dp), rh.ctypes.data_as(dp), re.ctypes.data_as(dp),
            bwp, ctypes.byref(prm), gx.data_ptr(), gy.data_ptr(), rows, nlon, levels.ctypes.data_as(dp), nz, FILL,
            comp.data_ptr(), None, cov.data_ptr(), blind.data_ptr(), cr.data_ptr(), None, st))

    steps = max(2, min(args.steps, 5))
    for _ in range(3):
        launch()
        flush.zero_()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        launch()
        b.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    # gather of the finished composite slabs (NCCL all_gather, padded slabs), timed on the device
    gather_ms = 0.0
    fused = None
    if dist is not None:
        maxrows = max(b[1] - b[0] for b in bounds)
        padded = torch.zeros((nz, maxrows, nlon), dtype=torch.float64, device=dev)
        padded[:, :rows] = comp
        parts = torch.empty((world, nz, maxrows, nlon), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(parts, padded)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(steps):
            dist.all_gather_into_tensor(parts, padded)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = float(g0.elapsed_time(g1))
        t = torch.tensor([ms, gather_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, gather_ms = float(t[0].item()), float(t[1].item())
        # the same product with the gather fused into the finalize kernel: every rank stores its slab straight
        # into the full composite grid of every peer over NVLink (IPC-mapped buffers, no NCCL on the data path)
        try:
            fused = _fused_gather(torch, dist, dev, rank, world, lib, launch_args=dict(
                n=n, handles=handles, rx=rx, ry=ry, rh=rh, re=re, bwp=bwp, prm=prm, gx=gx, gy=gy, rows=rows, nlon=nlon,
                levels=levels, nz=nz, cov=cov, blind=blind, cr=cr), nlat=nlat, r0=r0, steps=steps, flush=flush,
                reference=(parts, bounds))
        except Exception as exc:  # noqa: BLE001 — peer mapping unavailable: the NCCL numbers stand
            fused = {"error": repr(exc)[:200]}
        del parts, padded
    nrad = torch.tensor([n], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(nrad, op=dist.ReduceOp.SUM)
    voxels = nz * nlat * nlon
    valid_frac = float((comp != FILL).double().mean().item()) if rows else 0.0
    del comp, cov, blind, cr, flush
    torch.cuda.empty_cache()
    e2e = None
    if world == 1 and n:
        # end to end through the host-facing call of the network product (host grids in; composite with NaN in
        # its empty cells, coverage, blind mask, CR, VIL, ET and the topped flag out, every copy timed)
        import time

        from pycwr_b200 import interp
        from pycwr_b200.runtime import pinned_copy
        gx_p, gy_p = pinned_copy(gx_h), pinned_copy(gy_h)                                       # inputs in pinned host memory
        call = lambda: interp.network_compose(                                                  # noqa: E731
            vols, sites[mine], heights, gx_p, gy_p, levels, fillvalue=FILL, beam_widths=bws, missing=float("nan"),
            outputs=("coverage_count", "blind_mask", "CR"), column_products=(18.0, 56.0, 18.0))
        res = call()
        res = call()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            res = call()
        dt = (time.perf_counter() - t0) / reps
        out_bytes = sum(int(_np.asarray(a).nbytes) for k, a in res.items() if k != "tie_voxels")
        e2e = {"value": voxels / dt, "unit": UNIT, "ms_per_call": dt * 1e3,
               "h2d_bytes_per_step": int(gx_h.nbytes + gy_h.nbytes), "d2h_bytes_per_step": out_bytes,
               "api": "pycwr_b200.interp.network_compose (host grids in; composite(NaN), coverage, blind, CR, VIL, ET out)"}
        del res
    for v in vols:
        v.close()
    if world > 1:
        # the host-facing product over all GPUs of the box: one call from rank 0's process, one thread per GPU; the
        # other ranks have released their buffers and wait on a gloo barrier (no NCCL kernel spinning on their GPUs)
        torch.cuda.empty_cache()
        sync = dist.new_group(backend="gloo")
        dist.barrier(group=sync)
        if rank == 0:
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))    # the threads serve GPUs on every socket
            except Exception:                                          # noqa: BLE001
                pass
            try:
                e2e = network_e2e_devices(args, list(range(world)))
            except Exception as exc:                                   # noqa: BLE001 - reported, the device numbers stand
                e2e = {"error": repr(exc)[:300]}
        dist.barrier(group=sync)
    if rank != 0:
        return None
    peak, _ = read_peaks()
    # algorithmic bytes (network): composite + coverage + blind per voxel, grid + CR per
    # column, every radar volume once
    balg = voxels * (8 + 2 + 1) + nlat * nlon * (16 + 8) + args.network_radars * 8 * 9 * 366 * 1840
    kern_s = ms * 1e-3 / steps
    return {
        "workload": "C4_network_%dr_%dx%dx%d" % (args.network_radars, nlat, nlon, nz),
        "value": voxels * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
        "value_with_gather": voxels * steps / ((ms + gather_ms) * 1e-3), "gather_ms_per_step": gather_ms / steps,
        "fused_gather": None if fused is None else (fused if "error" in fused else dict(
            fused, value=voxels / (fused["ms_per_step"] * 1e-3), unit=UNIT)),
        "n_gpus": world, "scaling": "strong", "partition": "row slabs of the output grid balanced by (column, radar) pairs, radars culled per slab",
        "slab_rows": [b[1] - b[0] for b in bounds],
        "collective": "all_gather of finished composite slabs (NCCL)" if world > 1 else "none",
        "radars_per_rank_mean": float(nrad.item()) / world, "composite_method": "quality_weighted",
        "blind_method": "hybrid", "style": "stress", "gpu_launches": int(launches),
        "outputs": "composite, coverage_count, blind_mask, CR — the variables of run_radar_network_3d's dataset; "
                   "valid_count is not requested (the reference workflow discards it, RadarInterp.py:1957), so samples "
                   "with quality weight zero, which could only show there, are not taken",
        "valid_fraction_rank0": valid_frac, "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": balg / kern_s / 1e9, "peak": peak * world, "unit": "GB/s",
                     "frac": balg / kern_s / 1e9 / (peak * world), "algorithmic_bytes": int(balg),
                     "traffic": _static_traffic("C4_network_32r_3000x3000x30") if world == 1 and args.network_scale == 1.0
                     and args.network_radars == 32 else None,
                     "traffic_source": "profiles/traffic.json (one `ncu --set full` capture of this kernel on this workload; "
                                       "static, not measured in this run)",
                     "kernel": "pcg::net_tile_kernel<3,1> (one launch per level chunk)" if os.environ.get("PCG_NET") != "passes"
                     else "pcg::net_radar_kernel<3,1> x radars + net_cull_kernel + net_finalize_kernel<3>"},
    }


def bench_extra_config(cid, args, torch, dev, steps=10):
    """Kernel-only timing of another BASELINE config on this GPU (inputs resident, L2 flushed between steps):
    the same measurement as the headline `value`, reported under `other_configs` (rank 0 only)."""
    import ctypes

    from pycwr_b200 import RadarGridC as G
    from pycwr_b200 import _lib, synth
    from pycwr_b200.runtime import DeviceVolume
    cfg = synth.config(cid, style=args.style)
    vol, fields = cfg["volume"], cfg["fields"]
    nfield = len(fields)
    dv = DeviceVolume(vol["vol_azimuth"], vol["vol_range"], vol["fix_elevation"],
                      [vol["vol_value"][f] for f in fields] if nfield > 1 else vol["vol_value"][fields[0]],
                      fillvalue=FILL, value_dtype=args.value_dtype)
    gx = torch.from_numpy(cfg["GridX"]).to(dev)
    gy = torch.from_numpy(cfg["GridY"]).to(dev)
    nx, ny = cfg["GridX"].shape
    is_cr = cfg["kind"] == "cr"
    levels = cfg["levels"]
    nz = 1 if is_cr else int(levels.size)
    out = torch.empty((nfield, nz, nx, ny) if not is_cr else (nx, ny), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    blind = {"mask": 0, "hybrid": 3}[cfg["blind_method"]]

    def launch():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if is_cr:
            _lib.check(lib.pcg_get_cr_dev(dv.handle, vol["radar_height"], G.DEFAULT_EFFECTIVE_EARTH_RADIUS, gx.data_ptr(),
                                          gy.data_ptr(), nx, ny, FILL, 0.0, 0.0, out.data_ptr(), st))
        else:
            _lib.check(lib.pcg_get_cappi3d_dev(dv.handle, vol["radar_height"], G.DEFAULT_EFFECTIVE_EARTH_RADIUS,
                                               gx.data_ptr(), gy.data_ptr(), nx, ny, levels.ctypes.data_as(_lib.c_double_p),
                                               nz, FILL, blind, 1.0, 0.0, 0.0, 0, out.data_ptr(), None, st))

    for _ in range(3):
        launch()
        flush.zero_()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        launch()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    peak, _ = read_peaks()
    balg = algorithmic_bytes(cfg, nfield)
    voxels = nfield * nz * nx * ny
    dv.close()
    del out, flush, gx, gy
    torch.cuda.empty_cache()
    return {"workload": cfg["name"], "fields": fields, "grid": [nz, nx, ny], "ms_per_step": ms,
            "value": voxels / (ms * 1e-3), "unit": UNIT + (" (voxel-fields)" if nfield > 1 else ""),
            "roofline": {"achieved": balg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": balg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": int(balg),
                         "traffic": _static_traffic(cfg["name"]),
                         "traffic_source": "profiles/traffic.json (static ncu capture, when one exists for this workload)"}}

def bench_next_rows(args):
    """SURVEY §8(f) rows built beyond the gridding path, timed through their host-facing calls (host buffers,
    copies included): f3 vertical section through the C1-C3 volume, f4 local VVP on one sweep.  The VVP line
    carries a CPU figure from the oracle (= the reference's algorithm, NumPy + LAPACK per window) on a bounded
    sample of windows."""
    import time

    from pycwr_b200 import section, synth, wind
    out = {}
    vol = synth.make_volume(synth.BASE_SEED + 7, style="storm")
    sweeps = [(a, r, np.where(v == FILL, np.nan, v))
              for a, r, v in zip(vol["vol_azimuth"], vol["vol_range"], vol["vol_value"]["dBZ"])]
    sv = section.SectionVolume(sweeps, vol["fix_elevation"])
    x, y, _ = section._build_section_line((-150e3, -40e3), (150e3, 90e3), 250.0)
    for _ in range(3):
        sv.sample_line(x, y, vol["radar_height"])
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        vals = sv.sample_line(x, y, vol["radar_height"])[0]
    dt = (time.perf_counter() - t0) / reps
    sv.close()
    out["f3_section"] = {"workload": "section of %d points through %d sweeps (366 x 1840), volume resident" % (x.size, len(sweeps)),
                         "ms_per_call": dt * 1e3, "samples_per_s": vals.size / dt, "kernel": "pcg::section_kernel"}
    # f4: a 360 x 920 velocity sweep, the reference's default 91 x 9 window
    rng = np.random.default_rng(33)
    naz, nrange = 360, 920
    az = np.sort(np.mod(np.arange(naz) * 1.0 + rng.uniform(-0.3, 0.3, naz), 360.0))
    el = np.full(naz, 1.45)
    a = np.deg2rad(az)[:, None]
    g = np.arange(nrange)[None, :]
    vr = ((9.0 + 0.01 * g) * np.sin(a) + (-4.0 + 0.02 * g) * np.cos(a)) * np.cos(np.deg2rad(1.45)) \
        + rng.normal(0.0, 0.6, (naz, nrange))
    bad = rng.random(vr.shape) < 0.05
    vr[bad] += rng.uniform(8.0, 25.0, int(bad.sum()))
    vr[rng.random(vr.shape) < 0.2] = FILL
    for _ in range(2):
        res = wind._run_local_vvp(az, el, vr, 91, 9)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        res = wind._run_local_vvp(az, el, vr, 91, 9)
    dt = (time.perf_counter() - t0) / reps
    windows = naz * (nrange - 8)
    line = {"workload": "local VVP, %d x %d sweep, 91 x 9 window" % (naz, nrange), "ms_per_sweep": dt * 1e3,
            "windows_per_s": windows / dt, "fits": int(np.isfinite(res["u"]).sum()), "kernel": "pcg::vvp_kernel"}
    if not args.no_cpu:
        from oracle import vvp_oracle
        t0 = time.perf_counter()
        sub = vvp_oracle.run_local_vvp(az, el, vr[:, :24], 91, 9)          # 16 analysed gates x 360 rays
        cdt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": naz * 16 / cdt, "unit": "windows/s", "cores": 1, "kind": "port",
                                "sample": "gates 4..19 of every ray (%d windows)" % (naz * 16)}
        m = np.isfinite(sub["u"][:, 4:20])
        line["max_abs_diff_u_vs_oracle"] = float(np.max(np.abs(sub["u"][:, 4:20][m] - res["u"][:, 4:20][m]))) if m.any() else None
    out["f4_local_vvp"] = line
    return out


# --------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------

def run_product(args):
    import ctypes

    import torch

    from pycwr_b200 import RadarGridC as G
    from pycwr_b200 import _lib, synth
    from pycwr_b200.runtime import DeviceVolume, pinned_copy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _lib.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    _lib.check(_lib.load().pcg_set_device(local))
    numa = None
    if world > 1:
        # one process per GPU: keep it (and the pinned buffers it allocates) on the CPUs next to that GPU
        cpus = gpu_numa_cpus(torch, local)
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
                numa = "%d CPUs local to GPU %d" % (len(cpus), local)
            except Exception:                                          # noqa: BLE001
                numa = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    cid = WORKLOADS[args.workload]
    cfg = synth.config(cid, style=args.style)
    if world > 1:   # every rank grids its own radar (independent volume, same shapes)
        cfg["volume"] = synth.make_volume(synth.BASE_SEED + cid + 1000 * rank,
                                          naz=cfg["volume"]["vol_azimuth"][0].size,
                                          ngate=cfg["volume"]["vol_range"][0].size,
                                          gate_m=float(cfg["volume"]["vol_range"][0][0]),
                                          elevations=cfg["volume"]["fix_elevation"],
                                          fields=tuple(cfg["fields"]), style=args.style)
    vol = cfg["volume"]
    fields = cfg["fields"]
    nfield = len(fields)
    gx_h, gy_h = cfg["GridX"], cfg["GridY"]
    nx, ny = gx_h.shape
    levels = cfg["levels"]
    is_cr = cfg["kind"] == "cr"
    nz = 1 if is_cr else int(levels.size)
    blind = {"mask": 0, "hybrid": 3}[cfg["blind_method"]]
    h = vol["radar_height"]
    Re = G.DEFAULT_EFFECTIVE_EARTH_RADIUS
    dv = DeviceVolume(vol["vol_azimuth"], vol["vol_range"], vol["fix_elevation"],
                      [vol["vol_value"][f] for f in fields] if nfield > 1 else vol["vol_value"][fields[0]],
                      fillvalue=FILL, value_dtype=args.value_dtype)
    G.VALUE_DTYPE = args.value_dtype
    gx = torch.from_numpy(gx_h).to(dev)
    gy = torch.from_numpy(gy_h).to(dev)
    out = torch.empty((nfield, nz, nx, ny) if not is_cr else (nx, ny), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    lib = _lib.load()
    lv_p = levels.ctypes.data_as(_lib.c_double_p) if not is_cr else None

    def launch():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if is_cr:
            _lib.check(lib.pcg_get_cr_dev(dv.handle, h, Re, gx.data_ptr(), gy.data_ptr(), nx, ny, FILL,
                                          0.0, 0.0, out.data_ptr(), st))
        else:
            _lib.check(lib.pcg_get_cappi3d_dev(dv.handle, h, Re, gx.data_ptr(), gy.data_ptr(), nx, ny, lv_p,
                                               nz, FILL, blind, 1.0, 0.0, 0.0, 0, out.data_ptr(), None, st))

    voxels = nfield * nz * nx * ny
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover warm-up, the timed steps and a ~1 s loaded tail
    for _ in range(max(args.warmup, 3)):
        launch()
        flush.zero_()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        starts[i].record()
        launch()
        ends[i].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall
    if dist is not None:
        dist.barrier()
    launches = _lib.launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    if rank == 0:                # keep the GPU under the same load while nvidia-smi samples clocks
        t_tail = time.perf_counter()
        while time.perf_counter() - t_tail < 1.0:
            for _ in range(8):
                launch()
            torch.cuda.synchronize()
    dev_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * voxels * args.steps / (dev_ms * 1e-3)

    # ---- e2e through the public API with host buffers (pinned), every copy inside the timing
    USEP = True
    va_p = [pinned_copy(a) for a in vol["vol_azimuth"]] if USEP else vol["vol_azimuth"]
    vr_p = [pinned_copy(a) for a in vol["vol_range"]]
    vv_p = [[pinned_copy(a) for a in vol["vol_value"][f]] for f in fields]
    gx_p, gy_p = pinned_copy(gx_h), pinned_copy(gy_h)
    fe = vol["fix_elevation"]

    def e2e_once():
        if is_cr:
            return G.get_CR_xy(va_p, vr_p, fe, vv_p[0], h, gx_p, gy_p, FILL)
        if nfield == 1:
            return G.get_CAPPI_3d(va_p, vr_p, fe, vv_p[0], h, gx_p, gy_p, levels, FILL, None, cfg["blind_method"])
        dvol = DeviceVolume(va_p, vr_p, fe, vv_p, fillvalue=FILL, value_dtype=args.value_dtype)
        try:
            return G.get_CAPPI_3d(None, None, None, dvol, h, gx_p, gy_p, levels, FILL, None, cfg["blind_method"])
        finally:
            dvol.close()

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        res = e2e_once()
        del res
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        res = e2e_once()
        checksum += float(res.reshape(-1)[0])
        del res
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * voxels * e2e_steps / e2e_s
    h2d = 8 * nfield * dv.nvalues + dv.table_bytes + 16 * nx * ny
    d2h = 8 * voxels

    # free the single-radar buffers before the network leg
    del out, flush
    dv.close()
    torch.cuda.empty_cache()
    net = None
    if args.network:
        net = bench_network(args, torch, dist, dev, rank, world)
    others = None
    if args.other_configs and rank == 0 and world == 1:
        others = [bench_extra_config(c, args, torch, dev) for c in (1, 2, 5) if WORKLOADS[args.workload] != c]
    next_rows = bench_next_rows(args) if args.other_configs and rank == 0 and world == 1 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = read_peaks()
    balg = algorithmic_bytes(cfg, nfield)
    kern_ms = float(np.mean(step_ms))
    achieved = balg / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(cfg["name"], {}).get("bytes")
        except Exception:
            traffic = None
    cpu = None
    if not args.no_cpu:
        v, kind, cores, desc = cpu_sample(cfg, args.cpu_levels, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if dv.value_dtype == "f64" else "f64 geometry+indices / f32 interpolation",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "style": args.style, "grid": [nz, nx, ny], "fields": fields,
                   "value_storage": dv.value_dtype, "blind_method": cfg["blind_method"],
                   "l2": "256 MB flush write between timed steps (excluded from step time)",
                   "parallelism": "1 radar volume per GPU, no collective" if world > 1 else "single GPU",
                   "cpu_binding": numa,
                   "wall_s_timed_region": wall},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "api": "pycwr_b200.RadarGridC.%s (host buffers in, host array out)"
                % ("get_CR_xy" if is_cr else "get_CAPPI_3d")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": "profiles/traffic.json: dram__bytes_read+write of one `ncu --set full` capture "
                                       "of this kernel on this workload (static, not measured in this run)",
                     "peak_source": peak_src, "algorithmic_bytes": int(balg),
                     "kernel": ("pcg::cr_fast_kernel" if is_cr else
                                ("pcg::cappi3d_r2_kernel" if os.environ.get("PCG_KERNEL") == "r2" else "pcg::cappi3d_r3_kernel"))
                     if dv.value_dtype == "f32"
                     else ("pcg::cr_kernel<double>" if is_cr else "pcg::cappi3d_kernel<double>"),
                     "kernel_ms": kern_ms},
        "cpu_baseline": cpu,
    }
    if net is not None:
        line["network"] = net
    if others:
        line["other_configs"] = others
    if next_rows:
        line["next_rows"] = next_rows
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--style", default="stress", choices=["stress", "storm"])
    ap.add_argument("--value-dtype", default="f32", choices=["f64", "f32"],
                    help="device moment layout: f32 = production pair layout, f64 = reference arithmetic")
    ap.add_argument("--cpu-levels", type=int, default=40,
                    help="levels of the full grid timed on 1 host core (40 = all of C3, about 11 s)")
    ap.add_argument("--ref-levels", type=int, default=3, help="levels per worker for --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--network", dest="network", action="store_true", default=True,
                    help="also run the C4 network composite sharded over the ranks (default)")
    ap.add_argument("--no-network", dest="network", action="store_false")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false", default=True,
                    help="skip the kernel-only timing of C1 / C2 / C5 (N = 1 only)")
    ap.add_argument("--network-radars", type=int, default=32)
    ap.add_argument("--network-scale", type=float, default=1.0, help="1.0 = 3000 x 3000 x 30")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
