"""Multi-GPU sharding of the network composite: one process per GPU, the output
grid split into contiguous ROW slabs, no data-path collective until the finished slabs are gathered.

Every output voxel depends only on read-only radar volumes, so rank ``r`` runs the fused network
kernel over rows ``[r0, r1)`` of the grid with the radars whose range disc can reach that slab and
owns its slab of ``composite``, ``coverage_count``, ``blind_mask`` and ``CR``.  The only exchange is
``gather_slabs`` (``torch.distributed.all_gather`` — NCCL over NVLink on the GPU box, gloo in the
CPU tests).  The reference's counterpart is the fork pool over radars
(``pycwr/interp/RadarInterp.py:1816-1861``), which returns whole per-radar volumes through pipes.
"""
import numpy as np


def slab_bounds(nrows, world):
    """Balanced contiguous row slabs: ``[(r0, r1)] * world`` (empty slabs when world > nrows)."""
    nrows, world = int(nrows), int(world)
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(nrows, world)
    bounds, r0 = [], 0
    for r in range(world):
        r1 = r0 + base + (1 if r < extra else 0)
        bounds.append((r0, r1))
        r0 = r1
    return bounds


def row_costs(grid_x, grid_y, sites_xy, reach_m):
    """Work estimate per grid row: the total chord length of the radars' reach discs along the row
    (the number of (column, radar) pairs the composite has to grid there), from the row's mean
    ordinate and abscissa extent, plus a base cost per row for the passes that touch every voxel
    (accumulator reset, finalize: about a third of one radar's share on the C4 bench)."""
    gx = np.asarray(grid_x, dtype=np.float64)
    gy = np.asarray(grid_y, dtype=np.float64)
    if gx.size == 0:
        return np.zeros(gx.shape[0])
    y = gy.mean(axis=1)
    x0, x1 = gx.min(axis=1), gx.max(axis=1)
    width = np.maximum(x1 - x0, 1.0)
    cost = np.full(gx.shape[0], 0.3)
    for (sx, sy), r in zip(np.asarray(sites_xy, dtype=np.float64), reach_m):
        half = np.sqrt(np.maximum(float(r) ** 2 - (y - sy) ** 2, 0.0))
        lo, hi = np.maximum(sx - half, x0), np.minimum(sx + half, x1)
        cost += np.maximum(hi - lo, 0.0) / width
    return cost


def slab_bounds_weighted(costs, world):
    """Contiguous row slabs of (nearly) equal total cost: rank r owns rows [b[r][0], b[r][1])."""
    costs = np.asarray(costs, dtype=np.float64)
    world = int(world)
    if world < 1:
        raise ValueError("world must be >= 1")
    n = costs.size
    total = float(costs.sum())
    if n == 0 or not total > 0.0:
        return slab_bounds(n, world)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    cuts = [0]
    for r in range(1, world):
        cut = int(np.searchsorted(cum, total * r / world, side="left"))
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slab_bounds_shares(costs, shares):
    """Contiguous row slabs whose total costs are in the proportions ``shares`` (one entry per rank, any positive
    scale): rank r owns rows [b[r][0], b[r][1]).  ``slab_bounds_weighted`` is the equal-shares case."""
    costs = np.asarray(costs, dtype=np.float64)
    shares = np.asarray(shares, dtype=np.float64)
    world = int(shares.size)
    if world < 1 or not np.all(shares > 0.0):
        raise ValueError("shares must hold one positive entry per rank")
    n = costs.size
    total = float(costs.sum())
    if n == 0 or not total > 0.0:
        return slab_bounds(n, world)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    targets = total * np.cumsum(shares) / float(shares.sum())
    cuts = [0]
    for r in range(world - 1):
        cut = int(np.searchsorted(cum, targets[r], side="left"))
        cuts.append(min(max(cut, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def radar_reach_m(vol_range, margin=1.002, extra_m=1000.0):
    """Ground distance beyond which no gate of the radar can be reached (slant range >= 0.999 x
    ground distance for |z|, |h| << Re; the kernel applies the same cull per block)."""
    return max(float(r[-1]) for r in vol_range) * margin / 0.999 + extra_m


def radars_reaching(grid_x, grid_y, sites_xy, reach_m):
    """Indices of the radars whose disc of radius ``reach_m[i]`` intersects the bounding box of the
    (slab of the) grid — conservative, so dropping the others cannot change any output."""
    gx = np.asarray(grid_x)
    gy = np.asarray(grid_y)
    if gx.size == 0:
        return []
    x0, x1, y0, y1 = float(gx.min()), float(gx.max()), float(gy.min()), float(gy.max())
    keep = []
    for i, (sx, sy) in enumerate(np.asarray(sites_xy, dtype=np.float64)):
        dx = max(x0 - sx, 0.0, sx - x1)
        dy = max(y0 - sy, 0.0, sy - y1)
        if dx * dx + dy * dy <= float(reach_m[i]) ** 2:
            keep.append(i)
    return keep


def gather_slabs(local, bounds, dist=None, group=None, device=None):
    """All-gather row slabs of an array whose axis -2 is the sharded row axis ((nz, rows, ny) volumes
    or (rows, ny) planes).  ``local`` is this rank's slab (numpy or torch); returns the full numpy
    array on every rank.  Slabs may have different row counts: they are padded to the largest."""
    import torch

    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if isinstance(local, torch.Tensor):
        t = local.contiguous()
        np_dtype = np.dtype(str(t.dtype).replace("torch.", ""))
    else:
        arr = np.ascontiguousarray(local)
        np_dtype = arr.dtype
        t = torch.from_numpy(arr)
    # transported as bytes along the last axis: every dtype (int16 / int8 included) works on NCCL and gloo
    t = t.view(torch.uint8)
    if device is not None:
        t = t.to(device)
    rows = [b[1] - b[0] for b in bounds]
    if t.shape[-2] != rows[rank]:
        raise ValueError("local slab has %d rows, expected %d" % (t.shape[-2], rows[rank]))
    maxrows = max(rows)
    pad_shape = list(t.shape)
    pad_shape[-2] = maxrows
    padded = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
    padded[..., : rows[rank], :] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    full = torch.cat([p[..., : rows[r], :] for r, p in enumerate(parts)], dim=-2)
    return np.ascontiguousarray(full.cpu().numpy()).view(np_dtype)


def network_compose_sharded(compute, grid_x, grid_y, sites_xy, reach_m, dist=None, group=None, device=None,
                            keys=("composite", "valid_count", "coverage_count", "blind_mask", "CR"), gather=True,
                            balance=True):
    """Run ``compute(rows, radar_indices) -> dict`` on this rank's slab and gather the results.

    Slabs are balanced by ``row_costs`` (``balance=False``: equal row counts).
    ``compute`` is the per-slab worker (``pycwr_b200.interp.network_compose`` bound to its
    volumes on the GPU box).  Returns ``(result, (r0, r1), radar_indices)``; with ``gather`` the
    result holds full-grid arrays on every rank, otherwise this rank's slab only."""
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    # every rank derives the same bounds from the same inputs: no exchange needed
    bounds = (slab_bounds_weighted(row_costs(grid_x, grid_y, sites_xy, reach_m), world) if balance
              else slab_bounds(np.asarray(grid_x).shape[0], world))
    r0, r1 = bounds[rank]
    radars = radars_reaching(np.asarray(grid_x)[r0:r1], np.asarray(grid_y)[r0:r1], sites_xy, reach_m)
    local = compute((r0, r1), radars)
    if not gather:
        return local, (r0, r1), radars
    out = {}
    for key in keys:
        if key in local:
            out[key] = gather_slabs(np.asarray(local[key]), bounds, dist=dist, group=group, device=device)
    return out, (r0, r1), radars
