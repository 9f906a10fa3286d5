"""Host-side logic of the multi-GPU network composite on CPU: slab partition, radar culling and
the slab gather over a world-size-2 ``gloo`` group.  The per-slab worker is the NumPy oracle (the
CUDA worker has the same call shape); sharded + gathered must equal the unsharded result exactly."""
import os
import socket

import numpy as np
import pytest

from conftest import FILL, ROOT
from pycwr_b200 import sharding as S


def test_slab_bounds_balanced_and_contiguous():
    for nrows, world in ((3000, 8), (34, 2), (5, 8), (0, 2), (7, 1)):
        b = S.slab_bounds(nrows, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == nrows
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [r1 - r0 for r0, r1 in b]
        assert max(sizes) - min(sizes) <= 1
    assert S.slab_bounds(3000, 8)[0] == (0, 375)
    with pytest.raises(ValueError):
        S.slab_bounds(10, 0)


def test_weighted_slabs_balance_the_work():
    gx, gy = np.meshgrid(np.linspace(0.0, 3000e3, 301), np.linspace(0.0, 3000e3, 300), indexing="xy")   # (rows, cols)
    rng = np.random.default_rng(3)
    sites = np.stack([rng.uniform(200e3, 2800e3, 32), rng.uniform(800e3, 2200e3, 32)], axis=1)   # crowded middle
    cost = S.row_costs(gx, gy, sites, [462e3] * 32)
    assert cost.shape == (300,) and cost.min() > 0 and cost[150] > 10 * cost[0]
    for world in (1, 2, 8):
        b = S.slab_bounds_weighted(cost, world)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == 300
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = np.array([cost[r0:r1].sum() for r0, r1 in b])
        assert loads.max() <= 1.15 * loads.mean() + cost.max()
    equal = np.array([cost[r0:r1].sum() for r0, r1 in S.slab_bounds(300, 8)])
    weighted = np.array([cost[r0:r1].sum() for r0, r1 in S.slab_bounds_weighted(cost, 8)])
    assert weighted.max() < 0.8 * equal.max()
    assert S.slab_bounds_weighted(np.zeros(10), 4) == S.slab_bounds(10, 4)


def test_radar_culling_is_conservative():
    gx, gy = np.meshgrid(np.linspace(0.0, 100e3, 11), np.linspace(0.0, 50e3, 6), indexing="ij")
    sites = np.array([[50e3, 25e3], [-300e3, 25e3], [150e3, 80e3], [100e3 + 199e3, 0.0]])
    keep = S.radars_reaching(gx, gy, sites, [200e3] * 4)
    assert keep == [0, 2, 3]
    assert S.radars_reaching(gx[:0], gy[:0], sites, [200e3] * 4) == []
    assert S.radar_reach_m([np.array([1e3, 460e3])]) > 460e3 / 0.999


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, path, outdir):
    import torch.distributed as dist
    from oracle import network_oracle as N
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(path)
    vols, bws = [], []
    for i in range(int(g["N_nradar"])):
        pre = "N_r%d_" % i
        fe = g[pre + "elev"]
        vols.append(([g[pre + "az%d" % k] for k in range(fe.size)], [g[pre + "rng%d" % k] for k in range(fe.size)],
                     fe, [g[pre + "val%d" % k] for k in range(fe.size)], float(g[pre + "h"])))
        bws.append(g[pre + "bw"])
    gx, gy, lv, sites = g["N_gx"], g["N_gy"], g["N_levels"], g["N_sites"]
    reach = [S.radar_reach_m(v[1]) * (0.35 if i == 1 else 1.0) for i, v in enumerate(vols)]
    reach[1] = S.radar_reach_m(vols[1][1])          # keep the cull honest: full reach for every radar

    def compute(rows, radars):
        r0, r1 = rows
        if not radars:
            shape = (lv.size, r1 - r0, gx.shape[1])
            return {"composite": np.full(shape, FILL), "valid_count": np.zeros(shape, np.int16),
                    "coverage_count": np.zeros(shape, np.int16), "blind_mask": np.ones(shape, np.int8),
                    "CR": np.full(shape[1:], np.nan)}
        res = N.network_3d([vols[i] for i in radars], sites[radars], gx[r0:r1], gy[r0:r1], lv, FILL,
                           "quality_weighted", beam_widths=[bws[i] for i in radars])
        return res

    out, rows, radars = S.network_compose_sharded(compute, gx, gy, sites, reach, dist=dist)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), rows=np.array(rows), radars=np.array(radars, dtype=np.int64),
             **{k: out[k] for k in out})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_network_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    path = os.path.join(ROOT, "tests", "golden", "network_reference.npz")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    g = np.load(path)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert r0["rows"][0] == 0 and r0["rows"][1] == r1["rows"][0] and r1["rows"][1] == 30
    for key, ref in (("composite", "N_comp_quality_weighted"), ("valid_count", "N_count_quality_weighted"),
                     ("coverage_count", "N_coverage"), ("blind_mask", "N_blind"), ("CR", "N_CR")):
        assert np.array_equal(r0[key], g[ref], equal_nan=True), key      # gathered == reference's unsharded result
        assert np.array_equal(r0[key], r1[key], equal_nan=True), key     # every rank holds the same full grid
        assert r0[key].dtype == g[ref].dtype, key


def test_slab_bounds_follow_the_given_shares():
    """``slab_bounds_shares``: contiguous slabs whose costs are in the requested proportions (what the multi-GPU
    product uses to give the GPUs with the faster host link more rows); equal shares = ``slab_bounds_weighted``."""
    rng = np.random.default_rng(5)
    costs = rng.uniform(0.5, 3.0, 2000)
    assert S.slab_bounds_shares(costs, [1, 1, 1, 1]) == S.slab_bounds_weighted(costs, 4)
    b = S.slab_bounds_shares(costs, [1, 2, 1, 4])
    assert b[0][0] == 0 and b[-1][1] == costs.size and all(b[i][1] == b[i + 1][0] for i in range(3))
    got = np.array([costs[x:y].sum() for x, y in b]) / costs.sum()
    assert np.allclose(got, np.array([1, 2, 1, 4]) / 8.0, atol=2e-3)
    assert S.slab_bounds_shares(np.zeros(0), [1, 1]) == [(0, 0), (0, 0)]
    with pytest.raises(ValueError):
        S.slab_bounds_shares(costs, [1, 0])
