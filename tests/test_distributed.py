"""CPU, world_size 2 (gloo): the multi-GPU sharding logic of parallel.py.

Each rank takes its chunk range of a pupil grid (parallel.shard_chunks), computes
that range's per-(field, wvl) spot sums with the CPU oracle (standing in for the
GPU trace, which needs a device), and the ranks combine them with
parallel.gather_summaries (all_gather_into_tensor + combine).  The result must
equal the single-process sums over the whole grid."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_model
from rayoptics_b200 import _abi, table as T, engine as E, parallel as P


def tile_sums(spec, r, ray_begin):
    """[n_tiles, 16] partial sums of oracle results for rays starting at ray_begin."""
    out = np.zeros((spec.n_tiles, 16))
    out[:, 10] = out[:, 12] = np.inf
    out[:, 11] = out[:, 13] = -np.inf
    n = r['status'].shape[0]
    tiles = (ray_begin + np.arange(n))//spec.rays_per_tile
    for t in np.unique(tiles):
        m = tiles == t
        st = r['status'][m]
        ok = st == 0
        ax, ay = r['abr'][0][m][ok], r['abr'][1][m][ok]
        out[t, 0:5] = [ok.sum(), (st == 1).sum(), (st == 2).sum(), (st == 3).sum(), (st > 3).sum()]
        if ok.any():
            out[t, 5:10] = [ax.sum(), ay.sum(), (ax*ax).sum(), (ay*ay).sum(), (ax*ay).sum()]
            out[t, 10:14] = [ax.min(), ax.max(), ay.min(), ay.max()]
            out[t, 14] = r['op'][m][ok].sum()
    return out


def setup(num=40):
    from oracle import rt_oracle
    opm = load_model('dblgauss')
    descs, n_by_wvl, _ = T.describe_model(opm.seq_model)
    spec = E.grid_spec_for_model(opm, num)
    opts = _abi.make_opts(first_surf=1, last_surf=len(descs) - 2, check_apertures=True)
    return rt_oracle, spec, descs, n_by_wvl, opts


def worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        oracle, spec, descs, n_by_wvl, opts = setup()
        c0, c1 = P.shard_chunks(spec.n_chunks, rank, world)
        r0, r1 = spec.first_ray_of_chunk(c0), spec.first_ray_of_chunk(c1)
        r = oracle.trace_grid(spec.c_spec(), descs, n_by_wvl, r0, r1, opts)
        part = torch.from_numpy(tile_sums(spec, r, r0))
        comb = P.gather_summaries(part)
        q.put((rank, r1 - r0, comb.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_summaries_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oracle, spec, descs, n_by_wvl, opts = setup()
    whole = oracle.trace_grid(spec.c_spec(), descs, n_by_wvl, 0, spec.n_rays, opts)
    ref = tile_sums(spec, whole, 0)
    assert sum(g[1] for g in got) == spec.n_rays
    for rank, _, comb in got:
        assert np.array_equal(comb[:, 0:5], ref[:, 0:5])
        assert np.array_equal(comb[:, 10:14], ref[:, 10:14])
        np.testing.assert_allclose(comb[:, 5:10], ref[:, 5:10], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(comb[:, 14], ref[:, 14], rtol=1e-12)
    stats = E.spot_statistics(torch.from_numpy(got[0][2]))
    assert (stats['n_ok'] > 0).all() and (stats['rms_radius'] >= 0).all()


def test_single_process_gather_is_identity():
    x = torch.arange(32, dtype=torch.float64).reshape(2, 16)
    assert torch.equal(P.gather_summaries(x), x)
