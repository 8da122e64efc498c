#!/usr/bin/env python3
"""Multi-GPU check, launched with torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank traces its shard of the double-Gauss 200x200 grid through
analyses.spot_diagram(shard=...); the all-gathered spot sums must equal the sums
of an unsharded run on the same GPU, and the shard's rays must equal the
corresponding slice of the unsharded result, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rayoptics_b200 import model as M, table as T, analyses as A   # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                            timeout=datetime.timedelta(seconds=180))
    opm = M.OpticalModel.load(os.path.join(ROOT, 'tests', 'golden', 'models', 'dblgauss.json'))
    tab = T.SurfaceTable.from_model(opm.seq_model, device=local)
    whole = A.spot_diagram(opm, 200, table=tab)
    part = A.spot_diagram(opm, 200, table=tab, shard=(rank, world))
    a, b = part.first_ray, part.first_ray + part.status.shape[0]
    assert np.array_equal(part.status, whole.status[a:b])
    assert np.array_equal(part.abr, whole.abr[:, a:b], equal_nan=True)
    for k in ('n_ok', 'n_blocked', 'n_missed', 'n_tir', 'min_x', 'max_y'):
        assert np.array_equal(part.summary[k], whole.summary[k]), k
    for k in ('centroid_x', 'centroid_y', 'rms_radius', 'mean_op'):
        np.testing.assert_allclose(part.summary[k], whole.summary[k], rtol=1e-11, atol=1e-13)
    counts = torch.tensor([b - a], device='cuda')
    dist.all_reduce(counts)
    assert int(counts) == whole.n_rays_total
    dist.barrier()
    if rank == 0:
        print(f'mgpu_check ok: world={world}, rays={whole.n_rays_total}, '
              f'rms_radius[0]={whole.summary["rms_radius"][0]}')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
