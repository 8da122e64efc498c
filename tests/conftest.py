import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODEL_NAMES = ['singlet', 'dblgauss', 'triplet', 'rc', 'cellphone', 'cellphone_even',
               'evenasph', 'zoom52', 'thin_triplet', 'exotic', 'threemir', 'fisheye']
# models with diffractive phase elements: the reference evaluates x**k with libm pow(), so the
# device carries tolerance parity there (the oracle, on the same libm, stays bit-exact)
# finite-conjugate relays specified by an angular object-space pupil ('NA', 'f/#')
ANGULAR_MODEL_NAMES = ['relay_na', 'relay_fno']
PHASE_MODEL_NAMES = ['hybrid', 'diffractive', 'diffractive_wild']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    if os.environ.get('B200RT_DRYRUN') == '1':      # opt-in: exercise GPU test CODE on the oracle
        import torch
        if torch.cuda.is_available():
            # never let the oracle stand in for the kernels where the kernels can run
            raise pytest.UsageError('B200RT_DRYRUN=1 is refused on a machine with a CUDA device: '
                                    'the -m gpu tests must run the real engine here')
        import dryrun_engine
        dryrun_engine.install()


@pytest.fixture(scope='session')
def oracle():
    from oracle import rt_oracle
    rt_oracle.lib()
    return rt_oracle


def load_model(name):
    from rayoptics_b200 import model as M
    return M.OpticalModel.load(os.path.join(GOLDEN, 'models', name + '.json'))


def load_vectors(name):
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '.npz'))
    v = {k: z[k] for k in z.files}
    v['cases'] = json.loads(str(v['cases']))
    return v


def seeded_bundle(opm, n, rng):
    """Start rays around the model's pupil / field, some of them wild."""
    osp, sm = opm.optical_spec, opm.seq_model
    fod = osp.fod
    z_pupil = fod.obj_dist + fod.enp_dist
    p0 = np.zeros((3, n))
    d0 = np.zeros((3, n))
    scale = np.where(rng.random(n) < 0.8, 1.05, 3.0)
    aim = fod.enp_radius*scale*rng.uniform(-1, 1, (2, n))
    aims = [f.aim_info for f in osp.field_of_view.fields
            if f.aim_info is not None and np.ndim(f.aim_info) == 1]
    if aims:                                    # off-axis systems: the beam is where the chief rays are aimed
        aim += np.mean(np.array(aims, dtype=float), axis=0)[:, None]
    if abs(sm.gaps[0].thi) > 1e8:
        fmax = abs(osp.fov.max_field_value()) if osp.fov.key[1] == 'angle' else \
            np.degrees(abs(np.arctan(fod.pr_slp0)))
        ang = np.deg2rad(scale*max(fmax, 0.2)*rng.uniform(-1, 1, (2, n)))
        dd = np.array([np.sin(ang[0])*np.cos(ang[1]), np.sin(ang[1]),
                       np.cos(ang[0])*np.cos(ang[1])])
        p0[0], p0[1] = -z_pupil*dd[0]/dd[2], -z_pupil*dd[1]/dd[2]
    else:
        p0[:2] = scale*max(abs(fod.pr_ht0), 0.5)*rng.uniform(-1, 1, (2, n))
    v = np.array([aim[0] - p0[0], aim[1] - p0[1], z_pupil - p0[2]])
    d0 = v/np.sqrt((v*v).sum(0))
    wv = rng.integers(0, len(sm.wvlns), n).astype(np.int32)
    return p0, d0, wv
