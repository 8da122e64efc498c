import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODEL_NAMES = ['singlet', 'dblgauss', 'triplet', 'rc', 'cellphone', 'cellphone_even',
               'evenasph', 'zoom52', 'thin_triplet', 'exotic']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


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
