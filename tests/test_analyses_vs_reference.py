"""RayFan / RayList / RayGrid of rayoptics_b200.analyses against the SAME classes of the
reference (tests/golden/vectors/<model>_analyses.npz: rayoptics.raytr.analyses run unmodified on
a hybrid model with the reference's trace_raw, generator tests/golden/make_golden_analyses.py).

CPU: the classes' host logic (pupil sampling, vignetting, chief ray / reference sphere set-up,
refocus, wave conversion, result shapes) with the tile tracing fed by the oracle through the
`backend=` seam -- bit for bit, OPD included (the oracle keeps libm pow like the reference).
The GPU run of the same comparison is tests/test_zz_gpu_additions.py.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_model
from rayoptics_b200 import _abi, analyses as A, table as T


class OracleBackend:
    """backend= seam of the analysis classes: chief rays and tiles traced by oracle/rt_oracle.c"""

    def __init__(self, opm):
        from oracle import rt_oracle
        self.o = rt_oracle
        self.descs, self.n_by_wvl, self.wvls = T.describe_model(opm.seq_model)

    def _opts(self, check_apertures):
        return _abi.make_opts(first_surf=1, last_surf=len(self.descs) - 2,
                              check_apertures=check_apertures)

    def chief_rays(self, opt_model, fields, wvls):
        from rayoptics_b200 import engine as E
        osp, sm = opt_model.optical_spec, opt_model.seq_model
        recs, eprad, z_pupil = osp.grid_fields(fields)
        spec = E.PupilGridSpec(recs, [sm.index_for_wavelength(w) for w in wvls], [0.0], [0.0], eprad,
                               z_pupil, apply_vignetting=True, flip_z_dir=sm.z_dir[0])
        p, d, wv, _ = self.o.grid_start_rays(spec.c_spec(), 0, spec.n_rays)
        r = self.o.trace_bundle(self.descs, self.n_by_wvl, p, d, wv, self._opts(False), want_full=True,
                                wvls=self.wvls)
        return r['full'], r['op'], r['status']

    def trace_tile(self, opt_model, spec, want_opd, check_apertures):
        r = self.o.trace_grid(spec.c_spec(), self.descs, self.n_by_wvl, 0, spec.n_rays,
                              self._opts(check_apertures), wvls=self.wvls)
        return {'abr': r['abr'], 'status': r['status'], 'opd': r['opd'] if want_opd else None}


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'triplet', 'telecentric', 'cellphone'])
def test_analysis_classes_equal_the_references(name):
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_analyses.npz'))
    n_fan, n_list, n_grid = (int(x) for x in z['num'])
    opm = load_model(name)
    be = OracleBackend(opm)
    for ci, (f, wl) in enumerate(z['cases']):
        f, wl = int(f), (None if wl < 0 else float(wl))
        for xy in 'xy':
            fan = A.RayFan(opm, f=f, wl=wl, xyfan=xy, num_rays=n_fan, backend=be)
            pup = np.array([p for p, v in fan.fan], dtype=float).reshape(-1, 2)
            val = np.array([v for p, v in fan.fan], dtype=float).reshape(-1, 3)
            assert np.array_equal(pup, z[f'fan{xy}_pupil_{ci}'])
            assert np.array_equal(val, z[f'fan{xy}_vals_{ci}'])
        rl = A.RayList(opm, num_rays=n_list, f=f, wl=wl, backend=be)
        assert np.array_equal(rl.ray_abr, z[f'list_abr_{ci}'])
        rg = A.RayGrid(opm, f=f, wl=wl, num_rays=n_grid, backend=be)
        assert np.array_equal(rg.grid, z[f'grid_{ci}'], equal_nan=True)
