"""raytrace.install(batched=True): the reference's own analysis classes running on the batched
drivers.

CPU: rayoptics.raytr.analyses.RayFan / RayList / RayGrid (importable) are run on the hybrid
model (oracle/ref_model.py) with their per-ray loops `trace_ray_fan / trace_ray_list /
trace_ray_grid` rebound to rayoptics_b200.trace's batched stand-ins (oracle-fed through the
`tracer=` seam) and must reproduce, bit for bit, what the unpatched reference produced
(tests/golden/vectors/<model>_analyses.npz).  Skipped without /root/reference.
"""
import functools
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_model
from oracle import ref_harness as rh
from rayoptics_b200 import trace as TR
from test_trace_drivers import oracle_tracer

pytestmark = pytest.mark.skipif(not rh.available(), reason='/root/reference not present')


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'telecentric', 'fisheye', 'threemir'])
def test_reference_classes_on_batched_loops(name, monkeypatch):
    from oracle import ref_model
    RT, RA = ref_model.modules()
    for fn in ('trace_ray_fan', 'trace_ray_list', 'trace_ray_grid'):
        monkeypatch.setattr(RA, fn, functools.partial(getattr(TR, 'analyses_' + fn), tracer=oracle_tracer))
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_analyses.npz'))
    n_fan, n_list, n_grid = (int(x) for x in z['num'])
    opm = load_model(name)
    H = ref_model.HybridModel(opm)
    for ci, (f, wl) in enumerate(z['cases']):
        f, wl = int(f), (None if wl < 0 else float(wl))
        fld = opm.optical_spec.field_of_view.fields[f]
        aim = None if fld.aim_info is None else np.array(fld.aim_info)

        def keep_aim():          # as in make_golden_analyses.py: no re-aiming by the reference
            fld.aim_info = None if aim is None else (float(aim) if aim.ndim == 0 else aim.copy())
            fld.chief_ray = ((None, None, -1.0), None)

        for xy in 'xy':
            keep_aim()
            fan = RA.RayFan(H, f=f, wl=wl, xyfan=xy, num_rays=n_fan)
            assert np.array_equal(np.array([p for p, v in fan.fan], dtype=float).reshape(-1, 2),
                                  z[f'fan{xy}_pupil_{ci}'])
            assert np.array_equal(np.array([v for p, v in fan.fan], dtype=float).reshape(-1, 3),
                                  z[f'fan{xy}_vals_{ci}'])
        keep_aim()
        assert np.array_equal(RA.RayList(H, num_rays=n_list, f=f, wl=wl).ray_abr, z[f'list_abr_{ci}'])
        keep_aim()
        assert np.array_equal(RA.RayGrid(H, f=f, wl=wl, num_rays=n_grid).grid, z[f'grid_{ci}'],
                              equal_nan=True)


def test_install_batched_rebinds_and_restores():
    from oracle import ref_model
    from rayoptics_b200 import raytrace as B
    RT, RA = ref_model.modules()
    import rayoptics.raytr.raytrace as rt
    orig = (rt.trace, rt.trace_raw, RT.trace_fan, RT.trace_grid, RA.trace_ray_fan, RA.trace_ray_list,
            RA.trace_ray_grid)
    B.install(batched=True)
    try:
        assert rt.trace is B.trace and RA.trace_ray_grid is not orig[6] and RT.trace_fan is not orig[2]
        assert RA.trace_ray_fan.__wrapped__ is orig[4]
    finally:
        B.uninstall()
    assert (rt.trace, rt.trace_raw, RT.trace_fan, RT.trace_grid, RA.trace_ray_fan, RA.trace_ray_list,
            RA.trace_ray_grid) == orig
