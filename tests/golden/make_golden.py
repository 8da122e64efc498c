#!/usr/bin/env python3
"""Generate the golden ray vectors tests/golden/vectors/*.npz.

Runs ONLY in the build container.  Every vector is produced by the REFERENCE's
own code -- ``rayoptics.raytr.raytrace.trace_raw`` imported from
/root/reference/src through oracle/ref_harness.py -- on the model fixtures
written by make_models.py.  The committed .npz files are what pins the C oracle
(tests/test_oracle_golden.py, CPU) and the CUDA engine (tests/test_gpu_parity.py,
GPU box, where /root/reference does not exist).

Per model one file with
  p0, d0 [3, n]      start point / direction cosines (object interface coords)
  wvl_idx [n]        row of the model's wavelength list
  case [n]           index into `cases` (the trace_raw keyword sets below)
  last [10, n]       ray[-1] of the (possibly partial) RayPkg: p, d, dst, nrml
  op, status, fail_surf, n_seg [n]
  full [n_ifc, 10, n_full]  whole rays of the first n_full rays (NaN padded)
Also writes kat.json: the reference's own known-answer data for this path
(raytr/tests/marginal_ray.py:12-24 and elem/tests/test_profiles.py:127-154),
copied as numbers.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                 # noqa: E402
from rayoptics_b200 import model as M                # noqa: E402
from rayoptics_b200.engine import accumulated_steps  # noqa: E402

REF = '/root/reference/src/rayoptics'
OUT = os.path.join(HERE, 'vectors')
N_FULL = 48

# keyword sets of trace_raw exercised by the vectors ("cases")
def cases_for(n_ifc, extra=False):
    if extra == 'wide':   # wide-angle fields: rays start off the object surface (trace.py:299-300)
        return [dict(c, intersect_obj=False) for c in cases_for(n_ifc)]
    if extra:      # model 'exotic': additionally phantom filtering and a raw (not re-intersected) object
        return cases_for(n_ifc) + [
            dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=True, filter_out_phantoms=True),
            dict(first_surf=2, last_surf=n_ifc - 3, check_apertures=False, intersect_obj=False),
        ]
    return [
        dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=True),     # grid analyses
        dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=False),    # trace() default
        dict(first_surf=0, last_surf=None, check_apertures=False),         # trace_raw() default
        dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=True, pt_inside_fuzz=1e-3,
             eps=1e-9),
    ]


def _extra(opm):
    if opm.optical_spec.field_of_view.is_wide_angle:
        return 'wide'
    return opm.name == 'exotic'


def grid_rays(opm, num, case_id, rays):
    """The reference's square pupil grid (trace.py:563-605) for every field / wvl."""
    osp, sm = opm.optical_spec, opm.seq_model
    xs = accumulated_steps(-1.0, 1.0, num)
    for fld in osp.field_of_view.fields:
        for wi, wvl in enumerate(sm.wvlns):
            for i in range(num):
                for j in range(num):
                    pupil = fld.apply_vignetting(np.array([xs[i], xs[j]]))
                    pt0, dir0 = osp.ray_start_from_osp(pupil, fld, 'rel pupil')
                    if not osp.field_of_view.is_wide_angle and dir0[2]*sm.z_dir[0] < 0:
                        dir0 = -dir0
                    rays.append((pt0, dir0, wi, case_id))


def wild_rays(opm, n, rng, rays):
    """Rays with large pupil / field excursions: misses, TIR, clipping."""
    osp, sm = opm.optical_spec, opm.seq_model
    fod = osp.fod
    thi0 = sm.gaps[0].thi
    n_cases = len(cases_for(sm.get_num_surfaces(), _extra(opm)))
    for k in range(n):
        wi = int(rng.integers(len(sm.wvlns)))
        aim = fod.enp_radius*rng.uniform(-3.0, 3.0, 2)
        pt1 = np.array([aim[0], aim[1], fod.obj_dist + fod.enp_dist])
        if abs(thi0) > 1e8:
            ang = np.deg2rad(rng.uniform(-1.0, 1.0, 2)*(3*abs(osp.fov.max_field_value())
                                                        if osp.fov.key[1] == 'angle' else 2.0))
            d0 = np.array([np.sin(ang[0])*np.cos(ang[1]), np.sin(ang[1]),
                           np.cos(ang[0])*np.cos(ang[1])])
            pt0 = -(fod.obj_dist + fod.enp_dist)*np.array([d0[0]/d0[2], d0[1]/d0[2], 0.])
        else:
            pt0 = np.array([*(rng.uniform(-3.0, 3.0, 2)*max(abs(fod.pr_ht0), 1.0)), 0.0])
        v = pt1 - pt0
        dir0 = v/np.linalg.norm(v)
        rays.append((pt0, dir0, wi, int(rng.integers(n_cases))))


def trace_all(opm, rays):
    sm = opm.seq_model
    n_ifc = sm.get_num_surfaces()
    cases = cases_for(n_ifc, _extra(opm))
    paths = [rh.ref_path(sm, w) for w in sm.wvlns]
    n = len(rays)
    out = dict(p0=np.zeros((3, n)), d0=np.zeros((3, n)), wvl_idx=np.zeros(n, np.int32),
               case=np.zeros(n, np.int32), last=np.zeros((10, n)), op=np.zeros(n),
               status=np.zeros(n, np.int32), fail_surf=np.zeros(n, np.int32),
               n_seg=np.zeros(n, np.int32),
               full=np.full((n_ifc, 10, min(n, N_FULL)), np.nan))
    for k, (pt0, dir0, wi, ci) in enumerate(rays):
        r = rh.ref_trace(paths[wi], pt0, dir0, sm.wvlns[wi], **cases[ci])
        out['p0'][:, k], out['d0'][:, k] = pt0, dir0
        out['wvl_idx'][k], out['case'][k] = wi, ci
        out['op'][k], out['status'][k] = r['op'], r['status']
        out['fail_surf'][k], out['n_seg'][k] = r['fail_surf'], r['n_seg']
        if r['n_seg'] > 0:
            out['last'][:, k] = r['ray'][-1]
        if k < N_FULL:
            out['full'][:r['n_seg'], :, k] = r['ray']
    out['cases'] = np.array(json.dumps(cases))
    return out


def write_kat():
    spec = importlib.util.spec_from_file_location('mr', f'{REF}/raytr/tests/marginal_ray.py')
    mr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mr)
    spec = importlib.util.spec_from_file_location('dg', f'{REF}/raytr/tests/ag_dblgauss_s.py')
    dg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dg)
    kat = {
        'source': ['raytr/tests/marginal_ray.py:12-24', 'raytr/tests/ag_dblgauss_s.py',
                   'raytr/tests/test_sequential.py:23-77', 'elem/tests/test_profiles.py:127-154'],
        'ag_dblgauss': [list(r) for r in dg.ag_dblgauss],
        'marginal_ray_f1r2': mr.rayf1r2,
        'wvl': 587.6, 'epd_half': 25.0, 'rel_tol': 3e-6,
    }
    # elem/tests/test_profiles.py:127-154 test_dbgauss_s1: Spherical(c=1/r1), ray from
    # p=[0, 25, 0] along +z (p0 in the test is [0, 0, -1] -> s = 1)
    kat['profile_s1'] = {'r1': 56.20238, 'y0': 25.0, 's': 5.866433424372758, 'rtol': 1e-14}
    # codev/tests/threemrc.lis (CODE V's own ray listing of codev/tests/threemir.seq), axial field:
    # rays launched parallel to the axis at height y0 on surface 1; (x, y, z) at surfaces 1..5 and
    # TAN Y after them.  Pins the decenter / tilt convention (elem/surface.py:274-337,
    # elem/transform.py:145-166, transforms3d euler2mat 'rxyz') to the listing's 6 decimals.
    kat['threemir_lis'] = {
        'source': 'codev/tests/threemrc.lis:233-262',
        'abs_tol': 5e-6,
        'rays': [
            {'y0': 144.927530, 'xyz': [[0., 144.927530, 0.], [0., 97.957122, -4.527500], [0., 0., 0.],
                                       [0., -6.195034, -0.059224], [0., -113.854355, -14.557556]],
             'tan_y': [0., 0.344215, 0.309800, -0.363227, -0.165989]},
            {'y0': 194.927530, 'xyz': [[0., 194.927530, 0.], [0., 149.393568, -10.521171],
                                       [0., 23.585027, 0.], [0., 17.314156, -0.463156],
                                       [0., -75.816236, -6.384849]],
             'tan_y': [0., 0.454650, 0.417646, -0.307504, -0.040803]}]}
    with open(os.path.join(HERE, 'kat.json'), 'w') as f:
        json.dump(kat, f, indent=1)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(0)
    plan = {'singlet': (7, 60), 'dblgauss': (5, 150), 'triplet': (5, 80), 'rc': (5, 60),
            'cellphone': (3, 100), 'cellphone_even': (3, 100), 'evenasph': (3, 100),
            'zoom52': (3, 80), 'threemir': (7, 300), 'fisheye': (7, 300), 'thin_triplet': (5, 100), 'exotic': (7, 400), 'hybrid': (5, 300), 'diffractive': (7, 500), 'diffractive_wild': (9, 800)}
    # Seeded inputs, independent of which models are selected on the command line: the models of the
    # first generation share ONE stream in plan order (it is advanced for skipped models too); the
    # two fixtures added later (threemir, fisheye) were generated on their own and start a fresh
    # stream.  A full run reproduces every committed file (checked array by array).
    own_stream = ('threemir', 'fisheye')
    only = sys.argv[1:]
    for name, (num, n_wild) in plan.items():
        selected = not only or name in only
        if name in own_stream and not selected:
            continue
        opm = M.OpticalModel.load(os.path.join(HERE, 'models', name + '.json'))
        if not selected:
            wild_rays(opm, n_wild, rng, [])                  # keep the shared stream in step
            continue
        rays = []
        grid_rays(opm, num, 0, rays)
        if name in ('dblgauss', 'rc', 'cellphone'):
            grid_rays(opm, 3, 1, rays)
        if name == 'diffractive':
            grid_rays(opm, 9, 1, rays)
        if name == 'exotic':
            grid_rays(opm, 5, 4, rays)
            grid_rays(opm, 3, 5, rays)
        wild_rays(opm, n_wild, np.random.default_rng(0) if name in own_stream else rng, rays)
        out = trace_all(opm, rays)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
        hist = np.bincount(out['status'], minlength=6)
        print(f'{name:15s} rays={len(rays):5d} status hist={hist.tolist()}')
    write_kat()


if __name__ == '__main__':
    main()
