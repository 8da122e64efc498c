#!/usr/bin/env python3
"""Golden analysis results tests/golden/vectors/<model>_analyses.npz produced by the REFERENCE's
own analysis classes (build container only): rayoptics.raytr.analyses.RayFan / RayList / RayGrid
(/root/reference/src/rayoptics/raytr/analyses.py:121-187,343-434,584-663) run on a hybrid model
(oracle/ref_model.py) with the reference's trace_raw.

The reference's chief-ray RE-AIMING is bypassed: `get_chief_ray_pkg` (raytr/trace.py:660-687) calls
`aim_chief_ray` whenever `fld.chief_ray is None`, and at this commit `iterate_ray` starts from the
object point of `obj_coords`, which for infinite conjugates is the mirror image of the point
`ray_start_from_osp` launches from -- the aim point comes back with the wrong sign and the chief
ray misses the stop centre (dblgauss 10 deg: +0.1522 instead of -0.1522, 0.19 mm off; shown by
tests/test_trace_drivers.py::test_iterate_ray_is_the_references and DESIGN.md).  The fixtures'
aim points (chief ray through the stop centre, as stored in the reference's .roa files) are kept
by giving every field a placeholder `chief_ray` of another wavelength, which makes
`get_chief_ray_pkg` re-trace the chief ray without re-aiming.

Per model and (field, wavelength) case: the y- and x-fan (pupil coordinates, dx, dy, OPD in
waves), the default ray list (transverse aberrations of the rays that reach the image) and the
wavefront grid [3, n, n].
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_model                        # noqa: E402
from rayoptics_b200 import model as M               # noqa: E402

OUT = os.path.join(HERE, 'vectors')
CASES = {'dblgauss': [(0, 587.6), (1, 656.3), (2, 486.1)], 'rc': [(0, 550.0), (3, 550.0)],
         'triplet': [(1, None)], 'telecentric': [(0, 587.6), (2, 486.1)], 'cellphone': [(4, None)],
         'fisheye': [(0, 587.6), (2, 656.3), (3, 587.6)], 'threemir': [(1, None), (4, None)]}
NUM_FAN, NUM_LIST, NUM_GRID = 21, 15, 16
PSF_DIM = 64


def main():
    RT, RA = ref_model.modules()
    for name in (sys.argv[1:] or CASES):
        opm = M.OpticalModel.load(os.path.join(HERE, 'models', name + '.json'))
        H = ref_model.HybridModel(opm)
        out = {'cases': np.array([[f, -1.0 if w is None else w] for f, w in CASES[name]]),
               'num': np.array([NUM_FAN, NUM_LIST, NUM_GRID]), 'psf_dim': np.array(PSF_DIM)}
        for ci, (f, wl) in enumerate(CASES[name]):
            fld = opm.optical_spec.field_of_view.fields[f]
            aim = None if fld.aim_info is None else np.array(fld.aim_info)

            def keep_aim():
                fld.aim_info = None if aim is None else (float(aim) if aim.ndim == 0 else aim.copy())
                fld.chief_ray = ((None, None, -1.0), None)     # != wvl: re-trace, do not re-aim

            for xy in 'xy':
                keep_aim()
                fan = RA.RayFan(H, f=f, wl=wl, xyfan=xy, num_rays=NUM_FAN)
                out[f'fan{xy}_pupil_{ci}'] = np.array([p for p, v in fan.fan], dtype=float).reshape(-1, 2)
                out[f'fan{xy}_vals_{ci}'] = np.array([v for p, v in fan.fan], dtype=float).reshape(-1, 3)
            keep_aim()
            rl = RA.RayList(H, num_rays=NUM_LIST, f=f, wl=wl)
            out[f'list_abr_{ci}'] = np.array(rl.ray_abr, dtype=float)
            keep_aim()
            rg = RA.RayGrid(H, f=f, wl=wl, num_rays=NUM_GRID)
            out[f'grid_{ci}'] = np.array(rg.grid, dtype=float)
            out[f'psf_{ci}'] = RA.calc_psf(rg.grid[2], NUM_GRID, PSF_DIM)     # analyses.py:848-875
        np.savez_compressed(os.path.join(OUT, name + '_analyses.npz'), **out)
        print(f'{name:12s} cases={len(CASES[name])} fan rays={out["fany_vals_0"].shape[0]} '
              f'list rays={out["list_abr_0"].shape[1]} grid ok={np.isfinite(out["grid_0"][2]).sum()}')


if __name__ == '__main__':
    main()
