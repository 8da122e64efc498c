#!/usr/bin/env python3
"""The five BASELINE configurations at reduced pupil sampling, traced by the REFERENCE:
for every field and wavelength the body of `rayoptics.raytr.trace.trace_grid`
(/root/reference/src/rayoptics/raytr/trace.py:563-605: accumulated pupil stepping, x outer / y
inner, `trace_safe(..., check_apertures=True)` -> `trace_base` -> `ray_start_from_osp` ->
`rt.trace`) on the hybrid model of oracle/ref_model.py (reference Surface objects, reference
trace_raw).  Stored in the engine's ray order (field, wavelength, i, j): status, image-plane
intercept p, direction d, op_delta -> tests/golden/vectors/<model>_grid.npz.

The chief-ray re-aiming of the reference is not involved (trace_safe does not aim).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_model, ref_harness as rh       # noqa: E402
from rayoptics_b200 import model as M                 # noqa: E402

# model -> pupil samples per axis (BASELINE: 7, 512, 1024, 256, 1024)
CONFIGS = {'singlet': 7, 'dblgauss': 24, 'rc': 24, 'evenasph': 12, 'cellphone': 10, 'zoom52': 8,
           'fisheye': 16, 'threemir': 12}


def main():
    RT, RA = ref_model.modules()
    for name in (sys.argv[1:] or CONFIGS):
        num = CONFIGS[name]
        opm = M.OpticalModel.load(os.path.join(HERE, 'models', name + '.json'))
        H = ref_model.HybridModel(opm)
        sm, osp = opm.seq_model, opm.optical_spec
        fields, wvls = osp.field_of_view.fields, sm.wvlns
        n = len(fields)*len(wvls)*num*num
        status = np.zeros(n, np.int32)
        p, d, op = np.full((n, 3), np.nan), np.full((n, 3), np.nan), np.full(n, np.nan)
        k = 0
        for fld in fields:
            for wvl in wvls:
                start = np.array([-1., -1.])
                stop = np.array([1., 1.])
                step = np.array((stop - start)/(num - 1))
                for i in range(num):
                    for j in range(num):
                        pupil = np.array(start)
                        rr = RT.trace_safe(H, pupil, fld, wvl, None, 'summary', check_apertures=True)
                        if rr.err is not None:
                            status[k] = rh.STATUS[type(rr.err).__name__]
                        else:
                            ray, opd, _ = rr.pkg
                            p[k], d[k], op[k] = ray[-1][0], ray[-1][1], opd
                        k += 1
                        start[1] += step[1]
                    start[0] += step[0]
                    start[1] = -1.
        np.savez_compressed(os.path.join(HERE, 'vectors', name + '_grid.npz'), num=np.array(num),
                            status=status, p=p, d=d, op=op)
        print(f'{name:10s} {len(fields)} fields x {len(wvls)} wvls x {num}x{num}: status hist',
              np.bincount(status, minlength=4))


if __name__ == '__main__':
    main()
