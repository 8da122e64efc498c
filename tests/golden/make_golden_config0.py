#!/usr/bin/env python3
"""BASELINE configs[0] -- "singlet lens, 1 field, 1 wvl, 7x7 pupil grid on reference CPU path
(plumbing)": the reference's own grid loop `rayoptics.raytr.trace.trace_grid`
(/root/reference/src/rayoptics/raytr/trace.py:563-605) on the hybrid model of oracle/ref_model.py
(reference Surface objects, reference trace_raw).  Stored per ray, in the reference's x-outer /
y-inner order: recorded pupil coordinates, status (0 ok / 3 blocked ...), image-plane intercept
p, direction d and op_delta -> tests/golden/vectors/singlet_config0.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_model, ref_harness as rh       # noqa: E402
from rayoptics_b200 import model as M                 # noqa: E402

NUM = 7


def main():
    RT, RA = ref_model.modules()
    opm = M.OpticalModel.load(os.path.join(HERE, 'models', 'singlet.json'))
    H = ref_model.HybridModel(opm)
    fld = opm.optical_spec.field_of_view.fields[0]
    wvl = opm.seq_model.central_wavelength()
    rows = []

    def record(pupil, pkg_or_err):
        rows.append((np.array(pupil), pkg_or_err))
        return 0.0

    grid_def = [np.array([-1., -1.]), np.array([1., 1.]), NUM]
    RT.trace_grid(H, grid_def, fld, wvl, 0.0, img_filter=record, form='list')
    n = len(rows)
    out = dict(pupil=np.zeros((n, 2)), status=np.zeros(n, np.int32), p=np.full((n, 3), np.nan),
               d=np.full((n, 3), np.nan), op=np.full(n, np.nan), num=np.array(NUM), wvl=np.array(wvl))
    for k, (pupil, pkg) in enumerate(rows):
        out['pupil'][k] = pupil
        if pkg is None:                          # trace_safe hands errors to img_filter as None
            out['status'][k] = -1
            continue
        ray, op, _ = pkg
        out['p'][k], out['d'][k], out['op'][k] = ray[-1][0], ray[-1][1], op
    # statuses of the failed rays from a second pass that keeps the errors
    for k, (pupil, pkg) in enumerate(rows):
        if pkg is None:
            raw = [np.array([-1., -1.]), np.array([1., 1.]), NUM]
            i, j = divmod(k, NUM)
            start = np.array(raw[0])
            step = (raw[1] - raw[0])/(NUM - 1)
            for _ in range(i):
                start[0] += step[0]
            for _ in range(j):
                start[1] += step[1]
            rr = RT.trace_safe(H, np.array(start), fld, wvl, None, 'full', check_apertures=True)
            out['status'][k] = rh.STATUS[type(rr.err).__name__]
    np.savez_compressed(os.path.join(HERE, 'vectors', 'singlet_config0.npz'), **out)
    print('singlet 7x7: status hist', np.bincount(out['status'], minlength=4))


if __name__ == '__main__':
    main()
