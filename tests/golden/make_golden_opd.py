#!/usr/bin/env python3
"""Golden OPD vectors tests/golden/vectors/<model>_opd.npz, produced by the
REFERENCE's own wavefront code (build container only):

  rayoptics.raytr.waveabr.calculate_reference_sphere / transfer_to_exit_pupil /
  wave_abr_full_calc  (/root/reference/src/rayoptics/raytr/waveabr.py:24-305)
  on rays traced by rayoptics.raytr.raytrace.trace_raw.

Per model: for every field and wavelength the chief ray (pupil 0,0), its exit
pupil segment and reference sphere, then a 9x9 pupil grid of rays with their OPD
(mm) and transverse aberration.  Stored: the per-tile 24-double records the
engine's epilogue consumes (as the reference's numbers), start rays, OPD, abr.
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                 # noqa: E402
from rayoptics_b200 import model as M                # noqa: E402
from rayoptics_b200.engine import accumulated_steps  # noqa: E402

OUT = os.path.join(HERE, 'vectors')


def main():
    R = rh.ref()
    W = importlib.import_module('rayoptics.raytr.waveabr')
    num = 9
    names = sys.argv[1:] or ('dblgauss', 'rc', 'cellphone', 'triplet', 'telecentric')
    for name in names:
        opm = M.OpticalModel.load(os.path.join(HERE, 'models', name + '.json'))
        osp, sm = opm.optical_spec, opm.seq_model
        fod = types.SimpleNamespace(n_obj=osp.fod.n_obj, n_img=osp.fod.n_img, exp_dist=osp.fod.exp_dist)
        n_ifc = sm.get_num_surfaces()
        kw = dict(first_surf=1, last_surf=n_ifc - 2)
        foc = 0.0
        xs = accumulated_steps(-1.0, 1.0, num)
        recs, rays = [], []
        for fi, fld in enumerate(osp.fov.fields):
            for wi, wvl in enumerate(sm.wvlns):
                path = rh.ref_path(sm, wvl)
                ifc_k = path[-2][0]

                def start(pupil):
                    pt0, dir0 = osp.ray_start_from_osp(pupil, fld, 'rel pupil')
                    if dir0[2]*sm.z_dir[0] < 0:
                        dir0 = -dir0
                    return pt0, dir0

                pt0, dir0 = start(fld.apply_vignetting([0., 0.]))
                cr = R.raytrace.trace_raw(iter(path), pt0, dir0, wvl, **kw)
                cr_exp_seg = W.transfer_to_exit_pupil(ifc_k, (cr[0][-2][0], cr[0][-2][1]), fod.exp_dist)
                crp = (types.SimpleNamespace(ray=cr[0], op=cr[1], wvl=wvl), cr_exp_seg)
                # the functions unpack cr as a 3-tuple *and* use cr.ray: give them a namedtuple
                from rayoptics.raytr import RayPkg
                crp = (RayPkg(*cr), cr_exp_seg)
                ref_sphere = W.calculate_reference_sphere({'seq_model': sm}, fld, wvl, foc, crp)
                image_pt, ref_dir, radius, _ = ref_sphere
                rec = np.zeros(24)
                rec[0:3], rec[3:6] = cr[0][1][0], cr[0][0][1]
                if R.misc_math.is_kinda_big(radius):
                    # infinite-reference variant (waveabr.py:356-420): the chief-ray-only
                    # quantities, with the reference's own expressions (include/b200rt.h layout)
                    rt_, t_ = ref_sphere[3]
                    p_cr_b4, d_cr_b4 = rt_.dot(cr[0][-2][0] - t_), rt_.dot(cr[0][-2][1])
                    op_cr_b4 = W.ray_dist_to_perp_from_origin((p_cr_b4, d_cr_b4))
                    rec[6:9], rec[9:12] = cr[0][-1][0], cr[0][-1][1]
                    rec[12] = cr[1] + op_cr_b4
                    rec[13:16] = image_pt
                    rec[17:20], rec[20], rec[21] = d_cr_b4, t_[2], 0.0
                else:
                    rec[6:9], rec[9:12] = cr[0][-2][0], cr[0][-2][1]
                    rec[12] = cr[1]
                    rec[13:16], rec[16] = cr_exp_seg[0], cr_exp_seg[2]
                    rec[17:20], rec[20] = ref_dir, radius
                    rec[21] = -1.0 if ref_dir[2]*cr[0][-1][1][2] < 0 else 1.0
                rec[22], rec[23] = abs(fod.n_obj), abs(fod.n_img)
                recs.append(rec)
                for i in range(num):
                    for j in range(num):
                        pupil = fld.apply_vignetting(np.array([xs[i], xs[j]]))
                        p0, d0 = start(pupil)
                        r = rh.ref_trace(path, p0, d0, wvl, check_apertures=True, **kw)
                        opd, abr = np.nan, [np.nan, np.nan]
                        if r['status'] == 0:
                            ray = [[s[0:3], s[3:6], s[6], s[7:10]] for s in r['ray']]
                            opd = W.wave_abr_full_calc(fod, fld, wvl, foc, (ray, r['op'], wvl),
                                                       crp, ref_sphere)
                            dist = foc/ray[-1][1][2]
                            abr = (ray[-1][0] + dist*ray[-1][1] - image_pt)[:2]
                        rays.append((p0, d0, wi, len(recs) - 1, r['status'], opd, abr[0], abr[1]))
        n = len(rays)
        out = dict(wave=np.array(recs), p0=np.array([r[0] for r in rays]).T,
                   d0=np.array([r[1] for r in rays]).T,
                   wvl_idx=np.array([r[2] for r in rays], np.int32),
                   tile=np.array([r[3] for r in rays], np.int32),
                   status=np.array([r[4] for r in rays], np.int32),
                   opd=np.array([r[5] for r in rays]), abr=np.array([[r[6], r[7]] for r in rays]).T,
                   num=np.array(num))
        np.savez_compressed(os.path.join(OUT, name + '_opd.npz'), **out)
        ok = out['status'] == 0
        print(f'{name:10s} tiles={len(recs)} rays={n} ok={ok.sum()} '
              f'|opd| max={np.nanmax(np.abs(out["opd"])):.3e} mm')


if __name__ == '__main__':
    main()
