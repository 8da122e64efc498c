"""Chief-ray aiming and clear-aperture setting on top of a ``trace`` function.

Host-side model preparation (tens of rays per field, run once per model):

* ``aim_chief_ray``       /root/reference/src/rayoptics/raytr/trace.py:313-415,627-640
  (``iterate_ray``: find the entrance-pupil aim point that puts the ray through
  the centre of the stop surface)
* ``set_clear_apertures`` /root/reference/src/rayoptics/raytr/vigcalc.py:45-80
  (max radial height of the 5 boundary rays per field -> ``max_aperture``;
  the stop surface is set from the first field only)

Both take the tracer as an argument (``trace_fn(seq_model, pt0, dir0, wvl,
**kw) -> (ray, op, wvl)``): the product passes the GPU drop-in
``rayoptics_b200.raytrace.trace``; the fixture generator, which runs where no
GPU exists, passes the reference's own ``trace``.  Not on the parity path --
their results are inputs shared by oracle and engine.
"""
from __future__ import annotations

import math

import numpy as np


def _stop_xy(opt_model, trace_fn, fld, wvl, aim, stop):
    osp = opt_model.optical_spec
    saved = fld.aim_info
    fld.aim_info = np.array(aim, dtype=float)
    try:
        pt0, dir0 = osp.ray_start_from_osp(np.array([0., 0.]), fld, 'rel pupil')
    finally:
        fld.aim_info = saved
    sm = opt_model.seq_model
    if dir0[2]*sm.z_dir[0] < 0:
        dir0 = -dir0
    ray, _, _ = trace_fn(sm, pt0, dir0, wvl)
    p = ray[stop][0]
    return np.array([p[0], p[1]])


def aim_chief_ray(opt_model, fld, wvl, trace_fn, tol=1e-13, max_iter=30):
    """Aim point on the paraxial entrance pupil such that the (0, 0) pupil ray
    crosses the stop surface at its vertex.  Damped 2-D Newton iteration."""
    sm = opt_model.seq_model
    stop = sm.stop_surface
    if stop is None:
        return np.array([0., 0.])
    x = np.array([0., 0.])

    def f_at(v):
        try:
            return _stop_xy(opt_model, trace_fn, fld, wvl, v, stop)
        except Exception:       # TraceError: the trial ray did not reach the stop
            return None

    f = f_at(x)
    if f is None:
        return x
    h = 1e-4*max(1.0, opt_model.optical_spec.fod.enp_radius)
    for _ in range(max_iter):
        if np.max(np.abs(f)) < tol:
            break
        J = np.zeros((2, 2))
        ok = True
        for k in range(2):
            dx = np.zeros(2)
            dx[k] = h
            fk = f_at(x + dx)
            if fk is None:
                ok = False
                break
            J[:, k] = (fk - f)/h
        if not ok:
            break
        try:
            step = np.linalg.solve(J, -f)
        except np.linalg.LinAlgError:
            break
        # backtracking: halve the step until the trial ray traces and improves
        lam, accepted = 1.0, False
        for _bt in range(20):
            f_new = f_at(x + lam*step)
            if f_new is not None and np.max(np.abs(f_new)) < np.max(np.abs(f)):
                x, f, accepted = x + lam*step, f_new, True
                break
            lam *= 0.5
        if not accepted:
            break
    if fld.x == 0.0:
        x[0] = 0.0
    return x


def aim_all_fields(opt_model, trace_fn, wvl=None):
    osp = opt_model.optical_spec
    wvl = osp.spectral_region.central_wvl if wvl is None else wvl
    for fld in osp.field_of_view.fields:
        fld.aim_info = aim_chief_ray(opt_model, fld, wvl, trace_fn)


def trace_boundary_rays(opt_model, trace_fn, wvl=None):
    """raytr/trace.py:460-510: the pupil_rays of every field, with vignetting."""
    osp = opt_model.optical_spec
    sm = opt_model.seq_model
    wvl = osp.spectral_region.central_wvl if wvl is None else wvl
    rayset = []
    for fld in osp.field_of_view.fields:
        rim = []
        for pr in osp.pupil.pupil_rays:
            pupil = fld.apply_vignetting(list(pr))
            pt0, dir0 = osp.ray_start_from_osp(pupil, fld, 'rel pupil')
            if dir0[2]*sm.z_dir[0] < 0:
                dir0 = -dir0
            try:
                ray, _, _ = trace_fn(sm, pt0, dir0, wvl)
            except Exception as e:       # TraceError: keep the partial ray
                pkg = getattr(e, 'ray_pkg', None)
                ray = pkg[0] if pkg is not None else []
            rim.append(ray)
        rayset.append(rim)
    return rayset


def set_clear_apertures(opt_model, trace_fn, wvl=None):
    sm = opt_model.seq_model
    rayset = trace_boundary_rays(opt_model, trace_fn, wvl)
    n = sm.get_num_surfaces()
    stop = sm.stop_surface

    def max_ap(fields, i):
        m = None
        for rim in fields:
            for ray in rim:
                if len(ray) > i:
                    p = ray[i][0]
                    ap = math.sqrt(p[0]*p[0] + p[1]*p[1])
                    m = ap if m is None or ap > m else m
                else:
                    return None
        return m

    for i in range(n):
        m = max_ap([rayset[0]], i) if i == stop else max_ap(rayset, i)
        if m is not None:
            sm.ifcs[i].set_max_aperture(m)
