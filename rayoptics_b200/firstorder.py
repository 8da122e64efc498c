"""Paraxial first-order data needed to launch real rays.

Host-side, runs once per model.  A restatement of the reference's y-u paraxial trace and
``compute_first_order`` (/root/reference/src/rayoptics/parax/firstorder.py:132-214,277-479,
482-548) with its order of operations, so that the numbers the start-ray generation reads
(``obj_dist``, ``enp_dist``, ``enp_radius``, ``exp_dist``, ``n_obj``, ``n_img``, ``red``, ``m``,
the object-space chief-ray data ...) are the reference's, bit for bit -- including the
~1e-6 relative round-off its entrance pupil distance carries for objects at 1e10 mm
(the chief ray is traced from the object, so ``enp_dist`` is the small difference of
1e11-sized numbers).  tests/test_firstorder_vs_reference.py runs the reference's own
``compute_first_order`` (importable) on shims of the same model and compares every field.

Surface powers follow ``SequentialModel.update_model`` (seq/sequential.py:612-658):
``delta_n`` is the difference of the signed indices (negative while travelling in -z) at the
central wavelength and ``optical_power = delta_n * cv``; a thin lens contributes its own
power (oprops/thinlens.py:73-99).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

HT, SLP, AOI = 0, 1, 2          # optical/model_constants.py: ht, slp, aoi


def _parax_path(sm, wvl):
    """[(profile_cv, optical_power, interact_mode, thi | None, rndx | None, z_dir | None)] of the
    sequence at ``wvl``; powers from the CENTRAL wavelength (they are set by update_model)."""
    wi = sm.index_for_wavelength(wvl)
    ci = sm.ref_wvl
    n_gaps = len(sm.gaps)
    n_before = sm.rndx[0][ci]
    z_dir_before = sm.z_dir[0]
    out = []
    for i, ifc in enumerate(sm.ifcs):
        thin = type(ifc).__name__ == 'ThinLens'
        z_dir_after = int(math.copysign(1, z_dir_before))
        if ifc.interact_mode == 'reflect':
            z_dir_after = -z_dir_after
        delta_n = 0.0
        if i < n_gaps:
            n_after = sm.rndx[i][ci]
            if z_dir_after < 0:
                n_after = -n_after
            delta_n = n_after - n_before
            n_before = n_after
            z_dir_before = z_dir_after
        cv = ifc.optical_power if thin else ifc.profile.cv
        pwr = ifc.optical_power if thin else delta_n*ifc.profile.cv
        if i < n_gaps:
            out.append((cv, pwr, ifc.interact_mode, sm.gaps[i].thi, sm.rndx[i][wi], sm.z_dir[i]))
        else:
            out.append((cv, pwr, ifc.interact_mode, None, None, None))
    return out


def paraxial_trace(path, start, start_yu, start_yu_bar):
    """perform a paraxial raytrace of 2 linearly independent rays (firstorder.py:132-214)"""
    it = iter(path)
    p_ray, p_ray_bar = [], []
    b4_cv, _, _, b4_thi, b4_rndx, z_dir_before = next(it)
    n_before = b4_rndx if z_dir_before > 0 else -b4_rndx
    b4_yui, b4_yui_bar = list(start_yu), list(start_yu_bar)
    if start == 1:
        t0 = b4_thi
        if np.isinf(t0):
            obj_ht, obj_htb = 0., -np.inf
        else:
            obj_ht = start_yu[HT] - t0*start_yu[SLP]
            obj_htb = start_yu_bar[HT] - t0*start_yu_bar[SLP]
        b4_yui = [obj_ht, start_yu[SLP]]
        b4_yui_bar = [obj_htb, start_yu_bar[SLP]]
    b4_yui.append(b4_yui[SLP] + b4_yui[HT]*b4_cv)
    b4_yui_bar.append(b4_yui_bar[SLP] + b4_yui_bar[HT]*b4_cv)
    p_ray.append(b4_yui)
    p_ray_bar.append(b4_yui_bar)
    for cv, pwr, mode, thi, rndx, z_dir_after in it:
        if rndx is None:
            rndx = abs(n_before)
        if z_dir_after is None:
            z_dir_after = z_dir_before
        t = b4_thi
        cur_ht = b4_yui[HT] + t*b4_yui[SLP]
        cur_htb = b4_yui_bar[HT] + t*b4_yui_bar[SLP]
        if mode == 'dummy' or mode == 'phantom':
            cur_slp, cur_slpb = b4_yui[SLP], b4_yui_bar[SLP]
        else:
            n_after = rndx if z_dir_after > 0 else -rndx
            k = n_before/n_after
            cur_slp = k*b4_yui[SLP] - cur_ht*pwr/n_after
            cur_slpb = k*b4_yui_bar[SLP] - cur_htb*pwr/n_after
            n_before = n_after
            z_dir_before = z_dir_after
        yu = [cur_ht, cur_slp, cur_slp + cur_ht*cv]
        yu_bar = [cur_htb, cur_slpb, cur_slpb + cur_htb*cv]
        p_ray.append(yu)
        p_ray_bar.append(yu_bar)
        b4_yui, b4_yui_bar, b4_thi = yu, yu_bar, thi
    return p_ray, p_ray_bar


def _na2slp(na, n=1.0):
    return n*math.tan(math.asin(na/n))          # parax/etendue.py:27-29


def compute_first_order(sm, osp, wvl=None):
    """Return a namespace of first-order properties (the reference's ``FirstOrderData``
    fields plus ``pr_slp0`` / ``pr_ht0`` / ``ax_slp0``, and the paraxial rays ``ax_ray``,
    ``pr_ray``) for model ``sm`` under the optical specification ``osp``."""
    wvl = sm.central_wavelength() if wvl is None else wvl
    ci = sm.ref_wvl
    stop = sm.stop_surface
    path = _parax_path(sm, wvl)
    oal = 0
    for g in sm.gaps[1:-1]:
        oal += g.thi
    n_0 = sm.z_dir[0]*sm.rndx[0][ci]
    n_k = sm.z_dir[-1]*sm.rndx[len(sm.gaps) - 1][ci]
    # compute_principle_points, firstorder.py:482-548
    p_ray, q_ray = paraxial_trace(path, 1, [1., 0.], [0., 1/n_0])
    img = -2 if len(sm.ifcs) > 2 else -1
    ak1, bk1 = p_ray[img][HT], q_ray[img][HT]
    ck1, dk1 = n_k*p_ray[img][SLP], n_k*q_ray[img][SLP]
    Mk1 = np.array([[ak1, bk1], [ck1, dk1]])
    M1k = np.array([[dk1, -bk1], [-ck1, ak1]])

    if stop is None:                      # nothing pre-computed: assume the 1st surface
        enp_dist = 0.0
        eff_stop = None
        if _is_fuzzy_zero(sm.gaps[0].thi):
            for i, g in enumerate(sm.gaps):
                if not _is_fuzzy_zero(g.thi):
                    eff_stop = i + 1
                    enp_dist += g.thi
                    break
        else:
            eff_stop = 1
    else:
        eff_stop = stop
    if eff_stop is not None:
        as1, bs1 = p_ray[eff_stop][HT], q_ray[eff_stop][HT]
        ybar1, ubar1 = -bs1, as1
        n_0 = sm.gaps[0].medium.rindex(wvl)
        enp_dist = -ybar1/(n_0*ubar1)
    else:
        as1, bs1 = p_ray[1][HT], q_ray[1][HT]
        ybar1, ubar1 = -bs1, as1

    thi0 = sm.gaps[0].thi
    red = dk1 + thi0*ck1
    obj2enp_dist = thi0 + enp_dist

    pupil_oi_key, pupil_value_key = osp.pupil.key
    pv = osp.pupil.value
    n_obj_u, n_img_u = sm.rndx[0][ci], sm.rndx[len(sm.gaps) - 1][ci]     # obj_img_rindex()
    if 'NA' in pupil_value_key:
        pupil_key, pupil_value = 'slope', _na2slp(pv, n=(n_obj_u if pupil_oi_key == 'object' else n_img_u))
    elif 'f/#' in pupil_value_key:
        pupil_key, pupil_value = 'slope', -1/(2*pv)
    else:
        pupil_key, pupil_value = 'height', pv/2
    # PupilSpec.derive_parax_params (opticalspec.py:619-643) has turned the specification
    # into a paraxial height or slope; firstorder.py:353-375
    if pupil_oi_key == 'object':
        slp0 = pupil_value/obj2enp_dist if pupil_key == 'height' else pupil_value
    else:
        slpk = pupil_value/obj2enp_dist if pupil_key == 'height' else pupil_value
        slp0 = slpk/red
    yu = [0., slp0]

    fov_oi_key, fov_value_key = osp.field_of_view.key
    fov_value = osp.field_of_view.value if osp.field_of_view.value != 0 else 1.
    if 'angle' in fov_value_key:
        field_key, field_value = 'slope', math.tan(math.radians(fov_value))    # etendue.ang2slp
    else:
        field_key, field_value = 'height', fov_value
    if fov_oi_key == 'object':
        if field_key == 'slope':
            slpbar0 = field_value
            ybar0 = -slpbar0*obj2enp_dist
        else:
            ybar0 = field_value
            slpbar0 = -ybar0/obj2enp_dist
    else:
        aki, bki = p_ray[-1][HT], q_ray[-1][HT]
        cki, dki = n_k*p_ray[-1][SLP], n_k*q_ray[-1][SLP]
        M1i = np.array([[dki, -bki], [-cki, aki]])
        q_ray_k = np.matmul(Mk1, np.array([ybar1, ubar1]))
        exp_dist_i = -q_ray_k[HT]/(n_k*q_ray_k[SLP])
        img2exp_dist = exp_dist_i - sm.gaps[-1].thi
        if field_key == 'height':
            ht_i = field_value
            pr_ray_1 = np.matmul(M1i, np.array([ht_i, -ht_i/img2exp_dist]))
        else:
            slp_k = field_value
            pr_ray_1 = np.matmul(M1k, np.array([-slp_k*exp_dist_i, slp_k]))
        slpbar0 = pr_ray_1[SLP]
        ybar0 = -slpbar0*obj2enp_dist
    yu_bar = [ybar0, slpbar0]

    ax_ray, pr_ray = paraxial_trace(path, 0, yu, yu_bar)
    opt_inv = n_0*(ax_ray[1][HT]*pr_ray[0][SLP] - pr_ray[1][HT]*ax_ray[0][SLP])

    fod = SimpleNamespace()
    fod.opt_inv = opt_inv
    fod.obj_dist = obj_dist = sm.gaps[0].thi
    if ck1 == 0.0:
        fod.img_dist = img_dist = 1e10
        fod.power = 0.0
        fod.fl_obj = fl_obj = 0.0
        fod.fl_img = fl_img = 0.0
        fod.efl = 0.0
        fod.pp1 = 0.0
        fod.ppk = 0.0
    else:
        if ax_ray[img][SLP] != 0:
            fod.img_dist = img_dist = -ax_ray[img][HT]/ax_ray[img][SLP]
        else:
            fod.img_dist = img_dist = math.copysign(1e10, sm.gaps[-1].thi)
        fod.power = power = -ck1
        fod.fl_obj = fl_obj = n_0/power
        fod.fl_img = fl_img = n_k/power
        fod.efl = fl_img
        fod.pp1 = (1.0 - dk1)*(fl_obj)
        fod.ppk = (ak1 - 1.0)*(fl_img)
    fod.ffl = fod.pp1 + (-fl_obj)
    fod.bfl = fod.ppk + fl_img
    fod.pp_sep = oal - fod.pp1 + fod.ppk
    if ax_ray[img][SLP] != 0:
        fod.fno = -1.0/(2.0*n_k*ax_ray[-1][SLP])
        fod.img_ht = -fod.opt_inv/(n_k*ax_ray[-1][SLP])
    else:
        fod.fno = 1e10
        fod.img_ht = 1e10
    fod.m = ak1 + ck1*img_dist/n_k
    fod.red = dk1 + ck1*obj_dist
    fod.n_obj = n_0
    fod.n_img = n_k
    fod.obj_ang = math.degrees(math.atan(pr_ray[0][SLP]))
    if pr_ray[0][SLP] != 0:
        nu_pr0 = n_0*pr_ray[0][SLP]
        fod.enp_dist = -pr_ray[1][HT]/nu_pr0
        fod.enp_radius = abs(fod.opt_inv/nu_pr0)
    else:
        fod.enp_dist = -1e10
        fod.enp_radius = 1e10
    if pr_ray[-1][SLP] != 0:
        fod.exp_dist = -(pr_ray[-1][HT]/pr_ray[-1][SLP] - fod.img_dist)
        fod.exp_radius = abs(fod.opt_inv/(n_k*pr_ray[-1][SLP]))
    else:
        fod.exp_dist = -1e10
        fod.exp_radius = 1e10
    fod.obj_na = n_0*sm.z_dir[0]*ax_ray[0][SLP]
    fod.img_na = n_k*sm.z_dir[-1]*ax_ray[-1][SLP]
    # the object-space rays, as the start-ray code reads them from parax_data (pr[0], ax[0])
    fod.pr_ht0, fod.pr_slp0 = pr_ray[0][HT], pr_ray[0][SLP]
    fod.ax_slp0 = ax_ray[0][SLP]
    fod.ax_ray, fod.pr_ray = ax_ray, pr_ray
    return fod


def _is_fuzzy_zero(x, fuzz=1e-10):
    return abs(x) < fuzz
