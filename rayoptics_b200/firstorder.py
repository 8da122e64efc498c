"""Paraxial first-order data needed to launch real rays.

Host-side, runs once per model.  Provides the few quantities the start-ray
generation reads from the reference's ``FirstOrderData``
(/root/reference/src/rayoptics/parax/firstorder.py:277-479): ``obj_dist``,
``enp_dist``, ``enp_radius``, ``exp_dist``, ``img_dist``, ``n_obj``, ``n_img``,
``efl``, ``m``, ``red`` and the object-space chief / marginal ray slopes.

This is NOT on the parity path: oracle and engine both receive the numbers
computed here (SURVEY.md 8(d) "model-side inputs").  It is an independent y-nu
formulation (2x2 system matrix from surface 1 to the image), not a restatement
of the reference's routine.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np


def _signed_indices(sm, wi):
    """index following each interface, negative while travelling in -z"""
    ns = []
    for i, g in enumerate(sm.gaps):
        n = sm.rndx[i][wi]
        ns.append(n if sm.z_dir[i] > 0 else -n)
    return ns


def _matrix(sm, wi, first, last):
    """y-nu system matrix taking (y, nu) just before interface `first` to just
    before interface `last` (i.e. including refraction at first..last-1 and the
    transfers between)."""
    ns = _signed_indices(sm, wi)
    M = np.identity(2)
    for i in range(first, last):
        ifc = sm.ifcs[i]
        n_b, n_a = ns[i - 1], ns[i]
        if type(ifc).__name__ == 'ThinLens':
            pwr = ifc.optical_power                 # oprops/thinlens.py:93-95
        elif ifc.interact_mode in ('transmit', 'reflect'):
            pwr = (n_a - n_b)*ifc.profile.cv
        else:
            pwr, n_a = 0.0, n_b
        R = np.array([[1.0, 0.0], [-pwr, 1.0]])
        T = np.array([[1.0, sm.gaps[i].thi/n_a], [0.0, 1.0]])
        M = T @ R @ M
    return M


def compute_first_order(sm, osp, wvl=None):
    """Return a namespace of first-order properties for model ``sm`` under the
    optical specification ``osp`` (pupil / field definitions)."""
    wvl = sm.central_wavelength() if wvl is None else wvl
    wi = sm.index_for_wavelength(wvl)
    ns = _signed_indices(sm, wi)
    n_0, n_k = ns[0], ns[-1]
    n_ifc = len(sm.ifcs)
    img = n_ifc - 1
    thi0 = sm.gaps[0].thi

    # surface 1 -> image system matrix, and surface 1 -> stop
    M1k = _matrix(sm, wi, 1, img)
    stop = sm.stop_surface
    if stop is None:
        stop = 1
    M1s = _matrix(sm, wi, 1, stop)
    # chief ray through the stop centre: y_s = A y1 + B nu1 = 0
    A, B = M1s[0, 0], M1s[0, 1]
    # entrance pupil: object-space ray (y1, u1) with y_s = 0 crosses the axis at
    # distance enp_dist from surface 1: y1 + enp_dist*u1 = 0
    #   A*y1 + B*n_0*u1 = 0  ->  y1/u1 = -B*n_0/A  ->  enp_dist = -y1/u1 = B*n_0/A
    enp_dist = (B*n_0/A) if A != 0.0 else 0.0
    obj2enp = thi0 + enp_dist

    fod = SimpleNamespace()
    fod.obj_dist = thi0
    fod.enp_dist = enp_dist
    fod.n_obj, fod.n_img = n_0, n_k
    # power / reduction from the matrix *without* the last transfer
    Mlast = _matrix(sm, wi, 1, img - 1) if img - 1 >= 1 else np.identity(2)
    Rl = sm.ifcs[img - 1]
    n_b, n_a = ns[img - 2] if img - 2 >= 0 else n_0, ns[img - 1]
    if type(Rl).__name__ == 'ThinLens':
        pwr_l = Rl.optical_power
    else:
        pwr_l = (n_a - n_b)*Rl.profile.cv if Rl.interact_mode in ('transmit', 'reflect') else 0.0
    Mk = np.array([[1.0, 0.0], [-pwr_l, 1.0]]) @ Mlast   # to just after the last powered surface
    ck1, dk1, ak1 = Mk[1, 0], Mk[1, 1], Mk[0, 0]
    fod.power = -ck1
    fod.efl = n_k/fod.power if fod.power != 0.0 else 0.0
    fod.red = dk1 + thi0*ck1

    # ---- marginal (axial) ray from the pupil specification
    pupil_oi, pupil_key = osp.pupil.key
    pv = osp.pupil.value
    if pupil_oi == 'object':
        if pupil_key == 'epd':
            slp0 = 0.5*pv/obj2enp
        elif pupil_key == 'f/#':
            slp0 = -1.0/(2.0*pv)
        elif pupil_key == 'NA':
            slp0 = pv/n_0
        else:
            raise ValueError(f'pupil key {pupil_key}')
    else:
        if pupil_key == 'f/#':
            slpk = -1.0/(2.0*pv)
        elif pupil_key == 'NA':
            slpk = pv/n_k
        else:
            raise ValueError(f'pupil key {pupil_key}')
        slp0 = slpk/fod.red
    # axial ray at surface 1: from the axial object point
    y1 = thi0*slp0
    # height of the marginal ray in the entrance pupil plane
    fod.enp_radius = abs(slp0*obj2enp)
    yk, nuk = M1k @ np.array([y1, n_0*slp0])
    uk = nuk/n_k
    fod.img_dist = sm.gaps[-1].thi - (yk/uk if uk != 0.0 else 0.0)
    fod.ax_slp0 = slp0
    fod.fno = -1.0/(2.0*n_k*uk) if uk != 0.0 else 1e10
    fod.m = (n_0*slp0)/(n_k*uk) if uk != 0.0 else 0.0
    fod.obj_na = n_0*sm.z_dir[0]*slp0
    fod.img_na = n_k*sm.z_dir[-1]*uk

    # ---- chief ray of the maximum field
    fov_oi, fov_key = osp.fov.key
    fv = osp.fov.max_field_value()
    if fov_oi == 'object':
        if fov_key == 'angle':
            slpbar0 = math.tan(math.radians(fv))
            ybar0 = -slpbar0*obj2enp
        else:   # height
            ybar0 = fv
            slpbar0 = -ybar0/obj2enp
    else:
        # image height (or angle): scale the unit chief ray through the stop centre
        ybar1_u, ubar1_u = -enp_dist, 1.0          # unit-slope chief ray at surface 1
        yk_u, nuk_u = M1k @ np.array([ybar1_u, n_0*ubar1_u])
        if fov_key == 'height' or fov_key == 'real height':
            scale = fv/yk_u if yk_u != 0.0 else 0.0
        else:
            scale = math.tan(math.radians(fv))/(nuk_u/n_k) if nuk_u != 0.0 else 0.0
        slpbar0 = scale*ubar1_u
        ybar0 = -slpbar0*obj2enp
    fod.pr_slp0 = slpbar0
    fod.pr_ht0 = ybar0
    ybar1 = -enp_dist*slpbar0
    ybk, nubk = M1k @ np.array([ybar1, n_0*slpbar0])
    ubk = nubk/n_k
    # exit pupil distance measured from the last interface before the image
    fod.exp_dist = (sm.gaps[-1].thi - ybk/ubk) if ubk != 0.0 else -1e10
    fod.img_ht = ybk
    fod.opt_inv = n_0*(y1*slpbar0 - ybar1*slp0)
    return fod
