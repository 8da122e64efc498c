"""Host-side sequential model: the data the ray-trace engine is compiled from.

This is a *data* mirror of the parts of the reference's model layer that the
hot path reads -- it holds numbers, it does not trace rays:

* profile classes carry the attributes of ``rayoptics.elem.profiles``
  (``Spherical`` :218, ``Conic`` :449, ``EvenPolynomial`` :682,
  ``RadialPolynomial`` :891, ``YToroid`` :1119, ``XToroid`` :1375 of
  /root/reference/src/rayoptics/elem/profiles.py) under the *same class
  names*, because the surface-table builder (``table.py``) dispatches on
  ``type(profile).__name__`` and therefore accepts the reference's own objects
  and these mirrors interchangeably;
* ``Surface`` / ``Circular`` / ``Rectangular`` mirror
  /root/reference/src/rayoptics/elem/surface.py:38-271,397-469;
* ``SequentialModel.path(wl)`` yields the reference's path tuples
  ``(Intfc, Gap, Tfrm, Indx, Zdir)`` (seq/sequential.py:149-202,
  optical/model_constants.py:12).

Models are persisted as a small JSON "prescription" (``b200rt-model-v1``);
tools/make_models.py writes them from the reference's bundled lens files.
"""
from __future__ import annotations

import json
import math
from itertools import zip_longest

import numpy as np


# ----------------------------------------------------------------- profiles
class SurfaceProfile:
    """Base of the profile data holders (reference: elem/profiles.py:48)."""

    cv = 0.0

    def update(self):
        return self

    def to_dict(self):
        d = {'type': type(self).__name__}
        d.update({k: (list(v) if isinstance(v, (list, tuple, np.ndarray)) else v)
                  for k, v in vars(self).items() if k != 'max_nonzero_coef'})
        return d


class Spherical(SurfaceProfile):
    def __init__(self, c=0.0, r=None):
        if r is not None:
            c = 1.0/r if r != 0.0 else 0.0
        self.cv = c

    @property
    def r(self):
        return 1.0/self.cv if self.cv != 0.0 else 0.0


class Conic(SurfaceProfile):
    def __init__(self, c=0.0, cc=0.0, r=None, ec=None):
        if r is not None:
            c = 1.0/r if r != 0.0 else 0.0
        self.cv = c
        self.cc = cc if ec is None else ec - 1.0

    @property
    def ec(self):
        # same expression as the reference property (profiles.py:519-521)
        return self.cc + 1.0


def _max_nonzero_coef(coefs):
    k = -1
    for i, c in enumerate(coefs):
        if c != 0.0:
            k = i
    return k + 1


class EvenPolynomial(SurfaceProfile):
    """coefs[i] multiplies r**(2*(i+1)) (profiles.py:849-866)."""

    def __init__(self, c=0.0, cc=0.0, r=None, ec=None, coefs=None):
        if r is not None:
            c = 1.0/r if r != 0.0 else 0.0
        self.cv = c
        self.cc = cc if ec is None else ec - 1.0
        self.coefs = list(coefs) if coefs is not None else []
        self.update()

    @property
    def ec(self):
        return self.cc + 1.0

    def update(self):
        self.max_nonzero_coef = _max_nonzero_coef(self.coefs)
        return self


class RadialPolynomial(SurfaceProfile):
    """coefs[i] multiplies r**(i+1); stores ec, cc is derived (profiles.py:970-976)."""

    def __init__(self, c=0.0, cc=None, r=None, ec=1.0, coefs=None):
        if r is not None:
            c = 1.0/r if r != 0.0 else 0.0
        self.cv = c
        self.ec = ec if cc is None else cc + 1.0
        self.coefs = list(coefs) if coefs is not None else []
        self.update()

    @property
    def cc(self):
        return self.ec - 1.0

    def update(self):
        self.max_nonzero_coef = _max_nonzero_coef(self.coefs)
        return self


class YToroid(SurfaceProfile):
    def __init__(self, c=0.0, cR=0.0, cc=0.0, r=None, rR=None, ec=None, coefs=None):
        if r is not None:
            c = 1.0/r if r != 0.0 else 0.0
        if rR is not None:
            cR = 1.0/rR if rR != 0.0 else 0.0
        self.cv = c
        self.cR = cR
        self.cc = cc if ec is None else ec - 1.0
        self.coefs = list(coefs) if coefs is not None else []
        self.update()

    @property
    def ec(self):
        return self.cc + 1.0

    @property
    def rR(self):
        return 1.0/self.cR if self.cR != 0.0 else 0.0

    def update(self):
        self.max_nonzero_coef = _max_nonzero_coef(self.coefs)
        return self


class XToroid(YToroid):
    pass


_PROFILE_CLASSES = {c.__name__: c for c in
                    (Spherical, Conic, EvenPolynomial, RadialPolynomial, YToroid, XToroid)}


def profile_from_dict(d):
    d = dict(d)
    cls = _PROFILE_CLASSES[d.pop('type')]
    prf = cls.__new__(cls)
    for k, v in d.items():
        setattr(prf, k, v)
    return prf.update()


# ---------------------------------------------------------------- apertures
class Aperture:
    def __init__(self, x_offset=0.0, y_offset=0.0, rotation=0.0, is_obscuration=False):
        self.x_offset = x_offset
        self.y_offset = y_offset
        self.rotation = rotation
        self.is_obscuration = is_obscuration

    def to_dict(self):
        d = {'type': type(self).__name__}
        d.update(vars(self))
        return d


class Circular(Aperture):
    def __init__(self, radius=1.0, **kwargs):
        super().__init__(**kwargs)
        self.radius = radius

    def resize(self, half_size):            # set_dimension(x, y), elem/surface.py:410-411
        self.radius = half_size


class Rectangular(Aperture):
    def __init__(self, x_half_width=1.0, y_half_width=1.0, **kwargs):
        super().__init__(**kwargs)
        self.x_half_width = x_half_width
        self.y_half_width = y_half_width

    def resize(self, half_size):            # set_dimension, elem/surface.py:449-451,487-489
        self.x_half_width = self.y_half_width = abs(half_size)


class Elliptical(Rectangular):
    pass


_APERTURE_CLASSES = {c.__name__: c for c in (Circular, Rectangular, Elliptical)}


def aperture_from_dict(d):
    d = dict(d)
    cls = _APERTURE_CLASSES[d.pop('type')]
    return cls(**d)


# ------------------------------------------------------------------ decenters and tilts
def euler2mat_rxyz(ai, aj, ak):
    """``transforms3d.euler.euler2mat(ai, aj, ak, axes='rxyz')`` (transforms3d 0.4.x, un-vendored
    dependency of the reference, setup.cfg:36): the generic Euler-to-matrix construction of
    Gohlke's transformations.py for the axes tuple 'rxyz' = (firstaxis 2, parity 1, repetition 0,
    frame 1), restated from the published algorithm.  Angles in radians."""
    i, j, k = 2, 1, 0                      # firstaxis, NEXT[i + parity], NEXT[i - parity + 1]
    ai, ak = ak, ai                        # rotating frame
    ai, aj, ak = -ai, -aj, -ak             # odd parity
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs = ci*ck, ci*sk
    sc, ss = si*ck, si*sk
    M = np.eye(3)
    M[i, i] = cj*ck
    M[i, j] = sj*sc - cs
    M[i, k] = sj*cc + ss
    M[j, i] = cj*sk
    M[j, j] = sj*ss + cc
    M[j, k] = sj*cs - sc
    M[k, i] = -sj
    M[k, j] = cj*si
    M[k, k] = cj*ci
    return M


def euler2rot3d(euler):
    """util/misc_math.py:151-161: optical-design convention (alpha, beta left-handed), degrees"""
    e = np.asarray(euler, dtype=float)
    return euler2mat_rxyz(*np.deg2rad(np.array([-e[0], -e[1], e[2]])))


class DecenterData:
    """Data mirror of elem/surface.py:274-337: position / orientation changes at an interface.

    dtype: 'decenter' (applied before the surface), 'reverse' (applied after it, in reverse),
    'dec and return' (before, and undone after), 'bend' (fold mirror: before and after)."""

    def __init__(self, dtype, x=0., y=0., alpha=0., beta=0., gamma=0.):
        self.dtype = dtype
        self.dec = np.array([x, y, 0.])
        self.euler = np.array([alpha, beta, gamma], dtype=float)
        self.rot_pt = np.array([0., 0., 0.])
        self.rot_mat = None
        self.update()

    def update(self):
        self.rot_mat = euler2rot3d(self.euler) if self.euler.any() else None

    def tform_before_surf(self):
        if self.dtype != 'reverse':
            return self.rot_mat, self.dec
        return None, np.array([0., 0., 0.])

    def tform_after_surf(self):
        if self.dtype in ('reverse', 'dec and return'):
            rt = self.rot_mat
            if rt is not None:
                rt = rt.transpose()
            return rt, -self.dec
        if self.dtype == 'bend':
            return self.rot_mat, np.array([0., 0., 0.])
        return None, np.array([0., 0., 0.])

    def to_dict(self):
        return {'dtype': self.dtype, 'dec': self.dec.tolist(), 'euler': self.euler.tolist()}

    @classmethod
    def from_dict(cls, d):
        o = cls(d['dtype'], d['dec'][0], d['dec'][1], *d['euler'])
        o.dec[2] = d['dec'][2]
        return o


def forward_transform(s1, zdist, s2):
    """Rotation and translation from s1 coordinates to s2 coordinates, elem/transform.py:145-166
    (same numpy operations in the same order: the memory layout of the result decides which
    dgemv rounding the trace applies, table.py)."""
    t_orig = np.array([0., 0., zdist])
    r_after_s1 = r_before_s2 = None
    if getattr(s1, 'decenter', None):
        r_after_s1, t_after_s1 = s1.decenter.tform_after_surf()
        t_orig += t_after_s1
    if getattr(s2, 'decenter', None):
        r_before_s2, t_before_s2 = s2.decenter.tform_before_surf()
        t_orig += t_before_s2
    r_cascade = np.identity(3)
    if r_after_s1 is not None:
        t_orig = np.matmul(r_after_s1, t_orig)
        r_cascade = r_after_s1
        if r_before_s2 is not None:
            r_cascade = np.matmul(r_after_s1, r_before_s2)
    elif r_before_s2 is not None:
        r_cascade = r_before_s2
    return r_cascade, t_orig


def reverse_transform(s2, zdist, s1):
    """rotation and translation from s2 coordinates back to s1 coordinates, the decenters applied in
    reverse order (elem/transform.py:169-191)"""
    t_orig = np.array([0., 0., zdist])
    r_before_s2 = r_after_s1 = None
    if getattr(s2, 'decenter', None):
        r_before_s2, t_before_s2 = s2.decenter.tform_before_surf()
        t_orig += t_before_s2
    if getattr(s1, 'decenter', None):
        r_after_s1, t_after_s1 = s1.decenter.tform_after_surf()
        t_orig += t_after_s1
    r_cascade = np.identity(3)
    if r_before_s2 is not None:
        r_cascade = r_before_s2.transpose()
        t_orig = np.matmul(r_cascade, t_orig)
        if r_after_s1 is not None:
            r_cascade = np.matmul(r_cascade, r_after_s1.transpose())
    elif r_after_s1 is not None:
        r_cascade = r_after_s1.transpose()
    return r_cascade, t_orig


def compute_global_coords(ifcs, gaps, glo=1, origin=None):
    """``(rot, t)`` of every interface with respect to interface ``glo`` (elem/transform.py:18-76):
    ``rot.dot(p) + t`` takes a point from an interface's local frame to the global one --
    what ``trace.list_ray(ray, tfrms=...)`` prints.  ``origin``: optional ``(r, t)`` from the
    desired global origin to interface ``glo``."""
    r0, t0 = (np.identity(3), np.array([0., 0., 0.])) if origin is None else origin
    tfrms = [(r0, t0)]

    def accumulate(indices, calc, sign):
        r_prev, t_prev = r0, t0
        for b4, nxt, gap in indices:
            r, t = calc(ifcs[b4], sign*gaps[gap].thi, ifcs[nxt])
            t_new = np.matmul(r_prev, t) + t_prev
            r_new = np.matmul(r_prev, r)
            tfrms.append((r_new, t_new))
            r_prev, t_prev = r_new, t_new

    if glo > 0:                  # from the global surface back to the object surface
        accumulate([(i, i - 1, i - 1) for i in range(glo, 0, -1)], reverse_transform, -1)
        tfrms.reverse()
    accumulate([(i, i + 1, i) for i in range(glo, len(ifcs) - 1)], forward_transform, +1)
    return tfrms


def compute_local_transforms(ifcs, gaps, step=1):
    """elem/transform.py:79-107: ``(r.T, t)`` per interface, in path order (``step=1``) or -- for
    paths traced from the image back to the object -- from the last interface to the first
    (``step=-1``, thicknesses negated, decenters undone in reverse order)"""
    tfrms = []
    n = len(ifcs)
    if step == 1:
        for i in range(n - 1):
            r, t = forward_transform(ifcs[i], gaps[i].thi, ifcs[i + 1])
            tfrms.append((r.transpose(), t))
    elif step == -1:
        for i in range(n - 1, 0, -1):
            r, t = reverse_transform(ifcs[i], -gaps[i - 1].thi, ifcs[i - 1])
            tfrms.append((r.transpose(), t))
    else:
        raise ValueError('step must be 1 or -1')
    tfrms.append((np.identity(3), np.array([0., 0., 0.])))
    return tfrms


# ------------------------------------------------------------------ surface
class Surface:
    """Data mirror of elem/surface.py:38 ``Surface(Interface)``."""

    def __init__(self, lbl='', profile=None, interact_mode='transmit',
                 max_aperture=1.0, clear_apertures=None, **kwargs):
        self.label = lbl
        self.profile = profile if profile is not None else Spherical()
        self.interact_mode = interact_mode
        self.max_aperture = max_aperture
        self.clear_apertures = list(clear_apertures) if clear_apertures else []
        self.edge_apertures = []
        self.decenter = None
        self.delta_n = 0.0

    delta_n = 0.0       # index step across the interface at the reference wavelength, signed
                        # by the propagation direction (set by SequentialModel.update_model)

    @property
    def optical_power(self):            # elem/surface.py:124-130
        return self.delta_n*self.profile.cv

    @optical_power.setter
    def optical_power(self, pwr):
        self.profile.cv = pwr/self.delta_n if self.delta_n != 0.0 else 0.0

    def set_max_aperture(self, max_ap):
        """elem/surface.py:174-179: the clear apertures that are not obscurations follow"""
        self.max_aperture = max_ap
        for ca in self.clear_apertures:
            if not ca.is_obscuration:
                ca.resize(max_ap)

    def update(self):
        self.profile.update()
        if self.decenter is not None:
            self.decenter.update()
        return self


class HolographicElement:
    """Two-point hologram phase element: data mirror of oprops/doe.py:326-370."""

    def __init__(self, label='', ref_pt=None, ref_virtual=False, obj_pt=None, obj_virtual=False,
                 ref_wl=550.):
        self.label = label
        self.ref_pt = np.array([0., 0., -1e10]) if ref_pt is None else np.array(ref_pt, dtype=float)
        self.ref_virtual = ref_virtual
        self.obj_pt = np.array([0., 0., -1e10]) if obj_pt is None else np.array(obj_pt, dtype=float)
        self.obj_virtual = obj_virtual
        self.ref_wl = ref_wl

    def to_dict(self):
        return {'type': 'HolographicElement', 'label': self.label,
                'ref_pt': self.ref_pt.tolist(), 'ref_virtual': bool(self.ref_virtual),
                'obj_pt': self.obj_pt.tolist(), 'obj_virtual': bool(self.obj_virtual),
                'ref_wl': self.ref_wl}

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.pop('type', None)
        return cls(**d)


def radial_phase_fct(pt, coefficients):
    """Name-compatible stand-in for oprops/doe.py:28-54: the table compiles a
    DiffractiveElement by the NAME of its phase function (table.py); the phase itself
    is evaluated on the device (csrc/rt_device.cuh radial_doe_phase)."""
    raise NotImplementedError('evaluated on the device; see rayoptics_b200.table')


class DiffractionGrating:
    """Linear grating phase element: data mirror of oprops/doe.py:57-117."""

    def __init__(self, label='', order=1, grating_normal=None, grating_freq_um=1.0,
                 grating_lpmm=None, interact_mode='transmit'):
        self.label = label
        self.grating_normal = (np.array([0., 1., 0.]) if grating_normal is None
                               else np.array(grating_normal, dtype=float))
        # same expressions as the reference constructor / setter (doe.py:82-99)
        self.grating_lpmm = grating_lpmm if grating_lpmm is not None else 1/(grating_freq_um*1000)
        self.order = order
        self.interact_mode = interact_mode

    @property
    def grating_lpmm(self):
        return self._grating_lpmm

    @grating_lpmm.setter
    def grating_lpmm(self, grating_lpmm):
        self._grating_lpmm = grating_lpmm
        self._grating_spacing_nm = 1e6/grating_lpmm

    def to_dict(self):
        return {'type': 'DiffractionGrating', 'label': self.label, 'order': self.order,
                'grating_normal': self.grating_normal.tolist(), 'grating_lpmm': self._grating_lpmm,
                'interact_mode': self.interact_mode}

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.pop('type', None)
        return cls(**d)


class DiffractiveElement:
    """Phase-function DOE: data mirror of oprops/doe.py:214-270.  Only the
    reference's own ``radial_phase_fct`` is compiled into the surface table."""

    def __init__(self, label='', coefficients=None, ref_wl=550., order=1, phase_fct=None):
        self.label = label
        self.coefficients = [] if coefficients is None else list(coefficients)
        self.ref_wl = ref_wl
        self.order = order
        self.phase_fct = radial_phase_fct if phase_fct is None else phase_fct

    def to_dict(self):
        return {'type': 'DiffractiveElement', 'label': self.label,
                'coefficients': list(self.coefficients), 'ref_wl': self.ref_wl, 'order': self.order}

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.pop('type', None)
        return cls(**d)


def phase_element_from_dict(d):
    kinds = {'HolographicElement': HolographicElement, 'DiffractionGrating': DiffractionGrating,
             'DiffractiveElement': DiffractiveElement}
    return kinds[d.get('type', 'HolographicElement')].from_dict(d)


class ThinLens:
    """Thin lens interface (oprops/thinlens.py:17-140): a plane with a
    HolographicElement whose object point encodes the power."""

    def __init__(self, lbl='', power=0.0, ref_index=1.5, center_wvl=550., max_aperture=1.0,
                 phase_element=None, interact_mode='transmit'):
        self.label = lbl
        self.interact_mode = interact_mode
        self.max_aperture = max_aperture
        self.ref_index = ref_index
        self.decenter = None
        self.delta_n = 0.0
        self.phase_element = phase_element or HolographicElement(ref_wl=center_wvl)
        self._power = power
        if phase_element is None:
            self.optical_power = power

    @property
    def optical_power(self):
        return self._power

    @optical_power.setter
    def optical_power(self, pwr):          # thinlens.py:97-105
        self._power = pwr
        self.phase_element.obj_pt[2] = 1./pwr if pwr != 0 else 1e+10
        self.phase_element.obj_virtual = True if pwr > 0. else False

    def set_max_aperture(self, max_ap):
        self.max_aperture = max_ap

    def update(self):
        return self


# -------------------------------------------------------------------- media
class Medium:
    """Refractive-index source of a gap: ``rindex(wvl_nm)``.

    The reference delegates this to the un-vendored ``opticalglass`` package
    (seq/sequential.py:259-274); here the index table is an *input* shared by
    the oracle and the engine.
    """

    def rindex(self, wvl):
        raise NotImplementedError

    def to_dict(self):
        d = {'type': type(self).__name__}
        d.update(vars(self))
        return d


class Air(Medium):
    def rindex(self, wvl):
        return 1.0


class ConstantIndex(Medium):
    def __init__(self, n=1.0, label=''):
        self.n = n
        self.label = label

    def rindex(self, wvl):
        return self.n


class Sellmeier(Medium):
    """n**2 = 1 + sum B_i l**2/(l**2 - C_i), l in um; coefs = [B1,B2,B3,C1,C2,C3]
    as stored per glass in the reference's .roa files."""

    def __init__(self, coefs=None, label=''):
        self.coefs = list(coefs)
        self.label = label

    def rindex(self, wvl):
        w2 = (wvl*1.0e-3)**2
        B1, B2, B3, C1, C2, C3 = self.coefs
        n2 = 1.0 + B1*w2/(w2 - C1) + B2*w2/(w2 - C2) + B3*w2/(w2 - C3)
        return math.sqrt(n2)


class AbbeGlass(Medium):
    """Two-term Cauchy glass fitted through (n_d, V_d): n = A + B/l**2 with
    n_F - n_C = (n_d - 1)/V_d.  A documented synthetic dispersion model for
    prescriptions that give only n_d and V_d (raytr/tests/ag_dblgauss_s.py)."""

    L_D, L_F, L_C = 0.5875618, 0.4861327, 0.6562725

    def __init__(self, nd=1.5, vd=60.0, label=''):
        self.nd = nd
        self.vd = vd
        self.label = label

    def rindex(self, wvl):
        if self.vd == 0.0:
            return self.nd
        l2 = (wvl*1.0e-3)**2
        B = ((self.nd - 1.0)/self.vd)/(1.0/self.L_F**2 - 1.0/self.L_C**2)
        A = self.nd - B/self.L_D**2
        return A + B/l2


class TableIndex(Medium):
    def __init__(self, wvls=None, ns=None, label=''):
        self.wvls = list(wvls)
        self.ns = list(ns)
        self.label = label

    def rindex(self, wvl):
        for w, n in zip(self.wvls, self.ns):
            if w == wvl:
                return n
        return float(np.interp(wvl, self.wvls, self.ns))


_MEDIUM_CLASSES = {c.__name__: c for c in (Air, ConstantIndex, Sellmeier, AbbeGlass, TableIndex)}


def medium_from_dict(d):
    d = dict(d)
    cls = _MEDIUM_CLASSES[d.pop('type')]
    return cls(**d)


class Gap:
    """seq/gap.py:21 -- thickness + medium."""

    def __init__(self, t=0.0, med=None):
        self.thi = t
        self.medium = med if med is not None else Air()


# ---------------------------------------------------------- sequential model
class SequentialModel:
    """Interfaces, gaps and what ``path()`` needs (seq/sequential.py:41-202).

    ``lcl_tfrms[i] = (rt, t)`` takes interface ``i`` coordinates to interface
    ``i+1`` coordinates as ``rt.dot(p - t)`` (elem/transform.py:79-107).
    """

    def __init__(self, ifcs, gaps, z_dir=None, stop_surface=None, wvlns=None,
                 ref_wvl=0, lcl_tfrms=None):
        assert len(gaps) == len(ifcs) - 1
        self.ifcs = list(ifcs)
        self.gaps = list(gaps)
        self.stop_surface = stop_surface
        self.wvlns = list(wvlns) if wvlns is not None else [550.0]
        self.ref_wvl = ref_wvl
        self.z_dir = list(z_dir) if z_dir is not None else None
        self._tfrms_given = lcl_tfrms
        self._version = 0
        # sequential.py:89: True = the interfaces carry no aperture data, clear apertures follow
        # the boundary rays (OpticalModel.update_optical_properties).  The mirror's own files
        # store apertures, so it is off unless a reader of a foreign format turns it on;
        # ``input_ca_list``: interfaces whose aperture came from the file (cmdproc.py:88-94)
        self.do_apertures = False
        self.input_ca_list = None
        self.update_model()

    # -- reference API used by the hot path
    def get_num_surfaces(self):
        return len(self.ifcs)

    def central_wavelength(self):
        return self.wvlns[self.ref_wvl]

    def index_for_wavelength(self, wvl):
        return self.wvlns.index(wvl)

    def calc_ref_indices_for_spectrum(self, wvls):
        return [[g.medium.rindex(w) for w in wvls] for g in self.gaps]

    def update_model(self, **kwargs):
        """Rebuild index table, z_dir and transforms (seq/sequential.py:612-669)."""
        self.rndx = self.calc_ref_indices_for_spectrum(self.wvlns)
        if self.z_dir is None or len(self.z_dir) != len(self.gaps):
            z_before = 1
            self.z_dir = []
            for ifc in self.ifcs[:-1]:
                z_after = -z_before if ifc.interact_mode == 'reflect' else z_before
                self.z_dir.append(z_after)
                z_before = z_after
        # seq/sequential.py:628-657: indices stay unsigned, the sign of the propagation direction
        # goes into delta_n (used by the Coddington trace, trace.trace_coddington_fan)
        ref = self.index_for_wavelength(self.central_wavelength())
        n_before = self.rndx[0][ref]
        for i, ifc in enumerate(self.ifcs[:len(self.gaps)]):
            n_after = self.rndx[i][ref] if self.z_dir[i] > 0 else -self.rndx[i][ref]
            if not isinstance(ifc, ThinLens):
                ifc.delta_n = n_after - n_before
            n_before = n_after
        for ifc in self.ifcs:
            ifc.update()
        if self._tfrms_given is not None:
            self.lcl_tfrms = [(np.array(rt, dtype=float), np.array(t, dtype=float))
                              for rt, t in self._tfrms_given]
        elif any(getattr(ifc, 'decenter', None) for ifc in self.ifcs):
            self.lcl_tfrms = compute_local_transforms(self.ifcs, self.gaps)
        else:
            self.lcl_tfrms = [(np.identity(3), np.array([0., 0., g.thi])) for g in self.gaps]
            self.lcl_tfrms.append((np.identity(3), np.array([0., 0., 0.])))
        self._gbl_tfrms = None
        self._version += 1

    @property
    def gbl_tfrms(self):
        """global coordinates of the interfaces w.r.t. interface 1 (seq/sequential.py:659; lazily)"""
        if getattr(self, '_gbl_tfrms', None) is None:
            self._gbl_tfrms = compute_global_coords(self.ifcs, self.gaps, 1)
        return self._gbl_tfrms

    def compute_global_coords(self, glo=1, origin=None):
        return compute_global_coords(self.ifcs, self.gaps, glo, origin)

    def path(self, wl=None, start=None, stop=None, step=1):
        """Iterator of ``(Intfc, Gap, Tfrm, Indx, Zdir)`` (seq/sequential.py:149-202)."""
        if wl is None:
            wl = self.central_wavelength()
        wi = self.index_for_wavelength(wl)
        rndx = [n[wi] for n in self.rndx[start:stop:step]]
        return iter(list(zip_longest(self.ifcs[start:stop:step],
                                     self.gaps[start:stop:step],
                                     self.lcl_tfrms[start:stop:step],
                                     rndx,
                                     self.z_dir[start:stop:step])))

    def reverse_path(self, start=None, stop=None, step=-1, wl=None):
        """Iterator of path tuples from the image surface back to the object
        (seq/sequential.py:204-253 for the whole model, ``start = len(ifcs)``): the interfaces in
        reverse order, each with the gap that FOLLOWS it on the way back, the transform into the
        next interface's frame, that gap's index and the negated propagation direction.

        Transforms: the exact inverses of the forward local transforms -- forward
        ``p' = rt (p - t)`` gives backward ``p = rt^T (p' - (-rt t))``.  Between interfaces
        without decenters that is the reference's ``compute_local_transforms(step=-1)``
        (``[0, 0, -thi]``, kept bit for bit); for decentered / tilted interfaces the reference's
        ``reverse_transform`` (restated above, pinned to it) is NOT the inverse of its
        ``forward_transform`` -- a ray turned around at the image does not come back -- so it is
        not used here (tests/test_host.py::test_rays_retrace_themselves_on_the_reverse_path)."""
        if step != -1 or stop is not None or (start is not None and start < len(self.ifcs) - 1):
            raise NotImplementedError('reverse_path: the whole model, image to object')
        if wl is None:
            wl = self.central_wavelength()
        wi = self.index_for_wavelength(wl)
        n = len(self.ifcs)
        tfrms = []
        for i in range(n - 1, 0, -1):                     # from interface i back to interface i-1
            rt_f, t_f = self.lcl_tfrms[i - 1]
            plain = (np.array_equal(rt_f, np.identity(3)) and t_f[0] == 0.0 and t_f[1] == 0.0)
            if plain:
                tfrms.append((np.identity(3), np.array([0., 0., -t_f[2]])))
            else:
                tfrms.append((np.ascontiguousarray(rt_f.T), -np.matmul(rt_f, t_f)))
        tfrms.append((np.identity(3), np.array([0., 0., 0.])))
        path = []
        for i in range(n):
            g = n - 2 - i                     # gap between interface n-1-i and the one before it
            path.append((self.ifcs[n - 1 - i], self.gaps[g] if g >= 0 else None, tfrms[i],
                         self.rndx[g][wi] if g >= 0 else None, -self.z_dir[g] if g >= 0 else None))
        return iter(path)

    # -- analysis drivers of the reference's SequentialModel (seq/sequential.py:1006-1135):
    #    same arguments and return shapes, rays traced in one launch per (field, wavelength)
    def trace(self, pt0, dir0, wvl, **kwargs):
        from . import raytrace
        return raytrace.trace(self, pt0, dir0, wvl, **kwargs)

    def trace_fan(self, fct, fi, xy, num_rays=21, **kwargs):
        """xy determines whether x (=0) or y (=1) fan (sequential.py:1006-1056).
        Returns ``fans_x, fans_y, (max_rho_val, max_y_val), render_colors``."""
        from . import trace
        opm = self.opt_model
        osp = opm.optical_spec
        fld = osp.field_of_view.fields[fi]
        wvl = self.central_wavelength()
        foc = osp.defocus.get_focus()
        engine = {k: kwargs.pop(k) for k in ('table', 'device', 'tracer') if k in kwargs}
        rs_pkg, cr_pkg = trace.setup_pupil_coords(opm, fld, wvl, foc, **engine)
        fld.chief_ray = cr_pkg
        fld.ref_sphere = rs_pkg
        ref_img_pt = rs_pkg[0]      # central-wavelength image point for every wavelength
        wvls = osp.spectral_region
        fans_x, fans_y, rc = [], [], []
        fan_start, fan_stop = np.array([0., 0.]), np.array([0., 0.])
        fan_start[xy], fan_stop[xy] = -1.0, 1.0
        fan_def = [fan_start, fan_stop, num_rays]
        max_rho_val = max_y_val = 0.0
        for wi, wvl in enumerate(wvls.wavelengths):
            rc.append(wvls.render_colors[wi])
            rs_pkg, cr_pkg = trace.setup_pupil_coords(opm, fld, wvl, foc, image_pt=ref_img_pt,
                                                      **engine)
            fld.chief_ray = cr_pkg
            fld.ref_sphere = rs_pkg
            fan = trace.trace_fan(opm, fan_def, fld, wvl, foc,
                                  img_filter=lambda p, ray_pkg, wvl=wvl:
                                  fct(p, xy, ray_pkg, fld, wvl, foc), **engine, **kwargs)
            f_x, f_y = [], []
            for p, y_val in fan:
                f_x.append(p[xy])
                f_y.append(y_val)
                max_rho_val = max(max_rho_val, abs(p[xy]))
                max_y_val = max(max_y_val, abs(y_val))
            fans_x.append(f_x)
            fans_y.append(f_y)
        return np.array(fans_x), np.array(fans_y), (max_rho_val, max_y_val), rc

    def trace_grid(self, fct, fi, wl=None, num_rays=21, form='grid', append_if_none=True,
                   **kwargs):
        """fct is applied to the raw grid and returned as a grid (sequential.py:1058-1085).
        Returns ``grids, render_colors``."""
        from . import trace
        opm = self.opt_model
        osp = opm.optical_spec
        wvls = osp.spectral_region
        wvl = self.central_wavelength()
        wv_list = wvls.wavelengths if wl is None else [wl]
        fld = osp.field_of_view.fields[fi]
        foc = osp.defocus.get_focus()
        engine = {k: kwargs.pop(k) for k in ('table', 'device', 'tracer') if k in kwargs}
        rs_pkg, cr_pkg = trace.setup_pupil_coords(opm, fld, wvl, foc, **engine)
        fld.chief_ray = cr_pkg
        fld.ref_sphere = rs_pkg
        grids = []
        grid_def = [np.array([-1., -1.]), np.array([1., 1.]), num_rays]
        for wi, wvl in enumerate(wv_list):
            grid = trace.trace_grid(opm, grid_def, fld, wvl, foc, form=form,
                                    append_if_none=append_if_none,
                                    img_filter=lambda p, ray_pkg, wi=wi, wvl=wvl:
                                    fct(p, wi, ray_pkg, fld, wvl, foc), **engine, **kwargs)
            grids.append(grid)
        return grids, wvls.render_colors

    def trace_wavefront(self, fld, wvl, foc, num_rays=32, **engine):
        """``[num, num, 3]`` grid of (pupil x, pupil y, OPD in waves), 0 where the ray does
        not reach the image (sequential.py:1087-1119)."""
        from . import trace, waveabr
        opm = self.opt_model
        rs_pkg, cr_pkg = trace.setup_pupil_coords(opm, fld, wvl, foc, **engine)
        fld.chief_ray = cr_pkg
        fld.ref_sphere = rs_pkg
        fod = opm.optical_spec.fod

        def wave(p, ray_pkg):
            if ray_pkg is not None:
                opd = waveabr.wave_abr_full_calc(fod, fld, wvl, foc, ray_pkg, fld.chief_ray,
                                                 fld.ref_sphere)
                opd = opd/opm.nm_to_sys_units(wvl)
            else:
                opd = 0.0
            return np.array([p[0], p[1], opd])

        grid_def = (np.array([-1., -1.]), np.array([1., 1.]), num_rays)
        return trace.trace_grid(opm, grid_def, fld, wvl, foc, img_filter=wave, form='grid', **engine)

    # -- persistence
    def to_dict(self):
        ifcs = []
        for i, ifc in enumerate(self.ifcs):
            e = {'label': ifc.label, 'mode': ifc.interact_mode,
                 'max_aperture': ifc.max_aperture}
            if type(ifc).__name__ == 'ThinLens':
                e['thin_lens'] = {'power': ifc.optical_power, 'ref_index': ifc.ref_index}
            else:
                e['profile'] = ifc.profile.to_dict()
            if hasattr(ifc, 'phase_element'):
                e['phase_element'] = ifc.phase_element.to_dict()
            if getattr(ifc, 'clear_apertures', None):
                e['clear_apertures'] = [ca.to_dict() for ca in ifc.clear_apertures]
            if getattr(ifc, 'decenter', None):
                e['decenter'] = ifc.decenter.to_dict()
            if i < len(self.gaps):
                e['thi'] = self.gaps[i].thi
                e['medium'] = self.gaps[i].medium.to_dict()
                e['z_dir'] = self.z_dir[i]
                rt, t = self.lcl_tfrms[i]
                if self._tfrms_given is not None and (
                        not np.array_equal(rt, np.identity(3)) or t[0] != 0.0 or t[1] != 0.0):
                    rt = np.asarray(rt)
                    order = 'C' if (rt.flags['C_CONTIGUOUS'] and not rt.flags['F_CONTIGUOUS']) else 'F'
                    e['tfrm'] = {'rt': rt.tolist(), 't': np.asarray(t).tolist(), 'order': order}
            ifcs.append(e)
        return {'ifcs': ifcs, 'stop_surface': self.stop_surface,
                'wvls': self.wvlns, 'ref_wvl': self.ref_wvl}

    @classmethod
    def from_dict(cls, d):
        ifcs, gaps, z_dir, tfrms, any_tfrm = [], [], [], [], False
        n = len(d['ifcs'])
        for i, e in enumerate(d['ifcs']):
            if 'thin_lens' in e:
                s = ThinLens(lbl=e.get('label', ''), power=e['thin_lens']['power'],
                             ref_index=e['thin_lens'].get('ref_index', 1.5),
                             max_aperture=e.get('max_aperture', 1.0), interact_mode=e['mode'],
                             phase_element=HolographicElement.from_dict(e['phase_element']))
            else:
                s = Surface(lbl=e.get('label', ''), profile=profile_from_dict(e['profile']),
                            interact_mode=e['mode'], max_aperture=e.get('max_aperture', 1.0),
                            clear_apertures=[aperture_from_dict(a)
                                             for a in e.get('clear_apertures', [])])
                if 'phase_element' in e:
                    s.phase_element = phase_element_from_dict(e['phase_element'])
            if 'decenter' in e:
                s.decenter = DecenterData.from_dict(e['decenter'])
            ifcs.append(s)
            if i < n - 1:
                gaps.append(Gap(e['thi'], medium_from_dict(e['medium'])))
                z_dir.append(e.get('z_dir', 1))
                if 'tfrm' in e:
                    any_tfrm = True
                    # the memory order decides which dgemv rounding numpy applies (table.py)
                    rt = np.array(e['tfrm']['rt'], dtype=float)
                    if e['tfrm'].get('order', 'F') == 'F':
                        rt = np.asfortranarray(rt)
                    tfrms.append((rt, e['tfrm']['t']))
                else:
                    tfrms.append((np.identity(3), [0., 0., e['thi']]))
        tfrms.append((np.identity(3), [0., 0., 0.]))
        return cls(ifcs, gaps, z_dir=z_dir, stop_surface=d.get('stop_surface'),
                   wvlns=d.get('wvls', [550.0]), ref_wvl=d.get('ref_wvl', 0),
                   lcl_tfrms=tfrms if any_tfrm else None)


def gen_sequence(surf_data_list, wvls=(550.0,), ref_wvl=0, sd=None, stop_surface=None,
                 dispersion=True):
    """Build a model from ``[curvature, thickness, n_d, V_d]`` rows, the list
    form used by the reference's own hot-path test
    (seq/sequential.py:1182-1223, raytr/tests/test_sequential.py:38-41)."""
    ifcs, gaps = [], []
    prev_med = Air()
    for row in surf_data_list:
        s = Surface(profile=Spherical(c=row[0]))
        if len(row) > 2 and isinstance(row[2], str) and row[2].casefold() == 'refl':
            s.interact_mode = 'reflect'
            med = prev_med
        elif len(row) > 3 and row[3] != 0 and dispersion:
            med = AbbeGlass(row[2], row[3])
        elif len(row) > 2:
            med = ConstantIndex(row[2]) if row[2] != 1 else Air()
        else:
            med = Air()
        if sd is not None:
            s.set_max_aperture(sd)
        ifcs.append(s)
        gaps.append(Gap(row[1], med))
        prev_med = med
    ifcs[-1].interact_mode = 'dummy'
    return SequentialModel(ifcs, gaps[:-1], stop_surface=stop_surface,
                           wvlns=list(wvls), ref_wvl=ref_wvl)


# ------------------------------------------------------ optical specification
class Field:
    """raytr/opticalspec.py:1197 -- field point + vignetting + aim info."""

    def __init__(self, x=0.0, y=0.0, wt=1.0, vux=0.0, vuy=0.0, vlx=0.0, vly=0.0,
                 aim_pt=None, fov=None):
        if fov is not None:
            self.fov = fov
        self.x, self.y, self.wt = x, y, wt
        self.vux, self.vuy, self.vlx, self.vly = vux, vuy, vlx, vly
        # [x, y] aim point on the paraxial entrance pupil, or -- wide-angle fields -- the scalar
        # z position of the field's real entrance pupil (raytr/wideangle.py)
        self.aim_info = (None if aim_pt is None else float(aim_pt) if np.ndim(aim_pt) == 0
                         else np.array(aim_pt, dtype=float))
        self.chief_ray = None
        self.ref_sphere = None

    def apply_vignetting(self, pupil):
        """opticalspec.py:1339-1353 (including its in-place behaviour for ndarrays)."""
        vig_pupil = pupil[:]
        if pupil[0] < 0.0:
            if self.vlx != 0.0:
                vig_pupil[0] *= (1.0 - self.vlx)
        else:
            if self.vux != 0.0:
                vig_pupil[0] *= (1.0 - self.vux)
        if pupil[1] < 0.0:
            if self.vly != 0.0:
                vig_pupil[1] *= (1.0 - self.vly)
        else:
            if self.vuy != 0.0:
                vig_pupil[1] *= (1.0 - self.vuy)
        return vig_pupil

    # value / fractional access (opticalspec.py:1182-1222); `fov` is set by FieldSpec
    fov = None

    def _rel(self, v):
        fov = self.fov
        if fov is None or fov.is_relative:
            return v
        return v/fov.value if fov.value != 0 else 0.0

    @property
    def xf(self):
        return self._rel(self.x)

    @property
    def yf(self):
        return self._rel(self.y)

    @property
    def xv(self):
        return self.x*self.fov.value if (self.fov is not None and self.fov.is_relative) else self.x

    @xv.setter
    def xv(self, x_val):            # opticalspec.py:1188-1191: unscaled value in, stored per is_relative
        self.x = x_val/self.fov.value if (self.fov is not None and self.fov.is_relative) else x_val

    @property
    def yv(self):
        return self.y*self.fov.value if (self.fov is not None and self.fov.is_relative) else self.y

    @yv.setter
    def yv(self, y_val):
        self.y = y_val/self.fov.value if (self.fov is not None and self.fov.is_relative) else y_val

    def vignetting_bbox(self, pupil_spec, oversize=1.):
        """bbox of the vignetted pupil ray extents (opticalspec.py:1326-1333)"""
        poly = [self.apply_vignetting(pup_ray) for pup_ray in pupil_spec.pupil_rays]
        return oversize*np.array([np.min(poly, axis=0), np.max(poly, axis=0)])

    def clear_vignetting(self):
        self.vux = self.vuy = self.vlx = self.vly = 0.

    def to_dict(self):
        return {'x': self.x, 'y': self.y, 'wt': self.wt, 'vux': self.vux, 'vuy': self.vuy,
                'vlx': self.vlx, 'vly': self.vly,
                'aim_pt': (None if self.aim_info is None else float(self.aim_info)
                           if np.ndim(self.aim_info) == 0 else list(map(float, self.aim_info)))}


class OpticalModel:
    """Container with the reference's access keys (optical/opticalmodel.py:186-202):
    ``opm['seq_model']``, ``opm['optical_spec']``, ``opm['analysis_results']``."""

    def __init__(self, seq_model, optical_spec=None, name=''):
        self.name = name
        self.seq_model = seq_model
        self.optical_spec = optical_spec
        self.analysis_results = {'parax_data': None}
        seq_model.opt_model = self
        if optical_spec is not None:
            optical_spec.opt_model = self
            optical_spec.update_model()

    _keys = {'sm': 'seq_model', 'seq_model': 'seq_model', 'osp': 'optical_spec',
             'optical_spec': 'optical_spec', 'ar': 'analysis_results',
             'analysis_results': 'analysis_results'}

    def __getitem__(self, key):
        return getattr(self, self._keys[key])

    dimensions = 'mm'        # SystemSpec.dimensions (optical/opticalmodel.py:33-97)

    def nm_to_sys_units(self, nm):
        """SystemSpec.nm_to_sys_units, optical/opticalmodel.py:77-97 (the CODE V importer stores
        'inches', which that function does not know and passes through -- kept as is)"""
        d = self.dimensions
        if d == 'm':
            return 1e-9*nm
        if d == 'cm':
            return 1e-7*nm
        if d == 'mm':
            return 1e-6*nm
        if d == 'in':
            return 1e-6*nm/25.4
        if d == 'ft':
            return 1e-6*nm/304.8
        return nm

    def update_model(self, **kwargs):
        self.seq_model.update_model(**kwargs)
        if self.optical_spec is not None:
            self.optical_spec.update_model(**kwargs)

    def update_optical_properties(self, bundle_fn=None, do_aiming=None, trace_fn=None):
        """The ray-traced part of the reference's ``OpticalModel.update_model`` (optical/
        opticalmodel.py:318-354), which its importers run after reading a lens file: first-order data,
        chief-ray aiming of every field (``OpticalSpecs.update_optical_properties``,
        opticalspec.py:263-281) and -- when the interfaces carry no aperture data
        (``seq_model.do_apertures``) or only some do (``input_ca_list``, cmdproc.py:88-94) --
        clear apertures from the boundary rays of all fields (sequential.py:670-674).  All rays go
        through bundles (``vigcalc.aim_all_fields_batched`` / ``set_clear_apertures_batched``);
        ``bundle_fn`` (and ``trace_fn``, single rays of the wide-angle pupil search): test seams,
        default the CUDA engine.  Returns the number of interfaces whose aperture was set."""
        from . import vigcalc
        self.update_model()
        sm, osp = self.seq_model, self.optical_spec
        if sm.get_num_surfaces() <= 2:
            return 0
        if bundle_fn is None:
            bundle_fn = vigcalc.cuda_bundle_fn(self)
        if osp.do_aiming if do_aiming is None else do_aiming:
            vigcalc.aim_all_fields_batched(self, bundle_fn, trace_fn=trace_fn)
        given = list(sm.input_ca_list or [])
        if not (sm.do_apertures or given):
            return 0
        before = [ifc.max_aperture for ifc in sm.ifcs]
        vigcalc.set_clear_apertures_batched(self, bundle_fn, avoid_list=given or None)
        return sum(a != ifc.max_aperture for a, ifc in zip(before, sm.ifcs))

    def to_dict(self):
        d = {'format': 'b200rt-model-v1', 'name': self.name}
        if self.dimensions != 'mm':
            d['dimensions'] = self.dimensions
        d.update(self.seq_model.to_dict())
        if self.optical_spec is not None:
            d['optical_spec'] = self.optical_spec.to_dict()
        return d

    def save(self, path):
        with open(path, 'w') as f:
            json.dump(self.to_dict(), f, indent=1)

    @classmethod
    def from_dict(cls, d):
        from .opticalspec import OpticalSpecs
        sm = SequentialModel.from_dict(d)
        osp = OpticalSpecs.from_dict(d['optical_spec']) if 'optical_spec' in d else None
        opm = cls(sm, osp, name=d.get('name', ''))
        if 'dimensions' in d:
            opm.dimensions = d['dimensions']
        return opm

    @classmethod
    def load(cls, path):
        with open(path) as f:
            return cls.from_dict(json.load(f))
