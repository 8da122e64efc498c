"""Run the REAL reference hot path (/root/reference) as the golden source.

TEST INFRASTRUCTURE, and only usable in the build container: /root/reference
does not exist on the GPU box, so nothing here is imported by `-m gpu` tests,
smoke() or bench.py.  It is used by tools/make_golden.py (which commits the
vectors under tests/golden/) and by CPU tests that skip when the reference
tree is absent.

The reference's hot-path modules import with two `sys.modules` shims for
third-party packages that are not installed:
`anytree` (imported by rayoptics/elem/__init__.py:15) and `transforms3d`
(imported by rayoptics/util/misc_math.py:14, used only for tilted surfaces).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np

REF_SRC = os.environ.get('B200RT_REF_SRC', '/root/reference/src')   # the env override lets the suite be run as on a box without the reference


def available():
    return os.path.isdir(os.path.join(REF_SRC, 'rayoptics'))


_mods = None


def ref():
    """Import the reference's hot-path modules (once) and return them."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError('/root/reference is not present')
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if 'anytree' not in sys.modules:
        at = types.ModuleType('anytree')

        class Node:   # placeholder, never instantiated on the hot path
            def __init__(self, *a, **k):
                pass
        at.Node = Node
        sys.modules['anytree'] = at
    if 'transforms3d' not in sys.modules:
        t3 = types.ModuleType('transforms3d')
        t3.euler = types.ModuleType('transforms3d.euler')
        sys.modules['transforms3d'] = t3
        sys.modules['transforms3d.euler'] = t3.euler
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from rayoptics.raytr import raytrace, traceerror
        from rayoptics.elem import surface, profiles
        from rayoptics.util import misc_math
    _mods = types.SimpleNamespace(raytrace=raytrace, traceerror=traceerror,
                                  surface=surface, profiles=profiles, misc_math=misc_math)
    return _mods


def ref_profile(prf):
    """mirror profile (rayoptics_b200.model) -> reference profile object"""
    P = ref().profiles
    name = type(prf).__name__
    if name == 'Spherical':
        return P.Spherical(c=prf.cv)
    if name == 'Conic':
        return P.Conic(c=prf.cv, cc=prf.cc)
    if name == 'EvenPolynomial':
        return P.EvenPolynomial(c=prf.cv, cc=prf.cc, coefs=list(prf.coefs)).update()
    if name == 'RadialPolynomial':
        return P.RadialPolynomial(c=prf.cv, ec=prf.ec, coefs=list(prf.coefs)).update()
    if name == 'YToroid':
        return P.YToroid(c=prf.cv, cR=prf.cR, cc=prf.cc, coefs=list(prf.coefs)).update()
    if name == 'XToroid':
        return P.XToroid(c=prf.cv, cR=prf.cR, cc=prf.cc, coefs=list(prf.coefs)).update()
    raise ValueError(name)


def ref_surface(ifc):
    if type(ifc).__name__ == 'ThinLens':
        import importlib
        TL = importlib.import_module('rayoptics.oprops.thinlens')
        DOE = importlib.import_module('rayoptics.oprops.doe')
        pe = ifc.phase_element
        t = TL.ThinLens(power=ifc.optical_power, ref_index=ifc.ref_index, max_ap=ifc.max_aperture)
        t.interact_mode = ifc.interact_mode
        t.phase_element = DOE.HolographicElement(ref_pt=np.array(pe.ref_pt), ref_virtual=pe.ref_virtual,
                                                 obj_pt=np.array(pe.obj_pt), obj_virtual=pe.obj_virtual,
                                                 ref_wl=pe.ref_wl)
        return t
    S = ref().surface
    s = S.Surface(profile=ref_profile(ifc.profile), interact_mode=ifc.interact_mode,
                  max_ap=ifc.max_aperture)
    cas = []
    for ca in ifc.clear_apertures:
        kw = dict(x_offset=ca.x_offset, y_offset=ca.y_offset, rotation=ca.rotation,
                  is_obscuration=ca.is_obscuration)
        name = type(ca).__name__
        if name == 'Circular':
            cas.append(S.Circular(radius=ca.radius, **kw))
        elif name == 'Rectangular':
            cas.append(S.Rectangular(x_half_width=ca.x_half_width,
                                     y_half_width=ca.y_half_width, **kw))
        else:
            cas.append(S.Elliptical(x_half_width=ca.x_half_width,
                                    y_half_width=ca.y_half_width, **kw))
    s.clear_apertures = cas
    pe = getattr(ifc, 'phase_element', None)
    if pe is not None:
        import importlib
        DOE = importlib.import_module('rayoptics.oprops.doe')
        kind = type(pe).__name__
        if kind == 'DiffractionGrating':
            s.phase_element = DOE.DiffractionGrating(order=pe.order,
                                                     grating_normal=np.array(pe.grating_normal),
                                                     grating_lpmm=pe.grating_lpmm,
                                                     interact_mode=pe.interact_mode)
        elif kind == 'DiffractiveElement':
            s.phase_element = DOE.DiffractiveElement(coefficients=list(pe.coefficients),
                                                     ref_wl=pe.ref_wl, order=pe.order,
                                                     phase_fct=DOE.radial_phase_fct)
        elif kind == 'HolographicElement':
            s.phase_element = DOE.HolographicElement(ref_pt=np.array(pe.ref_pt),
                                                     ref_virtual=pe.ref_virtual,
                                                     obj_pt=np.array(pe.obj_pt),
                                                     obj_virtual=pe.obj_virtual, ref_wl=pe.ref_wl)
        else:
            raise ValueError(kind)
    return s


def ref_path(seq_model, wvl):
    """Path list of reference Surface objects for a mirror SequentialModel."""
    out = []
    for ifc, gap, tfrm, n, z_dir in seq_model.path(wvl):
        out.append([ref_surface(ifc), None, tfrm, n, z_dir])
    return out


STATUS = {'TraceMissedSurfaceError': 1, 'TraceTIRError': 2, 'TraceRayBlockedError': 3,
          'TraceEvanescentRayError': 4}


def ref_trace(path, pt0, dir0, wvl, **kwargs):
    """trace_raw on a path list -> dict(ray [n_seg,10], op, status, fail_surf)."""
    R = ref()
    pt0 = np.array(pt0, dtype=float)
    dir0 = np.array(dir0, dtype=float)
    status, fail_surf = 0, -1
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            ray, op, _ = R.raytrace.trace_raw(iter(path), pt0, dir0, wvl, **kwargs)
        except R.traceerror.TraceError as e:
            status = STATUS[type(e).__name__]
            fail_surf = e.surf if e.surf is not None else 0
            if e.ray_pkg is None:
                ray, op = [], 0.0
            else:
                ray, op, _ = e.ray_pkg
        except (ValueError, ZeroDivisionError):
            status, fail_surf, ray, op = 5, -2, [], 0.0
    segs = np.zeros((len(ray), 10))
    for k, (p, d, dst, nrml) in enumerate(ray):
        segs[k, 0:3] = p
        segs[k, 3:6] = d
        segs[k, 6] = dst
        segs[k, 7:10] = nrml
    return {'ray': segs, 'op': float(op), 'status': status, 'fail_surf': fail_surf,
            'n_seg': len(ray)}
