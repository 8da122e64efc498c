"""Self-contained reader of Zemax ``.zmx`` lens files.

The reference imports ``.zmx`` through ``zemax/zmxread.py:93-388`` with glass names
resolved by the un-vendored ``opticalglass`` catalogs.  This reader builds the
``OpticalModel`` mirror of this package from the keywords that describe a rotationally
symmetric sequential system:

  UNIT, NAME, ENPD / FNUM / OBNA, WAVM (or WAVL / WWGT), FTYP, XFLN / YFLN,
  VDXN VDYN VCXN VCYN, SURF with TYPE (STANDARD, EVENASPH, XOSPHERE, TOROIDAL, COORDBRK, PARAXIAL,
  DGRATING),
  CURV, DISZ, GLAS, DIAM, CONI, PARM, XDAT, STOP.

Coordinate breaks (zmxread.py:314-316,341-355): a COORDBRK surface becomes a phantom
interface carrying ``DecenterData('decenter')`` with PARM 1,2 = x, y decenter, PARM 3,4,5 =
tilts about x, y, z and a non-zero PARM 6 (order flag) turning it into a 'reverse' decenter;
the local transforms follow as in elem/transform.py (model.compute_local_transforms).

Conventions copied from the reference: wavelengths are collected in order of first
appearance, a trailing 550 nm filler is dropped and the reference wavelength is the middle
one (zmxread.py:251-257); vignetting is converted to the asymmetric form
``vly = vc + vd``, ``vuy = vc - vd`` (:266-275); an infinite object distance becomes 1e10
(:233-236); object and image surfaces are dummies.

Glasses: ``MIRROR``, a model glass (``GLAS ___BLANK 1 0 nd vd ...``), or a catalog name
looked up in ``glass_map`` (name -> Medium, index, or ``(n_d, V_d)``; case-insensitive, with
and without a ``_MOLD``-type suffix).  This package ships no glass catalog: unknown names
raise ``KeyError``.  Non-sequential data, floating apertures (FLOA) and other surface types
raise ``NotImplementedError``.
"""
from __future__ import annotations

import math

from . import model as M
from .opticalspec import OpticalSpecs, WvlSpec, PupilSpec, FieldSpec, FocusRange
from .seq import _medium


def _read_text(path):
    raw = open(path, 'rb').read()
    if raw[:2] in (b'\xff\xfe', b'\xfe\xff'):
        return raw.decode('utf-16')
    return raw.decode('latin-1')


def open_zmx(path, glass_map=None, do_update=False, bundle_fn=None):
    """Read a Zemax ``.zmx`` file into an ``OpticalModel`` mirror.  ``do_update``: run the ray-traced
    part of the reference's post-import ``update_model`` (``OpticalModel.update_optical_properties``:
    chief-ray aiming, and clear apertures from boundary rays unless the file fixes them with
    DIAM records of type 1 / 4 / 6, zmxread.py:247-249,399-440)."""
    title = ''
    pupil = None
    wvls, wts = [], []
    ftyp, n_fields = 0, None
    xf, yf = [], []
    vig = {}
    surfs, cur = [], None
    for line in _read_text(path).splitlines():
        parts = line.strip().split(None, 1)
        if not parts:
            continue
        cmd = parts[0]
        inputs = parts[1] if len(parts) == 2 else ''
        items = inputs.split()
        if cmd == 'UNIT':
            if items and items[0] != 'MM':
                raise NotImplementedError(f'UNIT {items[0]}')
        elif cmd == 'NAME':
            title = inputs.strip('"')
        elif cmd == 'ENPD':
            pupil = (('object', 'epd'), float(items[0]))
        elif cmd == 'FNUM':
            pupil = (('image', 'f/#'), float(items[0]))
        elif cmd == 'OBNA':
            pupil = (('object', 'NA'), float(items[0]))
        elif cmd == 'FLOA':
            raise NotImplementedError('FLOA: pupil defined by the stop size')
        elif cmd == 'WAVM':
            w = float(items[1])*1e3
            if w not in wvls:
                wvls.append(w)
                wts.append(float(items[2]))
        elif cmd == 'WAVL':
            wvls = [float(i)*1e3 for i in items]
        elif cmd == 'WWGT':
            wts = [float(i) for i in items]
        elif cmd == 'FTYP':
            ftyp = int(items[0])
            if len(items) > 2:
                n_fields = int(items[2])
        elif cmd == 'XFLN':
            xf = [float(i) for i in items]
        elif cmd == 'YFLN':
            yf = [float(i) for i in items]
        elif cmd in ('VDXN', 'VDYN', 'VCXN', 'VCYN'):
            vig[cmd[:3]] = [float(i) for i in items]
        elif cmd == 'SURF':
            cur = {'type': 'STANDARD', 'cv': 0.0, 'thi': 0.0, 'glass': None, 'model': None,
                   'stop': False, 'cc': 0.0, 'parm': {}, 'diam': None}
            surfs.append(cur)
        elif cur is None:
            continue
        elif cmd == 'TYPE':
            cur['type'] = items[0]
            if items[0] not in ('STANDARD', 'EVENASPH', 'XOSPHERE', 'TOROIDAL', 'COORDBRK', 'PARAXIAL', 'DGRATING'):
                raise NotImplementedError(f'.zmx surface TYPE {items[0]}')
        elif cmd == 'CURV':
            cur['cv'] = float(items[0])
        elif cmd == 'DISZ':
            cur['thi'] = math.inf if items[0].upper().startswith('INF') else float(items[0])
        elif cmd == 'GLAS':
            cur['glass'] = items[0]
            if items[0] == '___BLANK' and len(items) >= 5:
                cur['model'] = (float(items[3]), float(items[4]))
        elif cmd == 'STOP':
            cur['stop'] = True
        elif cmd == 'CONI':
            cur['cc'] = float(items[0])
        elif cmd == 'PARM':
            cur['parm'][int(items[0])] = float(items[1])
        elif cmd == 'XDAT':
            cur.setdefault('xdat', {})[int(items[0])] = float(items[1])
        elif cmd == 'DIAM':
            cur['diam'] = float(items[0])
            cur['diam_type'] = int(float(items[1])) if len(items) > 1 else 0
    if len(surfs) < 2:
        raise ValueError(f'{path}: no surfaces')
    if wvls and len(wvls) > 1 and wvls[-1] == 550.0:       # zmxread.py:251-254
        wvls.pop()
        wts = wts[:len(wvls)]
    if not wvls:
        wvls = [550.0]
    ref_wl = len(wvls)//2                                   # zmxread.py:255
    ifcs, gaps, z_dir = [], [], []
    stop_surface, z, medium_before = None, 1, None
    for i, s in enumerate(surfs):
        mode = 'transmit'
        g = s['glass']
        if g is not None and g.upper() == 'MIRROR':
            mode, g = 'reflect', None
        decenter = None
        if s['type'] == 'EVENASPH':
            k = max(s['parm']) if s['parm'] else 0
            coefs = [s['parm'].get(j + 1, 0.0) for j in range(max(k, 1))]
            prf = M.EvenPolynomial(c=s['cv'], cc=s['cc'], coefs=coefs)
        elif s['type'] == 'XOSPHERE':           # XDAT 1 = number of terms, 2 = norm radius, 3.. = r, r^2, ...
            xd = s.get('xdat', {})
            coefs = [xd[j] for j in sorted(xd) if j >= 3]
            prf = M.RadialPolynomial(c=s['cv'], cc=s['cc'], coefs=coefs)
        elif s['type'] == 'TOROIDAL':           # PARM 1 = radius of rotation, 2.. = y^2, y^4, ...
            pm = s['parm']
            coefs = [pm[j] for j in sorted(pm) if j > 1]
            rR = pm.get(1, 0.0)
            prf = M.YToroid(c=s['cv'], cR=(1.0/rR if rR != 0.0 else 0.0), cc=s['cc'], coefs=coefs)
        elif s['type'] == 'COORDBRK':
            pm = s['parm']
            decenter = M.DecenterData('reverse' if pm.get(6, 0.0) != 0 else 'decenter',
                                      x=pm.get(1, 0.0), y=pm.get(2, 0.0), alpha=pm.get(3, 0.0),
                                      beta=pm.get(4, 0.0), gamma=pm.get(5, 0.0))
            prf = M.Spherical(c=0.0)
            mode = 'phantom'
        elif s['cc'] != 0.0:
            prf = M.Conic(c=s['cv'], cc=s['cc'])
        else:
            prf = M.Spherical(c=s['cv'])
        if i == 0 or i == len(surfs) - 1:
            mode = 'dummy'
        if s['type'] == 'PARAXIAL':             # zmxread.py:317-320,364-366: PARM 1 = focal length
            f = s['parm'].get(1, 0.0)
            ifc = M.ThinLens(power=(1.0/f if f != 0.0 else 0.0), interact_mode=mode)
        else:
            ifc = M.Surface(profile=prf, interact_mode=mode)
            if s['type'] == 'DGRATING':         # zmxread.py:321-324,356-360: PARM 1 = lines/um, PARM 2 = order
                # the reference assigns through the grating_freq_um SETTER: lpmm = freq*1000 (doe.py:105-107)
                ifc.phase_element = M.DiffractionGrating(order=s['parm'].get(2, 1), interact_mode=mode)
                ifc.phase_element.grating_lpmm = s['parm'].get(1, 1.0)*1000
        ifc.decenter = decenter
        if s['diam'] is not None and s['diam'] != 0.0:
            ifc.max_aperture = s['diam']
            if s.get('diam_type', 0) == 2:          # circular obscuration (zmxread.py:416-417)
                ifc.clear_apertures = [M.Circular(radius=s['diam'], is_obscuration=True)]
        if s['stop']:
            stop_surface = i
        ifcs.append(ifc)
        if i < len(surfs) - 1:
            if mode == 'reflect':
                med = medium_before if medium_before is not None else M.Air()
                z = -z
            elif s['model'] is not None:
                med = M.AbbeGlass(s['model'][0], s['model'][1], label='model')
            else:
                med = _medium(None, glass_map) if g is None else _zmx_medium(g, glass_map)
            thi = 1e10 if math.isinf(s['thi']) else s['thi']
            gaps.append(M.Gap(thi, med))
            z_dir.append(z)
            medium_before = med
    sm = M.SequentialModel(ifcs, gaps, z_dir=z_dir, stop_surface=stop_surface, wvlns=wvls,
                           ref_wvl=ref_wl)
    if pupil is None:
        pupil = (('object', 'epd'), 1.0)
    fkey = {0: ('object', 'angle'), 1: ('object', 'height'), 2: ('image', 'height'),
            3: ('image', 'real height')}.get(ftyp, ('object', 'angle'))
    n = max(len(xf), len(yf))
    xf = xf + [0.0]*(n - len(xf))
    yf = yf + [0.0]*(n - len(yf))
    if n_fields is None:            # zmxread.py:260-263: fields up to the one of maximum extent
        mags = [math.hypot(a, b) for a, b in zip(xf, yf)]
        n_fields = (mags.index(max(mags)) + 1) if mags else 1
    fields = []
    for k in range(max(n_fields, 1)):
        get = lambda key: (vig.get(key, []) + [0.0]*(k + 1))[k]      # noqa: E731
        vcx, vdx, vcy, vdy = get('VCX'), get('VDX'), get('VCY'), get('VDY')
        fields.append(M.Field(x=xf[k] if k < n else 0.0, y=yf[k] if k < n else 0.0,
                              vlx=vcx + vdx, vux=vcx - vdx, vly=vcy + vdy, vuy=vcy - vdy))
    max_f = max((math.hypot(f.x, f.y) for f in fields), default=0.0)
    osp = OpticalSpecs(WvlSpec(wvls, ref_wl, wts if len(wts) == len(wvls) else None),
                       PupilSpec(*pupil), FieldSpec(fkey, max_f, fields), FocusRange(0.0, 0.0))
    opm = M.OpticalModel(sm, osp, name=title or str(path).rsplit('/', 1)[-1])
    from .seq import apply_wide_angle_rule
    apply_wide_angle_rule(opm)
    # zmxread.py:247-249: DIAM records of a user-defined kind (1 circular, 4 rectangular, 6 elliptical)
    # turn the automatic apertures off; Zemax' own semi-diameters (kind 0) are start values only
    sm.do_apertures = not any(s_.get('diam_type', 0) in (1, 4, 6) for s_ in surfs)
    if do_update:
        opm.update_optical_properties(bundle_fn)
    return opm


def _zmx_medium(name, glass_map):
    from .seq import SubstituteGlasses
    strict = dict(glass_map) if isinstance(glass_map, SubstituteGlasses) else glass_map
    try:
        return _medium(name, strict)
    except KeyError:
        base = name.rsplit('_', 1)[0]
        if base != name:
            try:
                return _medium(base, strict)
            except KeyError:
                pass
        if strict is not glass_map:
            return glass_map.substitute(name)
        raise
