"""Self-contained reader of CODE V ``.seq`` lens files.

The reference imports ``.seq`` files through ``codev/cmdproc.py:56-449`` and
``codev/reader.py``, resolving glass names with the un-vendored ``opticalglass``
catalogs.  This reader builds the same ``OpticalModel`` mirror the rest of this package
consumes (``model.py`` / ``opticalspec.py``) from the commands that describe a
rotationally symmetric sequential system:

  RDM, TITLE, DIM, EPD / FNO / NA / NAO, WL, REF, XAN YAN / XOB YOB / XIM YIM,
  VUX VLX VUY VLY, SO / S / SI (curvature or radius, thickness, glass | REFL),
  STO, CIR [EDG], K / CON, ASP with A..J (r**4 .. r**20), private catalogs (PRV / PWL / 'name' n.. /
  END: tabulated index, linear interpolation between the tabulated wavelengths), diffractive
  surfaces (DIF DOE, HOR, HWL, HCT R, HCO Cn: DiffractiveElement with the radial phase function,
  codev/cmdproc.py:579-618), '&' continuation lines.  SPS coefficients are not read.

Glasses: ``REFL`` / ``AIR`` / empty, a fictitious glass code ``nnn.vvv`` (n_d = 1.nnn,
V_d = vv.v, CODE V's six-digit form), or a catalog name looked up in ``glass_map``
(name -> Medium, index, or ``(n_d, V_d)``; matched case-insensitively, with and without
the ``_CATALOG`` suffix), then in ``glass_table.json`` (the 17 catalog glasses whose coefficients the
reference's bundled lens files carry).  There is no full glass catalog in this package: an unknown
name raises ``KeyError`` naming the glass, it is never guessed.

Tilts and decenters (codev/cmdproc.py:544-576): XDE YDE ZDE ADE BDE CDE create a
``DecenterData('decenter')`` on the current surface, DAR / BEN / REV change its type to
'dec and return' / 'bend' / 'reverse'; the local transforms follow from them as in
elem/transform.py (model.compute_local_transforms).

Not read (raise ``NotImplementedError``): BAS / RET (basic decenters, return to a surface),
zoom data, special surface types.  Solves (CCY, THC, PIM) and DER lines are ignored like
the reference does when it only builds the model.
"""
from __future__ import annotations

import re

import numpy as np

from . import model as M
from .opticalspec import OpticalSpecs, WvlSpec, PupilSpec, FieldSpec, FocusRange

_IGNORED = {'LEN', 'INI', 'WTW', 'WTF', 'CCY', 'THC', 'PIM', 'DER', 'GO', 'CUF', 'GL1', 'GL2',
            'SLB', 'THM', 'TEM', 'PRE', 'INF', 'MNR', 'MXR', 'CUM', 'VLZ', 'VUZ', 'RMD', 'GLB'}
_TILTS = {'ADE', 'BDE', 'CDE', 'XDE', 'YDE', 'ZDE', 'DAR', 'BEN', 'REV'}
_TILTS_UNREAD = {'BAS', 'RET'}
_ASP_COEFS = 'ABCDEFGHJ'


def _commands(text):
    """lines / ';'-separated commands, comments ('!') dropped, quoted strings kept whole,
    '&' continuation lines joined"""
    joined, pending = [], ''
    for line in text.splitlines():
        line = line.split('!', 1)[0].rstrip()
        if line.endswith('&'):
            pending += line[:-1]
            continue
        joined.append(pending + line)
        pending = ''
    if pending:
        joined.append(pending)
    for line in joined:
        for part in line.split(';'):
            toks = re.findall(r"'[^']*'|\"[^\"]*\"|\S+", part)
            if toks:
                yield toks


def _medium(token, glass_map):
    if token is None or token.upper() in ('AIR', ''):
        return M.Air()
    m = re.fullmatch(r'(\d{3})\.(\d{3})', token)
    if m:                                   # fictitious glass: n_d = 1.nnn, V_d = vv.v
        return M.AbbeGlass(1.0 + int(m.group(1))/1000.0, int(m.group(2))/10.0, label=token)
    keys = [token, token.upper(), token.upper().split('_')[0]]
    for k in keys:
        for gk, gv in (glass_map or {}).items():
            if gk.upper() == k.upper():
                if isinstance(gv, M.Medium):
                    return gv
                if isinstance(gv, (tuple, list)):
                    return M.AbbeGlass(gv[0], gv[1], label=token)
                return M.ConstantIndex(float(gv), label=token)
    m = _builtin_glass(token)
    if m is not None:
        return m
    if isinstance(glass_map, SubstituteGlasses):
        return glass_map.substitute(token)
    raise KeyError(f'glass {token!r}: neither in glass_map nor among the {len(_glass_table())} glasses of '
                   f'rayoptics_b200/glass_table.json (this package ships no full catalog)')


class SubstituteGlasses(dict):
    """``glass_map`` that never fails: names it does not hold resolve through the built-in table
    and then -- what the reference's importers do when a glass is in none of its catalogs
    (``GlassHandlerBase.find_glass``, seq/medium.py:165-203) -- become ``ConstantIndex(1.5,
    'not ' + name)``.  ``not_found`` lists those names after the import."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.not_found = []

    def substitute(self, token):
        if token not in self.not_found:
            self.not_found.append(token)
        return M.ConstantIndex(1.5, label='not ' + token)


_GLASS_TABLE = None


def _glass_table():
    """rayoptics_b200/glass_table.json: the catalog glasses that the reference's bundled lens files
    carry coefficients for (tools/make_glass_table.py)."""
    global _GLASS_TABLE
    if _GLASS_TABLE is None:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'glass_table.json')
        try:
            _GLASS_TABLE = json.load(open(path))
        except OSError:
            _GLASS_TABLE = {}
    return _GLASS_TABLE


def _builtin_glass(token):
    t = _glass_table()
    squeeze = lambda x: x.replace('-', '').replace(' ', '')          # noqa: E731  (CODE V writes NBK7_SCHOTT)
    by_squeezed = {squeeze(k): v for k, v in t.items()}
    for k in (token.upper(), token.upper().split('_')[0]):
        e = t.get(k) or by_squeezed.get(squeeze(k))
        if e is not None:
            if e['form'] == 'sellmeier':
                return M.Sellmeier(e['coefs'], label=e['name'])
            from .roa import PowerSeries
            return PowerSeries(e['coefs'], label=e['name'])
    return None


def open_seq(path, glass_map=None, do_update=False, bundle_fn=None):
    """Read a CODE V ``.seq`` file into an ``OpticalModel`` mirror."""
    with open(path) as f:
        text = f.read()
    radius_mode = False
    units = 'mm'
    title = ''
    wvls, ref_wl = [587.6], 0
    pupil = None
    fld = {}
    vig = {}
    surfs = []        # dicts: cv, thi, glass token, mode, stop, cc, coefs, cir
    cur = None
    private, private_wvls, in_private = {}, None, False
    for toks in _commands(text):
        tla = toks[0][:3].upper()
        args = toks[1:]
        if tla == 'PRV':                       # private catalog (codev/cmdproc.py:358-368, 330-356)
            in_private = True
            continue
        if in_private:
            if tla == 'END':
                in_private, private_wvls = False, None
            elif tla == 'PWL':
                private_wvls = [float(a) for a in args]
            elif toks[0][0] in '\'"' and args and args[0].upper() == 'LAU':
                # Laurent dispersion formula: n^2 = A0 + A1 l^2 + A2 l^-2 + A3 l^-4 + A4 l^-6 + A5 l^-8 (l in um)
                from .roa import PowerSeries
                coefs = ([float(a) for a in args[1:7]] + [0.0]*6)[:6]
                private[toks[0].strip('\'"').upper()] = PowerSeries(coefs, label=toks[0].strip('\'"'))
            elif toks[0][0] in '\'"' and private_wvls and args and re.fullmatch(r'[-+0-9.eE]+', args[0]):
                ns = [float(a) for a in args]
                order = np.argsort(private_wvls[:len(ns)])          # tabulated index, ascending wavelength
                private[toks[0].strip('\'"').upper()] = M.TableIndex(
                    [private_wvls[i] for i in order], [ns[i] for i in order], label=toks[0].strip('\'"'))
            continue
        if tla == 'RDM':
            radius_mode = (not args) or args[0].upper().startswith('Y')
        elif tla == 'TIT':
            title = ' '.join(args).strip('\'"')
        elif tla == 'DIM':
            # system units: numbers are kept as written (the reference stores the unit in system_spec,
            # codev/cmdproc.py:281-289, and traces in those units too); only nm -> system units needs it
            dim = args[0].upper()[0] if args else 'M'
            units = {'M': 'mm', 'C': 'cm', 'I': 'inches'}.get(dim, 'mm')
        elif toks[0].upper() in ('EPD', 'FNO', 'NA', 'NAO'):
            key = {'EPD': ('object', 'epd'), 'FNO': ('image', 'f/#'), 'NA': ('image', 'NA'),
                   'NAO': ('object', 'NA')}[toks[0].upper()]
            pupil = (key, float(args[0]))
        elif tla == 'WL':
            wvls = [float(a) for a in args]
        elif tla == 'REF':
            ref_wl = int(args[0]) - 1
        elif tla in ('XAN', 'YAN', 'XOB', 'YOB', 'XIM', 'YIM', 'XRI', 'YRI'):
            fld[toks[0].upper()[:3]] = [float(a) for a in args]
        elif tla in ('VUX', 'VLX', 'VUY', 'VLY'):
            vig[tla.lower()] = [float(a) for a in args]
        elif tla in ('SO', 'S', 'SI') and toks[0].upper() in ('SO', 'S', 'SI'):
            c = float(args[0]) if args else 0.0
            if radius_mode:
                c = 1.0/c if c != 0.0 else 0.0
            thi = float(args[1]) if len(args) > 1 else 0.0
            g = args[2] if len(args) > 2 else None
            cur = {'cv': c, 'thi': thi, 'glass': g, 'stop': False, 'cc': None, 'coefs': None,
                   'cir': None, 'kind': toks[0].upper()}
            surfs.append(cur)
        elif tla == 'STO':
            cur['stop'] = True
        elif tla == 'CIR':
            vals = [a for a in args if re.fullmatch(r'[-+0-9.eE]+', a)]
            if vals and (len(args) == 1 or args[0].upper() != 'OBS'):
                cur['cir'] = float(vals[0])
        elif tla == 'K' and toks[0].upper() == 'K':
            cur['cc'] = float(args[0])
        elif tla == 'CON':
            cur['cc'] = cur['cc'] if cur['cc'] is not None else 0.0
        elif tla == 'ASP':
            cur['coefs'] = cur['coefs'] or [0.0]*10
        elif toks[0].upper() in _ASP_COEFS and cur is not None:
            cur['coefs'] = cur['coefs'] or [0.0]*10
            cur['coefs'][_ASP_COEFS.index(toks[0].upper()) + 1] = float(args[0])   # A -> r**4
        elif toks[0].upper() in ('DIF', 'HOR', 'HWL', 'HCT', 'HCO') and cur is not None:
            # diffractive surface (codev/cmdproc.py:579-618): DIF DOE creates the element, HOR the
            # order, HWL the construction wavelength, HCT R the radial phase function, HCO Cn its
            # coefficients (HCC: optimisation controls, ignored)
            key = toks[0].upper()
            if key == 'DIF':
                if any(a.upper() == 'DOE' for a in args):
                    cur['doe'] = {'order': 1, 'ref_wl': 550.0, 'coefs': [], 'radial': False}
            elif 'doe' in cur:
                d = cur['doe']
                if key == 'HOR':
                    d['order'] = float(args[0])
                elif key == 'HWL':
                    d['ref_wl'] = float(args[0])
                elif key == 'HCT':
                    d['radial'] = d['radial'] or any(a.upper() == 'R' for a in args)
                elif key == 'HCO':
                    cidx = int(args[0].upper().lstrip('C'))
                    if cidx > len(d['coefs']):
                        d['coefs'].extend([0.]*(cidx - len(d['coefs'])))
                    d['coefs'][cidx - 1] = float(args[1])
        elif tla in _TILTS and toks[0].upper() == tla and cur is not None:
            dc = cur.setdefault('decenter', {'dtype': 'decenter', 'dec': [0., 0., 0.],
                                             'euler': [0., 0., 0.]})
            if tla in ('XDE', 'YDE', 'ZDE'):
                dc['dec']['XYZ'.index(tla[0])] = float(args[0])
            elif tla in ('ADE', 'BDE', 'CDE'):
                dc['euler']['ABC'.index(tla[0])] = float(args[0])
            else:
                dc['dtype'] = {'DAR': 'dec and return', 'BEN': 'bend', 'REV': 'reverse'}[tla]
        elif tla in _TILTS_UNREAD:
            raise NotImplementedError(f'.seq command {toks[0]}: basic decenters / returns are not read')
        elif tla in ('SPS', 'SCO'):
            raise NotImplementedError(f'.seq command {toks[0]}: special surface types are not read')
        # _IGNORED and anything else: not part of the path description (the reference logs
        # and skips unknown commands too)

    if len(surfs) < 2:
        raise ValueError(f'{path}: no surfaces')
    ifcs, gaps, z_dir = [], [], []
    stop_surface = None
    z = 1
    medium_before = None
    for i, s in enumerate(surfs):
        g = s['glass']
        mode = 'transmit'
        if g is not None and g.upper() in ('REFL', 'REFLECT'):
            mode, g = 'reflect', None
        if s['coefs'] is not None and any(c != 0.0 for c in s['coefs']):
            prf = M.EvenPolynomial(c=s['cv'], cc=s['cc'] or 0.0, coefs=s['coefs'])
        elif s['cc'] is not None and s['cc'] != 0.0:
            prf = M.Conic(c=s['cv'], cc=s['cc'])
        else:
            prf = M.Spherical(c=s['cv'])
        if s['kind'] in ('SO', 'SI'):
            mode = 'dummy'
        ifc = M.Surface(profile=prf, interact_mode=mode)
        if s['cir'] is not None:
            ifc.max_aperture = s['cir']
        if s.get('decenter'):
            ifc.decenter = M.DecenterData.from_dict(s['decenter'])
        if s.get('doe'):
            if not s['doe']['radial']:
                raise NotImplementedError('.seq diffractive surface without HCT R (radial phase function)')
            ifc.phase_element = M.DiffractiveElement(coefficients=s['doe']['coefs'], ref_wl=s['doe']['ref_wl'],
                                                     order=s['doe']['order'], phase_fct=M.radial_phase_fct)
        if s['stop']:
            stop_surface = i
        ifcs.append(ifc)
        if i < len(surfs) - 1:
            if mode == 'reflect':
                med = medium_before if medium_before is not None else M.Air()
                z = -z
            elif g is not None and g.strip('\'"').upper() in private:
                med = private[g.strip('\'"').upper()]
            else:
                med = _medium(g, glass_map)
            gaps.append(M.Gap(s['thi'], med))
            z_dir.append(z)
            medium_before = med
    sm = M.SequentialModel(ifcs, gaps, z_dir=z_dir, stop_surface=stop_surface, wvlns=wvls,
                           ref_wvl=ref_wl)
    # optical specification
    if pupil is None:
        pupil = (('object', 'epd'), 1.0)
    if 'XAN' in fld or 'YAN' in fld:
        fkey, fx, fy = ('object', 'angle'), fld.get('XAN'), fld.get('YAN')
    elif 'XOB' in fld or 'YOB' in fld:
        fkey, fx, fy = ('object', 'height'), fld.get('XOB'), fld.get('YOB')
    elif 'XIM' in fld or 'YIM' in fld:
        fkey, fx, fy = ('image', 'height'), fld.get('XIM'), fld.get('YIM')
    elif 'XRI' in fld or 'YRI' in fld:
        fkey, fx, fy = ('image', 'real height'), fld.get('XRI'), fld.get('YRI')
    else:
        fkey, fx, fy = ('object', 'angle'), [0.0], [0.0]
    n_f = max(len(fx or []), len(fy or []))
    fx = (fx or [0.0]*n_f) + [0.0]*(n_f - len(fx or []))
    fy = (fy or [0.0]*n_f) + [0.0]*(n_f - len(fy or []))
    fields = []
    for k in range(n_f):
        v = {key: (vals[k] if k < len(vals) else 0.0) for key, vals in vig.items()}
        fields.append(M.Field(x=fx[k], y=fy[k], **v))
    max_f = max((abs(a) for a in fx + fy), default=0.0)
    # the thickness on the SI line is the defocus from the image surface
    osp = OpticalSpecs(WvlSpec(wvls, ref_wl), PupilSpec(*pupil), FieldSpec(fkey, max_f, fields),
                       FocusRange(surfs[-1]['thi'], 0.0))
    opm = M.OpticalModel(sm, osp, name=title or str(path).rsplit('/', 1)[-1])
    opm.dimensions = units
    apply_wide_angle_rule(opm)
    # cmdproc.py:196-201: any CIR input turns the automatic apertures off, the interfaces without
    # one are then set once, around the imported ones (cmdproc.py:88-94)
    given = [i for i, s_ in enumerate(surfs) if s_['cir'] is not None]
    sm.do_apertures, sm.input_ca_list = (not given), (given or None)
    if do_update:
        opm.update_optical_properties(bundle_fn)
    return opm


def apply_wide_angle_rule(opm):
    """what both importers of the reference do last (cmdproc.py:210, zmxread.py:288):
    ``fov.is_wide_angle = fov.check_is_wide_angle()``"""
    osp = opm.optical_spec
    fov = osp.field_of_view
    wide = fov.check_is_wide_angle(optical_spec=osp)
    if wide != bool(fov.is_wide_angle):
        fov.is_wide_angle = wide
        opm.update_model()
    return wide
