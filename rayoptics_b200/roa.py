"""Self-contained reader of the reference's ``.roa`` model files.

``.roa`` is a json_tricks dump of the whole ``OpticalModel``
(/root/reference/src/rayoptics/optical/opticalmodel.py:238-264,
gui/roafile.py:61-123).  Plain ``json`` reads it: json_tricks only adds
``__instance_type__`` / ``__ndarray__`` wrappers.  This reader extracts what the
ray-trace path needs -- interfaces, profiles, gaps, media, z_dir, stop, optical
specification with stored aim points -- without ``opticalglass`` /
``json_tricks`` / ``anytree``.

Dispersion formulas restated from the published catalog conventions:
Schott / Ohara Sellmeier-3, Hoya / Hikari power series, and the Buchdahl
2-term model the reference stores for ``ModelGlass`` entries.
"""
from __future__ import annotations

import json
import math

from . import model as M
from .opticalspec import OpticalSpecs, WvlSpec, PupilSpec, FieldSpec, FocusRange


class PowerSeries(M.Medium):
    """n**2 = sum c_i * l**e_i (l in um): Hoya (6 terms) / Hikari (9 terms)."""
    EXPONENTS = {6: (0, 2, -2, -4, -6, -8), 9: (0, 2, 4, -2, -4, -6, -8, -10, -12)}

    def __init__(self, coefs=None, label=''):
        self.coefs = list(coefs)
        self.label = label

    def rindex(self, wvl):
        l = wvl*1.0e-3
        n2 = 0.0
        for c, e in zip(self.coefs, self.EXPONENTS[len(self.coefs)]):
            n2 += c*l**e
        return math.sqrt(n2)


class Buchdahl(M.Medium):
    """n = n0 + v1*w + v2*w**2, w = dl/(1 + 2.5 dl), dl = l - l0 (um)."""

    def __init__(self, rind0=1.5, wv0=0.5875618, coefs=(0.0, 0.0), label=''):
        self.rind0, self.wv0, self.coefs, self.label = rind0, wv0, list(coefs), label

    def rindex(self, wvl):
        dl = wvl*1.0e-3 - self.wv0
        om = dl/(1 + 2.5*dl)
        return self.rind0 + self.coefs[0]*om + self.coefs[1]*om**2


M._MEDIUM_CLASSES.update({'PowerSeries': PowerSeries, 'Buchdahl': Buchdahl})


def _val(x):
    """unwrap json_tricks ndarray wrappers"""
    if isinstance(x, dict) and '__ndarray__' in x:
        return x['__ndarray__']
    return x


def _medium(m):
    kind = m['__instance_type__'][1]
    a = m.get('attributes', {})
    if kind == 'Air':
        return M.Air()
    if kind in ('SchottGlass', 'OharaGlass', 'CDGMGlass', 'SumitaGlass'):
        return M.Sellmeier(_val(a['coefs']), label=a.get('gname', ''))
    if kind in ('HoyaGlass', 'HikariGlass'):
        return PowerSeries(_val(a['coefs']), label=a.get('gname', ''))
    if kind == 'ModelGlass':
        b = a.get('bdhl_model')
        if b is not None:
            ba = b['attributes']
            return Buchdahl(ba['rind0'], ba['wv0'], _val(ba['coefs']), label=a.get('label', ''))
        return M.AbbeGlass(a['n'], a['v'], label=a.get('label', ''))
    if kind in ('ConstantIndex', 'Medium'):
        return M.ConstantIndex(a.get('n', 1.0), label=a.get('label', ''))
    if kind == 'InterpolatedMedium':
        return M.TableIndex(_val(a['wvls']), _val(a['rndx']), label=a.get('label', ''))
    raise NotImplementedError(f'.roa medium type {kind}')


def _profile(p):
    kind = p['__instance_type__'][1]
    a = p['attributes']
    if kind == 'Spherical':
        return M.Spherical(c=a['cv'])
    if kind == 'Conic':
        return M.Conic(c=a['cv'], cc=a['cc'])
    if kind == 'EvenPolynomial':
        return M.EvenPolynomial(c=a['cv'], cc=a['cc'], coefs=_val(a.get('coefs', [])))
    if kind == 'RadialPolynomial':
        return M.RadialPolynomial(c=a['cv'], ec=a['ec'], coefs=_val(a.get('coefs', [])))
    if kind in ('YToroid', 'XToroid'):
        cls = M.YToroid if kind == 'YToroid' else M.XToroid
        return cls(c=a['cv'], cR=a['cR'], cc=a['cc'], coefs=_val(a.get('coefs', [])))
    raise NotImplementedError(f'.roa profile type {kind}')


def _aperture(c):
    kind = c['__instance_type__'][1]
    a = c['attributes']
    kw = dict(x_offset=a.get('x_offset', 0.0), y_offset=a.get('y_offset', 0.0),
              rotation=a.get('rotation', 0.0), is_obscuration=a.get('is_obscuration', False))
    if kind == 'Circular':
        return M.Circular(radius=a['radius'], **kw)
    if kind == 'Rectangular':
        return M.Rectangular(a['x_half_width'], a['y_half_width'], **kw)
    if kind == 'Elliptical':
        return M.Elliptical(a['x_half_width'], a['y_half_width'], **kw)
    raise NotImplementedError(f'.roa aperture type {kind}')


def _phase_element(pe):
    """phase element of a Surface (oprops/doe.py json encodings)"""
    if pe is None:
        return None
    kind = pe['__instance_type__'][1]
    a = pe['attributes']
    if kind == 'DiffractiveElement':
        if a.get('phase_fct_name', 'radial_phase_fct') != 'radial_phase_fct':
            raise NotImplementedError(f".roa DiffractiveElement with phase_fct {a['phase_fct_name']}")
        return M.DiffractiveElement(label=a.get('label', ''), coefficients=_val(a['coefficients']),
                                    ref_wl=a['ref_wl'], order=a['order'])
    if kind == 'DiffractionGrating':
        return M.DiffractionGrating(label=a.get('label', ''), order=a['order'],
                                    grating_normal=_val(a['grating_normal']),
                                    grating_lpmm=a['_grating_lpmm'],
                                    interact_mode=a.get('interact_mode', 'transmit'))
    if kind == 'HolographicElement':
        return M.HolographicElement(a.get('label', ''), _val(a['ref_pt']), a['ref_virtual'],
                                    _val(a['obj_pt']), a['obj_virtual'], a['ref_wl'])
    raise NotImplementedError(f'.roa phase element {kind}')


def open_roa(path):
    """Read a ``.roa`` file into an ``OpticalModel`` mirror."""
    with open(path) as f:
        d = json.load(f)
    om = d['optical_model']['attributes']
    sm_a = om['seq_model']['attributes']
    pdict = om.get('profile_dict', {})
    ifcs = []
    for i in sm_a['ifcs']:
        kind = i['__instance_type__'][1]
        a = i['attributes']
        if kind == 'ThinLens':
            pe = a['phase_element']['attributes']
            hoe = M.HolographicElement(pe.get('label', ''), _val(pe['ref_pt']), pe['ref_virtual'],
                                       _val(pe['obj_pt']), pe['obj_virtual'], pe['ref_wl'])
            ifcs.append(M.ThinLens(lbl=a.get('label', ''), power=a['_power'],
                                   ref_index=a.get('ref_index', 1.5),
                                   max_aperture=a.get('max_aperture', 1.0),
                                   interact_mode=a['interact_mode'], phase_element=hoe))
            continue
        if kind != 'Surface':
            raise NotImplementedError(f'.roa interface type {kind}')
        phase_element = _phase_element(a.get('phase_element'))
        decenter = None
        if a.get('decenter') is not None:       # DecenterData, elem/surface.py:274-337 (json_tricks: vars())
            da = a['decenter'].get('attributes', a['decenter'])
            dec, eul = _val(da['dec']), _val(da['euler'])
            decenter = M.DecenterData(da.get('_dtype', da.get('dtype', 'decenter')), dec[0], dec[1],
                                      eul[0], eul[1], eul[2])
            decenter.dec[2] = dec[2]
        prf = a.get('profile')
        prf = pdict[str(a['profile_id'])] if prf is None else prf
        ifcs.append(M.Surface(lbl=a.get('label', ''), profile=_profile(prf),
                              interact_mode=a['interact_mode'],
                              max_aperture=a.get('max_aperture', 1.0),
                              clear_apertures=[_aperture(c) for c in a.get('clear_apertures', [])]))
        ifcs[-1].decenter = decenter
        if phase_element is not None:       # the reference tests hasattr(ifc, 'phase_element')
            ifcs[-1].phase_element = phase_element
    gaps = [M.Gap(g['attributes']['thi'], _medium(g['attributes']['medium']))
            for g in sm_a['gaps']]
    osp_a = om['optical_spec']['attributes']
    wv = osp_a['spectral_region']['attributes']
    pup = osp_a['pupil']['attributes']
    fov = osp_a['field_of_view']['attributes']
    foc = osp_a.get('defocus', {}).get('attributes', {})
    fields = []
    for f in fov['fields']:
        a = f['attributes']
        aim = a.get('aim_pt', a.get('aim_info'))
        fields.append(M.Field(x=a['x'], y=a['y'], wt=a.get('wt', 1.0),
                              vux=a.get('vux', 0.0), vuy=a.get('vuy', 0.0),
                              vlx=a.get('vlx', 0.0), vly=a.get('vly', 0.0),
                              aim_pt=None if aim is None else _val(aim)))
    key = pup.get('_key', pup.get('key'))
    fkey = fov.get('key', fov.get('_key'))
    osp = OpticalSpecs(WvlSpec(_val(wv['wavelengths']), wv.get('reference_wvl', 0),
                               _val(wv.get('spectral_wts'))),
                       PupilSpec(key, pup['value']),
                       FieldSpec(fkey, fov['value'], fields, fov.get('is_relative', False),
                                 fov.get('is_wide_angle', False)),
                       FocusRange(foc.get('focus_shift', 0.0), foc.get('defocus_range', 0.0)))
    sm = M.SequentialModel(ifcs, gaps, z_dir=sm_a.get('z_dir'),
                           stop_surface=sm_a.get('stop_surface'),
                           wvlns=osp.spectral_region.wavelengths,
                           ref_wvl=osp.spectral_region.reference_wvl)
    return M.OpticalModel(sm, osp, name=path.rsplit('/', 1)[-1])
