"""Optical specification (wavelengths, pupil, fields, focus) and start rays.

Mirrors what the hot path's callers read from the reference's
``OpticalSpecs`` (/root/reference/src/rayoptics/raytr/opticalspec.py:41-400):

* ``obj_coords(fld)``            opticalspec.py:990-1091 (non wide-angle cases)
* ``ray_start_from_osp(...)``    opticalspec.py:289-400  (single ray, numpy,
  same expression order as the reference so the start ray has the same bits)
* ``grid_fields(...)``           the per-field constants of the 'epd' branch
  (opticalspec.py:354-366) that the grid kernel turns into start rays on device.
"""
from __future__ import annotations

import collections
import math

import numpy as np

from .model import Field
from .firstorder import compute_first_order


def rot_v1_into_v2(v1, v2):
    """util/misc_math.py:132-148: "equivalent angle" rotation matrix built from the cross
    product of v1 and v2 -- every term as the reference writes it, including the signs of the
    off-diagonal sine terms (the wide-angle start rays are defined by THIS matrix)."""
    rot_axis = -np.cross(v1, v2)
    s = np.linalg.norm(rot_axis)
    c = np.dot(v1, v2)
    v = 1 - c
    ax = normalize(rot_axis)
    return np.array(
        [[ax[0]*ax[0]*v + c, ax[0]*ax[1]*v - ax[2]*s, ax[0]*ax[2]*v + ax[1]*s],
         [ax[0]*ax[1]*v + ax[2]*s, ax[1]*ax[1]*v + c, ax[1]*ax[2]*v + ax[0]*s],
         [ax[0]*ax[2]*v + ax[1]*s, ax[1]*ax[2]*v + ax[0]*s, ax[2]*ax[2]*v + c]])


def normalize(v):
    """util/misc_math.py:48-54"""
    length = np.linalg.norm(v)
    if length == 0.0:
        return v
    return v/length


class WvlSpec:
    def __init__(self, wavelengths=(550.0,), ref_wl=0, spectral_wts=None):
        self.wavelengths = list(wavelengths)
        self.reference_wvl = ref_wl
        self.spectral_wts = list(spectral_wts) if spectral_wts else [1.0]*len(self.wavelengths)
        self.render_colors = [None]*len(self.wavelengths)   # plotting detail of the reference

    @property
    def central_wvl(self):
        return self.wavelengths[self.reference_wvl]


class PupilSpec:
    """key = (object|image, epd|f/#|NA); opticalspec.py:626-760"""
    default_pupil_rays = [[0., 0.], [1., 0.], [-1., 0.], [0., 1.], [0., -1.]]
    default_ray_labels = ['00', '+X', '-X', '+Y', '-Y']

    def __init__(self, key=('object', 'epd'), value=1.0):
        self.key = tuple(key[-2:])
        self.value = value
        self.pupil_rays = [list(r) for r in self.default_pupil_rays]
        self.ray_labels = list(self.default_ray_labels)


class FieldSpec:
    """key = (object|image, angle|height); opticalspec.py:820-1195"""

    def __init__(self, key=('object', 'angle'), value=0.0, fields=None, is_relative=False,
                 is_wide_angle=False):
        self.key = tuple(key[-2:])
        self.value = value
        self.is_relative = is_relative
        self.is_wide_angle = is_wide_angle
        self.fields = list(fields) if fields else [Field()]
        for f in self.fields:
            f.fov = self

    def check_is_wide_angle(self, angle_threshold=45., optical_spec=None):
        """the importers' rule (opticalspec.py:896-905, applied by zmxread.py:288 and cmdproc.py:210):
        object angles beyond the threshold, or real image heights with the object at infinity"""
        if tuple(self.key) == ('image', 'real height'):
            return bool(optical_spec is not None and optical_spec.conjugate_type('object') == 'infinite')
        if tuple(self.key) == ('object', 'angle'):
            return self.max_field()[0] > angle_threshold
        return False

    def new_field(self, x=0.0, y=0.0, **kwargs):
        """a Field tied to this specification, not added to ``fields`` (opticalspec.py FieldSpec)"""
        return Field(x=x, y=y, fov=self, **kwargs)

    def max_field(self):
        """(magnitude of the maximum field, index of that field), opticalspec.py:1093-1109"""
        max_fld, max_fld_sqrd = 0, -1.0
        for i, f in enumerate(self.fields):
            fld_sqrd = f.x*f.x + f.y*f.y
            if fld_sqrd > max_fld_sqrd:
                max_fld_sqrd, max_fld = fld_sqrd, i
        max_fld_value = math.sqrt(max_fld_sqrd)
        if self.is_relative:
            max_fld_value *= self.value
        return max_fld_value, max_fld

    @property
    def index_labels(self):
        """'axis', ' 0.70y', ..., 'edge' (opticalspec.py:907-932, index_label_type 'auto')"""
        field_norm = 1 if (self.is_relative or self.value == 0) else 1.0/self.value
        labels = []
        for f in self.fields:
            fldx = '{:5.2f}x'.format(field_norm*f.x) if f.x != 0.0 else ''
            fldy = '{:5.2f}y'.format(field_norm*f.y) if f.y != 0.0 else ''
            labels.append(fldx + fldy)
        labels[0] = 'axis'
        if len(labels) > 1:
            labels[-1] = 'edge'
        return labels

    def max_field_value(self):
        if self.is_relative:
            return self.value
        m = 0.0
        for f in self.fields:
            m = max(m, math.sqrt(f.x*f.x + f.y*f.y))
        return m if m != 0.0 else self.value


class FocusRange:
    def __init__(self, focus_shift=0.0, defocus_range=0.0):
        self.focus_shift = focus_shift
        self.defocus_range = defocus_range

    def get_focus(self, fr=0.0):
        """opticalspec.py FocusRange.get_focus: focus for the fractional range position"""
        return self.focus_shift + fr*self.defocus_range


class OpticalSpecs:
    def __init__(self, wvls=None, pupil=None, fov=None, focus=None, do_aiming=True):
        self.spectral_region = wvls or WvlSpec()
        self.pupil = pupil or PupilSpec()
        self.field_of_view = fov or FieldSpec()
        self.defocus = focus or FocusRange()
        self.do_aiming = do_aiming
        self.opt_model = None

    # reference access keys (opticalspec.py:78-95)
    _keys = {'wvls': 'spectral_region', 'pupil': 'pupil', 'fov': 'field_of_view',
             'focus': 'defocus'}

    def __getitem__(self, key):
        return getattr(self, self._keys[key])

    @property
    def fov(self):
        return self.field_of_view

    def update_model(self, **kwargs):
        sm = self.opt_model.seq_model
        fod = compute_first_order(sm, self, self.spectral_region.central_wvl)
        self.opt_model.analysis_results['parax_data'] = ParaxData(fod.ax_ray, fod.pr_ray, fod)

    @property
    def fod(self):
        return self.opt_model.analysis_results['parax_data'].fod

    def obj_img_rindex(self):
        return self.fod.n_obj, self.fod.n_img

    def lookup_fld_wvl_focus(self, fi, wl=None, fr=0.0):
        """field, wavelength (nm) and focus shift for the indices (opticalspec.py:403-427)"""
        wvl = (self.spectral_region.central_wvl if wl is None
               else self.spectral_region.wavelengths[wl])
        return self.field_of_view.fields[fi], wvl, self.defocus.get_focus(fr)

    def conjugate_type(self, space='object'):
        sm = self.opt_model.seq_model
        thi = sm.gaps[0].thi if space == 'object' else sm.gaps[-1].thi
        return 'infinite' if abs(thi) >= 1e8 or math.isinf(thi) else 'finite'

    # ---------------------------------------------------------- obj_coords
    def obj_coords(self, fld):
        """(pt, dir) characterising `fld` in object space, opticalspec.py:990-1091."""
        fov = self.field_of_view
        obj_img_key, value_key = fov.key
        fld_coord = np.array([fld.x, fld.y, 0.0])
        rel_fld_coord = np.array([fld.x, fld.y, 0.0])
        if fov.is_relative:
            fld_coord *= fov.value
        else:
            if fov.value != 0:
                rel_fld_coord /= fov.value
        fod = self.fod
        obj2enp_dist = fod.obj_dist + fod.enp_dist
        pt1 = np.array([0., 0., obj2enp_dist])
        if obj_img_key == 'image' and value_key == 'real height':
            # opticalspec.py:1018-1034,1060-1075: the real chief ray traced back from the image
            # point; the pupil offset it implies becomes the field's aim info (side effect kept)
            from .wideangle import eval_real_image_ht
            (obj_pt, obj_dir), z_enp = eval_real_image_ht(self.opt_model, fld,
                                                          self.spectral_region.central_wvl,
                                                          trace_raw_fn=getattr(self, '_trace_raw_fn', None))
            if fov.is_wide_angle:
                fld.aim_info = z_enp
            else:
                del_z = fod.enp_dist - z_enp
                if abs(obj_dir[2]) < 1e-14:
                    fld.aim_info = np.array([0., 0.])
                else:
                    fld.aim_info = del_z*np.array([obj_dir[0]/obj_dir[2], obj_dir[1]/obj_dir[2]])
            return obj_pt, obj_dir
        if self.conjugate_type('object') == 'infinite':
            if obj_img_key == 'image':
                max_field_ang = math.atan(fod.pr_slp0)
                fld_angle = max_field_ang*rel_fld_coord
            elif value_key == 'angle':
                fld_angle = np.deg2rad(fld_coord)
            else:
                obj_pt = fld_coord
                return obj_pt, normalize(pt1 - obj_pt)
            ang_x, ang_y = fld_angle[0], fld_angle[1]
            dir_cos = np.array([math.sin(ang_x)*math.cos(ang_y),
                                math.sin(ang_y),
                                math.cos(ang_x)*math.cos(ang_y)])
            if fov.is_wide_angle:               # opticalspec.py:1041-1043
                rot_mat = rot_v1_into_v2(np.array([0., 0., 1.]), dir_cos)
                obj_pt = np.matmul(rot_mat, -pt1) + pt1
            else:
                obj_pt = obj2enp_dist*np.array([dir_cos[0]/dir_cos[2], dir_cos[1]/dir_cos[2], 0.0])
            return obj_pt, dir_cos
        # finite conjugates
        if obj_img_key == 'image':
            obj_pt = fod.pr_ht0*rel_fld_coord
        elif value_key == 'angle':
            fld_angle = np.deg2rad(fld_coord)
            obj_dir = np.sin(fld_angle)
            obj_dir[2] = np.sqrt(1 - obj_dir[0]**2 - obj_dir[1]**2)
            obj_pt = obj2enp_dist*np.array([obj_dir[0]/obj_dir[2], obj_dir[1]/obj_dir[2], 0.0])
            return obj_pt, obj_dir
        else:
            obj_pt = fld_coord
        return obj_pt, normalize(pt1 - obj_pt)

    # -------------------------------------------------- single start ray
    def _epd_pupil(self):
        """(pupil_oi_key, pupil_value_key, pupil_value) after the image-space substitution of
        opticalspec.py:311-325"""
        return effective_pupil(self.opt_model)

    def ray_start_from_osp(self, pupil, fld, pupil_type='rel pupil'):
        """(pt0, dir0) for one ray, opticalspec.py:289-400."""
        pupil_oi_key, pupil_value_key, pupil_value = self._epd_pupil()
        n_obj, n_img = self.obj_img_rindex()
        p0, d0 = self.obj_coords(fld)
        fod = self.fod
        aim_info = fld.aim_info if getattr(fld, 'aim_info', None) is not None else None
        z_enp = fod.enp_dist
        if 'epd' == pupil_value_key:
            if pupil_type == 'aim pt':
                pt0 = p0
                pt1 = np.array([pupil[0], pupil[1], fod.obj_dist + z_enp])
            elif self.field_of_view.is_wide_angle:      # opticalspec.py:342-358
                eprad = pupil_value/2
                # the pupil is normal to the chief ray: pupil point rotated into surface-1 coordinates
                pupil_pt = eprad*np.array([pupil[0], pupil[1], 0.])
                rot_mat_d2s = rot_v1_into_v2(d0, np.array([0., 0., 1.]))
                pt1 = np.matmul(rot_mat_d2s, pupil_pt)
                if aim_info is not None:
                    z_enp = aim_info                    # real entrance pupil position of this field
                obj2enp_dist = -(fod.obj_dist + z_enp)
                if self.conjugate_type('object') == 'infinite':
                    enp_pt = np.array([0., 0., obj2enp_dist])
                    rot_mat_s2d = rot_v1_into_v2(np.array([0., 0., 1.]), d0)
                    pt0 = np.matmul(rot_mat_s2d, enp_pt) - enp_pt
                else:
                    pt0 = p0
                pt1[2] -= obj2enp_dist
            else:
                eprad = pupil_value/2
                aim_pt = [0., 0.] if aim_info is None else aim_info
                obj2enp_dist = -(fod.obj_dist + z_enp)
                pt1 = np.array([eprad*pupil[0] + aim_pt[0],
                                eprad*pupil[1] + aim_pt[1],
                                fod.obj_dist + z_enp])
                pt0 = obj2enp_dist*np.array([d0[0]/d0[2], d0[1]/d0[2], 0.])
            dir0 = normalize(pt1 - pt0)
        else:
            if pupil_type == 'aim dir':
                dir_tot = pupil
                pt0 = p0
            else:
                if 'NA' in pupil_value_key:
                    n = n_obj if pupil_oi_key == 'object' else n_img
                    sin_ang = pupil_value/n
                    pupil_dir = sin_ang*np.array([pupil[0], pupil[1]])
                else:
                    slope = -1/(2*pupil_value)
                    hypt = np.sqrt(1 + (pupil[0]*slope)**2 + (pupil[1]*slope)**2)
                    pupil_dir = np.array([slope*pupil[0]/hypt, slope*pupil[1]/hypt])
                pt0 = p0
                cr_dir = d0[:2]
                dir_tot = pupil_dir + cr_dir
            dir0 = np.array([dir_tot[0], dir_tot[1], np.sqrt(1 - np.dot(dir_tot, dir_tot))])
        return pt0, dir0

    # ------------------------------------------------- grid field records
    def grid_fields(self, fields=None):
        """see ``grid_fields_of``"""
        return grid_fields_of(self.opt_model, fields)

    # ------------------------------------------------------- persistence
    def to_dict(self):
        return {'wvls': self.spectral_region.wavelengths,
                'ref_wvl': self.spectral_region.reference_wvl,
                'pupil': {'key': list(self.pupil.key), 'value': self.pupil.value},
                'fov': {'key': list(self.field_of_view.key), 'value': self.field_of_view.value,
                        'is_relative': self.field_of_view.is_relative,
                        **({'is_wide_angle': True} if self.field_of_view.is_wide_angle else {}),
                        'fields': [f.to_dict() for f in self.field_of_view.fields]},
                'focus_shift': self.defocus.focus_shift}

    @classmethod
    def from_dict(cls, d):
        fov = d['fov']
        fields = [Field(**f) for f in fov['fields']]
        return cls(WvlSpec(d['wvls'], d.get('ref_wvl', 0)),
                   PupilSpec(d['pupil']['key'], d['pupil']['value']),
                   FieldSpec(fov['key'], fov['value'], fields, fov.get('is_relative', False),
                             is_wide_angle=fov.get('is_wide_angle', False)),
                   FocusRange(d.get('focus_shift', 0.0)))


# --- the same derivations on ANY optical model that follows the reference's interface
#     (the reference's own OpticalModel or the mirror): used by the batched drivers when they
#     are installed inside the reference (raytrace.install(batched=True)) --------------------
def effective_pupil(opt_model):
    """(pupil_oi_key, pupil_value_key, pupil_value) after the image-space substitution of
    ``ray_start_from_osp`` (opticalspec.py:311-325)"""
    osp = opt_model['optical_spec']
    fod = opt_model['analysis_results']['parax_data'].fod
    pupil_oi_key, pupil_value_key = osp['pupil'].key
    pupil_value = osp['pupil'].value
    if pupil_oi_key == 'image':
        if abs(fod.m) < 1e-10:
            pupil_value_key, pupil_value = 'epd', 2*fod.enp_radius
        elif abs(fod.enp_dist) > 1e10:      # telecentric entrance pupil
            pupil_value_key = 'NA'
            n_obj = osp.obj_img_rindex()[0]
            slp0 = fod.obj_na/n_obj                                  # etendue.na2slp_parax
            pupil_value = n_obj*math.sin(math.atan(slp0/n_obj))      # etendue.slp2na
        else:
            pupil_value_key, pupil_value = 'epd', 2*fod.enp_radius
    return pupil_oi_key, pupil_value_key, pupil_value


def grid_fields_of(opt_model, fields=None):
    """Per-field constants of the start-ray generation for the grid kernels: list of dicts
    (pt0, aim, vlx, vux, vly, vuy, pupil_kind) plus (scale, z_pupil).

    'epd' pupils (``rt_pupil_kind`` 0): pt0 = the ray origin in the object plane, aim = aim
    point, scale = entrance pupil radius.  Angular pupils ('NA' 1, 'f/#' 2;
    opticalspec.py:368-398): pt0 = object point, aim = chief-ray direction cosines ``d0[:2]``,
    scale = NA/n or -1/(2 f/#)."""
    osp = opt_model['optical_spec']
    fod = opt_model['analysis_results']['parax_data'].fod
    pupil_oi_key, pupil_value_key, pupil_value = effective_pupil(opt_model)
    z_pupil = fod.obj_dist + fod.enp_dist
    flds = fields if fields is not None else osp['fov'].fields
    out = []
    if pupil_value_key == 'epd' and osp['fov'].is_wide_angle:
        # rt_pupil_kind 3 (opticalspec.py:342-358): per field the start point, the matrix that
        # rotates the pupil plane normal to the chief ray into surface-1 coordinates, and the
        # distance object -> real entrance pupil of that field (aim_info = z_enp)
        kind, scale = 3, pupil_value/2
        infinite = osp.conjugate_type('object') == 'infinite'
        for fld in flds:
            p0, d0 = osp.obj_coords(fld)
            z_enp = fod.enp_dist if getattr(fld, 'aim_info', None) is None else float(fld.aim_info)
            obj2enp_dist = -(fod.obj_dist + z_enp)
            rot_d2s = rot_v1_into_v2(d0, np.array([0., 0., 1.]))
            if infinite:
                enp_pt = np.array([0., 0., obj2enp_dist])
                pt0 = np.matmul(rot_v1_into_v2(np.array([0., 0., 1.]), d0), enp_pt) - enp_pt
            else:
                pt0 = p0
            out.append({'pt0': pt0, 'aim': [0., 0.], 'rot': np.ascontiguousarray(rot_d2s).reshape(9),
                        'obj2enp': obj2enp_dist})
    elif pupil_value_key == 'epd':
        kind, scale = 0, pupil_value/2
        obj2enp_dist = -(fod.obj_dist + fod.enp_dist)
        for fld in flds:
            p0, d0 = osp.obj_coords(fld)
            pt0 = obj2enp_dist*np.array([d0[0]/d0[2], d0[1]/d0[2], 0.])
            aim = [0., 0.] if getattr(fld, 'aim_info', None) is None else fld.aim_info
            out.append({'pt0': pt0, 'aim': [float(aim[0]), float(aim[1])]})
    else:
        n_obj, n_img = osp.obj_img_rindex()
        if 'NA' in pupil_value_key:
            n = n_obj if pupil_oi_key == 'object' else n_img
            kind, scale = 1, pupil_value/n
        else:
            kind, scale = 2, -1/(2*pupil_value)
        for fld in flds:
            p0, d0 = osp.obj_coords(fld)
            out.append({'pt0': p0, 'aim': [float(d0[0]), float(d0[1])]})
    for rec, fld in zip(out, flds):
        rec.update(vlx=fld.vlx, vux=fld.vux, vly=fld.vly, vuy=fld.vuy, pupil_kind=kind)
    return out, scale, z_pupil


# parax/firstorder.py:29: ParaxData = namedtuple('ParaxData', ['ax_ray', 'pr_ray', 'fod'])
ParaxData = collections.namedtuple('ParaxData', ['ax_ray', 'pr_ray', 'fod'])