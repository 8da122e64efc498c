"""Real entrance pupil of wide-angle fields: the search of the reference's
``raytr/wideangle.py`` (``find_real_enp`` rev1, ``find_edge``, ``find_z_enp_on_interval``;
/root/reference/src/rayoptics/raytr/wideangle.py:46-93,105-330,333-352,355-446).

Fisheye lenses have strong pupil aberration: the chief ray of an oblique field crosses the
axis far from the paraxial entrance pupil.  The reference parameterises the pupil position by
its z offset from the first interface (``z_enp``), samples it from the paraxial value towards
the first vertex until rays get through and straddle the stop centre, and then iterates
(scipy secant ``newton``, ``brentq`` as the fallback) to the z for which the central ray hits
the stop at height 0.  ``fld.aim_info = z_enp`` then feeds the wide-angle start rays
(``opticalspec.ray_start_from_osp``, ``rt_pupil_kind`` 3 on the device).

Same decisions, same arithmetic, same scipy solvers as the reference -- the result is the same
double (``tests/test_trace_drivers.py::test_find_real_enp_is_the_references``).  Every ray goes
through ``trace_fn(seq_model, pt0, dir0, wvl, **kw) -> (ray, op, wvl)``: the drop-in GPU
``raytrace.trace`` by default, any callable with that signature in tests.
"""
from __future__ import annotations

import warnings

import numpy as np

from .opticalspec import rot_v1_into_v2


def solver_gave_up(e):
    from .vigcalc import solver_gave_up as f
    return f(e)


def _is_fuzzy_zero(a):
    return abs(a) < 1e-14          # util/misc_math.py: is_fuzzy_zero


def _default_trace_fn():
    from . import raytrace as RT
    return RT.trace


def _trace_errors():
    from .raytrace import TraceError, TraceMissedSurfaceError
    return TraceError, TraceMissedSurfaceError


def enp_z_coordinate(z_enp, seq_model, stop_idx, dir0, obj_dist, wvl, trace_fn=None):
    """wideangle.py:46-93: trace the ray through the centre of the pupil plane at ``z_enp``
    along ``dir0``; returns ``(point at the stop | zeros, RayResult(ray package, error | None))``."""
    from .trace import RayPkg, RayResult
    TraceError, _ = _trace_errors()
    trace_fn = _default_trace_fn() if trace_fn is None else trace_fn
    obj2enp_dist = (obj_dist + z_enp)
    pt1 = np.array([0., 0., obj2enp_dist])
    rot_mat = rot_v1_into_v2(np.array([0., 0., 1.]), dir0)
    pt0 = np.matmul(rot_mat, -pt1) + pt1
    try:
        pkg = RayPkg(*trace_fn(seq_model, pt0, dir0, wvl, intersect_obj=False))
    except TraceError as ray_error:
        pkg = ray_error.ray_pkg
        pkg = RayPkg(*pkg) if pkg is not None else RayPkg([], 0.0, wvl)
        return np.array([0., 0., 0.]), RayResult(pkg, ray_error)
    return pkg.ray[stop_idx][0], RayResult(pkg, None)


def find_edge(f, a, b, max_iter=3):
    """wideangle.py:333-352: binary search for the edge of the range where ``f`` evaluates"""
    fa = f(a)
    fb = f(b)
    for _ in range(max_iter):
        c = a + (b - a)/2
        fc = f(c)
        if fc is None:
            b = c
            fb = fc
        else:
            a = c
            fa = fc
    if fb is None:
        return a, fa
    return b, fb


def find_z_enp_on_interval(opt_model, stop_idx, start_z, end_z, z_estimate, fld, wvl, trace_fn):
    """wideangle.py:355-446: iterate ``z_enp`` until the ray crosses the stop at height 0.
    Returns ``(start_coords, RayResult of the last evaluation, converged)``."""
    from scipy.optimize import newton, brentq
    TraceError, _ = _trace_errors()
    sm, osp = opt_model['seq_model'], opt_model['optical_spec']
    fod = opt_model['analysis_results']['parax_data'].fod
    pt0, dir0 = osp.obj_coords(fld)
    last = {}

    def eval_z_enp(z_enp, *args):
        final_coord, rr = enp_z_coordinate(z_enp, sm, stop_idx, dir0, fod.obj_dist, wvl, trace_fn)
        last['rr'] = rr
        return final_coord[1] - 0.0

    if stop_idx is None:                       # floating stop: paraxial entrance pupil
        return np.array([0., 0., fod.enp_dist]), None, True
    z_enp, results = z_estimate, None
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            z_enp, results = newton(eval_z_enp, z_enp, rtol=1e-7, disp=False, full_output=True)
        except (RuntimeError, TraceError) as e:
            if isinstance(e, RuntimeError) and not solver_gave_up(e):
                raise                      # an engine failure, not an iteration that gave up
            z_enp = results.root if results is not None else z_enp
        converged = bool(results.converged) if results is not None else False
        ht_at_stop = last['rr'].pkg.ray[stop_idx][0][1]
        if abs(ht_at_stop) < 1e-6:
            converged = True
        start_coords = np.array([0., 0., z_enp])
        if not converged:
            try:
                z_enp, results = brentq(eval_z_enp, start_z, end_z, rtol=1e-7, disp=False,
                                        full_output=True)
            except RuntimeError as e:
                if not solver_gave_up(e):
                    raise
                z_enp = results.root
            start_coords = np.array([0., 0., z_enp])
            converged = bool(results.converged)
    return start_coords, last['rr'], converged


class _PupilScan:
    """Phase 1 of the real-pupil search: walk z from the paraxial pupil in steps of 1/16 of its
    distance, remembering the first, previous and latest pupil positions whose ray reached the stop
    (``first`` / ``prev`` / ``last``: ``(z, height at the stop)``).  The walk turns around at most
    once -- when the second hit is farther from the stop centre than the first, or when rays
    stop getting through -- and ends on a sign change between consecutive hits, on a failure
    after the turn, after two misses of the first surface, or after 64 samples."""

    MAX_SAMPLES = 64

    def __init__(self, z_paraxial):
        self.home = z_paraxial
        self.step = -z_paraxial/16
        self.z = z_paraxial
        self.first = self.prev = self.last = None
        self.hits = self.samples = self.front_misses = 0
        self.turned = self.finished = False

    def active(self):
        return not self.finished and self.samples < self.MAX_SAMPLES and self.front_misses < 2

    def _restart_reversed(self):
        self.step = -self.step
        self.z = self.home

    def _turn_around(self):
        self._restart_reversed()
        self.turned = True
        self.first, self.last = self.last, self.first

    def hit(self, height):
        self.hits += 1
        if self.first is None:
            self.first = (self.z, height)
        self.prev, self.last = self.last, (self.z, height)
        if self.hits > 1 and self.prev[1]*self.last[1] < 0:
            self.finished = True                         # the stop centre lies between the last two
        if self.hits == 2 and abs(self.first[1]) < abs(self.last[1]) and not self.turned:
            self._turn_around()                          # walking away from the centre

    def miss(self, err, missed_surface_type):
        if isinstance(err, missed_surface_type) and err.surf == 1:
            self._restart_reversed()
            self.front_misses += 1
        if self.first is not None:
            if self.turned:
                self.finished = True
            else:
                self._turn_around()

    def advance(self):
        self.z += self.step
        if _is_fuzzy_zero(self.z):                       # never sample the first vertex itself
            self.z = self.step/10
        self.samples += 1


def _bracket_stop_centre(scan, probe):
    """Phase 2: from the scan's hits to an interval ``(a, b)`` for the root finder, plus the two
    samples ``(lo, hi)`` the secant estimate is built from.  ``probe(z) -> (height | None, ray
    result)``.  Returns ``(a, b, lo, hi, None)``, or ``(None, None, None, None, (z, ray result))``
    when no ray through the stop centre exists (the answer is then the scan's last hit)."""
    (z_a, h_a), (z_b, h_b) = scan.first, scan.last
    step = scan.step
    if z_a == z_b:                                       # a single hit: resample around it, finer
        lo = hi = None
        for z in np.linspace(z_a - step, z_b + step, num=8):
            h, _ = probe(z)
            if h is not None:
                lo = (z, h) if lo is None else lo
                hi = (z, h)
        return lo[0], hi[0], lo, hi, None
    if h_a*h_b < 0:                                      # the scan already straddles the centre
        if scan.prev is not None and scan.prev[1]*h_b < 0:
            return scan.prev[0], z_b, scan.prev, scan.last, None
        return z_a, z_b, scan.first, scan.last, None
    # no sign change among the hits: the centre may sit between a hit and the edge of the beam
    height_only = lambda z: probe(z)[0]                  # noqa: E731
    edge_b = find_edge(height_only, z_b, z_b + step, max_iter=6)
    if edge_b[1]*h_b < 0:
        return z_b, edge_b[0], scan.last, edge_b, None
    edge_a = find_edge(height_only, z_a, z_a - step, max_iter=6)
    if edge_a[1]*h_a < 0:
        return z_a, edge_a[0], scan.first, edge_a, None
    _, rr = probe(edge_a[0] + (edge_b[0] - edge_a[0])/2)
    return None, None, None, None, (z_b, rr)


def find_real_enp(opt_model, stop_idx, fld, wvl, trace_fn=None):
    """z position, relative to the first interface, of the real entrance pupil of ``fld``
    (``find_real_enp`` / ``find_real_enp_rev1``, wideangle.py:86-312): the scan and bracketing above,
    a secant estimate from the bracketing samples, then ``find_z_enp_on_interval``.  Same samples,
    same decisions, same arithmetic as the reference (tests/test_trace_drivers.py drives both with
    scripted rays through every branch).  Returns ``(z_enp, RayResult of the last ray)``."""
    _, TraceMissedSurfaceError = _trace_errors()
    trace_fn = _default_trace_fn() if trace_fn is None else trace_fn
    sm, osp = opt_model['seq_model'], opt_model['optical_spec']
    fod = opt_model['analysis_results']['parax_data'].fod
    stop_idx = 1 if stop_idx is None else stop_idx
    _, dir0 = osp.obj_coords(fld)

    def probe(z):
        coord, rr = enp_z_coordinate(z, sm, stop_idx, dir0, fod.obj_dist, wvl, trace_fn)
        return (coord[1] if rr.err is None else None), rr

    def height_at(z):                                    # zeros when the ray failed, like the reference
        coord, rr = enp_z_coordinate(z, sm, stop_idx, dir0, fod.obj_dist, wvl, trace_fn)
        return coord[1], rr

    if fld.aim_info is not None:                         # a stored pupil position that still holds
        h, rr = height_at(fld.aim_info)
        if abs(h) < 1.48e-08:
            return fld.aim_info, rr
    if dir0[2] == 1:                                     # axial field: the paraxial pupil
        return fod.enp_dist, height_at(fod.enp_dist)[1]

    scan = _PupilScan(fod.enp_dist)
    while scan.active():
        h, rr = probe(scan.z)
        if rr.err is None:
            scan.hit(h)
        else:
            scan.miss(rr.err, TraceMissedSurfaceError)
        scan.advance()

    a, b, lo, hi, no_centre = _bracket_stop_centre(scan, probe)
    if no_centre is not None:
        return no_centre
    rise = hi[1] - lo[1]
    z_estimate = lo[0] if _is_fuzzy_zero(rise) else lo[0] - ((hi[0] - lo[0])/rise)*lo[1]
    start_coords, rr, _ = find_z_enp_on_interval(opt_model, stop_idx, a, b, z_estimate, fld, wvl,
                                                 trace_fn)
    return start_coords[2], rr


def find_z_enp(opt_model, stop_idx, z_enp_0, fld, wvl, trace_fn=None, **kwargs):
    """wideangle.py:559-617: secant iteration of ``z_enp`` from the estimate ``z_enp_0`` (which
    must give a ray that reaches the stop); ``(start_coords, RayResult, scipy results)``."""
    from scipy.optimize import newton
    TraceError, _ = _trace_errors()
    sm, osp = opt_model['seq_model'], opt_model['optical_spec']
    fod = opt_model['analysis_results']['parax_data'].fod
    pt0, dir0 = osp.obj_coords(fld)
    last = {'rr': None}

    def eval_z_enp(z_enp):
        final_coord, last['rr'] = enp_z_coordinate(z_enp, sm, stop_idx, dir0, fod.obj_dist, wvl,
                                                   trace_fn)
        return final_coord[1] - 0.
    if stop_idx is None:
        return np.array([0., 0., fod.enp_dist]), None, None
    z_enp, results = z_enp_0, None
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            z_enp, results = newton(eval_z_enp, z_enp, rtol=1e-7, disp=False, full_output=True)
        except (RuntimeError, TraceError) as e:
            if isinstance(e, RuntimeError) and not solver_gave_up(e):
                raise
            z_enp = results.root if results is not None else z_enp
    return np.array([0., 0., z_enp]), last['rr'], results


def eval_real_image_ht(opt_model, fld, wvl, trace_raw_fn=None):
    """Object-space chief ray of a field given as a REAL image height (wideangle.py:620-664): the
    chief ray is traced backwards, from the image point through the centre of the stop (iterated
    on the reverse path, ``vigcalc.iterate_ray_raw``), and leaves the first surface towards the
    object.  Returns ``((object point, direction), z_enp)`` -- the implementation of
    ``obj_coords`` for ('image', 'real height') fields."""
    from . import vigcalc
    sm, osp = opt_model['seq_model'], opt_model['optical_spec']
    fov = osp['fov']
    fod = opt_model['analysis_results']['parax_data'].fod
    not_wa = not fov.is_wide_angle
    stop_idx = 1 if sm.stop_surface is None else sm.stop_surface
    ifcx = len(sm.ifcs) - stop_idx - 1                    # the stop, counted from the image
    rpath_list = list(sm.reverse_path(wl=wvl, start=len(sm.ifcs), stop=None, step=-1))
    obj2pup_dist = fod.exp_dist - fod.img_dist            # image surface -> exit pupil
    p_exp = np.array([0, 0, obj2pup_dist])
    p_i = np.array([fld.x, fld.y, 0])
    if fov.is_relative:
        p_i = p_i*fov.value
    v = p_exp - p_i
    d_i = v/np.linalg.norm(v)
    start_coords, rrev_cr = vigcalc.iterate_ray_raw(rpath_list, ifcx, [0., 0.], p_i, d_i, obj2pup_dist,
                                                    fod.exp_radius, wvl, not_wa,
                                                    trace_raw_fn=trace_raw_fn)
    ray = rrev_cr[0][0]
    p_k, d_k = ray[-2][0], ray[-2][1]                     # at the first surface, heading back
    p_k01 = np.sqrt(p_k[0]**2 + p_k[1]**2)
    d_o = -d_k
    d_k01 = np.sqrt(d_k[0]**2 + d_k[1]**2)
    z_enp = fod.enp_dist if d_k01 == 0. else p_k[2] + p_k01*d_o[2]/d_k01
    p_o = ray[-1][0]
    if osp.conjugate_type('object') == 'infinite':
        obj2enp_dist = fod.obj_dist + z_enp
        enp_pt = np.array([0., 0., obj2enp_dist])
        p_o = enp_pt + obj2enp_dist*d_k
    return (p_o, d_o), z_enp


def eval_z_enp_curve(opm, printout=True, trace_fn=None, num_fields=21):
    """the z position of the real entrance pupil across the field of view (wideangle.py:667-705,
    ('object', 'angle') fields): ``(fields, object angles, image heights, z_enps)``."""
    import math
    sm, osp = opm['seq_model'], opm['optical_spec']
    fov = osp['fov']
    if tuple(fov.key) != ('object', 'angle'):
        raise NotImplementedError(f'field type {fov.key}: only (object, angle) fields')
    save_is_relative = fov.is_relative
    save_fields = [(f.x, f.y) for f in fov.fields]
    if not save_is_relative:                      # the fields are re-read as fractions
        fov.is_relative = True
    flds, z_enps, obj_angs, img_hts = [], [], [], []
    cwl = osp['wvls'].central_wvl
    if printout:
        print('frac fld     obj angle     img ht      z_enp')
    try:
        for fld_ht in np.linspace(0, 1, num_fields):
            fld = fov.new_field(y=fld_ht)
            z_enp, cr_rr = find_real_enp(opm, sm.stop_surface, fld, cwl, trace_fn)
            cr_ray = cr_rr.pkg.ray
            d0, img_ht = cr_ray[0][1], cr_ray[-1][0]
            ang_y = np.rad2deg(math.atan2(d0[1], d0[2]))
            if printout:
                print(f'{fld.yf:7.2f}     {ang_y:9.3f}     {img_ht[1]:7.2f}    {z_enp:8.4f}')
            flds.append(fld)
            obj_angs.append(ang_y)
            img_hts.append(img_ht[1])
            z_enps.append(z_enp)
    finally:
        fov.is_relative = save_is_relative
    return flds, obj_angs, img_hts, z_enps


def aim_wide_angle_fields(opt_model, wvl=None, trace_fn=None):
    """``aim_chief_ray`` (raytr/trace.py:627-640) for every field of a wide-angle specification:
    sets ``fld.aim_info = z_enp``; returns the list."""
    sm, osp = opt_model['seq_model'], opt_model['optical_spec']
    wvl = sm.central_wavelength() if wvl is None else wvl
    out = []
    for fld in osp['fov'].fields:
        z_enp, _ = find_real_enp(opt_model, sm.stop_surface, fld, wvl, trace_fn)
        fld.aim_info = float(z_enp)
        out.append(fld.aim_info)
    return out
