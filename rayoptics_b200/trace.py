"""Batched drivers with the call surface of ``rayoptics.raytr.trace``.

The reference's ``trace_fan`` / ``trace_grid`` (/root/reference/src/rayoptics/raytr/trace.py:537-605)
call ``trace_safe`` -> ``trace_base`` -> ``rt.trace`` once per pupil point and hand
every ray to a Python callback (``img_filter``).  Here all pupil points of one
(field, wavelength) go through ONE grid launch that writes whole rays; the result
containers (``RaySeg`` / ``RayPkg`` / ``RayResult``, raytr/__init__.py:24-40), the
``output_filter`` / ``rayerr_filter`` conventions of ``trace_safe`` (trace.py:159-219),
the accumulated pupil stepping and the shape of what is returned are the
reference's, so ``SequentialModel.trace_fan / trace_grid / trace_wavefront``
(seq/sequential.py:1006-1120) and the figure classes above them keep working.

The callbacks stay Python (they are user code); only the ray tracing is batched.
There is no CPU path: the default tracer is the CUDA engine.  ``tracer=`` is a seam
for tests (the CPU suite passes the oracle to check the host logic) -- same idea
as ``vigcalc``'s ``trace_fn``.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from . import engine as E
from . import waveabr as W
from .opticalspec import grid_fields_of

RayResult = namedtuple('RayResult', ['pkg', 'err'])
RayPkg = namedtuple('RayPkg', ['ray', 'op', 'wvl'])
RaySeg = namedtuple('RaySeg', ['p', 'd', 'dst', 'nrml'])

_TRACE_RAW_KEYS = ('eps', 'check_apertures', 'intersect_obj', 'filter_out_phantoms',
                   'first_surf', 'last_surf', 'pt_inside_fuzz')


def _table_for(opt_model, table=None, device=0):
    from .analyses import _table_for as tf
    return tf(opt_model, table, device)


def cuda_tracer(opt_model, table, fld, wvl, px, py, apply_vignetting, trace_kwargs):
    """Whole rays of the pupil points ``(px[k], py[k])`` of one field / wavelength:
    one paired-grid launch.  Returns host arrays ``full [n_ifc, 10, n]``, ``op``,
    ``status``, ``fail_surf``, ``n_seg``."""
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    recs, eprad, z_pupil = grid_fields_of(opt_model, [fld])
    grid = E.PupilGrid(recs, [table.wvl_index(wvl)], px, py, eprad, z_pupil,
                       apply_vignetting=apply_vignetting, flip_z_dir=sm.z_dir[0], paired=True,
                       device=table.device)
    kw = {k: v for k, v in trace_kwargs.items() if k in _TRACE_RAW_KEYS}
    kw.setdefault('check_apertures', False)          # trace_raw's default (raytrace.py:83)
    res = E.trace_grid(table, grid, outputs=('op', 'status', 'fail_surf', 'n_seg'), full=True,
                       summary=False, **kw)
    out = {'full': res.full.cpu().numpy(), 'op': res.op.cpu().numpy(),
           'status': res.status.cpu().numpy(), 'fail_surf': res.fail_surf.cpu().numpy(),
           'n_seg': res.n_seg.cpu().numpy()}
    grid.close()
    return out


def trace_pupil_rays(opt_model, pupils, fld, wvl, output_filter=None, rayerr_filter=None,
                     apply_vignetting=True, table=None, device=0, tracer=None, **kwargs):
    """``[trace_safe(opt_model, p, fld, wvl, output_filter, rayerr_filter, **kwargs) for p in
    pupils]`` (trace.py:159-219) with one launch.  Returns a list of ``RayResult``."""
    from . import raytrace as RT
    pupil_type = kwargs.pop('pupil_type', 'rel pupil')
    use_named_tuples = kwargs.get('use_named_tuples', False)
    if pupil_type != 'rel pupil':
        # 'aim pt' (points on the pupil plane) / 'aim dir' (object-space directions), trace.py:291-308:
        # the start rays are built on the host and go through one BUNDLE launch
        r = _trace_started_rays(opt_model, pupils, fld, wvl, pupil_type, table, device,
                                kwargs.pop('bundle_tracer', None), kwargs)
    else:
        pts = np.array([np.array(p, dtype=float) for p in pupils], dtype=float).reshape(-1, 2)
        if tracer is None:
            table = _table_for(opt_model, table, device)
            tracer = cuda_tracer
        r = tracer(opt_model, table, fld, wvl, pts[:, 0].copy(), pts[:, 1].copy(), apply_vignetting,
                   kwargs)
    segs = list(opt_model.seq_model.path(wvl))
    results = []
    for k in range(len(r['status'])):
        pkg, err = RT.package_ray(segs, r['full'][:, :, k], float(r['op'][k]), int(r['status'][k]),
                                  int(r['fail_surf'][k]), int(r['n_seg'][k]), wvl)
        if err is not None:
            if rayerr_filter == 'full':
                if err.ray_pkg is not None:
                    ray, op_delta, w = err.ray_pkg
                    err.ray_pkg = RayPkg([RaySeg(*rs) for rs in ray], op_delta, w)
                results.append(RayResult(err.ray_pkg, err))
            elif rayerr_filter == 'summary':
                err.ray_pkg = None
                results.append(RayResult(None, err))
            else:
                results.append(RayResult(None, None))
            continue
        if use_named_tuples:
            ray, op_delta, w = pkg
            pkg = RayPkg([RaySeg(*rs) for rs in ray], op_delta, w)
        if output_filter is None:
            results.append(RayResult(pkg, None))
        elif output_filter == 'last':
            ray, op_delta, w = pkg
            results.append(RayResult(RayPkg([ray[-1]], op_delta, w), None))
        else:
            results.append(RayResult(output_filter(pkg), None))
    return results


def _trace_started_rays(opt_model, pupils, fld, wvl, pupil_type, table, device, bundle_tracer, kwargs):
    """start rays from ``ray_start_from_osp(pupil, fld, pupil_type)`` (no vignetting: trace.py:291-295),
    the wide-angle / virtual-object rules of trace_base, one bundle launch.  ``bundle_tracer``: test
    seam with the signature of ``analyses._cuda_bundle_tracer``."""
    osp, sm = opt_model['optical_spec'], opt_model['seq_model']
    n = len(pupils)
    p0, d0 = np.zeros((3, n)), np.zeros((3, n))
    kw = {k: v for k, v in kwargs.items() if k in _TRACE_RAW_KEYS}
    kw.setdefault('first_surf', 1)
    kw.setdefault('last_surf', sm.get_num_surfaces() - 2)
    wide = bool(osp['fov'].is_wide_angle)
    if wide:
        kw['intersect_obj'] = False
    for k, pupil in enumerate(pupils):
        pt0, dir0 = osp.ray_start_from_osp(pupil, fld, pupil_type)
        if not wide and dir0[2]*sm.z_dir[0] < 0:
            dir0 = -dir0
        p0[:, k], d0[:, k] = pt0, dir0
    if bundle_tracer is None:
        from .analyses import _cuda_bundle_tracer as bundle_tracer
        table = _table_for(opt_model, table, device)
    row = table.wvl_index(wvl) if hasattr(table, 'wvl_index') else sm.index_for_wavelength(wvl)
    return bundle_tracer(opt_model, table, p0, d0, np.full(n, row, dtype=np.int32), kw)


def trace_safe(opt_model, pupil, fld, wvl, output_filter, rayerr_filter, **kwargs):
    """trace.py:159-219 for one pupil point."""
    return trace_pupil_rays(opt_model, [pupil], fld, wvl, output_filter, rayerr_filter, **kwargs)[0]


def trace_base(opt_model, pupil, fld, wvl, apply_vignetting=True, **kwargs):
    """trace.py:253-310 for one pupil point (``pupil_type``: 'rel pupil', 'aim pt', 'aim dir'): the ray
    package, or the TraceError raised."""
    res = trace_pupil_rays(opt_model, [pupil], fld, wvl, None, 'full',
                           apply_vignetting=apply_vignetting, **kwargs)[0]
    if res.err is not None:
        raise res.err
    return res.pkg


def _as_recorded(fld, pupils, kwargs):
    """The pupil coordinates the reference hands to ``img_filter`` and stores in its results:
    ``trace_base`` applies ``Field.apply_vignetting`` to the ndarray IN PLACE
    (``vig_pupil = pupil[:]`` is a view, opticalspec.py:1339-1353), so what the loops of
    trace.py:546-559,572-604 record are the VIGNETTED coordinates."""
    if not kwargs.get('apply_vignetting', True):
        return pupils
    return [fld.apply_vignetting(np.array(p)) for p in pupils]


def trace_fan(opt_model, fan_rng, fld, wvl, foc, img_filter=None, **kwargs):
    """trace.py:537-560: ``[[pupil, img_filter(pupil, ray_pkg)], ...]`` for the rays that
    yield a package; pupil coordinates are the accumulated ``start += step`` values."""
    output_filter = kwargs.pop('output_filter', None)
    rayerr_filter = kwargs.pop('rayerr_filter', None)
    start = np.array(fan_rng[0], dtype=float)
    stop = fan_rng[1]
    num = fan_rng[2]
    step = (stop - start)/(num - 1)
    pupils = []
    for _ in range(num):
        pupils.append(np.array(start))
        start += step
    results = trace_pupil_rays(opt_model, pupils, fld, wvl, output_filter, rayerr_filter, **kwargs)
    pupils = _as_recorded(fld, pupils, kwargs)
    fan = []
    for pupil, ray_result in zip(pupils, results):
        if ray_result.pkg is not None:
            if img_filter:
                fan.append([pupil, img_filter(pupil, ray_result.pkg)])
            else:
                fan.append([pupil, ray_result.pkg])
    return fan


def trace_grid(opt_model, grid_rng, fld, wvl, foc, img_filter=None, form='grid',
               append_if_none=True, **kwargs):
    """trace.py:563-605: x outer / y inner, ``check_apertures=True``, same nesting
    (``form='grid'`` rows or one flat ``'list'``) and ``append_if_none`` handling."""
    output_filter = kwargs.pop('output_filter', None)
    rayerr_filter = kwargs.pop('rayerr_filter', None)
    start = np.array(grid_rng[0], dtype=float)
    stop = grid_rng[1]
    num = grid_rng[2]
    step = np.array((stop - start)/(num - 1))
    pupils = []
    for i in range(num):
        for j in range(num):
            pupils.append(np.array(start))
            start[1] += step[1]
        start[0] += step[0]
        start[1] = grid_rng[0][1]
    kwargs['check_apertures'] = True
    results = trace_pupil_rays(opt_model, pupils, fld, wvl, output_filter, rayerr_filter, **kwargs)
    pupils = _as_recorded(fld, pupils, kwargs)
    grid = []
    k = 0
    for i in range(num):
        working_grid = grid if form == 'list' else []
        for j in range(num):
            pupil, ray_result = pupils[k], results[k]
            k += 1
            if ray_result.pkg is not None:
                if img_filter:
                    working_grid.append(img_filter(pupil, ray_result.pkg))
                else:
                    working_grid.append([pupil[0], pupil[1], ray_result.pkg])
            else:                                   # ray outside pupil or failed
                if img_filter:
                    result = img_filter(pupil, None)
                    if result is not None or append_if_none:
                        working_grid.append(result)
                elif append_if_none:
                    working_grid.append([pupil[0], pupil[1], None])
        if form == 'grid':
            grid.append(working_grid)
    try:
        return np.array(grid)
    except ValueError:                              # ragged / object entries (numpy >= 1.24)
        return np.array(grid, dtype=object)


def trace_chief_ray(opt_model, fld, wvl, foc, table=None, device=0, tracer=None):
    """trace.py:513-534: ``(chief_ray, cr_exp_seg)``; pupil (0, 0), apertures not checked."""
    res = trace_pupil_rays(opt_model, [np.array([0., 0.])], fld, wvl, None, 'full', table=table,
                           device=device, tracer=tracer)[0]
    if res.err is not None:
        raise res.err
    cr = RayPkg(*res.pkg)
    fod = opt_model['analysis_results']['parax_data'].fod
    cr_exp_seg = W.transfer_to_exit_pupil((cr.ray[-2][0], cr.ray[-2][1]), fod.exp_dist)
    return cr, cr_exp_seg


def setup_pupil_coords(opt_model, fld, wvl, foc, image_pt=None, image_delta=None, table=None,
                       device=0, tracer=None):
    """trace.py:607-624: ``(ref_sphere, chief_ray_pkg)`` of a field / wavelength."""
    chief_ray_pkg = trace_chief_ray(opt_model, fld, wvl, foc, table=table, device=device,
                                    tracer=tracer)
    ref_sphere = W.calculate_reference_sphere(opt_model, fld, wvl, foc, chief_ray_pkg,
                                              image_pt_2d=image_pt, image_delta=image_delta)
    return ref_sphere, chief_ray_pkg


def trace_ray(opt_model, pupil, fld, wvl, output_filter=None, rayerr_filter='full', **kwargs):
    """Trace a single ray via pupil, field and wavelength specs (trace.py:103-157):
    ``trace_safe`` with named tuples by default."""
    kwargs.setdefault('use_named_tuples', True)
    return trace_safe(opt_model, pupil, fld, wvl, output_filter, rayerr_filter, **kwargs)


def trace_boundary_rays_at_field(opt_model, fld, wvl, use_named_tuples=False, **kwargs):
    """list of RayPkgs of the boundary (pupil) rays of field ``fld`` (trace.py:441-458) -- the
    chief-ray setup plus ONE launch for all pupil rays"""
    rayerr_filter = kwargs.pop('rayerr_filter', 'full')
    output_filter = kwargs.pop('output_filter', None)
    engine = {k: kwargs[k] for k in ('table', 'device', 'tracer') if k in kwargs}
    ref_sphere, cr_pkg = setup_pupil_coords(opt_model, fld, wvl, 0.0, **engine)
    fld.chief_ray = cr_pkg
    fld.ref_sphere = ref_sphere
    results = trace_pupil_rays(opt_model, opt_model.optical_spec.pupil.pupil_rays, fld, wvl,
                               output_filter, rayerr_filter, use_named_tuples=use_named_tuples,
                               **kwargs)
    return [r.pkg for r in results]


def boundary_ray_dict(opt_model, rim_rays):
    """trace.py:461-465"""
    labels = getattr(opt_model.optical_spec.pupil, 'ray_labels', ['00', '+X', '-X', '+Y', '-Y'])
    return dict(zip(labels, rim_rays))


def trace_boundary_rays(opt_model, **kwargs):
    """boundary rays of every field at the central wavelength (trace.py:468-476)"""
    rayset = []
    wvl = opt_model.seq_model.central_wavelength()
    for fld in opt_model.optical_spec.field_of_view.fields:
        rim_rays = trace_boundary_rays_at_field(opt_model, fld, wvl, **kwargs)
        fld.pupil_rays = boundary_ray_dict(opt_model, rim_rays)
        rayset.append(rim_rays)
    return rayset


# --- batched stand-ins for the per-ray loops of rayoptics.raytr.analyses ------------------------
# (trace_ray_fan :212-230, trace_ray_list :437-455, trace_ray_grid :666-696): same arguments, same
# nested lists of [pupil_x, pupil_y, ray_pkg] out.  raytrace.install(batched=True) rebinds the
# reference's functions to these, so its RayFan / RayList / RayGrid classes trace each (field,
# wavelength) with one launch.
def _vignette_like_reference(fld, pupils, kwargs):
    """trace_base vignettes ndarray pupils IN PLACE (``vig_pupil = pupil[:]`` is a view; a list
    is copied): do the same to the caller's objects and return the coordinates the
    reference's loops then record"""
    if kwargs.get('apply_vignetting', True):
        for p in pupils:
            fld.apply_vignetting(p)
    return [(p[0], p[1]) for p in pupils]


def analyses_trace_ray_fan(opt_model, fan_rng, fld, wvl, foc, output_filter=None,
                           rayerr_filter=None, **kwargs):
    start = np.array(fan_rng[0], dtype=float)
    stop, num = fan_rng[1], fan_rng[2]
    step = (stop - start)/(num - 1)
    pupils = []
    for _ in range(num):
        pupils.append(np.array(start))
        start += step
    kwargs['use_named_tuples'] = True
    results = trace_pupil_rays(opt_model, [p.copy() for p in pupils], fld, wvl, output_filter,
                               rayerr_filter, **kwargs)
    rec = _vignette_like_reference(fld, pupils, kwargs)
    return [[px, py, r.pkg] for (px, py), r in zip(rec, results) if r.pkg is not None]


def analyses_trace_ray_list(opt_model, pupil_coords, fld, wvl, foc, append_if_none=False,
                            output_filter=None, rayerr_filter=None, **kwargs):
    pupils = list(pupil_coords)
    results = trace_pupil_rays(opt_model, [np.array(p, dtype=float) for p in pupils], fld, wvl,
                               output_filter, rayerr_filter, **kwargs)
    rec = _vignette_like_reference(fld, pupils, kwargs)
    ray_list = []
    for (px, py), r in zip(rec, results):
        if r.pkg is not None:
            ray_list.append([px, py, r.pkg])
        elif append_if_none:
            ray_list.append([px, py, None])
    return ray_list


def analyses_trace_ray_grid(opt_model, grid_rng, fld, wvl, foc, append_if_none=True,
                            output_filter=None, rayerr_filter=None, **kwargs):
    start = np.array(grid_rng[0], dtype=float)
    stop, num = grid_rng[1], grid_rng[2]
    step = np.array((stop - start)/(num - 1))
    kwargs['apply_vignetting'] = kwargs.get('apply_vignetting', False)
    pupils = []
    for i in range(num):
        for j in range(num):
            pupils.append(np.array(start))
            start[1] += step[1]
        start[0] += step[0]
        start[1] = grid_rng[0][1]
    results = trace_pupil_rays(opt_model, [p.copy() for p in pupils], fld, wvl, output_filter,
                               rayerr_filter, **kwargs)
    rec = _vignette_like_reference(fld, pupils, kwargs)
    grid, k = [], 0
    for i in range(num):
        row = []
        for j in range(num):
            (px, py), r = rec[k], results[k]
            k += 1
            if r.pkg is not None:
                row.append([px, py, r.pkg])
            elif append_if_none:
                row.append([px, py, None])
        grid.append(row)
    return grid


# --- the rest of rayoptics.raytr.trace's call surface around the path ---------------------------
def _engine(kwargs):
    return {k: kwargs.pop(k) for k in ('table', 'device', 'tracer') if k in kwargs}


def ray_pkg(ray_pkg):
    """a pandas Series holding a ray package (trace.py:31-33)"""
    import pandas as pd
    return pd.Series(ray_pkg, index=['ray', 'op', 'wvl'])


def ray_df(ray):
    """a pandas DataFrame holding the segments of a ray (trace.py:36-41)"""
    import pandas as pd
    r = pd.DataFrame(ray, columns=['inc_pt', 'after_dir', 'after_dst', 'normal'])
    r.index.names = ['intrfc']
    return r


def list_ray(ray_obj, tfrms=None, start=0):
    """pretty print a ray in local or (with ``tfrms``) global coordinates (trace.py:44-79).
    ``ray_obj``: the return of ``trace_ray`` (package, error), a ray package or a ray."""
    ray_err = None
    if isinstance(ray_obj, tuple):
        if len(ray_obj) == 2:
            ray_obj, ray_err = ray_obj
        ray = ray_obj[0]
    else:
        ray = ray_obj
    print('            X            Y            Z           L            M            N'
          '               Len')
    fmt = '{:3d}: {:12.5f} {:12.5f} {:12.5g} {:12.6f} {:12.6f} {:12.6f} {:12.5g}'
    for i, seg in enumerate(ray[start:], start=start):
        p, d = seg[0], seg[1]
        if tfrms is not None:
            rot, trns = tfrms[i]
            p, d = rot.dot(p) + trns, rot.dot(d)
        print(fmt.format(i, p[0], p[1], p[2], d[0], d[1], d[2], seg[2]))
    if ray_err is not None:
        print(f'ray failure: {type(ray_err).__name__}')


def list_in_out_dir(path, ray):
    """list the incident and exiting direction cosines at every interface (trace.py:82-100)"""
    lcl_tfrms = [seg[2] for seg in path]
    row = '{:10.6f} {:10.6f} {:10.6f}'
    print('                  in_dir              |              out_dir')
    before_dir = ray[0][1]
    print(f'{0:2d}:                                   |' + row.format(*before_dir))
    i = 0
    for i, (seg, tfrm) in enumerate(zip(ray[1:], lcl_tfrms), start=1):
        b4_dir = tfrm[0].dot(before_dir)
        print(f'{i:2d}: ' + row.format(*b4_dir) + '  |' + row.format(*seg[1]))
        before_dir = seg[1]
    b4_dir = lcl_tfrms[-1][0].dot(before_dir)
    print(f'{i+1:2d}: ' + row.format(*b4_dir) + '  |')


def aim_chief_ray(opt_model, fld, wvl=None, trace_fn=None):
    """aim the chief ray at the centre of the stop surface (trace.py:627-640): the aim point on
    the paraxial entrance pupil, or -- wide-angle specifications -- the z position of the real
    entrance pupil (raytr/wideangle.py)."""
    from . import raytrace as RT, vigcalc, wideangle
    sm = opt_model.seq_model
    wvl = sm.central_wavelength() if wvl is None else wvl
    if opt_model.optical_spec.field_of_view.is_wide_angle:
        z_enp, _ = wideangle.find_real_enp(opt_model, sm.stop_surface, fld, wvl, trace_fn)
        return float(z_enp)
    return vigcalc.aim_chief_ray(opt_model, fld, wvl, RT.trace if trace_fn is None else trace_fn)


def get_chief_ray_pkg(opt_model, fld, wvl, foc, trace_fn=None, **engine):
    """the chief ray package of ``fld``, computed (and aimed) when necessary (trace.py:660-687)"""
    if fld.chief_ray is None:
        fld.aim_info = aim_chief_ray(opt_model, fld, wvl=wvl, trace_fn=trace_fn)
        return trace_chief_ray(opt_model, fld, wvl, foc, **engine)
    if fld.chief_ray[0][2] != wvl:
        return trace_chief_ray(opt_model, fld, wvl, foc, **engine)
    return fld.chief_ray


def trace_with_opd(opt_model, pupil, fld, wvl, foc, **kwargs):
    """``(ray, ray_opl, wvl, opd)`` of one pupil point (trace.py:418-438)"""
    engine = {k: kwargs[k] for k in ('table', 'device', 'tracer') if k in kwargs}
    chief = get_chief_ray_pkg(opt_model, fld, wvl, foc, trace_fn=kwargs.pop('trace_fn', None),
                              **engine)
    ref_sphere = W.calculate_reference_sphere(opt_model, fld, wvl, foc, chief,
                                              image_pt_2d=kwargs.pop('image_pt', None),
                                              image_delta=kwargs.pop('image_delta', None))
    pkg, err = trace_ray(opt_model, pupil, fld, wvl, **kwargs)
    fld.chief_ray, fld.ref_sphere = chief, ref_sphere
    fod = opt_model['analysis_results']['parax_data'].fod
    opd = W.wave_abr_full_calc(fod, fld, wvl, foc, pkg, chief, ref_sphere)
    ray, ray_op, wvl = pkg
    return ray, ray_op, wvl, opd


def trace_ray_list_at_field(opt_model, ray_list, fld, wvl, foc, **kwargs):
    """a ray DataFrame per pupil point of ``ray_list`` at ``fld`` (trace.py:478-486); one launch"""
    kwargs.setdefault('use_named_tuples', True)
    results = trace_pupil_rays(opt_model, list(ray_list), fld, wvl, kwargs.pop('output_filter', None),
                               kwargs.pop('rayerr_filter', 'full'), **kwargs)
    return [ray_df(r.pkg[0]) for r in results]


def trace_field(opt_model, fld, wvl, foc, **engine):
    """DataFrame of the boundary rays of ``fld`` (trace.py:489-496)"""
    import pandas as pd
    pupil = opt_model.optical_spec.pupil
    rdf_list = trace_ray_list_at_field(opt_model, pupil.pupil_rays, fld, wvl, foc, **engine)
    return pd.concat(rdf_list, keys=pupil.ray_labels, names=['pupil'])


def trace_all_fields(opt_model, **engine):
    """DataFrame of the boundary rays of all fields (trace.py:499-510)"""
    import pandas as pd
    osp = opt_model.optical_spec
    _, wvl, foc = osp.lookup_fld_wvl_focus(0)
    fov = osp.field_of_view
    fset = [trace_field(opt_model, f, wvl, foc, **engine) for f in fov.fields]
    return pd.concat(fset, keys=fov.index_labels, names=['field'])


def refocus(opt_model, **engine):
    """focus shift that brings the axial marginal ray to the axis (trace.py:690-705)"""
    osp = opt_model['optical_spec']
    fld = osp['fov'].fields[0]
    wvl = osp['wvls'].central_wvl
    res = trace_safe(opt_model, [0., 1.], fld, wvl, output_filter=None, rayerr_filter='full',
                     use_named_tuples=True, **engine)
    ray = res.pkg[0]
    return -ray[-1].p[1]/(ray[-2].d[1]/ray[-2].d[2])


def intersect_2_lines(P1, V1, P2, V2):
    """distance from P1 along V1 to the intersection of two non-parallel lines (trace.py:779-789)"""
    Vx = np.cross(V1, V2)
    return np.dot(np.cross(P2 - P1, V1), Vx)/np.dot(Vx, Vx)


def trace_astigmatism(opt_model, fld, wvl, foc, dx=0.001, dy=0.001, **engine):
    """sagittal and tangential focus shifts at ``fld`` from close rays about the chief ray
    (trace.py:826-863); the five rays go through one launch"""
    pupils = [[0., 0.], [dx, 0.], [0., dy], [-dx, 0.], [0., -dy]]
    r = [x.pkg for x in trace_pupil_rays(opt_model, pupils, fld, wvl, None, 'full',
                                         use_named_tuples=True, **engine)]
    s = intersect_2_lines(r[1].ray[-1].p, r[1].ray[-1].d, r[3].ray[-1].p, r[3].ray[-1].d)
    s_foc = s*r[1].ray[-1].d[2]
    t = intersect_2_lines(r[2].ray[-1].p, r[2].ray[-1].d, r[4].ray[-1].p, r[4].ray[-1].d)
    t_foc = t*r[2].ray[-1].d[2]
    if foc is not None:
        s_foc -= foc
        t_foc -= foc
    return s_foc, t_foc


def trace_astigmatism_curve(opt_model, num_points=21, **kwargs):
    """astigmatism over a fan of fields from the axis to the maximum field (trace.py:792-823):
    ``(field values, sagittal focus shifts, tangential focus shifts)``"""
    from .model import Field
    engine = _engine(kwargs)
    osp = opt_model['optical_spec']
    fov = osp['fov']
    _, wvl, foc = osp.lookup_fld_wvl_focus(0)
    fld = Field(fov=fov)
    field_data, s_data, t_data = [], [], []
    for f in np.linspace(0., fov.max_field()[0], num=num_points):
        fld.yv = f
        ref_sphere, cr_pkg = setup_pupil_coords(opt_model, fld, wvl, foc, **engine)
        fld.chief_ray, fld.ref_sphere = cr_pkg, ref_sphere
        s_foc, t_foc = trace_astigmatism(opt_model, fld, wvl, foc, **kwargs, **engine)
        s_data.append(s_foc)
        t_data.append(t_foc)
        field_data.append(f)
    return field_data, s_data, t_data


def trace_coddington_fan(opt_model, ray_pkg, foc=None):
    """sagittal / tangential focus along a traced ray by Coddington's equations
    (trace.py:715-776; spherical surfaces only).  Host arithmetic on one ray package."""
    import math
    sm = opt_model.seq_model
    ray = ray_pkg[0]
    wl = sm.index_for_wavelength(ray_pkg[2])
    n_ifc = len(ray)
    rind = [sm.rndx[i][wl] if i < len(sm.rndx) else None for i in range(n_ifc)]
    before_rind = sm.rndx[0][wl]
    s_before = t_before = s_prime = t_prime = None
    for i, (pt, after_dir, after_dst, normal) in enumerate(ray):
        after_rind = rind[i] if rind[i] is not None else before_rind
        if i == 0:
            s_before = t_before = -after_dst
        else:
            cosI_prime = np.dot(after_dir, normal)/np.linalg.norm(normal)
            sinI_prime = math.sqrt(1.0 - cosI_prime**2)
            sinI = after_rind*sinI_prime/before_rind
            cosI = math.sqrt(1.0 - sinI**2)
            obl_power = sm.ifcs[i].optical_power
            if obl_power != 0.0:
                obl_power *= ((after_rind*cosI_prime - before_rind*cosI)/(after_rind - before_rind))
            s_prime = after_rind/(before_rind/s_before + obl_power)
            s_before = s_prime - after_dst
            t_prime = after_rind*cosI_prime**2/(before_rind*cosI**2/t_before + obl_power)
            t_before = t_prime - after_dst
        before_rind = after_rind
    s_dfoc = s_prime*after_dir[2] + pt[2]
    t_dfoc = t_prime*after_dir[2] + pt[2]
    if foc is not None:
        s_dfoc -= foc
        t_dfoc -= foc
    return s_dfoc, t_dfoc


def trace_astigmatism_coddington_fan(opt_model, fld, wvl, foc, **engine):
    """astigmatism by a Coddington trace along the chief ray of ``fld`` (trace.py:708-712)"""
    cr_pkg, _ = trace_ray(opt_model, [0., 0.], fld, wvl, **engine)
    return trace_coddington_fan(opt_model, cr_pkg, foc=foc)


def iterate_ray(opt_model, ifcx, xy_target, fld, wvl, **kwargs):
    """``(start_coords, (ray_pkg, error))`` as the reference returns it (trace.py:313-415);
    the solver is ``vigcalc.iterate_ray``"""
    from . import vigcalc
    return vigcalc.iterate_ray(opt_model, ifcx, xy_target, fld, wvl,
                               trace_fn=kwargs.get('trace_fn'), full=True)


def iterate_ray_raw(pthlist, ifcx, xy_target, pt0, d0, obj2pup_dist, eprad, wvl, not_wa, **kwargs):
    """trace.py:866-957, see ``vigcalc.iterate_ray_raw``"""
    from . import vigcalc
    return vigcalc.iterate_ray_raw(pthlist, ifcx, xy_target, pt0, d0, obj2pup_dist, eprad, wvl,
                                   not_wa, **kwargs)


def apply_paraxial_vignetting(opt_model):
    """trace.py:643-657, see ``vigcalc.apply_paraxial_vignetting``"""
    from . import vigcalc
    return vigcalc.apply_paraxial_vignetting(opt_model)
