"""Drop-in ``trace()`` / ``trace_raw()`` on the B200 engine.

Same signatures, return values and exceptions as
/root/reference/src/rayoptics/raytr/raytrace.py:51-264:

    trace(seq_model, pt0, dir0, wvl, **kwargs)        -> (ray, op_delta, wvl)
    trace_raw(path, pt0, dir0, wvl, eps=1e-12, check_apertures=False,
              intersect_obj=True, filter_out_phantoms=False, **kwargs)

``ray`` is a list with one ``[p, d, dst, nrml]`` entry per interface (numpy
3-vectors + float), indexed by the reference's ``mc.p/d/dst/nrml``.  Per-ray
failures come back from the kernel as data (status, failing surface) and are
re-raised here as the reference's ``TraceError`` subclasses with ``.surf``,
``.ifc``, ``.ray_pkg`` (partial ray), ``.int_pt`` / ``.prev_tfrm`` filled the
way raytrace.py:231-257 fills them.

``install()`` rebinds ``rayoptics.raytr.raytrace.trace`` / ``trace_raw`` so that
every caller in the reference (``trace_base``, ``iterate_ray``, ``vigcalc``,
``wideangle`` ... they all call ``rt.trace`` through the module attribute) runs
on the GPU without touching reference sources.

A single ray costs a kernel launch and a device->host copy (~0.1 ms): this entry
point exists for compatibility; throughput comes from ``engine.trace_bundle`` /
``engine.trace_grid`` / ``analyses``.  There is no CPU fallback: models with
interfaces the table cannot represent raise ``UnsupportedInterfaceError``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi, engine as E
from .table import SurfaceTable

try:                                    # the reference's own exception classes, when importable
    from rayoptics.raytr.traceerror import (TraceError, TraceMissedSurfaceError,   # type: ignore
                                            TraceTIRError, TraceRayBlockedError,
                                            TraceEvanescentRayError)
except Exception:                       # noqa: BLE001 - any import problem -> local mirrors
    from .traceerror import (TraceError, TraceMissedSurfaceError, TraceTIRError,
                             TraceRayBlockedError, TraceEvanescentRayError)


class TraceNumericError(TraceError):
    """The reference would have died with an uncaught ValueError/ZeroDivisionError
    inside a polynomial profile (status RT_RAY_NUMERIC)."""

    def __init__(self, ifc=None, prev_seg=None):
        self.ifc = ifc
        self.prev_seg = prev_seg


_DEVICE = 0
_PATH_CACHE = {}          # fingerprint -> SurfaceTable


def set_device(device):
    global _DEVICE
    _DEVICE = int(device)


def _phase_key(pe):
    if pe is None:
        return None
    fp = [type(pe).__name__, getattr(getattr(pe, 'phase_fct', None), '__name__', None)]
    for k in ('ref_pt', 'obj_pt', 'grating_normal', 'coefficients'):
        v = getattr(pe, k, None)
        fp.append(None if v is None else tuple(map(float, v)))
    for k in ('ref_virtual', 'obj_virtual', 'ref_wl', 'order', '_grating_spacing_nm'):
        fp.append(getattr(pe, k, None))
    return tuple(fp)


def _aperture_key(ca):
    return (type(ca).__name__, getattr(ca, 'radius', None), getattr(ca, 'x_half_width', None),
            getattr(ca, 'y_half_width', None), getattr(ca, 'x_offset', 0.0), getattr(ca, 'y_offset', 0.0),
            bool(getattr(ca, 'is_obscuration', False)))


def _fingerprint(segs):
    """Cache key of a path: every attribute ``table._describe_interface`` reads -- profile numbers and
    coefficient count, interact mode, max_aperture and each clear aperture's type / size / offsets /
    obscuration flag, transform (values AND numpy memory order, which selects the dgemv rounding),
    index, z_dir, phase-element parameters -- as a tuple of Python values (~25 us for 13 interfaces;
    compiling the descriptors costs 6x that).  An in-place edit of any of them yields a new table
    (the reference's own invalidation point is update_model(), seq/sequential.py:666-668)."""
    fp = []
    for seg in segs:
        ifc, _gap, tfrm, n, z_dir = (tuple(seg) + (None,)*5)[:5]
        prf = getattr(ifc, 'profile', None)
        coefs = getattr(prf, 'coefs', None)
        cas = getattr(ifc, 'clear_apertures', None)
        if tfrm is None:
            tk = None
        else:
            r = tfrm[0]
            tk = ((r.tobytes(), r.flags['C_CONTIGUOUS'], r.flags['F_CONTIGUOUS']) if hasattr(r, 'tobytes')
                  else repr(r), tuple(map(float, tfrm[1])))
        fp.append((type(ifc).__name__, type(prf).__name__, getattr(prf, 'cv', None),
                   getattr(prf, 'cc', None), getattr(prf, 'ec', None), getattr(prf, 'cR', None),
                   None if coefs is None else tuple(coefs), getattr(prf, 'max_nonzero_coef', None),
                   getattr(ifc, 'interact_mode', None), getattr(ifc, 'max_aperture', None),
                   None if not cas else tuple(_aperture_key(ca) for ca in cas), n, z_dir,
                   _phase_key(getattr(ifc, 'phase_element', None)) if hasattr(ifc, 'phase_element') else 0,
                   tk))
    return tuple(fp)


def _table_for_path(segs, wvl=None):
    """Surface table of a path, cached on the content it is compiled from."""
    fp = (_fingerprint(segs), wvl, _DEVICE)
    tab = _PATH_CACHE.get(fp)
    if tab is None:
        if len(_PATH_CACHE) > 64:
            _PATH_CACHE.clear()
        tab = SurfaceTable.from_path(segs, device=_DEVICE, wvl=wvl)
        _PATH_CACHE[fp] = tab
    return tab


def _ray_list(full, n_seg):
    ray = []
    for k in range(n_seg):
        s = full[k]
        ray.append([s[0:3].copy(), s[3:6].copy(), float(s[6]), s[7:10].copy()])
    return ray


def trace_raw(path, pt0, dir0, wvl, eps=1.0e-12, check_apertures=False,
              intersect_obj=True, filter_out_phantoms=False, **kwargs):
    """fundamental raytrace function (raytrace.py:83-264) -- one ray on the GPU."""
    segs = list(path)
    tab = _table_for_path(segs, wvl)
    first_surf = kwargs.get('first_surf', 0)
    last_surf = kwargs.get('last_surf', None)
    pt_inside_fuzz = kwargs.get('pt_inside_fuzz', None)
    p = np.asarray(pt0, dtype=np.float64).reshape(3, 1)
    d = np.asarray(dir0, dtype=np.float64).reshape(3, 1)
    res = E.trace_bundle(tab, p, d, full=True, outputs=('op', 'status', 'fail_surf', 'n_seg'),
                         eps=eps, check_apertures=check_apertures, intersect_obj=intersect_obj,
                         filter_out_phantoms=filter_out_phantoms, first_surf=first_surf,
                         last_surf=last_surf, pt_inside_fuzz=pt_inside_fuzz)
    full = res.full[:, :, 0].cpu().numpy()
    meta = torch.stack([res.status, res.fail_surf, res.n_seg]).cpu().numpy()[:, 0]
    op = float(res.op.cpu().numpy()[0])
    ray_pkg, err = package_ray(segs, full, op, int(meta[0]), int(meta[1]), int(meta[2]), wvl)
    if err is not None:
        raise err
    return ray_pkg


def package_ray(segs, full, op, status, surf, n_seg, wvl):
    """One ray of a bundle / grid result -> ``((ray, op, wvl), None)`` or
    ``(None, TraceError)`` with the error filled as raytrace.py:231-257 fills it.
    ``segs``: the path list, ``full``: ``[n_ifc, 10]`` segments of this ray."""
    ray = _ray_list(full, n_seg)
    if status == _abi.RAY_OK:
        return (ray, op, wvl), None
    ifc = segs[surf][0] if 0 <= surf < len(segs) else None
    if status == _abi.RAY_MISSED:
        err = TraceMissedSurfaceError(ifc, None)
        err.prev_tfrm = segs[surf - 1][2] if surf >= 1 else None
    elif status == _abi.RAY_TIR:
        n_in = segs[surf - 1][3]
        n_out = segs[surf][3]
        err = TraceTIRError(ray[-2][1] if len(ray) > 1 else None, ray[-1][3], n_in, n_out)
        err.ifc = ifc
        err.int_pt = ray[-1][0]
    elif status == _abi.RAY_BLOCKED:
        err = TraceRayBlockedError(ifc, ray[-1][0])
    elif status == _abi.RAY_EVANESCENT:
        err = TraceEvanescentRayError(ifc, ray[-1][0], None, None, None, None)
    else:
        err = TraceNumericError(ifc, None)
    err.surf = surf
    if surf == 0 and status in (_abi.RAY_MISSED, _abi.RAY_NUMERIC):
        # the reference intersects the object outside its try block
        # (raytrace.py:147-151): the error propagates without a ray package
        err.ray_pkg = None
    else:
        err.ray_pkg = ray, op, wvl
    return None, err


def trace(seq_model, pt0, dir0, wvl, **kwargs):
    """fundamental raytrace function (raytrace.py:51-80)."""
    path = seq_model.path(wvl)
    kwargs['first_surf'] = kwargs.get('first_surf', 1)
    kwargs['last_surf'] = kwargs.get('last_surf', seq_model.get_num_surfaces() - 2)
    return trace_raw(path, pt0, dir0, wvl, **kwargs)


# --- the module's small host-side helpers, for callers that use them directly
#     (oprops/doe.py:298,321 calls rt.bend; analyses call calc_optical_path) -------------------
def bend(d_in, normal, n_in, n_out):
    """refract incoming direction, d_in, about normal (raytrace.py:19-30)"""
    from math import sqrt, copysign
    try:
        normal_len = np.linalg.norm(normal)
        cosI = np.dot(d_in, normal)/normal_len
        sinI_sqr = 1.0 - cosI*cosI
        n_cosIp = copysign(sqrt(n_out*n_out - n_in*n_in*sinI_sqr), cosI)
        alpha = n_cosIp - n_in*cosI
        d_out = (n_in*d_in + alpha*normal)/n_out
        return d_out
    except ValueError:
        raise TraceTIRError(d_in, normal, n_in, n_out)


def reflect(d_in, normal):
    """reflect incoming direction, d_in, about normal (raytrace.py:33-38)"""
    normal_len = np.linalg.norm(normal)
    cosI = np.dot(d_in, normal)/normal_len
    d_out = d_in - 2.0*cosI*normal
    return d_out


def phase(ifc, inc_pt, d_in, normal, ifc_cntxt):
    """diffracted direction and phase at an interface with a phase element (raytrace.py:41-48): the
    interface's own ``phase`` method does the work (the reference's interfaces have one; in the
    kernels the same arithmetic is ``hoe_phase`` / ``grating_phase`` / ``radial_doe_phase``), an
    evanescent order -- a ``ValueError`` there -- becomes the trace error"""
    try:
        return ifc.phase(inc_pt, d_in, normal, ifc_cntxt)
    except ValueError:
        z_dir, wvl, n_in, n_out, interact_mode = ifc_cntxt
        raise TraceEvanescentRayError(ifc, inc_pt, d_in, normal, n_in, n_out)


def calc_optical_path(ray, path):
    """optical path between the first and last optical surfaces (raytrace.py:267-293)"""
    num_items = len(ray)
    ray_seq_iter = zip(ray, path)
    next(ray_seq_iter)
    ray_op = 0
    for i in range(1, num_items - 2):
        after_ray_seg, surf = next(ray_seq_iter)
        ray_op += surf[3]*after_ray_seg[2]
    return ray_op


_saved = {}


def install(batched=False):
    """Rebind the reference's ``rayoptics.raytr.raytrace.trace/trace_raw`` to the engine.

    ``batched=True`` additionally rebinds the per-ray LOOPS of the reference to the batched
    drivers of ``rayoptics_b200.trace`` -- ``rayoptics.raytr.trace.trace_fan / trace_grid`` and
    ``rayoptics.raytr.analyses.trace_ray_fan / trace_ray_list / trace_ray_grid``, plus the short
    lists ``trace.trace_boundary_rays_at_field / trace_astigmatism / trace_ray_list_at_field`` -- so that the
    reference's own ``RayFan`` / ``RayList`` / ``RayGrid`` / ``SequentialModel.trace_fan`` ...
    trace each (field, wavelength) in one launch.  Same arguments and results (checked
    against the unpatched reference in tests/test_dropin_batched.py, wide-angle fields
    included); ``pupil_type`` other than 'rel pupil' keeps the reference's loop on the drop-in
    ``trace``."""
    import rayoptics.raytr.raytrace as rt      # type: ignore
    if 'trace' not in _saved:
        _saved['trace'], _saved['trace_raw'] = rt.trace, rt.trace_raw
    rt.trace, rt.trace_raw = trace, trace_raw
    if batched:
        import rayoptics.raytr.trace as rtr        # type: ignore
        import rayoptics.raytr.analyses as ran     # type: ignore
        from . import trace as TR
        pairs = [(rtr, 'trace_fan', TR.trace_fan), (rtr, 'trace_grid', TR.trace_grid),
                 (ran, 'trace_ray_fan', TR.analyses_trace_ray_fan),
                 (ran, 'trace_ray_list', TR.analyses_trace_ray_list),
                 (ran, 'trace_ray_grid', TR.analyses_trace_ray_grid),
                 # short pupil-point lists of raytr/trace.py: one launch instead of 5 / n
                 (rtr, 'trace_boundary_rays_at_field', TR.trace_boundary_rays_at_field),
                 (rtr, 'trace_astigmatism', TR.trace_astigmatism),
                 (rtr, 'trace_ray_list_at_field', TR.trace_ray_list_at_field)]
        for mod, name, fn in pairs:
            key = (mod.__name__, name)
            if key not in _saved:
                _saved[key] = getattr(mod, name)
            setattr(mod, name, _batched_or_original(fn, _saved[key]))
    return rt


def _batched_or_original(batched_fn, original_fn):
    """the batched driver where it applies, the reference's own loop otherwise"""
    import functools

    @functools.wraps(original_fn)
    def wrapper(opt_model, *args, **kwargs):
        if kwargs.get('pupil_type', 'rel pupil') != 'rel pupil':
            return original_fn(opt_model, *args, **kwargs)
        return batched_fn(opt_model, *args, **kwargs)
    return wrapper


def uninstall():
    if _saved:
        import importlib
        import rayoptics.raytr.raytrace as rt  # type: ignore
        rt.trace, rt.trace_raw = _saved.pop('trace'), _saved.pop('trace_raw')
        for key in [k for k in _saved if isinstance(k, tuple)]:
            setattr(importlib.import_module(key[0]), key[1], _saved.pop(key))
