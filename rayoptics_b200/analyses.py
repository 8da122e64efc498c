"""Grid / list / fan analyses as single bundle launches.

The reference evaluates every analysis with a Python loop that traces one ray
per iteration (/root/reference/src/rayoptics/raytr/trace.py:537-605,
raytr/analyses.py:212-230,437-455,666-696; seq/sequential.py:1006-1085).  Here
the whole fields x wavelengths x pupil-grid index space is one ``rt_trace_grid``
launch; the start rays are generated on the device and only the requested
results come back to the host.

``spot_diagram`` is the engine-side equivalent of
``SequentialModel.trace_grid(spot, fi, num_rays=N, form='list',
append_if_none=False)`` evaluated for every field
(seq/sequential.py:1058-1085 with the ``spot`` filter of
mpl/axisarrayfigure.py:229-238): per field and wavelength the transverse ray
aberrations ``(dx, dy)`` of the rays that reach the image, referred to the
chief-ray image point at the central wavelength.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from . import engine as E
from .table import SurfaceTable
from .opticalspec import grid_fields_of


class SpotDiagram:
    """Result of ``spot_diagram``.

    ``abr`` ``[2, n]``: host (pinned) array over the rays this process traced, flattened
    (field, wvl, i, j) order starting at ``first_ray``; rays that do not reach the image
    hold NaNs whose payloads are their status / failing surface (``status``, ``fail_surf``
    decode them on first use).  ``grids[fi][wi]``: ``[n_ok, 2]`` arrays of the rays that
    reach the image, in the reference's order (x outer, y inner) -- the ``form='list',
    append_if_none=False`` shape of seq/sequential.py:1058-1085 -- built lazily (whole grid
    only); ``summary``: dict of ``[n_fields, n_wvls]`` arrays (n_ok, n_blocked,
    centroid_x/y, rms_radius ...), combined over all ranks; ``ref_img``: ``[n_fields, 2]``
    chief-ray image points."""

    def __init__(self, abr, status, summary, ref_img, num_rays, n_fields, n_wvls, first_ray,
                 n_rays_total, io_bytes):
        self.abr, self._status, self.summary, self.ref_img = abr, status, summary, ref_img
        self._fail_surf = None
        self.num_rays, self.n_fields, self.n_wvls = num_rays, n_fields, n_wvls
        self.first_ray, self.n_rays_total = first_ray, n_rays_total
        self.io_bytes = io_bytes            # {'h2d': ..., 'd2h': ...} of this call
        self._grids = None

    def _decode(self):
        if self._status is None:
            self._status, self._fail_surf = E.decode_nan_status(self.abr)

    @property
    def status(self):
        self._decode()
        return self._status

    @property
    def fail_surf(self):
        self._decode()
        return self._fail_surf

    @property
    def grids(self):
        if self._grids is None:
            if self.first_ray != 0 or self.abr.shape[1] != self.n_rays_total:
                raise ValueError('per-tile lists need the whole grid on one process')
            ok = (self._status == 0) if self._status is not None else ~np.isnan(self.abr[0])
            per = self.num_rays*self.num_rays
            self._grids = []
            for fi in range(self.n_fields):
                row = []
                for wi in range(self.n_wvls):
                    sl = slice((fi*self.n_wvls + wi)*per, (fi*self.n_wvls + wi + 1)*per)
                    m = ok[sl]
                    row.append(np.stack([self.abr[0, sl][m], self.abr[1, sl][m]], axis=1))
                self._grids.append(row)
        return self._grids


_STREAMS = {}
# Re-used device blocks are per THREAD: two threads running analyses on one device must not share a
# grid block (one would trace the other's description) or staging buffers.  Tables are immutable and
# stay shared (SURVEY.md 8(b) "Threading").
_TLS = threading.local()


def _per_thread(name):
    d = getattr(_TLS, name, None)
    if d is None:
        d = {}
        setattr(_TLS, name, d)
    return d


def _side_streams(dev):
    if dev not in _STREAMS:
        _STREAMS[dev] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    return _STREAMS[dev]


def _spec_key(opt_model, table, num_rays, fields, wvls, foc):
    """Everything the grid description is computed from, as a hashable key: while it does not
    change, the host-side description (start points, aim points, pupil tables) of the previous
    call is uploaded again instead of being recomputed."""
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    fod = osp.fod if hasattr(osp, 'fod') else opt_model['analysis_results']['parax_data'].fod
    fkey = tuple((f.x, f.y, f.vux, f.vuy, f.vlx, f.vly,
                  None if f.aim_info is None else tuple(np.asarray(f.aim_info, dtype=float).ravel()))
                 for f in fields)
    fov, pup = osp.field_of_view, osp.pupil
    return (id(table), num_rays, fkey, tuple(wvls), foc, tuple(pup.key), pup.value, tuple(fov.key),
            fov.value, fov.is_relative, fov.is_wide_angle, sm.z_dir[0], sm.gaps[0].thi) + tuple(
                getattr(fod, a, None) for a in ('obj_dist', 'enp_dist', 'enp_radius', 'm', 'obj_na',
                                                'pr_ht0', 'pr_slp0', 'n_obj', 'n_img'))


def _reusable_grid(opt_model, table, num_rays, fields, wvls, foc):
    """The PupilGrid of this description: the device block and its pinned staging are allocated
    once per shape (``rt_grid_create``); every call uploads the description again (one
    asynchronous copy, ``rt_grid_update``) -- no cudaMalloc / cudaFree per analysis call -- and
    the host-side description itself is recomputed only when its inputs changed."""
    key = _spec_key(opt_model, table, num_rays, fields, wvls, foc)
    hit = _SPECS.get(key)
    if hit is None:
        args, kw = E._grid_args(opt_model, table.wvl_index, num_rays, fields, wvls, foc, (-1.0, 1.0), True)
        spec = E.PupilGridSpec(*args, **kw)
        if len(_SPECS) > 32:
            _SPECS.clear()
        hit = _SPECS[key] = (spec, args, kw)
    spec, args, kw = hit
    gkey = (int(table.device), spec.n_fields, spec.n_wvls, spec.nx, spec.ny, spec.paired, spec.wave is not None)
    grids = _per_thread('grids')     # shape key -> PupilGrid whose device block / pinned staging is re-used
    grid = grids.get(gkey)
    if grid is None or grid._handle is None:
        if len(grids) > 16:
            for g in grids.values():
                g.close()
            grids.clear()
        grid = grids[gkey] = E.PupilGrid(*args, device=table.device, **kw)
    else:
        grid.upload(spec)
    return grid, spec


_SPECS = {}


def _table_for(opt_model, table=None, device=0):
    if table is not None:
        return table
    sm = opt_model.seq_model
    cached = getattr(sm, '_b200_table', None)
    version = getattr(sm, '_version', None)
    built = None
    if version is None:
        # a model of the reference: no edit counter -- key the cache on the compiled
        # descriptor records themselves (every number the kernels can see; the reference's
        # own invalidation point is update_model(), seq/sequential.py:666-668)
        from .table import describe_model
        built = describe_model(sm)
        version = (bytes(built[0]), built[1].tobytes(), tuple(built[2]))
    if cached is None or cached[0] != version or cached[1].device != device:
        tab = SurfaceTable(*built, device=device) if built else SurfaceTable.from_model(sm, device=device)
        cached = (version, tab)
        sm._b200_table = cached
    return cached[1]


def chief_ray_image_points(opt_model, table, fields, wvl=None, foc=0.0, io=None):
    """Image intercept of the (0, 0) pupil ray of every field at ``wvl``
    (default: central wavelength): ``ref_sphere[0]`` of
    raytr/waveabr.py:24-76 for ``image_pt_2d=None``."""
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    wvl = sm.central_wavelength() if wvl is None else wvl
    recs, eprad, z_pupil = grid_fields_of(opt_model, fields)
    g0 = E.PupilGrid(recs, [table.wvl_index(wvl)], [0.0], [0.0], eprad, z_pupil,
                     apply_vignetting=False, flip_z_dir=sm.z_dir[0], foc=foc, device=table.device)
    r0 = E.trace_grid(table, g0, outputs=('p',), summary=False, check_apertures=False)
    ref = r0.p[:2].t().contiguous().cpu().numpy()
    if io is not None:
        io['h2d'] += g0.host_bytes()
        io['d2h'] += ref.nbytes
    g0.close()
    return ref        # [n_fields, 2]


def spot_diagram(opt_model, num_rays=21, fields=None, wvls=None, foc=None, table=None,
                 device=0, pinned=None, shard=None, group=None, pieces=8, chunk_range=None, **kwargs):
    """Spot diagrams of all fields and wavelengths in one pass over the device.

    Host buffers in, host buffers out: the grid description goes to the device (one
    asynchronous copy into a re-used block), the chief-ray reference image points are
    computed there (``rt_grid_chief_ref``: no host round trip), ``rt_trace_grid`` generates
    and traces the rays, and the transverse aberrations come back into pinned host memory,
    16 B per ray (status / failing surface of the rays that do not arrive ride in the NaN
    payloads).  ``shard=(rank, world)`` traces only that rank's slice of the chunk space and
    all-gathers the per-(field, wvl) sums over ``group`` (parallel.py); ``chunk_range=(c0, c1)``
    overrides the equal-length slice (e.g. ``parallel.shard_chunks_weighted``).  ``pinned``:
    optional dict with a pinned host tensor ``abr`` ``[2, >=n]`` to re-use across calls.
    ``pieces``: the chunk range is traced in that many launches on two streams so that the
    device->host copy of one piece overlaps the trace of the next."""
    from .parallel import shard_chunks, gather_summaries
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    table = _table_for(opt_model, table, device)
    fields = list(osp.field_of_view.fields if fields is None else fields)
    wvls = list(sm.wvlns if wvls is None else wvls)
    foc = osp.defocus.focus_shift if foc is None else foc
    io = {'h2d': 0, 'd2h': 0}
    dev = torch.device('cuda', table.device)
    with torch.cuda.device(dev):
        grid, spec = _reusable_grid(opt_model, table, num_rays, fields, wvls, foc)
        io['h2d'] += spec.host_bytes()
        ws = _per_thread('work').setdefault(table.device, {})     # re-used device staging buffers
        if ws.get('ref') is None or ws['ref'].shape[0] != len(fields):
            ws['ref'] = torch.empty((len(fields), 2), dtype=torch.float64, device=dev)
        ref_dev = ws['ref']
        grid.chief_ref(table, table.wvl_index(sm.central_wavelength()), out=ref_dev)
        c0, c1 = (0, grid.n_chunks) if shard is None else shard_chunks(grid.n_chunks, *shard)
        if chunk_range is not None:          # this rank's range, e.g. parallel.shard_chunks_weighted
            c0, c1 = chunk_range
        n = grid.rays_in_chunks(c0, c1)
        if pinned is None:
            pinned = {'abr': torch.empty((2, max(n, 1)), dtype=torch.float64).pin_memory()}
        h_abr = pinned['abr'][:, :n]
        # the chunk range is traced in `pieces` launches on two streams of the grid handle, each
        # followed by its device->host copy (rt_trace_grid_to_host): one C call, no per-piece Python
        summ, ws['trace'] = E.trace_grid_to_host(table, grid, pinned['abr'], c0, c1,
                                                 pieces=max(1, min(pieces, (c1 - c0)//64)),
                                                 workspace=ws.get('trace'), **kwargs)
        if shard is not None:
            summ = gather_summaries(summ, group)
        tail = torch.cat([summ.reshape(-1), ref_dev.reshape(-1)]).cpu().numpy()   # one small copy; waits
        torch.cuda.current_stream(dev).synchronize()
    summ_host = tail[:summ.numel()].reshape(summ.shape)
    ref = tail[summ.numel():].reshape(len(fields), 2).copy()
    stats_host = {k: np.asarray(v).reshape(len(fields), len(wvls))
                  for k, v in E.spot_statistics(summ_host).items()}
    io['d2h'] += h_abr.numel()*8 + tail.nbytes
    return SpotDiagram(h_abr.numpy(), None, stats_host, ref, num_rays, len(fields), len(wvls),
                       grid.first_ray_of_chunk(c0), grid.n_rays, io)


# --------------------------------------------------------------------------
# RayFan / RayList / RayGrid: the reference's analysis classes
# (/root/reference/src/rayoptics/raytr/analyses.py:121-187,343-434,584-663) with
# the same constructor arguments and result attributes, each evaluated by one
# grid launch (chief rays: one more tiny launch).
# --------------------------------------------------------------------------
from . import waveabr as W                      # noqa: E402
from . import sampler                           # noqa: E402


def _resolve(opt_model, f, wl, foc):
    osp = opt_model.optical_spec
    fld = osp.field_of_view.fields[f] if isinstance(f, int) else f
    wvl = osp.spectral_region.central_wvl if wl is None else wl
    foc = osp.defocus.focus_shift if foc is None else foc
    return fld, wvl, foc


def _trace_pupil_points(opt_model, table, fld, wvl, foc, px, py, paired, apply_vignetting,
                        check_apertures, image_pt_2d, image_delta, want_opd, backend=None):
    """One (field, wvl) tile of pupil points -> host dict(pupil, abr, opd, status).
    ``backend``: test seam (object with ``chief_rays(opt_model, fields, wvls)`` and
    ``trace_tile(opt_model, spec, want_opd, check_apertures)``); None = the CUDA engine."""
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    wave, ref_img, pkgs = W.setup_tiles(opt_model, table, [fld], [wvl], foc, image_pt_2d, image_delta,
                                        chief_tracer=None if backend is None else backend.chief_rays)
    recs, eprad, z_pupil = grid_fields_of(opt_model, [fld])
    if backend is not None:
        spec = E.PupilGridSpec(recs, [sm.index_for_wavelength(wvl)], px, py, eprad, z_pupil,
                               ref_img=ref_img, apply_vignetting=apply_vignetting,
                               flip_z_dir=sm.z_dir[0], foc=foc, paired=paired,
                               wave=wave if want_opd else None)
        out = backend.trace_tile(opt_model, spec, want_opd, check_apertures)
        out.update(ref_sphere=pkgs[0][0][1], chief_ray=pkgs[0][0][0])
        return out
    grid = E.PupilGrid(recs, [table.wvl_index(wvl)], px, py, eprad, z_pupil, ref_img=ref_img,
                       apply_vignetting=apply_vignetting, flip_z_dir=sm.z_dir[0], foc=foc,
                       paired=paired, wave=wave if want_opd else None, device=table.device)
    outs = ('abr', 'status') + (('opd',) if want_opd else ())
    res = E.trace_grid(table, grid, outputs=outs, summary=False, check_apertures=check_apertures)
    out = {'abr': res.abr.cpu().numpy(), 'status': res.status.cpu().numpy(),
           'opd': res.opd.cpu().numpy() if want_opd else None, 'ref_sphere': pkgs[0][0][1],
           'chief_ray': pkgs[0][0][0]}
    grid.close()
    return out


def _vignetted(fld, px, py, apply_vignetting):
    """pupil coordinates as the reference records them (after Field.apply_vignetting
    when it is applied: trace.trace_grid forwards the vignetted values)."""
    if not apply_vignetting:
        return np.array(px, dtype=float), np.array(py, dtype=float)
    vx, vy = np.array(px, dtype=float), np.array(py, dtype=float)
    for k in range(len(vx)):
        v = fld.apply_vignetting(np.array([vx[k], vy[k]]))
        vx[k], vy[k] = v[0], v[1]
    return vx, vy


class Ray:
    """A ray at the given field and wavelength (analyses.py:46-118): ``ray_seg`` (the
    segment at ``srf_indx``), ``ray_pkg`` when ``srf_save='all'``, ``t_abr`` the
    transverse aberration w.r.t. the reference image point."""

    def __init__(self, opt_model, p, f=0, wl=None, foc=None, image_pt_2d=None, image_delta=None,
                 srf_indx=-1, srf_save='single', output_filter=None, rayerr_filter=None,
                 color=None, clip_rays=False, table=None, device=0, tracer=None):
        self.opt_model = opt_model
        self.pupil = p
        self.fld, self.wvl, self.foc = _resolve(opt_model, f, wl, foc)
        self.image_pt_2d, self.image_delta = image_pt_2d, image_delta
        self.output_filter, self.rayerr_filter = output_filter, rayerr_filter
        self.clip_rays = clip_rays
        self.color = color
        self.srf_save, self.srf_indx = srf_save, srf_indx
        self._engine = dict(table=table, device=device, tracer=tracer)
        self.update_data()

    def update_data(self, **kwargs):
        from . import trace
        ref_sphere, cr_pkg = trace.setup_pupil_coords(self.opt_model, self.fld, self.wvl, self.foc,
                                                      image_pt=self.image_pt_2d,
                                                      image_delta=self.image_delta, **self._engine)
        build = kwargs.pop('build', 'rebuild')
        if build == 'rebuild':
            ray_pkg, ray_err = trace.trace_safe(self.opt_model, self.pupil, self.fld, self.wvl,
                                                self.output_filter, self.rayerr_filter,
                                                use_named_tuples=True,
                                                check_apertures=self.clip_rays, **self._engine,
                                                **kwargs)
            if ray_pkg is None:
                raise ray_err if ray_err is not None else RuntimeError('ray failed')
            self.ray_seg = ray_pkg.ray[self.srf_indx]
            if self.srf_save == 'all':
                self.ray_pkg = ray_pkg
        ray_seg = self.ray_seg
        dist = self.foc/ray_seg[1][2]
        defocused_pt = ray_seg[0] + dist*ray_seg[1]
        reference_image_pt = ref_sphere[0]
        self.t_abr = defocused_pt[:2] - reference_image_pt[:2]
        return self


class RayFan:
    """A fan of rays across the pupil (analyses.py:121-187).

    ``fan``: list of ``((pupil_x, pupil_y), (dx, dy, opd))`` for the rays that
    reach the image, ``opd`` in waves -- the output of ``focus_fan``
    (analyses.py:313-339)."""

    def __init__(self, opt_model, f=0, wl=None, foc=None, image_pt_2d=None, image_delta=None,
                 num_rays=21, xyfan='y', output_filter=None, rayerr_filter=None, color=None,
                 clip_rays=False, table=None, device=0, backend=None, **kwargs):
        self.opt_model = opt_model
        self._backend = backend
        self.fld, self.wvl, self.foc = _resolve(opt_model, f, wl, foc)
        self.image_pt_2d, self.image_delta = image_pt_2d, image_delta
        self.num_rays = num_rays
        self.xyfan = 0 if xyfan == 'x' else (1 if xyfan == 'y' else int(xyfan))
        self.color = color
        self.clip_rays = clip_rays
        self._table = None if backend is not None else _table_for(opt_model, table, device)
        self.update_data()

    def update_data(self, **kwargs):
        t = E.accumulated_steps(-1.0, 1.0, self.num_rays)
        zeros = E.accumulated_steps(0.0, 0.0, self.num_rays)
        px, py = (t, zeros) if self.xyfan == 0 else (zeros, t)
        r = _trace_pupil_points(self.opt_model, self._table, self.fld, self.wvl, self.foc, px, py,
                                True, True, self.clip_rays, self.image_pt_2d, self.image_delta, True,
                                backend=self._backend)
        convert_to_opd = 1/self.opt_model.nm_to_sys_units(self.wvl)
        vx, vy = _vignetted(self.fld, px, py, True)
        self.fan = [((vx[k], vy[k]), (r['abr'][0, k], r['abr'][1, k], convert_to_opd*r['opd'][k]))
                    for k in range(self.num_rays) if r['status'][k] == 0]
        return self


class RayList:
    """Rays from a list / generator of pupil coordinates (analyses.py:343-434).

    ``ray_abr``: ``[2, n_ok]`` transverse aberrations of the rays that reach the
    image (``np.rollaxis(focus_pupil_coords(...), 1)``, analyses.py:432)."""

    def __init__(self, opt_model, pupil_gen=None, pupil_coords=None, num_rays=21, f=0, wl=None,
                 foc=None, image_pt_2d=None, image_delta=None, output_filter=None,
                 rayerr_filter=None, clip_rays=False, apply_vignetting=True, table=None, device=0,
                 backend=None, **kwargs):
        self.opt_model = opt_model
        self._backend = backend
        if pupil_coords is not None and pupil_gen is None:
            self.pupil_coords, self.pupil_gen = pupil_coords, None
        else:
            if pupil_gen is None:
                grid_def = [np.array([-1., -1.]), np.array([1., 1.]), num_rays]
                pupil_gen = (sampler.csd_grid_ray_generator, (grid_def,), {})
            self.pupil_gen = pupil_gen
        self.fld, self.wvl, self.foc = _resolve(opt_model, f, wl, foc)
        self.image_pt_2d, self.image_delta = image_pt_2d, image_delta
        self.apply_vignetting = apply_vignetting
        # trace_pupil_coords defaults check_apertures to True when the key is absent;
        # RayList always passes clip_rays (analyses.py:389,554)
        self.clip_rays = clip_rays
        self._table = None if backend is not None else _table_for(opt_model, table, device)
        self.update_data()

    def update_data(self, **kwargs):
        if self.pupil_gen:
            fct, args, kwa = self.pupil_gen
            if fct in sampler._ARRAY_FORM and not kwa:      # whole sample set as one array
                self.pupil_coords = pts = sampler.sample_points(fct, *args)
            else:
                self.pupil_coords = fct(*args, **kwa)
                pts = None
        else:
            pts = None
        if pts is None:
            pts = np.array([np.array(p, dtype=float) for p in self.pupil_coords]).reshape(-1, 2)
        r = _trace_pupil_points(self.opt_model, self._table, self.fld, self.wvl, self.foc,
                                pts[:, 0], pts[:, 1], True, self.apply_vignetting, self.clip_rays,
                                self.image_pt_2d, self.image_delta, False, backend=self._backend)
        ok = r['status'] == 0
        self.ray_abr = r['abr'][:, ok]
        self.pupil = pts[ok].T
        return self


class RayGrid:
    """Square grid of rays over the vignetted pupil -> wavefront map
    (analyses.py:584-663).  ``grid``: ``[3, num, num]`` = pupil x, pupil y, OPD
    in waves (``value_if_none`` where the ray does not reach the image)."""

    def __init__(self, opt_model, f=0, wl=None, foc=None, image_pt_2d=None, image_delta=None,
                 output_filter=None, rayerr_filter=None, num_rays=21, clip_rays=True,
                 value_if_none=np.nan, oversize=1., table=None, device=0, backend=None, **kwargs):
        self.opt_model = opt_model
        self._backend = backend
        self.fld, self.wvl, self.foc = _resolve(opt_model, f, wl, foc)
        self.image_pt_2d, self.image_delta = image_pt_2d, image_delta
        self.num_rays, self.value_if_none, self.oversize = num_rays, value_if_none, oversize
        self.clip_rays = clip_rays
        self._table = None if backend is not None else _table_for(opt_model, table, device)
        self.update_data()

    def vignetting_bbox(self):
        """Field.vignetting_bbox (raytr/opticalspec.py:1326-1333)."""
        poly = [self.fld.apply_vignetting(list(pr)) for pr in self.opt_model.optical_spec.pupil.pupil_rays]
        poly = np.array(poly)
        return self.oversize*np.array([poly.min(axis=0), poly.max(axis=0)])

    def update_data(self, **kwargs):
        bbox = self.vignetting_bbox()
        n = self.num_rays
        px = E.accumulated_steps(bbox[0][0], bbox[1][0], n)
        py = E.accumulated_steps(bbox[0][1], bbox[1][1], n)
        # trace_ray_grid: apply_vignetting defaults to False (analyses.py:674)
        r = _trace_pupil_points(self.opt_model, self._table, self.fld, self.wvl, self.foc, px, py,
                                False, False, self.clip_rays, self.image_pt_2d, self.image_delta, True,
                                backend=self._backend)
        convert_to_opd = 1/self.opt_model.nm_to_sys_units(self.wvl)
        opd = np.where(r['status'] == 0, convert_to_opd*r['opd'], self.value_if_none).reshape(n, n)
        gx, gy = np.meshgrid(px, py, indexing='ij')
        self.grid = np.stack([gx, gy, opd])
        return self


# --- the functional forms of the analyses (analyses.py:233-273,513-542,699-732): one launch each,
#     same arguments and return values ---------------------------------------------------------
def _engine_kw(kwargs):
    return {k: kwargs.pop(k) for k in ('table', 'device', 'backend') if k in kwargs}


def _tile(opt_model, fld, wvl, foc, px, py, paired, apply_vignetting, check_apertures,
          image_pt_2d, image_delta, want_opd, eng):
    backend = eng.get('backend')
    table = None if backend is not None else _table_for(opt_model, eng.get('table'),
                                                        eng.get('device', 0))
    r = _trace_pupil_points(opt_model, table, fld, wvl, foc, px, py, paired, apply_vignetting,
                            check_apertures, image_pt_2d, image_delta, want_opd, backend=backend)
    fld.chief_ray, fld.ref_sphere = r['chief_ray'], r['ref_sphere']
    return r


def eval_fan(opt_model, fld, wvl, foc, xy, image_pt_2d=None, image_delta=None, num_rays=21,
             output_filter=None, rayerr_filter=None, **kwargs):
    """Trace a fan of rays and evaluate dx, dy, & OPD across the fan (analyses.py:233-273)."""
    eng = _engine_kw(kwargs)
    t = E.accumulated_steps(-1.0, 1.0, num_rays)
    zeros = E.accumulated_steps(0.0, 0.0, num_rays)
    px, py = (t, zeros) if xy == 0 else (zeros, t)
    apply_vig = kwargs.get('apply_vignetting', True)
    r = _tile(opt_model, fld, wvl, foc, px, py, True, apply_vig, kwargs.get('check_apertures', False),
              image_pt_2d, image_delta, True, eng)
    convert_to_opd = 1/opt_model.nm_to_sys_units(wvl)
    vx, vy = _vignetted(fld, px, py, apply_vig)
    return [((vx[k], vy[k]), (r['abr'][0, k], r['abr'][1, k], convert_to_opd*r['opd'][k]))
            for k in range(num_rays) if r['status'][k] == 0]


def eval_pupil_coords(opt_model, fld, wvl, foc, image_pt_2d=None, image_delta=None, num_rays=21,
                      **kwargs):
    """Trace a square grid of rays and return the transverse aberrations ``[n_ok, 2]``
    (analyses.py:513-542)."""
    eng = _engine_kw(kwargs)
    grid_def = [np.array([-1., -1.]), np.array([1., 1.]), num_rays]
    pts = sampler.square_grid_points(grid_def)
    r = _tile(opt_model, fld, wvl, foc, pts[:, 0], pts[:, 1], True,
              kwargs.get('apply_vignetting', True), kwargs.get('check_apertures', True),
              image_pt_2d, image_delta, False, eng)
    ok = r['status'] == 0
    return r['abr'][:, ok].T.copy()


def eval_wavefront(opt_model, fld, wvl, foc, image_pt_2d=None, image_delta=None, num_rays=21,
                   value_if_none=np.nan, **kwargs):
    """Trace a grid of rays over the vignetted pupil and evaluate the OPD: ``[num, num, 3]`` of
    (pupil x, pupil y, OPD in waves) (analyses.py:699-732)."""
    eng = _engine_kw(kwargs)
    bbox = fld.vignetting_bbox(opt_model['optical_spec']['pupil'], oversize=kwargs.get('oversize', 1.))
    px = E.accumulated_steps(bbox[0][0], bbox[1][0], num_rays)
    py = E.accumulated_steps(bbox[0][1], bbox[1][1], num_rays)
    r = _tile(opt_model, fld, wvl, foc, px, py, False, kwargs.get('apply_vignetting', False),
              kwargs.get('check_apertures', True), image_pt_2d, image_delta, True, eng)
    convert_to_opd = 1/opt_model.nm_to_sys_units(wvl)
    opd = np.where(r['status'] == 0, convert_to_opd*r['opd'], value_if_none).reshape(num_rays, num_rays)
    gx, gy = np.meshgrid(px, py, indexing='ij')
    return np.stack([gx, gy, opd], axis=2)


def select_plot_data(fan, xyfan, data_type):
    """Given a fan of data, select the sample points and the resulting data (analyses.py:190-199)"""
    f_x = np.array([p[xyfan] for p, val in fan])
    f_y = np.array([val[data_type] for p, val in fan])
    return f_x, f_y


def smooth_plot_data(f_x, f_y, num_points=100):
    """Interpolate fan data points and return a smoothed version (analyses.py:202-209)"""
    from scipy.interpolate import interp1d
    interpolator = interp1d(f_x, f_y, kind='cubic', assume_sorted=True)
    x_sample = np.linspace(f_x.min(), f_x.max(), num_points)
    return x_sample, interpolator(x_sample)


def update_psf_data(pupil_grid, build='rebuild'):
    """analyses.py:878-883"""
    pupil_grid.update_data(build=build)
    return calc_psf(pupil_grid.grid[2], pupil_grid.num_rays, pupil_grid.maxdim)


# the reference's per-ray loops of this module, batched (rayoptics_b200/trace.py)
from .trace import (analyses_trace_ray_fan as trace_ray_fan,       # noqa: E402
                    analyses_trace_ray_list as trace_ray_list,
                    analyses_trace_ray_grid as trace_ray_grid)


# --- trace once, refocus often (analyses.py:276-339,545-580,735-791): the first stage keeps the
#     whole rays of ONE launch (plus the focus-independent part of every OPD), the second stage is
#     host arithmetic on them -- no retrace when only ``foc`` / the image point changes.
#     (``eval_*`` above do both stages on the device and copy back 16-24 B per ray; these exist
#     for callers that hold on to the traced rays, e.g. the reference's focus sliders.)
def _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs):
    from . import trace as TR
    eng = {k: kwargs[k] for k in ('table', 'device', 'tracer') if k in kwargs}
    ref_sphere, cr_pkg = TR.setup_pupil_coords(opt_model, fld, wvl, foc, image_pt=image_pt_2d,
                                               image_delta=image_delta, **eng)
    return ref_sphere, cr_pkg


def _refocused(ray_pkg, foc, image_pt):
    seg = ray_pkg[0][-1]
    dist = foc/seg[1][2]
    defocused_pt = seg[0] + dist*seg[1]
    return defocused_pt - image_pt


def _pre_calc(opt_model, fld, wvl, foc, ray_pkg, cr_pkg, ref_sphere):
    if ray_pkg is None or isinstance(ray_pkg, Exception):
        return None
    fod = opt_model['analysis_results']['parax_data'].fod
    return W.wave_abr_pre_calc(fod, fld, wvl, foc, ray_pkg, cr_pkg, ref_sphere)


def trace_fan(opt_model, fld, wvl, foc, xy, image_pt_2d=None, image_delta=None, num_rays=21,
              output_filter=None, rayerr_filter=None, **kwargs):
    """Trace a fan of rays and precalculate data for rapid refocus later (analyses.py:276-310):
    ``(fan, upd_fan)`` = ``[[pupil_x, pupil_y, ray_pkg], ...]`` and the matching
    ``wave_abr_pre_calc`` tuples."""
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    fld.chief_ray, fld.ref_sphere = cr_pkg, ref_sphere
    fan_start, fan_stop = np.array([0., 0.]), np.array([0., 0.])
    fan_start[xy], fan_stop[xy] = -1.0, 1.0
    fan = trace_ray_fan(opt_model, [fan_start, fan_stop, num_rays], fld, wvl, foc,
                        output_filter=output_filter, rayerr_filter=rayerr_filter, **kwargs)
    upd_fan = [_pre_calc(opt_model, fld, wvl, foc, fi[2], cr_pkg, ref_sphere) for fi in fan]
    return fan, upd_fan


def focus_fan(opt_model, fan_pkg, fld, wvl, foc, image_pt_2d=None, image_delta=None, **kwargs):
    """Refocus the fan of rays and return the transverse aberration and OPD (analyses.py:313-339):
    ``[((pupil_x, pupil_y), (dx, dy, opd in waves)), ...]``."""
    fod = opt_model['analysis_results']['parax_data'].fod
    fan, upd_fan = fan_pkg
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    convert_to_opd = 1/opt_model.nm_to_sys_units(wvl)
    fan_data = []
    for (pupil_x, pupil_y, ray_pkg), pre in zip(fan, upd_fan):
        if ray_pkg is None or isinstance(ray_pkg, Exception):
            fan_data.append((pupil_x, pupil_y, np.nan))
            continue
        t_abr = _refocused(ray_pkg, foc, ref_sphere[0])
        opd = convert_to_opd*W.wave_abr_calc(fod, fld, wvl, foc, ray_pkg, cr_pkg, pre, ref_sphere)
        fan_data.append(((pupil_x, pupil_y), (t_abr[0], t_abr[1], opd)))
    return fan_data


def trace_pupil_coords(opt_model, pupil_coords, fld, wvl, foc, image_pt_2d=None, image_delta=None,
                       **kwargs):
    """Trace a list of rays and return data needed for rapid refocus (analyses.py:545-558)."""
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    fld.chief_ray, fld.ref_sphere = cr_pkg, ref_sphere
    kwargs['check_apertures'] = kwargs.get('check_apertures', True)
    return trace_ray_list(opt_model, pupil_coords, fld, wvl, foc, **kwargs)


def focus_pupil_coords(opt_model, ray_list, fld, wvl, foc, image_pt_2d=None, image_delta=None,
                       **kwargs):
    """Given pre-traced rays and a reference sphere, return the transverse aberrations
    (analyses.py:561-580): ``[n, 2]`` (``nan`` entries for rays recorded as None)."""
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    data = []
    for pupil_x, pupil_y, ray_pkg in ray_list:
        if ray_pkg is None:
            data.append(np.nan)
        else:
            t_abr = _refocused(ray_pkg, foc, ref_sphere[0])
            data.append((t_abr[0], t_abr[1]))
    return np.array(data)


def trace_wavefront(opt_model, fld, wvl, foc, image_pt_2d=None, image_delta=None, num_rays=21,
                    **kwargs):
    """Trace a grid of rays over the vignetted pupil and pre-calculate data needed for rapid
    refocus (analyses.py:735-766): ``(grid, upd_grid)``, rows of ``[pupil_x, pupil_y, ray_pkg]``."""
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    fld.chief_ray, fld.ref_sphere = cr_pkg, ref_sphere
    vig_bbox = fld.vignetting_bbox(opt_model['optical_spec']['pupil'],
                                   oversize=kwargs.pop('oversize', 1.))
    kwargs['check_apertures'] = kwargs.get('check_apertures', True)
    grid = trace_ray_grid(opt_model, [vig_bbox[0], vig_bbox[1], num_rays], fld, wvl, foc, **kwargs)
    upd_grid = [[_pre_calc(opt_model, fld, wvl, foc, gij[2], cr_pkg, ref_sphere) for gij in row]
                for row in grid]
    return grid, upd_grid


def focus_wavefront(opt_model, grid_pkg, fld, wvl, foc, image_pt_2d=None, image_delta=None,
                    value_if_none=np.nan, **kwargs):
    """Given pre-traced rays and a reference sphere, return the rays' OPD (analyses.py:769-791):
    ``[num, num, 3]`` of (pupil x, pupil y, OPD in waves)."""
    fod = opt_model['analysis_results']['parax_data'].fod
    grid, upd_grid = grid_pkg
    ref_sphere, cr_pkg = _stage_setup(opt_model, fld, wvl, foc, image_pt_2d, image_delta, kwargs)
    convert_to_opd = 1/opt_model.nm_to_sys_units(wvl)
    out = []
    for row, upd_row in zip(grid, upd_grid):
        out_row = []
        for (pupil_x, pupil_y, ray_pkg), pre in zip(row, upd_row):
            if ray_pkg is None:
                out_row.append((pupil_x, pupil_y, value_if_none))
            else:
                opd = convert_to_opd*W.wave_abr_calc(fod, fld, wvl, foc, ray_pkg, cr_pkg, pre,
                                                     ref_sphere)
                out_row.append((pupil_x, pupil_y, opd))
        out.append(out_row)
    return np.array(out)


# --- raw ray list (analyses.py:458-510) ------------------------------------------------------
def _cuda_bundle_tracer(opt_model, table, p0, d0, wvl_idx, trace_kwargs):
    res = E.trace_bundle(table, p0, d0, wvl_idx=wvl_idx, full=True,
                         outputs=('op', 'status', 'fail_surf', 'n_seg'), **trace_kwargs)
    return {'full': res.full.cpu().numpy(), 'op': res.op.cpu().numpy(),
            'status': res.status.cpu().numpy(), 'fail_surf': res.fail_surf.cpu().numpy(),
            'n_seg': res.n_seg.cpu().numpy()}


def trace_list_of_rays(opt_model, rays, output_filter=None, rayerr_filter=None, table=None,
                       device=0, tracer=None, **kwargs):
    """Trace a list of rays ``(pt0, dir0, wvl)`` and return the ray packages in a list
    (analyses.py:458-510): one bundle launch instead of one ``trace()`` per ray, same
    ``output_filter`` / ``rayerr_filter`` conventions and the same list out."""
    from . import raytrace as RT
    from . import trace as TR
    sm = opt_model.seq_model
    rays = list(rays)
    if not rays:
        return []
    p0 = np.array([np.asarray(r[0], dtype=float) for r in rays]).T.copy()
    d0 = np.array([np.asarray(r[1], dtype=float) for r in rays]).T.copy()
    wvl_idx = np.array([sm.index_for_wavelength(r[2]) for r in rays], dtype=np.int32)
    kw = {k: v for k, v in kwargs.items() if k in TR._TRACE_RAW_KEYS}
    kw.setdefault('first_surf', 1)                       # raytrace.trace defaults (raytrace.py:77-79)
    kw.setdefault('last_surf', sm.get_num_surfaces() - 2)
    if tracer is None:
        table = _table_for(opt_model, table, device)
        tracer = _cuda_bundle_tracer
    r = tracer(opt_model, table, p0, d0, wvl_idx, kw)
    paths = {}
    ray_list = []
    for k, ray in enumerate(rays):
        wvl = ray[2]
        if wvl not in paths:
            paths[wvl] = list(sm.path(wvl))
        pkg, err = RT.package_ray(paths[wvl], r['full'][:, :, k], float(r['op'][k]),
                                  int(r['status'][k]), int(r['fail_surf'][k]), int(r['n_seg'][k]), wvl)
        if err is not None:
            if rayerr_filter == 'full':
                ray_list.append((ray, err))
            elif rayerr_filter == 'summary':
                err.ray_pkg = None
                ray_list.append((ray, err))
            continue
        pkg = TR.RayPkg(*pkg)
        if output_filter is None:
            ray_list.append(pkg)
        elif output_filter == 'last':
            rr, op_delta, w = pkg
            ray_list.append((rr[-1], op_delta, w))
        else:
            ray_list.append(output_filter(pkg))
    return ray_list


# --- PSF from a wavefront map (raytr/analyses.py:795-875) -------------------
def psf_sampling(n=None, n_pupil=None, n_airy=None):
    """Given 2 of (grid width, pupil samples, Airy-peak samples) compute the third
    (analyses.py:795-815)."""
    npa = n, n_pupil, n_airy
    i = npa.index(None)
    if i == 0:
        n = round((n_pupil*n_airy)/2.44)
    elif i == 1:
        n_pupil = round(2.44*n/n_airy)
    else:
        n_airy = round(2.44*n/n_pupil)
    return n, n_pupil, n_airy


def calc_psf_scaling(opt_model, fld, wvl, ndim, maxdim):
    """Input / output grid spacings of the FFT PSF (analyses.py:818-845): ``(delta_x, delta_xp)``,
    the linear grid spacing on the entrance pupil and on the image plane.  ``fld.ref_sphere``
    must be set (the analysis classes / ``trace.setup_pupil_coords`` do)."""
    fod = opt_model['analysis_results']['parax_data'].fod
    wl = opt_model.nm_to_sys_units(wvl)
    fill_factor = ndim/maxdim
    max_D = 2*fod.enp_radius/fill_factor
    delta_x = max_D/maxdim
    C = wl/fod.exp_radius
    delta_theta = (fill_factor*C)/2
    ref_sphere_radius = fld.ref_sphere[2]
    delta_xp = delta_theta*ref_sphere_radius
    return delta_x, delta_xp


def calc_psf(wavefront, ndim, maxdim, device=0):
    """Point spread function of a wavefront map (analyses.py:848-875): embed the
    ``ndim x ndim`` OPD map (waves, NaN = no data) in a ``maxdim`` grid, form the
    pupil function ``exp(2 pi i W)`` (entries equal to 1 are zeroed, as the
    reference does), FFT, normalise to the peak.  The FFT is ``torch.fft`` on the
    GPU (a plain library transform; tolerance vs numpy 1e-12)."""
    dev = torch.device('cuda', device)
    Wm = torch.zeros((maxdim, maxdim), dtype=torch.float64, device=dev)
    w = torch.as_tensor(np.nan_to_num(np.asarray(wavefront, dtype=np.float64)), device=dev)
    m2, nd2 = maxdim//2, ndim//2
    Wm[m2 - (nd2 - 1):m2 + (nd2 + 1), m2 - (nd2 - 1):m2 + (nd2 + 1)] = w
    phase = torch.exp(1j*2*np.pi*Wm.to(torch.complex128))
    phase = torch.where(phase == 1, torch.zeros_like(phase), phase)
    AP = torch.fft.fftshift(torch.fft.fft2(torch.fft.fftshift(phase))).abs()**2
    AP = AP/AP.max()
    return AP.cpu().numpy()
