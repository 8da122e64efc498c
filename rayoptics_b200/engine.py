"""Ray bundles and pupil grids on the GPU: PyTorch tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, streams); every ray is traced by the
hand-written sm_100a kernels in csrc/ through ``rt_trace_bundle`` /
``rt_trace_grid`` (include/b200rt.h).  Ray bundles are structure-of-arrays:
``p`` and ``d`` are ``[3, n]`` float64 tensors (row c = component c), so each
component is a contiguous array and every warp load/store is coalesced.

Replaces the per-ray Python loops of the reference:
``trace_list_of_rays`` (/root/reference/src/rayoptics/raytr/analyses.py:458-510),
``trace_grid`` (raytr/trace.py:563-605), ``trace_ray_grid`` (analyses.py:666-696).
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _abi
from ._abi import rt_out, rt_grid_spec, rt_field_desc, RT_SEG_DOUBLES, RT_SUMMARY_DOUBLES

SUMMARY_FIELDS = ('n_ok', 'n_missed', 'n_tir', 'n_blocked', 'n_other',
                  'sum_x', 'sum_y', 'sum_xx', 'sum_yy', 'sum_xy',
                  'min_x', 'max_x', 'min_y', 'max_y', 'sum_op', 'reserved')


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_soa(x, device, name):
    """[3, n] float64 device tensor (rows contiguous)."""
    if isinstance(x, (tuple, list)) and len(x) == 3 and torch.is_tensor(x[0]):
        x = torch.stack([xi.to(device=device, dtype=torch.float64) for xi in x])
    elif not torch.is_tensor(x):
        x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
    if x.dim() != 2:
        raise ValueError(f'{name} must be 2-D')
    if x.shape[0] != 3 and x.shape[1] == 3:
        x = x.t()
    if x.shape[0] != 3:
        raise ValueError(f'{name} must have shape [3, n] or [n, 3]')
    return x.to(device=device, dtype=torch.float64).contiguous()


BUNDLE_OUTPUTS = ('p', 'd', 'nrml', 'dst', 'op', 'status', 'fail_surf', 'n_seg')
GRID_OUTPUTS = ('p', 'd', 'op', 'status', 'fail_surf', 'abr')


class BundleResult:
    """Structure-of-arrays result of a bundle / grid trace (device tensors).

    ``p``, ``d``, ``nrml``: ``[3, n]`` last ray segment (``ray[-1]`` of the
    reference's RayPkg); ``dst``, ``op``: ``[n]``; ``status``, ``fail_surf``,
    ``n_seg``: ``[n]`` int32; ``full``: ``[n_ifc, 10, n]`` whole rays;
    ``abr``: ``[2, n]`` transverse aberration (grid traces).  Only the
    ``outputs`` asked for are allocated and written (others are None)."""

    def __init__(self, n, n_ifc, device, outputs=BUNDLE_OUTPUTS, nan_status=False):
        self.nan_status = bool(nan_status)     # RT_OUT_ABR_NAN_STATUS: status / fail_surf ride in abr's NaNs
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        want = set(outputs)
        unknown = want - set(BUNDLE_OUTPUTS) - {'full', 'abr', 'opd'}
        if unknown:
            raise ValueError(f'unknown outputs {sorted(unknown)}')
        self.n = n
        self.p = torch.empty((3, n), **f64) if 'p' in want else None
        self.d = torch.empty((3, n), **f64) if 'd' in want else None
        self.nrml = torch.empty((3, n), **f64) if 'nrml' in want else None
        self.dst = torch.empty(n, **f64) if 'dst' in want else None
        self.op = torch.empty(n, **f64) if 'op' in want else None
        self.status = torch.empty(n, **i32) if 'status' in want else None
        self.fail_surf = torch.empty(n, **i32) if 'fail_surf' in want else None
        self.n_seg = torch.empty(n, **i32) if 'n_seg' in want else None
        self.full = (torch.full((n_ifc, RT_SEG_DOUBLES, n), float('nan'), **f64)
                     if 'full' in want else None)
        self.abr = torch.empty((2, n), **f64) if 'abr' in want else None
        self.opd = torch.empty(n, **f64) if 'opd' in want else None
        self.summary = None

    def c_struct(self):
        o = rt_out()
        if self.p is not None:
            o.px, o.py, o.pz = (_ptr(self.p[i]) for i in range(3))
        if self.d is not None:
            o.dx, o.dy, o.dz = (_ptr(self.d[i]) for i in range(3))
        if self.nrml is not None:
            o.nx, o.ny, o.nz = (_ptr(self.nrml[i]) for i in range(3))
        o.dst, o.op = _ptr(self.dst), _ptr(self.op)
        o.status, o.fail_surf, o.n_seg = _ptr(self.status), _ptr(self.fail_surf), _ptr(self.n_seg)
        if self.full is not None:
            o.full = _ptr(self.full)
            o.full_stride = self.n
        if self.abr is not None:
            o.abr_x, o.abr_y = _ptr(self.abr[0]), _ptr(self.abr[1])
        o.opd = _ptr(self.opd)
        o.flags = _abi.RT_OUT_ABR_NAN_STATUS if self.nan_status else 0
        return o

    def bytes_per_ray(self):
        """bytes the kernel writes per ray for the allocated outputs"""
        b = 0
        for t in (self.p, self.d, self.nrml, self.abr, self.dst, self.op, self.status,
                  self.fail_surf, self.n_seg, self.opd):
            if t is not None:
                b += t.element_size()*(t.numel()//max(self.n, 1))
        return b

    def ok(self):
        return self.status == 0


def trace_bundle(table, p, d, wvl_idx=None, full=False, outputs=BUNDLE_OUTPUTS, **kwargs):
    """Trace ``n`` rays given start points / direction cosines in the object
    interface's coordinates.  ``kwargs`` are trace_raw's keyword arguments
    (eps, check_apertures, intersect_obj, filter_out_phantoms, first_surf,
    last_surf, pt_inside_fuzz) plus ``wvl`` / ``wvl_index`` for a
    single-wavelength bundle.  Asynchronous on the current CUDA stream."""
    lib = _abi.load_library()
    device = torch.device('cuda', table.device)
    p = _as_soa(p, device, 'p')
    d = _as_soa(d, device, 'd')
    n = p.shape[1]
    if d.shape[1] != n:
        raise ValueError('p and d must hold the same number of rays')
    wi = 0
    if 'wvl' in kwargs:
        wi = table.wvl_index(kwargs.pop('wvl'))
    wi = kwargs.pop('wvl_index', wi)
    opts = _abi.make_opts(wvl_idx=wi, **kwargs)
    if wvl_idx is not None:
        wvl_idx = torch.as_tensor(wvl_idx).to(device=device, dtype=torch.int32).contiguous()
        if wvl_idx.numel() != n:
            raise ValueError('wvl_idx must hold one entry per ray')
    res = BundleResult(n, table.n_ifc, device, tuple(outputs) + (('full',) if full else ()))
    out = res.c_struct()
    _abi.check(lib.rt_trace_bundle(table.handle, n, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]),
                                   _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(wvl_idx),
                                   C.byref(opts), C.byref(out), _stream_ptr(device)))
    res._keep = (p, d, wvl_idx)   # inputs must outlive the asynchronous launch
    return res


def accumulated_steps(start, stop, num):
    """The reference's pupil sampling: ``start += step`` repeated (NOT linspace),
    /root/reference/src/rayoptics/raytr/trace.py:567-604."""
    return _accumulated_steps(float(start), float(stop), int(num)).copy()


@functools.lru_cache(maxsize=64)
def _accumulated_steps(start, stop, num):
    vals = np.empty(num)
    step = np.array((np.float64(stop) - np.float64(start))/(num - 1)) if num > 1 else np.float64(0.)
    x = np.float64(start)
    for i in range(num):
        vals[i] = x
        x = x + step
    return vals


def chunk_rays():
    """rays per chunk = threads per CTA of the loaded library (``rt_chunk_rays``; 256)"""
    return int(_abi.load_library().rt_chunk_rays())



class PupilGridSpec:
    """Host-side description of fields x wavelengths x (nx x ny) pupil rays
    (the arrays behind ``rt_grid_spec``; no CUDA involved).

    ``fields``: list of dicts / objects with ``pt0`` (3), ``aim`` (2), ``vlx,
    vux, vly, vuy`` (and optionally ``pupil_kind``, see ``rt_pupil_kind``: for the
    angular kinds ``pt0`` is the object point, ``aim`` the chief-ray direction
    cosines and ``eprad`` the sine / slope scale); ``pupil_x`` / ``pupil_y``: ``[n_fields, nx]`` / ``[n_fields,
    ny]`` relative pupil coordinates before vignetting (or 1-D, shared by all
    fields); ``wvl_idx``: rows of the table's index table; ``ref_img``:
    ``[n_fields, n_wvls, 2]`` reference image points or None."""

    def __init__(self, fields, wvl_idx, pupil_x, pupil_y, eprad, z_pupil, ref_img=None,
                 apply_vignetting=True, flip_z_dir=1, foc=0.0, paired=False, wave=None,
                 pupil_kind=None):
        nf = len(fields)
        if pupil_kind is None:      # records of OpticalSpecs.grid_fields carry it
            f0 = fields[0]
            pupil_kind = f0.get('pupil_kind', 0) if isinstance(f0, dict) else getattr(f0, 'pupil_kind', 0)
        self.pupil_kind = int(pupil_kind)
        self.paired = int(bool(paired))
        self.n_fields = nf
        self.wvl_idx = np.ascontiguousarray(wvl_idx, dtype=np.int32)
        self.n_wvls = len(self.wvl_idx)
        px = np.ascontiguousarray(pupil_x, dtype=np.float64)
        py = np.ascontiguousarray(pupil_y, dtype=np.float64)
        if px.ndim == 1:
            px = np.ascontiguousarray(np.broadcast_to(px, (nf, px.shape[0])))
        if py.ndim == 1:
            py = np.ascontiguousarray(np.broadcast_to(py, (nf, py.shape[0])))
        self.pupil_x, self.pupil_y = px, py
        self.nx, self.ny = px.shape[1], py.shape[1]
        if self.paired:          # ray list: (pupil_x[i], pupil_y[i]), one "row" of nx rays
            if py.shape[1] != px.shape[1]:
                raise ValueError('paired pupil lists need as many y as x coordinates')
            self.ny = 1
        self.fields = (rt_field_desc*nf)()
        for i, f in enumerate(fields):
            get = (lambda k, f=f: f[k]) if isinstance(f, dict) else (lambda k, f=f: getattr(f, k))
            for c in range(3):
                self.fields[i].pt0[c] = float(get('pt0')[c])
            for c in range(2):
                self.fields[i].aim[c] = float(get('aim')[c])
            for k in ('vlx', 'vux', 'vly', 'vuy'):
                setattr(self.fields[i], k, float(get(k)))
            if self.pupil_kind == _abi.PUPIL_WIDE:
                rot = np.asarray(get('rot'), dtype=float).reshape(9)
                for c in range(9):
                    self.fields[i].rot[c] = float(rot[c])
                self.fields[i].obj2enp = float(get('obj2enp'))
        self.ref_img = None
        if ref_img is not None:
            self.ref_img = np.ascontiguousarray(ref_img, dtype=np.float64).reshape(nf, self.n_wvls, 2)
        self.wave = None
        if wave is not None:
            self.wave = np.ascontiguousarray(wave, dtype=np.float64).reshape(
                nf, self.n_wvls, _abi.RT_WAVE_DOUBLES)
        self.eprad, self.z_pupil, self.foc = float(eprad), float(z_pupil), float(foc)
        # wide-angle fields: no virtual-object flip (trace.py:299-303)
        self.apply_vignetting = int(bool(apply_vignetting))
        self.flip_z_dir = 0 if self.pupil_kind == _abi.PUPIL_WIDE else int(flip_z_dir)
        self.chunk_rays = chunk_rays()
        self.rays_per_tile = self.nx*self.ny
        self.n_tiles = self.n_fields*self.n_wvls
        self.chunks_per_tile = (self.rays_per_tile + self.chunk_rays - 1)//self.chunk_rays
        self.n_chunks = self.n_tiles*self.chunks_per_tile
        self.n_rays = self.n_tiles*self.rays_per_tile

    def host_bytes(self):
        """bytes rt_grid_create copies to the device"""
        return (C.sizeof(rt_field_desc)*self.n_fields + self.wvl_idx.nbytes + self.pupil_x.nbytes
                + self.pupil_y.nbytes + (0 if self.ref_img is None else self.ref_img.nbytes)
                + (0 if self.wave is None else self.wave.nbytes))

    def c_spec(self):
        """The rt_grid_spec (host pointers into arrays owned by this object)."""
        s = rt_grid_spec()
        s.n_fields, s.n_wvls, s.nx, s.ny = self.n_fields, self.n_wvls, self.nx, self.ny
        s.fields = C.cast(self.fields, C.POINTER(rt_field_desc))
        s.wvl_idx = self.wvl_idx.ctypes.data_as(_abi.c_int32_p)
        s.pupil_x = self.pupil_x.ctypes.data_as(_abi.c_double_p)
        s.pupil_y = self.pupil_y.ctypes.data_as(_abi.c_double_p)
        s.ref_img = None if self.ref_img is None else self.ref_img.ctypes.data_as(_abi.c_double_p)
        s.wave = None if self.wave is None else self.wave.ctypes.data_as(_abi.c_double_p)
        s.apply_vignetting, s.flip_z_dir = self.apply_vignetting, self.flip_z_dir
        s.paired = self.paired
        s.pupil_kind = self.pupil_kind
        s.eprad, s.z_pupil, s.foc = self.eprad, self.z_pupil, self.foc
        return s

    def first_ray_of_chunk(self, chunk):
        tile, lc = divmod(chunk, self.chunks_per_tile)
        return tile*self.rays_per_tile + min(lc*self.chunk_rays, self.rays_per_tile)

    def rays_in_chunks(self, chunk_begin, chunk_end):
        return self.first_ray_of_chunk(chunk_end) - self.first_ray_of_chunk(chunk_begin)


class PupilGrid(PupilGridSpec):
    """A PupilGridSpec uploaded to the device (``rt_grid*``)."""

    def __init__(self, *args, device=0, **kwargs):
        super().__init__(*args, **kwargs)
        lib = _abi.load_library()
        self.device = int(device)
        spec = self.c_spec()
        handle = C.c_void_p()
        _abi.check(lib.rt_grid_create(C.byref(spec), self.device, C.byref(handle)))
        self._handle, self._lib = handle, lib
        n_rays, n_chunks, chunk = C.c_int64(), C.c_int64(), C.c_int32()
        _abi.check(lib.rt_grid_dims(handle, C.byref(n_rays), C.byref(n_chunks), C.byref(chunk)))
        assert (n_rays.value, n_chunks.value, chunk.value) == (self.n_rays, self.n_chunks,
                                                              self.chunk_rays)

    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError('PupilGrid was destroyed')
        return self._handle

    def shape_key(self):
        return (self.device, self.n_fields, self.n_wvls, self.nx, self.ny, self.paired,
                self.wave is not None)

    def update(self, *args, **kwargs):
        """Replace the description by another one of the same shape (``rt_grid_update``: one
        asynchronous copy on the current stream, no allocation).  Arguments as the constructor."""
        kwargs.pop('device', None)
        old = self.shape_key()
        PupilGridSpec.__init__(self, *args, **kwargs)
        if self.shape_key() != old:
            raise ValueError('PupilGrid.update: the new description has a different shape')
        spec = self.c_spec()
        with torch.cuda.device(self.device):
            _abi.check(self._lib.rt_grid_update(self.handle, C.byref(spec),
                                                _stream_ptr(torch.device('cuda', self.device))))
        return self

    def upload(self, spec):
        """``rt_grid_update`` from a PupilGridSpec built earlier (same shape)."""
        c = spec.c_spec()
        with torch.cuda.device(self.device):
            _abi.check(self._lib.rt_grid_update(self.handle, C.byref(c),
                                                _stream_ptr(torch.device('cuda', self.device))))
        self.foc, self.apply_vignetting = spec.foc, spec.apply_vignetting
        return self

    def chief_ref(self, table, wvl_idx, out=None):
        """Reference image points = image intercepts of the chief rays at row ``wvl_idx`` of
        the table, computed and stored on the device (``rt_grid_chief_ref``); ``out``:
        optional ``[n_fields, 2]`` float64 device tensor that receives a copy."""
        with torch.cuda.device(self.device):
            _abi.check(self._lib.rt_grid_chief_ref(table.handle, self.handle, int(wvl_idx), _ptr(out),
                                                   _stream_ptr(torch.device('cuda', self.device))))
        return out

    def close(self):
        if getattr(self, '_handle', None) is not None:
            self._lib.rt_grid_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def trace_grid(table, grid, chunk_begin=0, chunk_end=None, outputs=GRID_OUTPUTS, full=False,
               summary=True, res=None, nan_status=False, **kwargs):
    """Trace chunks ``[chunk_begin, chunk_end)`` of a PupilGrid.

    Returns a BundleResult whose per-ray tensors (``outputs``; pass ``()`` for
    summary-only, ``res=`` to reuse buffers) cover the rays of those chunks in flattened (field, wvl, i, j) order and whose
    ``summary`` is the ``[n_tiles, 16]`` partial per-(field, wvl) spot sums
    (SUMMARY_FIELDS).  Defaults follow the grid analyses of the reference:
    ``check_apertures=True`` (trace.py:583), ``first_surf=1``,
    ``last_surf=n_ifc-2`` (raytrace.py:77-79)."""
    lib = _abi.load_library()
    device = torch.device('cuda', table.device)
    if chunk_end is None:
        chunk_end = grid.n_chunks
    kwargs.setdefault('check_apertures', True)
    kwargs.setdefault('first_surf', 1)
    kwargs.setdefault('last_surf', table.n_ifc - 2)
    if grid.pupil_kind == _abi.PUPIL_WIDE:
        kwargs['intersect_obj'] = False          # trace_base, trace.py:299-300
    opts = _abi.make_opts(**kwargs)
    n = grid.rays_in_chunks(chunk_begin, chunk_end)
    if res is None:
        res = BundleResult(n, table.n_ifc, device, tuple(outputs) + (('full',) if full else ()),
                           nan_status=nan_status)
    elif res.n != n:
        raise ValueError('res was allocated for a different number of rays')
    out = res.c_struct()
    summ = scratch = None
    if summary:
        summ = torch.empty((grid.n_tiles, RT_SUMMARY_DOUBLES), dtype=torch.float64, device=device)
        nbytes = lib.rt_grid_scratch_bytes(grid.handle, chunk_begin, chunk_end)
        scratch = torch.empty(max(nbytes//8, 1), dtype=torch.float64, device=device)
    _abi.check(lib.rt_trace_grid(table.handle, grid.handle, chunk_begin, chunk_end,
                                 C.byref(opts), C.byref(out), _ptr(summ), _ptr(scratch),
                                 _stream_ptr(device)))
    res.summary = summ
    res._keep = scratch
    return res


def trace_grid_to_host(table, grid, h_abr, chunk_begin=0, chunk_end=None, pieces=8, summary=True,
                       workspace=None, **kwargs):
    """Trace chunks ``[chunk_begin, chunk_end)`` of a PupilGrid and deliver the transverse
    aberrations to page-locked HOST memory (``rt_trace_grid_to_host``): ``pieces`` launches on
    two library-owned streams, each followed by its device->host copy.  ``h_abr``: pinned
    ``[2, >= n]`` float64 tensor; rays that do not reach the image hold NaNs coding status /
    failing surface (``decode_nan_status``).  Returns ``(summary [n_tiles, 16] device tensor or
    None, workspace)``; pass ``workspace`` back in to re-use the device staging buffers.
    Asynchronous: synchronise the current stream before reading ``h_abr``."""
    lib = _abi.load_library()
    device = torch.device('cuda', table.device)
    if chunk_end is None:
        chunk_end = grid.n_chunks
    kwargs.setdefault('check_apertures', True)
    kwargs.setdefault('first_surf', 1)
    kwargs.setdefault('last_surf', table.n_ifc - 2)
    if grid.pupil_kind == _abi.PUPIL_WIDE:
        kwargs['intersect_obj'] = False
    opts = _abi.make_opts(**kwargs)
    n = grid.rays_in_chunks(chunk_begin, chunk_end)
    if not (h_abr.is_pinned() and h_abr.dtype == torch.float64 and h_abr.shape[0] == 2
            and h_abr.shape[1] >= n and h_abr.stride(1) == 1):
        raise ValueError('h_abr must be a pinned float64 tensor [2, >= n] with contiguous rows')
    pieces = max(1, min(int(pieces), chunk_end - chunk_begin))
    nbytes = lib.rt_trace_grid_to_host_scratch_bytes(grid.handle, pieces)
    ws = workspace
    if ws is None or ws['abr'].shape[1] < n or ws['scratch'].numel()*8 < nbytes or ws['device'] != device:
        ws = {'abr': torch.empty((2, max(n, 1)), dtype=torch.float64, device=device),
              'scratch': torch.empty(max(nbytes//8, 1), dtype=torch.float64, device=device),
              'device': device}
    summ = (torch.empty((grid.n_tiles, RT_SUMMARY_DOUBLES), dtype=torch.float64, device=device)
            if summary else None)
    _abi.check(lib.rt_trace_grid_to_host(table.handle, grid.handle, chunk_begin, chunk_end, C.byref(opts),
                                         _ptr(ws['abr'][0]), _ptr(ws['abr'][1]), _ptr(h_abr[0]),
                                         _ptr(h_abr[1]), _ptr(summ), _ptr(ws['scratch']), pieces,
                                         _stream_ptr(device)))
    return summ, ws


def decode_nan_status(abr):
    """``(status, fail_surf)`` int32 arrays from an ``abr`` ``[2, n]`` array written with
    ``nan_status=True`` (numpy, host): rays that reach the image have finite aberrations and
    status 0 / fail_surf -1; the others carry both numbers in the NaN payloads."""
    bits = np.ascontiguousarray(abr).view(np.uint64)
    bad = np.isnan(abr[0])
    status = np.where(bad, bits[0] & np.uint64(0xFFFF), 0).astype(np.int32)
    fs = (bits[1] & np.uint64(0xFFFF)).astype(np.int32)
    fail_surf = np.where(bad, np.where(fs >= 0x8000, fs - 0x10000, fs), -1).astype(np.int32)
    return status, fail_surf


def combine_summaries(parts, out=None):
    """Combine partial ``[n_tiles, 16]`` summaries (from chunk ranges / ranks):
    sums add, min/max columns take min/max.  CUDA tensors: one ``rt_combine_summaries``
    launch on the current stream; CPU tensors (gloo tests): torch."""
    parts = torch.stack(list(parts)) if not torch.is_tensor(parts) else parts
    if parts.is_cuda:
        parts = parts.contiguous()
        if out is None:
            out = torch.empty(parts.shape[1:], dtype=parts.dtype, device=parts.device)
        with torch.cuda.device(parts.device):
            _abi.check(_abi.load_library().rt_combine_summaries(
                _ptr(parts), parts.shape[0], parts.shape[1], _ptr(out), _stream_ptr(parts.device)))
        out._keep = parts
        return out
    out = parts.sum(dim=0)
    for k in (10, 12):
        out[:, k] = parts[:, :, k].min(dim=0).values
    for k in (11, 13):
        out[:, k] = parts[:, :, k].max(dim=0).values
    return out


def spot_statistics(summary):
    """Per-(field, wvl) spot centroid and RMS radius from a combined summary
    (torch tensor or numpy array ``[n_tiles, 16]``; same type out)."""
    s = summary
    if torch.is_tensor(s):
        n = s[:, 0].clamp(min=1.0)
        sqrt0 = lambda v: v.clamp(min=0.0).sqrt()      # noqa: E731
    else:
        n = np.maximum(s[:, 0], 1.0)
        sqrt0 = lambda v: np.sqrt(np.maximum(v, 0.0))  # noqa: E731
    cx, cy = s[:, 5]/n, s[:, 6]/n
    var = (s[:, 7] + s[:, 8])/n - (cx*cx + cy*cy)
    return {'n_ok': s[:, 0], 'n_missed': s[:, 1], 'n_tir': s[:, 2], 'n_blocked': s[:, 3],
            'centroid_x': cx, 'centroid_y': cy, 'rms_radius': sqrt0(var),
            'min_x': s[:, 10], 'max_x': s[:, 11], 'min_y': s[:, 12], 'max_y': s[:, 13],
            'mean_op': s[:, 14]/n}


def measure_fp64_peak(device=0):
    """TFLOP/s of the fp64 vector pipe (DFMA microbenchmark in the library)."""
    lib = _abi.load_library()
    v = C.c_double()
    _abi.check(lib.rt_measure_fp64_peak(int(device), C.byref(v)))
    return v.value


def measure_fp64_latency(device=0):
    """cycles between two dependent DFMAs of one warp (``rt_measure_fp64_latency``)"""
    lib = _abi.load_library()
    v = C.c_double()
    _abi.check(lib.rt_measure_fp64_latency(int(device), C.byref(v)))
    return v.value


def launch_count():
    return int(_abi.load_library().rt_launch_count())


def _grid_args(opt_model, wvl_index, num_rays, fields, wvls, foc, pupil_range, apply_vignetting):
    osp, sm = opt_model.optical_spec, opt_model.seq_model
    fields = list(osp.field_of_view.fields if fields is None else fields)
    wvls = list(sm.wvlns if wvls is None else wvls)
    recs, eprad, z_pupil = osp.grid_fields(fields)
    foc = osp.defocus.focus_shift if foc is None else foc
    xs = accumulated_steps(pupil_range[0], pupil_range[1], num_rays)
    args = (recs, [wvl_index(w) for w in wvls], xs, xs, eprad, z_pupil)
    kw = dict(apply_vignetting=apply_vignetting, flip_z_dir=sm.z_dir[0], foc=foc)
    return args, kw


def grid_spec_for_model(opt_model, num_rays, fields=None, wvls=None, foc=None,
                        pupil_range=(-1.0, 1.0), apply_vignetting=True, ref_img=None):
    """Host-only PupilGridSpec of the reference's square-grid analyses (no CUDA)."""
    sm = opt_model.seq_model
    args, kw = _grid_args(opt_model, sm.index_for_wavelength, num_rays, fields, wvls, foc,
                          pupil_range, apply_vignetting)
    return PupilGridSpec(*args, ref_img=ref_img, **kw)


def grid_for_model(opt_model, table, num_rays, fields=None, wvls=None, foc=None,
                   pupil_range=(-1.0, 1.0), apply_vignetting=True, ref_img='chief'):
    """PupilGrid for the reference's square-grid analyses of a model:
    fields x wavelengths x (num_rays x num_rays) over relative pupil
    ``[-1, 1]^2`` with the reference's accumulated stepping
    (raytr/trace.py:563-605; seq/sequential.py:1058-1085).

    ``ref_img='chief'`` traces the (0, 0) pupil ray of every (field, wvl) first
    (one tiny grid launch) and uses its image intercept as the reference image
    point, as ``calculate_reference_sphere`` does (raytr/waveabr.py:24-76)."""
    args, kw = _grid_args(opt_model, table.wvl_index, num_rays, fields, wvls, foc, pupil_range,
                          apply_vignetting)
    ref = None
    if isinstance(ref_img, str) and ref_img == 'chief':
        g0 = PupilGrid(args[0], args[1], [0.0], [0.0], args[4], args[5], device=table.device,
                       **dict(kw, apply_vignetting=False))
        r0 = trace_grid(table, g0, outputs=('p',), summary=False, check_apertures=False)
        ref = r0.p[:2].t().contiguous().cpu().numpy().reshape(len(args[0]), len(args[1]), 2)
        g0.close()
    elif ref_img is not None:
        ref = ref_img
    return PupilGrid(*args, ref_img=ref, device=table.device, **kw)
