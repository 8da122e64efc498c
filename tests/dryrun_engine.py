"""Dry-run engine for the GPU test files (TEST INFRASTRUCTURE, opt-in, never used by the driver).

    B200RT_DRYRUN=1 python -m pytest tests/test_zz_gpu_additions.py -m gpu -q

replaces the CUDA entry points of rayoptics_b200.engine / table with stand-ins that compute the
same results with the oracle on CPU tensors, so that GPU test CODE written without access to a
GPU (API use, shapes, tolerances, fixtures) can be exercised before it first meets the hardware.
It says nothing about the kernels -- that is what the real `-m gpu` run is for.
"""
import numpy as np
import torch

from oracle import rt_oracle
from rayoptics_b200 import _abi, engine as E, table as T, analyses as A


class DryTable:
    def __init__(self, descs, n_by_wvl, wvls=None, device=0):
        self.descs, self.n_by_wvl = descs, np.ascontiguousarray(n_by_wvl, dtype=np.float64)
        self.n_ifc, self.n_wvl = len(descs), self.n_by_wvl.shape[0]
        self.wvls = list(wvls) if wvls is not None else list(range(self.n_wvl))
        self.device, self.handle = int(device), object()

    def wvl_index(self, wvl):
        return self.wvls.index(wvl)

    def close(self):
        pass


def _from_model(cls, seq_model, device=0, wvls=None):
    descs, n_by_wvl, wv = T.describe_model(seq_model, wvls)
    return DryTable(descs, n_by_wvl, wv, device)


def _from_path(cls, path, device=0, wvl=None):
    descs, ns = T.describe_path(path)
    return DryTable(descs, np.array([ns]), None if wvl is None else [float(wvl)], device)


class DryGrid(E.PupilGridSpec):
    def __init__(self, *args, device=0, **kwargs):
        super().__init__(*args, **kwargs)
        self.device, self.handle = int(device), object()

    def host_bytes(self):
        return 0

    def close(self):
        pass


def _result(n, n_ifc, outputs, full):
    return E.BundleResult(n, n_ifc, torch.device('cpu'), tuple(outputs) + (('full',) if full else ()))


def _fill(res, r, keys):
    t = torch.from_numpy
    if res.p is not None:
        res.p.copy_(t(r['last'][0:3].copy()))
    if res.d is not None:
        res.d.copy_(t(r['last'][3:6].copy()))
    if res.dst is not None:
        res.dst.copy_(t(r['last'][6].copy()))
    if res.nrml is not None:
        res.nrml.copy_(t(r['last'][7:10].copy()))
    for k in ('op', 'status', 'fail_surf', 'n_seg'):
        if getattr(res, k) is not None and k in r:
            getattr(res, k).copy_(t(np.ascontiguousarray(r[k])))
    if res.full is not None:
        res.full.copy_(t(r['full']))


def trace_bundle(table, p, d, wvl_idx=None, full=False, outputs=E.BUNDLE_OUTPUTS, **kwargs):
    p = np.ascontiguousarray(torch.as_tensor(p).cpu().numpy(), dtype=np.float64).reshape(3, -1)
    d = np.ascontiguousarray(torch.as_tensor(d).cpu().numpy(), dtype=np.float64).reshape(3, -1)
    n = p.shape[1]
    wi = table.wvl_index(kwargs.pop('wvl')) if 'wvl' in kwargs else 0
    wi = kwargs.pop('wvl_index', wi)
    wv = np.full(n, wi, np.int32) if wvl_idx is None else np.asarray(torch.as_tensor(wvl_idx).cpu().numpy(), np.int32)
    r = rt_oracle.trace_bundle(table.descs, table.n_by_wvl, p, d, wv, _abi.make_opts(**kwargs),
                               want_full=True, n_threads=4, wvls=table.wvls)
    res = _result(n, table.n_ifc, outputs, full)
    _fill(res, r, outputs)
    return res


def trace_grid(table, grid, chunk_begin=0, chunk_end=None, outputs=E.GRID_OUTPUTS, full=False,
               summary=True, res=None, nan_status=False, **kwargs):
    chunk_end = grid.n_chunks if chunk_end is None else chunk_end
    kwargs.setdefault('check_apertures', True)
    kwargs.setdefault('first_surf', 1)
    kwargs.setdefault('last_surf', table.n_ifc - 2)
    opts = _abi.make_opts(**kwargs)
    r0, r1 = grid.first_ray_of_chunk(chunk_begin), grid.first_ray_of_chunk(chunk_end)
    spec = grid.c_spec()
    g = rt_oracle.trace_grid(spec, table.descs, table.n_by_wvl, r0, r1, opts, n_threads=4, wvls=table.wvls)
    if full:
        p, d, wv, _ = rt_oracle.grid_start_rays(spec, r0, r1)
        b = rt_oracle.trace_bundle(table.descs, table.n_by_wvl, p, d, wv, opts, want_full=True,
                                   n_threads=4, wvls=table.wvls)
        g['full'], g['n_seg'] = b['full'], b['n_seg']
    if res is None:
        res = _result(r1 - r0, table.n_ifc, outputs, full)
    _fill(res, g, outputs)
    if res.abr is not None:
        abr = g['abr'].copy()
        if nan_status:
            bad = g['status'] != 0
            bits = abr.view(np.uint64)
            bits[0, bad] = _abi.RT_NAN_PAYLOAD_BASE | g['status'][bad].astype(np.uint64)
            bits[1, bad] = _abi.RT_NAN_PAYLOAD_BASE | g['fail_surf'][bad].astype(np.uint64)
        res.abr.copy_(torch.from_numpy(abr))
    if res.opd is not None and g['opd'] is not None:
        res.opd.copy_(torch.from_numpy(g['opd']))
    if summary:
        summ = np.zeros((grid.n_tiles, _abi.RT_SUMMARY_DOUBLES))
        tile = (np.arange(r0, r1)//grid.rays_per_tile)
        for s in range(4):
            np.add.at(summ[:, s], tile, g['status'] == s)
        ok = g['status'] == 0
        np.add.at(summ[:, 5], tile[ok], g['abr'][0, ok])
        np.add.at(summ[:, 6], tile[ok], g['abr'][1, ok])
        res.summary = torch.from_numpy(summ)
    return res


def calc_psf(wavefront, ndim, maxdim, device=0):
    from numpy.fft import fftshift, fft2
    W = np.zeros([maxdim, maxdim])
    m2, nd2 = maxdim//2, ndim//2
    W[m2 - (nd2 - 1):m2 + (nd2 + 1), m2 - (nd2 - 1):m2 + (nd2 + 1)] = np.nan_to_num(wavefront)
    phase = np.exp(1j*2*np.pi*W)
    phase[phase == 1] = 0
    AP = abs(fftshift(fft2(fftshift(phase))))**2
    return AP/np.nanmax(AP)


def install():
    T.SurfaceTable.from_model = classmethod(_from_model)
    T.SurfaceTable.from_path = classmethod(_from_path)
    E.PupilGrid = DryGrid
    E.trace_bundle, E.trace_grid = trace_bundle, trace_grid
    A.calc_psf = calc_psf
    torch.cuda.synchronize = lambda *a, **k: None
