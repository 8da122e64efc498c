"""ctypes wrapper of oracle/librt_oracle.so (the CPU parity oracle).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference arm.  Never imported by rayoptics_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rayoptics_b200 import _abi
from rayoptics_b200._abi import rt_surface_desc, rt_opts, rt_grid_spec, RT_SEG_DOUBLES

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(['make', '-s', '-C', _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'librt_oracle.so')
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        assert L.rto_sizeof_surface_desc() == C.sizeof(rt_surface_desc)
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def trace_ray(descs, n_row, pt0, dir0, opts, wvl=float('nan')):
    """One ray -> dict(ray [n_seg,10], op, status, fail_surf)."""
    n_ifc = len(descs)
    n_row = np.ascontiguousarray(n_row, dtype=np.float64)
    ray = np.zeros((n_ifc, RT_SEG_DOUBLES))
    last = np.zeros(RT_SEG_DOUBLES)
    p0 = np.ascontiguousarray(pt0, dtype=np.float64)
    d0 = np.ascontiguousarray(dir0, dtype=np.float64)
    n_seg, st, fs = C.c_int32(), C.c_int32(), C.c_int32()
    op = C.c_double()
    lib().rto_trace_ray(descs, C.c_int32(n_ifc), _dp(n_row), C.c_double(wvl), _dp(p0), _dp(d0), C.byref(opts),
                        _dp(ray), _dp(last), C.byref(n_seg), C.byref(op), C.byref(st), C.byref(fs))
    return {'ray': ray[:n_seg.value].copy(), 'last': last, 'op': op.value,
            'status': st.value, 'fail_surf': fs.value, 'n_seg': n_seg.value}


def trace_bundle(descs, n_by_wvl, p, d, wvl_idx, opts, want_full=False, n_threads=1, wvls=None):
    """p, d: [3, n] arrays.  Returns dict of SoA numpy arrays like rt_out."""
    n_ifc = len(descs)
    n = p.shape[1]
    n_by_wvl = np.ascontiguousarray(n_by_wvl, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    wv = None if wvl_idx is None else np.ascontiguousarray(wvl_idx, dtype=np.int32)
    wl = None if wvls is None else np.ascontiguousarray(wvls, dtype=np.float64)
    last = np.zeros((RT_SEG_DOUBLES, n))
    full = np.full((n_ifc, RT_SEG_DOUBLES, n), np.nan) if want_full else None
    op = np.zeros(n)
    status = np.zeros(n, dtype=np.int32)
    fail_surf = np.zeros(n, dtype=np.int32)
    n_seg = np.zeros(n, dtype=np.int32)
    lib().rto_trace_bundle(descs, C.c_int32(n_ifc), _dp(n_by_wvl), C.c_int64(n),
                           _dp(p[0]), _dp(p[1]), _dp(p[2]), _dp(d[0]), _dp(d[1]), _dp(d[2]),
                           _ip(wv), C.byref(opts), _dp(wl), _dp(last), _dp(full), C.c_int64(n),
                           _dp(op), _ip(status), _ip(fail_surf), _ip(n_seg), C.c_int32(n_threads))
    return {'last': last, 'full': full, 'op': op, 'status': status,
            'fail_surf': fail_surf, 'n_seg': n_seg}


def grid_start_rays(spec: rt_grid_spec, ray_begin, ray_end):
    """Start rays of a grid spec (host pointers) -> p[3,n], d[3,n], wvl_idx[n], pupil[2,n]."""
    n = ray_end - ray_begin
    p = np.zeros((3, n))
    d = np.zeros((3, n))
    wv = np.zeros(n, dtype=np.int32)
    pup = np.zeros((2, n))
    lib().rto_grid_start_rays(C.byref(spec), C.c_int64(ray_begin), C.c_int64(ray_end),
                              _dp(p[0]), _dp(p[1]), _dp(p[2]), _dp(d[0]), _dp(d[1]), _dp(d[2]),
                              _ip(wv), _dp(pup[0]), _dp(pup[1]))
    return p, d, wv, pup


def transverse_abr(px, py, dx, dy, dz, foc, ref_x, ref_y):
    n = len(px)
    ax, ay = np.zeros(n), np.zeros(n)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (px, py, dx, dy, dz)]
    lib().rto_transverse_abr(C.c_int64(n), *[_dp(a) for a in arrs], C.c_double(foc),
                             C.c_double(ref_x), C.c_double(ref_y), _dp(ax), _dp(ay))
    return ax, ay


def trace_grid(spec: rt_grid_spec, descs, n_by_wvl, ray_begin, ray_end, opts, n_threads=1,
               want_last=True, wvls=None):
    """Whole grid on the host (start rays + trace + transverse aberration)."""
    n = ray_end - ray_begin
    n_by_wvl = np.ascontiguousarray(n_by_wvl, dtype=np.float64)
    last = np.zeros((RT_SEG_DOUBLES, n)) if want_last else None
    op = np.zeros(n)
    status = np.zeros(n, dtype=np.int32)
    fail_surf = np.zeros(n, dtype=np.int32)
    ax, ay = np.zeros(n), np.zeros(n)
    opd = np.full(n, np.nan) if bool(spec.wave) else None
    wl = None if wvls is None else np.ascontiguousarray(wvls, dtype=np.float64)
    lib().rto_trace_grid(C.byref(spec), descs, C.c_int32(len(descs)), _dp(n_by_wvl),
                         C.c_int64(ray_begin), C.c_int64(ray_end), C.byref(opts),
                         _dp(last), _dp(op), _ip(status), _ip(fail_surf), _dp(ax), _dp(ay),
                         _dp(opd), _dp(wl), C.c_int32(n_threads))
    return {'last': last, 'op': op, 'status': status, 'fail_surf': fail_surf,
            'abr': np.stack([ax, ay]), 'opd': opd}


def wave_opd(W, p1, d0, pk, dk, ray_op, pl=None, dl=None):
    """wave_abr_full_calc for one ray (W: RT_WAVE_DOUBLES record; pl, dl: ray[-1],
    needed by the infinite-reference variant only)."""
    f = lib().rto_wave_opd
    f.restype = C.c_double
    if pl is None:
        pl, dl = np.zeros(3), np.zeros(3)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (W, p1, d0, pk, dk, pl, dl)]
    return f(*[_dp(a) for a in arrs], C.c_double(ray_op))


class GridRunner:
    """bench.py's CPU arm: the whole-grid trace on a persistent pthread pool with dynamic
    scheduling and output buffers that are allocated (and first touched) once, so that a timed
    call is nothing but ``rto_trace_grid_pool``."""

    def __init__(self, spec, descs, n_by_wvl, opts, n_rays_max, n_threads, block=2048):
        self.spec, self.descs, self.opts = spec, descs, opts
        self.c_spec = spec.c_spec() if hasattr(spec, 'c_spec') else spec
        self.n_by_wvl = np.ascontiguousarray(n_by_wvl, dtype=np.float64)
        self.n_threads, self.block = int(n_threads), int(block)
        n = int(n_rays_max)
        self.op = np.ones(n)
        self.status = np.ones(n, dtype=np.int32)
        self.fail_surf = np.ones(n, dtype=np.int32)
        self.ax, self.ay = np.ones(n), np.ones(n)
        lib().rto_pool_create(C.c_int32(self.n_threads))

    def run(self, ray_begin, ray_end):
        """trace rays [ray_begin, ray_end) into the buffers; returns seconds inside the C call"""
        import time
        n = ray_end - ray_begin
        assert n <= self.op.shape[0]
        L = lib()
        t0 = time.perf_counter()
        rc = L.rto_trace_grid_pool(C.byref(self.c_spec), self.descs, C.c_int32(len(self.descs)),
                                   _dp(self.n_by_wvl), C.c_int64(ray_begin), C.c_int64(ray_end),
                                   C.byref(self.opts), None, _dp(self.op), _ip(self.status),
                                   _ip(self.fail_surf), _dp(self.ax), _dp(self.ay), None, None,
                                   C.c_int64(self.block))
        dt = time.perf_counter() - t0
        assert rc == 0
        return dt

    def close(self):
        lib().rto_pool_destroy()
