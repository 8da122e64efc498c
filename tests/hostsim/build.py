"""Build + ctypes wrapper of tests/hostsim (TEST INFRASTRUCTURE, see cuda_runtime.h
there): the per-ray device headers compiled for the host."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, '_build', 'libhostsim.so')
SOURCES = [os.path.join(HERE, 'hostsim.cpp'), os.path.join(HERE, 'cuda_runtime.h'),
           os.path.join(ROOT, 'rayoptics_b200', 'csrc', 'rt_device.cuh'),
           os.path.join(ROOT, 'rayoptics_b200', 'csrc', 'rt_lean.cuh'),
           os.path.join(ROOT, 'rayoptics_b200', 'csrc', 'rt_grid.cuh'),
           os.path.join(ROOT, 'include', 'b200rt.h')]

_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    stale = force or not os.path.exists(LIB) or \
        any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SOURCES)
    if stale:
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-mfma', '-DRT_HOSTSIM',
                               '-fPIC', '-shared', '-I', HERE, '-o', LIB, SOURCES[0]])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hostsim_check_division.restype = C.c_int64
        _lib.hostsim_check_sqrt.restype = C.c_int64
        _lib.hostsim_wave_opd.restype = C.c_double
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def lean_kind(descs):
    """0 = general kernel only, 1 = lean, 2 = lean with out-of-line polynomial
    intersect -- the eligibility rule of rt_table_create (csrc/b200rt.cu)."""
    kind = 1
    for s in descs:
        if s.has_tfrm != 0 or s.n_apertures != 0 or s.phase_kind != 0 or s.profile == 6:
            return 0
        if s.profile > 1:
            kind = 2
    return kind


def trace_bundle(descs, n_by_wvl, p0, d0, wvl_idx, opts, kernel=0, out_kind=2, wvls=None):
    n = p0.shape[1]
    n_ifc = len(descs)
    n_by_wvl = np.ascontiguousarray(n_by_wvl, dtype=np.float64)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    d0 = np.ascontiguousarray(d0, dtype=np.float64)
    wv = np.ascontiguousarray(wvl_idx, dtype=np.int32)
    wl = None if wvls is None else np.ascontiguousarray(wvls, dtype=np.float64)
    last = np.zeros((10, n))
    op = np.zeros(n)
    status = np.zeros(n, dtype=np.int32)
    fail_surf = np.zeros(n, dtype=np.int32)
    n_seg = np.zeros(n, dtype=np.int32)
    full = np.full((n_ifc, 10, n), np.nan) if out_kind == 2 else None
    rc = lib().hostsim_trace_bundle(descs, C.c_int(n_ifc), _dp(n_by_wvl), C.c_int(n_by_wvl.shape[0]),
                                    _dp(wl), C.c_int64(n), _dp(p0), _dp(d0), _ip(wv), C.byref(opts),
                                    C.c_int(kernel), C.c_int(out_kind), _dp(last), _dp(op), _ip(status),
                                    _ip(fail_surf), _ip(n_seg), _dp(full))
    assert rc == 0
    return {'last': last, 'op': op, 'status': status, 'fail_surf': fail_surf, 'n_seg': n_seg,
            'full': full}


def check_division(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nf = C.c_int64(0)
    bad = lib().hostsim_check_division(C.c_int64(a.size), _dp(a), _dp(b), C.byref(nf))
    return int(bad), int(nf.value)


def check_sqrt(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    nf = C.c_int64(0)
    bad = lib().hostsim_check_sqrt(C.c_int64(x.size), _dp(x), C.byref(nf))
    return int(bad), int(nf.value)


def wave_opd(W, p1, d0, pk, dk, pl, dl, ray_op):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (W, p1, d0, pk, dk, pl, dl)]
    return lib().hostsim_wave_opd(*[_dp(a) for a in arrs], C.c_double(ray_op))


def grid_start_rays(spec, r0, r1, lean=False):
    """spec: rt_grid_spec (PupilGridSpec.c_spec()) -> p [3, n], d [3, n]"""
    n = r1 - r0
    p, d = np.zeros((3, n)), np.zeros((3, n))
    rc = lib().hostsim_grid_start_rays(C.byref(spec), C.c_int64(r0), C.c_int64(r1), C.c_int(int(lean)),
                                       _dp(p), _dp(d))
    assert rc == 0
    return p, d
