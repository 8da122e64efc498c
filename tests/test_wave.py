"""Wavefront (OPD) stage: the reference's raytr/waveabr.py:24-305.

CPU: the C restatement (oracle/rt_oracle.c rto_wave_opd) and the host-side
chief-ray / reference-sphere records (rayoptics_b200/waveabr.py) reproduce the
numbers the REFERENCE's own waveabr functions produced (tests/golden/vectors/
<model>_opd.npz, generator tests/golden/make_golden_opd.py) bit for bit.
GPU: the kernel epilogue against the oracle and the golden OPDs.  The GPU
evaluates F**2 as F*F where numpy's scalar power calls libm pow(): 1-ulp
differences in ~0.1 % of rays; tolerance 1e-12 mm (north_star: 1e-10 mm).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_model
from rayoptics_b200 import _abi, table as T, engine as E, waveabr as W

OPD_MODELS = ['dblgauss', 'rc', 'cellphone', 'triplet']
# 'telecentric': exit pupil at ~4e14 mm -> the axial tiles take the infinite-reference branch
# (raytr/waveabr.py:356-420).  Its GPU run lives in tests/test_zz_gpu_additions.py.
OPD_MODELS_CPU = OPD_MODELS + ['telecentric']


def load_opd(name):
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_opd.npz'))
    return {k: z[k] for k in z.files}


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('name', OPD_MODELS_CPU)
def test_oracle_opd_matches_reference(oracle, name):
    opm = load_model(name)
    v = load_opd(name)
    descs, n_by_wvl, _ = T.describe_model(opm.seq_model)
    n_ifc = len(descs)
    opts = _abi.make_opts(first_surf=1, last_surf=n_ifc - 2, check_apertures=True)
    r = oracle.trace_bundle(descs, n_by_wvl, v['p0'], v['d0'], v['wvl_idx'], opts, want_full=True)
    assert same(r['status'], v['status'])
    ok = np.nonzero(v['status'] == 0)[0]
    got = np.full(v['opd'].shape, np.nan)
    for k in ok:
        full = r['full'][:, :, k]
        got[k] = oracle.wave_opd(v['wave'][v['tile'][k]], full[1, 0:3], full[0, 3:6],
                                 full[n_ifc - 2, 0:3], full[n_ifc - 2, 3:6], r['op'][k],
                                 pl=full[n_ifc - 1, 0:3], dl=full[n_ifc - 1, 3:6])
    assert same(got, v['opd'])
    if name == 'telecentric':
        assert (v['wave'][:, 21] == 0).sum() == 3        # the axial tiles are infinite-reference


@pytest.mark.parametrize('name', OPD_MODELS_CPU)
def test_host_wave_records_match_reference(oracle, name):
    """chief ray -> exit pupil segment -> reference sphere -> 24-double record,
    computed by rayoptics_b200/waveabr.py from oracle-traced chief rays."""
    opm = load_model(name)
    osp, sm = opm.optical_spec, opm.seq_model
    v = load_opd(name)
    descs, n_by_wvl, _ = T.describe_model(sm)
    n_ifc = len(descs)
    opts = _abi.make_opts(first_surf=1, last_surf=n_ifc - 2)
    t = 0
    for fld in osp.fov.fields:
        for wi, wvl in enumerate(sm.wvlns):
            pt0, dir0 = osp.ray_start_from_osp(fld.apply_vignetting([0., 0.]), fld, 'rel pupil')
            r = oracle.trace_ray(descs, n_by_wvl[wi], pt0, dir0, opts)
            full = r['ray'].reshape(n_ifc, 10, 1)
            crp = W.chief_ray_pkg(opm, full, np.array([r['op']]), wvl, 0)
            rs = W.calculate_reference_sphere(opm, fld, wvl, 0.0, crp)
            assert same(W.wave_record(opm, crp, rs), v['wave'][t])
            t += 1
    assert t == v['wave'].shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize('name', OPD_MODELS)
def test_cuda_opd_matches_oracle_and_reference(oracle, name):
    import torch
    opm = load_model(name)
    osp, sm = opm.optical_spec, opm.seq_model
    tab = T.SurfaceTable.from_model(sm, device=0)
    fields, wvls = list(osp.fov.fields), list(sm.wvlns)
    v = load_opd(name)
    num = int(v['num'])
    wave, ref_img, _ = W.setup_tiles(opm, tab, fields, wvls, 0.0)
    # the GPU-traced chief rays give the reference's records bit for bit
    assert same(wave.reshape(-1, _abi.RT_WAVE_DOUBLES), v['wave'])
    recs, eprad, z_pupil = osp.grid_fields(fields)
    xs = E.accumulated_steps(-1.0, 1.0, num)
    grid = E.PupilGrid(recs, [tab.wvl_index(w) for w in wvls], xs, xs, eprad, z_pupil,
                       ref_img=ref_img, flip_z_dir=sm.z_dir[0], wave=wave, device=0)
    res = E.trace_grid(tab, grid, outputs=('abr', 'status', 'opd'))
    torch.cuda.synchronize()
    opd = res.opd.cpu().numpy()
    assert same(res.status.cpu().numpy(), v['status'])
    ok = v['status'] == 0
    assert same(res.abr.cpu().numpy()[:, ok], v['abr'][:, ok])
    assert np.isnan(opd[~ok]).all()
    np.testing.assert_allclose(opd[ok], v['opd'][ok], rtol=0, atol=1e-12)
    assert (opd[ok] == v['opd'][ok]).mean() > 0.95
    # oracle on the same grid spec (uses pow like the reference)
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    ref = oracle.trace_grid(grid.c_spec(), tab.descs, tab.n_by_wvl, 0, grid.n_rays, opts)
    assert same(ref['opd'], v['opd'])


@pytest.mark.gpu
def test_analysis_classes(oracle):
    """RayFan / RayList / RayGrid (raytr/analyses.py:121-187,343-434,584-663):
    constructor surface and result shapes, values against the golden OPDs."""
    from rayoptics_b200 import analyses as A
    opm = load_model('dblgauss')
    osp = opm.optical_spec
    v = load_opd('dblgauss')
    num = int(v['num'])
    fi, wvl = 1, opm.seq_model.wvlns[1]
    tile = fi*3 + 1
    m = v['tile'] == tile
    # RayGrid over [-1,1]^2 (oversize such that bbox == unit square is not the default:
    # use a field without vignetting for an exact comparison)
    fan = A.RayFan(opm, f=fi, wl=wvl, num_rays=num, xyfan='y')
    col = (np.arange(num*num)//num == num//2)          # x index = middle column -> pupil x = 0
    gold_opd = v['opd'][m][col]
    gold_abr = v['abr'][:, m][:, col]
    gold_ok = v['status'][m][col] == 0
    # the fan is traced without aperture clipping; rays that are inside the apertures agree
    assert len(fan.fan) >= gold_ok.sum()
    fan_by_y = {round(p[1], 12): val for p, val in fan.fan}
    xs = E.accumulated_steps(-1.0, 1.0, num)
    fld = osp.fov.fields[fi]
    conv = 1/opm.nm_to_sys_units(wvl)
    for j in range(num):
        if gold_ok[j]:
            py = fld.apply_vignetting(np.array([0.0, xs[j]]))[1]
            dx, dy, opd = fan_by_y[round(py, 12)]
            assert dx == gold_abr[0, j] and dy == gold_abr[1, j]
            assert abs(opd - conv*gold_opd[j]) <= 1e-12*conv
    lst = A.RayList(opm, num_rays=11, f=0, wl=wvl)
    assert lst.ray_abr.shape[0] == 2 and 60 < lst.ray_abr.shape[1] <= 121
    grid = A.RayGrid(opm, f=0, wl=wvl, num_rays=num)
    assert grid.grid.shape == (3, num, num)
    m0 = v['tile'] == 1                                  # field 0, wvl index 1: no vignetting
    gold = np.where(v['status'][m0] == 0, conv*v['opd'][m0], np.nan).reshape(num, num)
    np.testing.assert_allclose(grid.grid[2], gold, rtol=0, atol=1e-12*conv, equal_nan=True)
    assert same(grid.grid[0][:, 0], xs) and same(grid.grid[1][0, :], xs)


@pytest.mark.gpu
def test_psf_matches_numpy_restatement():
    """calc_psf (raytr/analyses.py:848-875) on torch.fft vs the same steps in numpy."""
    from numpy.fft import fftshift, fft2
    from rayoptics_b200 import analyses as A
    opm = load_model('dblgauss')
    ndim, maxdim = 32, 128
    grid = A.RayGrid(opm, f=1, wl=opm.seq_model.wvlns[1], num_rays=ndim)
    AP = A.calc_psf(grid.grid[2], ndim, maxdim)
    W = np.zeros([maxdim, maxdim])
    nd2, m2 = ndim//2, maxdim//2
    W[m2 - (nd2 - 1):m2 + (nd2 + 1), m2 - (nd2 - 1):m2 + (nd2 + 1)] = np.nan_to_num(grid.grid[2])
    phase = np.exp(1j*2*np.pi*W)
    phase[phase == 1] = 0
    ref = abs(fftshift(fft2(fftshift(phase))))**2
    ref = ref/np.nanmax(ref)
    assert AP.shape == (maxdim, maxdim) and AP.max() == 1.0
    np.testing.assert_allclose(AP, ref, rtol=0, atol=1e-12)
    assert A.psf_sampling(n=128, n_pupil=32) == (128, 32, round(2.44*128/32))
