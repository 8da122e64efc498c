"""CUDA parity for diffractive phase elements (DiffractionGrating, radial
DiffractiveElement; oprops/doe.py) -- TOLERANCE parity: the reference evaluates
``x**k`` with libm pow(), the device with exact-rounded products (rt_device.cuh
pow_int_rn), so ~0.1 % of the rays differ in the last bits.  Bar: status, failing
surface and segment counts exact; coordinates within 1e-11 mm (north star 1e-10).

Also here: the batched drivers of rayoptics_b200/trace.py through the CUDA engine.

Sorted last on purpose: this code was added after the last GPU session of the round in
which it was written; the host-compiled device source (tests/test_hostsim.py) and the
oracle-fed `tracer=` seam (tests/test_trace_drivers.py) are what pinned it first.
"""
import numpy as np
import pytest
import torch

from conftest import ANGULAR_MODEL_NAMES, PHASE_MODEL_NAMES, load_model, load_vectors, seeded_bundle
from rayoptics_b200 import _abi, engine as E, table as T, trace as TR, analyses as A
from test_trace_drivers import oracle_tracer
from test_wave import test_cuda_opd_matches_oracle_and_reference as _opd_check

pytestmark = pytest.mark.gpu
TOL_MM = 1e-11


def np_(t):
    return t.detach().cpu().numpy()


def check(r, ref, n_full=None, full_ref=None):
    assert np.array_equal(np_(r.status), ref['status'])
    assert np.array_equal(np_(r.fail_surf), ref['fail_surf'])
    assert np.array_equal(np_(r.n_seg), ref['n_seg'])
    last = np.concatenate([np_(r.p), np_(r.d), np_(r.dst)[None], np_(r.nrml)])
    assert np.array_equal(np.isnan(last), np.isnan(ref['last']))
    d_last = np.nan_to_num(np.abs(last - ref['last']))
    d_op = np.nan_to_num(np.abs(np_(r.op) - ref['op']))
    assert d_last.max(initial=0.0) <= TOL_MM and d_op.max(initial=0.0) <= TOL_MM*1e3
    return float(((d_last.max(0) == 0) & (d_op == 0)).mean())


@pytest.mark.parametrize('name', PHASE_MODEL_NAMES)
def test_cuda_phase_elements_match_reference_vectors(name):
    opm = load_model(name)
    tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    v = load_vectors(name)
    n_full = v['full'].shape[2]
    for ci, case in enumerate(v['cases']):
        idx = np.nonzero(v['case'] == ci)[0]
        if idx.size == 0:
            continue
        r = E.trace_bundle(tab, v['p0'][:, idx], v['d0'][:, idx], wvl_idx=v['wvl_idx'][idx],
                           full=True, **case)
        torch.cuda.synchronize()
        ref = {k: v[k][..., idx] for k in ('last', 'op', 'status', 'n_seg')}
        ref['fail_surf'] = np.where(v['status'][idx] == 0, -1, v['fail_surf'][idx])
        assert check(r, ref) > 0.98
        sel = idx < n_full
        np.testing.assert_allclose(np.nan_to_num(np_(r.full)[:, :, sel]),
                                   np.nan_to_num(v['full'][:, :, idx[sel]]), rtol=0, atol=TOL_MM)


@pytest.mark.parametrize('name', PHASE_MODEL_NAMES)
def test_cuda_phase_elements_match_oracle_bundle(oracle, name):
    opm = load_model(name)
    tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    rng = np.random.default_rng(3)
    p0, d0, wv = seeded_bundle(opm, 20000, rng)
    for ca in (True, False):
        case = dict(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=ca)
        ref = oracle.trace_bundle(tab.descs, tab.n_by_wvl, p0, d0, wv, _abi.make_opts(**case),
                                  n_threads=8, wvls=tab.wvls)
        r = E.trace_bundle(tab, p0, d0, wvl_idx=wv, **case)
        torch.cuda.synchronize()
        assert check(r, ref) > 0.98


def test_cuda_phase_grid(oracle):
    """grid launch (start rays on the device, spot sums) on the hybrid lens"""
    opm = load_model('diffractive')
    tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    grid = E.grid_for_model(opm, tab, 32)
    r = E.trace_grid(tab, grid)
    torch.cuda.synchronize()
    spec = grid.c_spec()
    p, d, wv, _ = oracle.grid_start_rays(spec, 0, grid.n_rays)
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    ref = oracle.trace_bundle(tab.descs, tab.n_by_wvl, p, d, wv, opts, n_threads=8, wvls=tab.wvls)
    assert np.array_equal(np_(r.status), ref['status'])
    ok = ref['status'] == 0
    assert ok.sum() > grid.n_rays//4
    assert np.abs(np_(r.p)[:, ok] - ref['last'][0:3][:, ok]).max() <= TOL_MM
    assert np.abs(np_(r.op)[ok] - ref['op'][ok]).max() <= TOL_MM*1e3
    summ = np_(r.summary)
    assert summ[:, 0].sum() == ok.sum()


def test_cuda_drivers_match_oracle_seam():
    """the same driver calls through the CUDA engine give the same packages"""
    opm = load_model('dblgauss')
    fld, wvl = opm.optical_spec.field_of_view.fields[1], 656.3
    fan_def = [np.array([0., -1.]), np.array([0., 1.]), 13]
    a = TR.trace_fan(opm, fan_def, fld, wvl, 0.0)
    b = TR.trace_fan(opm, [np.array([0., -1.]), np.array([0., 1.]), 13], fld, wvl, 0.0,
                     tracer=oracle_tracer)
    assert len(a) == len(b) == 13
    for (pa, ka), (pb, kb) in zip(a, b):
        assert np.array_equal(pa, pb) and ka[1] == kb[1]
        for sa, sb in zip(ka[0], kb[0]):
            assert all(np.array_equal(x, y) for x, y in zip(sa, sb))
    grid_def = [np.array([-1., -1.]), np.array([1., 1.]), 9]
    ga = TR.trace_grid(opm, grid_def, fld, wvl, 0.0, img_filter=lambda p, k: -1.0 if k is None else k[1])
    gb = TR.trace_grid(opm, [np.array([-1., -1.]), np.array([1., 1.]), 9], fld, wvl, 0.0,
                       img_filter=lambda p, k: -1.0 if k is None else k[1], tracer=oracle_tracer)
    assert np.array_equal(ga, gb) and ga.shape == (9, 9)
    ray = A.Ray(opm, [0.2, -0.4], f=2, wl=486.1)
    ref = A.Ray(opm, [0.2, -0.4], f=2, wl=486.1, tracer=oracle_tracer)
    assert np.array_equal(ray.t_abr, ref.t_abr)
    wf = opm.seq_model.trace_wavefront(fld, wvl, 0.0, num_rays=8)
    wf2 = opm.seq_model.trace_wavefront(fld, wvl, 0.0, num_rays=8, tracer=oracle_tracer)
    assert np.array_equal(wf, wf2)


def test_cuda_opd_infinite_reference(oracle):
    """telecentric image space: the axial tiles use wave_abr_full_calc_inf_ref"""
    _opd_check(oracle, 'telecentric')


@pytest.mark.parametrize('name', ANGULAR_MODEL_NAMES)
def test_cuda_angular_pupil_grid(oracle, name):
    """start rays from an angular object-space pupil ('NA' bit-exact; 'f/#' has one libm
    pow() in the reference -> <= 1e-11 mm), generated by the general grid kernel"""
    opm = load_model(name)
    tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    grid = E.grid_for_model(opm, tab, 24)
    assert grid.pupil_kind == (1 if name == 'relay_na' else 2)
    r = E.trace_grid(tab, grid)
    torch.cuda.synchronize()
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    ref = oracle.trace_grid(grid.c_spec(), tab.descs, tab.n_by_wvl, 0, grid.n_rays, opts, n_threads=4)
    assert np.array_equal(np_(r.status), ref['status'])
    ok = ref['status'] == 0
    assert ok.sum() > grid.n_rays//3
    got = np.concatenate([np_(r.p), np_(r.d)])
    if name == 'relay_na':
        assert np.array_equal(got, ref['last'][0:6]) and np.array_equal(np_(r.op), ref['op'])
        assert np.array_equal(np_(r.abr)[:, ok], ref['abr'][:, ok])
    else:
        assert np.abs(got - ref['last'][0:6]).max() <= TOL_MM
        assert (got == ref['last'][0:6]).mean() > 0.9


def test_cuda_trace_list_of_rays():
    from test_trace_drivers import oracle_bundle_tracer
    opm = load_model('triplet')
    sm = opm.seq_model
    v = load_vectors('triplet')
    idx = np.nonzero(v['case'] == 0)[0][:64]
    rays = [(v['p0'][:, k], v['d0'][:, k], sm.wvlns[v['wvl_idx'][k]]) for k in idx]
    ca = v['cases'][0]['check_apertures']
    a = A.trace_list_of_rays(opm, rays, output_filter='last', check_apertures=ca)
    b = A.trace_list_of_rays(opm, rays, output_filter='last', check_apertures=ca,
                             tracer=oracle_bundle_tracer)
    assert len(a) == len(b) > 0
    for (sa, oa, wa), (sb, ob, wb) in zip(a, b):
        assert oa == ob and wa == wb and all(np.array_equal(x, y) for x, y in zip(sa, sb))


def test_cuda_batched_aiming():
    """chief-ray aiming and clear apertures of all fields through bundle launches"""
    from rayoptics_b200 import vigcalc as V
    from test_trace_drivers import oracle_bundle_fn
    opm_a, opm_b = load_model('dblgauss'), load_model('dblgauss')
    a = V.aim_all_fields_batched(opm_a)
    b = V.aim_all_fields_batched(opm_b, oracle_bundle_fn(opm_b))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    V.set_clear_apertures_batched(opm_a)
    V.set_clear_apertures_batched(opm_b, oracle_bundle_fn(opm_b))
    assert [i.max_aperture for i in opm_a.seq_model.ifcs] == [i.max_aperture for i in opm_b.seq_model.ifcs]


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'triplet', 'telecentric', 'cellphone'])
def test_cuda_analysis_classes_match_the_references(name):
    """RayFan / RayList / RayGrid against the SAME classes of the reference run on its own
    trace_raw (tests/golden/vectors/<model>_analyses.npz, generator make_golden_analyses.py):
    pupil coordinates and transverse aberrations bit for bit; OPD within 1e-12 mm (F**2 is
    libm pow() in the reference) expressed in waves."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_analyses.npz'))
    n_fan, n_list, n_grid = (int(x) for x in z['num'])
    opm = load_model(name)
    for ci, (f, wl) in enumerate(z['cases']):
        f, wl = int(f), (None if wl < 0 else float(wl))
        wvl = opm.seq_model.central_wavelength() if wl is None else wl
        tol = 1e-12/opm.nm_to_sys_units(wvl)            # 1e-12 mm in waves
        for xy in 'xy':
            fan = A.RayFan(opm, f=f, wl=wl, xyfan=xy, num_rays=n_fan)
            pup = np.array([p for p, v in fan.fan], dtype=float).reshape(-1, 2)
            val = np.array([v for p, v in fan.fan], dtype=float).reshape(-1, 3)
            assert np.array_equal(pup, z[f'fan{xy}_pupil_{ci}'])
            gold = z[f'fan{xy}_vals_{ci}']
            assert np.array_equal(val[:, :2], gold[:, :2])
            assert np.abs(val[:, 2] - gold[:, 2]).max() <= tol
        rl = A.RayList(opm, num_rays=n_list, f=f, wl=wl)
        assert np.array_equal(rl.ray_abr, z[f'list_abr_{ci}'])
        rg = A.RayGrid(opm, f=f, wl=wl, num_rays=n_grid)
        gold = z[f'grid_{ci}']
        assert rg.grid.shape == gold.shape
        assert np.array_equal(rg.grid[0], gold[0]) and np.array_equal(rg.grid[1], gold[1])
        np.testing.assert_allclose(rg.grid[2], gold[2], rtol=0, atol=tol, equal_nan=True)
        assert (rg.grid[2][np.isfinite(gold[2])] == gold[2][np.isfinite(gold[2])]).mean() > 0.9


@pytest.mark.parametrize('name', ['dblgauss', 'telecentric'])
def test_cuda_psf_matches_the_references(name):
    """calc_psf (torch.fft on the device) on the reference's own wavefront grid against the
    reference's calc_psf (numpy fft): library transforms, 1e-12 of the normalised peak"""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_analyses.npz'))
    n_grid, dim = int(z['num'][2]), int(z['psf_dim'])
    for ci in range(len(z['cases'])):
        got = A.calc_psf(z[f'grid_{ci}'][2], n_grid, dim)
        np.testing.assert_allclose(got, z[f'psf_{ci}'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', ['singlet', 'dblgauss', 'rc', 'evenasph', 'cellphone', 'zoom52',
                                  'fisheye', 'threemir'])
def test_cuda_baseline_configs_grid_of_the_reference(name):
    """The BASELINE configurations at reduced pupil sampling: one grid launch of the engine
    against the body of the reference's own trace_grid loop (<model>_grid.npz), bit for bit --
    status of every ray, image intercept, direction and op_delta of the rays that arrive"""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_grid.npz'))
    opm = load_model(name)
    tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
    grid = E.grid_for_model(opm, tab, int(z['num']))
    assert grid.n_rays == z['status'].size
    r = E.trace_grid(tab, grid)
    torch.cuda.synchronize()
    assert np.array_equal(np_(r.status), z['status'])
    ok = z['status'] == 0
    assert np.array_equal(np_(r.p).T[ok], z['p'][ok]) and np.array_equal(np_(r.d).T[ok], z['d'][ok])
    assert np.array_equal(np_(r.op)[ok], z['op'][ok])


@pytest.mark.parametrize('lean', [0, 1, 2])
def test_cuda_random_systems(oracle, lean):
    """the randomised systems of tests/test_fuzz_three_way.py (reference == oracle == host-compiled
    device source on CPU) through the CUDA kernels: bit for bit against the oracle, whole rays"""
    import test_fuzz_three_way as F
    for seed in range(40):
        rng = np.random.default_rng(1000 + seed + 100000*lean)
        segs = F.random_system(rng, lean)
        tab = T.SurfaceTable.from_path(segs, device=0)
        n = 400
        p0 = np.zeros((3, n))
        p0[:2] = rng.uniform(-6, 6, (2, n))
        tgt = rng.uniform(-5, 5, (2, n))*np.where(rng.random(n) < 0.85, 1.0, 4.0)
        v = np.array([tgt[0] - p0[0], tgt[1] - p0[1], np.full(n, abs(segs[0][2][1][2]))])
        d0 = v/np.sqrt((v*v).sum(0))
        wv = np.zeros(n, dtype=np.int32)
        case = F.random_case(rng, tab.n_ifc)
        ref = oracle.trace_bundle(tab.descs, tab.n_by_wvl, p0, d0, wv, _abi.make_opts(**case),
                                  want_full=True, n_threads=4)
        r = E.trace_bundle(tab, p0, d0, wvl_idx=wv, full=True, **case)
        torch.cuda.synchronize()
        last = np.concatenate([np_(r.p), np_(r.d), np_(r.dst)[None], np_(r.nrml)])
        for got, key in ((np_(r.status), 'status'), (np_(r.fail_surf), 'fail_surf'),
                         (np_(r.n_seg), 'n_seg'), (np_(r.op), 'op'), (last, 'last'), (np_(r.full), 'full')):
            assert np.array_equal(got, ref[key], equal_nan=True), (lean, seed, key, case)
