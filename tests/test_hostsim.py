"""The per-ray device source (rayoptics_b200/csrc/rt_device.cuh, rt_lean.cuh)
compiled for the HOST by tests/hostsim and compared bit for bit with the oracle
and the reference's golden vectors.

What this pins without a GPU: the algebra of the general loop and of the lean
loop's exact shortcuts (shared-reciprocal division, the sqrt sequence, the
aperture band, the branch-free fast path and its fallbacks), on every model of
the golden set.  What it cannot pin: the MUFU seeds and ptxas' code generation --
those are covered by the `-m gpu` tests.  tests/hostsim is test infrastructure;
the product has no CPU path.
"""
import numpy as np
import pytest

from conftest import MODEL_NAMES, PHASE_MODEL_NAMES, load_model, load_vectors, seeded_bundle
from rayoptics_b200 import _abi, table as T
from hostsim import build as HS


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope='module')
def hostsim():
    HS.lib()
    return HS


def kernels_for(descs):
    kind = HS.lean_kind(descs)
    return [0] if kind == 0 else [0, kind]


def compare(r, ref, out_kind, n_full=None):
    assert same(r['status'], ref['status'])
    assert same(r['fail_surf'], ref['fail_surf'])
    assert same(r['n_seg'], ref['n_seg'])
    assert same(r['op'], ref['op'])
    assert same(r['last'][0:6], ref['last'][0:6])
    if out_kind >= 1:
        assert same(r['last'][6:10], ref['last'][6:10])
    if out_kind == 2 and ref.get('full') is not None:
        assert same(r['full'], ref['full'])


@pytest.mark.parametrize('name', MODEL_NAMES)
def test_device_source_matches_reference_vectors(hostsim, name):
    opm = load_model(name)
    descs, n_by_wvl, wvls = T.describe_model(opm.seq_model)
    v = load_vectors(name)
    n_full = v['full'].shape[2]
    for ci, case in enumerate(v['cases']):
        idx = np.nonzero(v['case'] == ci)[0]
        if idx.size == 0:
            continue
        opts = _abi.make_opts(**case)
        for kern in kernels_for(descs):
            r = hostsim.trace_bundle(descs, n_by_wvl, v['p0'][:, idx], v['d0'][:, idx],
                                     v['wvl_idx'][idx], opts, kernel=kern, out_kind=2, wvls=wvls)
            st = r['status']
            assert same(st, v['status'][idx])
            assert same(r['fail_surf'], np.where(st == 0, -1, v['fail_surf'][idx]))
            assert same(r['n_seg'], v['n_seg'][idx])
            assert same(r['op'], v['op'][idx])
            assert same(r['last'], v['last'][:, idx])
            sel = idx < n_full
            assert same(r['full'][:, :, sel], v['full'][:, :, idx[sel]])


@pytest.mark.parametrize('name', MODEL_NAMES)
def test_device_source_matches_oracle_bundle(hostsim, oracle, name):
    opm = load_model(name)
    descs, n_by_wvl, wvls = T.describe_model(opm.seq_model)
    rng = np.random.default_rng(11)
    n = 6000
    p0, d0, wv = seeded_bundle(opm, n, rng)
    n_ifc = len(descs)
    for case in (dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=True),
                 dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=False),
                 dict(first_surf=2, last_surf=n_ifc - 3, check_apertures=True,
                      filter_out_phantoms=True, intersect_obj=False)):
        opts = _abi.make_opts(**case)
        ref = oracle.trace_bundle(descs, n_by_wvl, p0, d0, wv, opts, want_full=True, n_threads=4,
                                  wvls=wvls)
        for kern in kernels_for(descs):
            for out_kind in ((2,) if kern == 0 else (0, 1, 2)):
                r = hostsim.trace_bundle(descs, n_by_wvl, p0, d0, wv, opts, kernel=kern,
                                         out_kind=out_kind, wvls=wvls)
                compare(r, ref, out_kind)


PHASE_TOL_MM = 1e-11        # north star: <= 1e-10 mm; observed: a few 1e-14


def close_records(r, ref, tol=PHASE_TOL_MM):
    """status / failing surface / segment count exact; positions, directions and path
    lengths within `tol`; returns the fraction of rays that is bit-identical."""
    assert same(r['status'], ref['status'])
    assert same(r['fail_surf'], ref['fail_surf'])
    assert same(r['n_seg'], ref['n_seg'])
    d_last = np.abs(r['last'] - ref['last'])
    d_op = np.abs(r['op'] - ref['op'])
    assert np.array_equal(np.isnan(r['last']), np.isnan(ref['last']))
    assert np.nanmax(d_last, initial=0.0) <= tol and np.nanmax(d_op, initial=0.0) <= tol*1e3
    return float(((np.nan_to_num(d_last).max(0) == 0) & (np.nan_to_num(d_op) == 0)).mean())


@pytest.mark.parametrize('name', PHASE_MODEL_NAMES)
def test_device_source_phase_elements(hostsim, oracle, name):
    """Diffractive phase elements (grating, radial DOE): x**k is libm pow() in the
    reference, exact-rounded products on the device -> tolerance parity, almost all
    rays still bit-identical."""
    opm = load_model(name)
    descs, n_by_wvl, wvls = T.describe_model(opm.seq_model)
    assert HS.lean_kind(descs) == 0
    v = load_vectors(name)
    for ci, case in enumerate(v['cases']):
        idx = np.nonzero(v['case'] == ci)[0]
        if idx.size == 0:
            continue
        r = hostsim.trace_bundle(descs, n_by_wvl, v['p0'][:, idx], v['d0'][:, idx],
                                 v['wvl_idx'][idx], _abi.make_opts(**case), kernel=0, out_kind=2,
                                 wvls=wvls)
        ref = {k: v[k][..., idx] for k in ('last', 'op', 'status', 'n_seg')}
        ref['fail_surf'] = np.where(v['status'][idx] == 0, -1, v['fail_surf'][idx])
        assert close_records(r, ref) > 0.98
    rng = np.random.default_rng(3)
    p0, d0, wv = seeded_bundle(opm, 8000, rng)
    opts = _abi.make_opts(first_surf=1, last_surf=len(descs) - 2, check_apertures=True)
    ref = oracle.trace_bundle(descs, n_by_wvl, p0, d0, wv, opts, want_full=True, n_threads=4, wvls=wvls)
    r = hostsim.trace_bundle(descs, n_by_wvl, p0, d0, wv, opts, kernel=0, out_kind=2, wvls=wvls)
    assert close_records(r, ref) > 0.98
    assert (ref['status'] == 0).sum() > 1000
    np.testing.assert_allclose(np.nan_to_num(r['full']), np.nan_to_num(ref['full']), rtol=0,
                               atol=PHASE_TOL_MM)


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'cellphone', 'triplet', 'telecentric'])
def test_device_opd_epilogue(hostsim, oracle, name):
    """wave_opd of rt_device.cuh (finite and infinite reference sphere) on the reference's
    golden OPDs.  Finite spheres: F**2 is libm pow() in the reference, F*F on the device
    (<= 1e-12 mm, > 95 % bit-equal); the infinite-reference variant has no power and is
    bit-exact."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'vectors', name + '_opd.npz'))
    v = {k: z[k] for k in z.files}
    opm = load_model(name)
    descs, n_by_wvl, _ = T.describe_model(opm.seq_model)
    n_ifc = len(descs)
    opts = _abi.make_opts(first_surf=1, last_surf=n_ifc - 2, check_apertures=True)
    r = hostsim.trace_bundle(descs, n_by_wvl, v['p0'], v['d0'], v['wvl_idx'], opts,
                             kernel=HS.lean_kind(descs), out_kind=2)
    assert same(r['status'], v['status'])
    ok = np.nonzero(v['status'] == 0)[0]
    got = np.full(v['opd'].shape, np.nan)
    for k in ok:
        full = r['full'][:, :, k]
        got[k] = hostsim.wave_opd(v['wave'][v['tile'][k]], full[1, 0:3], full[0, 3:6],
                                  full[n_ifc - 2, 0:3], full[n_ifc - 2, 3:6],
                                  full[n_ifc - 1, 0:3], full[n_ifc - 1, 3:6], r['op'][k])
    np.testing.assert_allclose(got[ok], v['opd'][ok], rtol=0, atol=1e-12)
    assert (got[ok] == v['opd'][ok]).mean() > 0.95
    inf = v['wave'][v['tile'], 21] == 0
    if name == 'telecentric':
        assert (inf & (v['status'] == 0)).sum() > 100
    assert same(got[inf], v['opd'][inf])


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'cellphone', 'exotic', 'relay_na', 'relay_fno'])
def test_device_grid_start_rays(hostsim, oracle, name):
    """grid_start_ray of rt_grid.cuh (both instances) == oracle == the numpy expressions of
    ray_start_from_osp (rayoptics_b200/opticalspec.py, reference order of operations),
    for spatial ('epd') and angular ('NA', 'f/#') pupil specifications."""
    from rayoptics_b200 import engine as E
    opm = load_model(name)
    osp = opm.optical_spec
    num = 11
    spec = E.grid_spec_for_model(opm, num)
    kind = {'relay_na': 1, 'relay_fno': 2}.get(name, 0)
    assert spec.pupil_kind == kind
    cs = spec.c_spec()
    p, d, wv, pup = oracle.grid_start_rays(cs, 0, spec.n_rays)
    # numpy restatement, ray by ray (first wavelength tile of every field)
    xs = E.accumulated_steps(-1.0, 1.0, num)
    n_w = spec.n_wvls
    for fi, fld in enumerate(osp.field_of_view.fields):
        base = fi*n_w*num*num
        for i in range(0, num, 2):
            for j in range(0, num, 3):
                pt0, dir0 = osp.ray_start_from_osp(fld.apply_vignetting(np.array([xs[i], xs[j]])), fld)
                if dir0[2]*opm.seq_model.z_dir[0] < 0:
                    dir0 = -dir0
                k = base + i*num + j
                assert same(p[:, k], pt0) and same(d[:, k], dir0)
    pg, dg = hostsim.grid_start_rays(cs, 0, spec.n_rays, lean=False)
    assert same(pg, p)
    if kind == 2:         # (pupil*slope)**2 is libm pow() in the reference
        np.testing.assert_allclose(dg, d, rtol=0, atol=2e-16)
        assert (dg == d).mean() > 0.99
    else:
        assert same(dg, d)
    if kind == 0:
        pl, dl = hostsim.grid_start_rays(cs, 0, spec.n_rays, lean=True)
        assert same(pl, p) and same(dl, d)


def test_division_and_sqrt_sequences(hostsim):
    """CPU sibling of rt_selftest_division (b200rt.cu): wherever a sequence reports
    'fast' it equals the IEEE operation; div_shared() and sqrt_near_one() always do."""
    rng = np.random.default_rng(5)
    n = 400000
    a = rng.standard_normal(n)*10.0**rng.uniform(-300, 300, n)
    b = rng.standard_normal(n)*10.0**rng.uniform(-300, 300, n)
    a[:2000] = 0.0
    a[2000:3000] = rng.standard_normal(1000)*1e-310           # denormal numerators
    b[3000:3500] = rng.standard_normal(500)*1e-310            # denormal denominators
    bad, n_fast = hostsim.check_division(a, b)
    assert bad == 0
    a2 = rng.standard_normal(n)
    b2 = rng.uniform(0.5, 2.0, n)
    bad, n_fast = hostsim.check_division(a2, b2)
    assert bad == 0 and n_fast > 0.99*n
    x = np.concatenate([rng.uniform(0, 4, n), 10.0**rng.uniform(-300, 300, n),
                        np.float64(1.0) + np.arange(-3000, 3000)*2.0**-53, [0.0, 1.0, np.inf]])
    bad, n_fast = hostsim.check_sqrt(x)
    assert bad == 0 and n_fast > 0.6*x.size
