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

from conftest import MODEL_NAMES, load_model, load_vectors, seeded_bundle
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
