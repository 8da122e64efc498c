"""GPU (B200 box): the CUDA engine, called through the C ABI, against

1. the committed golden vectors produced by the reference's own trace_raw
   (tests/golden/vectors, bit-exact: status, failing surface, every segment);
2. the C oracle on larger seeded bundles (bit-exact);
3. the C oracle on pupil grids generated on the device (start rays, clipping,
   transverse aberration; spot sums to 1e-12 relative);
4. size-independent properties at BASELINE.json's full sizes.

Tolerance: north_star asks <= 1e-10 mm RMS; these tests demand 0 (== on fp64).
"""
import numpy as np
import pytest
import torch

from conftest import MODEL_NAMES, load_model, load_vectors, seeded_bundle
from rayoptics_b200 import _abi, table as T, engine as E

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def np_(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope='module')
def tables():
    cache = {}

    def get(name):
        if name not in cache:
            opm = load_model(name)
            cache[name] = (opm, T.SurfaceTable.from_model(opm.seq_model, device=0))
        return cache[name]
    return get


@pytest.mark.parametrize('name', MODEL_NAMES)
def test_cuda_matches_reference_vectors(tables, name):
    opm, tab = tables(name)
    v = load_vectors(name)
    n_full = v['full'].shape[2]
    for ci, case in enumerate(v['cases']):
        idx = np.nonzero(v['case'] == ci)[0]
        if idx.size == 0:
            continue
        r = E.trace_bundle(tab, v['p0'][:, idx], v['d0'][:, idx], wvl_idx=v['wvl_idx'][idx],
                           full=True, **case)
        torch.cuda.synchronize()
        st = np_(r.status)
        assert same(st, v['status'][idx])
        assert same(np_(r.fail_surf), np.where(st == 0, -1, v['fail_surf'][idx]))
        assert same(np_(r.n_seg), v['n_seg'][idx])
        assert same(np_(r.op), v['op'][idx])
        last = np.concatenate([np_(r.p), np_(r.d), np_(r.dst)[None], np_(r.nrml)])
        assert same(last, v['last'][:, idx])
        sel = idx < n_full
        assert same(np_(r.full)[:, :, sel], v['full'][:, :, idx[sel]])


@pytest.mark.parametrize('name', MODEL_NAMES)
def test_cuda_matches_oracle_bundle(tables, oracle, name):
    opm, tab = tables(name)
    rng = np.random.default_rng(7)
    n = 20000
    p0, d0, wv = seeded_bundle(opm, n, rng)
    n_ifc = tab.n_ifc
    for case in (dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=True),
                 dict(first_surf=1, last_surf=n_ifc - 2, check_apertures=False)):
        opts = _abi.make_opts(**case)
        ref = oracle.trace_bundle(tab.descs, tab.n_by_wvl, p0, d0, wv, opts, want_full=True,
                                  n_threads=8, wvls=tab.wvls)
        r = E.trace_bundle(tab, p0, d0, wvl_idx=wv, full=True, **case)
        torch.cuda.synchronize()
        assert same(np_(r.status), ref['status'])
        assert same(np_(r.fail_surf), ref['fail_surf'])
        assert same(np_(r.n_seg), ref['n_seg'])
        assert same(np_(r.op), ref['op'])
        last = np.concatenate([np_(r.p), np_(r.d), np_(r.dst)[None], np_(r.nrml)])
        assert same(last, ref['last'])
        assert same(np_(r.full), ref['full'])
        assert (ref['status'] == 0).sum() > n//20      # the bundle is not degenerate
        # last-segment-only launch (no whole-ray output) gives the same records
        r2 = E.trace_bundle(tab, p0, d0, wvl_idx=wv, full=False, **case)
        torch.cuda.synchronize()
        assert same(np_(r2.p), np_(r.p)) and same(np_(r2.d), np_(r.d)) and same(np_(r2.op), np_(r.op))


def oracle_grid(oracle, tab, grid, r0, r1, opts):
    spec = grid.c_spec()
    p, d, wv, pup = oracle.grid_start_rays(spec, r0, r1)
    ref = oracle.trace_bundle(tab.descs, tab.n_by_wvl, p, d, wv, opts, n_threads=8, wvls=tab.wvls)
    ref['p0'], ref['d0'] = p, d
    return ref


@pytest.mark.parametrize('name,num', [('singlet', 7), ('dblgauss', 64), ('rc', 64),
                                      ('cellphone', 32), ('evenasph', 32), ('zoom52', 16),
                                      ('thin_triplet', 48), ('exotic', 24)])
def test_cuda_grid_matches_oracle(tables, oracle, name, num):
    opm, tab = tables(name)
    grid = E.grid_for_model(opm, tab, num)
    assert grid.n_rays == len(opm.optical_spec.fov.fields)*len(opm.seq_model.wvlns)*num*num
    r = E.trace_grid(tab, grid)
    torch.cuda.synchronize()
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    ref = oracle_grid(oracle, tab, grid, 0, grid.n_rays, opts)
    assert same(np_(r.status), ref['status'])
    assert same(np_(r.fail_surf), ref['fail_surf'])
    assert same(np_(r.op), ref['op'])
    assert same(np_(r.p), ref['last'][0:3]) and same(np_(r.d), ref['last'][3:6])
    # transverse aberration per tile + spot sums
    per_tile = num*num
    abr = np_(r.abr)
    summ = np_(r.summary)
    for t in range(grid.n_tiles):
        sl = slice(t*per_tile, (t + 1)*per_tile)
        rx, ry = grid.ref_img.reshape(-1, 2)[t]
        ax, ay = oracle.transverse_abr(ref['last'][0, sl], ref['last'][1, sl], ref['last'][3, sl],
                                       ref['last'][4, sl], ref['last'][5, sl], grid.foc, rx, ry)
        assert same(abr[0, sl], ax) and same(abr[1, sl], ay)
        ok = ref['status'][sl] == 0
        st = ref['status'][sl]
        assert summ[t, 0] == ok.sum() and summ[t, 1] == (st == 1).sum()
        assert summ[t, 2] == (st == 2).sum() and summ[t, 3] == (st == 3).sum()
        if ok.any():
            np.testing.assert_allclose(summ[t, 5], ax[ok].sum(), rtol=1e-12,
                                       atol=1e-13*np.abs(ax[ok]).sum() + 1e-300)
            np.testing.assert_allclose(summ[t, 7], (ax[ok]**2).sum(), rtol=1e-12, atol=1e-18)
            np.testing.assert_allclose(summ[t, 9], (ax[ok]*ay[ok]).sum(), rtol=1e-11,
                                       atol=1e-13*np.abs(ax[ok]*ay[ok]).sum() + 1e-300)
            assert summ[t, 10] == ax[ok].min() and summ[t, 13] == ay[ok].max()
            np.testing.assert_allclose(summ[t, 14], ref['op'][sl][ok].sum(), rtol=1e-12)
    assert summ[:, 0:5].sum() == grid.n_rays


def test_grid_sharding_invariance(tables):
    """Tracing the chunk range in pieces (what ranks do) gives the same rays and,
    combined, the same sums as one call."""
    opm, tab = tables('dblgauss')
    grid = E.grid_for_model(opm, tab, 100)        # 100*100 is not a multiple of the chunk size
    whole = E.trace_grid(tab, grid)
    cuts = [0, 7, grid.n_chunks//3, grid.n_chunks//3 + 1, grid.n_chunks - 5, grid.n_chunks]
    parts, pieces = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        r = E.trace_grid(tab, grid, a, b)
        parts.append(r.summary)
        pieces.append(r)
    torch.cuda.synchronize()
    assert sum(p.n for p in pieces) == grid.n_rays
    assert torch.equal(torch.cat([p.status for p in pieces]), whole.status)
    assert torch.equal(torch.cat([p.p for p in pieces], dim=1), whole.p)
    assert torch.equal(torch.cat([p.abr for p in pieces], dim=1), whole.abr)
    comb = E.combine_summaries(parts)
    assert torch.equal(comb[:, 0:5], whole.summary[:, 0:5])
    assert torch.equal(comb[:, 10:14], whole.summary[:, 10:14])
    torch.testing.assert_close(comb[:, 5:10], whole.summary[:, 5:10], rtol=1e-12, atol=1e-14)


def test_full_size_properties(tables, oracle):
    """BASELINE config 2 at full size (3 fields x 3 wvls x 512 x 512): a seeded
    sample against the oracle plus size-independent invariants."""
    opm, tab = tables('dblgauss')
    num = 512
    grid = E.grid_for_model(opm, tab, num)
    assert grid.n_rays == 2359296
    r = E.trace_grid(tab, grid)
    r2 = E.trace_grid(tab, grid)                   # idempotent / deterministic
    torch.cuda.synchronize()
    assert torch.equal(r.p, r2.p) and torch.equal(r.summary, r2.summary)
    summ = np_(r.summary)
    assert summ[:, 0:5].sum() == grid.n_rays
    st = np_(r.status)
    per_tile = num*num
    for t in range(grid.n_tiles):
        assert (st[t*per_tile:(t + 1)*per_tile] == 0).sum() == summ[t, 0]
    # every unvignetted on-axis pupil ray inside the unit circle gets through
    xs = E.accumulated_steps(-1, 1, num)
    inside = (xs[:, None]**2 + xs[None, :]**2) <= 0.98
    assert (st[:per_tile].reshape(num, num)[inside] == 0).all()
    # seeded sample of whole chunks against the oracle
    rng = np.random.default_rng(3)
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    for c in rng.integers(0, grid.n_chunks, 40):
        a, b = grid.first_ray_of_chunk(int(c)), grid.first_ray_of_chunk(int(c) + 1)
        ref = oracle_grid(oracle, tab, grid, a, b, opts)
        assert same(st[a:b], ref['status'])
        assert same(np_(r.p[:, a:b]), ref['last'][0:3])
        assert same(np_(r.d[:, a:b]), ref['last'][3:6])
        assert same(np_(r.op[a:b]), ref['op'])


def test_division_selftest():
    """The shared-reciprocal division of the specialised kernels is bit-identical
    to the IEEE division on ~1e9 random operand sets incl. extreme exponents,
    zero numerators, all-ones and power-of-two denominators."""
    import ctypes as C
    lib = _abi.load_library()
    bad = C.c_uint64(123)
    for seed in (1, 2):
        _abi.check(lib.rt_selftest_division(0, 1184, 2000, seed, C.byref(bad)))
        assert bad.value == 0


@pytest.mark.parametrize('name', ['dblgauss', 'rc', 'zoom52'])
def test_general_and_specialised_kernels_agree(name, monkeypatch):
    """Models that qualify for the specialised (lean) kernels give the same bits
    through the general kernels (B200RT_NO_LEAN forces them)."""
    opm = load_model(name)
    lean = T.SurfaceTable.from_model(opm.seq_model, device=0)
    monkeypatch.setenv('B200RT_NO_LEAN', '1')
    general = T.SurfaceTable.from_model(opm.seq_model, device=0)
    monkeypatch.delenv('B200RT_NO_LEAN')
    grid_a = E.grid_for_model(opm, lean, 48)
    grid_b = E.grid_for_model(opm, general, 48)
    outs = E.GRID_OUTPUTS + ('nrml', 'dst', 'n_seg')
    a = E.trace_grid(lean, grid_a, outputs=outs, full=True)
    b = E.trace_grid(general, grid_b, outputs=outs, full=True)
    torch.cuda.synchronize()
    for k in ('p', 'd', 'nrml', 'dst', 'op', 'status', 'fail_surf', 'n_seg', 'abr', 'full', 'summary'):
        assert same(np_(getattr(a, k)), np_(getattr(b, k))), k


def test_dropin_trace_matches_reference_vectors(tables):
    """rayoptics_b200.raytrace.trace()/trace_raw(): the reference's call surface
    (raytrace.py:51-264) -- ray lists and TraceError subclasses -- reproduce the
    reference's own outputs on the golden rays, one call per ray."""
    from rayoptics_b200 import raytrace as rt
    ERR = {1: rt.TraceMissedSurfaceError, 2: rt.TraceTIRError, 3: rt.TraceRayBlockedError}
    for name in ('dblgauss', 'rc', 'cellphone'):
        opm = load_model(name)
        sm = opm.seq_model
        v = load_vectors(name)
        n_full = v['full'].shape[2]
        seen = set()
        for k in list(range(0, n_full, 3)) + list(range(n_full, v['p0'].shape[1], 11)):
            case = dict(v['cases'][v['case'][k]])
            wvl = sm.wvlns[v['wvl_idx'][k]]
            st = int(v['status'][k])
            seen.add(st)
            try:
                if case['first_surf'] == 1 and case['last_surf'] == sm.get_num_surfaces() - 2 \
                        and len(case) == 3:
                    ray, op, w = rt.trace(sm, v['p0'][:, k], v['d0'][:, k], wvl,
                                          check_apertures=case['check_apertures'])
                else:
                    ray, op, w = rt.trace_raw(sm.path(wvl), v['p0'][:, k], v['d0'][:, k], wvl, **case)
                assert st == 0
            except rt.TraceError as e:
                assert isinstance(e, ERR[st]) and e.surf == v['fail_surf'][k]
                ray, op, w = e.ray_pkg
            assert w == wvl and op == v['op'][k] and len(ray) == v['n_seg'][k]
            if len(ray):
                last = np.concatenate([ray[-1][0], ray[-1][1], [ray[-1][2]], ray[-1][3]])
                assert same(last, v['last'][:, k])
            if k < n_full:
                got = np.array([np.concatenate([s[0], s[1], [s[2]], s[3]]) for s in ray])
                assert same(got, v['full'][:len(ray), :, k])
        assert {0, 3} <= seen


@pytest.mark.parametrize('name,num,n_rays', [
    ('rc', 1024, 5*1024*1024),                 # BASELINE configs[2]: reflect path, 5 fields x 1024^2
    ('evenasph', 256, 9*3*256*256),            # configs[3]: even-asphere lens, 9 fields x 3 wvls x 256^2
    ('cellphone', 256, 9*3*256*256),           # configs[3] alternative: 8 RadialPolynomial surfaces
    ('zoom52', 1024, 25*7*1024*1024//8),       # configs[4]: one rank's 1/8 share of 183.5 M rays
])
def test_full_size_configs(tables, oracle, name, num, n_rays):
    """BASELINE configs 3-5 at full size: seeded chunks against the oracle
    (bit-exact) + invariants (counts add up, idempotent, shard == whole)."""
    opm, tab = tables(name)
    grid = E.grid_for_model(opm, tab, num)
    if name == 'zoom52':
        c0, c1 = grid.n_chunks//8*3, grid.n_chunks//8*4          # the 4th of 8 ranks' shard
    else:
        c0, c1 = 0, grid.n_chunks
        assert grid.n_rays == n_rays
    r = E.trace_grid(tab, grid, c0, c1, outputs=('p', 'd', 'op', 'status', 'abr'))
    r2 = E.trace_grid(tab, grid, c0, c1, outputs=('status', 'abr'))
    torch.cuda.synchronize()
    assert r.n == n_rays
    assert torch.equal(r.abr, r2.abr) and torch.equal(r.summary, r2.summary)
    summ = np_(r.summary)
    assert summ[:, 0:5].sum() == r.n
    st = np_(r.status)
    assert (st == 0).sum() == summ[:, 0].sum() and (st == 0).mean() > 0.5
    rng = np.random.default_rng(11)
    opts = _abi.make_opts(first_surf=1, last_surf=tab.n_ifc - 2, check_apertures=True)
    base = grid.first_ray_of_chunk(c0)
    for c in rng.integers(c0, c1, 24):
        a, b = grid.first_ray_of_chunk(int(c)), grid.first_ray_of_chunk(int(c) + 1)
        ref = oracle.trace_grid(grid.c_spec(), tab.descs, tab.n_by_wvl, a, b, opts, n_threads=8,
                                wvls=tab.wvls)
        sl = slice(a - base, b - base)
        assert same(st[sl], ref['status'])
        assert same(np_(r.p[:, sl]), ref['last'][0:3]) and same(np_(r.d[:, sl]), ref['last'][3:6])
        assert same(np_(r.op[sl]), ref['op']) and same(np_(r.abr[:, sl]), ref['abr'])
