"""The reference arm of bench.py: the UNMODIFIED Python reference (mjhoptics/ray-optics) timed on
the box's host cores.

`baseline/_ref` (git-ignored, travels to the GPU box with the snapshot) is a plain
`pip install --no-deps --target baseline/_ref` of the reference (baseline/install_reference.sh).
Nothing here reads /root/reference.  The timed call is the reference's own public function for
the path,

    rayoptics.raytr.trace.trace_grid(opt_model, grid_rng, fld, wvl, foc, check_apertures=True)

(raytr/trace.py:563-605) -> trace_safe -> trace_base -> OpticalSpecs.ray_start_from_osp ->
raytrace.trace -> raytrace.trace_raw on the reference's own Surface / profile objects -- stock
code all the way down, none of this repo's kernels, engine or oracle on the path.

What is NOT the reference's: the container that hands those objects to it.  The reference's
SequentialModel / OpticalModel constructors need `opticalglass` (glass catalogs), `anytree` and
`json_tricks`, which are not installed and cannot be (no network), so `build_model` assembles
the model programmatically: reference `Surface`s and profiles in path tuples, the reference's
`OpticalSpecs` / `PupilSpec` / `FieldSpec` / `Field` objects filled with the prescription's
numbers, refractive indices and first-order data taken from this repo's model loader (those
are inputs shared by both arms; tests/test_firstorder_vs_reference.py pins them to the
reference's `compute_first_order`).  Missing third-party modules are satisfied by inert
stand-ins (`_stub_missing`), none of which is called on the path.

The Python reference is single-threaded; `PoolRunner` spreads sub-blocks of the pupil grid
over one worker process per host core (fork), each worker calling `trace_grid` on its block.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import time
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')
THIRD_PARTY = ('opticalglass', 'anytree', 'transforms3d', 'json_tricks', 'parsimonious',
               'deprecation', 'matplotlib', 'PySide6', 'qdarkstyle', 'qtconsole', 'IPython',
               'ipywidgets', 'traitlets', 'requests', 'packaging', 'pandas', 'scipy')


def available():
    return os.path.isdir(os.path.join(REF, 'rayoptics', 'raytr'))


class _Stub(types.ModuleType):
    """Stand-in for a third-party module that is not installed: CapitalCase attributes are empty
    classes (usable as base classes / annotations), lowercase ones are sub-modules."""

    def __call__(self, *a, **kw):
        if len(a) == 1 and callable(a[0]) and not kw:
            return a[0]
        return lambda fn: fn

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        if k[0].isupper():
            c = type(k, (), {'__init__': lambda self, *a, **kw: None})
            setattr(self, k, c)
            return c
        try:
            return importlib.import_module(self.__name__ + '.' + k)
        except Exception:
            raise AttributeError(k) from None


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_mods = None


def _stub_missing():
    missing = []
    for r in THIRD_PARTY:
        try:
            importlib.import_module(r)
        except Exception:
            missing.append(r)
    if missing:
        sys.meta_path.append(_StubFinder(missing))
    return missing


def modules():
    """Import the reference's hot-path modules from baseline/_ref (once)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError('baseline/_ref is not installed (baseline/install_reference.sh)')
    if REF not in sys.path:
        sys.path.insert(0, REF)
    stubs = _stub_missing()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = types.SimpleNamespace(
            raytrace=importlib.import_module('rayoptics.raytr.raytrace'),
            trace=importlib.import_module('rayoptics.raytr.trace'),
            opticalspec=importlib.import_module('rayoptics.raytr.opticalspec'),
            surface=importlib.import_module('rayoptics.elem.surface'),
            profiles=importlib.import_module('rayoptics.elem.profiles'),
            firstorder=importlib.import_module('rayoptics.parax.firstorder'),
            stubs=stubs)
    assert os.path.realpath(m.raytrace.__file__).startswith(os.path.realpath(REF)), m.raytrace.__file__
    _mods = m
    return m


# ------------------------------------------------------------------ model assembly
def _profile(P, prf):
    name = type(prf).__name__
    if name == 'Spherical':
        return P.Spherical(c=prf.cv)
    if name == 'Conic':
        return P.Conic(c=prf.cv, cc=prf.cc)
    if name == 'EvenPolynomial':
        return P.EvenPolynomial(c=prf.cv, cc=prf.cc, coefs=list(prf.coefs)).update()
    if name == 'RadialPolynomial':
        return P.RadialPolynomial(c=prf.cv, ec=prf.ec, coefs=list(prf.coefs)).update()
    raise NotImplementedError(f'reference arm: profile {name}')


def _surface(M, ifc):
    s = M.surface.Surface(profile=_profile(M.profiles, ifc.profile), interact_mode=ifc.interact_mode,
                          max_ap=ifc.max_aperture)
    if ifc.clear_apertures:
        raise NotImplementedError('reference arm: clear-aperture lists')
    return s


class _Gap:
    def __init__(self, thi):
        self.thi = thi


class RefSeq:
    """The attributes and methods of SequentialModel that the reference's trace / opticalspec
    code reads (seq/sequential.py:149-202,259-274), holding reference Surface objects."""

    def __init__(self, M, sm):
        self.ifcs = [_surface(M, ifc) for ifc in sm.ifcs]
        self.gaps = [_Gap(g.thi) for g in sm.gaps]
        self.z_dir = list(sm.z_dir)
        self.stop_surface = sm.stop_surface
        self.wvlns = list(sm.wvlns)
        self._ref_wvl = sm.central_wavelength()
        self._paths = {}
        for wl in self.wvlns:
            self._paths[wl] = [[self.ifcs[i], None, seg[2], seg[3], seg[4]]
                               for i, seg in enumerate(sm.path(wl))]
        self._n_central = [seg[3] for seg in self._paths[self._ref_wvl]]

    def path(self, wl=None, start=None, stop=None, step=1):
        return iter(self._paths[self._ref_wvl if wl is None else wl][start:stop:step])

    def get_num_surfaces(self):
        return len(self.ifcs)

    def central_wavelength(self):
        return self._ref_wvl

    def index_for_wavelength(self, wvl):
        return self.wvlns.index(wvl)

    def central_rndx(self, i):
        n = self._n_central[i]
        return self._n_central[-2] if n is None else n


class RefModel:
    def __init__(self, M, opm):
        sm, osp = opm.seq_model, opm.optical_spec
        OS = M.opticalspec
        self.seq_model = RefSeq(M, sm)
        fod = osp.fod
        PD = M.firstorder.ParaxData
        self.analysis_results = {'parax_data': PD(fod.ax_ray, fod.pr_ray, fod)}
        self._sub = {'seq_model': self.seq_model, 'sm': self.seq_model,
                     'analysis_results': self.analysis_results, 'ar': self.analysis_results}
        ros = OS.OpticalSpecs.__new__(OS.OpticalSpecs)      # __init__ would look up glass-catalog lines
        ros.opt_model = self
        ros._submodels = {}
        ros.do_aiming = False
        wv = OS.WvlSpec(do_init=False)
        wv.wavelengths = list(osp.spectral_region.wavelengths)
        wv.spectral_wts = list(osp.spectral_region.spectral_wts)
        wv.reference_wvl = osp.spectral_region.reference_wvl
        ros._submodels['wvls'] = wv
        ros._submodels['pupil'] = OS.PupilSpec(ros, key=tuple(osp.pupil.key), value=osp.pupil.value)
        fs = OS.FieldSpec(ros, key=tuple(osp.field_of_view.key), value=osp.field_of_view.value,
                          is_relative=osp.field_of_view.is_relative,
                          is_wide_angle=osp.field_of_view.is_wide_angle, do_init=False)
        fs.fields = []
        for f in osp.field_of_view.fields:
            rf = OS.Field(x=f.x, y=f.y, wt=getattr(f, 'wt', 1.0), fov=fs)
            rf.vux, rf.vuy, rf.vlx, rf.vly = f.vux, f.vuy, f.vlx, f.vly
            rf.aim_info = None if f.aim_info is None else np.array(f.aim_info, dtype=float)
            fs.fields.append(rf)
        ros._submodels['fov'] = fs
        ros._submodels['focus'] = OS.FocusRange(osp.defocus.focus_shift)
        self.optical_spec = ros
        self._sub.update(optical_spec=ros, osp=ros)

    def __getitem__(self, key):
        return self._sub[key]


def build_model(opm):
    """mirror OpticalModel (this repo's loader) -> model made of the reference's own objects"""
    return RefModel(modules(), opm)


STATUS = {'TraceMissedSurfaceError': 1, 'TraceTIRError': 2, 'TraceRayBlockedError': 3,
          'TraceEvanescentRayError': 4}


def _image_point(pupil, ray_pkg):
    """img_filter in the style of the reference's spot-diagram filter
    (mpl/axisarrayfigure.py:229-238): pupil coordinates and image intercept of a ray that
    arrives, nothing for one that does not."""
    if ray_pkg is None:
        return None
    seg = ray_pkg[0][-1]
    return np.array([pupil[0], pupil[1], seg[0][0], seg[0][1]])


def trace_block(R, fi, wi, x0, x1, y0, y1, num, want=False):
    """The reference's trace_grid over the pupil block [x0, x1] x [y0, y1], num x num samples,
    called the way the reference's own spot diagram calls it (form='list',
    append_if_none=False, seq/sequential.py:1058-1085).  Returns the number of rays traced
    (num*num) or, if `want`, the [n_ok, 4] array (pupil x, y, image x, y) of the rays that arrive."""
    M = modules()
    osp = R['osp']
    fld = osp['fov'].fields[fi]
    wvl = R['sm'].wvlns[wi]
    foc = osp['focus'].focus_shift
    grid_rng = (np.array([x0, y0]), np.array([x1, y1]), num)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        g = M.trace.trace_grid(R, grid_rng, fld, wvl, foc, img_filter=_image_point, form='list',
                               append_if_none=False, output_filter='last', rayerr_filter='summary')
    return g if want else num*num


# ------------------------------------------------------------------ process pool
_W = {}


def _worker_init(model_path):
    sys.path.insert(0, os.path.dirname(HERE))
    from rayoptics_b200 import model as MM
    _W['R'] = build_model(MM.OpticalModel.load(model_path))


def _worker_run(task):
    return trace_block(_W['R'], *task)


class PoolRunner:
    """One worker process per core, each holding its own copy of the model; a step = a list of
    pupil blocks traced by the reference's trace_grid."""

    def __init__(self, model_path, procs):
        import multiprocessing as mp
        self.procs = int(procs)
        ctx = mp.get_context('fork')
        self.pool = ctx.Pool(self.procs, initializer=_worker_init, initargs=(model_path,))

    def run(self, tasks):
        t0 = time.perf_counter()
        n = sum(self.pool.map(_worker_run, tasks, chunksize=1))
        return time.perf_counter() - t0, n

    def close(self):
        self.pool.close()
        self.pool.join()


def block_tasks(n_fields, n_wvls, blocks_per_side, num):
    """Every (field, wavelength) tile's pupil square [-1, 1]^2 cut into blocks_per_side^2 blocks
    of num x num samples: the sample the reference arm traces per step."""
    edges = np.linspace(-1.0, 1.0, blocks_per_side + 1)
    tasks = []
    for fi in range(n_fields):
        for wi in range(n_wvls):
            for bx in range(blocks_per_side):
                for by in range(blocks_per_side):
                    # interior end points stop one sample short so that blocks do not overlap
                    x1 = edges[bx + 1] - (edges[bx + 1] - edges[bx])/num*(bx < blocks_per_side - 1)
                    y1 = edges[by + 1] - (edges[by + 1] - edges[by])/num*(by < blocks_per_side - 1)
                    tasks.append((fi, wi, float(edges[bx]), float(x1), float(edges[by]), float(y1), num))
    return tasks


def host_info():
    """what the box really offers: logical CPUs, affinity mask, cgroup CPU quota, load"""
    info = {'cpu_count': os.cpu_count()}
    try:
        info['affinity'] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            info['cgroup:' + os.path.basename(f)] = open(f).read().strip()
        except Exception:
            pass
    try:
        info['loadavg'] = os.getloadavg()[0]
    except Exception:
        pass
    return info


def run_arm(model, steps, warmup, step_s=1.5, procs=None):
    """K timed steps of the reference on `procs` worker processes (default: one per core); each
    step traces every (field, wvl) tile's pupil square cut into blocks (block_tasks), the block
    size chosen from a calibration step so that a step lasts about `step_s` seconds."""
    sys.path.insert(0, os.path.dirname(HERE))
    from rayoptics_b200 import model as MM
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'models', model + '.json')
    opm = MM.OpticalModel.load(path)
    n_fields, n_wvls = len(opm.optical_spec.field_of_view.fields), len(opm.seq_model.wvlns)
    cores = int(procs or os.cpu_count() or 1)
    # one process, one block: the reference's single-core rate on this box
    R1 = build_model(opm)
    trace_block(R1, 0, 0, -1.0, 1.0, -1.0, 1.0, 8)
    t1 = time.perf_counter()
    n1 = trace_block(R1, n_fields - 1, n_wvls//2, -1.0, 1.0, -1.0, 1.0, 40)
    single = n1/(time.perf_counter() - t1)
    pool = PoolRunner(path, cores)
    tiles = n_fields*n_wvls
    bps = 1
    while tiles*bps*bps < 2*cores:
        bps += 1
    pool.run(block_tasks(n_fields, n_wvls, bps, 6))             # imports, page-in
    num_b = 8
    for _ in range(3):                                          # calibration: two refinements
        dt, n = pool.run(block_tasks(n_fields, n_wvls, bps, num_b))
        nxt = int(max(6, min(128, round((n/dt*step_s/(tiles*bps*bps))**0.5))))
        if abs(nxt - num_b) <= max(1, num_b//10):
            break
        num_b = nxt
    tasks = block_tasks(n_fields, n_wvls, bps, num_b)
    for _ in range(max(warmup, 1)):
        pool.run(tasks)
    times, n_step = [], 0
    for _ in range(steps):
        dt, n_step = pool.run(tasks)
        times.append(dt)
    pool.close()
    tot = float(np.sum(times))
    return {'value': n_step*steps/tot, 'unit': 'rays/s', 'cores': cores, 'kind': 'reference',
            'rays_per_step': int(n_step), 'rays_per_s_per_core': n_step*steps/tot/cores,
            'single_process': single, 'host': host_info(),
            'ms_per_step': 1e3*tot/steps,
            'sample': f'{steps} steps x {n_step} rays: all {tiles} (field, wvl) tiles, each pupil square cut '
                      f'into {bps}x{bps} blocks of {num_b}x{num_b} samples, every block traced by the unmodified '
                      f'rayoptics.raytr.trace.trace_grid (-> trace_safe -> trace_base -> ray_start_from_osp -> '
                      f'raytrace.trace_raw) from baseline/_ref, one worker process per core ({cores})'}


if __name__ == '__main__':
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='dblgauss')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--step-s', type=float, default=1.5)
    ap.add_argument('--procs', type=int, default=None)
    a = ap.parse_args()
    if not available():
        print(json.dumps({'unavailable': 'baseline/_ref is not installed'}))
    else:
        print(json.dumps(run_arm(a.model, a.steps, a.warmup, a.step_s, a.procs)))
