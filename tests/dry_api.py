"""The DEFAULT-engine code paths of the host functions added after the last GPU call, on the dry-run
engine (TEST INFRASTRUCTURE, build container): every call below is made the way a user makes it -- no
test seam -- with tests/dryrun_engine.py standing in for the CUDA entry points.  What this checks is the
plumbing (arguments the engine functions accept, shapes that come back); the numbers are the oracle's.
Run by tests/test_bench_contract.py::test_default_engine_paths_of_the_new_host_functions."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np, torch
import dryrun_engine as D
D.install()
torch.cuda.is_available = lambda: True
from conftest import load_model
from rayoptics_b200 import vigcalc as V, trace as TR, analyses as A, wideangle as W, seq, zmx, raytrace as RT
# 1. bisection in one launch, default CUDA tile_fn
b = load_model('dblgauss')
for f in b.optical_spec.field_of_view.fields: f.clear_vignetting()
print('bisection', V.set_vig_by_bisection(b))
print('  vig', [(round(f.vuy,4), round(f.vly,4)) for f in b.optical_spec.field_of_view.fields])
# 2. batched set_vig / apertures / aiming defaults
print('set_vig_batched launches', V.set_vig_batched(load_model('triplet')))
# 3. post-import update default engine
path = os.path.join(HERE, 'golden', 'samples', 'triplet_fict.seq')
opm = seq.open_seq(path, do_update=True)
print('update ok', [round(i.max_aperture,3) for i in opm.seq_model.ifcs])
# 4. aim pt pupils default bundle tracer
m = load_model('dblgauss'); fld = m.optical_spec.field_of_view.fields[1]; wvl = m.seq_model.central_wavelength()
pkg = TR.trace_base(m, np.array([0., 5.0]), fld, wvl, pupil_type='aim pt')
print('aim pt', pkg[0][-1][0])
# 5. two-stage analyses default engine
fan = A.trace_fan(m, fld, wvl, 0.0, 1, num_rays=7); print('focus_fan', A.focus_fan(m, fan, fld, wvl, 0.02)[3])
g = A.trace_wavefront(m, fld, wvl, 0.0, num_rays=6); print('focus_wavefront', A.focus_wavefront(m, g, fld, wvl, 0.0).shape)
# 6. trace module functions default engine
print('refocus', TR.refocus(m), 'astig', TR.trace_astigmatism(m, fld, wvl, 0.0))
print('coddington', TR.trace_astigmatism_coddington_fan(m, fld, wvl, 0.0))
print('curve', [round(x,4) for x in TR.trace_astigmatism_curve(m, num_points=3)[1]])
print(TR.trace_all_fields(m).shape)
# 7. real image height via the drop-in trace_raw (reverse path table)
import test_trace_drivers as TD
rh, hts = TD._real_height_model('triplet')
rh.optical_spec._trace_raw_fn = None
fl = rh.optical_spec.field_of_view.fields[-1]
(p, d), z = W.eval_real_image_ht(rh, fl, rh.seq_model.central_wavelength())
print('real height z_enp', z)
pk = TR.trace_base(rh, np.array([0., 0.]), fl, rh.seq_model.central_wavelength())
print('  lands', pk[0][-1][0][:2], 'want', hts[-1])
# 8. wide-angle aiming via the drop-in trace
fish = load_model('fisheye')
for f in fish.optical_spec.field_of_view.fields: f.aim_info = None
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    print('z_enp', [round(z,4) for z in W.aim_wide_angle_fields(fish)])
    print('curve', W.eval_z_enp_curve(fish, printout=False, num_fields=3)[3])
# 9. set_pupil / set_stop_aperture defaults
t = load_model('triplet'); V.set_stop_aperture(t); V.set_pupil(t); print('set_pupil', t.optical_spec.pupil.value)
print('ALL DEFAULT-ENGINE PATHS OK')
