import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import ref_harness as rh, rt_oracle as ro
from rayoptics_b200 import model as M, table as T, _abi
import importlib.util
spec = importlib.util.spec_from_file_location('dblg', '/root/reference/src/rayoptics/raytr/tests/ag_dblgauss_s.py')
dblg = importlib.util.module_from_spec(spec); spec.loader.exec_module(dblg)
rows = [list(r) for r in dblg.ag_dblgauss]
sm = M.gen_sequence(rows, wvls=[656.3, 587.6, 486.1], ref_wvl=1, sd=30.0)
rng = np.random.default_rng(0)

def compare(sm, wvl, rays, **kw):
    path = rh.ref_path(sm, wvl)
    descs, ns = T.describe_path(sm.path(wvl))
    opts = _abi.make_opts(**kw)
    nbad = 0; stats = {}
    for pt0, dir0 in rays:
        a = rh.ref_trace(path, pt0, dir0, wvl, **kw)
        b = ro.trace_ray(descs, ns, pt0, dir0, opts)
        stats[a['status']] = stats.get(a['status'], 0) + 1
        ok = (a['status'] == b['status'] and (a['status'] == 0 or a['fail_surf'] == b['fail_surf'])
              and a['n_seg'] == b['n_seg'] and np.array_equal(a['ray'], b['ray'], equal_nan=True) and (a['op'] == b['op'] or (np.isnan(a['op']) and np.isnan(b['op']))))
        if not ok:
            nbad += 1
            if nbad < 4:
                print('MISMATCH', a['status'], b['status'], a['fail_surf'], b['fail_surf'], a['n_seg'], b['n_seg'], a['op'], b['op'])
                if a['n_seg'] == b['n_seg']:
                    print(np.argwhere(a['ray'] != b['ray'])[:5], (a['ray']-b['ray'])[a['ray'] != b['ray']][:5])
    print('bad', nbad, 'of', len(rays), 'status hist', stats)
    return nbad

def inf_rays(n, thi, epr, maxang):
    rays = []
    for _ in range(n):
        ang = np.deg2rad(rng.uniform(-maxang, maxang, 2))
        d0 = np.array([np.sin(ang[0])*np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[0])*np.cos(ang[1])])
        pt0 = -thi*np.array([d0[0]/d0[2], d0[1]/d0[2], 0.])
        pt1 = np.array([*(epr*rng.uniform(-1,1,2)), thi])
        d = (pt1-pt0); d = d/np.sqrt(d.dot(d))
        rays.append((pt0, d))
    return rays
thi = rows[0][1]
for kw in [dict(first_surf=1, last_surf=11), dict(first_surf=1, last_surf=11, check_apertures=True)]:
    compare(sm, 587.6, inf_rays(300, thi, 40., 25.), **kw)
    compare(sm, 486.1, inf_rays(100, thi, 30., 15.), **kw)

print('--- RC-like conic mirrors (reflect)')
ifcs = [M.Surface(profile=M.Spherical(0.0), interact_mode='dummy', max_aperture=1e10),
        M.Surface(profile=M.Conic(c=-0.013727272717823692, cc=-1.303115101278525), interact_mode='reflect', max_aperture=3.79),
        M.Surface(profile=M.Conic(c=-0.017998163738521585, cc=-13.602974376524408), interact_mode='reflect', max_aperture=1.65),
        M.Surface(profile=M.Spherical(0.0), interact_mode='dummy', max_aperture=0.6)]
gaps = [M.Gap(1e10), M.Gap(-22.00000027666667), M.Gap(30.0)]
rc = M.SequentialModel(ifcs, gaps, z_dir=[1,-1,1], wvlns=[550.0])
for kw in [dict(first_surf=1, last_surf=2), dict(first_surf=1, last_surf=2, check_apertures=True)]:
    compare(rc, 550.0, inf_rays(300, 1e10, 5.0, 1.0), **kw)

print('--- aspheres')
def asph_model(kind):
    ifcs = [M.Surface(profile=M.Spherical(0.0), interact_mode='dummy', max_aperture=1e10)]
    gaps = [M.Gap(1e10)]
    for i, r in enumerate(rows[1:-1]):
        cv = r[0]
        if kind == 'even':
            prf = M.EvenPolynomial(c=cv, cc=-0.5, coefs=[0.0, 1e-7*(i+1), -2e-10, 0, 1e-16])
        elif kind == 'radial':
            prf = M.RadialPolynomial(c=cv, ec=0.7, coefs=[0.0, 0.0, 1e-6, 1e-7*(i+1), -2e-9, 1e-11])
        elif kind == 'ytor':
            prf = M.YToroid(c=cv, cR=cv*0.9, cc=-0.3, coefs=[0.0, 1e-7])
        elif kind == 'xtor':
            prf = M.XToroid(c=cv, cR=cv*1.1, cc=0.2, coefs=[0.0, -1e-7])
        elif kind == 'conic':
            prf = M.Conic(c=cv, cc=0.3*(-1)**i)
        med = M.AbbeGlass(r[2], r[3]) if r[3] else M.Air()
        ifcs.append(M.Surface(profile=prf, max_aperture=22.0)); gaps.append(M.Gap(r[1], med))
    ifcs.append(M.Surface(profile=M.Spherical(0.0), interact_mode='dummy', max_aperture=50.))
    return M.SequentialModel(ifcs, gaps, wvlns=[587.6, 486.1])
for kind in ['conic', 'even', 'radial', 'ytor', 'xtor']:
    m = asph_model(kind)
    print(kind)
    compare(m, 587.6, inf_rays(200, 1e10, 28., 16.), first_surf=1, last_surf=11, check_apertures=True)
    compare(m, 486.1, inf_rays(100, 1e10, 20., 10.), first_surf=1, last_surf=11)

print('--- apertures / phantom / tfrm / intersect_obj')
m = asph_model('conic')
m.ifcs[3].clear_apertures = [M.Rectangular(15., 10., x_offset=1.0, y_offset=-0.5)]
m.ifcs[5].clear_apertures = [M.Circular(18.0), M.Circular(3.0, is_obscuration=True, x_offset=0.5)]
m.ifcs[6].interact_mode = 'phantom'
m.ifcs[8].clear_apertures = [M.Elliptical(30., 30.)]
compare(m, 587.6, inf_rays(300, 1e10, 20., 8.), first_surf=1, last_surf=11, check_apertures=True, filter_out_phantoms=True)
m.ifcs[8].clear_apertures = []
compare(m, 587.6, inf_rays(300, 1e10, 20., 8.), first_surf=2, last_surf=9, check_apertures=True, filter_out_phantoms=True, pt_inside_fuzz=1e-3)
compare(m, 587.6, inf_rays(100, 1e10, 20., 8.), filter_out_phantoms=True)
# finite object, intersect_obj False
rays = [(np.array([*rng.uniform(-5,5,2), 0.0]), (lambda v: v/np.sqrt(v.dot(v)))(np.array([*rng.uniform(-.1,.1,2), 1.0]))) for _ in range(100)]
m.gaps[0].thi = 200.0; m.update_model()
compare(m, 587.6, rays, intersect_obj=False, first_surf=1, last_surf=11, check_apertures=True)
compare(m, 587.6, rays, intersect_obj=True, first_surf=1, last_surf=11)
print('--- tilted / decentered')
def rot(rng):
    q = rng.normal(size=4)*np.array([1,.03,.03,.03]); q/=np.linalg.norm(q)
    w,x,y,z=q
    return np.array([[1-2*(y*y+z*z),2*(x*y-z*w),2*(x*z+y*w)],[2*(x*y+z*w),1-2*(x*x+z*z),2*(y*z-x*w)],[2*(x*z-y*w),2*(y*z+x*w),1-2*(x*x+y*y)]])
m = asph_model('conic')
tf = []
for i,g in enumerate(m.gaps):
    R = rot(rng)
    rt = R.T if i%2==0 else np.ascontiguousarray(R)   # F-ordered view / C array
    tf.append((rt, np.array([rng.normal()*0.05, rng.normal()*0.05, g.thi])))
tf.append((np.identity(3), np.zeros(3)))
m._tfrms_given = tf; m.update_model()
print([d.has_tfrm for d in T.describe_path(m.path(587.6))[0]])
compare(m, 587.6, inf_rays(300, 1e10, 15., 5.), first_surf=1, last_surf=11, check_apertures=True)
m2 = M.SequentialModel.from_dict(m.to_dict())
print([d.has_tfrm for d in T.describe_path(m2.path(587.6))[0]])
compare(m2, 587.6, inf_rays(100, 1e10, 15., 5.), first_surf=1, last_surf=11, check_apertures=True)
