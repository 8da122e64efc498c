/*
 * tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE (see cuda_runtime.h here).
 * The per-ray device functions of rayoptics_b200/csrc compiled for the host and
 * driven ray by ray; what a CTA does around them (plan staging, store_result)
 * is restated in a few lines below.
 */
#define RT_HOSTSIM 1
#include "cuda_runtime.h"
#include "../../rayoptics_b200/csrc/rt_lean.cuh"
#include "../../rayoptics_b200/csrc/rt_grid.cuh"
#include <vector>

using namespace b200rt;

namespace {

struct Out {
    double *last, *op, *full;
    int32_t *status, *fail_surf, *n_seg;
    int64_t n;
};

void store(const Out &o, int64_t k, const RayResult &R, bool has_n)
{
    double *l = o.last + k;
    l[0] = R.p.x; l[o.n] = R.p.y; l[2*o.n] = R.p.z;
    l[3*o.n] = R.d.x; l[4*o.n] = R.d.y; l[5*o.n] = R.d.z;
    l[6*o.n] = has_n ? R.dst : 0.0;
    l[7*o.n] = has_n ? R.n.x : 0.0; l[8*o.n] = has_n ? R.n.y : 0.0; l[9*o.n] = has_n ? R.n.z : 0.0;
    o.op[k] = R.op; o.status[k] = R.status; o.fail_surf[k] = R.fail_surf; o.n_seg[k] = R.n_seg;
}

template <int OUT, bool POLY>
void run_lean(const rt_surface_desc *surfs, int n_ifc, const double *n_by_wvl, int n_wvl, int64_t n,
              const double *p0, const double *d0, const int32_t *wvl_idx, const rt_opts &o, const Out &out)
{
    std::vector<LeanSurf> ls(n_ifc);
    std::vector<LeanIdx> li((size_t)n_ifc*n_wvl);
    std::vector<LeanPoly> lp(n_ifc);
    build_plan(surfs, n_by_wvl, n_ifc, n_wvl, o, ls.data(), li.data());
    if (POLY) build_poly_plan(surfs, n_ifc, lp.data());
    for (int64_t r = 0; r < n; r++) {
        Vec3 p = {p0[r], p0[n + r], p0[2*n + r]}, d = {d0[r], d0[n + r], d0[2*n + r]};
        const int w = wvl_idx ? wvl_idx[r] : o.wvl_idx;
        FullWriter fw = {OUT == 2 ? out.full + r : nullptr, n};
        RayResult R;
        trace_ray_lean<OUT, false, POLY>(ls.data(), li.data() + (int64_t)w*n_ifc, lp.data(), surfs, n_ifc, o, p, d, fw, R);
        store(out, r, R, OUT >= 1);
    }
}

template <bool FULL>
void run_general(const rt_surface_desc *surfs, int n_ifc, const double *n_by_wvl, const double *wvls,
                 int64_t n, const double *p0, const double *d0, const int32_t *wvl_idx, const rt_opts &o,
                 const Out &out)
{
    for (int64_t r = 0; r < n; r++) {
        Vec3 p = {p0[r], p0[n + r], p0[2*n + r]}, d = {d0[r], d0[n + r], d0[2*n + r]};
        const int w = wvl_idx ? wvl_idx[r] : o.wvl_idx;
        FullWriter fw = {FULL ? out.full + r : nullptr, n};
        RayResult R;
        trace_ray<FULL>(surfs, n_by_wvl + (int64_t)w*n_ifc, wvls ? wvls[w] : 0.0, n_ifc, o, p, d, fw, R);
        store(out, r, R, true);
    }
}

}  // namespace

extern "C" {

/* kernel: 0 = general trace_ray, 1 = trace_ray_lean, 2 = trace_ray_lean with POLY.
 * out_kind (lean only): 0 p,d / 1 + normal,dst / 2 whole ray.  p0, d0: [3][n]. */
int hostsim_trace_bundle(const rt_surface_desc *surfs, int n_ifc, const double *n_by_wvl, int n_wvl,
                         const double *wvls, int64_t n, const double *p0, const double *d0,
                         const int32_t *wvl_idx, const rt_opts *o, int kernel, int out_kind,
                         double *last, double *op, int32_t *status, int32_t *fail_surf, int32_t *n_seg,
                         double *full)
{
    Out out = {last, op, full, status, fail_surf, n_seg, n};
    if (out_kind == 2 && !full) return -1;
    if (kernel == 0) {
        if (out_kind == 2) run_general<true>(surfs, n_ifc, n_by_wvl, wvls, n, p0, d0, wvl_idx, *o, out);
        else run_general<false>(surfs, n_ifc, n_by_wvl, wvls, n, p0, d0, wvl_idx, *o, out);
        return 0;
    }
#define LEAN(OUT) do { if (kernel == 2) run_lean<OUT, true>(surfs, n_ifc, n_by_wvl, n_wvl, n, p0, d0, wvl_idx, *o, out); \
                       else run_lean<OUT, false>(surfs, n_ifc, n_by_wvl, n_wvl, n, p0, d0, wvl_idx, *o, out); } while (0)
    if (out_kind == 0) LEAN(0);
    else if (out_kind == 1) LEAN(1);
    else LEAN(2);
#undef LEAN
    return 0;
}

/* building blocks against the plain IEEE operations (the CPU sibling of
 * rt_selftest_division): returns the number of mismatches where the fast flag is set */
int64_t hostsim_check_division(int64_t n, const double *a, const double *b, int64_t *n_fast)
{
    int64_t bad = 0, fast_cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        bool f;
        double q = quot_seq(a[i], b[i], rcp_refined(b[i]), f);
        if (f) { fast_cnt++; if (!(q == a[i]/b[i])) bad++; }
        double q2 = div_shared(a[i], b[i], rcp_refined(b[i])), q3 = a[i]/b[i];
        if (std::memcmp(&q2, &q3, 8) != 0 && !(q2 != q2 && q3 != q3)) bad++;
    }
    *n_fast = fast_cnt;
    return bad;
}

/* start rays of grid rays [r0, r1): grid_start_ray of rt_grid.cuh (lean: the lean kernels'
 * instance, 'epd' pupils only).  p, d: [3][r1 - r0]. */
int hostsim_grid_start_rays(const rt_grid_spec *g, int64_t r0, int64_t r1, int lean, double *p, double *d)
{
    GridDev G;
    G.n_wvls = g->n_wvls; G.nx = g->nx; G.ny = g->ny;
    G.apply_vignetting = g->apply_vignetting; G.flip_z_dir = g->flip_z_dir; G.paired = g->paired;
    G.eprad = g->eprad; G.z_pupil = g->z_pupil; G.foc = g->foc;
    G.fields = g->fields; G.wvl_idx = g->wvl_idx;
    G.pupil_x = g->pupil_x; G.pupil_y = g->pupil_y; G.ref_img = g->ref_img; G.wave = g->wave;
    G.rays_per_tile = (int64_t)g->nx*g->ny; G.chunks_per_tile = 0;
    if (lean && g->pupil_kind != RT_PUPIL_EPD) return -1;
    const int64_t n = r1 - r0;
    for (int64_t r = r0; r < r1; r++) {
        const int64_t tile = r/G.rays_per_tile, loc = r - tile*G.rays_per_tile;
        const int f = (int)(tile/G.n_wvls);
        Vec3 p0, d0;
        if (lean) grid_start_ray<true>(G, RT_PUPIL_EPD, f, loc, p0, d0);
        else grid_start_ray<false>(G, g->pupil_kind, f, loc, p0, d0);
        const int64_t k = r - r0;
        p[k] = p0.x; p[n + k] = p0.y; p[2*n + k] = p0.z;
        d[k] = d0.x; d[n + k] = d0.y; d[2*n + k] = d0.z;
    }
    return 0;
}

/* the OPD epilogue of the grid kernels (csrc/b200rt.cu grid_chunk_loop -> wave_opd) */
double hostsim_wave_opd(const double *W, const double *p1, const double *d0, const double *pk,
                        const double *dk, const double *pl, const double *dl, double ray_op)
{
    const Vec3 a = {p1[0], p1[1], p1[2]}, b = {d0[0], d0[1], d0[2]}, c = {pk[0], pk[1], pk[2]};
    const Vec3 d = {dk[0], dk[1], dk[2]}, e = {pl[0], pl[1], pl[2]}, f = {dl[0], dl[1], dl[2]};
    return wave_opd(W, a, b, c, d, e, f, ray_op);
}

int64_t hostsim_check_sqrt(int64_t n, const double *x, int64_t *n_fast)
{
    int64_t bad = 0, fast_cnt = 0;
    for (int64_t i = 0; i < n; i++) {
        bool f;
        double s = sqrt_seq(x[i], f);
        if (f) { fast_cnt++; if (!(s == std::sqrt(x[i]))) bad++; }
        double s1 = sqrt_near_one(x[i]);
        if (x[i] >= 0 && !(s1 == std::sqrt(x[i]))) bad++;
    }
    *n_fast = fast_cnt;
    return bad;
}

}
