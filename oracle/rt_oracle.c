/*
 * rt_oracle.c -- CPU restatement of the reference's sequential real-ray trace.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for libb200rt.so:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load it.  Nothing under rayoptics_b200/ links, imports or calls it.
 *
 * It restates, scalar and one ray at a time, the algorithm of
 *   /root/reference/src/rayoptics/raytr/raytrace.py:19-38   bend, reflect
 *   /root/reference/src/rayoptics/raytr/raytrace.py:83-264  trace_raw
 *   /root/reference/src/rayoptics/elem/profiles.py:155-186  intersect_spencer
 *   /root/reference/src/rayoptics/elem/profiles.py:310-362  Spherical
 *   /root/reference/src/rayoptics/elem/profiles.py:569-609  Conic
 *   /root/reference/src/rayoptics/elem/profiles.py:849-885  EvenPolynomial
 *   /root/reference/src/rayoptics/elem/profiles.py:1070-1113 RadialPolynomial
 *   /root/reference/src/rayoptics/elem/profiles.py:1317-1369,1429-1437 Y/XToroid
 *   /root/reference/src/rayoptics/elem/surface.py:198-208,416-419,453-457 point_inside
 *   /root/reference/src/rayoptics/seq/interface.py:113-122  default point_inside
 *   /root/reference/src/rayoptics/util/misc_math.py:48-54   normalize
 * and, for grids, the start-ray generation of
 *   /root/reference/src/rayoptics/raytr/opticalspec.py:289-366,1339-1353
 *   /root/reference/src/rayoptics/raytr/trace.py:289-308
 *
 * Arithmetic contract (pinned by tests/test_oracle_golden.py
 * against vectors produced by the reference itself in tests/golden/):
 *   - numpy's 3-vector dot / linalg.norm are the FMA chain
 *       dot3(a,b) = fma(a2,b2, fma(a1,b1, a0*b0))
 *   - every other expression is unfused IEEE binary64 in source order;
 *   - sqrt and / are correctly rounded.
 * Build with -ffp-contract=off (see Makefile); fma() below is the only fusion.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/b200rt.h"

#define ST_OK RT_RAY_OK
#define ST_MISS RT_RAY_MISSED
#define ST_TIR RT_RAY_TIR
#define ST_BLOCKED RT_RAY_BLOCKED
#define ST_NUMERIC RT_RAY_NUMERIC

static inline double dot3(const double a[3], const double b[3])
{
    /* OpenBLAS ddot for n=3 as seen from numpy: sequential FMA accumulation */
    return fma(a[2], b[2], fma(a[1], b[1], a[0]*b[0]));
}

/* misc_math.normalize: v/norm(v), norm = sqrt(dot(v,v)); zero vector unchanged */
static inline void normalize3(const double v[3], double out[3])
{
    double len = sqrt(dot3(v, v));
    if (len == 0.0) {
        out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    } else {
        out[0] = v[0]/len; out[1] = v[1]/len; out[2] = v[2]/len;
    }
}

/* ---- quadric root shared by Spherical.intersect and Conic.intersect
 * s = cx2/(z_dir*sqrt(b*b - ax2*cx2) - b) with the reference's special cases */
static inline int quadric_root(double ax2, double cx2, double b, double z_dir, double *s)
{
    if (!(b == 0) || !(cx2 == 0) || !(ax2 == 0)) {
        double disc = b*b - ax2*cx2;
        if (disc < 0.0)
            return ST_MISS;                 /* math.sqrt ValueError */
        double den = z_dir*sqrt(disc) - b;
        if (den == 0.0 && cx2 != 0.0 && !isnan(cx2) && !isinf(cx2))
            *s = 0.0;                       /* FloatingPointError (divide) -> s = 0 */
        else
            *s = cx2/den;                   /* includes 0/0 -> nan ('invalid' is not raised) */
    } else {
        *s = 0.0;
    }
    return ST_OK;
}

/* ---- polynomial / toroid pieces ------------------------------------------*/

/* EvenPolynomial.sag (profiles.py:849-866) */
static int evenpoly_sag(const rt_surface_desc *S, double x, double y, double *z_tot)
{
    double r2 = x*x + y*y;
    double arg = 1. - S->ec*S->cv*S->cv*r2;
    if (arg < 0.0) return ST_MISS;
    double z = S->cv*r2/(1. + sqrt(arg));
    double z_asp = 0.0, r_pow = r2;
    for (int i = 0; i < S->n_coefs; i++) {
        z_asp += S->coefs[i]*r_pow;
        r_pow *= r2;
    }
    *z_tot = z + z_asp;
    return ST_OK;
}

/* EvenPolynomial.df (profiles.py:870-885) */
static int evenpoly_df(const rt_surface_desc *S, const double p[3], double g[3])
{
    double r2 = p[0]*p[0] + p[1]*p[1];
    double arg = 1. - S->ec*S->cv*S->cv*r2;
    if (arg < 0.0) return ST_NUMERIC;       /* uncaught ValueError in the reference */
    double sq = sqrt(arg);
    if (sq == 0.0) return ST_NUMERIC;       /* ZeroDivisionError */
    double e = S->cv/sq;
    double r_pow = 1.0, e_asp = 0.0, c_coef = 2.0;
    for (int i = 0; i < S->n_coefs; i++) {
        e_asp += c_coef*S->coefs[i]*r_pow;
        c_coef += 2.0;
        r_pow *= r2;
    }
    double e_tot = e + e_asp;
    g[0] = -e_tot*p[0]; g[1] = -e_tot*p[1]; g[2] = 1.0;
    return ST_OK;
}

/* RadialPolynomial.sag (profiles.py:1070-1088) */
static int radpoly_sag(const rt_surface_desc *S, double x, double y, double *z_tot)
{
    double r2 = x*x + y*y;
    double r = sqrt(r2);
    double arg = 1. - S->ec*S->cv*S->cv*r2;
    if (arg < 0.0) return ST_MISS;
    double z = S->cv*r2/(1. + sqrt(arg));
    double z_asp = 0.0, r_pow = r;
    for (int i = 0; i < S->n_coefs; i++) {
        z_asp += S->coefs[i]*r_pow;
        r_pow *= r;
    }
    *z_tot = z + z_asp;
    return ST_OK;
}

/* RadialPolynomial.df (profiles.py:1093-1113) */
static int radpoly_df(const rt_surface_desc *S, const double p[3], double g[3])
{
    double r2 = p[0]*p[0] + p[1]*p[1];
    double r = sqrt(r2);
    double arg = 1. - S->ec*S->cv*S->cv*r2;
    if (arg < 0.0) return ST_NUMERIC;
    double sq = sqrt(arg);
    if (sq == 0.0) return ST_NUMERIC;
    double e = S->cv/sq;
    double e_asp = 0.0;
    double r_pow = (r == 0.0) ? 1.0 : 1/r;
    double c_coef = 1.0;
    for (int i = 0; i < S->n_coefs; i++) {
        e_asp += c_coef*S->coefs[i]*r_pow;
        c_coef += 1.0;
        r_pow *= r;
    }
    double e_tot = e + e_asp;
    g[0] = -e_tot*p[0]; g[1] = -e_tot*p[1]; g[2] = 1.0;
    return ST_OK;
}

/* YToroid.fY (profiles.py:1329-1345) */
static int ytoroid_fY(const rt_surface_desc *S, double y, double *out)
{
    double y2 = y*y;
    double arg = 1. - S->ec*S->cv*S->cv*y2;
    if (arg < 0.0) return ST_MISS;
    double z = S->cv*y2/(1. + sqrt(arg));
    double z_asp = 0.0, y_pow = y2;
    for (int i = 0; i < S->n_coefs; i++) {
        z_asp += S->coefs[i]*y_pow;
        y_pow *= y2;
    }
    *out = z + z_asp;
    return ST_OK;
}

/* YToroid.f (profiles.py:1347-1349) */
static int ytoroid_f(const rt_surface_desc *S, const double p[3], double *f)
{
    double fY;
    int st = ytoroid_fY(S, p[1], &fY);
    if (st) return st;
    *f = p[2] - fY - S->cR*(p[0]*p[0] + p[2]*p[2] - fY*fY)/2;
    return ST_OK;
}

/* YToroid.df (profiles.py:1351-1369) */
static int ytoroid_df(const rt_surface_desc *S, const double p[3], double g[3])
{
    double y2 = p[1]*p[1];
    double arg = 1. - S->ec*S->cv*S->cv*y2;
    if (arg < 0.0) return ST_NUMERIC;
    double sq = sqrt(arg);
    if (sq == 0.0) return ST_NUMERIC;
    double e = S->cv/sq;
    double e_asp = 0.0, y_pow = 1.0, c_coef = 2.0;
    for (int i = 0; i < S->n_coefs; i++) {
        e_asp += c_coef*S->coefs[i]*y_pow;
        c_coef += 2.0;
        y_pow *= y2;
    }
    double dfdY = e + e_asp;
    double fY;
    int st = ytoroid_fY(S, p[1], &fY);
    if (st) return st;
    g[0] = -S->cR*p[0];
    g[1] = (S->cR*fY - 1)*(dfdY)*p[1];
    g[2] = 1 - S->cR*p[2];
    return ST_OK;
}

/* profile.f(p) for the iterated profiles */
static int prof_f(const rt_surface_desc *S, const double p[3], double *f)
{
    double z;
    int st;
    switch (S->profile) {
    case RT_PROFILE_EVENPOLY:
        st = evenpoly_sag(S, p[0], p[1], &z);
        if (st) return st;
        *f = p[2] - z;
        return ST_OK;
    case RT_PROFILE_RADIALPOLY:
        st = radpoly_sag(S, p[0], p[1], &z);
        if (st) return st;
        *f = p[2] - z;
        return ST_OK;
    case RT_PROFILE_YTOROID:
        return ytoroid_f(S, p, f);
    case RT_PROFILE_XTOROID: {
        double q[3] = {p[1], p[0], p[2]};   /* profiles.py:1432-1433 */
        return ytoroid_f(S, q, f);
    }
    }
    return ST_NUMERIC;
}

/* profile.df(p) for every profile */
static int prof_df(const rt_surface_desc *S, const double p[3], double g[3])
{
    switch (S->profile) {
    case RT_PROFILE_SPHERICAL:              /* profiles.py:360-362 */
        g[0] = -S->cv*p[0]; g[1] = -S->cv*p[1]; g[2] = 1.0 - S->cv*p[2];
        return ST_OK;
    case RT_PROFILE_CONIC:                  /* profiles.py:605-609 */
        g[0] = -S->cv*p[0]; g[1] = -S->cv*p[1]; g[2] = 1.0 - S->ec*S->cv*p[2];
        return ST_OK;
    case RT_PROFILE_EVENPOLY:
        return evenpoly_df(S, p, g);
    case RT_PROFILE_RADIALPOLY:
        return radpoly_df(S, p, g);
    case RT_PROFILE_YTOROID:
        return ytoroid_df(S, p, g);
    case RT_PROFILE_XTOROID: {              /* profiles.py:1435-1437 */
        double q[3] = {p[1], p[0], p[2]}, h[3];
        int st = ytoroid_df(S, q, h);
        if (st) return st;
        g[0] = h[1]; g[1] = h[0]; g[2] = h[2];
        return ST_OK;
    }
    }
    return ST_NUMERIC;
}

/* SurfaceProfile.intersect_spencer (profiles.py:155-186).  Note the returned
 * point is the last *evaluated* iterate, not p0 + s1*d. */
static int intersect_spencer(const rt_surface_desc *S, const double p0[3], const double d[3],
                             double eps, double *s_out, double p_out[3])
{
    double p[3] = {p0[0], p0[1], p0[2]};
    double f, g[3];
    int st = prof_f(S, p, &f);
    if (st) return st;
    st = prof_df(S, p, g);
    if (st) return st;
    double s1 = -f/dot3(d, g);
    double delta = fabs(s1);
    int iter = 0;
    while (delta > eps && iter < 1000) {
        p[0] = p0[0] + s1*d[0]; p[1] = p0[1] + s1*d[1]; p[2] = p0[2] + s1*d[2];
        st = prof_f(S, p, &f);
        if (st) return st;
        st = prof_df(S, p, g);
        if (st) return st;
        double s2 = s1 - f/dot3(d, g);
        delta = fabs(s2 - s1);
        s1 = s2;
        iter++;
    }
    *s_out = s1;
    p_out[0] = p[0]; p_out[1] = p[1]; p_out[2] = p[2];
    return ST_OK;
}

/* ifc.intersect(p, d, eps, z_dir) -> (s, p1) */
static int ifc_intersect(const rt_surface_desc *S, const double p[3], const double d[3],
                         double eps, double z_dir, double *s, double p1[3])
{
    int st;
    switch (S->profile) {
    case RT_PROFILE_THINLENS: {             /* oprops/thinlens.py:133-136 */
        double s1 = -p[2]/d[2];
        *s = s1;
        p1[0] = p[0] + s1*d[0]; p1[1] = p[1] + s1*d[1]; p1[2] = p[2] + s1*d[2];
        return ST_OK;
    }
    case RT_PROFILE_SPHERICAL: {            /* profiles.py:310-336 */
        double ax2 = S->cv;
        double cx2 = S->cv*dot3(p, p) - 2*p[2];
        double b = S->cv*dot3(d, p) - d[2];
        st = quadric_root(ax2, cx2, b, z_dir, s);
        if (st) return st;
        break;
    }
    case RT_PROFILE_CONIC: {                /* profiles.py:569-593 */
        double ax2 = S->cv*(1. + S->cc*d[2]*d[2]);
        double cx2 = S->cv*(p[0]*p[0] + p[1]*p[1] + S->ec*p[2]*p[2]) - 2.0*p[2];
        double b = S->cv*(d[0]*p[0] + d[1]*p[1] + S->ec*d[2]*p[2]) - d[2];
        st = quadric_root(ax2, cx2, b, z_dir, s);
        if (st) return st;
        break;
    }
    default:
        return intersect_spencer(S, p, d, eps, s, p1);
    }
    p1[0] = p[0] + (*s)*d[0]; p1[1] = p[1] + (*s)*d[1]; p1[2] = p[2] + (*s)*d[2];
    return ST_OK;
}

static int ifc_normal(const rt_surface_desc *S, const double p[3], double n[3])
{
    double g[3];
    if (S->profile == RT_PROFILE_THINLENS) {    /* oprops/thinlens.py:130-131 */
        n[0] = 0.; n[1] = 0.; n[2] = 1.;
        return ST_OK;
    }
    int st = prof_df(S, p, g);
    if (st) return st;
    normalize3(g, n);
    return ST_OK;
}

/* Surface.point_inside / Interface.point_inside */
static int point_inside(const rt_surface_desc *S, double x, double y, double fuzz)
{
    if (S->n_apertures > 0) {
        for (int k = 0; k < S->n_apertures; k++) {
            const rt_aperture_desc *A = &S->apertures[k];
            double xa = x - A->x_offset, ya = y - A->y_offset;  /* Aperture.tform, surface.py:391-394 */
            int ans;
            if (A->type == RT_APERTURE_CIRCULAR) {
                ans = sqrt(xa*xa + ya*ya) <= A->a + fuzz;
            } else if (A->type == RT_APERTURE_RECTANGULAR) {
                ans = (fabs(xa) <= A->a + fuzz) && (fabs(ya) <= A->b + fuzz);
            } else {
                return 0;   /* Elliptical has no point_inside: the base returns None */
            }
            if (A->is_obscuration) ans = !ans;
            if (!ans) return 0;
        }
        return 1;
    }
    return sqrt(x*x + y*y) <= S->max_aperture + fuzz;
}

static inline void apply_tfrm(const rt_surface_desc *S, const double p[3], const double d[3],
                              double bp[3], double bd[3])
{
    double q[3] = {p[0] - S->t[0], p[1] - S->t[1], p[2] - S->t[2]};
    if (S->has_tfrm == 2) {
        /* rt C-contiguous: OpenBLAS dgemv 't' path as seen from numpy,
         * y_i = fma(a_i2,v2, fma(a_i0,v0, a_i1*v1))  (probed, tools/probe_blas.py) */
        for (int r = 0; r < 3; r++) {
            const double *a = &S->rt[3*r];
            bp[r] = fma(a[2], q[2], fma(a[0], q[0], a[1]*q[1]));
            bd[r] = fma(a[2], d[2], fma(a[0], d[0], a[1]*d[1]));
        }
    } else if (S->has_tfrm) {
        /* rt Fortran-ordered (r.transpose() of a C array, elem/transform.py:86):
         * dgemv 'n' path, y_i = fma(a_i2,v2, fma(a_i1,v1, a_i0*v0)) */
        for (int r = 0; r < 3; r++) {
            bp[r] = dot3(&S->rt[3*r], q);
            bd[r] = dot3(&S->rt[3*r], d);
        }
    } else {
        bp[0] = q[0]; bp[1] = q[1]; bp[2] = q[2];
        bd[0] = d[0]; bd[1] = d[1]; bd[2] = d[2];
    }
}

/* HolographicElement.phase (oprops/doe.py:372-395); returns ST_OK or
 * RT_RAY_EVANESCENT where math.sqrt raises (raytrace.py:41-48) */
static int hoe_phase(const rt_surface_desc *S, const double pt[3], const double in_dir[3],
                     const double srf_nrml[3], double z_dir, double wvl, double out_dir[3])
{
    double normal[3], ref_dir[3], obj_dir[3], v[3];
    normalize3(srf_nrml, normal);
    for (int c = 0; c < 3; c++) v[c] = pt[c] - S->phase_ref_pt[c];
    normalize3(v, ref_dir);
    if (S->phase_flags & 1) for (int c = 0; c < 3; c++) ref_dir[c] = -ref_dir[c];
    double ref_cosI = dot3(ref_dir, normal);
    for (int c = 0; c < 3; c++) v[c] = pt[c] - S->phase_obj_pt[c];
    normalize3(v, obj_dir);
    if (S->phase_flags & 2) for (int c = 0; c < 3; c++) obj_dir[c] = -obj_dir[c];
    double obj_cosI = dot3(obj_dir, normal);
    double in_cosI = dot3(in_dir, normal);
    double mu = wvl/S->phase_ref_wl;
    double b = in_cosI + mu*(obj_cosI - ref_cosI);
    double refp_cosI = dot3(ref_dir, in_dir);
    double objp_cosI = dot3(obj_dir, in_dir);
    double ro_cosI = dot3(ref_dir, obj_dir);
    double c_ = mu*(mu*(1.0 - ro_cosI) + (objp_cosI - refp_cosI));
    double rad = b*b - 2*c_;
    if (rad < 0.0) return RT_RAY_EVANESCENT;
    double Q = -b + z_dir*sqrt(rad);
    for (int c = 0; c < 3; c++)
        out_dir[c] = in_dir[c] + mu*(obj_dir[c] - ref_dir[c]) + Q*normal[c];
    return ST_OK;
}

/* bend (raytrace.py:19-30) as called from DiffractiveElement.phase: math.sqrt of a
 * negative raises ValueError, which raytrace.phase() turns into the evanescent error */
static int bend_for_phase(const double d_in[3], const double normal[3], double n_in, double n_out,
                          double d_out[3])
{
    double normal_len = sqrt(dot3(normal, normal));
    double cosI = dot3(d_in, normal)/normal_len;
    double sinI_sqr = 1.0 - cosI*cosI;
    double arg = n_out*n_out - n_in*n_in*sinI_sqr;
    if (arg < 0.0) return RT_RAY_TIR;   /* bend() raises TraceTIRError itself (raytrace.py:28-30) */
    double n_cosIp = copysign(sqrt(arg), cosI);
    double alpha = n_cosIp - n_in*cosI;
    for (int c = 0; c < 3; c++) d_out[c] = (n_in*d_in[c] + alpha*normal[c])/n_out;
    return ST_OK;
}

/* np.cross for two 3-vectors: multiply, multiply, subtract (numpy/_core/numeric.py cross) */
static inline void cross3(const double a[3], const double b[3], double o[3])
{
    double t;
    o[0] = a[1]*b[2]; t = a[2]*b[1]; o[0] -= t;
    o[1] = a[2]*b[0]; t = a[0]*b[2]; o[1] -= t;
    o[2] = a[0]*b[1]; t = a[1]*b[0]; o[2] -= t;
}

/* DiffractionGrating.phase -> phase_ludwig (oprops/doe.py:119-172).  `x**2` on a Python /
 * numpy float is libm pow(x, 2.0): kept as pow() here so that this oracle, on the same
 * glibc, reproduces the reference bit for bit (pow(x, 2.0) != x*x for ~0.08 % of x).
 * np.sqrt of a negative gives NaN (no exception): the NaN direction propagates;
 * math.sqrt of a negative raises ValueError -> evanescent. */
static int grating_phase(const rt_surface_desc *S, const double in_dir[3], const double srf_nrml[3],
                         double z_dir, double wvl, double n_in, double n_out,
                         double out_dir[3], double *dW)
{
    const int reflect = S->mode == RT_MODE_REFLECT;
    const double refl = reflect ? -1.0 : 1.0;
    double normal[3], P[3], D[3], c[3];
    normalize3(srf_nrml, normal);
    for (int i = 0; i < 3; i++) normal[i] = z_dir*normal[i];
    cross3(S->phase_ref_pt, normal, P);
    cross3(normal, P, c);
    normalize3(c, D);
    const double spacing = S->phase_ref_wl;
    double mu = n_in/n_out;
    double T = refl*(wvl*S->phase_order)/(spacing*n_out);
    double in_cosI = dot3(in_dir, normal);
    double V = mu*in_cosI;
    double W = pow(mu, 2.0) - 1 + pow(T, 2.0) - 2*mu*T*dot3(D, in_dir);
    double result = sqrt(pow(V, 2.0) - W);          /* np.sqrt: NaN when negative */
    double Q1 = result - V, Q2 = -result - V, Q;
    if (!reflect) Q = (Q2 > Q1) ? Q2 : Q1;           /* max(Q1, Q2) */
    else Q = (Q2 < Q1) ? Q2 : Q1;                    /* min(Q1, Q2) */
    for (int i = 0; i < 3; i++) out_dir[i] = mu*in_dir[i] - T*D[i] + Q*normal[i];
    double a0 = 1 - pow(out_dir[0], 2.0) - pow(out_dir[1], 2.0);
    if (a0 < 0.0) return RT_RAY_EVANESCENT;          /* math.sqrt */
    out_dir[2] = copysign(sqrt(a0), out_dir[2]);
    double a1 = 1 - pow(in_cosI, 2.0);
    if (a1 < 0.0) return RT_RAY_EVANESCENT;
    double in_sinI = sqrt(a1);
    double out_cosI = dot3(out_dir, normal);
    double a2 = 1 - pow(out_cosI, 2.0);
    if (a2 < 0.0) return RT_RAY_EVANESCENT;
    double out_sinI = sqrt(a2);
    *dW = (spacing/wvl)*(n_in*in_sinI + refl*n_out*out_sinI);
    return ST_OK;
}

/* DiffractiveElement.phase with radial_phase_fct (oprops/doe.py:28-54,272-323) */
static int radial_doe_phase(const rt_surface_desc *S, const double pt[3], const double in_dir[3],
                            const double srf_nrml[3], double z_dir, double wvl, double n_in,
                            double n_out, double out_dir[3], double *dW_out)
{
    const double order = S->phase_order;
    double normal[3], inc_dir[3];
    normalize3(srf_nrml, normal);
    for (int i = 0; i < 3; i++) inc_dir[i] = in_dir[i];
    if (n_in != 1.0) {
        int st = bend_for_phase(in_dir, srf_nrml, n_in, 1.0, inc_dir);
        if (st) return st;
    }
    double in_cosI = dot3(inc_dir, normal);
    double mu = wvl/S->phase_ref_wl;
    const double x = pt[0], y = pt[1];
    double r_sqr = x*x + y*y;
    double dW = 0, dWdX = 0, dWdY = 0;
    for (int i = 0; i < S->n_phase_coefs; i++) {
        const double c = S->phase_coefs[i];
        dW += c*pow(r_sqr, (double)(i + 1));
        double r_exp = pow(r_sqr, (double)i);
        double factor = 2*(i + 1);
        dWdX += factor*c*x*r_exp;
        dWdY += factor*c*y*r_exp;
    }
    double b = in_cosI + order*mu*(normal[0]*dWdX + normal[1]*dWdY);
    double c_ = mu*(mu*(pow(dWdX, 2.0) + pow(dWdY, 2.0))/2 + order*(inc_dir[0]*dWdX + inc_dir[1]*dWdY));
    double rad = b*b - 2*c_;
    if (rad < 0.0) return RT_RAY_EVANESCENT;         /* math.sqrt */
    double Q = -b + z_dir*sqrt(rad);
    const double om = order*mu;
    const double g[3] = {om*dWdX, om*dWdY, om*0.0};
    for (int i = 0; i < 3; i++) out_dir[i] = inc_dir[i] + g[i] + Q*normal[i];
    dW *= mu;
    if (n_in != 1.0) {
        double t[3] = {out_dir[0], out_dir[1], out_dir[2]};
        int st = bend_for_phase(t, srf_nrml, 1.0, n_out, out_dir);
        if (st) return st;
    }
    *dW_out = dW;
    return ST_OK;
}

static inline void put_seg(double *ray, int k, const double p[3], const double d[3],
                           double dst, const double n[3])
{
    if (!ray) return;
    double *s = ray + (size_t)k*RT_SEG_DOUBLES;
    s[0] = p[0]; s[1] = p[1]; s[2] = p[2];
    s[3] = d[0]; s[4] = d[1]; s[5] = d[2];
    s[6] = dst;
    s[7] = n[0]; s[8] = n[1]; s[9] = n[2];
}

/* trace_raw (raytrace.py:83-264) for one ray.
 *  n_row[i]  : refractive index following interface i (path tuple Indx)
 *  ray       : [n_ifc][RT_SEG_DOUBLES] or NULL
 *  last      : [RT_SEG_DOUBLES] copy of ray[-1] or NULL                       */
int rto_trace_ray(const rt_surface_desc *surfs, int32_t n_ifc, const double *n_row, double wvl,
                  const double pt0[3], const double dir0[3], const rt_opts *o,
                  double *ray, double *last, int32_t *n_seg_out, double *op_out,
                  int32_t *status_out, int32_t *fail_surf_out)
{
    const double fuzz = (o->pt_inside_fuzz < 0.0) ? 1e-5 : o->pt_inside_fuzz;
    const int first_surf = o->first_surf;
    const int last_surf = o->last_surf;     /* <0: None */
    int n_seg = 0, status = ST_OK, fail_surf = -1;
    double opl = 0.0;
    double phs_sum = 0.0;     /* op_delta before `op_delta += opl` (raytrace.py:210,260) */
    double before_pt[3], before_dir[3], before_nrml[3];
    double inc_pt[3] = {0, 0, 0}, normal[3] = {0, 0, 1}, after_dir[3] = {0, 0, 0};
    double lseg[RT_SEG_DOUBLES];
    memset(lseg, 0, sizeof lseg);
    int b4_mode = RT_MODE_DUMMY;

    const rt_surface_desc *before = &surfs[0];
    if (o->intersect_obj) {
        double s;
        b4_mode = before->mode;
        /* raytrace.py:150: eps takes intersect()'s default 1e-12 */
        int st = ifc_intersect(before, pt0, dir0, 1.0e-12, (double)before->z_dir, &s, before_pt);
        if (!st) st = ifc_normal(before, before_pt, before_nrml);
        if (st) { status = st; fail_surf = 0; goto done; }
    } else {
        before_pt[0] = pt0[0]; before_pt[1] = pt0[1]; before_pt[2] = pt0[2];
        before_nrml[0] = 0.; before_nrml[1] = 0.; before_nrml[2] = 1.;
    }
    before_dir[0] = dir0[0]; before_dir[1] = dir0[1]; before_dir[2] = dir0[2];
    double z_dir_before = (double)before->z_dir;

    for (int surf = 1; surf < n_ifc; surf++) {
        const rt_surface_desc *ifc = &surfs[surf];
        double n_before = n_row[surf - 1];
        double b4_pt[3], b4_dir[3], pp_pt[3];
        apply_tfrm(before, before_pt, before_dir, b4_pt, b4_dir);
        double pp_dst = -dot3(b4_pt, b4_dir);
        pp_pt[0] = b4_pt[0] + pp_dst*b4_dir[0];
        pp_pt[1] = b4_pt[1] + pp_dst*b4_dir[1];
        pp_pt[2] = b4_pt[2] + pp_dst*b4_dir[2];

        double s;
        int st = ifc_intersect(ifc, pp_pt, b4_dir, o->eps, z_dir_before, &s, inc_pt);
        if (st) {
            /* TraceMissedSurfaceError, raytrace.py:231-237.  ST_NUMERIC (where the
             * reference dies with an uncaught ValueError) is packaged the same way. */
            put_seg(ray, n_seg, before_pt, before_dir, pp_dst, before_nrml);
            put_seg(lseg, 0, before_pt, before_dir, pp_dst, before_nrml);
            n_seg++;
            status = st; fail_surf = surf;
            goto done;
        }
        double dst_b4 = pp_dst + s;

        if (b4_mode == RT_MODE_PHANTOM && o->filter_out_phantoms && n_seg > 0) {
            if (ray) ray[(size_t)(n_seg - 1)*RT_SEG_DOUBLES + 6] += dst_b4;
            lseg[6] += dst_b4;
        } else {
            put_seg(ray, n_seg, before_pt, before_dir, dst_b4, before_nrml);
            put_seg(lseg, 0, before_pt, before_dir, dst_b4, before_nrml);
            n_seg++;
        }

        /* in_gap_range(surf-1), raytrace.py:123-132 */
        {
            int g = surf - 1, in_gap;
            if (first_surf == last_surf) in_gap = 0;
            else if (g < first_surf) in_gap = 0;
            else if (last_surf < 0) in_gap = 1;
            else in_gap = g < last_surf;
            if (in_gap) opl += n_before*dst_b4;
        }

        st = ifc_normal(ifc, inc_pt, normal);
        if (st) { status = st; fail_surf = surf; goto done; }

        if (o->check_apertures && surf >= first_surf && (last_surf < 0 || surf <= last_surf)
            && ifc->mode != RT_MODE_PHANTOM) {
            if (!point_inside(ifc, inc_pt[0], inc_pt[1], fuzz)) {
                /* raytrace.py:247-251 */
                const double zero = 0.0;
                put_seg(ray, n_seg, inc_pt, before_dir, zero, normal);
                put_seg(lseg, 0, inc_pt, before_dir, zero, normal);
                n_seg++;
                status = ST_BLOCKED; fail_surf = surf;
                goto done;
            }
        }

        if (ifc->phase_kind != RT_PHASE_NONE) {
            /* raytrace.py:205-210: the phase element sets after_dir (phs = 0 for a HOE) */
            double phs = 0.0;
            if (ifc->phase_kind == RT_PHASE_HOE)
                st = hoe_phase(ifc, inc_pt, b4_dir, normal, z_dir_before, wvl, after_dir);
            else if (ifc->phase_kind == RT_PHASE_GRATING)
                st = grating_phase(ifc, b4_dir, normal, z_dir_before, wvl, n_before, n_row[surf],
                                   after_dir, &phs);
            else
                st = radial_doe_phase(ifc, inc_pt, b4_dir, normal, z_dir_before, wvl, n_before,
                                      n_row[surf], after_dir, &phs);
            if (!st) phs_sum += phs;
            if (st) {
                /* TraceEvanescentRayError, raytrace.py:253-257 */
                put_seg(ray, n_seg, inc_pt, before_dir, 0.0, normal);
                put_seg(lseg, 0, inc_pt, before_dir, 0.0, normal);
                n_seg++;
                status = st; fail_surf = surf;
                goto done;
            }
        } else if (ifc->mode == RT_MODE_REFLECT) {
            /* reflect, raytrace.py:33-38 */
            double normal_len = sqrt(dot3(normal, normal));
            double cosI = dot3(b4_dir, normal)/normal_len;
            double k = 2.0*cosI;
            after_dir[0] = b4_dir[0] - k*normal[0];
            after_dir[1] = b4_dir[1] - k*normal[1];
            after_dir[2] = b4_dir[2] - k*normal[2];
        } else if (ifc->mode == RT_MODE_TRANSMIT) {
            /* bend, raytrace.py:19-30 */
            double n_in = n_before, n_out = n_row[surf];
            double normal_len = sqrt(dot3(normal, normal));
            double cosI = dot3(b4_dir, normal)/normal_len;
            double sinI_sqr = 1.0 - cosI*cosI;
            double arg = n_out*n_out - n_in*n_in*sinI_sqr;
            if (arg < 0.0) {
                /* raytrace.py:239-245 */
                put_seg(ray, n_seg, inc_pt, before_dir, 0.0, normal);
                put_seg(lseg, 0, inc_pt, before_dir, 0.0, normal);
                n_seg++;
                status = ST_TIR; fail_surf = surf;
                goto done;
            }
            double n_cosIp = copysign(sqrt(arg), cosI);
            double alpha = n_cosIp - n_in*cosI;
            after_dir[0] = (n_in*b4_dir[0] + alpha*normal[0])/n_out;
            after_dir[1] = (n_in*b4_dir[1] + alpha*normal[1])/n_out;
            after_dir[2] = (n_in*b4_dir[2] + alpha*normal[2])/n_out;
        } else {
            after_dir[0] = b4_dir[0]; after_dir[1] = b4_dir[1]; after_dir[2] = b4_dir[2];
        }

        for (int c = 0; c < 3; c++) {
            before_pt[c] = inc_pt[c];
            before_nrml[c] = normal[c];
            before_dir[c] = after_dir[c];
        }
        z_dir_before = (double)ifc->z_dir;
        b4_mode = ifc->mode;
        before = ifc;
    }
    /* StopIteration, raytrace.py:259-262 */
    if (n_ifc > 1) {
        put_seg(ray, n_seg, inc_pt, after_dir, 0.0, normal);
        put_seg(lseg, 0, inc_pt, after_dir, 0.0, normal);
        n_seg++;
    }

done:
    if (last) memcpy(last, lseg, sizeof lseg);
    *n_seg_out = n_seg;
    /* success: op_delta (sum of phases) += opl, raytrace.py:260; failure: ray_pkg carries opl */
    *op_out = (status == ST_OK) ? phs_sum + opl : opl;
    *status_out = status;
    *fail_surf_out = fail_surf;
    return 0;
}

/* ---- bundle: loop of rto_trace_ray over SoA inputs.  n_by_wvl: [n_wvl][n_ifc].
 * full: [n_ifc][10][stride] SoA (like rt_out.full) or NULL; last: [10][n_rays] SoA or NULL.
 * n_threads > 1 splits the rays over pthreads (the host-cores CPU baseline). */
typedef struct {
    const rt_surface_desc *surfs; int32_t n_ifc; const double *n_by_wvl; int64_t n_rays;
    const double *px, *py, *pz, *dx, *dy, *dz; const int32_t *wvl_idx; const rt_opts *o;
    const double *wvl_nm;
    double *last, *full; int64_t full_stride;
    double *op; int32_t *status, *fail_surf, *n_seg;
    int64_t r0, r1;
} bundle_job;

static void *bundle_worker(void *arg)
{
    bundle_job *J = (bundle_job *)arg;
    const int32_t n_ifc = J->n_ifc;
    double *ray = (double *)malloc(sizeof(double)*RT_SEG_DOUBLES*(size_t)n_ifc);
    double lseg[RT_SEG_DOUBLES];
    for (int64_t r = J->r0; r < J->r1; r++) {
        double p0[3] = {J->px[r], J->py[r], J->pz[r]}, d0[3] = {J->dx[r], J->dy[r], J->dz[r]};
        int w = J->wvl_idx ? J->wvl_idx[r] : J->o->wvl_idx;
        int32_t ns, st, fs;
        double opl;
        rto_trace_ray(J->surfs, n_ifc, J->n_by_wvl + (size_t)w*n_ifc, J->wvl_nm ? J->wvl_nm[w] : NAN,
                      p0, d0, J->o, J->full ? ray : NULL, lseg, &ns, &opl, &st, &fs);
        if (J->op) J->op[r] = opl;
        if (J->status) J->status[r] = st;
        if (J->fail_surf) J->fail_surf[r] = fs;
        if (J->n_seg) J->n_seg[r] = ns;
        if (J->last)
            for (int c = 0; c < RT_SEG_DOUBLES; c++) J->last[(size_t)c*J->n_rays + r] = lseg[c];
        if (J->full)
            for (int k = 0; k < ns; k++)
                for (int c = 0; c < RT_SEG_DOUBLES; c++)
                    J->full[((size_t)k*RT_SEG_DOUBLES + c)*J->full_stride + r] = ray[k*RT_SEG_DOUBLES + c];
    }
    free(ray);
    return NULL;
}

int rto_trace_bundle(const rt_surface_desc *surfs, int32_t n_ifc, const double *n_by_wvl,
                     int64_t n_rays,
                     const double *px, const double *py, const double *pz,
                     const double *dx, const double *dy, const double *dz,
                     const int32_t *wvl_idx, const rt_opts *o, const double *wvl_nm,
                     double *last, double *full, int64_t full_stride,
                     double *op, int32_t *status, int32_t *fail_surf, int32_t *n_seg,
                     int32_t n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    bundle_job jobs[256];
    pthread_t tids[256];
    int64_t per = (n_rays + n_threads - 1)/n_threads;
    for (int t = 0; t < n_threads; t++) {
        bundle_job J = {surfs, n_ifc, n_by_wvl, n_rays, px, py, pz, dx, dy, dz, wvl_idx, o, wvl_nm,
                        last, full, full_stride, op, status, fail_surf, n_seg, 0, 0};
        J.r0 = t*per; J.r1 = (t + 1)*per;
        if (J.r0 > n_rays) J.r0 = n_rays;
        if (J.r1 > n_rays) J.r1 = n_rays;
        jobs[t] = J;
    }
    if (n_threads == 1) { bundle_worker(&jobs[0]); return 0; }
    for (int t = 0; t < n_threads; t++) pthread_create(&tids[t], NULL, bundle_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(tids[t], NULL);
    return 0;
}

/* ---- grid start rays: Field.apply_vignetting (opticalspec.py:1339-1353) +
 * ray_start_from_osp 'epd' branch (opticalspec.py:354-366) + the virtual-object
 * direction flip of trace_base (trace.py:305-308).
 * Writes pt0/dir0 SoA for rays [ray_begin, ray_end) of the flattened grid. */
int rto_grid_start_rays(const rt_grid_spec *g, int64_t ray_begin, int64_t ray_end,
                        double *px, double *py, double *pz,
                        double *dx, double *dy, double *dz, int32_t *wvl_idx,
                        double *pupx, double *pupy)
{
    const int64_t per_tile = (int64_t)g->nx*g->ny;
    for (int64_t r = ray_begin; r < ray_end; r++) {
        int64_t tile = r/per_tile, loc = r - tile*per_tile;
        int f = (int)(tile/g->n_wvls), w = (int)(tile - (int64_t)f*g->n_wvls);
        int i = (int)(loc/g->ny), j = (int)(loc - (int64_t)i*g->ny);
        const rt_field_desc *F = &g->fields[f];
        double pup[2] = {g->pupil_x[(size_t)f*g->nx + i],
                         g->paired ? g->pupil_y[(size_t)f*g->nx + i] : g->pupil_y[(size_t)f*g->ny + j]};
        if (g->apply_vignetting) {
            if (pup[0] < 0.0) { if (F->vlx != 0.0) pup[0] *= (1.0 - F->vlx); }
            else              { if (F->vux != 0.0) pup[0] *= (1.0 - F->vux); }
            if (pup[1] < 0.0) { if (F->vly != 0.0) pup[1] *= (1.0 - F->vly); }
            else              { if (F->vuy != 0.0) pup[1] *= (1.0 - F->vuy); }
        }
        double d[3];
        if (g->pupil_kind == RT_PUPIL_EPD) {
            double pt1[3] = {g->eprad*pup[0] + F->aim[0], g->eprad*pup[1] + F->aim[1], g->z_pupil};
            double v[3] = {pt1[0] - F->pt0[0], pt1[1] - F->pt0[1], pt1[2] - F->pt0[2]};
            normalize3(v, d);
        } else if (g->pupil_kind == RT_PUPIL_WIDE) {
            /* wide-angle fields, opticalspec.py:342-358: pupil plane normal to the chief ray.
             * np.matmul of the C-contiguous 3x3 with a 3-vector: fma(a2,v2, fma(a0,v0, a1*v1)) */
            const double *a = F->rot;
            double vv[3] = {g->eprad*pup[0], g->eprad*pup[1], g->eprad*0.0}, pt1[3];
            for (int c = 0; c < 3; c++)
                pt1[c] = fma(a[3*c + 2], vv[2], fma(a[3*c], vv[0], a[3*c + 1]*vv[1]));
            pt1[2] -= F->obj2enp;
            double v[3] = {pt1[0] - F->pt0[0], pt1[1] - F->pt0[1], pt1[2] - F->pt0[2]};
            normalize3(v, d);
        } else {
            /* angular pupil, opticalspec.py:368-398: dir_tot = pupil_dir + cr_dir (aim = d0[:2]) */
            double pd[2];
            if (g->pupil_kind == RT_PUPIL_NA) {
                pd[0] = g->eprad*pup[0]; pd[1] = g->eprad*pup[1];
            } else {
                const double slope = g->eprad;
                double hypt = sqrt(1 + pow(pup[0]*slope, 2.0) + pow(pup[1]*slope, 2.0));
                pd[0] = slope*pup[0]/hypt; pd[1] = slope*pup[1]/hypt;
            }
            d[0] = pd[0] + F->aim[0]; d[1] = pd[1] + F->aim[1];
            d[2] = sqrt(1 - fma(d[1], d[1], d[0]*d[0]));     /* np.sqrt(1 - np.dot(dir_tot, dir_tot)) */
        }
        if (d[2]*(double)g->flip_z_dir < 0) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; }
        int64_t k = r - ray_begin;
        px[k] = F->pt0[0]; py[k] = F->pt0[1]; pz[k] = F->pt0[2];
        dx[k] = d[0]; dy[k] = d[1]; dz[k] = d[2];
        if (wvl_idx) wvl_idx[k] = g->wvl_idx[w];
        if (pupx) pupx[k] = pup[0];
        if (pupy) pupy[k] = pup[1];
    }
    return 0;
}

/* transverse aberration, analyses.py:561-580: p + (foc/d_z) d - image_pt */
int rto_transverse_abr(int64_t n, const double *px, const double *py,
                       const double *dx, const double *dy, const double *dz,
                       double foc, double ref_x, double ref_y, double *ax, double *ay)
{
    for (int64_t r = 0; r < n; r++) {
        double dist = foc/dz[r];
        ax[r] = (px[r] + dist*dx[r]) - ref_x;
        ay[r] = (py[r] + dist*dy[r]) - ref_y;
    }
    return 0;
}

/* ---- wave_abr_full_calc_finite_pup (raytr/waveabr.py:255-305) for an interface
 * k = -2 without decenter; eic_distance (waveabr.py:117-132).  W: the tile's
 * RT_WAVE_DOUBLES record (layout in include/b200rt.h).  `F**2` on a numpy float64
 * scalar is libm pow(F, 2.0) -- kept as pow() here so that this restatement is
 * bit-identical to the reference on the build box (tests/golden/vectors/<model>_opd.npz). */
static double eic_distance(const double p[3], const double d[3], const double p0[3], const double d0[3])
{
    double a[3] = {d[0] + d0[0], d[1] + d0[1], d[2] + d0[2]};
    double b[3] = {p[0] - p0[0], p[1] - p0[1], p[2] - p0[2]};
    return dot3(a, b)/(1. + dot3(d, d0));
}

/* wave_abr_full_calc_inf_ref (raytr/waveabr.py:356-420) for a last gap without tilt or
 * decenter (lcl_tfrm_last = (identity, [0, 0, tz])).  Record layout of this variant
 * (flag W[21] == 0): W[0:3] cr ray[1].p, W[3:6] cr ray[0].d, W[6:9] cr ray[-1].p,
 * W[9:12] cr ray[-1].d, W[12] V_BE = cr_op + op_cr_b4, W[13:16] image_pt,
 * W[17:20] d_cr_b4, W[20] tz, W[22] n_obj, W[23] n_img. */
static double wave_opd_inf_ref(const double *W, const double p1[3], const double d0[3],
                               const double pk[3], const double dk[3], const double pl[3],
                               const double dl[3], double ray_op)
{
    const double *cr_p1 = W, *cr_d0 = W + 3, *cr_pl = W + 6, *cr_dl = W + 9;
    const double V_BE = W[12], *image_pt = W + 13, *d_cr_b4 = W + 17, tz = W[20];
    const double n_obj = W[22], n_img = W[23];
    double e1 = eic_distance(p1, d0, cr_p1, cr_d0);
    const double p_b4[3] = {pk[0] - 0.0, pk[1] - 0.0, pk[2] - tz};
    const double *d_b4 = dk;
    const double mp[3] = {-p_b4[0], -p_b4[1], -p_b4[2]};
    double op_b4 = dot3(d_b4, mp);                       /* ray_dist_to_perp_from_origin */
    /* dist_to_shortest_join((cr[-1].p, cr[-1].d), (ray[-1].p, ray[-1].d)), waveabr.py:163-187 */
    double del_p[3], n[3], P1[3], P2[3];
    for (int c = 0; c < 3; c++) del_p[c] = pl[c] - cr_pl[c];
    cross3(cr_dl, dl, n);
    double nn = dot3(n, n);
    if (nn == 0) {
        double q[3] = {cr_pl[0] - pl[0], cr_pl[1] - pl[1], cr_pl[2] - pl[2]};
        double t2 = dot3(q, cr_dl)*dot3(cr_dl, dl);
        for (int c = 0; c < 3; c++) { P1[c] = cr_pl[c]; P2[c] = pl[c] + t2*dl[c]; }
    } else {
        double c2n[3], c1n[3];
        cross3(dl, n, c2n);
        cross3(cr_dl, n, c1n);
        double t1 = dot3(c2n, del_p)/nn, t2 = dot3(c1n, del_p)/nn;
        for (int c = 0; c < 3; c++) { P1[c] = cr_pl[c] + t1*cr_dl[c]; P2[c] = pl[c] + t2*dl[c]; }
    }
    double rF0[3], dd[3], ta[3], v[3];
    for (int c = 0; c < 3; c++) {
        rF0[c] = (P1[c] + P2[c])/2;
        dd[c] = d_b4[c] - d_cr_b4[c];
        ta[c] = pl[c] - image_pt[c];
    }
    double V_B = ray_op + op_b4;
    double W0 = V_B - V_BE + n_img*dot3(dd, rF0);
    double dbc = dot3(d_b4, d_cr_b4);
    for (int c = 0; c < 3; c++) v[c] = d_cr_b4[c] - d_b4[c]*dbc;
    double numer = dot3(v, ta);
    double denom = 1 + dot3(d_b4, d_cr_b4);
    double W_inf = W0 + n_img*numer/denom;
    return -n_obj*e1 - W_inf;
}

/* wave_abr_full_calc (raytr/waveabr.py:206-253): finite reference sphere (:255-305) or,
 * for a record flagged W[21] == 0, the infinite-reference variant.  pl, dl: ray[-1]. */
double rto_wave_opd(const double *W, const double p1[3], const double d0[3],
                    const double pk[3], const double dk[3], const double pl[3],
                    const double dl[3], double ray_op)
{
    if (W[21] == 0.0) return wave_opd_inf_ref(W, p1, d0, pk, dk, pl, dl, ray_op);
    const double *cr_p1 = W, *cr_d0 = W + 3, *cr_pk = W + 6, *cr_dk = W + 9;
    const double cr_op = W[12], *cr_exp_pt = W + 13, cr_exp_dist = W[16], *ref_dir = W + 17;
    const double R = W[20], sign_soln = W[21], n_obj = W[22], n_img = W[23];
    double e1 = eic_distance(p1, d0, cr_p1, cr_d0);
    double ekp = eic_distance(pk, dk, cr_pk, cr_dk);
    double dst = ekp - cr_exp_dist;
    double pc[3];
    for (int c = 0; c < 3; c++) pc[c] = (pk[c] - dst*dk[c]) - cr_exp_pt[c];
    double F = dot3(ref_dir, dk) - dot3(dk, pc)/R;
    double J = dot3(pc, pc)/R - 2.0*dot3(ref_dir, pc);
    double rad = pow(F, 2.0) + J/R;
    if (rad < 0.0) return NAN;              /* math.sqrt ValueError in the reference */
    double denom = F + sign_soln*sqrt(rad);
    double ep = (denom == 0) ? 0 : J/denom;
    return -n_obj*e1 - ray_op + n_img*ekp + cr_op - n_img*ep;
}

/* ---- whole grid on the host: start rays + trace + transverse aberration, split
 * over n_threads pthreads.  This is the CPU baseline / reference arm of bench.py
 * (the reference evaluates the same thing with a Python loop, trace.py:563-605).
 * Outputs are SoA over rays [ray_begin, ray_end): last [10][n], op, status,
 * fail_surf, abr_x, abr_y (any may be NULL). */
typedef struct {
    const rt_grid_spec *g; const rt_surface_desc *surfs; int32_t n_ifc; const double *n_by_wvl;
    const rt_opts *o; int64_t ray_begin, n; int64_t r0, r1;
    double *last, *op, *abr_x, *abr_y, *opd; int32_t *status, *fail_surf; const double *wvl_nm;
} grid_job;

static void *grid_worker(void *arg)
{
    grid_job *J = (grid_job *)arg;
    const rt_grid_spec *g = J->g;
    const int64_t per_tile = (int64_t)g->nx*g->ny;
    double *ray = J->opd ? (double *)malloc(sizeof(double)*RT_SEG_DOUBLES*(size_t)J->n_ifc) : NULL;
    for (int64_t r = J->r0; r < J->r1; r++) {
        double px, py, pz, dx, dy, dz;
        int32_t w;
        rto_grid_start_rays(g, r, r + 1, &px, &py, &pz, &dx, &dy, &dz, &w, NULL, NULL);
        double p0[3] = {px, py, pz}, d0[3] = {dx, dy, dz}, lseg[RT_SEG_DOUBLES], opl;
        int32_t ns, st, fs;
        rto_trace_ray(J->surfs, J->n_ifc, J->n_by_wvl + (size_t)w*J->n_ifc,
                      J->wvl_nm ? J->wvl_nm[w] : NAN, p0, d0, J->o, ray, lseg, &ns, &opl, &st, &fs);
        int64_t k = r - J->ray_begin;
        if (J->opd) {
            const double *s1 = ray + RT_SEG_DOUBLES, *sk = ray + (size_t)(J->n_ifc - 2)*RT_SEG_DOUBLES;
            J->opd[k] = (st == 0) ? rto_wave_opd(g->wave + (r/per_tile)*RT_WAVE_DOUBLES, s1, ray + 3,
                                                 sk, sk + 3, lseg, lseg + 3, opl)
                                  : NAN;
        }
        if (J->last) for (int c = 0; c < RT_SEG_DOUBLES; c++) J->last[(size_t)c*J->n + k] = lseg[c];
        if (J->op) J->op[k] = opl;
        if (J->status) J->status[k] = st;
        if (J->fail_surf) J->fail_surf[k] = fs;
        if (J->abr_x) {
            int64_t tile = r/per_tile;
            double rx = g->ref_img ? g->ref_img[tile*2] : 0.0, ry = g->ref_img ? g->ref_img[tile*2 + 1] : 0.0;
            double dist = g->foc/lseg[5];
            J->abr_x[k] = (lseg[0] + dist*lseg[3]) - rx;
            J->abr_y[k] = (lseg[1] + dist*lseg[4]) - ry;
        }
    }
    free(ray);
    return NULL;
}

int rto_trace_grid(const rt_grid_spec *g, const rt_surface_desc *surfs, int32_t n_ifc,
                   const double *n_by_wvl, int64_t ray_begin, int64_t ray_end, const rt_opts *o,
                   double *last, double *op, int32_t *status, int32_t *fail_surf,
                   double *abr_x, double *abr_y, double *opd, const double *wvl_nm, int32_t n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    grid_job jobs[256];
    pthread_t tids[256];
    int64_t n = ray_end - ray_begin, per = (n + n_threads - 1)/n_threads;
    for (int t = 0; t < n_threads; t++) {
        grid_job J = {g, surfs, n_ifc, n_by_wvl, o, ray_begin, n, 0, 0, last, op, abr_x, abr_y,
                      (opd && g->wave) ? opd : NULL, status, fail_surf, wvl_nm};
        J.r0 = ray_begin + t*per; J.r1 = J.r0 + per;
        if (J.r0 > ray_end) J.r0 = ray_end;
        if (J.r1 > ray_end) J.r1 = ray_end;
        jobs[t] = J;
    }
    if (n_threads == 1) { grid_worker(&jobs[0]); return 0; }
    for (int t = 0; t < n_threads; t++) pthread_create(&tids[t], NULL, grid_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(tids[t], NULL);
    return 0;
}

/* ---- the same grid trace on a PERSISTENT thread pool with dynamic scheduling: bench.py's CPU
 * arm.  rto_trace_grid() above creates and joins its threads on every call and splits the
 * rays statically, which under-uses the host (clipped rays end early: the blocks are uneven);
 * here the workers live across calls and pull blocks of `block` rays from a shared counter. */
typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t go, done;
    pthread_t *tids;
    int n_threads, generation, running, quit;
    grid_job job;                 /* r0/r1 unused: blocks come from `next` */
    int64_t ray_end, block;
    volatile int64_t next;
} rto_pool;

static rto_pool *g_pool = NULL;

static void *pool_worker(void *arg)
{
    rto_pool *P = (rto_pool *)arg;
    int seen = 0;
    for (;;) {
        pthread_mutex_lock(&P->mu);
        while (P->generation == seen && !P->quit) pthread_cond_wait(&P->go, &P->mu);
        if (P->quit) { pthread_mutex_unlock(&P->mu); return NULL; }
        seen = P->generation;
        pthread_mutex_unlock(&P->mu);
        for (;;) {
            int64_t b = __atomic_fetch_add(&P->next, P->block, __ATOMIC_RELAXED);
            if (b >= P->ray_end) break;
            grid_job J = P->job;
            J.r0 = b; J.r1 = b + P->block < P->ray_end ? b + P->block : P->ray_end;
            grid_worker(&J);
        }
        pthread_mutex_lock(&P->mu);
        if (--P->running == 0) pthread_cond_signal(&P->done);
        pthread_mutex_unlock(&P->mu);
    }
}

int rto_pool_destroy(void)
{
    rto_pool *P = g_pool;
    if (!P) return 0;
    pthread_mutex_lock(&P->mu);
    P->quit = 1;
    pthread_cond_broadcast(&P->go);
    pthread_mutex_unlock(&P->mu);
    for (int t = 0; t < P->n_threads; t++) pthread_join(P->tids[t], NULL);
    free(P->tids); free(P);
    g_pool = NULL;
    return 0;
}

int rto_pool_create(int32_t n_threads)
{
    if (g_pool && g_pool->n_threads == n_threads) return 0;
    rto_pool_destroy();
    if (n_threads < 1) n_threads = 1;
    rto_pool *P = (rto_pool *)calloc(1, sizeof *P);
    if (!P) return -1;
    pthread_mutex_init(&P->mu, NULL);
    pthread_cond_init(&P->go, NULL);
    pthread_cond_init(&P->done, NULL);
    P->n_threads = n_threads;
    P->tids = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) pthread_create(&P->tids[t], NULL, pool_worker, P);
    g_pool = P;
    return 0;
}

/* as rto_trace_grid, on the pool created by rto_pool_create(); block: rays per work item */
int rto_trace_grid_pool(const rt_grid_spec *g, const rt_surface_desc *surfs, int32_t n_ifc,
                        const double *n_by_wvl, int64_t ray_begin, int64_t ray_end, const rt_opts *o,
                        double *last, double *op, int32_t *status, int32_t *fail_surf,
                        double *abr_x, double *abr_y, double *opd, const double *wvl_nm, int64_t block)
{
    rto_pool *P = g_pool;
    if (!P) return -1;
    grid_job J = {g, surfs, n_ifc, n_by_wvl, o, ray_begin, ray_end - ray_begin, 0, 0, last, op, abr_x, abr_y,
                  (opd && g->wave) ? opd : NULL, status, fail_surf, wvl_nm};
    pthread_mutex_lock(&P->mu);
    P->job = J; P->ray_end = ray_end; P->block = block < 64 ? 64 : block;
    P->next = ray_begin;
    P->running = P->n_threads;
    P->generation++;
    pthread_cond_broadcast(&P->go);
    while (P->running > 0) pthread_cond_wait(&P->done, &P->mu);
    pthread_mutex_unlock(&P->mu);
    return 0;
}

int rto_sizeof_surface_desc(void) { return (int)sizeof(rt_surface_desc); }
