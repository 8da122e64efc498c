/*
 * b200rt.h -- C ABI of libb200rt.so, the B200 (sm_100a) sequential real-ray
 * trace engine that replaces the hot path of mjhoptics/ray-optics:
 *
 *   rayoptics.raytr.raytrace.trace()/trace_raw()   src/rayoptics/raytr/raytrace.py:51-264
 *   and the per-ray loops stacked on it             src/rayoptics/raytr/trace.py:537-605,
 *                                                   src/rayoptics/raytr/analyses.py:212-230,437-455,666-696
 *
 * The reference is pure Python and has no FFI of its own; the binding a
 * maintainer would add is a ctypes stub.  Every entry
 * point below takes plain pointers and sizes -- no torch / numpy types.
 *
 * Conventions
 *   - all functions return RT_OK (0) or a negative rt_error code; the message
 *     for the calling thread is available from rt_last_error().
 *   - per-ray failures (missed surface, TIR, blocked by an aperture) are DATA
 *     (`status`, `fail_surf` arrays), never C errors.  They map 1:1 onto the
 *     reference's TraceError subclasses (src/rayoptics/raytr/traceerror.py:11-52).
 *   - pointers documented "DEVICE" must be device pointers on the table's
 *     device; pointers documented "HOST" are host pointers that are read
 *     before the call returns.
 *   - the caller owns every ray / result buffer; the library owns only the
 *     immutable table / grid handles.  Launches are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - all floating point is IEEE binary64; the arithmetic contract (which
 *     operations are fused) is stated in DESIGN.md and is what makes results
 *     bit-identical to the reference's numpy path.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 4
#define RT_MAX_COEFS 20     /* EvenPolynomial uses <=10, RadialPolynomial <=20 */
#define RT_MAX_PHASE_COEFS 10
#define RT_MAX_APERTURES 4  /* Surface.clear_apertures entries honoured per interface */
#define RT_SEG_DOUBLES 10   /* one ray segment = p[3], d[3], dst, nrml[3]  (raytr/__init__.py:36) */
#define RT_SUMMARY_DOUBLES 16
#define RT_WAVE_DOUBLES 24   /* per-(field, wvl) chief-ray / reference-sphere record, see rt_grid_spec.wave */

/* error codes (function return values) */
enum rt_error {
    RT_OK = 0,
    RT_ERR_INVALID = -1,     /* bad argument */
    RT_ERR_CUDA = -2,        /* CUDA runtime error, text in rt_last_error() */
    RT_ERR_UNSUPPORTED = -3, /* interface type not representable in the table */
    RT_ERR_NOMEM = -4
};

/* SurfaceProfile subclasses, src/rayoptics/elem/profiles.py */
enum rt_profile {
    RT_PROFILE_SPHERICAL = 0,  /* profiles.py:218-416 */
    RT_PROFILE_CONIC = 1,      /* profiles.py:449-680 */
    RT_PROFILE_EVENPOLY = 2,   /* profiles.py:682-889 */
    RT_PROFILE_RADIALPOLY = 3, /* profiles.py:891-1116 */
    RT_PROFILE_YTOROID = 4,    /* profiles.py:1119-1372 */
    RT_PROFILE_XTOROID = 5,    /* profiles.py:1375-1437 */
    RT_PROFILE_THINLENS = 6    /* oprops/thinlens.py:130-137: plane z=0, normal (0,0,1) */
};

/* phase elements (Interface.phase_element), src/rayoptics/oprops/doe.py */
enum rt_phase {
    RT_PHASE_NONE = 0,
    RT_PHASE_HOE = 1,          /* HolographicElement, doe.py:326-395 (also every ThinLens) */
    RT_PHASE_GRATING = 2,      /* DiffractionGrating.phase_ludwig, doe.py:57-172 */
    RT_PHASE_RADIAL = 3        /* DiffractiveElement with radial_phase_fct, doe.py:28-54,214-323 */
};

/* Interface.interact_mode, src/rayoptics/seq/interface.py:42-49 */
enum rt_mode {
    RT_MODE_TRANSMIT = 0,
    RT_MODE_REFLECT = 1,
    RT_MODE_DUMMY = 2,
    RT_MODE_PHANTOM = 3
};

/* per-ray status; 1..4 are the TraceError subclasses of traceerror.py */
enum rt_status {
    RT_RAY_OK = 0,
    RT_RAY_MISSED = 1,     /* TraceMissedSurfaceError */
    RT_RAY_TIR = 2,        /* TraceTIRError */
    RT_RAY_BLOCKED = 3,    /* TraceRayBlockedError */
    RT_RAY_EVANESCENT = 4, /* TraceEvanescentRayError (phase element: sqrt of a negative, raytrace.py:41-48) */
    RT_RAY_NUMERIC = 5     /* the reference would raise an uncaught ValueError /
                              ZeroDivisionError (sqrt of a negative in
                              EvenPolynomial.df, profiles.py:870-873) */
};

/* pupil specification after the image-space substitution of opticalspec.py:311-325 */
enum rt_pupil_kind {
    RT_PUPIL_EPD = 0,   /* spatial: aim at the entrance pupil plane, opticalspec.py:329-366.
                           rt_field_desc: pt0 = -(obj_dist + z_enp) [d0x/d0z, d0y/d0z, 0], aim = aim_pt */
    RT_PUPIL_NA = 1,    /* angular, object-space NA: dir_tot = sin_ang*pupil + cr_dir, :368-398.
                           rt_field_desc: pt0 = object point p0, aim = chief ray direction d0[:2] */
    RT_PUPIL_FNO = 2,   /* angular, object-space f/#: pupil_dir = slope*pupil/hypt, :378-384 */
    RT_PUPIL_WIDE = 3   /* spatial, wide-angle fields (fov.is_wide_angle), opticalspec.py:342-358: the pupil
                           plane is normal to the chief ray.  rt_field_desc: pt0 = start point, rot = the
                           matrix rot_v1_into_v2(d0, z) that takes the pupil plane into surface-1
                           coordinates, obj2enp = -(obj_dist + z_enp) with z_enp the field's real entrance
                           pupil position (fld.aim_info, raytr/wideangle.py).  Traced with
                           intersect_obj = 0 and without the virtual-object flip (trace.py:299-303). */
};

/* Aperture subclasses, src/rayoptics/elem/surface.py:340-494 */
enum rt_aperture_type {
    RT_APERTURE_CIRCULAR = 1,    /* surface.py:397-431 */
    RT_APERTURE_RECTANGULAR = 2, /* surface.py:434-469 */
    RT_APERTURE_ELLIPTICAL = 3   /* surface.py:472-494: no point_inside() -> always blocks */
};

typedef struct rt_aperture_desc {
    int32_t type;           /* rt_aperture_type */
    int32_t is_obscuration; /* Aperture.is_obscuration */
    double a;               /* Circular.radius | x_half_width */
    double b;               /* y_half_width (unused for Circular) */
    double x_offset;        /* Aperture.x_offset */
    double y_offset;        /* Aperture.y_offset */
} rt_aperture_desc;

/* One entry of SequentialModel.path(wvl): (Intfc, Gap, Tfrm, Indx, Zdir),
 * src/rayoptics/optical/model_constants.py:12, seq/sequential.py:149-202.
 * The refractive index lives in the separate n_by_wvl table. */
typedef struct rt_surface_desc {
    int32_t profile;      /* rt_profile */
    int32_t mode;         /* rt_mode */
    int32_t z_dir;        /* path tuple Zdir: +1 / -1 */
    int32_t n_coefs;      /* profile.max_nonzero_coef (polynomial profiles) */
    int32_t has_tfrm;     /* 0: Tfrm rotation is the identity (only `t` is applied);
                             1: rotation stored Fortran-ordered in numpy (r.transpose() of a C array,
                                elem/transform.py:86) -> y_i = fma(a_i2,v2, fma(a_i1,v1, a_i0*v0));
                             2: C-contiguous in numpy -> y_i = fma(a_i2,v2, fma(a_i0,v0, a_i1*v1))
                             (the two roundings numpy/OpenBLAS dgemv produce) */
    int32_t n_apertures;  /* len(ifc.clear_apertures), 0 -> max_aperture test */
    double cv;            /* profile.cv */
    double cc;            /* profile.cc */
    double ec;            /* profile.ec (= cc + 1.0 evaluated by the host) */
    double cR;            /* toroid sweep curvature */
    double max_aperture;  /* Interface.max_aperture */
    double coefs[RT_MAX_COEFS];
    double rt[9];         /* Tfrm[0], row-major: applied as rt . (p - t) on the way to the NEXT interface */
    double t[3];          /* Tfrm[1] */
    rt_aperture_desc apertures[RT_MAX_APERTURES];
    /* phase element (hasattr(ifc, 'phase_element'), raytrace.py:205-210) */
    int32_t phase_kind;   /* rt_phase */
    int32_t phase_flags;  /* HOE: bit 0 ref_virtual, bit 1 obj_virtual */
    double phase_ref_wl;  /* HOE / radial DOE: ref_wl (nm); grating: _grating_spacing_nm */
    double phase_ref_pt[3];   /* HOE: ref_pt; grating: grating_normal */
    double phase_obj_pt[3];   /* HOE: obj_pt */
    double phase_order;   /* grating / radial DOE: order */
    int32_t n_phase_coefs;    /* radial DOE: len(coefficients) */
    int32_t phase_pad;
    double phase_coefs[RT_MAX_PHASE_COEFS];   /* radial DOE: r**2, r**4, ... coefficients */
} rt_surface_desc;

/* keyword arguments of trace_raw(), raytrace.py:83-121 */
typedef struct rt_opts {
    double eps;                  /* eps=1e-12 */
    double pt_inside_fuzz;       /* pt_inside_fuzz; <0 means None -> the 1e-5 default of point_inside() */
    int32_t check_apertures;     /* check_apertures=False */
    int32_t intersect_obj;       /* intersect_obj=True */
    int32_t filter_out_phantoms; /* filter_out_phantoms=False */
    int32_t first_surf;          /* first_surf (trace() default 1, trace_raw() default 0) */
    int32_t last_surf;           /* last_surf; <0 means None */
    int32_t wvl_idx;             /* row of n_by_wvl used when the per-ray wvl_idx pointer is NULL */
} rt_opts;

/* Result buffers of a bundle / grid trace; every pointer is DEVICE and
 * optional (NULL = not wanted).  Arrays hold one element per ray. */
typedef struct rt_out {
    /* last ray segment, i.e. ray[-1] of the (possibly partial) RayPkg */
    double *px, *py, *pz;   /* RaySeg.p    */
    double *dx, *dy, *dz;   /* RaySeg.d    */
    double *nx, *ny, *nz;   /* RaySeg.nrml */
    double *dst;            /* RaySeg.dst (0 unless the ray missed: then pp_dst, raytrace.py:232) */
    double *op;             /* op_delta on success, opl so far on failure (raytrace.py:236,261) */
    int32_t *status;        /* rt_status */
    int32_t *fail_surf;     /* TraceError.surf; -1 on success */
    int32_t *n_seg;         /* number of valid segments in `full` */
    /* whole ray, structure of arrays: full[(seg*RT_SEG_DOUBLES + c)*full_stride + ray],
     * c = 0..2 p, 3..5 d, 6 dst, 7..9 nrml; seg < n_ifc.  Segments >= n_seg are not written. */
    double *full;
    int64_t full_stride;    /* >= n_rays */
    /* transverse ray aberration at the image (grid traces only):
     * p + (foc/d_z) d - ref_img, analyses.py:561-580 */
    double *abr_x, *abr_y;
    /* optical path difference w.r.t. the chief ray on the reference sphere (grid traces
     * with rt_grid_spec.wave): wave_abr_full_calc_finite_pup, raytr/waveabr.py:255-305.
     * System units (mm); NaN for rays that do not reach the image. */
    double *opd;
    /* RT_OUT_* bits */
    int32_t flags;
    int32_t pad_;
} rt_out;

/* rt_out.flags */
#define RT_OUT_ABR_NAN_STATUS 1  /* grid traces: rays that do not reach the image get abr_x = quiet NaN
                                    whose low mantissa bits hold rt_status and abr_y = quiet NaN whose low
                                    bits hold fail_surf, so a consumer that only reads abr_x/abr_y (16 B/ray
                                    instead of 20) still gets both: status = isnan(x) ? bits(x) & 0xFFFF : 0 */
#define RT_NAN_PAYLOAD_BASE 0x7FF8000000000000ull

typedef struct rt_table rt_table;
typedef struct rt_grid rt_grid;

/* ---- table: the compiled form of SequentialModel.path() for all wavelengths */

/* surfs: HOST [n_ifc]; n_by_wvl: HOST [n_wvl][n_ifc] refractive index following
 * each interface (path tuple Indx, unsigned; seq/sequential.py:259-274). */
int rt_table_create(const rt_surface_desc *surfs, int32_t n_ifc,
                    const double *n_by_wvl, int32_t n_wvl,
                    int32_t device, rt_table **out);
int rt_table_destroy(rt_table *table);
int rt_table_dims(const rt_table *table, int32_t *n_ifc, int32_t *n_wvl, int32_t *device);
/* wavelengths (nm) of the rows of n_by_wvl: needed only by phase elements (mu = wvl/ref_wl,
 * doe.py:384).  wvl_nm: HOST [n_wvl]. */
int rt_table_set_wavelengths(rt_table *table, const double *wvl_nm);

/* ---- bundle trace: replaces a Python loop over rt.trace()/trace_raw()
 * (raytrace.py:51-264; callers raytr/trace.py:250,310; analyses.py:458-510).
 * px..dz: DEVICE [n_rays] start point / direction cosines in the object
 * interface's coordinates; wvl_idx: DEVICE [n_rays] or NULL. */
int rt_trace_bundle(const rt_table *table, int64_t n_rays,
                    const double *px, const double *py, const double *pz,
                    const double *dx, const double *dy, const double *dz,
                    const int32_t *wvl_idx, const rt_opts *opts,
                    const rt_out *out, void *stream);

/* ---- grid trace: replaces trace.trace_grid / analyses.trace_ray_grid /
 * trace_ray_list / trace_ray_fan (raytr/trace.py:537-605, analyses.py:212-230,
 * 437-455,666-696) including the start-ray generation of
 * OpticalSpecs.ray_start_from_osp 'epd' branch (raytr/opticalspec.py:289-366),
 * Field.apply_vignetting (opticalspec.py:1339-1353) and the refocus /
 * transverse-aberration step (analyses.py:561-580). */

typedef struct rt_field_desc {
    double pt0[3];   /* ray start point on the object interface (opticalspec.py:361) */
    double aim[2];   /* fld.aim_info: aim point on the paraxial entrance pupil (opticalspec.py:357) */
    double vlx, vux, vly, vuy; /* Field vignetting factors */
    double rot[9];   /* RT_PUPIL_WIDE: rot_v1_into_v2(d0, [0,0,1]), row-major (C-contiguous in numpy) */
    double obj2enp;  /* RT_PUPIL_WIDE: -(fod.obj_dist + z_enp) */
} rt_field_desc;

typedef struct rt_grid_spec {
    int32_t n_fields, n_wvls;  /* tiles = n_fields * n_wvls, ordered field-major */
    int32_t nx, ny;            /* pupil samples per tile: x outer, y inner (trace.py:572-604) */
    const rt_field_desc *fields; /* HOST [n_fields] */
    const int32_t *wvl_idx;    /* HOST [n_wvls] rows of the table's n_by_wvl */
    const double *pupil_x;     /* HOST [n_fields][nx] relative pupil x before vignetting */
    const double *pupil_y;     /* HOST [n_fields][ny] */
    const double *ref_img;     /* HOST [n_fields][n_wvls][2] reference image point (ref_sphere[0]) or NULL (=0;
                                  rt_grid_chief_ref() can fill it on the device afterwards) */
    const double *wave;        /* HOST [n_fields][n_wvls][RT_WAVE_DOUBLES] or NULL.  Chief ray and reference
                                  sphere of each tile, what wave_abr_full_calc_finite_pup reads
                                  (raytr/waveabr.py:255-305, 24-76, 79-113):
                                  0-2 cr.ray[1].p   3-5 cr.ray[0].d   6-8 cr.ray[-2].p  9-11 cr.ray[-2].d
                                  12 cr_op  13-15 cr_exp_pt  16 cr_exp_dist  17-19 ref_dir
                                  20 ref_sphere_radius  21 sign_soln (+1/-1)  22 |n_obj|  23 |n_img|
                                  A record with [21] == 0 selects wave_abr_full_calc_inf_ref (waveabr.py:356-420,
                                  exit pupil beyond 1e8; image gap without tilt/decenter):
                                  0-2 cr.ray[1].p  3-5 cr.ray[0].d  6-8 cr.ray[-1].p  9-11 cr.ray[-1].d
                                  12 V_BE = cr_op + op_cr_b4  13-15 image_pt  17-19 d_cr_b4  20 t_z of the
                                  image gap  22 |n_obj|  23 |n_img| */
    int32_t apply_vignetting;  /* trace_base(apply_vignetting=...) trace.py:289-292 */
    int32_t flip_z_dir;        /* seq_model.z_dir[0]: dir0 is negated when dir0.z*z_dir < 0 (trace.py:305-308) */
    int32_t paired;            /* 0: product grid pupil_x[i] x pupil_y[j]; 1: ray list -- ny must be 1 and
                                  pupil_y is [n_fields][nx]: ray i uses (pupil_x[i], pupil_y[i])
                                  (trace_ray_list / trace_ray_fan, analyses.py:212-230,437-455) */
    int32_t pupil_kind;        /* rt_pupil_kind: which branch of ray_start_from_osp generates the rays */
    double eprad;              /* RT_PUPIL_EPD: pupil_value/2 (opticalspec.py:340); RT_PUPIL_NA: sin_ang = NA/n
                                  (:373-377); RT_PUPIL_FNO: slope = -1/(2 f/#) (:378-380) */
    double z_pupil;            /* fod.obj_dist + z_enp: z of the aim plane (opticalspec.py:360) */
    double foc;                /* focus shift used for abr_x/abr_y (analyses.py:572) */
} rt_grid_spec;

int rt_grid_create(const rt_grid_spec *spec, int32_t device, rt_grid **out);
int rt_grid_destroy(rt_grid *grid);
/* Replace the contents of `grid` by another description of the SAME shape (n_fields, n_wvls,
 * nx, ny, paired, wave present or not): one asynchronous host->device copy on `stream` from
 * the handle's pinned staging block, no allocation.  `spec` is read before the call returns;
 * launches on `stream` issued afterwards see the new description (launches on other streams
 * must be ordered after it by the caller). */
int rt_grid_update(rt_grid *grid, const rt_grid_spec *spec, void *stream);
/* total rays and the chunk geometry used for sharding / summaries */
int rt_grid_dims(const rt_grid *grid, int64_t *n_rays, int64_t *n_chunks, int32_t *chunk_rays);

/* Trace chunks [chunk_begin, chunk_end) of the grid.  A chunk is `chunk_rays`
 * consecutive rays of one (field, wvl) tile (the last chunk of a tile may be
 * short); flattened ray index = ((f*n_wvls + w)*nx + i)*ny + j.  Per-ray
 * outputs are indexed by (flattened ray index - first ray of chunk_begin).
 * summary: DEVICE [n_fields*n_wvls][RT_SUMMARY_DOUBLES] or NULL; receives this
 * call's partial per-tile sums (deterministic order):
 *   0 n_ok 1 n_missed 2 n_tir 3 n_blocked 4 n_other
 *   5 sum_x 6 sum_y 7 sum_xx 8 sum_yy 9 sum_xy (of abr_x/abr_y over ok rays)
 *   10 min_x 11 max_x 12 min_y 13 max_y 14 sum_op 15 reserved
 * scratch: DEVICE, rt_grid_scratch_bytes() bytes, needed when summary != NULL. */
int64_t rt_grid_scratch_bytes(const rt_grid *grid, int64_t chunk_begin, int64_t chunk_end);
int rt_trace_grid(const rt_table *table, const rt_grid *grid,
                  int64_t chunk_begin, int64_t chunk_end,
                  const rt_opts *opts, const rt_out *out,
                  double *summary, void *scratch, void *stream);

/* Grid trace with the results delivered to HOST memory -- the data path of the grid analyses
 * (spot diagrams: seq/sequential.py:1058-1085 evaluated for every field) in one call.  Chunks
 * [chunk_begin, chunk_end) are traced in n_pieces launches alternating between two streams
 * owned by the grid handle; each piece's transverse aberrations (RT_OUT_ABR_NAN_STATUS coding)
 * are copied device -> host right behind its trace, so copies and traces overlap.
 * d_abr_x/y: DEVICE staging [rays of the range]; h_abr_x/y: HOST, page-locked, same length;
 * summary: DEVICE [n_tiles][RT_SUMMARY_DOUBLES] or NULL; scratch: DEVICE,
 * rt_trace_grid_to_host_scratch_bytes(grid, n_pieces) bytes.  Work already queued on `stream`
 * (rt_grid_update, rt_grid_chief_ref) is waited for; when the call returns, `stream` waits for
 * all pieces: synchronising `stream` makes h_abr_* and summary valid. */
int64_t rt_trace_grid_to_host_scratch_bytes(const rt_grid *grid, int32_t n_pieces);
int rt_trace_grid_to_host(const rt_table *table, rt_grid *grid, int64_t chunk_begin, int64_t chunk_end,
                          const rt_opts *opts, double *d_abr_x, double *d_abr_y,
                          double *h_abr_x, double *h_abr_y, double *summary, void *scratch,
                          int32_t n_pieces, void *stream);

/* Reference image points without a host round trip: trace the (0, 0) pupil ray of every
 * field at row `wvl_idx` of the table (no vignetting, apertures not checked) and store its
 * image intercept (x, y) as the reference image point of every (field, wvl) tile of `grid` --
 * ref_sphere[0] of calculate_reference_sphere for image_pt_2d=None (raytr/waveabr.py:24-76,
 * raytr/trace.py:627-687 without the re-aiming).  One small launch on `stream`; later
 * rt_trace_grid calls on the same stream see the new points.  ref_out: DEVICE [n_fields][2]
 * or NULL, receives a copy. */
int rt_grid_chief_ref(const rt_table *table, rt_grid *grid, int32_t wvl_idx,
                      double *ref_out, void *stream);

/* Combine n_parts partial summaries (chunk ranges of one grid, or the ranks' rows of an
 * all-gather): parts DEVICE [n_parts][n_tiles][RT_SUMMARY_DOUBLES] -> out DEVICE
 * [n_tiles][RT_SUMMARY_DOUBLES]; sums add in part order, the min / max columns take
 * min / max.  One launch on `stream`. */
int rt_combine_summaries(const double *parts, int32_t n_parts, int64_t n_tiles,
                         double *out, void *stream);

/* ---- misc */
const char *rt_last_error(void);
int rt_abi_version(void);
/* rays per chunk (= threads per CTA of this build): the unit of rt_trace_grid's chunk ranges */
int32_t rt_chunk_rays(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t rt_launch_count(void);
/* fp64 vector-pipe peak of `device` in TFLOP/s, measured with a chain of
 * independent DFMAs (the roofline denominator of the register-resident trace) */
int rt_measure_fp64_peak(int32_t device, double *tflops);
/* cycles between two dependent fp64 FMAs of one warp (dependent-issue latency of the fp64 pipe) */
int rt_measure_fp64_latency(int32_t device, double *cycles_per_dependent_dfma);
/* self-test of the shared-reciprocal division of the specialised kernels against
 * the IEEE division (n_blocks x 256 threads x n_per_thread operand sets);
 * *mismatches must come back 0 */
int rt_selftest_division(int32_t device, int32_t n_blocks, int64_t n_per_thread, uint64_t seed,
                         uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
