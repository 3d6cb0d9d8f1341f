/*
 * rays_b200.h — C ABI of the B200-native ray-integration engine for ORNL-Fusion/RAYS.
 *
 * This is the drop-in boundary for the reference's hot path: everything lexically inside
 * `ray_loop` of `trace_rays` (RAYS_project/RAYS_lib/ray_tracing.f90:67-264).  The reference has
 * no FFI of its own (it is one Fortran program whose modules talk through module variables), so
 * each entry point below names the Fortran routine / module data it replaces.  A Fortran host
 * binds these with ISO_C_BINDING (see fortran/rays_b200_m.f90 and INTEGRATION.md).
 *
 * All structs are plain-old-data, laid out with natural C alignment (ISO_C_BINDING `bind(C)`
 * interoperable).  Arrays indexed by species are 0:nspec0 = 0..5 like the reference
 * (species_m.f90:21, electron = 0).  Pointers in the structs are HOST pointers unless the field
 * name starts with `d_`.  Matrices are stored the way the Fortran host holds them (column-major);
 * the C index expression is given next to each field.
 *
 * Nothing in this header refers to torch or to any C++ type.
 */
#ifndef RAYS_B200_H
#define RAYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAYS_NSPEC0 5                 /* species_m.f90:21  nspec0            */
#define RAYS_NSPECIES (RAYS_NSPEC0 + 1) /* arrays are 0:nspec0                 */
#define RAYS_NV_MAX 19                /* 7 + 1 + (1+nspec0) + 5, ode_m.f90:160-173 */
#define RAYS_FLAG_LEN 60              /* ray_results_m.f90 ray_stop_flag char*60 */

/* ---- selections that the reference makes with namelist strings + select case ---------- */
enum rays_equilib_model {             /* equilibrium_m.f90:177-195 */
    RAYS_EQ_SLAB = 1, RAYS_EQ_SOLOVEV = 2, RAYS_EQ_AXISYM_TOROID = 3, RAYS_EQ_MULTIPLE_MIRROR = 4
};
enum rays_ode_solver { RAYS_ODE_RK4 = 1, RAYS_ODE_SG = 2 };          /* ode_m.f90:232-245 */
enum rays_ray_deriv { RAYS_DERIV_COLD = 1, RAYS_DERIV_NUM = 2 };      /* eqn_ray.f90:106-123 */
enum rays_ray_param { RAYS_PARAM_ARCL = 1, RAYS_PARAM_TIME = 2 };     /* eqn_ray.f90:148-186 */
enum rays_damping_model { RAYS_DAMP_NONE = 0, RAYS_DAMP_FUND_ECH = 1 }; /* damping_m.f90:93-105 */
enum rays_wave_mode { RAYS_MODE_PLUS = 1, RAYS_MODE_MINUS = 2, RAYS_MODE_FAST = 3, RAYS_MODE_SLOW = 4 };
                                                                       /* dispersion_solvers_m.f90:66-79 */
/* profile models (slab_eq_m.f90:172-300, solovev_eq_m.f90:196-262, axisym_toroid_eq_m.f90:300-352,
 * multiple_mirror_eq_m.f90:290-368) */
enum rays_prof_model {
    RAYS_PROF_ZERO = 0, RAYS_PROF_CONSTANT = 1, RAYS_PROF_LINEAR = 2, RAYS_PROF_LINEAR_2 = 3,
    RAYS_PROF_PARABOLIC = 4, RAYS_PROF_GAUSSIAN = 5, RAYS_PROF_HYPERBOLIC = 6,
    RAYS_PROF_SPLINE = 7              /* 'density_spline_interp' / 'temperature_spline_interp' (axisym_toroid) */
};
enum rays_slab_b_model {              /* slab_eq_m.f90:172-215 */
    RAYS_SLAB_B_ZERO = 0, RAYS_SLAB_B_CONSTANT = 1, RAYS_SLAB_B_TOROID = 2,
    RAYS_SLAB_B_LINEAR_SHEAR = 3, RAYS_SLAB_B_LINEAR = 4, RAYS_SLAB_B_LINEAR_2 = 5
};
enum rays_axisym_magnetics { RAYS_MAG_SOLOVEV = 1, RAYS_MAG_EQDSK_SPLINE = 2 };  /* axisym_toroid_eq_m.f90:280-291 */

/* ---- per-ray stop codes <-> the reference's ode_stop_flag strings ------ */
enum rays_stop_code {
    RAYS_STOP_NONE = 0,
    RAYS_STOP_SOUT_GT_SMAX = 1,        /* 'sout > s_max'            ray_tracing.f90:144 */
    RAYS_STOP_NSTEP_MAX = 2,           /* ' nstep > nstep_max'      ray_tracing.f90:152 */
    RAYS_STOP_INFINITE_VG_RHS = 3,     /* 'infinite Vg'             eqn_ray.f90:142     */
    RAYS_STOP_RAY_STALLED = 4,         /* 'ray stalled'             eqn_ray.f90:168     */
    RAYS_STOP_DISP_RESIDUAL = 5,       /* 'dispersion_residual'     check_save.f90:70   */
    RAYS_STOP_INFINITE_VG_CHECK = 6,   /* 'infinite_Vg'             check_save.f90:108  */
    RAYS_STOP_TOTAL_ABSORPTION = 7,    /* 'total_absorption'        check_save.f90:123  */
    RAYS_STOP_ODE_TOTAL_ERROR = 8,     /* 'ODE total error'         SG_ode_m.f90:144    */
    RAYS_STOP_SG_MAXNUM = 9,           /* 'step number .ge. maxnum' ode_RAYS.f90:536    */
    RAYS_STOP_SG_STIFF = 10,           /* 'equations stiff'         ode_RAYS.f90:540    */
    RAYS_STOP_SG_T_EQ_TOUT = 11,       /* 't == tout'               ode_RAYS.f90:431    */
    RAYS_STOP_SG_BAD_TOL = 12,         /* 'relerr or abserr < 0'    ode_RAYS.f90:437    */
    RAYS_STOP_SG_EPS_LE_0 = 13,        /* 'eps <= 0'                ode_RAYS.f90:445    */
    RAYS_STOP_X_OUT_OF_BOUNDS = 20,    /* slab_eq_m.f90:163 */
    RAYS_STOP_Y_OUT_OF_BOUNDS = 21,    /* slab_eq_m.f90:164 */
    RAYS_STOP_Z_OUT_OF_BOUNDS = 22,    /* slab_eq_m.f90:165 */
    RAYS_STOP_R_OUT_OF_BOX_SOLOVEV = 23, /* 'R out_of_box'  solovev_eq_m.f90:155 */
    RAYS_STOP_Z_OUT_OF_BOX_SOLOVEV = 24, /* 'z out_of_box'  solovev_eq_m.f90:156 */
    RAYS_STOP_R_OUT_OF_BOX = 25,       /* 'R_out_of_box' axisym_toroid_eq_m.f90:261, multiple_mirror_eq_m.f90:270 */
    RAYS_STOP_Z_OUT_OF_BOX = 26,       /* 'Z_out_of_box' */
    RAYS_STOP_OUT_OF_PLASMA = 27,      /* 'out_of_plasma' */
    RAYS_STOP_R_OUT_OF_BOUNDS_SOLMAG = 28, /* 'R out_of_bounds' solovev_magnetics_m.f90:155 */
    RAYS_STOP_Z_OUT_OF_BOUNDS_SOLMAG = 29, /* 'z out_of_bounds' solovev_magnetics_m.f90:156 */
    RAYS_STOP_NEGATIVE_DENS = 30,      /* 'negative_dens' */
    RAYS_STOP_NEGATIVE_TEMP = 31,      /* 'negative_temp' */
    RAYS_STOP_CODE_MAX = 32
};

/* run-level error codes (return values; the reference would `stop 1`) */
enum rays_status {
    RAYS_OK = 0,
    RAYS_ERR_INVALID_CONFIG = 1,
    RAYS_ERR_CUDA = 2,
    RAYS_ERR_NOT_INITIALIZED = 3,
    RAYS_ERR_IM_DET = 4,              /* deriv_num.f90:137-140 / check_save.f90:221-224 `stop 1` */
    RAYS_ERR_ALLOC = 5,
    RAYS_ERR_IO = 6
};

/* ---- spline tables (splines_lib/quick_cube_splines_m.f90 types) ------------------------ */
typedef struct rays_spline1d {        /* cube_spline_function_1D, uniform grid (ilinx = 1)   */
    int32_t nx;
    int32_t pad_;
    const double *x_grid;             /* [nx]                                                */
    const double *fspl;               /* Fortran fspl(4,nx): C index fspl[4*i + c]           */
} rays_spline1d;

typedef struct rays_spline2d {        /* cube_spline_function_2D, uniform grids              */
    int32_t nx, ny;
    const double *x_grid;             /* [nx]                                                */
    const double *y_grid;             /* [ny]                                                */
    const double *fspl;               /* Fortran fspl(4,4,nx,ny): C index fspl[((j*nx+i)*4+cy)*4+cx] */
} rays_spline2d;

/* ---- equilibrium model data (the selected model's module variables) -------------------- */
typedef struct rays_slab_eq {         /* slab_eq_m.f90:35-85 */
    double xmin, xmax, ymin, ymax, zmin, zmax;
    double rmaj, rmin, x0;
    int32_t bx_prof_model, by_prof_model, bz_prof_model, dens_prof_model;
    double bx0, by0, bz0, LBy_shear_scale, LBz_scale, dBzdx;
    double Ln_scale, dndx, alphan1, alphan2, n_min;
    int32_t t_prof_model[RAYS_NSPECIES];
    double LT_scale, dtdx;
    double alphat1[RAYS_NSPECIES], alphat2[RAYS_NSPECIES], T_min[RAYS_NSPECIES];
} rays_slab_eq;

typedef struct rays_solovev_eq {      /* solovev_eq_m.f90:17-45 (+ derived psiB :83-84) */
    double rmaj, kappa, bphi0, iota0, outer_bound, psiB;
    double inner_bound, vert_bound, r_Zmax;
    double box_rmin, box_rmax, box_zmin, box_zmax;
    int32_t dens_prof_model;
    int32_t t_prof_model[RAYS_NSPECIES];
    int32_t pad_;
    double alphan1, alphan2;
    double alphat1[RAYS_NSPECIES], alphat2[RAYS_NSPECIES];
} rays_solovev_eq;

typedef struct rays_axisym_eq {       /* axisym_toroid_eq_m.f90:56-94 + solovev_magnetics_m.f90:19-35 */
    int32_t magnetics_model, density_prof_model;
    int32_t temperature_prof_model[RAYS_NSPECIES];
    double r_axis, z_axis;
    double box_rmin, box_rmax, box_zmin, box_zmax;
    double inner_bound, outer_bound, upper_bound, lower_bound;
    double plasma_psi_limit, alphan1, alphan2, d_scrape_off, T_scrape_off;
    double alphat1[RAYS_NSPECIES], alphat2[RAYS_NSPECIES];
    /* solovev_magnetics_m module data */
    double sm_rmaj, sm_kappa, sm_bphi0, sm_iota0, sm_psiB;
    double sm_box_rmin, sm_box_rmax, sm_box_zmin, sm_box_zmax;
    /* density_spline_interp_m.f90:22-27, temperature_spline_interp_m.f90:20-23: profiles on psi_N in [0,1],
     * normalised to 1 on axis (ne_profile_N, Te_profileN, Ti_profileN); used when the model is RAYS_PROF_SPLINE */
    rays_spline1d ne_spline, Te_spline, Ti_spline;
    /* eqdsk_magnetics_spline_interp_m.f90:30-55 (magnetics_model = RAYS_MAG_EQDSK_SPLINE): Psi_profile on the g-file's
     * (R,Z) grid, shifted to 0 on the axis; T_profile = R*Bphi, splined on the R grid as the reference does (:184);
     * eq_psibound = PSIBOUND - PSIAXIS */
    rays_spline2d Psi_spline;
    rays_spline1d T_spline;
    double eq_psibound;
} rays_axisym_eq;

typedef struct rays_mirror_eq {       /* multiple_mirror_eq_m.f90:63-106 + mirror_magnetics_spline_interp_m.f90:32-41 */
    int32_t density_prof_model;
    int32_t temperature_prof_model[RAYS_NSPECIES];
    int32_t pad_;
    double box_rmax, box_zmin, box_zmax;
    double r_LUFS, z_LUFS, Aphi_LUFS;
    double plasma_AphiN_limit, alphan1, alphan2, AphiN0_d, delta_d, d_scrape_off, T_scrape_off;
    double alphat1[RAYS_NSPECIES], alphat2[RAYS_NSPECIES];
    double AphiN0_t[RAYS_NSPECIES], delta_t[RAYS_NSPECIES];
    rays_spline2d Br_spline, Bz_spline, Aphi_spline;
} rays_mirror_eq;

/* ---- the whole marshalled module state (SURVEY.md §8b "inputs to marshal") -------------- */
typedef struct rays_cfg {
    /* constants_m.f90:36-60 (values carry the reference's float32-literal rounding) */
    double clight, eps0;
    /* rf_m.f90:18-51 */
    double omgrf, k0, dispersion_resid_limit;
    int32_t ray_param, wave_mode, k0_sign;
    /* species_m.f90:25-80 (ms, qs already scaled by me, e; t0s in joule) */
    int32_t nspec;
    double qs[RAYS_NSPECIES], ms[RAYS_NSPECIES], n0s[RAYS_NSPECIES], t0s[RAYS_NSPECIES],
        eta[RAYS_NSPECIES];
    /* ode_m.f90:20,89-107 and SG_ode_m.f90:26-31 */
    int32_t ode_solver, ray_deriv, nv, nstep_max;
    double ds, s_max;
    double rel_err0, abs_err0, SG_error_limit;
    /* damping_m.f90:30-40 */
    int32_t damping_model, multi_spec_damping;
    double total_damping_limit;
    /* diagnostics_m.f90:99-101 */
    int32_t integrate_eq_gradients;
    /* equilibrium_m.f90:62-68 */
    int32_t equilib_model;
    rays_slab_eq slab;
    rays_solovev_eq solovev;
    rays_axisym_eq axisym;
    rays_mirror_eq mirror;
    /* zfunctions_m.f90:22-31  x_grid(2001), fsplRe(4,2001) */
    rays_spline1d zfun_re;
} rays_cfg;

/* ---- launch fan (ray_init_m.f90:47-53) ------------------------------------------------- */
typedef struct rays_fan {
    int64_t nray;
    const double *rvec0;              /* Fortran rvec0(3,nray):       C rvec0[3*iray + i]       */
    const double *rindex_vec0;        /* Fortran rindex_vec0(3,nray): C rindex_vec0[3*iray + i] */
    const double *ray_pwr_wt;         /* [nray]                                                 */
} rays_fan;

/* ---- results (ray_results_m.f90:44-58); caller-owned, reference layout ------------------ */
typedef struct rays_results {
    int64_t nray;
    int32_t nv;
    int32_t npoints_alloc;            /* = nstep_max+1 : 2nd extent of ray_vec / 1st of residual */
    double *ray_vec;                  /* Fortran ray_vec(nv,npoints_alloc,nray): C [(iray*npoints_alloc+ip)*nv+iv]; may be NULL */
    double *residual;                 /* Fortran residual(npoints_alloc,nray): C [iray*npoints_alloc+ip]; may be NULL */
    int32_t *npoints;                 /* [nray] */
    int32_t *ray_stop_code;           /* [nray] enum rays_stop_code */
    char *ray_stop_flag;              /* [nray][60] blank padded, no NUL; may be NULL */
    double *initial_ray_power;        /* [nray] */
    double *ray_trace_time;           /* [nray] filled with total_trace_time/nray (no per-ray GPU time) */
    double *end_residuals;            /* [nray] */
    double *max_residuals;            /* [nray] */
    double *end_ray_parameter;        /* [nray] */
    double *start_ray_vec;            /* Fortran (nv,nray) */
    double *end_ray_vec;              /* Fortran (nv,nray) */
    double total_trace_time;          /* seconds, device time of the trace (out) */
    int64_t total_ray_steps;          /* sum(npoints-1) (out) */
} rays_results;

/* ---- deposition profile (post_process_lib/deposition_profiles_m.f90:228-292) ------------ */
typedef struct rays_deposition {
    int32_t n_bins;
    int32_t pad_;
    double grid_min, grid_max;
    double *profile;                  /* [n_bins] out */
    double Q_sum;                     /* out */
} rays_deposition;

/* ======================= life cycle ====================================================== */
/* Select the CUDA device this process drives (one process per GPU).  Replaces
 * initialize_openmp_m (openmp_m.f90:39-71).  Fails (RAYS_ERR_CUDA) when no GPU is present:
 * there is no CPU fallback. */
int rays_b200_init(int device);
int rays_b200_finalize(void);
const char *rays_b200_last_error(void);
/* blank-padded reference string for a stop code -> buf[len] (no NUL), SURVEY.md A.4 */
int rays_b200_stop_string(int code, char *buf, int len);

/* Upload config + spline tables to the device (module state -> constant/global memory).
 * Replaces the implicit "module variables are visible to trace_rays". */
int rays_b200_set_config(const rays_cfg *cfg);

/* ======================= the hot path ==================================================== */
/* trace_rays (ray_tracing.f90:1-290) with HOST buffers: H2D of the fan, trace, D2H of results
 * in the reference's layout.  This is what the Fortran `trace_rays` replacement calls. */
int rays_b200_trace(const rays_cfg *cfg, const rays_fan *fan, rays_results *res);

/* Device-resident variant: the fan is already in HBM (see rays_b200_fan_upload /
 * rays_b200_launch_fan_*), results stay in HBM until rays_b200_results_download. */
int rays_b200_fan_upload(const rays_fan *fan);
int rays_b200_trace_device(int store_trajectories);
int rays_b200_results_download(rays_results *res);
/* device time (ms) of the last rays_b200_trace_device kernel, its launch count, ray-steps */
int rays_b200_last_trace_stats(double *kernel_ms, int64_t *ray_steps, int32_t *n_launches);
/* RHS evaluations of the last trace (SG accounting), the kernel specialisation that ran, its grid and
 * resident CTAs per SM (for profiles/ and bench.py) */
int rays_b200_last_trace_info(int64_t *rhs_evals, char *kernel_name, int name_len, int32_t *grid, int32_t *blocks_per_sm);
/* Kernel time of the last trace (ms, CUDA events): the first pass over the fan and the resume passes over the
 * rays the first pass suspended (time slicing packs the long rays into full warps), and the number of passes */
int rays_b200_last_trace_breakdown(double *first_pass_ms, double *resume_pass_ms, int32_t *n_passes);
/* sharding helper: keep rays iray with iray % world == rank */
int rays_b200_fan_shard(int rank, int world);

/* ======================= launch fans (row f1) ============================================ */
typedef struct rays_solovev_launch {  /* solovev_ray_init_nphi_ntheta_m.f90:17-35 (also used for axisym) */
    int32_t n_r_launch, n_theta_launch, n_rindex_theta, n_rindex_phi;
    double r_launch0, dr_launch, theta_launch0, dtheta_launch;
    double rindex_theta0, delta_rindex_theta, rindex_phi0, delta_rindex_phi;
} rays_solovev_launch;
typedef struct rays_axisym_launch {   /* axisym_toroid_ray_init_R_Z_nphi_ntheta_m.f90:18-34 */
    int32_t n_R_launch, n_Z_launch, n_rindex_theta, n_rindex_phi;
    double R_launch0, Z_launch0;
    double rindex_theta0, delta_rindex_theta, rindex_phi0, delta_rindex_phi;
} rays_axisym_launch;
typedef struct rays_slab_launch {     /* simple_slab_ray_init_m.f90:17-40 */
    int32_t n_x_launch, n_y_launch, n_z_launch, n_ky_launch, n_kz_launch, pad_;
    double x_launch0, dx_launch, y_launch0, dy_launch, z_launch0, dz_launch;
    double rindex_y0, delta_rindex_y0, rindex_z0, delta_rindex_z0;
} rays_slab_launch;
/* Build the fan on the device (equilibrium + dispersion root per candidate, evanescent
 * candidates dropped, order preserved).  *nray_out = surviving rays.  The fan stays in HBM
 * as the current fan; rays_b200_fan_download copies it to host arrays sized >= nray_out. */
int rays_b200_launch_fan_solovev(const rays_solovev_launch *p, int64_t *nray_out);
int rays_b200_launch_fan_axisym(const rays_axisym_launch *p, int64_t *nray_out);
int rays_b200_launch_fan_slab(const rays_slab_launch *p, int64_t *nray_out);
/* file_input_ray_init / one_ray_init_XYZ_n_direction: positions + directions -> n from n(theta) */
int rays_b200_launch_fan_directions(int64_t n_in, const double *rvec_in, const double *nvec_in,
                                    int64_t *nray_out);
int rays_b200_fan_download(double *rvec0, double *rindex_vec0, double *ray_pwr_wt);

/* ======================= deposition (row f2) ============================================= */
/* calculate_deposition_profiles (deposition_profiles_m.f90:228-260) for Ptotal_x (slab) or
 * Ptotal_psi (axisym_toroid) on the trajectories of the last device trace; fills dep->profile
 * (host).  If d_profile_out != NULL the per-GPU partial profile (n_bins doubles + Q_sum) is
 * also left at that DEVICE address for an NCCL reduce by the caller. */
int rays_b200_deposition(rays_deposition *dep, double *d_profile_out);
/* Reproducible sums (the reference sums bins over rays in ray order, deposition_profiles_m.f90:251).  The
 * device bins into 64-bit FIXED-POINT accumulators (per-CTA partial profiles in shared memory, added to one global vector
 * at the end of the trace kernel): integer addition is associative, so the profile is bitwise independent of the order in
 * which rays, CTAs and GPUs contribute.  One count is *unit = 2^(e-62) with 2^e > sum |ray_pwr_wt| of the fan as it was
 * uploaded / launched (rays_b200_fan_shard keeps that total, so every shard of one fan uses the same unit).
 * rays_b200_deposition_fixed returns this GPU's raw bins (acc_out: host, d_acc_out: device, n_bins int64 each, either may
 * be NULL) for an exact integer reduction across GPUs; profile = sum(acc) * unit.  A host that uploads pre-sharded fans
 * sets the whole fan's weight with rays_b200_deposition_set_total_weight before tracing. */
int rays_b200_deposition_fixed(rays_deposition *dep, int64_t *acc_out, void *d_acc_out, double *unit);
int rays_b200_deposition_set_total_weight(double total_weight);
/* Per-ray summaries of the last trace packed on the device for the gather across GPUs: d_out (DEVICE,
 * rows_capacity rows) receives one row of 6 + 2 nv doubles per ray: npoints, stop code, initial_ray_power, end_residuals,
 * max_residuals, end_ray_parameter, start_ray_vec(nv), end_ray_vec(nv) (ray_tracing.f90:252-260).  d_out == NULL only
 * reports *rows and *row_doubles. */
int rays_b200_summaries_pack(void *d_out, int64_t rows_capacity, int64_t *rows, int32_t *row_doubles);
/* Same, fused into the trace: bins while tracing, no trajectory storage needed. */
int rays_b200_trace_device_binned(int n_bins, double grid_min, double grid_max, int store_trajectories);

/* ======================= one host process, several GPUs ============ */
/* `program rays` is a single host process (RAYS_code/RAYS.f90:9-15) whose `call trace_rays` (ray_tracing.f90:62-64: one OpenMP
 * loop over all rays) has no notion of ranks, so the library itself drives the GPUs of the box: rays_b200_init_multi creates one
 * context + stream per GPU and an NCCL communicator over them (ncclCommInitAll; NCCL is dlopen'ed, a 1-GPU host needs none);
 * ngpu <= 0 takes every visible GPU.  rays_b200_trace_multi is rays_b200_trace over all of them: GPU g integrates rays
 * g, g + ngpu, ... (interleaved: ray length varies smoothly along the launch loops), no traffic during integration, results land
 * in the caller's arrays in fan order.  rays_b200_trace_multi_binned traces without trajectory storage (res->ray_vec =
 * res->residual = NULL), bins the deposition on every GPU while tracing and finishes with the two collectives of the north star:
 * ONE ncclReduce (int64 sum, exact) of the n_bins fixed-point bins to GPU 0 -> dep->profile, dep->Q_sum, and ONE ncclAllGather of
 * the packed per-ray summaries (rays_b200_summaries_pack rows), which rays_b200_multi_summaries exposes per GPU (device pointer:
 * ngpu blocks of *rows_per_gpu rows, block g = rays g, g + ngpu, ...). */
int rays_b200_init_multi(int ngpu);
int rays_b200_ngpu(void);
int rays_b200_trace_multi(const rays_cfg *cfg, const rays_fan *fan, rays_results *res);
int rays_b200_trace_multi_binned(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, rays_deposition *dep);
const double *rays_b200_multi_summaries(int gpu, int64_t *rows_per_gpu, int32_t *row_doubles);
int rays_b200_finalize_multi(void);

/* ======================= O-X mode conversion analysis (row f4) ============================ */
/* One record per ray: type OX_conv + the per-ray outcome flags of analyze_OX_conv
 * (post_process_lib/OX_conv_analysis_m.f90:32-47, 91-198). */
typedef struct rays_ox_conv {
    double x_max[3], k_max[3];        /* saved point of maximum alpha_e on the ray, k there (find_x_max_ray :202-252)  */
    double alpha_max;                 /* omega_pe^2/omega^2 at x_max                                                   */
    double x_cut[3];                  /* point on the O-mode cutoff surface next to x_max (find_x_cutoff_ray :256-311)  */
    double conv_coeff;                /* O-X conversion coefficient (OX_conv_coeff :315-407); 0 unless `converted`      */
    double nvecx_c[3], nvecy_c[3], nvecz_c[3];
    int32_t ray_number, step_number;  /* 1-based, as the reference counts                                             */
    int32_t found_max, found_cutoff, converted, iteration;
} rays_ox_conv;
/* analyze_OX_conv on the trajectories of the last device trace (one thread per ray); out[nray] on the HOST.
 * The reference's OX_conv_data is the sub-list with converted != 0 (conv_coeff > 1e-4), in ray order. */
int rays_b200_ox_conv_analysis(rays_ox_conv *out, int64_t *n_converted);

/* ======================= mirror coil fields (row f4) ====================================== */
/* One coil of the mirror coil set: coil_type + /coil_data_list/ + /current_data_list/
 * (mirror_magnetics_lib/mirror_magnetics_m.f90:62-75, 101-115).  The conductors are modelled as
 * n_r_layers x n_z_slices circular filaments (positions: :222-239). */
typedef struct rays_coil {
    double inner_radius, outer_radius, z_center, z_width;
    double I_coil;                    /* current per turn [A] */
    int64_t n_turns;                  /* integer(KIND=8) in the reference */
    int32_t n_r_layers, n_z_slices;
} rays_coil;
/* calculate_B_on_rz_grid (mirror_magnetics_m.f90:324-368): Br, Bz, Aphi of all coils on the uniform grid
 * r_grid(n_r) x z_grid(n_z), one thread per grid point, coils and filaments summed in the reference's order
 * (coil_Brz_field_1Amp :246-287, mirror_Brz_field :291-320, Brz_loop_scaled B_loop_m.f90:191-246, complete
 * elliptic integrals by Carlson's RF/RD complete_elliptic_int_m.f90:38-507).  Outputs are HOST arrays in the
 * order of the netCDF field file, [n_z][n_r] (= Fortran Br(n_r,n_z)); r_grid/z_grid may be NULL. */
int rays_b200_mirror_brz_grid(const rays_coil *coils, int32_t n_coils, int32_t n_r, double r_min, double r_max,
                              int32_t n_z, double z_min, double z_max, double *r_grid, double *z_grid,
                              double *Br, double *Bz, double *Aphi);

/* ======================= one-point probes for unit parity tests ========================== */
/* equilibrium(rvec, eq) (equilibrium_m.f90:135-272) at n points.
 * out[n][RAYS_EQ_OUT] = bvec3, gradbtensor9 (Fortran order (i,j)->[i+3j]), ns6, gradns18 ([i+3s]),
 * ts6, gradts18, bmag, gradbmag3, bunit3, gradbunit9, omgc6, omgp2 6, alpha6, gamma6; err[n] stop code */
#define RAYS_EQ_OUT 106
int rays_b200_probe_equilibrium(int64_t n, const double *rvec, double *out, int32_t *err);
/* eqn_ray (eqn_ray.f90:1-236): dvds[n][nv], stop[n] */
int rays_b200_probe_rhs(int64_t n, const double *v, double *dvds, int32_t *stop);
/* check_save (check_save.f90:1-161): resid[n], stop[n] */
int rays_b200_probe_check_save(int64_t n, const double *v, double *resid, int32_t *stop);

/* ======================= measurement helpers ============================================= */
/* DFMA-only microbenchmark: measured fp64 FMA throughput of this GPU in TFLOP/s (2 flop/FMA)
 * and the SM clock (MHz) NVML/driver reported under load (0 if unavailable). */
int rays_b200_fp64_peak(double *tflops, double *sm_mhz);
/* Self-test of the kernels' exact arithmetic helpers (correctly rounded reciprocal and square root without
 * range checks, quotient by reciprocal + FMA correction) against IEEE 1.0/d, sqrt(x), x/d on n pseudo-random
 * operand pairs on the device; mismatch[0..2] receive the number of differing results of each (must be 0),
 * mismatch[3] the count of the one documented exception: the reciprocal of a divisor whose 52-bit significand
 * is all ones is 1 ulp off (probability 2^-52 per reciprocal on real data). */
int rays_b200_selftest_arith(int64_t n, uint64_t seed, int64_t *mismatch);
/* page-locked host memory for the result arrays (allocate_ray_results, ray_results_m.f90:132-142):
 * lets the trajectory copy-out of rays_b200_trace run at PCIe rate; pageable arrays work too */
int rays_b200_host_alloc(void **p, size_t bytes);
int rays_b200_host_free(void *p);
/* the CUDA stream (cudaStream_t as void*) the library launches on, for external event timing */
void *rays_b200_stream(void);
int rays_b200_version(void);
/* sizeof() of the ABI structs in the order rays_cfg, rays_fan, rays_results, rays_deposition,
 * rays_solovev_launch, rays_axisym_launch, rays_slab_launch, rays_spline1d, rays_spline2d, rays_slab_eq,
 * rays_solovev_eq, rays_axisym_eq, rays_mirror_eq: lets a binding in another language check its layout */
int rays_b200_struct_sizes(int32_t *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* RAYS_B200_H */
