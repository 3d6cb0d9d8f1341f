// physics.cuh — device-side physics of the RAYS hot path, one fp64 ray per thread (sm_100a).
//
// Everything a ray needs per right-hand-side evaluation lives in registers; the only memory
// traffic is constant-bank reads of the run configuration, read-only (LDG) lookups of the spline
// tables (mirror: 3 x 16 doubles per equilibrium call from L2-resident 1.3 MB tables; damping: one
// 4-double row of the 64 KB Z table) and the trajectory writeback.
//
// What is computed (reference file:line, all under RAYS_project/):
//   equilibrium + derived quantities      RAYS_lib/equilibrium_m.f90:135-272
//   slab / solovev / axisym_toroid+solovev_magnetics / multiple_mirror+spline models
//                                         RAYS_lib/slab_eq_m.f90:125-309, solovev_eq_m.f90:122-322,
//                                         axisym_toroid_eq_m.f90:215-362, solovev_magnetics_m.f90:124-253,
//                                         multiple_mirror_eq_m.f90:223-376, mirror_magnetics_spline_interp_m.f90:132-204
//   bicubic / cubic spline evaluation     splines_lib/bcspeval.f90:128-255,368-407, cspeval.f90:93-295
//   deriv_cold                            RAYS_lib/deriv_cold.f90:40-171
//   deriv_num + determ                    RAYS_lib/deriv_num.f90:40-153
//   residual (check_save)                 RAYS_lib/check_save.f90:163-235
//   damp_fund_ECH + Z-function lookup     RAYS_lib/damp_fund_ECH.f90:2-128, math_functions_lib/zfunctions_m.f90:351-432
//   cold dispersion roots (launch)        RAYS_lib/disp_solve_cold_n1sq_vs_n3.f90, disp_solve_cold_nsq_vs_theta.f90
//
// Design notes (B200-first, not a translation):
//  * The cold dielectric tensor is eps = [[S,-iD,0],[iD,S,0],[0,0,P]], so the reference's complex
//    3x3 determinant collapses to a real closed form (its imaginary part is exactly zero, hence the
//    reference's `stop 1` on |Im det| can never fire for the cold model); the 14 determinants of
//    deriv_num share one evaluation of (S,D,P) per equilibrium point.
//  * Species count is a template parameter (NS = 2 is the common electron+ion case) so every
//    species loop unrolls into registers; NS = 0 selects a run-time count (<= 6).
//  * Reciprocals that the Fortran recomputes (1/r, 1/bmag, 1/k0, 1/omega) are hoisted.
//  * IEEE semantics are kept where the reference relies on them: NaN from 0*0/0 outside the
//    Solov'ev plasma (deriv_cold.f90:86) must stop the ray with 'infinite_Vg' exactly as on the CPU.
#pragma once
#include <cfloat>
#include <cstdint>

#include "../../include/rays_b200.h"

namespace rays_dev {

struct DevCfg {
    rays_cfg c;            // table pointers inside are DEVICE pointers
    double inv_k0, inv_omgrf, inv_omgrf2;
    int pow_int_n[2];      // alphan1, alphan2 as small integers (0,1,2) if exactly so, else -1
};

#ifndef RAYS_DEV_CFG_DEFINED
#define RAYS_DEV_CFG_DEFINED
static __constant__ DevCfg g_dc;
#endif

#define RD_INLINE __device__ __forceinline__

template <int NS_> struct NSpec {
    static constexpr int MAX = NS_ > 0 ? NS_ : RAYS_NSPECIES;
    RD_INLINE static int n() { return NS_ > 0 ? NS_ : g_dc.c.nspec + 1; }
};

// pow with the exponents that occur in practice resolved to multiplications.  glibc's pow is
// correctly rounded for these (x**1 = x, x**2 = x*x, x**0 = 1), so this matches the CPU bit for bit.
RD_INLINE double pow_fast(double x, double a) {
    if (a == 1.0) return x;
    if (a == 0.0) return 1.0;
    if (a == 2.0) return x * x;
    return pow(x, a);
}

// parabolic_prof (axisym_toroid_eq_m.f90:505-521 and twins)
RD_INLINE void parabolic_prof(double rho, double f_min, double a1, double a2, double &f, double &fp) {
    f = 0.0;
    fp = 0.0;
    if (rho < 1.0) {
        const double base = 1.0 - pow_fast(rho, a2);
        f = pow_fast(base, a1);
        fp = -a1 * a2 * pow_fast(rho, a2 - 1.0) * pow_fast(base, a1 - 1.0);
    }
    if (f < f_min) { f = f_min; fp = 0.0; }
}
// hyperbolic_prof (multiple_mirror_eq_m.f90:486-505)
RD_INLINE void hyperbolic_prof(double rho, double f_min, double rho0, double delta, double &f, double &fp) {
    const double id = 1.0 / delta;
    const double tp = tanh((rho + rho0) * id), tm = tanh((rho - rho0) * id), t0 = tanh(rho0 * id);
    f = (tp - tm) / 2.0 / t0;
    // 1/cosh^2 = 1 - tanh^2
    fp = ((1.0 - tp * tp) - (1.0 - tm * tm)) / (2.0 * delta) / t0;
    f = (1.0 - f_min) * f + f_min;
    fp = (1.0 - f_min) * fp;
}

// ---- spline evaluation on uniform grids ------------------------------------------------------
// cell index as bcspevxy/cspevx compute it: truncation + min + one-cell fix-up (bcspeval.f90:205-218)
RD_INLINE int spline_cell(double z, const double *__restrict__ x, int nx, double &dx) {
    const int nxm = nx - 1;
    const double x1 = __ldg(x), xn = __ldg(x + nxm);
    if (z < x1 || z > xn) {  // range fix-up within 4e-7 relative tolerance (:150-170); beyond it the
        z = z < x1 ? x1 : xn;  // reference prints an error and returns garbage — callers' box tests prevent it
    }
    int ii = 1 + (int)((double)nxm * (z - x1) / (xn - x1));
    int i = ii < nxm ? ii : nxm;
    if (i < 1) i = 1;
    if (z < __ldg(x + i - 1)) i = i - 1;
    else if (z > __ldg(x + i)) i = i + 1;
    i = i < 1 ? 1 : (i > nxm ? nxm : i);
    dx = z - __ldg(x + i - 1);
    return i;
}
// eval_2D_fp: f, fx, fy of one bicubic table (quick_cube_splines_m.f90:277-300 -> bcspevfn ict=(1,1,1,0,0,0))
RD_INLINE void bicubic_fp(const rays_spline2d &s, int i, int j, double dx, double dy, double &f, double &fx, double &fy) {
    const double2 *c = reinterpret_cast<const double2 *>(s.fspl + (size_t)((j - 1) * s.nx + (i - 1)) * 16);
    // row cy holds the 4 x-coefficients of y-power cy: F(cx,cy) = c[cy*4+cx]
    double r[4][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double2 t = __ldg(c + k);
        r[k >> 1][(k & 1) * 2] = t.x;
        r[k >> 1][(k & 1) * 2 + 1] = t.y;
    }
    // a_cx(dy) = sum_cy F(cx,cy) dy^(cy-1);  b_cx(dy) = d/dy
    double a[4], b[4];
#pragma unroll
    for (int cx = 0; cx < 4; ++cx) {
        a[cx] = r[0][cx] + dy * (r[1][cx] + dy * (r[2][cx] + dy * r[3][cx]));
        b[cx] = r[1][cx] + dy * (2.0 * r[2][cx] + dy * 3.0 * r[3][cx]);
    }
    f = a[0] + dx * (a[1] + dx * (a[2] + dx * a[3]));
    fx = a[1] + 2.0 * dx * (a[2] + 1.5 * dx * a[3]);
    fy = b[0] + dx * (b[1] + dx * (b[2] + dx * b[3]));
}
// cspeval f only (cspeval.f90:248-256), 4 coefficients per cell
RD_INLINE double cubic_f(const rays_spline1d &s, double z) {
    double dx;
    const int i = spline_cell(z, s.x_grid, s.nx, dx);
    const double2 *c = reinterpret_cast<const double2 *>(s.fspl + 4 * (size_t)(i - 1));
    const double2 c01 = __ldg(c), c23 = __ldg(c + 1);
    return c01.x + dx * (c01.y + dx * (c23.x + dx * c23.y));
}

// ---- equilibrium ------------------------------------------------------------------------------
template <int NSM> struct Eq {
    double bvec[3], g[3][3];  // g[i][j] = gradbtensor(i+1,j+1) = dB_j/dx_i
    double ns[NSM], gradns[3][NSM], ts[NSM], gradts0[3];
    double bmag, bunit[3], gradbmag[3], gradbunit[3][3];
    double alpha[NSM], gamma[NSM], omgc0;
    int err;
};

// Solov'ev flux function and field (solovev_eq_m.f90:165-190,280-322 == solovev_magnetics_m.f90:166-253)
struct SolovevGeom { double rmaj, kappa, bphi0, iota0, psiB; };
template <bool GRAD>
RD_INLINE void solovev_field(const SolovevGeom &q, double x, double y, double z, double r, double bvec[3], double g[3][3],
                             double &psiN, double gradpsiN[3]) {
    const double bp0 = q.bphi0 * q.iota0;
    const double rk = q.rmaj * q.kappa;
    const double irk2 = 1.0 / (rk * rk), irm2 = 1.0 / (q.rmaj * q.rmaj);
    const double ir = 1.0 / r;
    const double br = -bp0 * r * z * irk2;
    const double bz = bp0 * (z * z * irk2 + 0.5 * (r * r * irm2 - 1.0));
    const double bphi = q.bphi0 * q.rmaj * ir;
    const double cx = x * ir, cy = y * ir;
    bvec[0] = br * cx - bphi * cy;
    bvec[1] = br * cy + bphi * cx;
    bvec[2] = bz;
    const double a = r * z / rk;
    const double b = r * r - q.rmaj * q.rmaj;
    const double psi = 0.5 * bp0 * (a * a + b * b * irm2 * 0.25);
    const double ipsiB = 1.0 / q.psiB;
    psiN = psi * ipsiB;
    if (GRAD) {
        gradpsiN[0] = x * bz * ipsiB;
        gradpsiN[1] = y * bz * ipsiB;
        gradpsiN[2] = -r * br * ipsiB;
        const double dbrdr = br * ir;
        const double dbrdz = -bp0 * r * irk2;
        const double dbzdr = bp0 * r * irm2;
        const double dbzdz = 2.0 * bp0 * z * irk2;
        const double dbphidr = -bphi * ir;
        const double ir2 = ir * ir;
        const double x2 = x * x, y2 = y * y, xy = x * y;
        const double bri = br * ir, bphii = bphi * ir;
        g[0][0] = (dbrdr * x2 + bri * y2 + (-dbphidr + bphii) * xy) * ir2;
        g[1][0] = ((dbrdr - bri) * xy - dbphidr * y2 - bphii * x2) * ir2;
        g[2][0] = dbrdz * cx;
        g[0][1] = ((dbrdr - bri) * xy + dbphidr * x2 + bphii * y2) * ir2;
        g[1][1] = (dbrdr * y2 + bri * x2 + (dbphidr - bphii) * xy) * ir2;
        g[2][1] = dbrdz * cy;
        g[0][2] = dbzdr * cx;
        g[1][2] = dbzdr * cy;
        g[2][2] = dbzdz;
    }
}

template <int NSM, bool GRAD> RD_INLINE void eq_zero(Eq<NSM> &e) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        e.bvec[i] = 0.0;
        e.gradts0[i] = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) e.g[i][j] = 0.0;
#pragma unroll
        for (int s = 0; s < NSM; ++s) e.gradns[i][s] = 0.0;
    }
#pragma unroll
    for (int s = 0; s < NSM; ++s) { e.ns[s] = 0.0; e.ts[s] = 0.0; }
    e.err = 0;
}

// min over the active species, as minval(ns) < 0 tests need
template <int NSM> RD_INLINE bool any_negative(const double (&a)[NSM], int ns) {
    bool neg = false;
#pragma unroll
    for (int s = 0; s < NSM; ++s) if (s < ns) neg = neg || (a[s] < 0.0);
    return neg;
}

template <int NS_, bool GRAD> RD_INLINE void model_slab(double x, double y, double z, Eq<NSpec<NS_>::MAX> &e) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const rays_slab_eq &p = g_dc.c.slab;
    const rays_cfg &c = g_dc.c;
    eq_zero<NSM, GRAD>(e);
    if (x < p.xmin || x > p.xmax) e.err = RAYS_STOP_X_OUT_OF_BOUNDS;
    if (y < p.ymin || y > p.ymax) e.err = RAYS_STOP_Y_OUT_OF_BOUNDS;
    if (z < p.zmin || z > p.zmax) e.err = RAYS_STOP_Z_OUT_OF_BOUNDS;
    if (e.err) return;
    switch (p.by_prof_model) {
        case RAYS_SLAB_B_CONSTANT: e.bvec[1] = p.by0; break;
        case RAYS_SLAB_B_TOROID: e.bvec[1] = p.by0 / (1.0 + x / p.rmaj); e.g[0][1] = -e.bvec[1] / (p.rmaj + x); break;
        case RAYS_SLAB_B_LINEAR_SHEAR: e.bvec[1] = p.by0 * x / p.LBy_shear_scale; e.g[0][1] = p.by0 / p.LBy_shear_scale; break;
        default: break;
    }
    switch (p.bz_prof_model) {
        case RAYS_SLAB_B_CONSTANT: e.bvec[2] = p.bz0; break;
        case RAYS_SLAB_B_TOROID: e.bvec[2] = p.bz0 / (1.0 + x / p.rmaj); e.g[0][2] = -e.bvec[2] / (p.rmaj + x); break;
        case RAYS_SLAB_B_LINEAR: e.bvec[2] = p.bz0 * (1.0 + x / p.LBz_scale); e.g[0][2] = p.bz0 / p.LBz_scale; break;
        case RAYS_SLAB_B_LINEAR_2: e.bvec[2] = p.bz0 + p.dBzdx * (x - p.x0); e.g[0][2] = p.dBzdx; break;
        default: break;
    }
    double f = 1.0, fp = 0.0;
    bool mult = true;  // ns = n0s*f, gradns = n0s*fp
    switch (p.dens_prof_model) {
        case RAYS_PROF_CONSTANT: f = 1.0; fp = 0.0; break;
        case RAYS_PROF_LINEAR: f = 1.0 + x / p.Ln_scale; fp = 1.0 / p.Ln_scale; break;
        case RAYS_PROF_LINEAR_2: mult = false; break;
        case RAYS_PROF_PARABOLIC: parabolic_prof(x, p.n_min, p.alphan1, p.alphan2, f, fp); break;
        case RAYS_PROF_GAUSSIAN: { const double xr = x / p.rmin; f = exp(-3.0 * p.alphan1 * (xr * xr)); fp = f * (-6.0 * p.alphan1 * x / (p.rmin * p.rmin)); } break;
        default: break;
    }
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            if (mult) { e.ns[s] = c.n0s[s] * f; e.gradns[0][s] = c.n0s[s] * fp; }
            else { e.ns[s] = c.n0s[s] + p.dndx * c.eta[s] * (x - p.x0); e.gradns[0][s] = c.n0s[s] * p.dndx; }
        }
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            double t = 0.0, tp = 0.0;
            switch (p.t_prof_model[s]) {
                case RAYS_PROF_CONSTANT: t = c.t0s[s]; break;
                case RAYS_PROF_LINEAR: t = c.t0s[s] * (1.0 + x / p.LT_scale); tp = c.t0s[s] * (1.0 / p.LT_scale); break;
                case RAYS_PROF_LINEAR_2: t = c.t0s[s] + p.dtdx * (x - p.x0); tp = c.t0s[s] * p.dtdx; break;
                case RAYS_PROF_PARABOLIC: { double ff, ffp; parabolic_prof(x - p.x0, p.T_min[s], p.alphat1[s], p.alphat2[s], ff, ffp); t = c.t0s[s] * ff; tp = c.t0s[s] * ffp; } break;
                default: break;
            }
            e.ts[s] = t;
            if (s == 0) e.gradts0[0] = tp;
        }
    if (any_negative<NSM>(e.ns, ns)) e.err = RAYS_STOP_NEGATIVE_DENS;
    if (any_negative<NSM>(e.ts, ns)) e.err = RAYS_STOP_NEGATIVE_TEMP;
}

template <int NS_, bool GRAD> RD_INLINE void model_solovev(double x, double y, double z, Eq<NSpec<NS_>::MAX> &e) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const rays_solovev_eq &p = g_dc.c.solovev;
    const rays_cfg &c = g_dc.c;
    eq_zero<NSM, GRAD>(e);
    const double r = sqrt(x * x + y * y);
    if (r < p.box_rmin || r > p.box_rmax) e.err = RAYS_STOP_R_OUT_OF_BOX_SOLOVEV;
    if (z < p.box_zmin || z > p.box_zmax) e.err = RAYS_STOP_Z_OUT_OF_BOX_SOLOVEV;
    if (e.err) return;
    SolovevGeom q{p.rmaj, p.kappa, p.bphi0, p.iota0, p.psiB};
    double psiN, gpN[3];
    solovev_field<GRAD>(q, x, y, z, r, e.bvec, e.g, psiN, gpN);
    double prof = 1.0, dd = 0.0;
    if (p.dens_prof_model == RAYS_PROF_PARABOLIC) {
        prof = 0.0;
        if (psiN < 1.0) {
            const double base = 1.0 - pow_fast(psiN, p.alphan2);
            prof = pow_fast(base, p.alphan1);
            dd = -p.alphan1 * p.alphan2 * pow_fast(psiN, p.alphan2 - 1.0) * pow_fast(base, p.alphan1 - 1.0);
        }
    }
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            e.ns[s] = c.n0s[s] * prof;
            if (GRAD) {
#pragma unroll
                for (int i = 0; i < 3; ++i) e.gradns[i][s] = c.n0s[s] * dd * gpN[i];
            }
        }
    // temperature with the reference's quirks (solovev_eq_m.f90:236-262): 'constant' resets the
    // DENSITY, 'parabolic' zeroes all species' T inside the species loop, gradient exponent alphat1
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const int m = p.t_prof_model[s];
            if (m == RAYS_PROF_CONSTANT) {
#pragma unroll
                for (int q2 = 0; q2 < NSM; ++q2) {
                    if (q2 < ns) e.ns[q2] = c.n0s[q2];
#pragma unroll
                    for (int i = 0; i < 3; ++i) e.gradns[i][q2] = 0.0;
                }
            } else if (m == RAYS_PROF_PARABOLIC) {
#pragma unroll
                for (int q2 = 0; q2 < NSM; ++q2) e.ts[q2] = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) e.gradts0[i] = 0.0;
                if (psiN < 1.0) {
                    const double base = 1.0 - pow_fast(psiN, p.alphat2[s]);
                    const double pw = pow_fast(base, p.alphat1[s]);
                    e.ts[s] = c.t0s[s] * pw;
                    if (s == 0) {
                        const double ddt = -p.alphat1[s] * p.alphat2[s] * pow_fast(psiN, p.alphat2[s] - 1.0) * pw;
#pragma unroll
                        for (int i = 0; i < 3; ++i) e.gradts0[i] = c.t0s[s] * ddt * gpN[i];
                    }
                }
            } else {
                e.ts[s] = 0.0;
                if (s == 0) { e.gradts0[0] = 0.0; e.gradts0[1] = 0.0; e.gradts0[2] = 0.0; }
            }
        }
    if (any_negative<NSM>(e.ns, ns)) e.err = RAYS_STOP_NEGATIVE_DENS;
    if (any_negative<NSM>(e.ts, ns)) e.err = RAYS_STOP_NEGATIVE_TEMP;
}

template <int NS_, bool GRAD> RD_INLINE void model_axisym(double x, double y, double z, Eq<NSpec<NS_>::MAX> &e) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const rays_axisym_eq &p = g_dc.c.axisym;
    const rays_cfg &c = g_dc.c;
    const double Tiny = 10.0e-14;
    eq_zero<NSM, GRAD>(e);
    const double r = sqrt(x * x + y * y);
    if (r < p.box_rmin - Tiny || r > p.box_rmax + Tiny) e.err = RAYS_STOP_R_OUT_OF_BOX;
    if (z < p.box_zmin - Tiny || z > p.box_zmax + Tiny) e.err = RAYS_STOP_Z_OUT_OF_BOX;
    if (e.err) return;
    if (r < p.sm_box_rmin || r > p.sm_box_rmax) e.err = RAYS_STOP_R_OUT_OF_BOUNDS_SOLMAG;
    if (z < p.sm_box_zmin || z > p.sm_box_zmax) e.err = RAYS_STOP_Z_OUT_OF_BOUNDS_SOLMAG;
    if (e.err) return;
    SolovevGeom q{p.sm_rmaj, p.sm_kappa, p.sm_bphi0, p.sm_iota0, p.sm_psiB};
    double psiN, gpN[3];
    solovev_field<GRAD>(q, x, y, z, r, e.bvec, e.g, psiN, gpN);
    if (psiN > p.plasma_psi_limit) e.err = RAYS_STOP_OUT_OF_PLASMA;
    double dens = 1.0, dd = 0.0;
    if (p.density_prof_model == RAYS_PROF_PARABOLIC) parabolic_prof(psiN, p.d_scrape_off, p.alphan1, p.alphan2, dens, dd);
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            e.ns[s] = c.n0s[s] * dens;
            if (GRAD) {
#pragma unroll
                for (int i = 0; i < 3; ++i) e.gradns[i][s] = c.n0s[s] * dd * gpN[i];
            }
        }
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const int m = p.temperature_prof_model[s];
            if (m == RAYS_PROF_CONSTANT) {
                e.ts[s] = c.t0s[s];
                e.gradts0[0] = 0.0; e.gradts0[1] = 0.0; e.gradts0[2] = 0.0;  // gradts = 0. (whole array)
            } else if (m == RAYS_PROF_PARABOLIC) {
                double t, dt;
                parabolic_prof(psiN, p.T_scrape_off, p.alphat1[s], p.alphat2[s], t, dt);
                e.ts[s] = c.t0s[s] * t;
                if (s == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) e.gradts0[i] = c.t0s[s] * dt * gpN[i];
                }
            }
        }
    if (any_negative<NSM>(e.ns, ns)) e.err = RAYS_STOP_NEGATIVE_DENS;
    if (any_negative<NSM>(e.ts, ns)) e.err = RAYS_STOP_NEGATIVE_TEMP;
}
// psi_N only (deposition evaluator, deposition_profiles_m.f90:455-470 -> axisym_toroid_psi)
RD_INLINE double axisym_psiN(double x, double y, double z) {
    const rays_axisym_eq &p = g_dc.c.axisym;
    const double bp0 = p.sm_bphi0 * p.sm_iota0;
    const double r = sqrt(x * x + y * y);
    const double rk = p.sm_rmaj * p.sm_kappa;
    const double a = r * z / rk;
    const double b = r * r - p.sm_rmaj * p.sm_rmaj;
    const double psi = 0.5 * bp0 * (a * a + b * b / (p.sm_rmaj * p.sm_rmaj) / 4.0);
    return psi / p.sm_psiB;
}

template <int NS_, bool GRAD> RD_INLINE void model_mirror(double x, double y, double z, Eq<NSpec<NS_>::MAX> &e) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const rays_mirror_eq &p = g_dc.c.mirror;
    const rays_cfg &c = g_dc.c;
    eq_zero<NSM, GRAD>(e);
    const double r = sqrt(x * x + y * y);
    if (r > p.box_rmax) e.err = RAYS_STOP_R_OUT_OF_BOX;
    if (z < p.box_zmin || z > p.box_zmax) e.err = RAYS_STOP_Z_OUT_OF_BOX;
    if (e.err) return;
    // the three fields share one grid: one cell search, 3 x 16 coefficient loads (L2-resident tables)
    double dx, dy;
    const int i = spline_cell(r, p.Br_spline.x_grid, p.Br_spline.nx, dx);
    const int j = spline_cell(z, p.Br_spline.y_grid, p.Br_spline.ny, dy);
    double br, dbrdr, dbrdz, bz, dbzdr, dbzdz, Aphi, dAdr, dAdz;
    bicubic_fp(p.Br_spline, i, j, dx, dy, br, dbrdr, dbrdz);
    bicubic_fp(p.Bz_spline, i, j, dx, dy, bz, dbzdr, dbzdz);
    bicubic_fp(p.Aphi_spline, i, j, dx, dy, Aphi, dAdr, dAdz);
    double gA[3] = {0.0, 0.0, 0.0};
    if (r < 2.0 * DBL_MIN) {
        e.bvec[2] = bz;
        e.g[0][0] = -dbzdz / 2.0; e.g[1][1] = -dbzdz / 2.0; e.g[2][2] = dbzdz;
        Aphi = 0.0;
    } else {
        const double ir = 1.0 / r;
        const double cx = x * ir, cy = y * ir, bri = br * ir;
        e.bvec[0] = cx * br; e.bvec[1] = cy * br; e.bvec[2] = bz;
        if (GRAD) {
            e.g[0][0] = (1.0 - cx * cx) * bri + cx * cx * dbrdr;
            e.g[1][0] = cx * cy * (dbrdr - bri);
            e.g[2][0] = dbrdz * cx;
            e.g[0][1] = e.g[1][0];
            e.g[1][1] = (1.0 - cy * cy) * bri + cy * cy * dbrdr;
            e.g[2][1] = dbrdz * cy;
            e.g[0][2] = dbzdr * cx;
            e.g[1][2] = dbzdr * cy;
            e.g[2][2] = dbzdz;
        }
        gA[0] = dAdr * cx; gA[1] = dAdr * cy; gA[2] = dAdz;
    }
    const double iA = 1.0 / p.Aphi_LUFS;
    const double AphiN = Aphi * iA;
    if (AphiN > p.plasma_AphiN_limit) e.err = RAYS_STOP_OUT_OF_PLASMA;
    double dens = 1.0, dd = 0.0;
    if (p.density_prof_model == RAYS_PROF_PARABOLIC) parabolic_prof(AphiN, p.d_scrape_off, p.alphan1, p.alphan2, dens, dd);
    else if (p.density_prof_model == RAYS_PROF_HYPERBOLIC) hyperbolic_prof(AphiN, p.d_scrape_off, p.AphiN0_d, p.delta_d, dens, dd);
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            e.ns[s] = c.n0s[s] * dens;
            if (GRAD) {
#pragma unroll
                for (int k = 0; k < 3; ++k) e.gradns[k][s] = c.n0s[s] * dd * (gA[k] * iA);
            }
        }
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const int m = p.temperature_prof_model[s];
            if (m == RAYS_PROF_CONSTANT) {
                e.ts[s] = c.t0s[s];
                e.gradts0[0] = 0.0; e.gradts0[1] = 0.0; e.gradts0[2] = 0.0;
            } else if (m == RAYS_PROF_PARABOLIC || m == RAYS_PROF_HYPERBOLIC) {
                double t, dt;
                if (m == RAYS_PROF_PARABOLIC) parabolic_prof(AphiN, p.T_scrape_off, p.alphat1[s], p.alphat2[s], t, dt);
                else hyperbolic_prof(AphiN, p.T_scrape_off, p.AphiN0_t[s], p.delta_t[s], t, dt);
                e.ts[s] = c.t0s[s] * t;
                if (s == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) e.gradts0[k] = c.t0s[s] * dt * (gA[k] * iA);
                }
            }
        }
    if (any_negative<NSM>(e.ns, ns)) e.err = RAYS_STOP_NEGATIVE_DENS;
    if (any_negative<NSM>(e.ts, ns)) e.err = RAYS_STOP_NEGATIVE_TEMP;
}

// equilibrium(rvec, eq): model + derived quantities (equilibrium_m.f90:238-269).
// GRAD = false drops every gradient (used by deriv_num's displaced points, which only feed determ).
template <int EQ_, int NS_, bool GRAD>
RD_INLINE void equilibrium(double x, double y, double z, double inv_omgrf, Eq<NSpec<NS_>::MAX> &e) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const rays_cfg &c = g_dc.c;
    if (EQ_ == RAYS_EQ_SLAB) model_slab<NS_, GRAD>(x, y, z, e);
    else if (EQ_ == RAYS_EQ_SOLOVEV) model_solovev<NS_, GRAD>(x, y, z, e);
    else if (EQ_ == RAYS_EQ_AXISYM_TOROID) model_axisym<NS_, GRAD>(x, y, z, e);
    else model_mirror<NS_, GRAD>(x, y, z, e);
    if (e.err) return;
    const double bmag = sqrt(e.bvec[0] * e.bvec[0] + e.bvec[1] * e.bvec[1] + e.bvec[2] * e.bvec[2]);
    const double ib = 1.0 / bmag;
    e.bmag = bmag;
#pragma unroll
    for (int i = 0; i < 3; ++i) e.bunit[i] = e.bvec[i] * ib;
    if (GRAD) {
#pragma unroll
        for (int i = 0; i < 3; ++i) e.gradbmag[i] = e.g[i][0] * e.bunit[0] + e.g[i][1] * e.bunit[1] + e.g[i][2] * e.bunit[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) e.gradbunit[i][j] = (e.g[i][j] - e.gradbmag[i] * e.bunit[j]) * ib;
    }
    const double io2 = inv_omgrf * inv_omgrf;
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const double omgc = c.qs[s] * bmag / c.ms[s];
            const double omgp2 = e.ns[s] * (c.qs[s] * c.qs[s]) / (c.eps0 * c.ms[s]);
            e.alpha[s] = omgp2 * io2;
            e.gamma[s] = omgc * inv_omgrf;
            if (s == 0) e.omgc0 = omgc;
        } else { e.alpha[s] = 0.0; e.gamma[s] = 0.0; }
}

// ---- cold plasma S, D, P and the dispersion determinant ----------------------------------------
template <int NSM> RD_INLINE void stix_SDP(const Eq<NSM> &e, int ns, double &S, double &D, double &P, double &prod) {
    double s1 = 0.0, d1 = 0.0, p1 = 0.0;
    prod = 1.0;
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const double one_m = 1.0 - e.gamma[s] * e.gamma[s];
            const double a = e.alpha[s] / one_m;
            s1 += a;
            d1 += a * e.gamma[s];
            p1 += e.alpha[s];
            prod *= one_m;
        }
    S = 1.0 - s1; D = d1; P = 1.0 - p1;
}
// Re det(eps_h + n n - n^2 I), n = (n1, 0, n3)
RD_INLINE double disp_det(double S, double D, double P, double n1sq, double n3sq) {
    const double nsq = n1sq + n3sq;
    const double a11 = S - n3sq, a22 = S - nsq, a33 = P - n1sq;
    return a33 * (a11 * a22 - D * D) - n1sq * n3sq * a22;
}
// k_par^2, k_perp^2 w.r.t. bunit
RD_INLINE void kpar_kperp2(const double k[3], const double b[3], double &k3, double &k1sq) {
    k3 = k[0] * b[0] + k[1] * b[1] + k[2] * b[2];
    const double d0 = k[0] - k3 * b[0], d1 = k[1] - k3 * b[1], d2 = k[2] - k3 * b[2];
    k1sq = d0 * d0 + d1 * d1 + d2 * d2;
}

// ---- deriv_cold (deriv_cold.f90:40-171) -----------------------------------------------------------
template <int NS_>
RD_INLINE void deriv_cold(const Eq<NSpec<NS_>::MAX> &e, const double nvec[3], double dddx[3], double dddk[3], double &dddw) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const double inv_k0 = g_dc.inv_k0, io = g_dc.inv_omgrf;
    const double n3 = nvec[0] * e.bunit[0] + nvec[1] * e.bunit[1] + nvec[2] * e.bunit[2];
    double dperp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dperp[i] = nvec[i] - n3 * e.bunit[i];
    const double n1sq = dperp[0] * dperp[0] + dperp[1] * dperp[1] + dperp[2] * dperp[2];
    const double n3sq = n3 * n3, n3p4 = n3sq * n3sq, n1p4 = n1sq * n1sq;
    double dn3dx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dn3dx[i] = e.gradbunit[i][0] * nvec[0] + e.gradbunit[i][1] * nvec[1] + e.gradbunit[i][2] * nvec[2];
    double p = 0.0, t = 1.0;
#pragma unroll
    for (int s = 0; s < NSM; ++s) if (s < ns) { p += e.alpha[s]; t *= 1.0 - e.gamma[s] * e.gamma[s]; }
    p = 1.0 - p;
    double dq1da[NSM], dq2da[NSM];
#pragma unroll
    for (int s1 = 0; s1 < NSM; ++s1) {
        double a = 1.0, b = 1.0;
#pragma unroll
        for (int s = 0; s < NSM; ++s) if (s < ns && s != s1) { a *= 1.0 + e.gamma[s]; b *= 1.0 - e.gamma[s]; }
        dq1da[s1] = a; dq2da[s1] = b;
    }
    double q1 = 0.0, q2 = 0.0, su = 0.0;
#pragma unroll
    for (int s = 0; s < NSM; ++s) if (s < ns) { q1 += e.alpha[s] * dq1da[s]; q2 += e.alpha[s] * dq2da[s]; su += e.alpha[s] * dq1da[s] * dq2da[s]; }
    const double u = t - su;
    const double q = 2.0 * u - t + q1 * q2;
    const double ibmag = 1.0 / e.bmag;
    double accx[3] = {0.0, 0.0, 0.0}, accw = 0.0;
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) {
            const double al = e.alpha[s], ga = e.gamma[s];
            const double duda = -dq1da[s] * dq2da[s];
            const double dqda = 2.0 * duda + dq1da[s] * q2 + q1 * dq2da[s];
            const double ddda = -t * n3p4 + (2.0 * (u - p * duda) + (-t + duda) * n1sq) * n3sq - q + p * dqda -
                                (dqda - u + p * duda) * n1sq + duda * n1p4;
            // sums over s1 of alpha(s1)*gp/gm/gpm(s1,s): products over species other than s1 and s
            double sgp = 0.0, sgm = 0.0, sgpm = 0.0;
#pragma unroll
            for (int s1 = 0; s1 < NSM; ++s1)
                if (s1 < ns) {
                    double gp = 1.0, gm = 1.0;
#pragma unroll
                    for (int s2 = 0; s2 < NSM; ++s2) if (s2 < ns && s2 != s1 && s2 != s) { gp *= 1.0 + e.gamma[s2]; gm *= 1.0 - e.gamma[s2]; }
                    sgp += e.alpha[s1] * gp; sgm += e.alpha[s1] * gm; sgpm += e.alpha[s1] * (gp * gm);
                }
            const double dtdg = 2.0 * ga * duda;
            const double dudg = dtdg + 2.0 * ga * (sgpm + al * duda);
            const double dq1dg = sgp - al * dq1da[s];
            const double dq2dg = -sgm + al * dq2da[s];
            const double dqdg = 2.0 * dudg - dtdg + dq1dg * q2 + q1 * dq2dg;
            const double dddg = dtdg * p * n3p4 + (-2.0 * p * dudg + (dtdg * p + dudg) * n1sq) * n3sq + p * dqdg -
                                (dqdg + p * dudg) * n1sq + dudg * n1p4;
            // dadx = alpha*gradns/ns keeps the reference's 0*0/0 = NaN outside the plasma
#pragma unroll
            for (int i = 0; i < 3; ++i) accx[i] += ddda * (al * e.gradns[i][s] / e.ns[s]) + dddg * (ga * e.gradbmag[i] * ibmag);
            accw += ddda * (-2.0 * io * al) + dddg * (-io * ga);
        }
    const double dddn3 = (4.0 * t * p * n3sq + 2.0 * (-2.0 * p * u + (t * p + u) * n1sq)) * n3;
    const double dddn12 = (t * p + u) * n3sq - (q + p * u) + 2.0 * u * n1sq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dddk[i] = (dddn3 * e.bunit[i] + dddn12 * 2.0 * dperp[i]) * inv_k0;
        dddx[i] = accx[i] + dddn3 * dn3dx[i] + dddn12 * (-2.0 * n3 * dn3dx[i]);
    }
    dddw = accw + dddn3 * (-n3 * io) + dddn12 * (-2.0 * io * n1sq);
}

// ---- deriv_num (deriv_num.f90:40-98) ----------------------------------------------------------------
// Central differences of D = Re det * prod(1-gamma^2); omega and k0 are thread-local arguments
// (the Fortran perturbs module globals, which is why its OpenMP run is racy).
template <int NSM> RD_INLINE double determ(const Eq<NSM> &e, int ns, const double k[3], double inv_k0) {
    double S, D, P, prod, k3, k1sq;
    stix_SDP<NSM>(e, ns, S, D, P, prod);
    kpar_kperp2(k, e.bunit, k3, k1sq);
    const double ik2 = inv_k0 * inv_k0;
    return disp_det(S, D, P, k1sq * ik2, k3 * k3 * ik2) * prod;
}
template <int EQ_, int NS_>
RD_INLINE void deriv_num(const Eq<NSpec<NS_>::MAX> &e0, const double r0[3], const double k0v[3], double dddx[3], double dddk[3],
                         double &dddw, int &pert_err) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const double delta = (double)1.e-6f;  // single-precision literal (deriv_num.f90:37)
    const double io = g_dc.inv_omgrf, ik0 = g_dc.inv_k0;
    pert_err = 0;
    Eq<NSM> ep;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const double hx = i == 0 ? delta : 0.0, hy = i == 1 ? delta : 0.0, hz = i == 2 ? delta : 0.0;
        equilibrium<EQ_, NS_, false>((r0[0] + hx), (r0[1] + hy), (r0[2] + hz), io, ep);
        if (ep.err && !pert_err) pert_err = ep.err;
        const double dp = determ<NSM>(ep, ns, k0v, ik0);
        equilibrium<EQ_, NS_, false>((r0[0] - hx), (r0[1] - hy), (r0[2] - hz), io, ep);
        if (ep.err && !pert_err) pert_err = ep.err;
        const double dm = determ<NSM>(ep, ns, k0v, ik0);
        const double d = (dp - dm) / (2.0 * delta);
        if (i == 0) dddx[0] = d; else if (i == 1) dddx[1] = d; else dddx[2] = d;
    }
    {   // k derivatives share (S, D, P) of the unperturbed point
        double S, D, P, prod;
        stix_SDP<NSM>(e0, ns, S, D, P, prod);
        const double ik2 = ik0 * ik0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double change = fmax(delta, fabs(delta * k0v[i])) / 2.0;
            double kp[3] = {k0v[0], k0v[1], k0v[2]}, km[3] = {k0v[0], k0v[1], k0v[2]};
            kp[i] = k0v[i] + change;
            km[i] = k0v[i] - change;
            double k3, k1sq;
            kpar_kperp2(kp, e0.bunit, k3, k1sq);
            const double dp = disp_det(S, D, P, k1sq * ik2, k3 * k3 * ik2) * prod;
            kpar_kperp2(km, e0.bunit, k3, k1sq);
            const double dm = disp_det(S, D, P, k1sq * ik2, k3 * k3 * ik2) * prod;
            dddk[i] = (dp - dm) / (2.0 * change);
        }
    }
    {   // omega derivative: alpha ~ 1/w^2, gamma ~ 1/w, k0 ~ w; geometry (bunit) unchanged
        const double omgrf0 = g_dc.c.omgrf;
        double det[2];
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
            const double omg = omgrf0 * (sgn == 0 ? (1.0 + delta / 2.0) : (1.0 - delta / 2.0));
            const double k0p = omg / g_dc.c.clight;
            const double iop = 1.0 / omg;
            Eq<NSM> ew;
#pragma unroll
            for (int i = 0; i < 3; ++i) ew.bunit[i] = e0.bunit[i];
#pragma unroll
            for (int s = 0; s < NSM; ++s) {
                // alpha = omgp2/omg^2, gamma = omgc/omg with omgp2 = alpha0*omgrf0^2, omgc = gamma0*omgrf0
                const double rw = omgrf0 * iop;
                ew.alpha[s] = e0.alpha[s] * rw * rw;
                ew.gamma[s] = e0.gamma[s] * rw;
            }
            det[sgn] = determ<NSM>(ew, ns, k0v, 1.0 / k0p);
        }
        dddw = (det[0] - det[1]) / (omgrf0 * delta);
    }
}

// ---- damping: damp_fund_ECH (damp_fund_ECH.f90:2-128) -----------------------------------------------
// returns ksi(0) = ki; all other species contribute 0.  D_WARM and DELTA are single-precision
// complex in the reference (:36), which quantises k_i to ~6e-8 relative: reproduced with float casts.
template <int NSM> RD_INLINE double damp_fund_ECH(const Eq<NSM> &e, const double kvec[3], const double vg[3]) {
    const rays_cfg &c = g_dc.c;
    const double ik0 = g_dc.inv_k0;
    double k3, k1sq;
    kpar_kperp2(kvec, e.bunit, k3, k1sq);
    const double R3 = k3 * ik0;
    const double R1S = k1sq * ik0 * ik0, R3S = R3 * R3, RS = R1S + R3S;
    const double B1 = e.gamma[0], BETAE = B1 * B1;
    if (R3 == 0.0) return 0.0;
    const double vth = sqrt(2.0 * e.ts[0] / c.ms[0]);
    const double VT = vth / c.clight;
    const double xi = (c.omgrf + e.omgc0) / (k3 * vth);
    if (fabs(xi) > 5.0) return 0.0;
    // zfun0_real_arg(xi, k3): Z(xi) for k3 > 0, -Z(-xi) for k3 < 0; Re from the spline table,
    // Im = sqrt(pi) exp(-x^2) (zfunctions_m.f90:351-432)
    const double sqrt_pi = 1.7724538509055159;
    const double xa = k3 > 0.0 ? xi : -xi;
    double zr = cubic_f(c.zfun_re, xa);
    double zi = sqrt_pi * exp(-(xa * xa));
    if (!(k3 > 0.0)) { zr = -zr; zi = -zi; }
    const double P = e.alpha[0];
    const double Q = P / 2.0 / (1.0 - B1);
    const double omq = 1.0 - Q, omp = 1.0 - P, om2q = 1.0 - 2.0 * Q;
    const double L1 = omq * RS * R1S + omp * RS * R3S - omq * omp * (RS + R3S) - om2q * R1S + om2q * omp;
    const double L2 = -P / B1 * (RS * R1S - om2q * R1S) + (P * P) / 4.0 / BETAE * R1S / R3S * (RS + R3S - 2.0 * om2q);
    const double L5 = P * (RS * R3S - omq * (RS + R3S) + om2q);
    const double fac = -(1.0 - B1) * R3 * VT * (L1 + L2 + R1S / 2.0 / R3 / BETAE * VT * xi * L5);
    // xi + 1/zf
    const double zden = 1.0 / (zr * zr + zi * zi);
    const double pr = xi + zr * zden, pi = -zi * zden;
    const float dwr = (float)(fac * pr), dwi = (float)(fac * pi);
    const double A = 1.0 - P - BETAE;
    const double B = -(omp * A + omp * omp - BETAE) + (A + omp * (1.0 - BETAE)) * R3S;
    const double DDNX2 = 2.0 * A * R1S + B;
    const double DDNZ = 2.0 * R3 * ((A + omp * (1.0 - BETAE)) * R1S + omp * (2.0 * (1.0 - BETAE) * R3S - 2.0 * A));
    const double vgn = sqrt(vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]);
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double dnperp2 = 2.0 * (kvec[i] * ik0 - R3 * e.bunit[i]);
        dot += (DDNX2 * dnperp2 + DDNZ * e.bunit[i]) * (vg[i] / vgn);
    }
    (void)dwr;
    const float deli = (float)(-(double)dwi / dot);
    return c.k0 * (double)deli;
}

// ---- residual of the dispersion relation at a saved point (check_save.f90:163-235) -----------------
template <int NSM> RD_INLINE double residual(const Eq<NSM> &e, int ns, const double kvec[3]) {
    double S, D, P, prod, k3, k1sq;
    stix_SDP<NSM>(e, ns, S, D, P, prod);
    kpar_kperp2(kvec, e.bunit, k3, k1sq);
    const double ik2 = g_dc.inv_k0 * g_dc.inv_k0;
    const double n1sq = k1sq * ik2, n3sq = k3 * k3 * ik2;
    const double det = disp_det(S, D, P, n1sq, n3sq);
    const double e11 = fabs(S) + n1sq, e22 = fabs(S), e33 = fabs(P) + n3sq, e12 = fabs(D);
    const double e13sq = n1sq * n3sq;  // |n1 n3|^2
    const double den = e33 * (e11 * e22) + e33 * (e12 * e12) + e13sq * e22;
    return fabs(det) / den;
}

// ---- cold dispersion roots for the launch kernels -----------------------------------------------------
struct cplx { double re, im; };
RD_INLINE cplx csqrt_dev(cplx z) {
    cplx r;
    if (z.im == 0.0) {
        if (z.re >= 0.0) { r.re = sqrt(z.re); r.im = z.im; }
        else { r.re = 0.0; r.im = copysign(sqrt(-z.re), z.im); }
        return r;
    }
    const double m = hypot(z.re, z.im);
    if (z.re >= 0.0) { r.re = sqrt(0.5 * (m + z.re)); r.im = z.im / (2.0 * r.re); }
    else { r.im = copysign(sqrt(0.5 * (m - z.re)), z.im); r.re = z.im / (2.0 * r.im); }
    return r;
}
// Stix R, L, S, D, P as RLSDP_cold returns them (suscep_m.f90:180-219)
template <int NSM> RD_INLINE void stix_RLSP(const Eq<NSM> &e, int ns, double &S, double &P, double &R, double &L) {
    double r1 = 0.0, l1 = 0.0, p1 = 0.0;
#pragma unroll
    for (int s = 0; s < NSM; ++s)
        if (s < ns) { r1 -= e.alpha[s] / (1.0 + e.gamma[s]); l1 -= e.alpha[s] / (1.0 - e.gamma[s]); p1 -= e.alpha[s]; }
    R = 1.0 + r1; L = 1.0 + l1; S = (R + L) / 2.0; P = 1.0 + p1;
}
// solve_n1_vs_n2_n3 (dispersion_solvers_m.f90:49-112 + disp_solve_cold_n1sq_vs_n3.f90:1-90)
template <int NSM> RD_INLINE cplx solve_n1_vs_n2_n3(const Eq<NSM> &e, int ns, double n2, double n3) {
    double S, P, R, L;
    stix_RLSP<NSM>(e, ns, S, P, R, L);
    const double n3s = n3 * n3;
    const double a = S, b = -R * L - P * S + n3s * (P + S), cc = P * (n3s - R) * (n3s - L);
    const double discr = b * b - 4.0 * a * cc;
    cplx sq = csqrt_dev(cplx{discr, 0.0});
    cplx plus, minus;
    if (copysign(1.0, b) < 0.0) {
        const cplx num{-b + sq.re, sq.im};
        plus = cplx{num.re / (2.0 * a), num.im / (2.0 * a)};
        const double d2 = num.re * num.re + num.im * num.im;
        minus = cplx{2.0 * cc * num.re / d2, -2.0 * cc * num.im / d2};
    } else {
        const cplx num{-b - sq.re, -sq.im};
        minus = cplx{num.re / (2.0 * a), num.im / (2.0 * a)};
        const double d2 = num.re * num.re + num.im * num.im;
        plus = cplx{2.0 * cc * num.re / d2, -2.0 * cc * num.im / d2};
    }
    cplx sel;
    const int mode = g_dc.c.wave_mode;
    const bool plus_is_fast = hypot(plus.re, plus.im) <= hypot(minus.re, minus.im);
    if (mode == RAYS_MODE_PLUS) sel = plus;
    else if (mode == RAYS_MODE_MINUS) sel = minus;
    else if (mode == RAYS_MODE_FAST) sel = plus_is_fast ? plus : minus;
    else sel = plus_is_fast ? minus : plus;
    sel.re -= n2 * n2;
    cplx r = csqrt_dev(sel);
    const double ks = (double)g_dc.c.k0_sign;
    return cplx{ks * r.re, ks * r.im};
}
// solve_n_vs_theta (dispersion_solvers_m.f90:157-231 + disp_solve_cold_nsq_vs_theta.f90): real n (NaN if n^2 < 0)
template <int NSM> RD_INLINE bool solve_n_vs_theta(const Eq<NSM> &e, int ns, double theta, double &n_out) {
    double S, P, R, L;
    stix_RLSP<NSM>(e, ns, S, P, R, L);
    const double ct = cos(theta), cos2 = ct * ct, sin2 = 1.0 - cos2;
    const double a = S * sin2 + P * cos2, b = -R * L * sin2 - P * S * (1.0 + cos2), cc = P * R * L;
    const double discr = b * b - 4.0 * a * cc;
    if (discr < 0.0) return false;
    const double sq = sqrt(discr);
    double plus, minus;
    if (copysign(1.0, b) < 0.0) { plus = (-b + sq) / (2.0 * a); minus = 2.0 * cc / (-b + sq); }
    else { minus = (-b - sq) / (2.0 * a); plus = 2.0 * cc / (-b - sq); }
    const int mode = g_dc.c.wave_mode;
    const bool plus_is_fast = fabs(plus) <= fabs(minus);
    double sel;
    if (mode == RAYS_MODE_PLUS) sel = plus;
    else if (mode == RAYS_MODE_MINUS) sel = minus;
    else if (mode == RAYS_MODE_FAST) sel = plus_is_fast ? plus : minus;
    else sel = plus_is_fast ? minus : plus;
    n_out = (double)g_dc.c.k0_sign * sqrt(sel);
    return true;
}

}  // namespace rays_dev
