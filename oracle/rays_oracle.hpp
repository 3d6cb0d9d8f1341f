// rays_oracle.hpp — CPU ORACLE for the RAYS ray-integration hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may compile, link or call this.  The product (rays_b200/csrc) never does.
//
// This is a statement-by-statement C++ restatement of the reference's Fortran under
// /root/reference/RAYS_project (abbreviated L/ = RAYS_lib/, S/ = splines_lib/, M/ =
// math_functions_lib/, P/ = post_process_lib/).  Every function cites the file:line it follows.
// Evaluation order, single-precision literals, integer powers by multiplication, complex
// arithmetic expansion and copysign semantics follow SURVEY.md Appendix A.
//
// PARITY STATUS: the reference cannot be compiled here (no Fortran compiler, no netCDF) and ships no
// numeric ray files -- but its example directories ship vector PDFs of the rays the Fortran traced.
// PINNED on (a) those figures: 29 rays of 7 inputs (slab + Solov'ev, all SG_ODE; 2 slab inputs reconstructed), 2 164 plotted
// trajectory points reproduced to the PDFs' resolution of 1e-6 pt (1.3e-9 ... 8.7e-9 m; 3e-14 m in z on
// the equatorial-plane runs), ray lengths included (tests/golden/ref_plot_vectors.json,
// tests/test_reference_plots.py); (b) the plasma Z function table M/"Splined Z function results.txt":48-85
// (tests/golden/zfun_kat.json); (c) the cold dispersion roots of the launchers on the slab example's kx-profile figures
// (13 pages x 404 values, 1e-4 ... 1e-6 pt: tests/golden/ref_kx_profiles.json); (d) coarsely, RK4_ODE + the mirror equilibrium on the MPEX example's raster figure
// (11 rays, 0.87 mm pixels: tests/golden/ref_raster_vectors.json).  NOT pinned by reference output: deriv_num,
// damping (k and power are pinned only through the positions they drive), the axisym/eqdsk equilibria.
//
// Where the Fortran has undefined behaviour the oracle makes a documented deterministic choice
// (marked "(X)" as in SURVEY.md A.5); the CUDA path makes the same choice.
#pragma once
#include <cfloat>
#include <cstring>
#include <vector>

#include "../include/rays_b200.h"
#include "oracle_real.hpp"

namespace rays_oracle {

constexpr int NS0 = RAYS_NSPECIES;  // 0:nspec0

// type eq_point  (L/equilibrium_m.f90:39-59).  Fortran a(i,j) -> C a[i-1][j-1]; species (…,is) -> [..][is]
template <class R> struct EqPoint {
    R bvec[3], bmag, gradbmag[3], bunit[3], gradbunit[3][3], gradbtensor[3][3];
    R ns[NS0], gradns[3][NS0];
    R ts[NS0], gradts[3][NS0];
    R omgc[NS0], omgp2[NS0], alpha[NS0], gamma[NS0];
    int equib_err;  // enum rays_stop_code, 0 = ''
};

// type ode_stop (L/ode_m.f90:24-29)
template <class R> struct OdeStop {
    bool stop_ode;
    int ode_stop_flag;
    R rel_err, abs_err;
    int run_error;  // run-level `stop 1` conditions (A.5 last item) -> enum rays_status
};

// ============================ splines: evaluation ============================================
// cspevx + cspevfn, uniform grid branch (S/cspeval.f90:93-176,180-295), ict = (1,0,0) or (1,1,0)
template <class R>
inline int cspeval(R xget, const rays_spline1d &s, R &f, R *fp = nullptr) {
    const double *x = s.x_grid;
    const int nx = s.nx;
    R zxget = xget;
    if ((xget < x[0]) || (xget > x[nx - 1])) {
        double zxtol = 4.0E-7 * std::fmax(std::fabs(x[0]), std::fabs(x[nx - 1]));
        if ((xget < x[0] - zxtol) || (xget > x[nx - 1] + zxtol)) return 1;  // ier=1, fval untouched
        if (xget < x[0]) zxget = R(x[0]);
        else zxget = R(x[nx - 1]);
    }
    const int nxm = nx - 1;
    // ii=1+nxm*(zxget-x(1))/(x(nx)-x(1))  real->integer truncation (S/cspeval.f90:153)
    R t = R((double)nxm) * (zxget - x[0]) / (x[nx - 1] - x[0]);
    int ii = 1 + (int)val(t);
    int i = ii < nxm ? ii : nxm;  // 1-based
    if (zxget < x[i - 1]) i = i - 1;
    else if (zxget > x[i]) i = i + 1;
    R dx = zxget - x[i - 1];
    const double *c = s.fspl + 4 * (i - 1);
    f = c[0] + dx * (c[1] + dx * (c[2] + dx * c[3]));
    if (fp) *fp = c[1] + dx * (2.0 * c[2] + dx * 3.0 * c[3]);
    return 0;
}

// bcspevxy + bcspevfn with ict = (1,1,1,0,0,0) (S/bcspeval.f90:128-255, :368-407),
// as called by eval_2D_fp (S/quick_cube_splines_m.f90:277-300)
template <class R>
inline int bcspeval_fp(R xget, R yget, const rays_spline2d &s, R &f, R &fx, R &fy) {
    const double *x = s.x_grid, *y = s.y_grid;
    const int nx = s.nx, ny = s.ny;
    R zxget = xget, zyget = yget;
    int ier = 0;
    if ((xget < x[0]) || (xget > x[nx - 1])) {
        double zxtol = 4.0E-7 * std::fmax(std::fabs(x[0]), std::fabs(x[nx - 1]));
        if ((xget < x[0] - zxtol) || (xget > x[nx - 1] + zxtol)) ier = 1;
        else if (xget < x[0]) zxget = R(x[0]);
        else zxget = R(x[nx - 1]);
    }
    if ((yget < y[0]) || (yget > y[ny - 1])) {
        double zytol = 4.0E-7 * std::fmax(std::fabs(y[0]), std::fabs(y[ny - 1]));
        if ((yget < y[0] - zytol) || (yget > y[ny - 1] + zytol)) ier = 1;
        else if (yget < y[0]) zyget = R(y[0]);
        else zyget = R(y[ny - 1]);
    }
    if (ier != 0) return ier;
    const int nxm = nx - 1, nym = ny - 1;
    int ii = 1 + (int)val(R((double)nxm) * (zxget - x[0]) / (x[nx - 1] - x[0]));
    int i = ii < nxm ? ii : nxm;
    if (zxget < x[i - 1]) i = i - 1;
    else if (zxget > x[i]) i = i + 1;
    int jj = 1 + (int)val(R((double)nym) * (zyget - y[0]) / (y[ny - 1] - y[0]));
    int j = jj < nym ? jj : nym;
    if (zyget < y[j - 1]) j = j - 1;
    else if (zyget > y[j]) j = j + 1;
    R dx = zxget - x[i - 1];
    R dy = zyget - y[j - 1];
    // f(cx,cy,i,j) -> fspl[((j*nx+i)*4+cy)*4+cx]
    const double *c = s.fspl + (size_t)((j - 1) * nx + (i - 1)) * 16;
#define F(cx, cy) c[((cy)-1) * 4 + ((cx)-1)]
    f = F(1, 1) + dy * (F(1, 2) + dy * (F(1, 3) + dy * F(1, 4))) +
        dx * (F(2, 1) + dy * (F(2, 2) + dy * (F(2, 3) + dy * F(2, 4))) +
              dx * (F(3, 1) + dy * (F(3, 2) + dy * (F(3, 3) + dy * F(3, 4))) +
                    dx * (F(4, 1) + dy * (F(4, 2) + dy * (F(4, 3) + dy * F(4, 4))))));
    fx = F(2, 1) + dy * (F(2, 2) + dy * (F(2, 3) + dy * F(2, 4))) +
         2.0 * dx *
             (F(3, 1) + dy * (F(3, 2) + dy * (F(3, 3) + dy * F(3, 4))) +
              1.5 * dx * (F(4, 1) + dy * (F(4, 2) + dy * (F(4, 3) + dy * F(4, 4)))));
    fy = F(1, 2) + dy * (2.0 * F(1, 3) + dy * 3.0 * F(1, 4)) +
         dx * (F(2, 2) + dy * (2.0 * F(2, 3) + dy * 3.0 * F(2, 4)) +
               dx * (F(3, 2) + dy * (2.0 * F(3, 3) + dy * 3.0 * F(3, 4)) +
                     dx * (F(4, 2) + dy * (2.0 * F(4, 3) + dy * 3.0 * F(4, 4)))));
#undef F
    return 0;
}

// bcspevxy + bcspevfn with ict = (1,1,1,1,1,1) (S/bcspeval.f90:128-255, :368-455), as called by eval_2D_fpp
// (S/quick_cube_splines_m.f90:305-332): f, fx, fy, fxx, fyy, fxy.  (X) on a range error the Fortran prints and
// uses fval uninitialised; here the outputs are left as the caller set them.
template <class R>
inline int bcspeval_fpp(R xget, R yget, const rays_spline2d &s, R &f, R &fx, R &fy, R &fxx, R &fyy, R &fxy) {
    if (bcspeval_fp(xget, yget, s, f, fx, fy) != 0) return 1;
    const double *x = s.x_grid, *y = s.y_grid;
    const int nx = s.nx, ny = s.ny, nxm = nx - 1, nym = ny - 1;
    R zxget = xget, zyget = yget;   // same zone lookup as bcspeval_fp
    if (xget < x[0]) zxget = R(x[0]); else if (xget > x[nx - 1]) zxget = R(x[nx - 1]);
    if (yget < y[0]) zyget = R(y[0]); else if (yget > y[ny - 1]) zyget = R(y[ny - 1]);
    int ii = 1 + (int)val(R((double)nxm) * (zxget - x[0]) / (x[nx - 1] - x[0]));
    int i = ii < nxm ? ii : nxm;
    if (zxget < x[i - 1]) i = i - 1;
    else if (zxget > x[i]) i = i + 1;
    int jj = 1 + (int)val(R((double)nym) * (zyget - y[0]) / (y[ny - 1] - y[0]));
    int j = jj < nym ? jj : nym;
    if (zyget < y[j - 1]) j = j - 1;
    else if (zyget > y[j]) j = j + 1;
    R dx = zxget - x[i - 1];
    R dy = zyget - y[j - 1];
    const double *c = s.fspl + (size_t)((j - 1) * nx + (i - 1)) * 16;
#define F(cx, cy) c[((cy)-1) * 4 + ((cx)-1)]
    fxx = 2.0 * (F(3, 1) + dy * (F(3, 2) + dy * (F(3, 3) + dy * F(3, 4)))) +
          6.0 * dx * (F(4, 1) + dy * (F(4, 2) + dy * (F(4, 3) + dy * F(4, 4))));
    fyy = 2.0 * F(1, 3) + 6.0 * dy * F(1, 4) +
          dx * (2.0 * F(2, 3) + 6.0 * dy * F(2, 4) +
                dx * (2.0 * F(3, 3) + 6.0 * dy * F(3, 4) +
                      dx * (2.0 * F(4, 3) + 6.0 * dy * F(4, 4))));
    fxy = F(2, 2) + dy * (2.0 * F(2, 3) + dy * 3.0 * F(2, 4)) +
          2.0 * dx * (F(3, 2) + dy * (2.0 * F(3, 3) + dy * 3.0 * F(3, 4)) +
                      1.5 * dx * (F(4, 2) + dy * (2.0 * F(4, 3) + dy * 3.0 * F(4, 4))));
#undef F
    return 0;
}

// ============================ profile helpers ================================================
// parabolic_prof (L/slab_eq_m.f90:354-381, L/axisym_toroid_eq_m.f90:505-521, L/multiple_mirror_eq_m.f90:465-481)
// (X) fp is left undefined by the Fortran when rho >= 1 and f >= f_min; here it is 0.
template <class R> inline void parabolic_prof(R rho, R f_min, R alpha1, R alpha2, R &f, R &fp) {
    f = R(0.0);
    fp = R(0.0);
    if (rho < 1.0) {
        f = Pow(1.0 - Pow(rho, alpha2), alpha1);
        fp = -alpha1 * alpha2 * Pow(rho, alpha2 - 1.0) * Pow(1.0 - Pow(rho, alpha2), alpha1 - 1.0);
    }
    if (f < f_min) {
        f = f_min;
        fp = R(0.0);
    }
}
// hyperbolic_prof (L/multiple_mirror_eq_m.f90:486-505)
template <class R> inline void hyperbolic_prof(R rho, R f_min, R rho0, R delta, R &f, R &fp) {
    f = (Tanh((rho + rho0) / delta) - Tanh((rho - rho0) / delta)) / 2.0 / Tanh(rho0 / delta);
    R cp = Cosh((rho + rho0) / delta), cm = Cosh((rho - rho0) / delta);
    fp = (1.0 / (cp * cp) - 1.0 / (cm * cm)) / (2.0 * delta) / Tanh(rho0 / delta);
    f = (1.0 - f_min) * f + f_min;
    fp = (1.0 - f_min) * fp;
}

// ============================ equilibrium models =============================================
template <class R> struct ModelOut {
    R bvec[3], gradbtensor[3][3], ns[NS0], gradns[3][NS0], ts[NS0], gradts[3][NS0];
    int equib_err;
};
template <class R> inline void zero_model(ModelOut<R> &m) {
    for (int i = 0; i < 3; ++i) {
        m.bvec[i] = R(0.0);
        for (int j = 0; j < 3; ++j) m.gradbtensor[i][j] = R(0.0);
        for (int s = 0; s < NS0; ++s) { m.gradns[i][s] = R(0.0); m.gradts[i][s] = R(0.0); }
    }
    for (int s = 0; s < NS0; ++s) { m.ns[s] = R(0.0); m.ts[s] = R(0.0); }
    m.equib_err = 0;
}
template <class R> inline R minval(const R *a, int n) {
    R m = a[0];
    for (int i = 1; i < n; ++i) if (a[i] < m) m = a[i];
    return m;
}

// slab_eq (L/slab_eq_m.f90:125-309)
template <class R> inline void slab_eq(const rays_cfg &c, const R rvec[3], ModelOut<R> &m) {
    const rays_slab_eq &p = c.slab;
    const int nspec = c.nspec;
    zero_model(m);
    R x = rvec[0], y = rvec[1], z = rvec[2];
    if (x < p.xmin || x > p.xmax) m.equib_err = RAYS_STOP_X_OUT_OF_BOUNDS;
    if (y < p.ymin || y > p.ymax) m.equib_err = RAYS_STOP_Y_OUT_OF_BOUNDS;
    if (z < p.zmin || z > p.zmax) m.equib_err = RAYS_STOP_Z_OUT_OF_BOUNDS;
    if (m.equib_err != 0) return;
    m.bvec[0] = R(0.0);  // bx 'zero'
    switch (p.by_prof_model) {
        case RAYS_SLAB_B_ZERO: m.bvec[1] = R(0.0); break;
        case RAYS_SLAB_B_CONSTANT: m.bvec[1] = R(p.by0); break;
        case RAYS_SLAB_B_TOROID:
            m.bvec[1] = p.by0 / (1.0 + x / p.rmaj);
            m.gradbtensor[0][1] = -m.bvec[1] / (p.rmaj + x);
            break;
        case RAYS_SLAB_B_LINEAR_SHEAR:
            m.bvec[1] = p.by0 * x / p.LBy_shear_scale;
            m.gradbtensor[0][1] = R(p.by0 / p.LBy_shear_scale);
            break;
    }
    switch (p.bz_prof_model) {
        case RAYS_SLAB_B_CONSTANT: m.bvec[2] = R(p.bz0); break;
        case RAYS_SLAB_B_TOROID:
            m.bvec[2] = p.bz0 / (1.0 + x / p.rmaj);
            m.gradbtensor[0][2] = -m.bvec[2] / (p.rmaj + x);
            break;
        case RAYS_SLAB_B_LINEAR:
            m.bvec[2] = p.bz0 * (1.0 + x / p.LBz_scale);
            m.gradbtensor[0][2] = R(p.bz0 / p.LBz_scale);
            break;
        case RAYS_SLAB_B_LINEAR_2:
            m.bvec[2] = p.bz0 + p.dBzdx * (x - p.x0);
            m.gradbtensor[0][2] = R(p.dBzdx);
            break;
    }
    switch (p.dens_prof_model) {
        case RAYS_PROF_CONSTANT:
            for (int s = 0; s <= nspec; ++s) m.ns[s] = R(c.n0s[s]);
            break;
        case RAYS_PROF_LINEAR:
            for (int s = 0; s <= nspec; ++s) {
                m.ns[s] = c.n0s[s] * (1.0 + x / p.Ln_scale);
                m.gradns[0][s] = R(c.n0s[s] * (1.0 / p.Ln_scale));
            }
            break;
        case RAYS_PROF_LINEAR_2:
            for (int s = 0; s <= nspec; ++s) {
                m.ns[s] = c.n0s[s] + p.dndx * c.eta[s] * (x - p.x0);
                m.gradns[0][s] = R(c.n0s[s] * p.dndx);
            }
            break;
        case RAYS_PROF_PARABOLIC: {
            R f, fp;
            parabolic_prof(x, R(p.n_min), R(p.alphan1), R(p.alphan2), f, fp);
            for (int s = 0; s <= nspec; ++s) { m.ns[s] = c.n0s[s] * f; m.gradns[0][s] = c.n0s[s] * fp; }
        } break;
        case RAYS_PROF_GAUSSIAN:
            for (int s = 0; s <= nspec; ++s) {
                R xr = x / p.rmin;
                m.ns[s] = c.n0s[s] * Exp(-3.0 * p.alphan1 * (xr * xr));
                m.gradns[0][s] = m.ns[s] * (-6.0 * p.alphan1 * x / (p.rmin * p.rmin));
            }
            break;
    }
    for (int s = 0; s <= nspec; ++s) {
        switch (p.t_prof_model[s]) {
            case RAYS_PROF_ZERO: m.ts[s] = R(0.0); break;
            case RAYS_PROF_CONSTANT: m.ts[s] = R(c.t0s[s]); break;
            case RAYS_PROF_LINEAR:
                m.ts[s] = c.t0s[s] * (1.0 + x / p.LT_scale);
                m.gradts[0][s] = R(c.t0s[s] * (1.0 / p.LT_scale));
                break;
            case RAYS_PROF_LINEAR_2:
                m.ts[s] = c.t0s[s] + p.dtdx * (x - p.x0);
                m.gradts[0][s] = R(c.t0s[s] * p.dtdx);
                break;
            case RAYS_PROF_PARABOLIC: {
                R f, fp;
                parabolic_prof(x - p.x0, R(p.T_min[s]), R(p.alphat1[s]), R(p.alphat2[s]), f, fp);
                m.ts[s] = c.t0s[s] * f;
                m.gradts[0][s] = c.t0s[s] * fp;
            } break;
        }
    }
    if (minval(m.ns, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_DENS;
    if (minval(m.ts, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_TEMP;
}

// Solov'ev field + grad(B) tensor, shared text of L/solovev_eq_m.f90:165-190 and
// L/solovev_magnetics_m.f90:166-190
template <class R>
inline void solovev_field(R x, R y, R z, R r, double bphi0, double bp0, double rmaj, double kappa,
                          R bvec[3], R g[3][3]) {
    const double rk = rmaj * kappa;
    R br = -bp0 * r * z / (rk * rk);
    R zk = z / rk, rr = r / rmaj;
    R bz = bp0 * ((zk * zk) + .5 * ((rr * rr) - 1.0));
    R bphi = bphi0 * rmaj / r;
    R dbrdr = br / r;
    R dbrdz = -bp0 * r / (rk * rk);
    R dbzdr = bp0 * r / (rmaj * rmaj);
    R dbzdz = bp0 * 2.0 * z / (rk * rk);
    R dbphidr = -bphi / r;
    bvec[0] = br * x / r - bphi * y / r;
    bvec[1] = br * y / r + bphi * x / r;
    bvec[2] = bz;
    R r2 = r * r;
    g[0][0] = (dbrdr * (x * x) + br * (y * y) / r + (-dbphidr + bphi / r) * x * y) / r2;
    g[1][0] = ((dbrdr - br / r) * x * y - dbphidr * (y * y) - bphi * (x * x) / r) / r2;
    g[2][0] = dbrdz * x / r;
    g[0][1] = ((dbrdr - br / r) * x * y + dbphidr * (x * x) + bphi * (y * y) / r) / r2;
    g[1][1] = (dbrdr * (y * y) + br * (x * x) / r + (dbphidr - bphi / r) * x * y) / r2;
    g[2][1] = dbrdz * y / r;
    g[0][2] = dbzdr * x / r;
    g[1][2] = dbzdr * y / r;
    g[2][2] = dbzdz;
}
// solovev_psi / solovev_magnetics_psi (L/solovev_eq_m.f90:280-322, L/solovev_magnetics_m.f90:211-253)
template <class R>
inline void solovev_psi(const R rvec[3], double bphi0, double iota0, double rmaj, double kappa,
                        double psiB, R &psi, R gradpsi[3], R &psiN, R gradpsiN[3]) {
    R x = rvec[0], y = rvec[1], z = rvec[2];
    R Rr = Sqrt(x * x + y * y);
    double bp0 = bphi0 * iota0;
    const double rk = rmaj * kappa;
    R a = Rr * z / rk;
    R b = (Rr * Rr) - rmaj * rmaj;
    psi = .5 * bp0 * ((a * a) + ((b * b)) / (rmaj * rmaj) / 4.0);
    R br = -bp0 * Rr * z / (rk * rk);
    R zk = z / rk, rr = Rr / rmaj;
    R bz = bp0 * ((zk * zk) + .5 * ((rr * rr) - 1.0));
    gradpsi[0] = x * bz;
    gradpsi[1] = y * bz;
    gradpsi[2] = -Rr * br;
    psiN = psi / psiB;
    for (int i = 0; i < 3; ++i) gradpsiN[i] = gradpsi[i] / psiB;
}

// TEST SWITCH: the example figures of ECH_90GHz_solovev_SG_eq_plane predate RAYS_project; the older generation of the same code
// (RAYS_code/, OLD/ in SURVEY.md) differs from the current one in how a ray ENDS at the plasma edge:
//   * RAYS_code/solovev_eq_m.f90:140   if (psiN > 1.) equib_err = 'psi >1 out_of_plasma'   (the current solovev_eq has no such test),
//   * RAYS_code/ray_tracing.f90:131-140: a point whose check_save raised ANY flag ends the ray and is not counted
//     (npoints = nstep, :150-153), where the current loop tests stop_ode only.
// With the switch on the oracle ends rays like that generation did; tests/test_reference_plots.py uses it to show that the two
// figure rays that are one point shorter than the current code's were drawn by it.  Off (0) everywhere else.
inline int &oracle_old_generation() { static int flag = 0; return flag; }

// solovev_eq (L/solovev_eq_m.f90:122-276) including its temperature-profile bugs
template <class R> inline void solovev_eq(const rays_cfg &c, const R rvec[3], ModelOut<R> &m) {
    const rays_solovev_eq &p = c.solovev;
    const int nspec = c.nspec;
    zero_model(m);  // (X) intent(out) arrays are not zeroed by the Fortran; zero is the benign choice
    R x = rvec[0], y = rvec[1], z = rvec[2];
    R r = Sqrt(x * x + y * y);
    if (r < p.box_rmin || r > p.box_rmax) m.equib_err = RAYS_STOP_R_OUT_OF_BOX_SOLOVEV;
    if (z < p.box_zmin || z > p.box_zmax) m.equib_err = RAYS_STOP_Z_OUT_OF_BOX_SOLOVEV;
    double bp0 = p.bphi0 * p.iota0;
    R psi, gradpsi[3], psiN, gradpsiN[3];
    solovev_psi(rvec, p.bphi0, p.iota0, p.rmaj, p.kappa, p.psiB, psi, gradpsi, psiN, gradpsiN);
    if (oracle_old_generation() && val(psiN) > 1.0) m.equib_err = RAYS_STOP_OUT_OF_PLASMA;   // RAYS_code/solovev_eq_m.f90:140
    if (m.equib_err != 0) return;
    solovev_field(x, y, z, r, p.bphi0, bp0, p.rmaj, p.kappa, m.bvec, m.gradbtensor);
    switch (p.dens_prof_model) {
        case RAYS_PROF_CONSTANT:
            for (int s = 0; s <= nspec; ++s) m.ns[s] = R(c.n0s[s]);
            break;
        case RAYS_PROF_PARABOLIC:
            if (psiN < 1.0) {
                R a1(p.alphan1), a2(p.alphan2);
                R base = 1.0 - Pow(psiN, a2);
                R prof = Pow(base, a1);
                R dd_psi = -a1 * a2 * Pow(psiN, a2 - 1.0) * Pow(1.0 - Pow(psiN, a2), a1 - 1.0);
                for (int s = 0; s <= nspec; ++s) {
                    m.ns[s] = c.n0s[s] * prof;
                    for (int i = 0; i < 3; ++i) m.gradns[i][s] = c.n0s[s] * dd_psi * gradpsiN[i];
                }
            }
            break;
    }
    for (int s = 0; s <= nspec; ++s) {
        switch (p.t_prof_model[s]) {
            case RAYS_PROF_ZERO:
                m.ts[s] = R(0.0);
                for (int i = 0; i < 3; ++i) m.gradts[i][s] = R(0.0);
                break;
            case RAYS_PROF_CONSTANT:  // (R) sets DENSITY (solovev_eq_m.f90:243-245)
                for (int q = 0; q <= nspec; ++q) m.ns[q] = R(c.n0s[q]);
                for (int i = 0; i < 3; ++i) for (int q = 0; q < NS0; ++q) m.gradns[i][q] = R(0.0);
                break;
            case RAYS_PROF_PARABOLIC:  // (R) zeroes ALL species (:251-252), exponent alphat1 (:256-257)
                for (int q = 0; q < NS0; ++q) {
                    m.ts[q] = R(0.0);
                    for (int i = 0; i < 3; ++i) m.gradts[i][q] = R(0.0);
                }
                if (psiN < 1.0) {
                    R a1(p.alphat1[s]), a2(p.alphat2[s]);
                    m.ts[s] = c.t0s[s] * Pow(1.0 - Pow(psiN, a2), a1);
                    R dd_psi = -a1 * a2 * Pow(psiN, a2 - 1.0) * Pow(1.0 - Pow(psiN, a2), a1);
                    for (int i = 0; i < 3; ++i) m.gradts[i][s] = c.t0s[s] * dd_psi * gradpsiN[i];
                }
                break;
        }
    }
    if (minval(m.ns, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_DENS;
    if (minval(m.ts, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_TEMP;
}

// eqdsk_magnetics_spline_interp (L/eqdsk_magnetics_spline_interp_m.f90:206-282): psi(R,Z) as a bicubic spline of
// the g-file's grid (shifted to 0 on axis), R*Bphi as a cubic spline -- on the R grid, as the reference builds it
// (:184, (R)) -- B = grad(psi) x grad(phi) + RBphi grad(phi)
template <class R>
inline void eqdsk_magnetics(const rays_axisym_eq &p, R x, R y, R z, R r, R bvec[3], R g[3][3], R &psi, R gradpsi[3], R &psiN, R gradpsiN[3]) {
    R PsiR(0.0), PsiZ(0.0), PsiRR(0.0), PsiRZ(0.0), PsiZZ(0.0), RBphi(0.0), RBphiR(0.0);
    psi = R(0.0);
    bcspeval_fpp(r, z, p.Psi_spline, psi, PsiR, PsiZ, PsiRR, PsiZZ, PsiRZ);
    cspeval(r, p.T_spline, RBphi, &RBphiR);
    R br = PsiZ / r;
    R bz = -PsiR / r;
    R bphi = RBphi / r;
    gradpsi[0] = -x * bz; gradpsi[1] = -y * bz; gradpsi[2] = r * br;
    psiN = psi / p.eq_psibound;
    for (int i = 0; i < 3; ++i) gradpsiN[i] = gradpsi[i] / p.eq_psibound;
    R dbrdr = -br / r + PsiRZ / r;
    R dbrdz = PsiZZ / r;
    R dbzdr = -bz / r - PsiRR / r;
    R dbzdz = -PsiRZ / r;
    R dbphidr = (RBphiR - bphi) / r;
    bvec[0] = br * x / r - bphi * y / r;
    bvec[1] = br * y / r + bphi * x / r;
    bvec[2] = bz;
    g[0][0] = (dbrdr * (x * x) + br * (y * y) / r + (-dbphidr + bphi / r) * x * y) / (r * r);
    g[1][0] = ((dbrdr - br / r) * x * y - dbphidr * (y * y) - bphi * (x * x) / r) / (r * r);
    g[2][0] = dbrdz * x / r;
    g[0][1] = ((dbrdr - br / r) * x * y + dbphidr * (x * x) + bphi * (y * y) / r) / (r * r);
    g[1][1] = (dbrdr * (y * y) + br * (x * x) / r + (dbphidr - bphi / r) * x * y) / (r * r);
    g[2][1] = dbrdz * y / r;
    g[0][2] = dbzdr * x / r;
    g[1][2] = dbzdr * y / r;
    g[2][2] = dbzdz;
}

// axisym_toroid_eq + solovev_magnetics (L/axisym_toroid_eq_m.f90:215-362, L/solovev_magnetics_m.f90:124-207)
template <class R> inline void axisym_toroid_eq(const rays_cfg &c, const R rvec[3], ModelOut<R> &m) {
    const rays_axisym_eq &p = c.axisym;
    const int nspec = c.nspec;
    const double Tiny = 10.0e-14;
    zero_model(m);
    R x = rvec[0], y = rvec[1], z = rvec[2];
    R r = Sqrt(x * x + y * y);
    if (r < p.box_rmin - Tiny || r > p.box_rmax + Tiny) m.equib_err = RAYS_STOP_R_OUT_OF_BOX;
    if (z < p.box_zmin - Tiny || z > p.box_zmax + Tiny) m.equib_err = RAYS_STOP_Z_OUT_OF_BOX;
    if (m.equib_err != 0) return;
    R psi, gradpsi[3], psiN, grad_psiN[3];
    if (p.magnetics_model == RAYS_MAG_EQDSK_SPLINE) {
        eqdsk_magnetics(p, x, y, z, r, m.bvec, m.gradbtensor, psi, gradpsi, psiN, grad_psiN);
    } else {   // solovev_magnetics
        if (r < p.sm_box_rmin || r > p.sm_box_rmax) m.equib_err = RAYS_STOP_R_OUT_OF_BOUNDS_SOLMAG;
        if (z < p.sm_box_zmin || z > p.sm_box_zmax) m.equib_err = RAYS_STOP_Z_OUT_OF_BOUNDS_SOLMAG;
        if (m.equib_err != 0) {
            // (X) the Fortran returns with psiN undefined and carries on; the only defined outcome
            // is that an error is reported.  Stop here with the magnetics flag.
            return;
        }
        double bp0 = p.sm_bphi0 * p.sm_iota0;
        solovev_psi(rvec, p.sm_bphi0, p.sm_iota0, p.sm_rmaj, p.sm_kappa, p.sm_psiB, psi, gradpsi, psiN, grad_psiN);
        solovev_field(x, y, z, r, p.sm_bphi0, bp0, p.sm_rmaj, p.sm_kappa, m.bvec, m.gradbtensor);
    }
    if (psiN > p.plasma_psi_limit) m.equib_err = RAYS_STOP_OUT_OF_PLASMA;
    switch (p.density_prof_model) {
        case RAYS_PROF_CONSTANT:
            for (int s = 0; s <= nspec; ++s) m.ns[s] = R(c.n0s[s]);
            break;
        case RAYS_PROF_PARABOLIC: {
            R dens, dd_psi;
            parabolic_prof(psiN, R(p.d_scrape_off), R(p.alphan1), R(p.alphan2), dens, dd_psi);
            for (int s = 0; s <= nspec; ++s) {
                m.ns[s] = c.n0s[s] * dens;
                for (int i = 0; i < 3; ++i) m.gradns[i][s] = c.n0s[s] * dd_psi * grad_psiN[i];
            }
        } break;
        case RAYS_PROF_SPLINE: {   // density_spline_interp (L/density_spline_interp_m.f90:107-127)
            R dens(0.0), dd_psi(0.0);   // (X) dd_psi is undefined in the Fortran for psi > 1 with d_scrape_off = 0
            if (psiN <= 1.0) cspeval(psiN, p.ne_spline, dens, &dd_psi);
            if (dens < p.d_scrape_off) { dens = R(p.d_scrape_off); dd_psi = R(0.0); }
            for (int s = 0; s <= nspec; ++s) {
                m.ns[s] = c.n0s[s] * dens;
                for (int i = 0; i < 3; ++i) m.gradns[i][s] = c.n0s[s] * dd_psi * grad_psiN[i];
            }
        } break;
    }
    for (int s = 0; s <= nspec; ++s) {
        switch (p.temperature_prof_model[s]) {
            case RAYS_PROF_ZERO: break;
            case RAYS_PROF_CONSTANT:
                m.ts[s] = R(c.t0s[s]);
                for (int i = 0; i < 3; ++i) for (int q = 0; q < NS0; ++q) m.gradts[i][q] = R(0.0);  // gradts = 0. (all)
                break;
            case RAYS_PROF_PARABOLIC: {
                R t_prof, dt_dpsi;
                parabolic_prof(psiN, R(p.T_scrape_off), R(p.alphat1[s]), R(p.alphat2[s]), t_prof, dt_dpsi);
                m.ts[s] = c.t0s[s] * t_prof;
                for (int i = 0; i < 3; ++i) m.gradts[i][s] = c.t0s[s] * dt_dpsi * grad_psiN[i];
            } break;
            case RAYS_PROF_SPLINE: {   // temperature_spline_interp (L/temperature_spline_interp_m.f90:82-108): Te for s = 0, Ti else
                R t(0.0), dt(0.0);
                if (psiN <= 1.0) cspeval(psiN, s == 0 ? p.Te_spline : p.Ti_spline, t, &dt);
                if (t < p.T_scrape_off) { t = R(p.T_scrape_off); dt = R(0.0); }
                m.ts[s] = c.t0s[s] * t;
                for (int i = 0; i < 3; ++i) m.gradts[i][s] = c.t0s[s] * dt * grad_psiN[i];
            } break;
        }
    }
    if (minval(m.ns, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_DENS;
    if (minval(m.ts, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_TEMP;
}
// axisym_toroid_psi -> solovev_magnetics_psi (L/axisym_toroid_eq_m.f90:366-396)
template <class R>
inline void axisym_toroid_psi(const rays_cfg &c, const R rvec[3], R &psi, R gradpsi[3], R &psiN, R gradpsiN[3]) {
    const rays_axisym_eq &p = c.axisym;
    if (p.magnetics_model == RAYS_MAG_EQDSK_SPLINE) {   // eqdsk_magnetics_spline_interp_psi (L/eqdsk_magnetics_spline_interp_m.f90:286-318)
        R x = rvec[0], y = rvec[1], z = rvec[2];
        R r = Sqrt(x * x + y * y);
        R PsiR(0.0), PsiZ(0.0);
        psi = R(0.0);
        bcspeval_fp(r, z, p.Psi_spline, psi, PsiR, PsiZ);
        R br = PsiZ / r;
        R bz = -PsiR / r;
        gradpsi[0] = -x * bz; gradpsi[1] = -y * bz; gradpsi[2] = r * br;
        psiN = psi / p.eq_psibound;
        for (int i = 0; i < 3; ++i) gradpsiN[i] = gradpsi[i] / p.eq_psibound;
        return;
    }
    solovev_psi(rvec, p.sm_bphi0, p.sm_iota0, p.sm_rmaj, p.sm_kappa, p.sm_psiB, psi, gradpsi, psiN, gradpsiN);
}

// multiple_mirror_eq + mirror_magnetics_spline_interp
// (L/multiple_mirror_eq_m.f90:223-376, L/mirror_magnetics_spline_interp_m.f90:132-204)
template <class R> inline void multiple_mirror_eq(const rays_cfg &c, const R rvec[3], ModelOut<R> &m) {
    const rays_mirror_eq &p = c.mirror;
    const int nspec = c.nspec;
    zero_model(m);
    R x = rvec[0], y = rvec[1], z = rvec[2];
    R r = Sqrt(x * x + y * y);
    if (r > p.box_rmax) m.equib_err = RAYS_STOP_R_OUT_OF_BOX;
    if (z < p.box_zmin || z > p.box_zmax) m.equib_err = RAYS_STOP_Z_OUT_OF_BOX;
    if (m.equib_err != 0) return;
    R br, dbrdr, dbrdz, bz, dbzdr, dbzdz, Aphi, dAphidr, dAphidz;
    bcspeval_fp(r, z, p.Br_spline, br, dbrdr, dbrdz);
    bcspeval_fp(r, z, p.Bz_spline, bz, dbzdr, dbzdz);
    bcspeval_fp(r, z, p.Aphi_spline, Aphi, dAphidr, dAphidz);
    R gradAphi[3] = {R(0.0), R(0.0), R(0.0)};
    if (r < 2.0 * DBL_MIN) {  // on axis
        m.bvec[0] = R(0.0); m.bvec[1] = R(0.0); m.bvec[2] = bz;
        m.gradbtensor[0][0] = -dbzdz / 2.0;
        m.gradbtensor[1][1] = -dbzdz / 2.0;
        m.gradbtensor[2][2] = dbzdz;
        Aphi = R(0.0);
    } else {
        m.bvec[0] = x * br / r; m.bvec[1] = y * br / r; m.bvec[2] = bz;
        R xr = x / r, yr = y / r;
        m.gradbtensor[0][0] = (1.0 - (xr * xr)) * br / r + (xr * xr) * dbrdr;
        m.gradbtensor[1][0] = x * y / (r * r) * (dbrdr - br / r);
        m.gradbtensor[2][0] = dbrdz * x / r;
        m.gradbtensor[0][1] = x * y / (r * r) * (dbrdr - br / r);
        m.gradbtensor[1][1] = (1.0 - (yr * yr)) * br / r + (yr * yr) * dbrdr;
        m.gradbtensor[2][1] = dbrdz * y / r;
        m.gradbtensor[0][2] = dbzdr * x / r;
        m.gradbtensor[1][2] = dbzdr * y / r;
        m.gradbtensor[2][2] = dbzdz;
        gradAphi[0] = dAphidr * x / r;
        gradAphi[1] = dAphidr * y / r;
        gradAphi[2] = dAphidz;
    }
    R AphiN = Aphi / p.Aphi_LUFS;
    R gradAphiN[3];
    for (int i = 0; i < 3; ++i) gradAphiN[i] = gradAphi[i] / p.Aphi_LUFS;
    if (AphiN > p.plasma_AphiN_limit) m.equib_err = RAYS_STOP_OUT_OF_PLASMA;
    R dens(0.0), dd_rho(0.0);
    switch (p.density_prof_model) {
        case RAYS_PROF_CONSTANT:
            for (int s = 0; s <= nspec; ++s) m.ns[s] = R(c.n0s[s]);
            break;
        case RAYS_PROF_PARABOLIC:
            parabolic_prof(AphiN, R(p.d_scrape_off), R(p.alphan1), R(p.alphan2), dens, dd_rho);
            break;
        case RAYS_PROF_HYPERBOLIC:
            hyperbolic_prof(AphiN, R(p.d_scrape_off), R(p.AphiN0_d), R(p.delta_d), dens, dd_rho);
            break;
    }
    if (p.density_prof_model != RAYS_PROF_CONSTANT)
        for (int s = 0; s <= nspec; ++s) {
            m.ns[s] = c.n0s[s] * dens;
            for (int i = 0; i < 3; ++i) m.gradns[i][s] = c.n0s[s] * dd_rho * gradAphiN[i];
        }
    for (int s = 0; s <= nspec; ++s) {
        R t_prof, dt_drho;
        switch (p.temperature_prof_model[s]) {
            case RAYS_PROF_ZERO: break;
            case RAYS_PROF_CONSTANT:
                m.ts[s] = R(c.t0s[s]);
                for (int i = 0; i < 3; ++i) for (int q = 0; q < NS0; ++q) m.gradts[i][q] = R(0.0);
                break;
            case RAYS_PROF_PARABOLIC:
                parabolic_prof(AphiN, R(p.T_scrape_off), R(p.alphat1[s]), R(p.alphat2[s]), t_prof, dt_drho);
                m.ts[s] = c.t0s[s] * t_prof;
                for (int i = 0; i < 3; ++i) m.gradts[i][s] = c.t0s[s] * dt_drho * gradAphiN[i];
                break;
            case RAYS_PROF_HYPERBOLIC:
                hyperbolic_prof(AphiN, R(p.T_scrape_off), R(p.AphiN0_t[s]), R(p.delta_t[s]), t_prof, dt_drho);
                m.ts[s] = c.t0s[s] * t_prof;
                for (int i = 0; i < 3; ++i) m.gradts[i][s] = c.t0s[s] * dt_drho * gradAphiN[i];
                break;
        }
    }
    if (minval(m.ns, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_DENS;
    if (minval(m.ts, nspec + 1) < 0.0) m.equib_err = RAYS_STOP_NEGATIVE_TEMP;
}

// equilibrium (L/equilibrium_m.f90:135-272).  omgrf is an argument because deriv_num perturbs it
// (the Fortran mutates the module variable, L/deriv_num.f90:72-84).
template <class R> inline void equilibrium(const rays_cfg &c, const R rvec[3], R omgrf, EqPoint<R> &eq) {
    ModelOut<R> m;
    switch (c.equilib_model) {
        case RAYS_EQ_SLAB: slab_eq(c, rvec, m); break;
        case RAYS_EQ_SOLOVEV: solovev_eq(c, rvec, m); break;
        case RAYS_EQ_AXISYM_TOROID: axisym_toroid_eq(c, rvec, m); break;
        case RAYS_EQ_MULTIPLE_MIRROR: multiple_mirror_eq(c, rvec, m); break;
        default: zero_model(m); m.equib_err = -1;
    }
    eq.equib_err = m.equib_err;
    // (X) on error the Fortran leaves every other field undefined; zero them so nothing downstream is UB
    const int nspec = c.nspec;
    for (int i = 0; i < 3; ++i) {
        eq.bvec[i] = m.bvec[i];
        for (int j = 0; j < 3; ++j) eq.gradbtensor[i][j] = m.gradbtensor[i][j];
        for (int s = 0; s < NS0; ++s) { eq.gradns[i][s] = m.gradns[i][s]; eq.gradts[i][s] = m.gradts[i][s]; }
    }
    for (int s = 0; s < NS0; ++s) {
        eq.ns[s] = m.ns[s]; eq.ts[s] = m.ts[s];
        eq.omgc[s] = R(0.0); eq.omgp2[s] = R(0.0); eq.alpha[s] = R(0.0); eq.gamma[s] = R(0.0);
    }
    if (m.equib_err != 0) {
        eq.bmag = R(0.0);
        for (int i = 0; i < 3; ++i) { eq.bunit[i] = R(0.0); eq.gradbmag[i] = R(0.0); for (int j = 0; j < 3; ++j) eq.gradbunit[i][j] = R(0.0); }
        return;
    }
    R bmag = Sqrt(m.bvec[0] * m.bvec[0] + m.bvec[1] * m.bvec[1] + m.bvec[2] * m.bvec[2]);
    eq.bmag = bmag;
    for (int i = 0; i < 3; ++i) eq.bunit[i] = m.bvec[i] / bmag;
    for (int i = 0; i < 3; ++i)
        eq.gradbmag[i] = m.gradbtensor[i][0] * eq.bunit[0] + m.gradbtensor[i][1] * eq.bunit[1] + m.gradbtensor[i][2] * eq.bunit[2];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            eq.gradbunit[i][j] = (m.gradbtensor[i][j] - eq.gradbmag[i] * eq.bunit[j]) / bmag;
    for (int s = 0; s <= nspec; ++s) {
        eq.omgc[s] = c.qs[s] * bmag / c.ms[s];
        eq.omgp2[s] = m.ns[s] * (c.qs[s] * c.qs[s]) / (c.eps0 * c.ms[s]);
        eq.alpha[s] = eq.omgp2[s] / (omgrf * omgrf);
        eq.gamma[s] = eq.omgc[s] / omgrf;
    }
}

// ============================ susceptibility / dielectric ====================================
// suscep_cold + dielectric_cold (L/suscep_m.f90:53-86,142-176): eps(i,j) -> e[i-1][j-1]
template <class R> inline void dielectric_cold(const rays_cfg &c, const EqPoint<R> &eq, Cx<R> e[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) e[i][j] = Cx<R>(R(0.0), R(0.0));
    for (int s = 0; s <= c.nspec; ++s) {
        R al = eq.alpha[s], ga = eq.gamma[s];
        R den = 1.0 - ga * ga;
        Cx<R> chi11(-al / den, R(0.0));
        Cx<R> chi33(-al, R(0.0));
        // chi(1,2) = -zi*alphas*gammas/(1-gammas**2), zi = (0,1): ((-zi*al)*ga)/den = (0, -(al*ga)/den)
        Cx<R> chi12(R(0.0), -(al * ga) / den);
        Cx<R> chi21 = -chi12;
        e[0][0] = e[0][0] + chi11;
        e[1][1] = e[1][1] + chi11;
        e[2][2] = e[2][2] + chi33;
        e[0][1] = e[0][1] + chi12;
        e[1][0] = e[1][0] + chi21;
    }
    for (int i = 0; i < 3; ++i) e[i][i].re = e[i][i].re + 1.0;
}
// RLSDP_cold (L/suscep_m.f90:180-219)
template <class R> inline void RLSDP_cold(const rays_cfg &c, const EqPoint<R> &eq, R &S, R &D, R &P, R &Rr, R &L) {
    Rr = R(0.0); L = R(0.0); S = R(0.0); D = R(0.0); P = R(0.0);
    for (int s = 0; s <= c.nspec; ++s) {
        R al = eq.alpha[s], ga = eq.gamma[s];
        Rr = Rr - al / (1.0 + ga);
        L = L - al / (1.0 - ga);
        S = S - al / (1.0 - ga * ga);
        D = D - al * ga / (1.0 - ga * ga);
        P = P - al;
    }
    Rr = 1.0 + Rr;
    L = 1.0 + L;
    S = (Rr + L) / 2.0;
    D = (Rr - L) / 2.0;
    P = 1.0 + P;
}

// det( eps_h + n n - n^2 I ) with n = (n1,0,n3): common text of `residual` (L/check_save.f90:163-235)
// and `determ` (L/deriv_num.f90:99-153)
template <class R>
inline Cx<R> disp_det(const rays_cfg &c, const EqPoint<R> &eq, R n1, R n3, Cx<R> eps_h[3][3], R n[3]) {
    Cx<R> eps[3][3];
    dielectric_cold(c, eq, eps);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) eps_h[i][j] = cscale(R(.5), eps[i][j] + conjg(eps[j][i]));
    n[0] = n1; n[1] = R(0.0); n[2] = n3;
    R nsq = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    Cx<R> epsn[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            R nn = n[i] * n[j];
            R kron = R(i == j ? 1.0 : 0.0);
            // eps_h(i,j) + n(i)*n(j) - delta*sum(n**2): complex + real - real
            epsn[i][j] = Cx<R>((eps_h[i][j].re + nn) - kron * nsq, eps_h[i][j].im);
        }
#define E(i, j) epsn[(i)-1][(j)-1]
    Cx<R> ctmp = E(3, 3) * (E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2)) -
                 E(3, 2) * (E(1, 1) * E(2, 3) - E(2, 1) * E(1, 3)) +
                 E(3, 1) * (E(1, 2) * E(2, 3) - E(2, 2) * E(1, 3));
#undef E
    return ctmp;
}

// ============================ deriv_cold (L/deriv_cold.f90:1-228) ============================
template <class R>
inline void deriv_cold(const rays_cfg &c, const EqPoint<R> &eq, const R nvec[3], R dddx[3], R dddk[3], R &dddw) {
    const int ns = c.nspec + 1;
    const double omgrf = c.omgrf, k0 = c.k0;
    R alpha[NS0], gamma[NS0];
    for (int s = 0; s < ns; ++s) { alpha[s] = eq.alpha[s]; gamma[s] = eq.gamma[s]; }
    R n3 = nvec[0] * eq.bunit[0] + nvec[1] * eq.bunit[1] + nvec[2] * eq.bunit[2];
    R d0 = nvec[0] - n3 * eq.bunit[0], d1 = nvec[1] - n3 * eq.bunit[1], d2 = nvec[2] - n3 * eq.bunit[2];
    R n1 = Sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    R dn3dk[3], dn12dk[3], dn3dx[3], dn12dx[3];
    for (int i = 0; i < 3; ++i) dn3dk[i] = eq.bunit[i] / k0;
    dn12dk[0] = (2.0 / k0) * d0; dn12dk[1] = (2.0 / k0) * d1; dn12dk[2] = (2.0 / k0) * d2;
    for (int i = 0; i < 3; ++i)
        dn3dx[i] = eq.gradbunit[i][0] * nvec[0] + eq.gradbunit[i][1] * nvec[1] + eq.gradbunit[i][2] * nvec[2];
    for (int i = 0; i < 3; ++i) dn12dx[i] = -2.0 * n3 * dn3dx[i];
    R dadx[3][NS0], dgdx[3][NS0];
    for (int i = 0; i < 3; ++i)
        for (int s = 0; s < ns; ++s) {
            dadx[i][s] = eq.alpha[s] * eq.gradns[i][s] / eq.ns[s];
            dgdx[i][s] = gamma[s] * eq.gradbmag[i] / eq.bmag;
        }
    R dn3dw = -n3 / omgrf;
    R dn12dw = (-2.0 / omgrf) * (n1 * n1);
    R dadw[NS0], dgdw[NS0];
    for (int s = 0; s < ns; ++s) { dadw[s] = -(2.0 / omgrf * alpha[s]); dgdw[s] = -(1.0 / omgrf * gamma[s]); }
    R sa(0.0);
    for (int s = 0; s < ns; ++s) sa = sa + alpha[s];
    R p = 1.0 - sa;
    R t(1.0);
    for (int s = 0; s < ns; ++s) t = t * (1.0 - gamma[s] * gamma[s]);
    R dq1da[NS0], dq2da[NS0];
    for (int s1 = 0; s1 < ns; ++s1) {
        dq1da[s1] = R(1.0); dq2da[s1] = R(1.0);
        for (int s = 0; s < ns; ++s)
            if (s != s1) { dq1da[s1] = dq1da[s1] * (1.0 + gamma[s]); dq2da[s1] = dq2da[s1] * (1.0 - gamma[s]); }
    }
    R q1(0.0), q2(0.0), su(0.0);
    for (int s = 0; s < ns; ++s) q1 = q1 + alpha[s] * dq1da[s];
    for (int s = 0; s < ns; ++s) q2 = q2 + alpha[s] * dq2da[s];
    for (int s = 0; s < ns; ++s) su = su + alpha[s] * dq1da[s] * dq2da[s];
    R u = t - su;
    R q = 2.0 * u - t + q1 * q2;
    R n1sq = n1 * n1, n3sq = n3 * n3;
    R n3p4 = n3sq * n3sq, n1p4 = n1sq * n1sq;  // x**4 = (x*x)*(x*x)
    R duda[NS0], dqda[NS0], ddda[NS0];
    for (int s = 0; s < ns; ++s) {
        duda[s] = -dq1da[s] * dq2da[s];
        dqda[s] = 2.0 * duda[s] + dq1da[s] * q2 + q1 * dq2da[s];
        ddda[s] = -t * n3p4 + (2.0 * (u - p * duda[s]) + (-t + duda[s]) * n1sq) * n3sq - q + p * dqda[s] -
                  (dqda[s] - u + p * duda[s]) * n1sq + duda[s] * n1p4;
    }
    R gp[NS0][NS0], gm[NS0][NS0], gpm[NS0][NS0];
    for (int s1 = 0; s1 < ns; ++s1)
        for (int s2 = 0; s2 < ns; ++s2) {
            gp[s1][s2] = R(1.0); gm[s1][s2] = R(1.0);
            for (int s = 0; s < ns; ++s)
                if (s != s1 && s != s2) { gp[s1][s2] = gp[s1][s2] * (1.0 + gamma[s]); gm[s1][s2] = gm[s1][s2] * (1.0 - gamma[s]); }
        }
    for (int s1 = 0; s1 < ns; ++s1) for (int s2 = 0; s2 < ns; ++s2) gpm[s1][s2] = gp[s1][s2] * gm[s1][s2];
    R dtdg[NS0], dudg[NS0], dq1dg[NS0], dq2dg[NS0], dqdg[NS0], dddg[NS0];
    for (int s = 0; s < ns; ++s) dtdg[s] = 2.0 * gamma[s] * duda[s];
    for (int s = 0; s < ns; ++s) {
        R a(0.0);
        for (int s1 = 0; s1 < ns; ++s1) a = a + alpha[s1] * gpm[s1][s];
        dudg[s] = a;
    }
    for (int s = 0; s < ns; ++s) dudg[s] = dtdg[s] + 2.0 * gamma[s] * (dudg[s] + alpha[s] * duda[s]);
    for (int s = 0; s < ns; ++s) {
        R a(0.0);
        for (int s1 = 0; s1 < ns; ++s1) a = a + alpha[s1] * gp[s1][s];
        dq1dg[s] = a;
    }
    for (int s = 0; s < ns; ++s) dq1dg[s] = dq1dg[s] - alpha[s] * dq1da[s];
    for (int s = 0; s < ns; ++s) {
        R a(0.0);
        for (int s1 = 0; s1 < ns; ++s1) a = a + alpha[s1] * gm[s1][s];
        dq2dg[s] = a;
    }
    for (int s = 0; s < ns; ++s) dq2dg[s] = -dq2dg[s] + alpha[s] * dq2da[s];
    for (int s = 0; s < ns; ++s) {
        dqdg[s] = 2.0 * dudg[s] - dtdg[s] + dq1dg[s] * q2 + q1 * dq2dg[s];
        dddg[s] = dtdg[s] * p * n3p4 + (-2.0 * p * dudg[s] + (dtdg[s] * p + dudg[s]) * n1sq) * n3sq + p * dqdg[s] -
                  (dqdg[s] + p * dudg[s]) * n1sq + dudg[s] * n1p4;
    }
    R dddn3 = (4.0 * t * p * n3sq + 2.0 * (-2.0 * p * u + (t * p + u) * n1sq)) * n3;
    R dddn12 = (t * p + u) * n3sq - (q + p * u) + 2.0 * u * n1sq;
    for (int i = 0; i < 3; ++i) dddk[i] = dddn3 * dn3dk[i] + dddn12 * dn12dk[i];
    for (int i = 0; i < 3; ++i) {
        R a(0.0);
        for (int s = 0; s < ns; ++s) a = a + (ddda[s] * dadx[i][s] + dddg[s] * dgdx[i][s]);
        dddx[i] = a;
    }
    for (int i = 0; i < 3; ++i) dddx[i] = dddx[i] + dddn3 * dn3dx[i] + dddn12 * dn12dx[i];
    R a(0.0);
    for (int s = 0; s < ns; ++s) a = a + (ddda[s] * dadw[s] + dddg[s] * dgdw[s]);
    dddw = a + dddn3 * dn3dw + dddn12 * dn12dw;
}

// ============================ deriv_num (L/deriv_num.f90:1-156) ==============================
// determ(eq) with host-associated kvec, k0
template <class R>
inline R determ(const rays_cfg &c, const EqPoint<R> &eq, const R kvec[3], R k0, int &run_error) {
    R k3 = kvec[0] * eq.bunit[0] + kvec[1] * eq.bunit[1] + kvec[2] * eq.bunit[2];
    R d0 = kvec[0] - k3 * eq.bunit[0], d1 = kvec[1] - k3 * eq.bunit[1], d2 = kvec[2] - k3 * eq.bunit[2];
    R k1 = Sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    Cx<R> eps_h[3][3];
    R n[3];
    Cx<R> ctmp = disp_det(c, eq, k1 / k0, k3 / k0, eps_h, n);
    if (Fabs(ctmp.im) > f32lit(1.e-7)) run_error = RAYS_ERR_IM_DET;
    R prod(1.0);
    for (int s = 0; s < NS0; ++s) prod = prod * (1.0 - eq.gamma[s] * eq.gamma[s]);  // product(1.-eq%gamma**2) over 0:nspec0
    return ctmp.re * prod;
}
template <class R>
inline void deriv_num(const rays_cfg &c, const EqPoint<R> &eq0, const R *v, R dddx[3], R dddk[3], R &dddw,
                      int &pert_err, int &run_error) {
    R rvec0[3] = {v[0], v[1], v[2]}, kvec0[3] = {v[3], v[4], v[5]};
    const double omgrf0 = c.omgrf;
    const double delta = f32lit(1.e-6);
    R kvec[3] = {kvec0[0], kvec0[1], kvec0[2]};
    EqPoint<R> eq_plus, eq_minus;
    pert_err = 0;
    for (int i = 0; i < 3; ++i) {
        R rvec[3] = {rvec0[0], rvec0[1], rvec0[2]};
        R change(delta);
        rvec[i] = rvec0[i] + change;
        equilibrium(c, rvec, R(omgrf0), eq_plus);
        rvec[i] = rvec0[i] - change;
        equilibrium(c, rvec, R(omgrf0), eq_minus);
        // (X) the Fortran ignores equib_err of the displaced points and differentiates garbage;
        // here the ray is stopped with that flag (same choice in the CUDA path)
        if (eq_plus.equib_err && !pert_err) pert_err = eq_plus.equib_err;
        if (eq_minus.equib_err && !pert_err) pert_err = eq_minus.equib_err;
        R det_plus = determ(c, eq_plus, kvec, R(c.k0), run_error);
        R det_minus = determ(c, eq_minus, kvec, R(c.k0), run_error);
        dddx[i] = (det_plus - det_minus) / (2.0 * change);
    }
    for (int i = 0; i < 3; ++i) {
        kvec[0] = kvec0[0]; kvec[1] = kvec0[1]; kvec[2] = kvec0[2];
        R change = Fmax(R(delta), Fabs(delta * kvec[i])) / 2.0;
        kvec[i] = kvec0[i] + change;
        R det_plus = determ(c, eq0, kvec, R(c.k0), run_error);
        kvec[i] = kvec0[i] - change;
        R det_minus = determ(c, eq0, kvec, R(c.k0), run_error);
        dddk[i] = (det_plus - det_minus) / (2.0 * change);
    }
    kvec[0] = kvec0[0]; kvec[1] = kvec0[1]; kvec[2] = kvec0[2];
    double omg = omgrf0 * (1.0 + delta / 2.0);
    double k0p = omg / c.clight;
    equilibrium(c, rvec0, R(omg), eq_plus);
    R det_plus = determ(c, eq_plus, kvec, R(k0p), run_error);
    omg = omgrf0 * (1.0 - delta / 2.0);
    k0p = omg / c.clight;
    equilibrium(c, rvec0, R(omg), eq_minus);
    R det_minus = determ(c, eq_minus, kvec, R(k0p), run_error);
    dddw = (det_plus - det_minus) / (omgrf0 * delta);
}

// ============================ damping (L/damping_m.f90:74-117, L/damp_fund_ECH.f90:2-128) =====
// zfun0_real_arg_D -> zfun_real_arg_spline_D (M/zfunctions_m.f90:351-432); |z| <= 5 always here
template <class R> inline Cx<R> zfun_real_arg_spline(const rays_cfg &c, R z) {
    const double sqrt_pi = std::sqrt(std::atan2(0.0, -1.0));  // pi = atan(zero,-one) true double pi (:16)
    R re(0.0);
    if (Fabs(z) <= 10.0) {
        cspeval(z, c.zfun_re, re);
    } else {
        static const double A[6] = {1.0, 1.0 / 2.0, 3.0 / 4.0, 15.0 / 8.0, 105.0 / 16.0, 945.0 / 32.0};
        R z_inv = 1.0 / z;
        for (int i = 1; i <= 6; ++i) {
            R pw(1.0);
            for (int k = 0; k < 2 * i - 1; ++k) pw = pw * z_inv;
            re = re - pw * A[i - 1];
        }
    }
    R im = sqrt_pi * Exp(-(z * z));
    return Cx<R>(re, im);
}
template <class R> inline Cx<R> zfun0_real_arg(const rays_cfg &c, R z, R kz) {
    if (kz > 0.0) return zfun_real_arg_spline(c, z);
    return -zfun_real_arg_spline(c, -z);  // kz < 0 (kz == 0 excluded by the caller)
}
inline float f32(double x) { return (float)x; }
template <class R>
inline void damp_fund_ECH(const rays_cfg &c, const EqPoint<R> &eq, const R *v_kx, const R vg[3], R ksi[NS0], R &ki) {
    const int nspec = c.nspec;
    const double k0 = c.k0, omgrf = c.omgrf, clight = c.clight;
    for (int s = 0; s <= nspec; ++s) ksi[s] = R(0.0);
    ki = ksi[0];
    R kvec[3] = {v_kx[3], v_kx[4], v_kx[5]};
    R nvec[3] = {kvec[0] / k0, kvec[1] / k0, kvec[2] / k0};
    R k3 = kvec[0] * eq.bunit[0] + kvec[1] * eq.bunit[1] + kvec[2] * eq.bunit[2];
    R d0 = kvec[0] - k3 * eq.bunit[0], d1 = kvec[1] - k3 * eq.bunit[1], d2 = kvec[2] - k3 * eq.bunit[2];
    R k1 = Sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    R R3 = k3 / k0, R1 = k1 / k0;
    R R1S = R1 * R1, R3S = R3 * R3, RS = R1S + R3S;
    R B1 = eq.gamma[0];
    R BETAE = B1 * B1;
    if (R3 == 0.0) return;
    R vth = Sqrt(2.0 * eq.ts[0] / c.ms[0]);
    R VT = vth / clight;
    R xi = (omgrf + eq.omgc[0]) / (k3 * vth);
    if (Fabs(xi) > 5.0) return;
    Cx<R> zf = zfun0_real_arg(c, xi, k3);
    R P = eq.alpha[0];
    R Q = P / 2.0 / (1.0 - B1);
    R LAMBDA1 = (1.0 - Q) * RS * R1S + (1.0 - P) * RS * R3S - (1.0 - Q) * (1.0 - P) * (RS + R3S) -
                (1.0 - 2.0 * Q) * R1S + (1.0 - 2.0 * Q) * (1.0 - P);
    R LAMBDA2 = -P / B1 * (RS * R1S - (1.0 - 2.0 * Q) * R1S) +
                (P * P) / 4.0 / BETAE * R1S / R3S * (RS + R3S - 2.0 * (1.0 - 2.0 * Q));
    R LAMBDA5 = P * (RS * R3S - (1.0 - Q) * (RS + R3S) + (1.0 - 2.0 * Q));
    // D_WARM = real_expr * (xi + 1./zf); COMPLEX D_WARM is SINGLE precision (damp_fund_ECH.f90:36)
    R fac = -(1.0 - B1) * R3 * VT * (LAMBDA1 + LAMBDA2 + R1S / 2.0 / R3 / BETAE * VT * xi * LAMBDA5);
    Cx<R> inv = Cx<R>(R(1.0), R(0.0)) / zf;          // 1./zf : real -> complex, Smith division
    Cx<R> par(xi + inv.re, inv.im);                 // xi + complex
    Cx<R> dw(fac * par.re, fac * par.im);           // real * complex
    float dwr = f32(val(dw.re)), dwi = f32(val(dw.im));
    R A = 1.0 - P - BETAE;
    R B = -((1.0 - P) * A + (1.0 - P) * (1.0 - P) - BETAE) + (A + (1.0 - P) * (1.0 - BETAE)) * R3S;
    (void)B;
    R DDNX2 = 2.0 * A * R1S + B;
    R DDNZ = 2.0 * R3 * ((A + (1.0 - P) * (1.0 - BETAE)) * R1S + (1.0 - P) * (2.0 * (1.0 - BETAE) * R3S - 2.0 * A));
    R DDN[3];
    for (int i = 0; i < 3; ++i) {
        R dnperp2 = 2.0 * (nvec[i] - R3 * eq.bunit[i]);
        DDN[i] = DDNX2 * dnperp2 + DDNZ * eq.bunit[i];
    }
    R vgn = Sqrt(vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]);
    R vgu[3] = {vg[0] / vgn, vg[1] / vgn, vg[2] / vgn};
    R dot = DDN[0] * vgu[0] + DDN[1] * vgu[1] + DDN[2] * vgu[2];
    // DELTA = -D_WARM / dot : single complex promoted to double, divided by a double real
    // (promoted to complex -> Smith division with zero imaginary part), rounded to single again
    Cx<R> num(R(-(double)dwr), R(-(double)dwi));
    Cx<R> delta = num / Cx<R>(dot, R(0.0));
    float deli = f32(val(delta.im));
    ksi[0] = k0 * (double)deli;
    for (int s = 1; s <= nspec; ++s) ksi[s] = R(0.0);
    ki = ksi[0];
}
template <class R>
inline void damping(const rays_cfg &c, const EqPoint<R> &eq, const R *v, const R vg[3], R ksi[NS0], R &ki) {
    if (c.damping_model == RAYS_DAMP_FUND_ECH) damp_fund_ECH(c, eq, v, vg, ksi, ki);
    else { for (int s = 0; s <= c.nspec; ++s) ksi[s] = R(0.0); ki = R(0.0); }
}

// ============================ eqn_ray (L/eqn_ray.f90:1-236) ==================================
template <class R> inline void eqn_ray(const rays_cfg &c, R s, const R *v, R *dvds, OdeStop<R> &ray_stop) {
    (void)s;
    const int nspec = c.nspec;
    const double k0 = c.k0;
    R rvec[3] = {v[0], v[1], v[2]}, kvec[3] = {v[3], v[4], v[5]};
    R nvec[3] = {kvec[0] / k0, kvec[1] / k0, kvec[2] / k0};
    EqPoint<R> eq;
    equilibrium(c, rvec, R(c.omgrf), eq);
    if (eq.equib_err != 0) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = eq.equib_err; return; }
    R dddx[3], dddk[3], dddw;
    if (c.ray_deriv == RAYS_DERIV_COLD) deriv_cold(c, eq, nvec, dddx, dddk, dddw);
    else {
        int pert_err = 0;
        deriv_num(c, eq, v, dddx, dddk, dddw, pert_err, ray_stop.run_error);
        if (pert_err) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = pert_err; return; }
    }
    R vg[3], vg0, vg_unit[3];
    if (dddw != 0.0) {
        for (int i = 0; i < 3; ++i) vg[i] = -dddk[i] / dddw;
        vg0 = Sqrt(vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]);
        for (int i = 0; i < 3; ++i) vg_unit[i] = vg[i] / vg0;
    } else { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_INFINITE_VG_RHS; return; }
    R dsd_ray_param;
    if (c.ray_param == RAYS_PARAM_ARCL) {
        if (dddk[0] != 0.0 || dddk[1] != 0.0 || dddk[2] != 0.0) {
            R sg = Copysign(R(1.0), dddw);
            for (int i = 0; i < 3; ++i) {
                dvds[i] = -sg * dddk[i] / Sqrt(dddk[0] * dddk[0] + dddk[1] * dddk[1] + dddk[2] * dddk[2]);
                dvds[3 + i] = sg * dddx[i] / Sqrt(dddk[0] * dddk[0] + dddk[1] * dddk[1] + dddk[2] * dddk[2]);
            }
            dsd_ray_param = R(1.0);
        } else { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_RAY_STALLED; return; }
    } else {
        for (int i = 0; i < 3; ++i) { dvds[i] = -dddk[i] / dddw; dvds[3 + i] = dddx[i] / dddw; }
        dsd_ray_param = vg0;
    }
    dvds[6] = dsd_ray_param;
    int nv0 = 7;  // 1-based count of filled slots
    if (c.damping_model != RAYS_DAMP_NONE) {
        R ksi[NS0], ki;
        damping(c, eq, v, vg, ksi, ki);
        nv0 = nv0 + 1;
        dvds[nv0 - 1] = dsd_ray_param * 2.0 * ki * (1.0 - v[nv0 - 1]);
        if (c.multi_spec_damping) {
            for (int is = 0; is <= nspec; ++is) dvds[nv0 + is] = dsd_ray_param * 2.0 * ksi[is] * (1.0 - v[nv0 - 1]);
            nv0 = nv0 + 1 + nspec;
        }
    }
    if (c.integrate_eq_gradients) {
        for (int j = 0; j < 3; ++j)
            dvds[nv0 + j] = dsd_ray_param * vg_unit[0] * eq.gradbtensor[0][j] + dsd_ray_param * vg_unit[1] * eq.gradbtensor[1][j] +
                            dsd_ray_param * vg_unit[2] * eq.gradbtensor[2][j];
        dvds[nv0 + 3] = dsd_ray_param * vg_unit[0] * eq.gradns[0][0] + dsd_ray_param * vg_unit[1] * eq.gradns[1][0] +
                        dsd_ray_param * vg_unit[2] * eq.gradns[2][0];
        dvds[nv0 + 4] = dsd_ray_param * vg_unit[0] * eq.gradts[0][0] + dsd_ray_param * vg_unit[1] * eq.gradts[1][0] +
                        dsd_ray_param * vg_unit[2] * eq.gradts[2][0];
    }
}

// ============================ check_save (L/check_save.f90:1-161 + residual :163-235) =========
template <class R> inline void check_save(const rays_cfg &c, R s, const R *v, R &resid, OdeStop<R> &ray_stop) {
    (void)s;
    const double k0 = c.k0;
    R rvec[3] = {v[0], v[1], v[2]};
    EqPoint<R> eq;
    equilibrium(c, rvec, R(c.omgrf), eq);
    if (eq.equib_err != 0) {
        // (X) SURVEY A.5: flag set, stop NOT set, eq undefined.  Deterministic choice: residual 0,
        // no residual/Vg tests; the absorption test below does not depend on eq and is kept.
        ray_stop.ode_stop_flag = eq.equib_err;
        resid = R(0.0);
        if (c.damping_model != RAYS_DAMP_NONE && v[7] > c.total_damping_limit) {
            ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_TOTAL_ABSORPTION;
        }
        return;
    }
    R kvec[3] = {v[3], v[4], v[5]};
    R k3 = kvec[0] * eq.bunit[0] + kvec[1] * eq.bunit[1] + kvec[2] * eq.bunit[2];
    R d0 = kvec[0] - k3 * eq.bunit[0], d1 = kvec[1] - k3 * eq.bunit[1], d2 = kvec[2] - k3 * eq.bunit[2];
    R k1 = Sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    R nvec[3] = {kvec[0] / k0, kvec[1] / k0, kvec[2] / k0};
    {   // residual(eq,k1,k3)
        Cx<R> eps_h[3][3];
        R n[3];
        Cx<R> ctmp = disp_det(c, eq, k1 / k0, k3 / k0, eps_h, n);
        if (Fabs(ctmp.im) > f32lit(1.e-6)) ray_stop.run_error = RAYS_ERR_IM_DET;
        R en[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) en[i][j] = cabs(eps_h[i][j]) + Fabs(n[i] * n[j]);
#define N(i, j) en[(i)-1][(j)-1]
        R den = N(3, 3) * (N(1, 1) * N(2, 2)) + N(3, 3) * (N(2, 1) * N(1, 2)) + N(3, 2) * (N(1, 1) * N(2, 3)) +
                N(3, 2) * (N(2, 1) * N(1, 3)) + N(3, 1) * (N(1, 2) * N(2, 3)) + N(3, 1) * (N(2, 2) * N(1, 3));
#undef N
        // abs(ctmp) / complex(den,0): Smith division with a zero imaginary part == real division
        resid = cabs(ctmp) / den;
    }
    if (resid > c.dispersion_resid_limit) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_DISP_RESIDUAL; }
    R dddx[3], dddk[3], dddw;
    deriv_cold(c, eq, nvec, dddx, dddk, dddw);  // (R) always cold (check_save.f90:79-87)
    R vg[3] = {R(0.0), R(0.0), R(0.0)};
    if (Fabs(dddw) > DBL_MIN) {
        for (int i = 0; i < 3; ++i) vg[i] = -dddk[i] / dddw;
    } else { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_INFINITE_VG_CHECK; }
    if (c.damping_model != RAYS_DAMP_NONE) {
        // damping() result is unused here (diagnostic only) -> not evaluated
        R total_absorption = v[7];
        if (total_absorption > c.total_damping_limit) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_TOTAL_ABSORPTION; }
    }
}

// ============================ initialize_ode_vector (L/initialize_ode_vector.f90:1-57) ========
template <class R> inline void initialize_ode_vector(const rays_cfg &c, const double *rvec0, const double *rindex_vec0, R *v) {
    for (int i = 0; i < 3; ++i) { v[i] = R(rvec0[i]); v[3 + i] = R(c.k0 * rindex_vec0[i]); }
    v[6] = R(0.0);
    int nv0 = 7;
    if (c.damping_model != RAYS_DAMP_NONE) {
        v[7] = R(0.0); nv0 = 8;
        if (c.multi_spec_damping) { for (int s = 0; s <= c.nspec; ++s) v[8 + s] = R(0.0); nv0 = nv0 + 1 + c.nspec; }
    }
    if (c.integrate_eq_gradients) {
        EqPoint<R> eq;
        R r[3] = {v[0], v[1], v[2]};
        equilibrium(c, r, R(c.omgrf), eq);
        for (int i = 0; i < 3; ++i) v[nv0 + i] = eq.bvec[i];
        v[nv0 + 3] = eq.ns[0];
        v[nv0 + 4] = eq.ts[0];
    }
}

// ============================ RK4_ode (L/RK4_ode_m.f90:59-94) ================================
template <class R> inline void RK4_ode(const rays_cfg &c, R *v, R &s, R sout, OdeStop<R> &ray_stop) {
    const int nv = c.nv;
    R ds = sout - s;
    R f1[RAYS_NV_MAX], f2[RAYS_NV_MAX], f3[RAYS_NV_MAX], f4[RAYS_NV_MAX], w[RAYS_NV_MAX];
    eqn_ray(c, s, v, f1, ray_stop);
    if (ray_stop.stop_ode) return;
    for (int i = 0; i < nv; ++i) w[i] = v[i] + ds * f1[i] / 2.0;
    eqn_ray(c, s + ds / 2.0, w, f2, ray_stop);
    if (ray_stop.stop_ode) return;
    for (int i = 0; i < nv; ++i) w[i] = v[i] + ds * f2[i] / 2.0;
    eqn_ray(c, s + ds / 2.0, w, f3, ray_stop);
    if (ray_stop.stop_ode) return;
    for (int i = 0; i < nv; ++i) w[i] = v[i] + ds * f3[i];
    eqn_ray(c, s + ds, w, f4, ray_stop);
    if (ray_stop.stop_ode) return;
    for (int i = 0; i < nv; ++i) v[i] = v[i] + ds * (f1[i] + 2.0 * f2[i] + 2.0 * f3[i] + f4[i]) / 6.0;
    s = sout;
}

}  // namespace rays_oracle

#include "rays_oracle_sg.hpp"
#include "rays_oracle_trace.hpp"
