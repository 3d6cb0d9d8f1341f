// rays_oracle_ox.hpp — CPU ORACLE (test infrastructure only, see rays_oracle.hpp) for the O-X mode conversion analysis of
// stored trajectories: P/OX_conv_analysis_m.f90:91-198 (analyze_OX_conv), :202-252 (find_x_max_ray), :256-311
// (find_x_cutoff_ray), :315-407 (OX_conv_coeff)   (P/ = RAYS_project/post_process_lib/).
// PARITY STATUS: unpinned by reference output (the example that runs it ships no numbers).
// Notes: norm2 is evaluated as sqrt(a*a + b*b + c*c) (libgfortran uses a scaled sum: last-bit differences);
// (X) find_x_max_ray leaves x_max/k_max undefined and analyze_OX_conv reads found_cutoff uninitialised when no maximum is
// found: zeros / false here.  equilibrium() forms alpha, gamma, bunit whatever equib_err says (L/equilibrium_m.f90:229-268):
// a saved point outside the plasma boundary (every model output defined) is evaluated like that; outside the box (model
// outputs undefined in the Fortran) alpha = 0.
#pragma once
#include <cmath>

#include "rays_oracle.hpp"

namespace rays_oracle {

struct OxEq { double alpha0, gamma0, ns0, gradns0[3], bunit[3]; };

inline void ox_equilibrium(const rays_cfg &c, const double rvec[3], OxEq &q) {
    EqPoint<double> eq;
    equilibrium<double>(c, rvec, c.omgrf, eq);
    const int err = eq.equib_err;
    const bool hard = err != 0 && err != RAYS_STOP_OUT_OF_PLASMA && err != RAYS_STOP_NEGATIVE_DENS && err != RAYS_STOP_NEGATIVE_TEMP;
    q = OxEq{};
    if (hard) return;
    const double bmag = std::sqrt(eq.bvec[0] * eq.bvec[0] + eq.bvec[1] * eq.bvec[1] + eq.bvec[2] * eq.bvec[2]);
    for (int i = 0; i < 3; ++i) { q.bunit[i] = eq.bvec[i] / bmag; q.gradns0[i] = eq.gradns[i][0]; }
    q.ns0 = eq.ns[0];
    const double omgc = c.qs[0] * bmag / c.ms[0];
    const double omgp2 = eq.ns[0] * (c.qs[0] * c.qs[0]) / (c.eps0 * c.ms[0]);
    q.alpha0 = omgp2 / (c.omgrf * c.omgrf);
    q.gamma0 = omgc / c.omgrf;
}
inline double ox_norm2(const double a[3]) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// analyze_OX_conv for one ray: v = ray_vec(:, :, iray) as [ip*nv + iv]
inline void ox_conv_ray(const rays_cfg &c, const double *v, int nv, int npoints, int ray_number, rays_ox_conv &o) {
    const double zero = 0.0, one = 1.0, two = 2.0;
    const double pi = (double)3.1415926535897932385f;              // constants_m
    const double conversion_threshold = (double)0.0001f;            // default-real parameter (:27)
    o = rays_ox_conv{};
    o.ray_number = ray_number;
    // ---- find_x_max_ray
    bool found_max = false;
    int step_number = 1;
    double alpha_max = zero, x_max[3] = {0, 0, 0}, k_max[3] = {0, 0, 0};
    OxEq q;
    ox_equilibrium(c, v, q);
    double alpha_low = q.alpha0;
    for (int i = 2; i <= npoints; ++i) {
        const double *p = v + (size_t)(i - 1) * nv;
        ox_equilibrium(c, p, q);
        const double alpha_high = q.alpha0;
        if (alpha_high < alpha_low) {
            found_max = true;
            alpha_max = alpha_low;
            step_number = i - 1;
            const double *pm = v + (size_t)(i - 2) * nv;
            for (int k = 0; k < 3; ++k) { x_max[k] = pm[k]; k_max[k] = pm[3 + k]; }
            break;
        }
        alpha_low = alpha_high;
    }
    o.found_max = found_max ? 1 : 0;
    if (!found_max) return;           // x_max, k_max, alpha_max = 0, step_number = 0 (:122-126)
    for (int k = 0; k < 3; ++k) { o.x_max[k] = x_max[k]; o.k_max[k] = k_max[k]; }
    o.alpha_max = alpha_max;
    o.step_number = step_number;
    // ---- find_x_cutoff_ray
    const double alpha_tolerence = one / (10.0 * 10.0 * 10.0 * 10.0);
    bool found_cutoff = false;
    double x_temp[3] = {x_max[0], x_max[1], x_max[2]};
    ox_equilibrium(c, x_temp, q);
    double alpha_temp = q.alpha0;
    int iteration;
    for (iteration = 1; iteration <= 10; ++iteration) {
        if (std::fabs(alpha_temp - one) <= alpha_tolerence) { found_cutoff = true; break; }
        ox_equilibrium(c, x_temp, q);
        alpha_temp = q.alpha0;
        const double ng = ox_norm2(q.gradns0);
        const double mod_grad_alpha = ng * (alpha_temp / q.ns0);
        const double xy[3] = {x_temp[0], x_temp[1], 0.0};
        const double r = std::sqrt(xy[0] * xy[0] + xy[1] * xy[1]);
        const double delta = (one - q.alpha0) / mod_grad_alpha;
        const double step = std::fmin(delta, 0.25 * r);
        for (int k = 0; k < 3; ++k) x_temp[k] = x_temp[k] + q.gradns0[k] / ng * step;
    }
    o.iteration = iteration;
    o.found_cutoff = found_cutoff ? 1 : 0;
    if (!found_cutoff) return;        // x_cut = 0 (:140-144)
    for (int k = 0; k < 3; ++k) o.x_cut[k] = x_temp[k];
    // ---- OX_conv_coeff
    ox_equilibrium(c, x_temp, q);
    const double ng = ox_norm2(q.gradns0);
    double xc[3], yc[3], zc[3], vt[3];
    for (int k = 0; k < 3; ++k) xc[k] = q.gradns0[k] / ng;
    vt[0] = q.bunit[1] * xc[2] - q.bunit[2] * xc[1];   // cross_product(bunit, xc_unit)
    vt[1] = q.bunit[2] * xc[0] - q.bunit[0] * xc[2];
    vt[2] = q.bunit[0] * xc[1] - q.bunit[1] * xc[0];
    const double nvt = ox_norm2(vt);
    for (int k = 0; k < 3; ++k) yc[k] = vt[k] / nvt;
    zc[0] = xc[1] * yc[2] - xc[2] * yc[1];
    zc[1] = xc[2] * yc[0] - xc[0] * yc[2];
    zc[2] = xc[0] * yc[1] - xc[1] * yc[0];
    auto dot = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const double theta = std::acos(dot(xc, q.bunit));
    const double gamma = std::fabs(q.gamma0);
    const double L = q.ns0 / ng;
    const double n_vertical = dot(k_max, xc) / c.k0;
    const double nz_c = dot(k_max, zc) / c.k0;
    const double ny_c = dot(k_max, yc) / c.k0;
    const double n_crit = std::sin(theta) * std::sqrt(gamma / (one + gamma));
    const double ct = std::cos(theta), st = std::sin(theta);
    const double F = 0.5 * (one + gamma) * std::sqrt(gamma) / std::pow((one + gamma) * (ct * ct) + (st * st) / two, 1.5);
    const double G = 0.5 * std::sqrt(gamma) / std::sqrt((one + gamma) * (ct * ct) + (st * st) / two);
    const double dz = std::fabs(nz_c) - n_crit, ay = std::fabs(ny_c);
    const double conv_coeff = std::exp(-pi * c.k0 * L * (F * (dz * dz) + G * (ay * ay)));
    if (conv_coeff > conversion_threshold) {
        o.converted = 1;
        o.conv_coeff = conv_coeff;
        for (int k = 0; k < 3; ++k) { o.nvecx_c[k] = n_vertical * xc[k]; o.nvecy_c[k] = ny_c * yc[k]; o.nvecz_c[k] = nz_c * zc[k]; }
    }
}

}  // namespace rays_oracle
