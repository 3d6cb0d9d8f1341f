// rays_oracle_trace.hpp — trace_rays, launch fans, deposition binning.  TEST INFRASTRUCTURE.
//   trace_rays        L/ray_tracing.f90:1-290
//   launchers         L/simple_slab_ray_init_m.f90:58-185, L/solovev_ray_init_nphi_ntheta_m.f90:61-209,
//                     L/axisym_toroid_ray_init_R_Z_nphi_ntheta_m.f90:67-244,
//                     L/one_ray_init_XYZ_k_direction_m.f90:131-180, L/file_input_ray_init_m.f90:74-207
//   dispersion roots  L/disp_solve_cold_n1sq_vs_n3.f90:1-90, L/disp_solve_cold_nsq_vs_theta.f90:1-73,
//                     L/dispersion_solvers_m.f90:49-231
//   binner            M/bin_to_uniform_grid_m.f90:155-266, P/deposition_profiles_m.f90:228-292,438-499
#pragma once
#include <cstdio>

namespace rays_oracle {

struct RayOut {  // one ray's slice of ray_results_m (caller-owned, reference layout)
    double *ray_vec;    // [npoints_alloc][nv]
    double *residual;   // [npoints_alloc]
    int npoints;
    int stop_code;
    double end_residual, max_residual, end_ray_parameter;
    long nrhs;          // RHS evaluations (SG accounting)
    bool started;       // false for "did not start" rays (summary block skipped, A.5 (R))
    double *end_vec;    // [nv] end_ray_vec = v at loop exit (may be NULL)
    int run_error;
};

// one iteration of ray_loop (L/ray_tracing.f90:67-264)
template <class R>
inline void trace_one_ray(const rays_cfg &c, const double *rvec0, const double *rindex_vec0, RayOut &o) {
    const int nv = c.nv;
    int nstep = 0;
    R s(0.0), sout(0.0), resid(0.0);
    OdeStop<R> ray_stop;
    ray_stop.stop_ode = false; ray_stop.ode_stop_flag = 0; ray_stop.run_error = 0;
    ray_stop.rel_err = R(c.rel_err0); ray_stop.abs_err = R(c.abs_err0);  // ray_init_SG_ode
    o.nrhs = 0; o.started = true;
    R v[RAYS_NV_MAX];
    initialize_ode_vector(c, rvec0, rindex_vec0, v);
    if (o.ray_vec) for (int i = 0; i < nv; ++i) o.ray_vec[i] = val(v[i]);
    if (o.residual) o.residual[0] = 0.0;
    double resid_prev = 0.0, resid_last = 0.0, resid_max = 0.0;  // residual(nstep), residual(nstep+1), maxval over 1:nstep
    check_save(c, sout, v, resid, ray_stop);
    if (ray_stop.stop_ode) {  // "did not start": only npoints, flag, first point (A.5 (R))
        o.npoints = 1; o.stop_code = ray_stop.ode_stop_flag; o.run_error = ray_stop.run_error; o.started = false;
        o.end_residual = 0.0; o.max_residual = 0.0; o.end_ray_parameter = 0.0;
        return;
    }
    for (;;) {
        s = sout;
        sout = sout + c.ds;
        if (sout > c.s_max) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_SOUT_GT_SMAX; break; }
        if (nstep + 1 > c.nstep_max) { ray_stop.stop_ode = true; ray_stop.ode_stop_flag = RAYS_STOP_NSTEP_MAX; nstep = c.nstep_max; break; }
        if (c.ode_solver == RAYS_ODE_RK4) { RK4_ode(c, v, s, sout, ray_stop); o.nrhs += 4; }
        else SG_ode(c, v, s, sout, ray_stop, o.nrhs);
        if (ray_stop.stop_ode) break;
        check_save(c, s, v, resid, ray_stop);
        if (ray_stop.stop_ode) break;
        if (oracle_old_generation() && ray_stop.ode_stop_flag != 0) break;   // RAYS_code/ray_tracing.f90:131-140 (see oracle_old_generation)
        nstep = nstep + 1;
        if (o.ray_vec) for (int i = 0; i < nv; ++i) o.ray_vec[(size_t)nstep * nv + i] = val(v[i]);
        if (o.residual) o.residual[nstep] = val(resid);
        // bookkeeping for end_residuals = residual(nstep) and max over residual(1:nstep) (1-based):
        // after this store residual(nstep+1) = resid; residual(nstep) = previous
        resid_prev = resid_last;
        resid_last = val(resid);
        if (std::fabs(resid_prev) > resid_max) resid_max = std::fabs(resid_prev);
    }
    o.npoints = nstep + 1;
    // end_residuals = residual(nstep,iray) [(R) one before the last; (X) nstep=0 -> 0]
    o.end_residual = nstep >= 1 ? resid_prev : 0.0;
    o.max_residual = nstep >= 1 ? resid_max : -DBL_MAX;  // maxval of a zero-size array = -huge
    o.end_ray_parameter = val(v[6]);
    if (o.end_vec) for (int i = 0; i < nv; ++i) o.end_vec[i] = val(v[i]);
    o.stop_code = ray_stop.ode_stop_flag;
    o.run_error = ray_stop.run_error;
    // end_ray_vec = v : returned through ray_vec's caller (v is the last saved point unless the
    // stepper advanced it; RK4/SG leave v untouched on a failed step)
}

// ---------------- dispersion roots -----------------------------------------------------------
// solve_cold_n1sq_vs_n3 (L/disp_solve_cold_n1sq_vs_n3.f90:1-90)
template <class R> inline void solve_cold_n1sq_vs_n3(const rays_cfg &c, const EqPoint<R> &eq, R n3, Cx<R> n1sq[5]) {
    R S, D, P, Rr, L;
    RLSDP_cold(c, eq, S, D, P, Rr, L);
    R n3s = n3 * n3;
    R a = S;
    R b = -Rr * L - P * S + n3s * (P + S);
    R cc = P * (n3s - Rr) * (n3s - L);
    R discr = b * b - 4.0 * a * cc;
    R sgn_b = Copysign(R(1.0), b);
    Cx<R> sqrt_d = csqrt_(Cx<R>(discr, R(0.0)));
    Cx<R> mb(-b, R(0.0));
    if (sgn_b < 0.0) {
        Cx<R> num = mb + sqrt_d;
        n1sq[1] = num / Cx<R>(2.0 * a, R(0.0));
        n1sq[2] = Cx<R>(2.0 * cc, R(0.0)) / num;
    } else {
        Cx<R> num = mb - sqrt_d;
        n1sq[2] = num / Cx<R>(2.0 * a, R(0.0));
        n1sq[1] = Cx<R>(2.0 * cc, R(0.0)) / num;
    }
    if (cabs(n1sq[1]) <= cabs(n1sq[2])) { n1sq[3] = n1sq[1]; n1sq[4] = n1sq[2]; }
    else { n1sq[3] = n1sq[2]; n1sq[4] = n1sq[1]; }
}
// solve_n1_vs_n2_n3 (L/dispersion_solvers_m.f90:49-112)
template <class R> inline Cx<R> solve_n1_vs_n2_n3(const rays_cfg &c, const EqPoint<R> &eq, R n2, R n3) {
    Cx<R> nperp_sq[5];
    solve_cold_n1sq_vs_n3(c, eq, n3, nperp_sq);
    Cx<R> z = nperp_sq[c.wave_mode];
    z.re = z.re - n2 * n2;
    Cx<R> r = csqrt_(z);
    R ks((double)c.k0_sign);
    return Cx<R>(ks * r.re, ks * r.im);
}
// solve_cold_nsq_vs_theta (L/disp_solve_cold_nsq_vs_theta.f90:1-73); returns false if discr < 0
template <class R> inline bool solve_cold_nsq_vs_theta(const rays_cfg &c, const EqPoint<R> &eq, R theta, R nsq[5]) {
    R S, D, P, Rr, L;
    RLSDP_cold(c, eq, S, D, P, Rr, L);
    R ct = Cos(theta);
    R cos2 = ct * ct;
    R sin2 = 1.0 - cos2;
    R a = S * sin2 + P * cos2;
    R b = -Rr * L * sin2 - P * S * (1.0 + cos2);
    R cc = P * Rr * L;
    R sgn_b = Copysign(R(1.0), b);
    R discr = b * b - 4.0 * a * cc;
    if (discr < 0.0) return false;
    R sq = Sqrt(discr);
    if (sgn_b < 0.0) { nsq[1] = (-b + sq) / (2.0 * a); nsq[2] = 2.0 * cc / (-b + sq); }
    else { nsq[2] = (-b - sq) / (2.0 * a); nsq[1] = 2.0 * cc / (-b - sq); }
    if (Fabs(nsq[1]) <= Fabs(nsq[2])) { nsq[3] = nsq[1]; nsq[4] = nsq[2]; }
    else { nsq[3] = nsq[2]; nsq[4] = nsq[1]; }
    return true;
}

struct FanOut {
    std::vector<double> rvec0, rindex_vec0, ray_pwr_wt;
    long nray = 0;
    void push(const double r[3], const double n[3]) {
        for (int i = 0; i < 3; ++i) { rvec0.push_back(r[i]); rindex_vec0.push_back(n[i]); }
        ++nray;
    }
};

// simple_slab_ray_init (L/simple_slab_ray_init_m.f90:58-185) incl. (R) z uses dy_launch,
// weights divided by nray twice
inline void simple_slab_ray_init(const rays_cfg &c, const rays_slab_launch &p, FanOut &f) {
    typedef double R;
    for (int iz = 1; iz <= p.n_z_launch; ++iz) {
        double z = p.z_launch0 + (iz - 1) * p.dy_launch;
        for (int iy = 1; iy <= p.n_y_launch; ++iy) {
            double y = p.y_launch0 + (iy - 1) * p.dy_launch;
            for (int ix = 1; ix <= p.n_x_launch; ++ix) {
                double x = p.x_launch0 + (ix - 1) * p.dx_launch;
                double rvec[3] = {x, y, z};
                for (int iky = 1; iky <= p.n_ky_launch; ++iky) {
                    double ny = p.rindex_y0 + (iky - 1) * p.delta_rindex_y0;
                    for (int ikz = 1; ikz <= p.n_kz_launch; ++ikz) {
                        double nz = p.rindex_z0 + (ikz - 1) * p.delta_rindex_z0;
                        EqPoint<R> eq;
                        equilibrium<R>(c, rvec, c.omgrf, eq);
                        if (eq.equib_err != 0) continue;
                        // solve_nx_vs_ny_nz_by_bz (L/dispersion_solvers_m.f90:116-153)
                        double n2 = ny * eq.bunit[2] - nz * eq.bunit[1];
                        double n3 = ny * eq.bunit[1] + nz * eq.bunit[2];
                        Cx<R> nx = solve_n1_vs_n2_n3<R>(c, eq, n2, n3);
                        if (nx.im != 0.0) continue;  // evanescent
                        double nvec[3] = {nx.re, ny, nz};
                        f.push(rvec, nvec);
                    }
                }
            }
        }
    }
    f.ray_pwr_wt.assign(f.nray, f.nray ? 1.0 / (double)f.nray / (double)f.nray : 0.0);
}

// shared body of the (nphi, ntheta) launchers: one candidate at rvec (y = 0)
template <class PsiFn>
inline bool nphi_ntheta_candidate(const rays_cfg &c, const double rvec[3], double rindex_theta, double rindex_phi,
                                  PsiFn psifn, bool strict_zero_im, double nout[3]) {
    typedef double R;
    EqPoint<R> eq;
    equilibrium<R>(c, rvec, c.omgrf, eq);
    if (eq.equib_err != 0) return false;
    double psi, gradpsi[3], psiN, gradpsiN[3];
    psifn(rvec, psi, gradpsi, psiN, gradpsiN);
    double gn = std::sqrt(gradpsi[0] * gradpsi[0] + gradpsi[1] * gradpsi[1] + gradpsi[2] * gradpsi[2]);
    double psi_unit[3] = {gradpsi[0] / gn, gradpsi[1] / gn, gradpsi[2] / gn};
    double phi_unit[3] = {0.0, 1.0, 0.0};
    double theta_unit[3] = {-gradpsi[2], 0.0, gradpsi[0]};
    double tn = std::sqrt(theta_unit[0] * theta_unit[0] + theta_unit[1] * theta_unit[1] + theta_unit[2] * theta_unit[2]);
    for (int i = 0; i < 3; ++i) theta_unit[i] = theta_unit[i] / tn;
    double trans_unit[3] = {eq.bunit[1] * psi_unit[2] - eq.bunit[2] * psi_unit[1],
                            eq.bunit[2] * psi_unit[0] - eq.bunit[0] * psi_unit[2],
                            eq.bunit[0] * psi_unit[1] - eq.bunit[1] * psi_unit[0]};
    double rindex_vec[3];
    for (int i = 0; i < 3; ++i) rindex_vec[i] = rindex_phi * phi_unit[i] + rindex_theta * theta_unit[i];
    double n3 = eq.bunit[0] * rindex_vec[0] + eq.bunit[1] * rindex_vec[1] + eq.bunit[2] * rindex_vec[2];
    double n2 = trans_unit[0] * rindex_vec[0] + trans_unit[1] * rindex_vec[1] + trans_unit[2] * rindex_vec[2];
    Cx<R> npsi = solve_n1_vs_n2_n3<R>(c, eq, n2, n3);
    if (strict_zero_im) { if (npsi.im != 0.0) return false; }
    else if (std::fabs(npsi.im) > 10.0 * DBL_MIN) return false;
    for (int i = 0; i < 3; ++i) nout[i] = rindex_vec[i] - npsi.re * psi_unit[i];
    return true;
}
// ray_init_solovev_nphi_ntheta (L/solovev_ray_init_nphi_ntheta_m.f90:61-209)
// (X) ray_pwr_wt: the Fortran initialises one entry per outer loop; the defined intent 1/nray is used.
inline void solovev_ray_init(const rays_cfg &c, const rays_solovev_launch &p, FanOut &f) {
    const rays_solovev_eq &q = c.solovev;
    auto psifn = [&](const double r[3], double &psi, double g[3], double &psiN, double gN[3]) {
        solovev_psi<double>(r, q.bphi0, q.iota0, q.rmaj, q.kappa, q.psiB, psi, g, psiN, gN);
    };
    for (int iray = 1; iray <= p.n_r_launch; ++iray)
        for (int itheta = 1; itheta <= p.n_theta_launch; ++itheta) {
            double theta = p.theta_launch0 + (itheta - 1) * p.dtheta_launch;
            double rmin_launch = p.r_launch0 + (iray - 1) * p.dr_launch;
            double rvec[3] = {q.rmaj + rmin_launch * std::cos(theta), 0.0, rmin_launch * std::sin(theta)};
            for (int it = 1; it <= p.n_rindex_theta; ++it) {
                double rindex_theta = p.rindex_theta0 + (it - 1) * p.delta_rindex_theta;
                for (int ip = 1; ip <= p.n_rindex_phi; ++ip) {
                    double rindex_phi = p.rindex_phi0 + (ip - 1) * p.delta_rindex_phi;
                    double n[3];
                    if (nphi_ntheta_candidate(c, rvec, rindex_theta, rindex_phi, psifn, true, n)) f.push(rvec, n);
                }
            }
        }
    f.ray_pwr_wt.assign(f.nray, f.nray ? 1.0 / (double)f.nray : 0.0);
}
// ray_init_axisym_toroid_R_Z_nphi_ntheta (L/axisym_toroid_ray_init_R_Z_nphi_ntheta_m.f90:67-244)
inline void axisym_ray_init(const rays_cfg &c, const rays_axisym_launch &p, FanOut &f) {
    auto psifn = [&](const double r[3], double &psi, double g[3], double &psiN, double gN[3]) {
        axisym_toroid_psi<double>(c, r, psi, g, psiN, gN);
    };
    for (int iR = 1; iR <= p.n_R_launch; ++iR)
        for (int iZ = 1; iZ <= p.n_Z_launch; ++iZ) {
            double rvec[3] = {p.R_launch0, 0.0, p.Z_launch0};  // (R) same point for every i_R, i_Z
            for (int it = 1; it <= p.n_rindex_theta; ++it) {
                double rindex_theta = p.rindex_theta0 + (it - 1) * p.delta_rindex_theta;
                for (int ip = 1; ip <= p.n_rindex_phi; ++ip) {
                    double rindex_phi = p.rindex_phi0 + (ip - 1) * p.delta_rindex_phi;
                    double n[3];
                    if (nphi_ntheta_candidate(c, rvec, rindex_theta, rindex_phi, psifn, false, n)) f.push(rvec, n);
                }
            }
        }
    f.ray_pwr_wt.assign(f.nray, f.nray ? 1.0 / (double)f.nray : 0.0);
}
// ray_init_XYZ_k_direction (L/one_ray_init_XYZ_k_direction_m.f90:131-180) over n_in rays, as
// file_input_ray_init does (L/file_input_ray_init_m.f90:160-202); (R) weights: 1/nray if all input
// weights are zero, else ray_pwr_wt_temp (never filled: zeros) * n_rays_in/nray.
inline void directions_ray_init(const rays_cfg &c, long n_in, const double *rvec_in, const double *nvec_in,
                                bool all_weights_zero, FanOut &f) {
    typedef double R;
    for (long i = 0; i < n_in; ++i) {
        const double *rvec = rvec_in + 3 * i;
        double nvec[3] = {nvec_in[3 * i], nvec_in[3 * i + 1], nvec_in[3 * i + 2]};
        EqPoint<R> eq;
        equilibrium<R>(c, rvec, c.omgrf, eq);
        if (eq.equib_err != 0) continue;
        double nn = std::sqrt(nvec[0] * nvec[0] + nvec[1] * nvec[1] + nvec[2] * nvec[2]);
        for (int k = 0; k < 3; ++k) nvec[k] = nvec[k] / nn;
        double cos_theta = eq.bunit[0] * nvec[0] + eq.bunit[1] * nvec[1] + eq.bunit[2] * nvec[2];
        double theta = std::acos(cos_theta);
        double nsq[5];
        if (!solve_cold_nsq_vs_theta<R>(c, eq, theta, nsq)) continue;  // (X) nsq undefined in the Fortran
        // solve_n_vs_theta: n_out = k_sign*sqrt(nsq(i_mode)) on a REAL -> NaN if negative (A.5 (R))
        double n = (double)c.k0_sign * std::sqrt(nsq[c.wave_mode]);
        for (int k = 0; k < 3; ++k) nvec[k] = n * nvec[k];
        f.push(rvec, nvec);
    }
    f.ray_pwr_wt.assign(f.nray, (f.nray && all_weights_zero) ? 1.0 / (double)f.nray : 0.0);
}

// ---------------- deposition -----------------------------------------------------------------
// binner_real (M/bin_to_uniform_grid_m.f90:155-266); binned_Q is zeroed on entry like the Fortran
inline int binner_real(const double *Q, const double *xQ, int nx, double xmin, double xmax, double *binned_Q, int n_bins) {
    int ierr = 0;
    for (int i = 0; i < n_bins; ++i) binned_Q[i] = 0.0;
    double x_range = xmax - xmin;
    double x_bin_width = x_range / n_bins;
    for (int is = 2; is <= nx; ++is) {
        double x_low = std::fmin(xQ[is - 2], xQ[is - 1]);
        double x_high = std::fmax(xQ[is - 2], xQ[is - 1]);
        double ix_low = (x_low - xmin) / x_bin_width;
        double ix_high = (x_high - xmin) / x_bin_width;
        double delta_ix = ix_high - ix_low;
        int index_low = (int)std::floor(ix_low) + 1;
        int index_high = (int)std::floor(ix_high) + 1;
        if (x_high >= xmax) index_high = n_bins;
        int delta_i = index_high - index_low;
        double delta_Q = Q[is - 1] - Q[is - 2];
        double Q_density = delta_Q / delta_ix;
        if (std::fabs(delta_Q) < 4.0 * DBL_MIN) continue;
        if (x_high < xmin || x_low > xmax) continue;
        if (x_low < xmin) {
            double fraction_in = ix_high / delta_ix;
            delta_Q = delta_Q * fraction_in;
            ix_low = 0.0; index_low = 1;
            delta_i = index_high - index_low;
            ierr = 1;
        }
        if (x_high > xmax) {
            double fraction_in = ((double)n_bins - ix_low) / delta_ix;
            delta_Q = delta_Q * fraction_in;
            ix_high = (double)n_bins; index_high = n_bins;
            delta_i = index_high - index_low;
            ierr = 2;
        }
        if (delta_i == 0) binned_Q[index_low - 1] += delta_Q;
        else if (delta_i > 0) {
            double fraction_low = ((double)index_low - ix_low) / delta_ix;
            binned_Q[index_low - 1] += delta_Q * fraction_low;
            double fraction_high = (ix_high - (double)(index_high - 1)) / delta_ix;
            binned_Q[index_high - 1] += delta_Q * fraction_high;
            if (delta_i > 1) for (int i = index_low + 1; i <= index_high - 1; ++i) binned_Q[i - 1] += Q_density;
        }
    }
    return ierr;
}
// Ptotal_x (slab) / Ptotal_psi (axisym) evaluator + bin_a_ray for one ray
inline void bin_a_ray(const rays_cfg &c, const double *ray_vec, int npoints, double ray_power, double grid_min,
                      double grid_max, double *work, int n_bins) {
    std::vector<double> Q(npoints), xQ(npoints);
    const int nv = c.nv;
    for (int ip = 0; ip < npoints; ++ip) {
        const double *v = ray_vec + (size_t)ip * nv;
        if (c.equilib_model == RAYS_EQ_SLAB) xQ[ip] = v[0];
        else {
            double psi, g[3], psiN, gN[3];
            axisym_toroid_psi<double>(c, v, psi, g, psiN, gN);
            xQ[ip] = psiN;
        }
        Q[ip] = v[7] * ray_power;
    }
    binner_real(Q.data(), xQ.data(), npoints, grid_min, grid_max, work, n_bins);
}

}  // namespace rays_oracle
