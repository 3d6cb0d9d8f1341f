// oracle_capi.cpp — extern "C" face of the CPU oracle for ctypes (tests/, smoke(), bench.py
// cpu_baseline / --impl reference ONLY).  TEST INFRASTRUCTURE: see rays_oracle.hpp header.
// Build: make -C oracle   (g++ -O2 -fopenmp -ffp-contract=off, no fast-math)
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstring>

#include "rays_oracle.hpp"
#include "rays_oracle_coils.hpp"
#include "rays_oracle_ox.hpp"

using namespace rays_oracle;

static const char *kStopStrings[RAYS_STOP_CODE_MAX] = {
    "", "sout > s_max", " nstep > nstep_max", "infinite Vg", "ray stalled", "dispersion_residual",
    "infinite_Vg", "total_absorption", "ODE total error", "step number .ge. maxnum", "equations stiff",
    "t == tout", "relerr or abserr < 0", "eps <= 0", "", "", "", "", "", "",
    "x out_of_bounds", "y out_of_bounds", "z out_of_bounds", "R out_of_box", "z out_of_box",
    "R_out_of_box", "Z_out_of_box", "out_of_plasma", "R out_of_bounds", "z out_of_bounds",
    "negative_dens", "negative_temp"};

extern "C" {

int oracle_stop_string(int code, char *buf, int len) {
    const char *s = (code >= 0 && code < RAYS_STOP_CODE_MAX) ? kStopStrings[code] : "";
    int n = (int)std::strlen(s);
    for (int i = 0; i < len; ++i) buf[i] = i < n ? s[i] : ' ';
    return 0;
}

// trace_rays (L/ray_tracing.f90:1-290) over the whole fan, OpenMP over rays like the reference
// (:62-64) but schedule(dynamic) and all state thread-private.  nthreads <= 0 -> all cores.
// Returns the run-level status (RAYS_OK / RAYS_ERR_IM_DET).
int oracle_trace(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, int nthreads, long *nrhs_total) {
    const rays_cfg &c = *cfg;
    const int nv = c.nv;
    const long nray = fan->nray;
    const int npa = res->npoints_alloc;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    int run_error = 0;
    long nrhs = 0, steps = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : nrhs, steps) reduction(max : run_error)
    for (long iray = 0; iray < nray; ++iray) {
        RayOut o;
        double endv[RAYS_NV_MAX];
        std::vector<double> scratch;
        o.ray_vec = res->ray_vec ? res->ray_vec + (size_t)iray * npa * nv : nullptr;
        o.residual = res->residual ? res->residual + (size_t)iray * npa : nullptr;
        o.end_vec = endv;
        trace_one_ray<double>(c, fan->rvec0 + 3 * iray, fan->rindex_vec0 + 3 * iray, o);
        res->npoints[iray] = o.npoints;
        if (res->ray_stop_code) res->ray_stop_code[iray] = o.stop_code;
        if (res->ray_stop_flag) oracle_stop_string(o.stop_code, res->ray_stop_flag + (size_t)iray * RAYS_FLAG_LEN, RAYS_FLAG_LEN);
        if (o.started) {
            if (res->initial_ray_power) res->initial_ray_power[iray] = fan->ray_pwr_wt ? fan->ray_pwr_wt[iray] : 0.0;
            if (res->end_residuals) res->end_residuals[iray] = o.end_residual;
            if (res->max_residuals) res->max_residuals[iray] = o.max_residual;
            if (res->end_ray_parameter) res->end_ray_parameter[iray] = o.end_ray_parameter;
            if (res->start_ray_vec) {
                double v0[RAYS_NV_MAX];
                initialize_ode_vector<double>(c, fan->rvec0 + 3 * iray, fan->rindex_vec0 + 3 * iray, v0);
                for (int i = 0; i < nv; ++i) res->start_ray_vec[(size_t)iray * nv + i] = v0[i];
            }
            if (res->end_ray_vec) for (int i = 0; i < nv; ++i) res->end_ray_vec[(size_t)iray * nv + i] = endv[i];
        } else {  // summary block skipped: arrays keep their initial zeros (A.5 (R))
            if (res->initial_ray_power) res->initial_ray_power[iray] = 0.0;
            if (res->end_residuals) res->end_residuals[iray] = 0.0;
            if (res->max_residuals) res->max_residuals[iray] = 0.0;
            if (res->end_ray_parameter) res->end_ray_parameter[iray] = 0.0;
            if (res->start_ray_vec) for (int i = 0; i < nv; ++i) res->start_ray_vec[(size_t)iray * nv + i] = 0.0;
            if (res->end_ray_vec) for (int i = 0; i < nv; ++i) res->end_ray_vec[(size_t)iray * nv + i] = 0.0;
        }
        nrhs += o.nrhs;
        steps += o.npoints - 1;
        run_error = std::max(run_error, o.run_error);
    }
    auto t1 = std::chrono::steady_clock::now();
    res->total_trace_time = std::chrono::duration<double>(t1 - t0).count();
    res->total_ray_steps = steps;
    if (res->ray_trace_time) for (long i = 0; i < nray; ++i) res->ray_trace_time[i] = nray ? res->total_trace_time / nray : 0.0;
    if (nrhs_total) *nrhs_total = nrhs;
    return run_error;
}

// exact algorithmic flop count: trace rays [first, first+count) with the counting
// scalar; returns total flops, ray-steps and RHS evaluations.
int oracle_count_flops(const rays_cfg *cfg, const rays_fan *fan, long first, long count, double *flops, long *steps, long *nrhs) {
    uint64_t f = 0;
    long st = 0, nr = 0;
    for (long iray = first; iray < first + count && iray < fan->nray; ++iray) {
        RayOut o;
        o.ray_vec = nullptr; o.residual = nullptr; o.end_vec = nullptr;
        FlopCounter::n() = 0;
        trace_one_ray<CountReal>(*cfg, fan->rvec0 + 3 * iray, fan->rindex_vec0 + 3 * iray, o);
        f += FlopCounter::n();
        st += o.npoints - 1;
        nr += o.nrhs;
    }
    *flops = (double)f; *steps = st; *nrhs = nr;
    return 0;
}

static void pack_eq(const EqPoint<double> &eq, double *o) {
    int k = 0;
    for (int i = 0; i < 3; ++i) o[k++] = eq.bvec[i];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) o[k++] = eq.gradbtensor[i][j];   // Fortran order [i+3j]
    for (int s = 0; s < NS0; ++s) o[k++] = eq.ns[s];
    for (int s = 0; s < NS0; ++s) for (int i = 0; i < 3; ++i) o[k++] = eq.gradns[i][s];
    for (int s = 0; s < NS0; ++s) o[k++] = eq.ts[s];
    for (int s = 0; s < NS0; ++s) for (int i = 0; i < 3; ++i) o[k++] = eq.gradts[i][s];
    o[k++] = eq.bmag;
    for (int i = 0; i < 3; ++i) o[k++] = eq.gradbmag[i];
    for (int i = 0; i < 3; ++i) o[k++] = eq.bunit[i];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) o[k++] = eq.gradbunit[i][j];
    for (int s = 0; s < NS0; ++s) o[k++] = eq.omgc[s];
    for (int s = 0; s < NS0; ++s) o[k++] = eq.omgp2[s];
    for (int s = 0; s < NS0; ++s) o[k++] = eq.alpha[s];
    for (int s = 0; s < NS0; ++s) o[k++] = eq.gamma[s];
}

int oracle_probe_equilibrium(const rays_cfg *cfg, long n, const double *rvec, double *out, int32_t *err) {
    for (long i = 0; i < n; ++i) {
        EqPoint<double> eq;
        equilibrium<double>(*cfg, rvec + 3 * i, cfg->omgrf, eq);
        pack_eq(eq, out + (size_t)i * RAYS_EQ_OUT);
        err[i] = eq.equib_err;
    }
    return 0;
}
int oracle_probe_rhs(const rays_cfg *cfg, long n, const double *v, double *dvds, int32_t *stop) {
    const int nv = cfg->nv;
    for (long i = 0; i < n; ++i) {
        OdeStop<double> st; st.stop_ode = false; st.ode_stop_flag = 0; st.run_error = 0;
        double d[RAYS_NV_MAX];
        for (int k = 0; k < nv; ++k) d[k] = 0.0;
        eqn_ray<double>(*cfg, 0.0, v + (size_t)i * nv, d, st);
        for (int k = 0; k < nv; ++k) dvds[(size_t)i * nv + k] = st.stop_ode ? 0.0 : d[k];
        stop[i] = st.stop_ode ? st.ode_stop_flag : 0;
    }
    return 0;
}
int oracle_probe_check_save(const rays_cfg *cfg, long n, const double *v, double *resid, int32_t *stop) {
    const int nv = cfg->nv;
    for (long i = 0; i < n; ++i) {
        OdeStop<double> st; st.stop_ode = false; st.ode_stop_flag = 0; st.run_error = 0;
        double r = 0.0;
        check_save<double>(*cfg, 0.0, v + (size_t)i * nv, r, st);
        resid[i] = r;
        stop[i] = st.stop_ode ? st.ode_stop_flag : 0;
    }
    return 0;
}

// launch fans: fill caller arrays (capacity cap rays); return surviving count (or -needed if cap too small)
static long emit_fan(const FanOut &f, long cap, double *rvec0, double *rindex_vec0, double *ray_pwr_wt) {
    if (f.nray > cap) return -f.nray;
    std::copy(f.rvec0.begin(), f.rvec0.end(), rvec0);
    std::copy(f.rindex_vec0.begin(), f.rindex_vec0.end(), rindex_vec0);
    std::copy(f.ray_pwr_wt.begin(), f.ray_pwr_wt.end(), ray_pwr_wt);
    return f.nray;
}
long oracle_launch_fan_solovev(const rays_cfg *cfg, const rays_solovev_launch *p, long cap, double *r, double *n, double *w) {
    FanOut f; solovev_ray_init(*cfg, *p, f); return emit_fan(f, cap, r, n, w);
}
long oracle_launch_fan_axisym(const rays_cfg *cfg, const rays_axisym_launch *p, long cap, double *r, double *n, double *w) {
    FanOut f; axisym_ray_init(*cfg, *p, f); return emit_fan(f, cap, r, n, w);
}
long oracle_launch_fan_slab(const rays_cfg *cfg, const rays_slab_launch *p, long cap, double *r, double *n, double *w) {
    FanOut f; simple_slab_ray_init(*cfg, *p, f); return emit_fan(f, cap, r, n, w);
}
long oracle_launch_fan_directions(const rays_cfg *cfg, long n_in, const double *rvec_in, const double *nvec_in,
                                  int all_weights_zero, long cap, double *r, double *n, double *w) {
    FanOut f; directions_ray_init(*cfg, n_in, rvec_in, nvec_in, all_weights_zero != 0, f); return emit_fan(f, cap, r, n, w);
}

// calculate_deposition_profiles (P/deposition_profiles_m.f90:228-260): per-ray work column, then
// profile = sum(work, 2) in ray order, Q_sum = sum(profile)
int oracle_deposition(const rays_cfg *cfg, const rays_results *res, rays_deposition *dep) {
    const int nb = dep->n_bins;
    const long nray = res->nray;
    const int nv = cfg->nv, npa = res->npoints_alloc;
    std::vector<double> work((size_t)nb * nray);
#pragma omp parallel for schedule(dynamic, 64)
    for (long iray = 0; iray < nray; ++iray)
        bin_a_ray(*cfg, res->ray_vec + (size_t)iray * npa * nv, res->npoints[iray], res->initial_ray_power[iray],
                  dep->grid_min, dep->grid_max, work.data() + (size_t)iray * nb, nb);
    for (int b = 0; b < nb; ++b) {
        double s = 0.0;
        for (long iray = 0; iray < nray; ++iray) s += work[(size_t)iray * nb + b];
        dep->profile[b] = s;
    }
    double q = 0.0;
    for (int b = 0; b < nb; ++b) q += dep->profile[b];
    dep->Q_sum = q;
    return 0;
}
// calculate_B_on_rz_grid (MM/mirror_magnetics_m.f90:324-368) + one loop field for unit checks
int oracle_mirror_Brz_grid(const rays_coil *coils, int n_coils, int n_r, double r_min, double r_max, int n_z, double z_min, double z_max,
                           double *r_grid, double *z_grid, double *Br, double *Bz, double *Aphi) {
    mirror_Brz_grid(coils, n_coils, n_r, r_min, r_max, n_z, z_min, z_max, r_grid, z_grid, Br, Bz, Aphi);
    return 0;
}
int oracle_Brz_loop_scaled(long n, const double *r, const double *z, double *Br, double *Bz, double *Aphi) {
    for (long i = 0; i < n; ++i) Brz_loop_scaled(r[i], z[i], Br[i], Bz[i], Aphi[i]);
    return 0;
}
int oracle_elliptic(long n, const double *m, double *K, double *E) {
    for (long i = 0; i < n; ++i) { K[i] = elliptic_Km(m[i]); E[i] = elliptic_Em(m[i]); }
    return 0;
}
// analyze_OX_conv (P/OX_conv_analysis_m.f90:91-198) on host result arrays
int oracle_ox_conv(const rays_cfg *cfg, const rays_results *res, rays_ox_conv *out) {
    const int nv = cfg->nv, npa = res->npoints_alloc;
#pragma omp parallel for schedule(dynamic, 16)
    for (long iray = 0; iray < res->nray; ++iray)
        ox_conv_ray(*cfg, res->ray_vec + (size_t)iray * npa * nv, nv, res->npoints[iray], (int)iray + 1, out[iray]);
    return 0;
}
// one row of write_kx_profiles (P/slab_processor_m.f90:783-819): k0*nx of the cold root `mode` (1 plus, 2 minus, 3 fast,
// 4 slow; k0_sign = +1) at rvec = (x, 0, 0) through solve_nx_vs_ny_nz_by_bz (L/dispersion_solvers_m.f90:116-153), cast to
// single precision the way the Fortran writes it
int oracle_kx_profile(const rays_cfg *cfg, long n, const double *x, double ny, double nz, int mode, double *kx_re, double *kx_im) {
    rays_cfg c = *cfg;
    c.wave_mode = mode;
    c.k0_sign = 1;
    for (long i = 0; i < n; ++i) {
        const double rvec[3] = {x[i], 0.0, 0.0};
        EqPoint<double> eq;
        equilibrium<double>(c, rvec, c.omgrf, eq);
        const double n2 = ny * eq.bunit[2] - nz * eq.bunit[1];
        const double n3 = ny * eq.bunit[1] + nz * eq.bunit[2];
        const Cx<double> nx = solve_n1_vs_n2_n3<double>(c, eq, n2, n3);
        kx_re[i] = (double)((float)c.k0 * (float)nx.re);
        kx_im[i] = (double)((float)c.k0 * (float)nx.im);
    }
    return 0;
}
int oracle_binner(const double *Q, const double *xQ, int nx, double xmin, double xmax, double *binned, int n_bins) {
    return binner_real(Q, xQ, nx, xmin, xmax, binned, n_bins);
}

// zfun0_real_arg (M/zfunctions_m.f90:351-432) for the Z-function known-answer test
int oracle_zfun(const rays_cfg *cfg, long n, const double *x, const double *kz, double *re, double *im) {
    for (long i = 0; i < n; ++i) {
        Cx<double> z = zfun0_real_arg<double>(*cfg, x[i], kz[i]);
        re[i] = z.re; im[i] = z.im;
    }
    return 0;
}
int oracle_cspeval(const rays_spline1d *s, long n, const double *x, double *f, double *fp) {
    for (long i = 0; i < n; ++i) { double a = 0, b = 0; cspeval<double>(x[i], *s, a, &b); f[i] = a; fp[i] = b; }
    return 0;
}
int oracle_bcspeval(const rays_spline2d *s, long n, const double *x, const double *y, double *f, double *fx, double *fy) {
    for (long i = 0; i < n; ++i) { double a = 0, b = 0, c = 0; bcspeval_fp<double>(x[i], y[i], *s, a, b, c); f[i] = a; fx[i] = b; fy[i] = c; }
    return 0;
}
// dispersion roots (launch kernel parity): n1(n2,n3) complex and nsq(theta)
int oracle_solve_n1(const rays_cfg *cfg, const double *rvec, double n2, double n3, double *re, double *im) {
    EqPoint<double> eq; equilibrium<double>(*cfg, rvec, cfg->omgrf, eq);
    if (eq.equib_err) return eq.equib_err;
    Cx<double> z = solve_n1_vs_n2_n3<double>(*cfg, eq, n2, n3);
    *re = z.re; *im = z.im; return 0;
}
int oracle_solve_nsq_theta(const rays_cfg *cfg, const double *rvec, double theta, double *nsq4) {
    EqPoint<double> eq; equilibrium<double>(*cfg, rvec, cfg->omgrf, eq);
    if (eq.equib_err) return eq.equib_err;
    double nsq[5] = {0, 0, 0, 0, 0};
    if (!solve_cold_nsq_vs_theta<double>(*cfg, eq, theta, nsq)) return -1;
    for (int i = 0; i < 4; ++i) nsq4[i] = nsq[i + 1];
    return 0;
}
// test switch: end rays at the plasma edge like the older generation of the code did (see oracle_old_generation)
int oracle_set_old_generation(int on) { const int was = oracle_old_generation(); oracle_old_generation() = on; return was; }
int oracle_num_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
