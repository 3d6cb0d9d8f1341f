// rays_oracle_sg.hpp — Shampine–Gordon Adams PECE integrator, restated from
// L/ode_RAYS.f90 (ode :1-231, de :232-593, step :595-1234, intrp :1235-1362) and its driver
// L/SG_ode_m.f90:89-159.  TEST INFRASTRUCTURE (see rays_oracle.hpp header).
//
// Arrays keep the Fortran's 1-based indexing (element 0 unused) so that the index arithmetic
// can be compared line by line.  phi(l,i) -> phi[i][l] (l = 0..neqn-1).
#pragma once

namespace rays_oracle {

template <class R> struct SGWork {  // the `work`/`iwork` arrays of SG_ode, fresh per call (A.3)
    R yy[RAYS_NV_MAX], wt[RAYS_NV_MAX], p[RAYS_NV_MAX], yp[RAYS_NV_MAX], ypout[RAYS_NV_MAX];
    R phi[17][RAYS_NV_MAX];
    R alpha[13], beta[13], sig[14], v[13], w[13], g[14], psi[13];
    R x, h, hold, told, delsgn;
    bool start, phase1, nornd;
    int ns, k, kold, isnold;
};

inline const double *sg_gstr() {  // L/ode_RAYS.f90:776-779, single-precision literals
    static const double g[14] = {0.0,
                                 f32lit(0.50e+00), f32lit(0.0833e+00), f32lit(0.0417e+00), f32lit(0.0264e+00),
                                 f32lit(0.0188e+00), f32lit(0.0143e+00), f32lit(0.0114e+00), f32lit(0.00936e+00),
                                 f32lit(0.00789e+00), f32lit(0.00679e+00), f32lit(0.00592e+00), f32lit(0.00524e+00),
                                 f32lit(0.00468e+00)};
    return g;
}

// step (L/ode_RAYS.f90:595-1234)
template <class R>
inline void sg_step(const rays_cfg &c, int neqn, SGWork<R> &W, R &eps, bool &crash, OdeStop<R> &ray_stop, long &nrhs) {
    static const double two[14] = {0.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0, 1024.0, 2048.0, 4096.0, 8192.0};
    const double *gstr = sg_gstr();
    R &x = W.x, &h = W.h, &hold = W.hold;
    R *y = W.yy, *wt = W.wt, *p = W.p, *yp = W.yp;
    R(*phi)[RAYS_NV_MAX] = W.phi;
    R *alpha = W.alpha, *beta = W.beta, *sig = W.sig, *v = W.v, *w = W.w, *g = W.g, *psi = W.psi;
    int &k = W.k, &kold = W.kold, &ns = W.ns;
    bool &start = W.start, &phase1 = W.phase1, &nornd = W.nornd;

    const double twou = 2.0 * DBL_EPSILON;
    const double fouru = 2.0 * twou;
    crash = true;
    if (Fabs(h) < fouru * Fabs(x)) { h = Copysign(fouru * Fabs(x), h); return; }
    R p5eps = 0.5 * eps;
    R sum(0.0);
    for (int l = 0; l < neqn; ++l) { R q = y[l] / wt[l]; sum = sum + q * q; }
    R round = twou * Sqrt(sum);
    if (p5eps < round) { eps = 2.0 * round * (1.0 + fouru); return; }
    crash = false;
    g[1] = R(1.0); g[2] = R(0.5); sig[1] = R(1.0);
    R absh;
    if (start) {
        eqn_ray(c, x, y, yp, ray_stop); ++nrhs;
        if (ray_stop.stop_ode) return;
        R tot(0.0);
        for (int l = 0; l < neqn; ++l) {
            phi[1][l] = yp[l]; phi[2][l] = R(0.0);
            R q = yp[l] / wt[l]; tot = tot + q * q;
        }
        R total = Sqrt(tot);
        absh = Fabs(h);
        if (eps < 16.0 * total * h * h) absh = 0.25 * Sqrt(eps / total);
        h = Copysign(Fmax(absh, R(fouru * val(Fabs(x)))), h);
        hold = R(0.0);
        k = 1; kold = 0;
        start = false; phase1 = true; nornd = true;
        if (p5eps <= 100.0 * round) {
            nornd = false;
            for (int l = 0; l < neqn; ++l) phi[15][l] = R(0.0);
        }
    }
    int ifail = 0;
    int kp1, kp2, km1, km2, knew;
    R erkm2, erkm1, erk, err, xold;
    for (;;) {
        kp1 = k + 1; kp2 = k + 2; km1 = k - 1; km2 = k - 2;
        if (h != hold) ns = 0;
        if (ns <= kold) ns = ns + 1;
        int nsp1 = ns + 1;
        if (ns <= k) {
            beta[ns] = R(1.0);
            alpha[ns] = 1.0 / R((double)ns);
            R temp1 = h * R((double)ns);
            sig[nsp1] = R(1.0);
            for (int i = nsp1; i <= k; ++i) {
                R temp2 = psi[i - 1];
                psi[i - 1] = temp1;
                beta[i] = beta[i - 1] * psi[i - 1] / temp2;
                temp1 = temp2 + h;
                alpha[i] = h / temp1;
                sig[i + 1] = R((double)i) * alpha[i] * sig[i];
            }
            psi[k] = temp1;
            if (ns <= 1) {
                for (int iq = 1; iq <= k; ++iq) { v[iq] = 1.0 / R((double)(iq * (iq + 1))); w[iq] = v[iq]; }
            } else {
                if (kold < k) {
                    v[k] = 1.0 / R((double)(k * kp1));
                    for (int j = 1; j <= ns - 2; ++j) { int i = k - j; v[i] = v[i] - alpha[j + 1] * v[i + 1]; }
                }
                for (int iq = 1; iq <= kp1 - ns; ++iq) { v[iq] = v[iq] - alpha[ns] * v[iq + 1]; w[iq] = v[iq]; }
                g[nsp1] = w[1];
            }
            for (int i = ns + 2; i <= kp1; ++i) {
                for (int iq = 1; iq <= kp2 - i; ++iq) w[iq] = w[iq] - alpha[i - 1] * w[iq + 1];
                g[i] = w[1];
            }
        }
        for (int i = nsp1; i <= k; ++i)
            for (int l = 0; l < neqn; ++l) phi[i][l] = beta[i] * phi[i][l];
        for (int l = 0; l < neqn; ++l) { phi[kp2][l] = phi[kp1][l]; phi[kp1][l] = R(0.0); p[l] = R(0.0); }
        for (int j = 1; j <= k; ++j) {
            int i = kp1 - j;
            for (int l = 0; l < neqn; ++l) {
                p[l] = p[l] + phi[i][l] * g[i];
                phi[i][l] = phi[i][l] + phi[i + 1][l];
            }
        }
        if (!nornd) {
            for (int l = 0; l < neqn; ++l) {
                R tau = h * p[l] - phi[15][l];
                p[l] = y[l] + tau;
                phi[16][l] = (p[l] - y[l]) - tau;
            }
        } else {
            for (int l = 0; l < neqn; ++l) p[l] = y[l] + h * p[l];
        }
        xold = x;
        x = x + h;
        absh = Fabs(h);
        eqn_ray(c, x, p, yp, ray_stop); ++nrhs;
        if (ray_stop.stop_ode) return;
        erkm2 = R(0.0); erkm1 = R(0.0); erk = R(0.0);
        for (int l = 0; l < neqn; ++l) {
            if (0 < km2) { R q = (phi[km1][l] + yp[l] - phi[1][l]) / wt[l]; erkm2 = erkm2 + q * q; }
            if (0 <= km2) { R q = (phi[k][l] + yp[l] - phi[1][l]) / wt[l]; erkm1 = erkm1 + q * q; }
            R q = (yp[l] - phi[1][l]) / wt[l];
            erk = erk + q * q;
        }
        if (0 < km2) erkm2 = absh * sig[km1] * gstr[km2] * Sqrt(erkm2);
        if (0 <= km2) erkm1 = absh * sig[k] * gstr[km1] * Sqrt(erkm1);
        err = absh * Sqrt(erk) * (g[k] - g[kp1]);
        erk = absh * Sqrt(erk) * sig[kp1] * gstr[k];
        knew = k;
        if (0 < km2) {
            if (Fmax(erkm1, erkm2) <= erk) knew = km1;
        } else if (0 == km2) {
            if (erkm1 <= 0.5 * erk) knew = km1;
        }
        if (err <= eps) break;
        // step failed
        phase1 = false;
        x = xold;
        for (int i = 1; i <= k; ++i)
            for (int l = 0; l < neqn; ++l) phi[i][l] = (phi[i][l] - phi[i + 1][l]) / beta[i];
        for (int i = 2; i <= k; ++i) psi[i - 1] = psi[i] - h;
        ifail = ifail + 1;
        R temp2(0.5);
        if (3 < ifail) { if (p5eps < 0.25 * erk) temp2 = Sqrt(p5eps / erk); }
        if (3 <= ifail) knew = 1;
        h = temp2 * h;
        k = knew;
        if (Fabs(h) < fouru * Fabs(x)) {
            crash = true;
            h = Copysign(fouru * Fabs(x), h);
            eps = eps + eps;
            return;
        }
    }
    kold = k;
    hold = h;
    if (!nornd) {
        for (int l = 0; l < neqn; ++l) {
            R rho = h * g[kp1] * (yp[l] - phi[1][l]) - phi[16][l];
            y[l] = p[l] + rho;
            phi[15][l] = (y[l] - p[l]) - rho;
        }
    } else {
        for (int l = 0; l < neqn; ++l) y[l] = p[l] + h * g[kp1] * (yp[l] - phi[1][l]);
    }
    eqn_ray(c, x, y, yp, ray_stop); ++nrhs;
    if (ray_stop.stop_ode) return;
    for (int l = 0; l < neqn; ++l) {
        phi[kp1][l] = yp[l] - phi[1][l];
        phi[kp2][l] = phi[kp1][l] - phi[kp2][l];
    }
    for (int i = 1; i <= k; ++i)
        for (int l = 0; l < neqn; ++l) phi[i][l] = phi[i][l] + phi[kp1][l];
    R erkp1(0.0);
    if (knew == km1 || k == 12) phase1 = false;
    if (phase1) {
        k = kp1; erk = erkp1;
    } else if (knew == km1) {
        k = km1; erk = erkm1;
    } else if (kp1 <= ns) {
        for (int l = 0; l < neqn; ++l) { R q = phi[kp2][l] / wt[l]; erkp1 = erkp1 + q * q; }
        erkp1 = absh * gstr[kp1] * Sqrt(erkp1);
        if (k == 1) {
            if (erkp1 < 0.5 * erk) { k = kp1; erk = erkp1; }
        } else if (erkm1 <= Fmin(erk, erkp1)) {
            k = km1; erk = erkm1;
        } else if (erkp1 < erk && k < 12) {
            k = kp1; erk = erkp1;
        }
    }
    R hnew = h + h;
    if (!phase1) {
        if (p5eps < erk * two[k + 1]) {
            hnew = h;
            if (p5eps < erk) {
                R temp2((double)(k + 1));
                R r = Pow(p5eps / erk, 1.0 / temp2);
                hnew = absh * Fmax(R(0.5), Fmin(R(f32lit(0.9)), r));
                hnew = Copysign(Fmax(hnew, R(fouru * val(Fabs(x)))), h);
            }
        }
    }
    h = hnew;
}

// intrp (L/ode_RAYS.f90:1235-1362)
template <class R> inline void sg_intrp(int neqn, SGWork<R> &W, R xout, R *yout) {
    R g[14], rho[14], w[14];
    R hi = xout - W.x;
    int ki = W.kold + 1;
    for (int i = 1; i <= ki; ++i) w[i] = 1.0 / R((double)i);
    g[1] = R(1.0); rho[1] = R(1.0);
    R term(0.0);
    for (int j = 2; j <= ki; ++j) {
        R psijm1 = W.psi[j - 1];
        R gamma = (hi + term) / psijm1;
        R eta = hi / psijm1;
        for (int i = 1; i <= ki + 1 - j; ++i) w[i] = gamma * w[i] - eta * w[i + 1];
        g[j] = w[1];
        rho[j] = gamma * rho[j - 1];
        term = psijm1;
    }
    for (int l = 0; l < neqn; ++l) { W.ypout[l] = R(0.0); yout[l] = R(0.0); }
    for (int j = 1; j <= ki; ++j) {
        int i = ki + 1 - j;
        for (int l = 0; l < neqn; ++l) {
            yout[l] = yout[l] + g[i] * W.phi[i][l];
            W.ypout[l] = W.ypout[l] + rho[i] * W.phi[i][l];
        }
    }
    for (int l = 0; l < neqn; ++l) yout[l] = W.yy[l] + hi * yout[l];
}

// ode + de with iflag = 1 on entry and fresh work arrays (SG_ode_m.f90:105-120).
// Returns iflag.
template <class R>
inline int sg_de(const rays_cfg &c, int neqn, R *y, R &t, R tout, R &relerr, R &abserr, OdeStop<R> &ray_stop, long &nrhs) {
    SGWork<R> W{};
    const int maxnum = 500;
    const double fouru = 4.0 * DBL_EPSILON;
    if (neqn < 1) return 6;
    if (t == tout) { ray_stop.ode_stop_flag = RAYS_STOP_SG_T_EQ_TOUT; return 6; }
    if (relerr < 0.0 || abserr < 0.0) { ray_stop.ode_stop_flag = RAYS_STOP_SG_BAD_TOL; return 6; }
    R eps = Fmax(relerr, abserr);
    if (eps <= 0.0) { ray_stop.ode_stop_flag = RAYS_STOP_SG_EPS_LE_0; return 6; }
    const int isn = 1;  // sign(1, iflag) with iflag = 1
    int iflag = 1;
    R del = tout - t;
    R absdel = Fabs(del);
    R tend = t + 10.0 * del;  // isn > 0
    int nostep = 0, kle4 = 0;
    bool stiff = false;
    R releps = relerr / eps;
    R abseps = abserr / eps;
    // iflag == 1 -> restart
    W.start = true;
    W.x = t;
    for (int l = 0; l < neqn; ++l) W.yy[l] = y[l];
    W.delsgn = Copysign(R(1.0), del);
    W.h = Copysign(Fmax(Fabs(tout - W.x), R(fouru * val(Fabs(W.x)))), tout - W.x);
    W.ns = 0; W.k = 0; W.kold = 0; W.hold = R(0.0); W.phase1 = false; W.nornd = true;
    for (;;) {
        if (absdel <= Fabs(W.x - t)) {
            sg_intrp(neqn, W, tout, y);
            iflag = 2;
            t = tout;
            break;
        }
        // (isn > 0: the "cannot pass tout" extrapolation branch is never taken)
        if (maxnum <= nostep) {
            iflag = isn * 4;
            ray_stop.ode_stop_flag = RAYS_STOP_SG_MAXNUM;
            if (stiff) { iflag = isn * 5; ray_stop.ode_stop_flag = RAYS_STOP_SG_STIFF; }
            for (int l = 0; l < neqn; ++l) y[l] = W.yy[l];
            t = W.x;
            break;
        }
        W.h = Copysign(Fmin(Fabs(W.h), Fabs(tend - W.x)), W.h);
        for (int l = 0; l < neqn; ++l) W.wt[l] = releps * Fabs(W.yy[l]) + abseps;
        bool crash;
        sg_step(c, neqn, W, eps, crash, ray_stop, nrhs);
        if (ray_stop.stop_ode) return iflag;
        if (crash) {
            iflag = isn * 3;
            relerr = eps * releps;
            abserr = eps * abseps;
            for (int l = 0; l < neqn; ++l) y[l] = W.yy[l];
            t = W.x;
            break;
        }
        nostep = nostep + 1;
        kle4 = kle4 + 1;
        if (4 < W.kold) kle4 = 0;
        if (50 <= kle4) stiff = true;
    }
    return iflag;
}

// SG_ode (L/SG_ode_m.f90:89-159)
template <class R> inline void SG_ode(const rays_cfg &c, R *v, R &s, R &sout, OdeStop<R> &ray_stop, long &nrhs) {
    R rel_err = ray_stop.rel_err, abs_err = ray_stop.abs_err;
    for (;;) {
        int iflag = sg_de(c, c.nv, v, s, sout, rel_err, abs_err, ray_stop, nrhs);
        ray_stop.rel_err = rel_err;
        ray_stop.abs_err = abs_err;
        if (ray_stop.stop_ode) { sout = s; break; }
        if (iflag == 2) break;
        else if (iflag == 3) {
            R total_error = Fabs(rel_err) + Fabs(abs_err);
            if (total_error > c.SG_error_limit) {
                ray_stop.ode_stop_flag = RAYS_STOP_ODE_TOTAL_ERROR;
                ray_stop.stop_ode = true;
                break;
            }
            continue;
        } else { ray_stop.stop_ode = true; break; }
    }
}

}  // namespace rays_oracle
