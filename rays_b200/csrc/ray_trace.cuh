// ray_trace.cuh — per-ray engine on the device: eqn_ray, check_save, initialize_ode_vector, the RK4
// and Shampine-Gordon steppers and the persistent trace kernel that replaces the OpenMP ray loop of
// trace_rays (RAYS_project/RAYS_lib/ray_tracing.f90:62-266).
//
// Execution model: one fp64 ray per thread, ray state in registers.  Each CTA is persistent: a lane
// whose ray has ended takes the next ray index from a global counter (one warp-aggregated atomicAdd
// per refill: __ballot_sync compacts the requests), so ray-length divergence and the adaptive
// stepper's step-count divergence do not idle lanes until the fan is exhausted.  Trajectory points
// are stored straight into the reference's result layout ray_vec(nv, npoints, iray)
// (ray_results_m.f90:44-58): each lane writes the nv contiguous doubles of one point with 16-byte
// stores, consecutive points of a ray are adjacent, so L2 write-back merges them into full sectors
// and no transpose is needed before the copy to the host.
#pragma once
#include "ray_physics.cuh"

namespace rays_dev {

// compile-time description of one kernel specialisation; DAMP_/GRADS_ = -1 selects the run-time
// ("generic") form that handles every nv layout of ode_m.f90:160-173
template <int EQ_, int NS_, int DERIV_, int DAMP_, int GRADS_> struct Traits {
    static constexpr int EQ = EQ_, NS = NS_, DERIV = DERIV_;
    static constexpr bool GENERIC = DAMP_ < 0;
    static constexpr int DAMP = DAMP_, GRADS = GRADS_;
    static constexpr int NV = GENERIC ? RAYS_NV_MAX : 7 + (DAMP_ > 0 ? 1 : 0) + (GRADS_ > 0 ? 5 : 0);
    RD_INLINE static bool damp() { return GENERIC ? g_dc.c.damping_model != RAYS_DAMP_NONE : DAMP_ == 1; }
    RD_INLINE static bool grads() { return GENERIC ? g_dc.c.integrate_eq_gradients != 0 : GRADS_ == 1; }
    RD_INLINE static bool multi() { return GENERIC ? g_dc.c.multi_spec_damping != 0 : false; }
    RD_INLINE static int nv() { return GENERIC ? g_dc.c.nv : NV; }
};

// ---- eqn_ray (eqn_ray.f90:1-236), in three pieces so that the fused RK4 kernel can share one equilibrium
// evaluation between check_save and the first stage: derivatives of D, then the ray equations.
template <class T>
RD_INLINE int ray_derivs(const Eq<NSpec<T::NS>::MAX> &e, const double *v, double dddx[3], double dddk[3], double &dddw) {
    if (T::DERIV == RAYS_DERIV_COLD) {
        const Rcp K0 = g_dc.rc_k0;
        const double nvec[3] = {qdiv(v[3], K0), qdiv(v[4], K0), qdiv(v[5], K0)};
        deriv_cold<T::NS>(e, nvec, dddx, dddk, dddw);
        return 0;
    }
    int pert_err;
    const double r0[3] = {v[0], v[1], v[2]}, kv[3] = {v[3], v[4], v[5]};
    deriv_num<T::EQ, T::NS>(e, r0, kv, dddx, dddk, dddw, pert_err);
    return pert_err;
}
// eqn_ray.f90:126-229: group velocity, dr/ds, dk/ds, arc length, damping and gradient diagnostics
template <class T>
RD_INLINE int ray_equations(const Eq<NSpec<T::NS>::MAX> &e, const double *v, const double dddx[3], const double dddk[3], double dddw, double *dvds) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    const rays_cfg &c = g_dc.c;
    if (dddw == 0.0) return RAYS_STOP_INFINITE_VG_RHS;
    const Rcp W = rcp_of(dddw);
    const double vg[3] = {qdiv(-dddk[0], W), qdiv(-dddk[1], W), qdiv(-dddk[2], W)};
    const double vg0 = sqrt_rn(vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2]);
    double dsd;
    if (c.ray_param == RAYS_PARAM_ARCL) {
        if (dddk[0] != 0.0 || dddk[1] != 0.0 || dddk[2] != 0.0) {
            const double sg = copysign(1.0, dddw);
            const Rcp nrm = rcp_of(sqrt(dddk[0] * dddk[0] + dddk[1] * dddk[1] + dddk[2] * dddk[2]));
#pragma unroll
            for (int i = 0; i < 3; ++i) { dvds[i] = qdiv(-sg * dddk[i], nrm); dvds[3 + i] = qdiv(sg * dddx[i], nrm); }
            dsd = 1.0;
        } else return RAYS_STOP_RAY_STALLED;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) { dvds[i] = vg[i]; dvds[3 + i] = qdiv(dddx[i], W); }
        dsd = vg0;
    }
    dvds[6] = dsd;
    int nv0 = 7;
    if (T::damp()) {
        const double kv[3] = {v[3], v[4], v[5]};
        const double ki = damp_fund_ECH<NSM>(e, kv, vg);
        dvds[7] = dsd * 2.0 * ki * (1.0 - v[7]);
        nv0 = 8;
        if (T::multi()) {  // ksi(0) = ki, ksi(1:nspec) = 0 (damp_fund_ECH.f90:121-125)
            dvds[8] = dsd * 2.0 * ki * (1.0 - v[7]);
            for (int is = 1; is <= c.nspec; ++is) dvds[8 + is] = dsd * 2.0 * 0.0 * (1.0 - v[7]);
            nv0 = 9 + c.nspec;
        }
    }
    if (T::grads()) {
        const Rcp V0 = rcp_of(vg0);
        const double u0 = qdiv(vg[0], V0), u1 = qdiv(vg[1], V0), u2 = qdiv(vg[2], V0);
#pragma unroll
        for (int j = 0; j < 3; ++j) dvds[nv0 + j] = dsd * u0 * e.g[0][j] + dsd * u1 * e.g[1][j] + dsd * u2 * e.g[2][j];
        dvds[nv0 + 3] = dsd * u0 * e.gradns[0][0] + dsd * u1 * e.gradns[1][0] + dsd * u2 * e.gradns[2][0];
        dvds[nv0 + 4] = dsd * u0 * e.gradts0[0] + dsd * u1 * e.gradts0[1] + dsd * u2 * e.gradts0[2];
    }
    return 0;
}
// eqn_ray: returns 0 or the stop code
template <class T> RD_INLINE int eqn_ray(const double *v, double *dvds) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    Eq<NSM> e;
    equilibrium<T::EQ, T::NS, true>(v[0], v[1], v[2], e);
    if (e.err) return e.err;
    double dddx[3], dddk[3], dddw;
    const int pert_err = ray_derivs<T>(e, v, dddx, dddk, dddw);
    if (pert_err) return pert_err;
    return ray_equations<T>(e, v, dddx, dddk, dddw, dvds);
}

// ---- check_save (check_save.f90:1-161): residual + stop tests at a saved point -------------------------
// flag/stop follow ode_stop semantics: a flag may be set without stopping (equilibrium error, A.5 (X)).
// check_save_tests works on an equilibrium that is already evaluated; dddw_cold is dD/d(omega) of
// deriv_cold at the point (check_save always differentiates with deriv_cold, check_save.f90:79-87).
template <class T>
RD_INLINE void check_save_resid(const Eq<NSpec<T::NS>::MAX> &e, const double *v, double &resid, bool &stop, int &flag) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    const int ns = NSpec<T::NS>::n();
    const double kv[3] = {v[3], v[4], v[5]};
    double k3, k1;
    kpar_kperp(kv, e.bunit, k3, k1);
    resid = residual<NSM>(e, ns, k1, k3);
    if (resid > g_dc.c.dispersion_resid_limit) { stop = true; flag = RAYS_STOP_DISP_RESIDUAL; }
}
template <class T> RD_INLINE void check_save_tail(const double *v, double dddw_cold, bool &stop, int &flag) {
    if (!(fabs(dddw_cold) > DBL_MIN)) { stop = true; flag = RAYS_STOP_INFINITE_VG_CHECK; }
    if (T::damp() && v[7] > g_dc.c.total_damping_limit) { stop = true; flag = RAYS_STOP_TOTAL_ABSORPTION; }
}
template <class T> RD_INLINE void check_save(const double *v, double &resid, bool &stop, int &flag) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    const rays_cfg &c = g_dc.c;
    const Rcp K0 = g_dc.rc_k0;
    Eq<NSM> e;
    equilibrium<T::EQ, T::NS, true>(v[0], v[1], v[2], e);
    if (e.err) {
        flag = e.err;
        resid = 0.0;
        if (T::damp() && v[7] > c.total_damping_limit) { stop = true; flag = RAYS_STOP_TOTAL_ABSORPTION; }
        return;
    }
    check_save_resid<T>(e, v, resid, stop, flag);
    const double nvec[3] = {qdiv(v[3], K0), qdiv(v[4], K0), qdiv(v[5], K0)};
    double dddx[3], dddk[3], dddw;
    deriv_cold<T::NS>(e, nvec, dddx, dddk, dddw);
    check_save_tail<T>(v, dddw, stop, flag);
}

// ---- initialize_ode_vector (initialize_ode_vector.f90:1-57) ---------------------------------------------
template <class T> RD_INLINE void initialize_ode_vector(const double *rvec0, const double *rindex_vec0, double *v) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    const rays_cfg &c = g_dc.c;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v[i] = rvec0[i]; v[3 + i] = c.k0 * rindex_vec0[i]; }
    v[6] = 0.0;
    int nv0 = 7;
    if (T::damp()) {
        v[7] = 0.0; nv0 = 8;
        if (T::multi()) { for (int s = 0; s <= c.nspec; ++s) v[8 + s] = 0.0; nv0 = 9 + c.nspec; }
    }
    if (T::grads()) {
        Eq<NSM> e;
        equilibrium<T::EQ, T::NS, false>(v[0], v[1], v[2], e);
        const bool ok = e.err == 0;   // on error the reference leaves eq undefined; zeros here (oracle choice)
        v[nv0] = ok ? e.bvec[0] : 0.0; v[nv0 + 1] = ok ? e.bvec[1] : 0.0; v[nv0 + 2] = ok ? e.bvec[2] : 0.0;
        v[nv0 + 3] = ok ? e.ns[0] : 0.0;
        v[nv0 + 4] = ok ? e.ts[0] : 0.0;
    }
}

// ---- RK4_ode (RK4_ode_m.f90:59-94): returns the stop code, v and s untouched on a stop ----------------
template <class T> RD_INLINE int RK4_ode(double *v, double &s, double sout) {
    constexpr int NV = T::NV;
    const int nv = T::nv();
    const double ds = sout - s;
    double w[NV], acc[NV], f[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { w[i] = v[i]; acc[i] = 0.0; f[i] = 0.0; }
    int code = 0;
#pragma unroll 1
    for (int stage = 0; stage < 4; ++stage) {
        code = eqn_ray<T>(w, f);
        if (code) break;
        // (f1 + 2.0*f2 + 2.0*f3 + f4), left to right; v + ds*f/2.0 for stages 1,2 and v + ds*f for 3
        const double ca = (stage == 1 || stage == 2) ? 2.0 : 1.0;
        const double cw = stage == 2 ? 1.0 : 0.5;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < nv) {
                acc[i] = stage == 0 ? f[i] : acc[i] + ca * f[i];
                w[i] = v[i] + ds * f[i] * cw;
            }
    }
    if (code) return code;
#pragma unroll
    for (int i = 0; i < NV; ++i) if (i < nv) v[i] = v[i] + qdiv(ds * acc[i], g_dc.rc_six);
    s = sout;
    return 0;
}

// ---- SG_ode (SG_ode_m.f90:89-159) + ode/de/step/intrp (ode_RAYS.f90:1-1362) -----------------------------
// `work`/`iwork` are automatic in SG_ode and iflag = 1 on every call, so each ds segment restarts the
// integrator at order 1; the work arrays therefore live only inside sg_de, in local
// memory (phi is indexed by the run-time order k).  Arrays keep the Fortran's 1-based subscripts.
template <int NV> struct SGWork {
    double yy[NV], wt[NV], p[NV], yp[NV];
    double phi[17][NV];
    double alpha[13], beta[13], sig[14], v[13], w[13], g[14], psi[13];
    double x, h, hold;
    bool start, phase1, nornd;
    int ns, k, kold;
    // values the reference keeps in locals of `step` across its derivative evaluations
    double round, absh, xold, erk, erkm1;
    int ifail, knew;
};
__device__ const double kSGgstr[14] = {  // ode_RAYS.f90:776-779 (single-precision literals)
    0.0, (double)0.50e+00f, (double)0.0833e+00f, (double)0.0417e+00f, (double)0.0264e+00f, (double)0.0188e+00f,
    (double)0.0143e+00f, (double)0.0114e+00f, (double)0.00936e+00f, (double)0.00789e+00f, (double)0.00679e+00f,
    (double)0.00592e+00f, (double)0.00524e+00f, (double)0.00468e+00f};
// 1/i and 1/(i(i+1)) as the compiler's IEEE division forms them: the same bits as the run-time 1.0/real(i)
__device__ const double kSGinv[15] = {0.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 3.0, 1.0 / 4.0, 1.0 / 5.0, 1.0 / 6.0, 1.0 / 7.0, 1.0 / 8.0, 1.0 / 9.0, 1.0 / 10.0,
                                      1.0 / 11.0, 1.0 / 12.0, 1.0 / 13.0, 1.0 / 14.0};
__device__ const double kSGinvTri[14] = {0.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 12.0, 1.0 / 20.0, 1.0 / 30.0, 1.0 / 42.0, 1.0 / 56.0, 1.0 / 72.0, 1.0 / 90.0,
                                         1.0 / 110.0, 1.0 / 132.0, 1.0 / 156.0, 1.0 / 182.0};
__device__ const double kSGtwo[14] = {0.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0, 1024.0, 2048.0, 4096.0, 8192.0};

// step (ode_RAYS.f90:595-1234) is cut at its three right-hand-side evaluations so that a warp evaluates the
// RHS of all its lanes at ONE place per loop iteration, whatever phase of the integrator each lane is in
// (the straight transcription called eqn_ray from three divergent sites and ran at 1 % of the FP64 peak).
// The arithmetic of every piece is the reference's, statement by statement.
enum SGPhase { SG_IDLE = 0, SG_CHECK, SG_DE_BEGIN, SG_AFTER_CHECK, SG_AFTER_R1, SG_AFTER_R2, SG_AFTER_R3, SG_RETRY };

// IEEE quotient / square root through the branch-free exact helpers of ray_physics.cuh (9 and 10 instructions inline
// instead of ~20 / ~25 with a slow-path call each; the SG kernel is instruction-cache bound).  A zero divisor can only be
// an error weight wt = releps*|y| + abseps with abs_err0 = 0 and y = 0: x/(+0) = x*inf, as IEEE has it.
RD_INLINE double sg_quot(double x, const Rcp &c) { return c.d == 0.0 ? x * __longlong_as_double(0x7ff0000000000000LL) : qdiv(x, c); }
RD_INLINE double sg_sqrt(double x) { return x == __longlong_as_double(0x7ff0000000000000LL) ? x : sqrt_rn(x); }   // sqrt_rn is for finite radicands
RD_INLINE double sg_div(double x, double d) { return sg_quot(x, rcp_of(d)); }
// (p5eps/erk)**(1/(k+1)) of the step-size formula (ode_RAYS.f90:1222): the square root is the correctly rounded value of
// x**0.5 and the fourth root two of them (within an ulp of it, like CUDA's pow), at a tenth of pow's instructions; 0 < x < 1 here
RD_INLINE double sg_root(double x, int kp1) {
    if (kp1 == 2) return sqrt_rn(x);
    if (kp1 == 4) return sqrt_rn(sqrt_rn(x));
    return pow_ool(x, kSGinv[kp1]);
}
// step, first block (:840-852): tests for too small a step / tolerance; returns true on crash
template <int NV> RD_INLINE bool sg_block0(int neqn, SGWork<NV> &W, double &eps) {
    const double twou = 2.0 * DBL_EPSILON, fouru = 2.0 * twou;
    if (fabs(W.h) < fouru * fabs(W.x)) { W.h = copysign(fouru * fabs(W.x), W.h); return true; }
    const double p5eps = 0.5 * eps;
    double sum = 0.0;
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) { const double q = sg_div(W.yy[l], W.wt[l]); sum = sum + q * q; }
    W.round = twou * sg_sqrt(sum);
    if (p5eps < W.round) { eps = 2.0 * W.round * (1.0 + fouru); return true; }
    W.g[1] = 1.0; W.g[2] = 0.5; W.sig[1] = 1.0;
    W.ifail = 0;
    return false;
}
// step, start block after the first derivative evaluation (:858-885)
template <int NV> RD_INLINE void sg_after_start(int neqn, SGWork<NV> &W, double eps) {
    const double fouru = 4.0 * DBL_EPSILON;
    const double p5eps = 0.5 * eps;
    double tot = 0.0;
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) {
        W.phi[1][l] = W.yp[l]; W.phi[2][l] = 0.0;
        const double q = sg_div(W.yp[l], W.wt[l]); tot = tot + q * q;
    }
    const double total = sg_sqrt(tot);
    double absh = fabs(W.h);
    if (eps < 16.0 * total * W.h * W.h) absh = 0.25 * sg_sqrt(sg_div(eps, total));
    W.h = copysign(fmax(absh, fouru * fabs(W.x)), W.h);
    W.hold = 0.0;
    W.k = 1; W.kold = 0;
    W.start = false; W.phase1 = true; W.nornd = true;
    if (p5eps <= 100.0 * W.round) {
        W.nornd = false;
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) W.phi[15][l] = 0.0;
    }
}
// step, blocks 1 and 2 (:896-1015): coefficients for this step size/order, then the predicted solution p at x + h
template <int NV> RD_INLINE void sg_predict(int neqn, SGWork<NV> &W) {
    double *alpha = W.alpha, *beta = W.beta, *sig = W.sig, *v = W.v, *w = W.w, *g = W.g, *psi = W.psi;
    double(*phi)[NV] = W.phi;
    const int k = W.k, kold = W.kold;
    int ns = W.ns;
    const double h = W.h;
    const int kp1 = k + 1, kp2 = k + 2;
    if (h != W.hold) ns = 0;
    if (ns <= kold) ns = ns + 1;
    const int nsp1 = ns + 1;
    if (ns <= k) {
        beta[ns] = 1.0;
        alpha[ns] = kSGinv[ns];
        double temp1 = h * (double)ns;
        sig[nsp1] = 1.0;
        #pragma unroll 1
        for (int i = nsp1; i <= k; ++i) {
            const double temp2 = psi[i - 1];
            psi[i - 1] = temp1;
            beta[i] = sg_div(beta[i - 1] * psi[i - 1], temp2);
            temp1 = temp2 + h;
            alpha[i] = sg_div(h, temp1);
            sig[i + 1] = (double)i * alpha[i] * sig[i];
        }
        psi[k] = temp1;
        if (ns <= 1) {
            #pragma unroll 1
            for (int iq = 1; iq <= k; ++iq) { v[iq] = kSGinvTri[iq]; w[iq] = v[iq]; }
        } else {
            if (kold < k) {
                v[k] = kSGinvTri[k];
                #pragma unroll 1
                for (int j = 1; j <= ns - 2; ++j) { const int i = k - j; v[i] = v[i] - alpha[j + 1] * v[i + 1]; }
            }
            #pragma unroll 1
            for (int iq = 1; iq <= kp1 - ns; ++iq) { v[iq] = v[iq] - alpha[ns] * v[iq + 1]; w[iq] = v[iq]; }
            g[nsp1] = w[1];
        }
        #pragma unroll 1
        for (int i = ns + 2; i <= kp1; ++i) {
            #pragma unroll 1
            for (int iq = 1; iq <= kp2 - i; ++iq) w[iq] = w[iq] - alpha[i - 1] * w[iq + 1];
            g[i] = w[1];
        }
    }
    W.ns = ns;
    #pragma unroll 1
    for (int i = nsp1; i <= k; ++i)
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) phi[i][l] = beta[i] * phi[i][l];
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) {   // component by component (same operation order per component as the reference's
        phi[kp2][l] = phi[kp1][l];     // i-outer loops): the running sum and phi(l,i+1) stay in registers
        phi[kp1][l] = 0.0;
        double pl = 0.0, up = 0.0;
        #pragma unroll 1
        for (int i = k; i >= 1; --i) {
            const double f = phi[i][l];
            pl = pl + f * g[i];
            up = f + up;
            phi[i][l] = up;
        }
        W.p[l] = pl;
    }
    if (!W.nornd) {
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) {
            const double tau = h * W.p[l] - phi[15][l];
            W.p[l] = W.yy[l] + tau;
            phi[16][l] = (W.p[l] - W.yy[l]) - tau;
        }
    } else {
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) W.p[l] = W.yy[l] + h * W.p[l];
    }
    W.xold = W.x;
    W.x = W.x + h;
    W.absh = fabs(h);
}
// step, after the derivative at the predicted point (:1022-1110): error estimates, accept or fail.
// returns 0 = accepted (corrected solution formed in yy, evaluate the derivative there next),
//         1 = failed, retry with the reduced step, 2 = crash (step size too small: eps doubled)
template <int NV> RD_INLINE int sg_after_predict(int neqn, SGWork<NV> &W, double &eps) {
    const double fouru = 4.0 * DBL_EPSILON;
    double(*phi)[NV] = W.phi;
    const double *yp = W.yp, *wt = W.wt, *sig = W.sig, *g = W.g;
    const int k = W.k, kp1 = k + 1, km1 = k - 1, km2 = k - 2;
    const double absh = W.absh, p5eps = 0.5 * eps;
    double erkm2 = 0.0, erkm1 = 0.0, erk = 0.0;
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) {
        const Rcp wl = rcp_of(wt[l]);   // up to three quotients by the same weight
        if (0 < km2) { const double q = sg_quot(phi[km1][l] + yp[l] - phi[1][l], wl); erkm2 = erkm2 + q * q; }
        if (0 <= km2) { const double q = sg_quot(phi[k][l] + yp[l] - phi[1][l], wl); erkm1 = erkm1 + q * q; }
        const double q = sg_quot(yp[l] - phi[1][l], wl);
        erk = erk + q * q;
    }
    if (0 < km2) erkm2 = absh * sig[km1] * kSGgstr[km2] * sg_sqrt(erkm2);
    if (0 <= km2) erkm1 = absh * sig[k] * kSGgstr[km1] * sg_sqrt(erkm1);
    const double rt_erk = sg_sqrt(erk);
    const double err = absh * rt_erk * (g[k] - g[kp1]);
    erk = absh * rt_erk * sig[kp1] * kSGgstr[k];
    int knew = k;
    if (0 < km2) {
        if (fmax(erkm1, erkm2) <= erk) knew = km1;
    } else if (0 == km2) {
        if (erkm1 <= 0.5 * erk) knew = km1;
    }
    W.knew = knew; W.erk = erk; W.erkm1 = erkm1;
    if (err <= eps) {   // accepted: correct (:1123-1141)
        const double h = W.h;
        W.kold = k;
        W.hold = h;
        if (!W.nornd) {
            #pragma unroll 1
            for (int l = 0; l < neqn; ++l) {
                const double rho = h * g[kp1] * (yp[l] - phi[1][l]) - phi[16][l];
                W.yy[l] = W.p[l] + rho;
                phi[15][l] = (W.yy[l] - W.p[l]) - rho;
            }
        } else {
            #pragma unroll 1
            for (int l = 0; l < neqn; ++l) W.yy[l] = W.p[l] + h * g[kp1] * (yp[l] - phi[1][l]);
        }
        return 0;
    }
    // step failed (:1076-1110): restore x, phi, psi; halve (or more) the step
    W.phase1 = false;
    W.x = W.xold;
    #pragma unroll 1
    for (int i = 1; i <= k; ++i) {
        const Rcp bi = rcp_of(W.beta[i]);
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) phi[i][l] = sg_quot(phi[i][l] - phi[i + 1][l], bi);
    }
    #pragma unroll 1
    for (int i = 2; i <= k; ++i) W.psi[i - 1] = W.psi[i] - W.h;
    W.ifail = W.ifail + 1;
    double temp2 = 0.5;
    if (3 < W.ifail) { if (p5eps < 0.25 * erk) temp2 = sg_sqrt(sg_div(p5eps, erk)); }
    if (3 <= W.ifail) knew = 1;
    W.h = temp2 * W.h;
    W.k = knew;
    if (fabs(W.h) < fouru * fabs(W.x)) {
        W.h = copysign(fouru * fabs(W.x), W.h);
        eps = eps + eps;
        return 2;
    }
    return 1;
}
// step, after the derivative at the corrected point (:1147-1231): update differences, choose order and step
template <int NV> RD_INLINE void sg_after_correct(int neqn, SGWork<NV> &W, double eps) {
    const double fouru = 4.0 * DBL_EPSILON;
    double(*phi)[NV] = W.phi;
    const double *yp = W.yp, *wt = W.wt;
    int k = W.k;
    const int kp1 = k + 1, kp2 = k + 2, km1 = k - 1, knew = W.knew, ns = W.ns;
    const double absh = W.absh, p5eps = 0.5 * eps, h = W.h, erkm1 = W.erkm1;
    double erk = W.erk;
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) {
        phi[kp1][l] = yp[l] - phi[1][l];
        phi[kp2][l] = phi[kp1][l] - phi[kp2][l];
    }
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) {
        const double d = phi[kp1][l];
        #pragma unroll 1
        for (int i = 1; i <= k; ++i) phi[i][l] = phi[i][l] + d;
    }
    double erkp1 = 0.0;
    if (knew == km1 || k == 12) W.phase1 = false;
    if (W.phase1) {
        k = kp1; erk = erkp1;
    } else if (knew == km1) {
        k = km1; erk = erkm1;
    } else if (kp1 <= ns) {
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) { const double q = sg_div(phi[kp2][l], wt[l]); erkp1 = erkp1 + q * q; }
        erkp1 = absh * kSGgstr[kp1] * sg_sqrt(erkp1);
        if (k == 1) {
            if (erkp1 < 0.5 * erk) { k = kp1; erk = erkp1; }
        } else if (erkm1 <= fmin(erk, erkp1)) {
            k = km1; erk = erkm1;
        } else if (erkp1 < erk && k < 12) {
            k = kp1; erk = erkp1;
        }
    }
    double hnew = h + h;
    if (!W.phase1) {
        if (p5eps < erk * kSGtwo[k + 1]) {
            hnew = h;
            if (p5eps < erk) {
                const double r = sg_root(sg_div(p5eps, erk), k + 1);
                hnew = absh * fmax(0.5, fmin((double)0.9f, r));
                hnew = copysign(fmax(hnew, fouru * fabs(W.x)), h);
            }
        }
    }
    W.k = k;
    W.h = hnew;
}

// intrp (ode_RAYS.f90:1235-1362); ypout is not used by the caller
template <int NV> RD_INLINE void sg_intrp(int neqn, const SGWork<NV> &W, double xout, double *yout) {
    double g[14], rho[14], w[14];
    const double hi = xout - W.x;
    const int ki = W.kold + 1;
    #pragma unroll 1
    for (int i = 1; i <= ki; ++i) w[i] = kSGinv[i];
    g[1] = 1.0; rho[1] = 1.0;
    double term = 0.0;
    #pragma unroll 1
    for (int j = 2; j <= ki; ++j) {
        const double psijm1 = W.psi[j - 1];
        const Rcp pj = rcp_of(psijm1);
        const double gamma = sg_quot(hi + term, pj);
        const double eta = sg_quot(hi, pj);
        #pragma unroll 1
        for (int i = 1; i <= ki + 1 - j; ++i) w[i] = gamma * w[i] - eta * w[i + 1];
        g[j] = w[1];
        rho[j] = gamma * rho[j - 1];
        term = psijm1;
    }
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) yout[l] = 0.0;
    #pragma unroll 1
    for (int j = 1; j <= ki; ++j) {
        const int i = ki + 1 - j;
        #pragma unroll 1
        for (int l = 0; l < neqn; ++l) yout[l] = yout[l] + g[i] * W.phi[i][l];
    }
    #pragma unroll 1
    for (int l = 0; l < neqn; ++l) yout[l] = W.yy[l] + hi * yout[l];
}

// ---- the trace kernel ------------------------------------------------------------------------------------
struct TraceArgs {
    long long nray;                 // rays in this launch
    const double *rvec0;            // [nray][3]   (Fortran rvec0(3,nray))
    const double *rindex_vec0;      // [nray][3]
    const double *ray_pwr_wt;       // [nray]
    double *ray_vec;                // [nray][npoints_alloc][nv] or NULL (no trajectory storage)
    double *residual;               // [nray][npoints_alloc] or NULL
    int npoints_alloc;
    // streaming copy-out: when host_ray_vec/host_residual are set (device-accessible pinned host memory in
    // the reference layout, row pitch host_npoints_alloc), ray_vec/residual are per-LANE staging rows
    // ([grid*block][npoints_alloc][nv]) and the warp copies each finished ray to the host as it ends
    double *host_ray_vec, *host_residual;
    int host_npoints_alloc;
    long long host_ray0;            // row of this launch's first ray in the host arrays ...
    long long host_ray_stride;      // ... and the row distance of consecutive rays (1; ngpu for the interleaved shards of rays_b200_trace_multi)
    int *npoints;                   // [nray]
    int *stop_code;                 // [nray]
    double *initial_ray_power, *end_residuals, *max_residuals, *end_ray_parameter;  // [nray]
    double *start_ray_vec, *end_ray_vec;   // [nray][nv]
    const int *order;               // optional: queue position -> ray index (resume list / explicit schedule)
    // Time slicing.  Ray lengths are very unequal (3.4 % of the bench fan runs 5x the mean) and the long rays
    // are scattered over all warps, so after the queue empties every warp would crawl on with one or two live
    // lanes (measured: 19 % of all lane slots idle).  With slice_steps > 0 a ray that has taken that many steps
    // in this launch is suspended: its state goes to cont_state[iray], its index to cont_list; the host then
    // launches the kernel again over that list (resume = 1), where the survivors are packed into full warps.
    int slice_steps;
    int resume;
    int sg_align;                   // SG kernel: 1 = lanes advance in alternating predictor / corrector slots (see trace_sg_kernel)
    double *sg_state;               // slot-machine SG kernel (ray_trace_sg2.cuh): global slot records, grid * sg_state_bytes_per_cta bytes
    int sg_slots;                   // ... and the ray slots of one CTA (their hot records are in dynamic shared memory)
    int sg_mixed;                   // ... 1: every warp of an iteration takes a batch of whatever kind is waiting; 0: one kind per iteration; 2: predictor + corrector
    double *cont_state;             // [nray][kContStride]
    int *cont_list;
    unsigned long long *cont_count;
    unsigned long long *queue;      // next ray index to hand out
    unsigned long long *counters;   // [0] ray-steps, [1] RHS evaluations
    // copy-out by a concurrent copier kernel (copy_out_kernel): a ray that has ENDED (not merely been suspended) appends its index
    // to done_list after its trajectory, npoints and summaries are written (done_list is preset to -1; the copier polls the entries)
    int *done_list;                 // [nray] or NULL
    unsigned long long *done_count;
    int *trace_started;             // set by the first trace kernel that runs: tells the copier that it is NOT being run alone
    // fused deposition binning (bin_to_uniform_grid_m.f90:155-266); dep_acc == NULL disables it.  Bins are 64-bit FIXED-POINT
    // accumulators (one unit = 1/dep_scale): integer addition is associative, so the profile does not depend on the order in
    // which rays, CTAs or GPUs contribute (SURVEY.md 8e "reproducibility": 1/2/4/8-GPU profiles are bitwise identical).  Each
    // CTA bins into shared memory (dep_smem = n_bins * 8 bytes of dynamic shared memory) and adds its partial profile to
    // dep_acc once, at the end of the kernel; dep_smem = 0 (more bins than shared memory holds) bins straight into dep_acc.
    unsigned long long *dep_acc;    // [n_bins] global accumulator
    int n_bins;
    int dep_smem;
    double grid_min, grid_max, dep_scale;
};
struct DepBins { unsigned long long *acc; int n_bins; double xmin, xmax, scale; };
RD_INLINE void dep_add(const DepBins &b, int i, double x) {
    atomicAdd(b.acc + i, (unsigned long long)__double2ll_rn(x * b.scale));
}

RD_INLINE void store_point(double *dst, const double *v, int nv) {
    // dst is 8-byte aligned; pair up into 16-byte stores when the row start allows it
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        int i = 0;
        for (; i + 1 < nv; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(v[i], v[i + 1]);
        if (i < nv) dst[i] = v[i];
    } else {
        dst[0] = v[0];
        int i = 1;
        for (; i + 1 < nv; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(v[i], v[i + 1]);
        if (i < nv) dst[i] = v[i];
    }
}
template <int NV> RD_INLINE void store_point_fixed(double *dst, const double (&v)[NV]) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int i = 0; i + 1 < NV; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(v[i], v[i + 1]);
        if (NV & 1) dst[NV - 1] = v[NV - 1];
    } else {
        dst[0] = v[0];
#pragma unroll
        for (int i = 1; i + 1 < NV; i += 2) *reinterpret_cast<double2 *>(dst + i) = make_double2(v[i], v[i + 1]);
        if (!(NV & 1)) dst[NV - 1] = v[NV - 1];
    }
}

// binner_real for one segment (bin_to_uniform_grid_m.f90:179-262): spreads delta_Q = Q1 - Q0 uniformly
// over the bins spanned by [x0, x1]
RD_INLINE void bin_segment(const DepBins &bins, double xa, double xb, double Qa, double Qb) {
    const int n_bins = bins.n_bins;
    const double xmin = bins.xmin, xmax = bins.xmax;
    const double x_bin_width = (xmax - xmin) / n_bins;
    const double x_low = fmin(xa, xb), x_high = fmax(xa, xb);
    double ix_low = (x_low - xmin) / x_bin_width, ix_high = (x_high - xmin) / x_bin_width;
    const double delta_ix = ix_high - ix_low;
    int index_low = (int)floor(ix_low) + 1, index_high = (int)floor(ix_high) + 1;
    if (x_high >= xmax) index_high = n_bins;
    int delta_i = index_high - index_low;
    double delta_Q = Qb - Qa;
    const double Q_density = delta_Q / delta_ix;
    if (fabs(delta_Q) < 4.0 * DBL_MIN) return;
    if (x_high < xmin || x_low > xmax) return;
    if (x_low < xmin) {
        delta_Q = delta_Q * (ix_high / delta_ix);
        ix_low = 0.0; index_low = 1;
        delta_i = index_high - index_low;
    }
    if (x_high > xmax) {
        delta_Q = delta_Q * (((double)n_bins - ix_low) / delta_ix);
        ix_high = (double)n_bins; index_high = n_bins;
        delta_i = index_high - index_low;
    }
    if (index_low < 1 || index_high > n_bins) return;   // NaN abscissa (a point outside the equilibrium's domain): nothing to bin
    if (delta_i == 0) dep_add(bins, index_low - 1, delta_Q);
    else if (delta_i > 0) {
        dep_add(bins, index_low - 1, delta_Q * (((double)index_low - ix_low) / delta_ix));
        dep_add(bins, index_high - 1, delta_Q * ((ix_high - (double)(index_high - 1)) / delta_ix));
        for (int i = index_low + 1; i <= index_high - 1; ++i) dep_add(bins, i - 1, Q_density);
    }
}
// per-CTA partial profile in dynamic shared memory: cleared at kernel start, added to the global accumulator at kernel end
extern __shared__ unsigned long long s_dep_bins[];
RD_INLINE DepBins dep_begin(const TraceArgs &a, bool binning) {
    DepBins b{nullptr, a.n_bins, a.grid_min, a.grid_max, a.dep_scale};
    if (!binning) return b;
    if (a.dep_smem > 0) {
        for (int i = threadIdx.x; i < a.n_bins; i += blockDim.x) s_dep_bins[i] = 0ULL;
        __syncthreads();
        b.acc = s_dep_bins;
    } else b.acc = a.dep_acc;
    return b;
}
RD_INLINE void dep_end(const TraceArgs &a, bool binning) {
    if (!binning || a.dep_smem <= 0) return;
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_bins; i += blockDim.x) { const unsigned long long q = s_dep_bins[i]; if (q) atomicAdd(a.dep_acc + i, q); }
}
// abscissa of the deposition profile: Ptotal_x (slab) / Ptotal_psi (axisym_toroid)
// (deposition_profiles_m.f90:438-499)
template <int EQ_> RD_INLINE double dep_abscissa(const double *v) {
    if (EQ_ == RAYS_EQ_SLAB) return v[0];
    if (EQ_ == RAYS_EQ_AXISYM_TOROID && g_dc.c.axisym.magnetics_model == RAYS_MAG_EQDSK_SPLINE) return eqdsk_psiN(v[0], v[1], v[2]);
    return solovev_psiN(v[0], v[1], v[2]);
}

constexpr int kTraceBlock = 128;
// Shampine-Gordon slot machine (ray_trace_sg2.cuh): kSgCtas CTAs of kSgBlock threads per SM, each owning as many ray slots as its
// share of the SM's shared memory holds (TraceArgs::sg_slots)
#ifndef RAYS_SG_BLOCK
#define RAYS_SG_BLOCK 256
#endif
#ifndef RAYS_SG2_MIN_CTAS
#define RAYS_SG2_MIN_CTAS 1
#endif
constexpr int kSgBlock = RAYS_SG_BLOCK;
constexpr int kSgCtas = RAYS_SG2_MIN_CTAS;
constexpr int kContStride = RAYS_NV_MAX + 11;   // v[nv], s, sout, nstep, flag, resid_prev/last/max, dep_x, dep_Q, rel_err, abs_err

// one warp copies n doubles HBM/L2 -> pinned host memory: 8 loads in flight per lane (2 KB per warp) before the
// 256-byte coalesced stores, so that the copy is bandwidth- rather than latency-bound (a 13 KB ray takes ~7
// round trips instead of ~50; the latency-bound version cost the warp about one ray-step per finished ray)
#ifndef RAYS_COPY_ALIGN
#define RAYS_COPY_ALIGN 128
#endif
RD_INLINE void copy_row_to_host(double *__restrict__ dst, const double *__restrict__ src, int n, unsigned lane) {
    // 16-byte stores on a destination aligned to RAYS_COPY_ALIGN bytes (512 contiguous bytes = four whole 128-byte lines per warp
    // instruction): the head elements go first, one per lane, when the row starts elsewhere (rows of the reference layout start
    // on 8-byte boundaries only).  Measured end to end, interleaved runs on one box: 16-byte alignment 96.2 / 95.9 ms, 128-byte
    // 91.5 / 92.1 ms, 256-byte 94.3 / 92.8 ms (262k-ray fan); 317.4 -> 306.5 ms on the 1M-ray fan.
    int head = (int)((RAYS_COPY_ALIGN - (reinterpret_cast<uintptr_t>(dst) & (RAYS_COPY_ALIGN - 1))) & (RAYS_COPY_ALIGN - 1)) >> 3;
    if (head > n) head = n;
    if ((int)lane < head) dst[lane] = __ldcg(src + lane);
    dst += head; src += head; n -= head;
    const int n2 = n >> 1;     // pairs
    double2 *__restrict__ d2 = reinterpret_cast<double2 *>(dst);
    int i = (int)lane;
    for (; i + 32 * 3 < n2; i += 32 * 4) {
        double2 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = 2 * (i + 32 * k); t[k] = make_double2(__ldcg(src + j), __ldcg(src + j + 1)); }
#pragma unroll
        for (int k = 0; k < 4; ++k) d2[i + 32 * k] = t[k];
    }
    for (; i < n2; i += 32) d2[i] = make_double2(__ldcg(src + 2 * i), __ldcg(src + 2 * i + 1));
    if ((n & 1) && lane == 0) dst[n - 1] = __ldcg(src + n - 1);
}

// suspended-ray record (time slicing): everything a lane needs to carry on with the ray in a later launch
struct RayCarry { double s, sout, resid_prev, resid_last, resid_max, dep_x, dep_Q, rel_err, abs_err; int nstep, flag; };
RD_INLINE void suspend_ray(const TraceArgs &a, long long iray, const double *v, int nv, const RayCarry &c) {
    double *r = a.cont_state + (size_t)iray * kContStride;
    for (int i = 0; i < nv; ++i) r[i] = v[i];
    double *q = r + RAYS_NV_MAX;
    q[0] = c.s; q[1] = c.sout; q[2] = (double)c.nstep; q[3] = (double)c.flag; q[4] = c.resid_prev; q[5] = c.resid_last;
    q[6] = c.resid_max; q[7] = c.dep_x; q[8] = c.dep_Q; q[9] = c.rel_err; q[10] = c.abs_err;
    const unsigned long long pos = atomicAdd(a.cont_count, 1ULL);
    a.cont_list[pos] = (int)iray;
}
RD_INLINE void resume_ray(const TraceArgs &a, long long iray, double *v, int nv, RayCarry &c) {
    const double *r = a.cont_state + (size_t)iray * kContStride;
    for (int i = 0; i < nv; ++i) v[i] = r[i];
    const double *q = r + RAYS_NV_MAX;
    c.s = q[0]; c.sout = q[1]; c.nstep = (int)q[2]; c.flag = (int)q[3]; c.resid_prev = q[4]; c.resid_last = q[5];
    c.resid_max = q[6]; c.dep_x = q[7]; c.dep_Q = q[8]; c.rel_err = q[9]; c.abs_err = q[10];
}

// Warp-cooperative copy-out of the rays that ended in this iteration: every lane of the warp moves a
// slice of each finished ray's staged trajectory (HBM/L2) to the caller's arrays in pinned host memory,
// 256-byte coalesced stores over PCIe, overlapped with the integration of the other rays.
RD_INLINE void flush_finished_rays(const TraceArgs &a, bool finished, long long iray, int npts, int p0, size_t row, int nv, unsigned lane) {
    unsigned m = __ballot_sync(0xffffffffu, finished);
    if (RAYS_USUAL(m == 0u || a.host_ray_vec == nullptr && a.host_residual == nullptr)) return;
    __syncwarp();   // orders the finished lanes' trajectory stores before the other lanes' loads
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const long long ir = __shfl_sync(0xffffffffu, iray, l);
        const int np = __shfl_sync(0xffffffffu, npts, l);          // points staged in this lane's row ...
        const int pf = __shfl_sync(0xffffffffu, p0, l);            // ... the first of which is point pf of the ray
        const unsigned long long rw = __shfl_sync(0xffffffffu, (unsigned long long)row, l);
        if (a.host_ray_vec) {
            const double *src = a.ray_vec + (size_t)rw * a.npoints_alloc * nv;
            double *dst = a.host_ray_vec + ((size_t)(a.host_ray0 + ir * a.host_ray_stride) * a.host_npoints_alloc + pf) * nv;
            copy_row_to_host(dst, src, np * nv, lane);
        }
        if (a.host_residual) {
            const double *src = a.residual + (size_t)rw * a.npoints_alloc;
            double *dst = a.host_residual + (size_t)(a.host_ray0 + ir * a.host_ray_stride) * a.host_npoints_alloc + pf;
            copy_row_to_host(dst, src, np, lane);
        }
    }
}

// ---- Shampine-Gordon trace kernel ---------------------------------------------------------------------------
// SG_ode (SG_ode_m.f90:89-159) -> ode/de (ode_RAYS.f90:1-593) -> step/intrp as a per-lane state machine.
// One loop iteration evaluates at most one right-hand side per lane, at a single convergence point of the
// warp; the integrator bookkeeping between two evaluations is lane-private.  `work`/`iwork` are automatic in
// SG_ode and iflag = 1 on every call, so each ds segment restarts the integrator at order 1:
// the history W lives in per-thread local memory and is re-initialised per segment.
#ifndef RAYS_SG_MIN_CTAS
#define RAYS_SG_MIN_CTAS 2
#endif
template <class T>
__global__ void __launch_bounds__(kTraceBlock, RAYS_SG_MIN_CTAS) trace_sg_kernel(const TraceArgs a) {
    constexpr int NV = T::NV;
    constexpr int NSM = NSpec<T::NS>::MAX;
    const int nv = T::nv();
    const rays_cfg &c = g_dc.c;
    const unsigned lane = threadIdx.x & 31;
    const int maxnum = 500;
    const double fouru = 4.0 * DBL_EPSILON;
    // ray state
    double v[NV];
    double s = 0.0, sout = 0.0, rel_err = 0.0, abs_err = 0.0;
    double resid_prev = 0.0, resid_last = 0.0, resid_max = 0.0;
    double dep_x = 0.0, dep_Q = 0.0, pwr = 0.0;
    long long iray = -1;
    int nstep = 0, flag = 0;
    int st = SG_IDLE;
    bool exhausted = false, first = false;
    // de state of the current segment
    SGWork<NV> W;
    double eps = 0.0, absdel = 0.0, tend = 0.0, releps = 0.0, abseps = 0.0, t = 0.0;
    int nostep = 0, kle4 = 0;
    bool stiff = false;
    unsigned long long my_steps = 0;
    unsigned my_rhs = 0;
    const bool binning = a.dep_acc != nullptr && T::damp();
    const DepBins dbins = dep_begin(a, binning);
    const bool streaming = a.host_ray_vec != nullptr || a.host_residual != nullptr;
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t row = 0;
    bool fin = false;
    int fin_np = 0;
    int p0 = 0;          // index (within the ray) of the first point staged in this lane's row (streaming + resume)
    int slice_n = 0;     // steps this ray has taken in this launch
    bool have_f1 = false;     // W.yp holds the derivative at v evaluated together with check_save
    int f1_code = 0;
    bool slot_b = false;      // warp-uniform: type of the current slot (see the alignment note in the loop)

    for (;;) {
        // ---- refill from the work queue (one atomic per warp)
        const unsigned want = __ballot_sync(0xffffffffu, st == SG_IDLE && !exhausted);
        if (want) {
            unsigned long long base = 0;
            const int leader = __ffs(want) - 1;
            if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (st == SG_IDLE && !exhausted) {
                const long long idx = (long long)(base + __popc(want & ((1u << lane) - 1u)));
                if (idx >= a.nray) exhausted = true;
                else {
                    iray = a.order ? (long long)a.order[idx] : idx;
                    row = streaming ? slot : (size_t)iray;
                    pwr = a.ray_pwr_wt ? a.ray_pwr_wt[iray] : 0.0;
                    slice_n = 0; have_f1 = false; st = SG_CHECK;
                    if (a.resume) {   // a ray suspended by an earlier launch: its last point is still to be checked and saved
                        RayCarry k;
                        resume_ray(a, iray, v, nv, k);
                        s = k.s; sout = k.sout; nstep = k.nstep; flag = k.flag; resid_prev = k.resid_prev; resid_last = k.resid_last;
                        resid_max = k.resid_max; dep_x = k.dep_x; dep_Q = k.dep_Q; rel_err = k.rel_err; abs_err = k.abs_err;
                        first = false;
                        p0 = streaming ? nstep + 1 : 0;
                    } else {
                        nstep = 0; s = 0.0; sout = 0.0; flag = 0; p0 = 0;
                        rel_err = c.rel_err0; abs_err = c.abs_err0;   // ray_init_SG_ode (SG_ode_m.f90:73-85)
                        resid_prev = 0.0; resid_last = 0.0; resid_max = 0.0;
                        initialize_ode_vector<T>(a.rvec0 + 3 * iray, a.rindex_vec0 + 3 * iray, v);
                        if (a.ray_vec) {
                            double *dst = a.ray_vec + row * a.npoints_alloc * nv;
                            if (T::GENERIC) store_point(dst, v, nv); else store_point_fixed<NV>(dst, v);
                        }
                        if (a.residual) a.residual[row * a.npoints_alloc] = 0.0;
                        if (a.start_ray_vec) for (int i = 0; i < nv; ++i) a.start_ray_vec[(size_t)iray * nv + i] = v[i];
                        first = true;
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, st != SG_IDLE) == 0u) {
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;
            continue;
        }
        bool stop = false, did_not_start = false;

        // Slot alignment.  Left alone, every lane is in its own integrator phase and the bookkeeping blocks below run
        // one after the other with a quarter of the lanes each (ncu: 65 % of all warp instructions at 6-9 active
        // lanes).  An internal step is predictor work -> derivative at p -> corrector work -> derivative at yy, so the
        // warp alternates two slot types: in an A slot only lanes whose next block is predictor-side work take part
        // (segment start, after_correct, retry after a failed step), in a B slot only lanes due for after_predict.
        // A lane that is out of phase (after a segment's check_save, after a failure) sits out one slot and is in
        // step with the others from then on; check_save itself needs no bookkeeping and runs in either slot.
        // (Re-synchronising at segment boundaries instead was 2x slower: segments differ in internal step count.)
        bool go = true;
        if (a.sg_align) {
            const bool b_type = st == SG_AFTER_R2;
            go = st == SG_CHECK || (st != SG_IDLE && (slot_b ? b_type : !b_type));
            if (__ballot_sync(0xffffffffu, go) == 0u) { slot_b = !slot_b; continue; }
            slot_b = !slot_b;
        }
        // ---- lane-private integrator bookkeeping up to the next derivative evaluation
        int req = 0;               // 1: derivative at yy (start), 2: at the predicted p, 3: at the corrected yy,
                                   // 4: check_save of the new point v + the start derivative of the next segment there
        bool need_predict = false, de_top = false, crashed = false;
        if (go) {
            if (st == SG_CHECK) req = 4;
            else if (st == SG_DE_BEGIN || st == SG_AFTER_CHECK) {   // ode/de entry with iflag = 1 (ode_RAYS.f90:425-505)
                t = s;
                if (t == sout) { stop = true; flag = RAYS_STOP_SG_T_EQ_TOUT; }
                else if (rel_err < 0.0 || abs_err < 0.0) { stop = true; flag = RAYS_STOP_SG_BAD_TOL; }
                else {
                    eps = fmax(rel_err, abs_err);
                    if (eps <= 0.0) { stop = true; flag = RAYS_STOP_SG_EPS_LE_0; }
                    else {
                        const double del = sout - t;
                        absdel = fabs(del);
                        tend = t + 10.0 * del;
                        nostep = 0; kle4 = 0; stiff = false;
                        releps = rel_err / eps;
                        abseps = abs_err / eps;
                        W.start = true;
                        W.x = t;
                        for (int l = 0; l < nv; ++l) W.yy[l] = v[l];
                        W.h = copysign(fmax(fabs(sout - W.x), fouru * fabs(W.x)), sout - W.x);
                        W.ns = 0; W.k = 0; W.kold = 0; W.hold = 0.0; W.phase1 = false; W.nornd = true;
                        de_top = true;
                    }
                }
            } else if (st == SG_AFTER_R1) {
                sg_after_start<NV>(nv, W, eps);
                need_predict = true;
            } else if (st == SG_AFTER_R2) {
                const int r = sg_after_predict<NV>(nv, W, eps);
                if (r == 0) req = 3;
                else if (r == 1) { if (a.sg_align) st = SG_RETRY; else need_predict = true; }   // failed: predict again with the reduced step
                else crashed = true;
            } else if (st == SG_RETRY) {
                need_predict = true;
            } else if (st == SG_AFTER_R3) {
                sg_after_correct<NV>(nv, W, eps);
                nostep = nostep + 1;       // de: step counter and stiffness test (ode_RAYS.f90:578-590)
                kle4 = kle4 + 1;
                if (4 < W.kold) kle4 = 0;
                if (50 <= kle4) stiff = true;
                de_top = true;
            }
            if (de_top) {   // top of de's loop (ode_RAYS.f90:509-562)
                if (absdel <= fabs(W.x - t)) {           // past the output point: interpolate, segment done (iflag = 2)
                    sg_intrp<NV>(nv, W, sout, v);
                    s = sout;
                    st = SG_CHECK;
                    ++slice_n;
                    if (a.slice_steps > 0 && slice_n >= a.slice_steps) {   // suspend: packed into full warps by the next launch
                        RayCarry k{s, sout, resid_prev, resid_last, resid_max, dep_x, dep_Q, rel_err, abs_err, nstep, flag};
                        suspend_ray(a, iray, v, nv, k);
                        st = SG_IDLE;
                        fin = true; fin_np = nstep + 1 - p0;
                    } else req = 4;
                } else if (maxnum <= nostep) {             // iflag = 4 / 5: error return, the ray stops with y = yy, t = x
                    flag = stiff ? RAYS_STOP_SG_STIFF : RAYS_STOP_SG_MAXNUM;
                    for (int l = 0; l < nv; ++l) v[l] = W.yy[l];
                    s = W.x;
                    stop = true;
                } else {
                    W.h = copysign(fmin(fabs(W.h), fabs(tend - W.x)), W.h);
                    for (int l = 0; l < nv; ++l) W.wt[l] = releps * fabs(W.yy[l]) + abseps;
                    if (sg_block0<NV>(nv, W, eps)) crashed = true;
                    else if (W.start) {
                        if (have_f1) {   // the start derivative was evaluated together with check_save at this very point
                            have_f1 = false;
                            my_rhs += 1;
                            if (f1_code) { flag = f1_code; sout = s; stop = true; }
                            else { sg_after_start<NV>(nv, W, eps); need_predict = true; }
                        } else req = 1;
                    } else need_predict = true;
                }
            }
            if (crashed) {   // iflag = 3: tolerances raised (ode_RAYS.f90:566-575), then SG_ode's test (SG_ode_m.f90:138-149)
                rel_err = eps * releps;
                abs_err = eps * abseps;
                for (int l = 0; l < nv; ++l) v[l] = W.yy[l];
                s = W.x;
                const double total_error = fabs(rel_err) + fabs(abs_err);
                if (total_error > c.SG_error_limit) { flag = RAYS_STOP_ODE_TOTAL_ERROR; stop = true; }
                else st = SG_DE_BEGIN;           // SG_ode loops: ode again from the current s to sout
            }
        }
        // ---- predictor (convergent for every lane that needs it, whatever state it came from)
        if (need_predict && !stop) { sg_predict<NV>(nv, W); req = 2; }
        // ---- equilibrium + derivative evaluation: ONE site for the whole warp.  A lane with req = 4 also runs
        // check_save on the same equilibrium (the reference evaluates it twice at this point: check_save.f90:38
        // and eqn_ray.f90:87 of the next segment's first derivative), saves the point and runs the loop-top tests.
        if (req && !stop) {
            const double *u = req == 2 ? W.p : (req == 4 ? v : W.yy);
            double uu[NV], ff[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) { uu[i] = i < nv ? u[i] : 0.0; ff[i] = 0.0; }
            Eq<NSM> e;
            equilibrium<T::EQ, T::NS, true>(uu[0], uu[1], uu[2], e);
            double dddx[3], dddk[3], dddw = 0.0;
            bool have_derivs = false;
            if (req == 4) {
                double resid = 0.0;
                if (e.err) {
                    flag = e.err;
                    if (T::damp() && v[7] > c.total_damping_limit) { stop = true; flag = RAYS_STOP_TOTAL_ABSORPTION; }
                } else {
                    check_save_resid<T>(e, uu, resid, stop, flag);
                    double dddw_cold;
                    if (T::DERIV == RAYS_DERIV_COLD) {
                        ray_derivs<T>(e, uu, dddx, dddk, dddw);
                        have_derivs = true;
                        dddw_cold = dddw;
                    } else {
                        const Rcp K0 = g_dc.rc_k0;
                        const double nvec[3] = {qdiv(uu[3], K0), qdiv(uu[4], K0), qdiv(uu[5], K0)};
                        double tx[3], tk[3];
                        deriv_cold<T::NS>(e, nvec, tx, tk, dddw_cold);
                    }
                    check_save_tail<T>(uu, dddw_cold, stop, flag);
                }
                if (stop) did_not_start = first;
                else {
                    if (!first) {   // the point passed check_save: save it (ray_tracing.f90:237-243)
                        nstep = nstep + 1;
                        if (a.ray_vec) {
                            double *dst = a.ray_vec + (row * a.npoints_alloc + (nstep - p0)) * nv;
                            if (T::GENERIC) store_point(dst, v, nv); else store_point_fixed<NV>(dst, v);
                        }
                        if (a.residual) a.residual[row * a.npoints_alloc + (nstep - p0)] = resid;
                        resid_prev = resid_last;
                        resid_last = resid;
                        if (fabs(resid_prev) > resid_max) resid_max = fabs(resid_prev);
                        if (binning) {
                            const double xn = dep_abscissa<T::EQ>(v), Qn = v[7] * pwr;
                            bin_segment(dbins, dep_x, xn, dep_Q, Qn);
                            dep_x = xn; dep_Q = Qn;
                        }
                        ++my_steps;
                    } else if (binning) { dep_x = dep_abscissa<T::EQ>(v); dep_Q = v[7] * pwr; }
                    first = false;
                    s = sout;                       // top of the trajectory loop (ray_tracing.f90:118-172)
                    sout = sout + c.ds;
                    if (sout > c.s_max) { stop = true; flag = RAYS_STOP_SOUT_GT_SMAX; }
                    else if (nstep + 1 > c.nstep_max) { stop = true; flag = RAYS_STOP_NSTEP_MAX; }
                }
            }
            if (!stop) {
                int code = e.err;
                if (!code && !have_derivs) code = ray_derivs<T>(e, uu, dddx, dddk, dddw);
                if (!code) code = ray_equations<T>(e, uu, dddx, dddk, dddw, ff);
                if (req == 4) {          // keep it for the start of the next segment (consumed after de's entry tests)
                    have_f1 = true; f1_code = code;
                    if (!code) {
#pragma unroll
                        for (int i = 0; i < NV; ++i) if (i < nv) W.yp[i] = ff[i];
                    }
                    st = SG_AFTER_CHECK;
                } else {
                    my_rhs += 1;
                    if (code) { flag = code; sout = s; stop = true; }     // SG_ode: stop_ode set in eqn_ray -> sout = s
                    else {
#pragma unroll
                        for (int i = 0; i < NV; ++i) if (i < nv) W.yp[i] = ff[i];
                        st = req == 1 ? SG_AFTER_R1 : (req == 2 ? SG_AFTER_R2 : SG_AFTER_R3);
                    }
                }
            }
        }
        if (stop) {
            a.stop_code[iray] = flag;
            if (did_not_start) {   // only npoints, flag and the first point are set (ray_tracing.f90:101-112)
                a.npoints[iray] = 1;
                if (a.initial_ray_power) a.initial_ray_power[iray] = 0.0;
                if (a.end_residuals) a.end_residuals[iray] = 0.0;
                if (a.max_residuals) a.max_residuals[iray] = 0.0;
                if (a.end_ray_parameter) a.end_ray_parameter[iray] = 0.0;
                if (a.start_ray_vec) for (int i = 0; i < nv; ++i) a.start_ray_vec[(size_t)iray * nv + i] = 0.0;
                if (a.end_ray_vec) for (int i = 0; i < nv; ++i) a.end_ray_vec[(size_t)iray * nv + i] = 0.0;
            } else {               // summary block (ray_tracing.f90:252-260)
                a.npoints[iray] = nstep + 1;
                if (a.initial_ray_power) a.initial_ray_power[iray] = pwr;
                if (a.end_residuals) a.end_residuals[iray] = nstep >= 1 ? resid_prev : 0.0;
                if (a.max_residuals) a.max_residuals[iray] = nstep >= 1 ? resid_max : -DBL_MAX;
                if (a.end_ray_parameter) a.end_ray_parameter[iray] = v[6];
                if (a.end_ray_vec) for (int i = 0; i < nv; ++i) a.end_ray_vec[(size_t)iray * nv + i] = v[i];
            }
            st = SG_IDLE;
            fin = true; fin_np = did_not_start ? 1 : nstep + 1 - p0;
        }
        if (streaming) { flush_finished_rays(a, fin, iray, fin_np, p0, row, nv, lane); fin = false; }
    }
    dep_end(a, binning);
    unsigned long long stt = my_steps, rh = my_rhs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { stt += __shfl_down_sync(0xffffffffu, stt, o); rh += __shfl_down_sync(0xffffffffu, rh, o); }
    if (lane == 0) { atomicAdd(a.counters, stt); atomicAdd(a.counters + 1, rh); }
}

// ---- fused RK4 trace kernel ---------------------------------------------------------------------------------
// One loop iteration = one ray-step of every active lane.  check_save of the point a step produced is
// deferred to the top of the next iteration, where the point is also the first RK4 stage: the two share one
// equilibrium evaluation (and, for ray_deriv_name='cold', one deriv_cold), which is bit-identical to the
// reference's two evaluations at the same point.  The four stages run through ONE copy of the equilibrium /
// derivative code (a real loop), which keeps the hot loop inside the instruction cache: the straight-line
// version was fetch-bound (ncu: 4.5 issue slots lost to "no instruction" per issued instruction).
// Sequence per iteration, in the reference's order (ray_tracing.f90:116-245):
//   check_save(v) -> [stop: point not saved] -> save point, nstep++ -> s = sout, sout += ds -> s_max / nstep_max
//   tests -> RK4 stages -> v advanced (or ray stopped by the RHS with v untouched).
// Resident CTAs per SM (sets the register budget: 4 -> 128, 3 -> 168 registers per thread), 1M-ray bench fan, deriv_num:
//   looped determinants:   2 -> 284 ms, 3 -> 238, 4 -> 234, 5 -> 246, 6 -> 269
//   unrolled determinants (RAYS_DN_UK = 6): 4 -> 154.6 ms, 3 -> 147.3 ms: with more independent FP64 work per warp the
//   kernel prefers fewer warps that do not spill (stack 232 -> 48 bytes); 2 -> 176.5 ms
// deriv_cold on the same fan: 4 -> 42.7 ms, 3 -> 39.3, 2 -> 37.7 (3.9e9 ray-steps/s; 352 bytes of spills at 128 registers, none at 255)
// deriv_cold + damping (config 5: damp_fund_ECH + fused binning): 42 KB of the kernel's 99 KB are hot, more than the SM's 32 KB
// instruction cache, and with warps drifting apart every warp fetched every line from L2 for itself (hit rate 63 %, 3.0 of 6.5 stall
// cycles per issue on instruction fetch).  Rk4Sync: ONE CTA per SM whose warps start every ray-step together (__syncthreads_or at
// the loop top), so a line fetched by one warp serves all of them: 185.7 -> 144.8 ms per 8.39 M-ray fan with 256 threads (hit rate
// 88 %); what was left was `wait` (dependency chains at 2 warps per scheduler), and without the fetch stalls a third warp per
// scheduler pays for its spills: 384 threads x 168 registers 124.2 ms (320: 150.9, 448: 140.4 - uneven over the four schedulers -
// 512 x 128 registers: 124.4).
// Measured on the other families, where the hot loops fit the cache, the barrier only costs: mirror 251 -> 269 ms, headline
// 147.4 -> 149.9 (3 x 128) / 150.8 (1 x 384), deriv_cold 37.8 -> 39.3 (2 x 128) / 42.7 (1 x 256); 256 threads without the barrier
// gain 2 % (181.6 ms).
#ifndef RAYS_RK4_SYNC_ALL
#define RAYS_RK4_SYNC_ALL 0   // measurement aid: 1 = the barrier in every RK4 kernel of the translation unit
#endif
template <class T> struct Rk4Sync { static constexpr bool value = RAYS_RK4_SYNC_ALL || (!T::GENERIC && T::DAMP == 1 && T::DERIV == RAYS_DERIV_COLD); };
#ifdef RAYS_RK4_MIN_CTAS
template <class T> struct Rk4Ctas { static constexpr int value = RAYS_RK4_MIN_CTAS; };
#else
template <class T> struct Rk4Ctas { static constexpr int value = Rk4Sync<T>::value ? 1 : (T::GENERIC ? 3 : (T::DERIV == RAYS_DERIV_NUM ? 3 : 2)); };
#endif
#ifdef RAYS_RK4_BLOCK
template <class T> struct Rk4Block { static constexpr int value = RAYS_RK4_BLOCK; };
#else
template <class T> struct Rk4Block { static constexpr int value = Rk4Sync<T>::value ? 3 * kTraceBlock : kTraceBlock; };
#endif
template <class T>
__global__ void __launch_bounds__(Rk4Block<T>::value, Rk4Ctas<T>::value) trace_rk4_kernel(const TraceArgs a) {
    constexpr int NV = T::NV;
    constexpr int NSM = NSpec<T::NS>::MAX;
    const int nv = T::nv();
    const rays_cfg &c = g_dc.c;
    const unsigned lane = threadIdx.x & 31;
    double v[NV];
    double s = 0.0, sout = 0.0;
    double resid_prev = 0.0, resid_last = 0.0, resid_max = 0.0;
    double dep_x = 0.0, dep_Q = 0.0, pwr = 0.0;
    long long iray = -1;
    int nstep = 0, flag = 0;
    bool active = false, exhausted = false, first = false;
    unsigned long long my_steps = 0;
    unsigned my_rhs = 0;
    const bool binning = a.dep_acc != nullptr && T::damp();
    const DepBins dbins = dep_begin(a, binning);
    const bool streaming = a.host_ray_vec != nullptr || a.host_residual != nullptr;
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t row = 0;
    bool fin = false;      // this lane's ray ended (or was suspended) in the current iteration
    int fin_np = 0;
    int p0 = 0;          // index (within the ray) of the first point staged in this lane's row (streaming + resume)
    int slice_n = 0;     // steps this ray has taken in this launch
    if (a.trace_started && blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile int *>(a.trace_started) = 1;
    bool warp_done = false;   // Rk4Sync: this warp has no ray left and only keeps the CTA's barrier company
    for (;;) {
        if (Rk4Sync<T>::value) {   // the warps of the CTA start every ray-step together (instruction cache, see Rk4Sync)
            if (!__syncthreads_or(!warp_done)) break;
            if (warp_done) continue;
        }
        // ---- refill from the work queue (one atomic per warp)
        const unsigned want = __ballot_sync(0xffffffffu, !active && !exhausted);
        if (RAYS_RARE(want)) {
            unsigned long long base = 0;
            const int leader = __ffs(want) - 1;
            if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!active && !exhausted) {
                const long long idx = (long long)(base + __popc(want & ((1u << lane) - 1u)));
                if (idx >= a.nray) exhausted = true;
                else {
                    iray = a.order ? (long long)a.order[idx] : idx;
                    row = streaming ? slot : (size_t)iray;
                    pwr = a.ray_pwr_wt ? a.ray_pwr_wt[iray] : 0.0;
                    slice_n = 0; active = true;
                    if (a.resume) {   // a ray suspended by an earlier launch: its last point is still to be checked and saved
                        RayCarry k;
                        resume_ray(a, iray, v, nv, k);
                        s = k.s; sout = k.sout; nstep = k.nstep; flag = k.flag; resid_prev = k.resid_prev; resid_last = k.resid_last;
                        resid_max = k.resid_max; dep_x = k.dep_x; dep_Q = k.dep_Q;
                        first = false;
                        p0 = streaming ? nstep + 1 : 0;
                    } else {
                        nstep = 0; s = 0.0; sout = 0.0; flag = 0; p0 = 0;
                        resid_prev = 0.0; resid_last = 0.0; resid_max = 0.0;
                        initialize_ode_vector<T>(a.rvec0 + 3 * iray, a.rindex_vec0 + 3 * iray, v);
                        if (a.ray_vec) {
                            double *dst = a.ray_vec + row * a.npoints_alloc * nv;
                            if (T::GENERIC) store_point(dst, v, nv); else store_point_fixed<NV>(dst, v);
                        }
                        if (a.residual) a.residual[row * a.npoints_alloc] = 0.0;
                        if (a.start_ray_vec) for (int i = 0; i < nv; ++i) a.start_ray_vec[(size_t)iray * nv + i] = v[i];
                        first = true;
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) {
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) {
                if (!Rk4Sync<T>::value) break;
                warp_done = true;
            }
            continue;
        }
        if (active) {
            const double ds = c.ds;    // RK4_ode: ds = sout - s with sout = s + ds formed below; see note (*)
            double w[NV], acc[NV], f[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) { w[i] = v[i]; acc[i] = 0.0; f[i] = 0.0; }
            int code = 0;          // RHS stop code (ray ends, v untouched)
            bool stop = false;     // stop raised by check_save or the loop-top tests
            bool did_not_start = false;
            double h = 0.0;        // sout - s of this step
#pragma unroll 1
            for (int stage = 0; stage < 4; ++stage) {
                Eq<NSM> e;
                equilibrium<T::EQ, T::NS, true>(w[0], w[1], w[2], e);
                double dddx[3], dddk[3], dddw = 0.0;
                bool have_derivs = false;
                if (stage == 0) {
                    // ---- check_save(v) (check_save.f90:1-161)
                    double resid = 0.0;
                    if (e.err) {
                        flag = e.err;
                        if (T::damp() && v[7] > c.total_damping_limit) { stop = true; flag = RAYS_STOP_TOTAL_ABSORPTION; }
                    } else {
                        check_save_resid<T>(e, v, resid, stop, flag);
                        double dddw_cold;
                        if (T::DERIV == RAYS_DERIV_COLD) {
                            ray_derivs<T>(e, v, dddx, dddk, dddw);     // shared with the first stage below
                            have_derivs = true;
                            dddw_cold = dddw;
                        } else {
                            const Rcp K0 = g_dc.rc_k0;
                            const double nvec[3] = {qdiv(v[3], K0), qdiv(v[4], K0), qdiv(v[5], K0)};
                            double tx[3], tk[3];
                            deriv_cold<T::NS>(e, nvec, tx, tk, dddw_cold);
                        }
                        check_save_tail<T>(v, dddw_cold, stop, flag);
                    }
                    if (stop) { did_not_start = first; break; }
                    if (!first) {   // the point passed check_save: save it (ray_tracing.f90:237-243)
                        nstep = nstep + 1;
                        if (a.ray_vec) {
                            double *dst = a.ray_vec + (row * a.npoints_alloc + (nstep - p0)) * nv;
                            if (T::GENERIC) store_point(dst, v, nv); else store_point_fixed<NV>(dst, v);
                        }
                        if (a.residual) a.residual[row * a.npoints_alloc + (nstep - p0)] = resid;
                        resid_prev = resid_last;
                        resid_last = resid;
                        if (fabs(resid_prev) > resid_max) resid_max = fabs(resid_prev);
                        if (binning) {
                            const double xn = dep_abscissa<T::EQ>(v), Qn = v[7] * pwr;
                            bin_segment(dbins, dep_x, xn, dep_Q, Qn);
                            dep_x = xn; dep_Q = Qn;
                        }
                        ++my_steps;
                    } else if (binning) { dep_x = dep_abscissa<T::EQ>(v); dep_Q = v[7] * pwr; }
                    first = false;
                    // ---- top of the trajectory loop (ray_tracing.f90:118-172)
                    s = sout;
                    sout = sout + ds;
                    if (sout > c.s_max) { stop = true; flag = RAYS_STOP_SOUT_GT_SMAX; break; }
                    if (nstep + 1 > c.nstep_max) { stop = true; flag = RAYS_STOP_NSTEP_MAX; break; }
                    h = sout - s;   // RK4_ode_m.f90:74
                }
                // ---- eqn_ray at w (eqn_ray.f90:87-229)
                if (e.err) { code = e.err; break; }
                if (!have_derivs) {
                    code = ray_derivs<T>(e, w, dddx, dddk, dddw);
                    if (code) break;
                }
                code = ray_equations<T>(e, w, dddx, dddk, dddw, f);
                if (code) break;
                const double ca = (stage == 1 || stage == 2) ? 2.0 : 1.0;
                const double cw = stage == 2 ? 1.0 : 0.5;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (i < nv) {
                        acc[i] = stage == 0 ? f[i] : acc[i] + ca * f[i];
                        w[i] = v[i] + h * f[i] * cw;
                    }
                my_rhs += 1;
            }
            if (RAYS_USUAL(!stop && code == 0)) {
#pragma unroll
                for (int i = 0; i < NV; ++i) if (i < nv) v[i] = v[i] + qdiv(h * acc[i], g_dc.rc_six);
                s = sout;
                ++slice_n;
                if (RAYS_RARE(a.slice_steps > 0 && slice_n >= a.slice_steps)) {   // suspend: packed into full warps by the next launch
                    RayCarry k{s, sout, resid_prev, resid_last, resid_max, dep_x, dep_Q, 0.0, 0.0, nstep, flag};
                    suspend_ray(a, iray, v, nv, k);
                    active = false;
                    fin = true; fin_np = nstep + 1 - p0;
                }
            } else {
                if (code) flag = code;
                a.stop_code[iray] = flag;
                if (did_not_start) {   // only npoints, flag and the first point are set (ray_tracing.f90:101-112)
                    a.npoints[iray] = 1;
                    if (a.initial_ray_power) a.initial_ray_power[iray] = 0.0;
                    if (a.end_residuals) a.end_residuals[iray] = 0.0;
                    if (a.max_residuals) a.max_residuals[iray] = 0.0;
                    if (a.end_ray_parameter) a.end_ray_parameter[iray] = 0.0;
                    if (a.start_ray_vec) for (int i = 0; i < nv; ++i) a.start_ray_vec[(size_t)iray * nv + i] = 0.0;
                    if (a.end_ray_vec) for (int i = 0; i < nv; ++i) a.end_ray_vec[(size_t)iray * nv + i] = 0.0;
                } else {               // summary block (ray_tracing.f90:252-260)
                    a.npoints[iray] = nstep + 1;
                    if (a.initial_ray_power) a.initial_ray_power[iray] = pwr;
                    if (a.end_residuals) a.end_residuals[iray] = nstep >= 1 ? resid_prev : 0.0;
                    if (a.max_residuals) a.max_residuals[iray] = nstep >= 1 ? resid_max : -DBL_MAX;
                    if (a.end_ray_parameter) a.end_ray_parameter[iray] = v[6];
                    if (a.end_ray_vec) for (int i = 0; i < nv; ++i) a.end_ray_vec[(size_t)iray * nv + i] = v[i];
                }
                active = false;
                fin = true; fin_np = did_not_start ? 1 : nstep + 1 - p0;
                if (a.done_list) {   // hand the finished ray to the copier kernel: everything written above must be visible first
                    __threadfence();
                    const unsigned long long pos = atomicAdd(a.done_count, 1ULL);
                    reinterpret_cast<volatile int *>(a.done_list)[pos] = (int)iray;
                }
            }
        }
        if (streaming) { flush_finished_rays(a, fin, iray, fin_np, p0, row, nv, lane); fin = false; }
    }
    dep_end(a, binning);
    unsigned long long st = my_steps, rh = my_rhs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { st += __shfl_down_sync(0xffffffffu, st, o); rh += __shfl_down_sync(0xffffffffu, rh, o); }
    if (lane == 0) { atomicAdd(a.counters, st); atomicAdd(a.counters + 1, rh); }
}

// ---- arithmetic self-test: rcp_rn / sqrt_rn / qdiv against IEEE 1.0/d, sqrt(x), x/d on pseudo-random operands
static __global__ void selftest_arith_kernel(unsigned long long seed, int per_thread, unsigned long long *mismatch) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad_rcp = 0, bad_sqrt = 0, bad_div = 0;
    for (int it = 0; it < per_thread; ++it) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        unsigned long long ma = s;
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        unsigned long long mb = s;
        const int mode = it & 7;
        int ea = (int)((ma >> 52) % 400) - 200, eb = (int)((mb >> 52) % 400) - 200;
        if (mode == 0) { mb |= 0xFFF0000000000ULL; ma |= 0xFF00000000000ULL; ea = eb = 0; }
        else if (mode == 1) { mb = 0xFFFFFFFFFFFFFULL - (mb & 0xFFFFF); ma = 0xFFFFFFFFFFFFFULL - (ma & 0xFFFFFF); ea = eb = 0; }
        else if (mode == 2) { mb &= 0xFFFULL; }
        const double a = __longlong_as_double((long long)(((unsigned long long)(1023 + ea) << 52) | (ma & 0xFFFFFFFFFFFFFULL)));
        double b = __longlong_as_double((long long)(((unsigned long long)(1023 + eb) << 52) | (mb & 0xFFFFFFFFFFFFFULL)));
        if (mb & (1ULL << 60)) b = -b;
        const double r = rcp_rn(b);
        // Known exception of the final correction step (Markstein): a divisor whose 52-bit significand is
        // all ones (probability 2^-52 for real data; CUDA's __drcp_rn routes that case to its slow path).
        // Such operands are counted separately (mismatch[3]) and excluded from the three exact counts.
        if ((mb & 0xFFFFFFFFFFFFFULL) == 0xFFFFFFFFFFFFFULL) { if (r != 1.0 / b) atomicAdd(mismatch + 3, 1ULL); continue; }
        if (r != 1.0 / b) ++bad_rcp;
        if (sqrt_rn(a) != sqrt(a)) ++bad_sqrt;
        Rcp rc; rc.d = b; rc.r = r;
        if (qdiv(a, rc) != a / b) ++bad_div;
    }
    if (bad_rcp) atomicAdd(mismatch, bad_rcp);
    if (bad_sqrt) atomicAdd(mismatch + 1, bad_sqrt);
    if (bad_div) atomicAdd(mismatch + 2, bad_div);
}

// ---- one-point probes (unit parity tests through the C ABI) -----------------------------------------------
template <class T> __global__ void probe_equilibrium_kernel(long long n, const double *rvec, double *out, int *err) {
    constexpr int NSM = NSpec<T::NS>::MAX;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Eq<NSM> e;
    equilibrium<T::EQ, T::NS, true>(rvec[3 * i], rvec[3 * i + 1], rvec[3 * i + 2], e);
    double *o = out + (size_t)i * RAYS_EQ_OUT;
    for (int k = 0; k < RAYS_EQ_OUT; ++k) o[k] = 0.0;
    err[i] = e.err;
    if (e.err) return;
    int k = 0;
    for (int q = 0; q < 3; ++q) o[k++] = e.bvec[q];
    for (int jj = 0; jj < 3; ++jj) for (int ii = 0; ii < 3; ++ii) o[k++] = e.g[ii][jj];
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.ns[s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) for (int ii = 0; ii < 3; ++ii) o[k++] = s < NSM ? e.gradns[ii][s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.ts[s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) for (int ii = 0; ii < 3; ++ii) o[k++] = s == 0 ? e.gradts0[ii] : 0.0;  // electrons only
    o[k++] = e.bmag;
    for (int q = 0; q < 3; ++q) o[k++] = e.gradbmag[q];
    for (int q = 0; q < 3; ++q) o[k++] = e.bunit[q];
    for (int jj = 0; jj < 3; ++jj) for (int ii = 0; ii < 3; ++ii) o[k++] = e.gradbunit[ii][jj];
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.omgc[s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.omgp2[s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.alpha[s < NSM ? s : 0] : 0.0;
    for (int s = 0; s < RAYS_NSPECIES; ++s) o[k++] = s < NSM ? e.gamma[s < NSM ? s : 0] : 0.0;
}
template <class T> __global__ void probe_rhs_kernel(long long n, const double *v, double *dvds, int *stop) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nv = T::nv();
    double w[T::NV], f[T::NV];
    for (int k = 0; k < T::NV; ++k) { w[k] = k < nv ? v[(size_t)i * nv + k] : 0.0; f[k] = 0.0; }
    const int code = eqn_ray<T>(w, f);
    for (int k = 0; k < nv; ++k) dvds[(size_t)i * nv + k] = code ? 0.0 : f[k];
    stop[i] = code;
}
template <class T> __global__ void probe_check_save_kernel(long long n, const double *v, double *resid, int *stop) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nv = T::nv();
    double w[T::NV];
    for (int k = 0; k < T::NV; ++k) w[k] = k < nv ? v[(size_t)i * nv + k] : 0.0;
    double r = 0.0;
    bool st = false;
    int flag = 0;
    check_save<T>(w, r, st, flag);
    resid[i] = r;
    stop[i] = st ? flag : 0;
}

}  // namespace rays_dev
