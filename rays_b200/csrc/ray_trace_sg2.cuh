// ray_trace_sg2.cuh — Shampine-Gordon trace kernel as a per-CTA SLOT MACHINE with the integrator state in SHARED MEMORY.
//
// SG_ode (SG_ode_m.f90:89-159) -> ode/de (ode_RAYS.f90:1-593) -> step (:595-1234) / intrp (:1235-1362), every decision of
// the reference replayed with the reference's arithmetic, as in the per-lane state machine of ray_trace.cuh — what changes
// is WHO executes WHAT.  Measured shape of the work (oracle statistics on the 1M-ray Solov'ev fan, tol 1e-6): a ds segment
// restarts the integrator at order 1 and takes 18.7 attempted internal steps, 11.6 of which FAIL and 7.1
// are accepted, at order k <= 3 (k = 4: 0.05 %): 27 right-hand sides per ray-step, and between two of them each ray needs
// one of four different bookkeeping blocks.  With one ray pinned to one lane the blocks of a warp run one after another with
// a third of the lanes each (ncu on the round-1 kernel: 13 of 32 lanes active, 129 KB of divergent code thrashing the
// instruction cache).
//
// Here a CTA owns S ray SLOTS (more than it has threads).  A slot's live integrator state — yy, wt, the divided differences
// phi(1..5), the step coefficients up to order 3, x, h, ... : SgHot, 87 doubles for nv = 7 — lives in SHARED memory, one
// contiguous record per slot (odd length: a warp's accesses to one field of 32 consecutive slots are conflict-free); what is
// touched once per ray-step (s, sout, residual history, deposition carry) and the rows only orders k > 3 reach live in a
// global, L2-resident record (SgLayout).  The first slot machine kept everything in global memory: ncu showed 1.9 of its
// 10 stall cycles per issue on those loads, 52 GB of DRAM writes for 2.4 GB of trajectory (the 70 MB hot set thrashing L2)
// and 2.7 on its three CTA barriers per iteration.
//
// Scheduling: every slot sits in exactly one shared-memory RING per kind of macro-step it needs next
//   PRED   predictor + f(p) + error test            (order <= 3: register-resident bookkeeping, sg3_*)
//   CORR   f(yy) + history update, next order/step
//   CHECK  segment boundary: interpolate + check_save + start derivative of the next segment (fused: same point)
//   START  restart after a tolerance raise / start derivative on its own (rare)
//   GPRED / GCORR   the same macro-steps for a slot at order k > 3 or with propagated-roundoff control on: the general,
//          looped code (sg2_*, out of line).  A few steps in ten thousand — but ONE such slot in an iteration made its warp
//          three times slower and the whole CTA wait at the barrier, so they are batched as kinds of their own.
//   FIN    ray ended: streaming copy-out of its row by a warp;  IDLE: refill from the global work queue
// and an iteration of the CTA is: ONE barrier; every thread reads the ring counters and comes to the same choice (the kind
// most slots wait for); thread j pops the j-th entry; one macro-step = bookkeeping -> ONE right-hand side -> bookkeeping;
// the thread pushes the slot onto the ring of its next kind (warp-aggregated shared-memory atomics).  All warps of the CTA
// therefore run the same code at the same time (one instruction stream per SM for the instruction cache), lanes are full
// whenever 32 slots of a kind exist, and the once-per-segment work (1/27 of the macro-steps) accumulates until it is the
// largest group instead of running with one or two lanes.
#pragma once
#include <climits>
#include "ray_trace.cuh"

namespace rays_dev {

enum SgKind : int { K_IDLE = 0, K_PRED, K_CORR, K_CHECK, K_START, K_BEGIN, K_FIN };
enum SgRing : int { Q_PRED = 0, Q_CORR, Q_CHECK, Q_START, Q_GPRED, Q_GCORR, Q_FIN, Q_IDLE, kSgNQ };
enum SgBits : int { B_FIRST = 1, B_START = 4, B_PHASE1 = 8, B_NORND = 16, B_STIFF = 32, B_INTRP = 64 };
constexpr int kSgKM = 3;            // highest order of the register-resident fast path
constexpr int kSgHotRows = kSgKM + 2;   // phi(1..k+2)
constexpr int kSgRingCap = 512;     // entries per ring (a power of two >= slots per CTA)
constexpr int kSgStride = kSgRingCap;   // slot stride of the global records

// global (cold) record of one CTA, field-major: double field f of slot j at D[f * kSgStride + j].  The fields the hot record
// holds (SgHot) keep their place here but are never touched.
template <int NV> struct SgLayout {
    enum : int {
        V = 0, YY = V + NV, WT = YY + NV, P = WT + NV, YP = P + NV, PHI = YP + NV,      // phi(l,i) at PHI + i*NV + l, i = 0..16; P, YP: general path only
        ALPHA = PHI + 17 * NV, BETA = ALPHA + 13, SIG = BETA + 13, VV = SIG + 14, WW = VV + 13, G = WW + 14, PSI = G + 14,
        S_ = PSI + 13, SOUT, REL, ABS, RPREV, RLAST, RMAX, DEPX, DEPQ, PWR, IRAY, NDBL
    };
    enum : int { NSTEP = 0, FLAG, P0, SLICE, FINNP, NINT = 6 };
};
// shared-memory (hot) record of one slot: contiguous, NDBL odd
template <int NV> struct SgHot {
    enum : int {
        YY = 0, WT = NV, PHI = 2 * NV,                     // phi(i, l), i = 1..kSgHotRows, at PHI + (i-1)*NV + l
        PSI = PHI + kSgHotRows * NV,                       // psi(1..3), alpha(1..3), beta(1..3), sig(1..4), vv(1..3), g(1..4): index i at base + i - 1
        ALPHA = PSI + kSgKM, BETA = ALPHA + kSgKM, SIG = BETA + kSgKM, VV = SIG + kSgKM + 1, G = VV + kSgKM,
        X = G + kSgKM + 1, H, HOLD, EPS, XOLD, ERK, ERKM1, T0, ABSDEL, TEND, RELEPS, ABSEPS, ROUND, ABSH,
        INTS,                                              // 8 ints
        NDBL = (INTS + 4) | 1
    };
    enum : int { I_NS = 0, I_K, I_KOLD, I_IFAIL, I_KNEW, I_NOSTEP, I_KLE4, I_BITS };
};
// component loops are unrolled (independent loads in flight) for the specialised kernels, real loops for the generic one
#ifndef RAYS_SG_UNROLL
#define RAYS_SG_UNROLL 1
#endif
template <int NV> struct SgUnroll { static constexpr int L = (RAYS_SG_UNROLL && NV <= 13) ? NV : 1; };

// accessors of one slot.  The ray vector v shares the storage of yy: intrp forms v from yy when a segment ends and the next
// segment starts the integrator afresh from v (iflag = 1, ode_RAYS.f90:425-505), so the two are never live at the same time.
template <int NV> struct SgSlot {
    using L = SgLayout<NV>;
    using HL = SgHot<NV>;
    double *h;   // shared: this slot's hot record
    double *d;   // global: &D[slot]
    int *n;      // global: &I[slot]
    RD_INLINE double &hs(int field) const { return h[field]; }
    RD_INLINE int &hi(int idx) const { return reinterpret_cast<int *>(h + HL::INTS)[idx]; }
    RD_INLINE double &f(int field) const { return d[(size_t)field * kSgStride]; }
    RD_INLINE int &i(int field) const { return n[(size_t)field * kSgStride]; }
    RD_INLINE double &v(int l) const { return h[HL::YY + l]; }
    RD_INLINE double &yy(int l) const { return h[HL::YY + l]; }
    RD_INLINE double &wt(int l) const { return h[HL::WT + l]; }
    RD_INLINE double &p(int l) const { return f(L::P + l); }
    RD_INLINE double &yp(int l) const { return f(L::YP + l); }
    RD_INLINE double &phi(int i_, int l) const { return i_ <= kSgHotRows ? h[HL::PHI + (i_ - 1) * NV + l] : f(L::PHI + i_ * NV + l); }
    RD_INLINE double &alpha(int i_) const { return i_ <= kSgKM ? h[HL::ALPHA + i_ - 1] : f(L::ALPHA + i_); }
    RD_INLINE double &beta(int i_) const { return i_ <= kSgKM ? h[HL::BETA + i_ - 1] : f(L::BETA + i_); }
    RD_INLINE double &sig(int i_) const { return i_ <= kSgKM + 1 ? h[HL::SIG + i_ - 1] : f(L::SIG + i_); }
    RD_INLINE double &vv(int i_) const { return i_ <= kSgKM ? h[HL::VV + i_ - 1] : f(L::VV + i_); }
    RD_INLINE double &ww(int i_) const { return f(L::WW + i_); }
    RD_INLINE double &g(int i_) const { return i_ <= kSgKM + 1 ? h[HL::G + i_ - 1] : f(L::G + i_); }
    RD_INLINE double &psi(int i_) const { return i_ <= kSgKM ? h[HL::PSI + i_ - 1] : f(L::PSI + i_); }
};

// ---- the pieces of `step` on slot memory: the statements of sg_block0 ... sg_intrp (ray_trace.cuh), with the loops over the
// history index i outside and the loops over the components inside and unrolled (per component the operation order is the
// reference's).  The predicted solution p and the derivative yp live in registers: both are produced and consumed within
// one macro-step.
// step, first block (ode_RAYS.f90:840-852); returns true on crash
template <int NV> RD_INLINE bool sg2_block0(int neqn, const SgSlot<NV> &W, double &eps) {
    using HL = SgHot<NV>;
    const double twou = 2.0 * DBL_EPSILON, fouru = 2.0 * twou;
    const double h = W.hs(HL::H), x = W.hs(HL::X);
    if (fabs(h) < fouru * fabs(x)) { W.hs(HL::H) = copysign(fouru * fabs(x), h); return true; }
    const double p5eps = 0.5 * eps;
    double sum = 0.0;
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) { const double q = sg_div(W.yy(l), W.wt(l)); sum = sum + q * q; }
    const double round = twou * sg_sqrt(sum);
    W.hs(HL::ROUND) = round;
    if (p5eps < round) { eps = 2.0 * round * (1.0 + fouru); return true; }
    W.g(1) = 1.0; W.g(2) = 0.5; W.sig(1) = 1.0;
    W.hi(HL::I_IFAIL) = 0;
    return false;
}
// step, start block after the first derivative evaluation (:858-885)
template <int NV> RD_INLINE void sg2_after_start(int neqn, const SgSlot<NV> &W, double eps, int &bits, const double (&yp)[NV]) {
    using HL = SgHot<NV>;
    const double fouru = 4.0 * DBL_EPSILON;
    const double p5eps = 0.5 * eps;
    double tot = 0.0;
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) {
        W.phi(1, l) = yp[l]; W.phi(2, l) = 0.0;
        const double q = sg_div(yp[l], W.wt(l)); tot = tot + q * q;
    }
    const double total = sg_sqrt(tot);
    const double h = W.hs(HL::H);
    double absh = fabs(h);
    if (eps < 16.0 * total * h * h) absh = 0.25 * sg_sqrt(sg_div(eps, total));
    W.hs(HL::H) = copysign(fmax(absh, fouru * fabs(W.hs(HL::X))), h);
    W.hs(HL::HOLD) = 0.0;
    W.hi(HL::I_K) = 1; W.hi(HL::I_KOLD) = 0;
    bits = (bits & ~B_START) | B_PHASE1 | B_NORND;
    if (p5eps <= 100.0 * W.hs(HL::ROUND)) {
        bits &= ~B_NORND;
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) W.phi(15, l) = 0.0;
    }
}
// step, blocks 1 and 2 (:896-1015): coefficients for this step size/order, then the predicted solution p at x + h
template <int NV> RD_INLINE void sg2_predict(int neqn, const SgSlot<NV> &W, int bits, double (&p)[NV]) {
    using HL = SgHot<NV>;
    const int k = W.hi(HL::I_K), kold = W.hi(HL::I_KOLD);
    int ns = W.hi(HL::I_NS);
    const double h = W.hs(HL::H);
    const int kp1 = k + 1, kp2 = k + 2;
    if (h != W.hs(HL::HOLD)) ns = 0;
    if (ns <= kold) ns = ns + 1;
    const int nsp1 = ns + 1;
    if (ns <= k) {
        W.beta(ns) = 1.0;
        W.alpha(ns) = kSGinv[ns];
        double temp1 = h * (double)ns;
        W.sig(nsp1) = 1.0;
        double beta_prev = 1.0, sig_prev = 1.0;
        #pragma unroll 1
        for (int i = nsp1; i <= k; ++i) {
            const double temp2 = W.psi(i - 1);
            W.psi(i - 1) = temp1;
            beta_prev = sg_div(beta_prev * temp1, temp2);
            W.beta(i) = beta_prev;
            temp1 = temp2 + h;
            const double al = sg_div(h, temp1);
            W.alpha(i) = al;
            sig_prev = (double)i * al * sig_prev;
            W.sig(i + 1) = sig_prev;
        }
        W.psi(k) = temp1;
        if (ns <= 1) {
            #pragma unroll 1
            for (int iq = 1; iq <= k; ++iq) { W.vv(iq) = kSGinvTri[iq]; W.ww(iq) = kSGinvTri[iq]; }
        } else {
            if (kold < k) {
                W.vv(k) = kSGinvTri[k];
                #pragma unroll 1
                for (int j = 1; j <= ns - 2; ++j) { const int i = k - j; W.vv(i) = W.vv(i) - W.alpha(j + 1) * W.vv(i + 1); }
            }
            const double al = W.alpha(ns);
            #pragma unroll 1
            for (int iq = 1; iq <= kp1 - ns; ++iq) { const double t = W.vv(iq) - al * W.vv(iq + 1); W.vv(iq) = t; W.ww(iq) = t; }
            W.g(nsp1) = W.ww(1);
        }
        #pragma unroll 1
        for (int i = ns + 2; i <= kp1; ++i) {
            const double al = W.alpha(i - 1);
            #pragma unroll 1
            for (int iq = 1; iq <= kp2 - i; ++iq) W.ww(iq) = W.ww(iq) - al * W.ww(iq + 1);
            W.g(i) = W.ww(1);
        }
    }
    W.hi(HL::I_NS) = ns;
    #pragma unroll 1
    for (int i = nsp1; i <= k; ++i) {
        const double b = W.beta(i);
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) W.phi(i, l) = b * W.phi(i, l);
    }
    double up[NV];
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) {
        W.phi(kp2, l) = W.phi(kp1, l);
        W.phi(kp1, l) = 0.0;
        p[l] = 0.0; up[l] = 0.0;
    }
    #pragma unroll 1
    for (int i = k; i >= 1; --i) {
        const double gi = W.g(i);
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) {
            const double f = W.phi(i, l);
            p[l] = p[l] + f * gi;
            up[l] = f + up[l];
            W.phi(i, l) = up[l];
        }
    }
    const bool nornd = (bits & B_NORND) != 0;
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) {
        const double y = W.yy(l);
        if (!nornd) {
            const double tau = h * p[l] - W.phi(15, l);
            const double pp = y + tau;
            p[l] = pp;
            W.phi(16, l) = (pp - y) - tau;
        } else p[l] = y + h * p[l];
    }
    const double x = W.hs(HL::X);
    W.hs(HL::XOLD) = x;
    W.hs(HL::X) = x + h;
    W.hs(HL::ABSH) = fabs(h);
}
// step, after the derivative at the predicted point (:1022-1110): error estimates, accept or fail.
// returns 0 = accepted (corrected solution formed in yy), 1 = failed: retry with the reduced step, 2 = crash (eps doubled)
template <int NV> RD_INLINE int sg2_after_predict(int neqn, const SgSlot<NV> &W, double &eps, int &bits, const double (&p)[NV], const double (&yp)[NV]) {
    using HL = SgHot<NV>;
    const double fouru = 4.0 * DBL_EPSILON;
    const int k = W.hi(HL::I_K), kp1 = k + 1, km1 = k - 1, km2 = k - 2;
    const double absh = W.hs(HL::ABSH), p5eps = 0.5 * eps;
    double erkm2 = 0.0, erkm1 = 0.0, erk = 0.0;
    double dl[NV];   // yp - phi(1)
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) {
        const Rcp wl = rcp_of(W.wt(l));   // up to three quotients by the same weight
        const double ph1 = W.phi(1, l);
        if (0 < km2) { const double q = sg_quot(W.phi(km1, l) + yp[l] - ph1, wl); erkm2 = erkm2 + q * q; }
        if (0 <= km2) { const double q = sg_quot(W.phi(k, l) + yp[l] - ph1, wl); erkm1 = erkm1 + q * q; }
        dl[l] = yp[l] - ph1;
        const double q = sg_quot(dl[l], wl);
        erk = erk + q * q;
    }
    if (0 < km2) erkm2 = absh * W.sig(km1) * kSGgstr[km2] * sg_sqrt(erkm2);
    if (0 <= km2) erkm1 = absh * W.sig(k) * kSGgstr[km1] * sg_sqrt(erkm1);
    const double rt_erk = sg_sqrt(erk);
    const double gkp1 = W.g(kp1);
    const double err = absh * rt_erk * (W.g(k) - gkp1);
    erk = absh * rt_erk * W.sig(kp1) * kSGgstr[k];
    int knew = k;
    if (0 < km2) {
        if (fmax(erkm1, erkm2) <= erk) knew = km1;
    } else if (0 == km2) {
        if (erkm1 <= 0.5 * erk) knew = km1;
    }
    W.hi(HL::I_KNEW) = knew; W.hs(HL::ERK) = erk; W.hs(HL::ERKM1) = erkm1;
    const double h = W.hs(HL::H);
    if (err <= eps) {   // accepted: correct (:1123-1141)
        W.hi(HL::I_KOLD) = k;
        W.hs(HL::HOLD) = h;
        const bool nornd = (bits & B_NORND) != 0;
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) {
            if (!nornd) {
                const double rho = h * gkp1 * dl[l] - W.phi(16, l);
                const double y = p[l] + rho;
                W.yy(l) = y;
                W.phi(15, l) = (y - p[l]) - rho;
            } else W.yy(l) = p[l] + h * gkp1 * dl[l];
        }
        return 0;
    }
    // step failed (:1076-1110): restore x, phi, psi; halve (or more) the step
    bits &= ~B_PHASE1;
    const double x = W.hs(HL::XOLD);
    W.hs(HL::X) = x;
    double nxt[NV];   // phi(i+1) as it was: row i+1 is restored after row i reads it, so read the rows top-down once
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) nxt[l] = W.phi(kp1, l);
    // the reference restores i = 1..k in ascending order, each row from the NOT YET restored row above it: identical to
    // walking down from i = k with the original value of row i + 1 carried along
    #pragma unroll 1
    for (int i = k; i >= 1; --i) {
        const Rcp bi = rcp_of(W.beta(i));
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) {
            const double cur = W.phi(i, l);
            W.phi(i, l) = sg_quot(cur - nxt[l], bi);
            nxt[l] = cur;
        }
    }
    #pragma unroll 1
    for (int i = 2; i <= k; ++i) W.psi(i - 1) = W.psi(i) - h;
    const int ifail = W.hi(HL::I_IFAIL) + 1;
    W.hi(HL::I_IFAIL) = ifail;
    double temp2 = 0.5;
    if (3 < ifail) { if (p5eps < 0.25 * erk) temp2 = sg_sqrt(sg_div(p5eps, erk)); }
    if (3 <= ifail) knew = 1;
    double hn = temp2 * h;
    W.hi(HL::I_K) = knew;
    if (fabs(hn) < fouru * fabs(x)) {
        W.hs(HL::H) = copysign(fouru * fabs(x), hn);
        eps = eps + eps;
        return 2;
    }
    W.hs(HL::H) = hn;
    return 1;
}
// step, after the derivative at the corrected point (:1147-1231): update differences, choose order and step
template <int NV> RD_INLINE void sg2_after_correct(int neqn, const SgSlot<NV> &W, double eps, int &bits, const double (&yp)[NV]) {
    using HL = SgHot<NV>;
    const double fouru = 4.0 * DBL_EPSILON;
    int k = W.hi(HL::I_K);
    const int kp1 = k + 1, kp2 = k + 2, km1 = k - 1, knew = W.hi(HL::I_KNEW), ns = W.hi(HL::I_NS);
    const double absh = W.hs(HL::ABSH), p5eps = 0.5 * eps, h = W.hs(HL::H), erkm1 = W.hs(HL::ERKM1);
    double erk = W.hs(HL::ERK);
    if (knew == km1 || k == 12) bits &= ~B_PHASE1;
    const bool phase1 = (bits & B_PHASE1) != 0;
    const bool want_erkp1 = !phase1 && knew != km1 && kp1 <= ns;
    double erkp1 = 0.0;
    double d[NV];
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) {
        d[l] = yp[l] - W.phi(1, l);
        W.phi(kp1, l) = d[l];
        const double e2 = d[l] - W.phi(kp2, l);
        W.phi(kp2, l) = e2;
        if (want_erkp1) { const double q = sg_div(e2, W.wt(l)); erkp1 = erkp1 + q * q; }
    }
    #pragma unroll 1
    for (int i = 1; i <= k; ++i) {
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) W.phi(i, l) = W.phi(i, l) + d[l];
    }
    if (phase1) {
        k = kp1; erk = 0.0;
    } else if (knew == km1) {
        k = km1; erk = erkm1;
    } else if (kp1 <= ns) {
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
    if (!phase1) {
        if (p5eps < erk * kSGtwo[k + 1]) {
            hnew = h;
            if (p5eps < erk) {
                const double r = sg_root(sg_div(p5eps, erk), k + 1);
                hnew = absh * fmax(0.5, fmin((double)0.9f, r));
                hnew = copysign(fmax(hnew, fouru * fabs(W.hs(HL::X))), h);
            }
        }
    }
    W.hi(HL::I_K) = k;
    W.hs(HL::H) = hnew;
}
// intrp (ode_RAYS.f90:1235-1362) into the ray vector v; the segment is over, so g/ww of the slot serve as scratch
template <int NV> RD_INLINE void sg2_intrp(int neqn, const SgSlot<NV> &W, double xout) {
    using HL = SgHot<NV>;
    const double hi = xout - W.hs(HL::X);
    const int ki = W.hi(HL::I_KOLD) + 1;
    #pragma unroll 1
    for (int i = 1; i <= ki; ++i) W.ww(i) = kSGinv[i];
    W.g(1) = 1.0;
    double term = 0.0;
    #pragma unroll 1
    for (int j = 2; j <= ki; ++j) {
        const double psijm1 = W.psi(j - 1);
        const Rcp pj = rcp_of(psijm1);
        const double gamma = sg_quot(hi + term, pj);
        const double eta = sg_quot(hi, pj);
        #pragma unroll 1
        for (int i = 1; i <= ki + 1 - j; ++i) W.ww(i) = gamma * W.ww(i) - eta * W.ww(i + 1);
        W.g(j) = W.ww(1);
        term = psijm1;
    }
    double yo[NV];
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) yo[l] = 0.0;
    #pragma unroll 1
    for (int j = 1; j <= ki; ++j) {
        const int i = ki + 1 - j;
        const double gi = W.g(i);
        #pragma unroll (SgUnroll<NV>::L)
        for (int l = 0; l < NV; ++l) if (l < neqn) yo[l] = yo[l] + gi * W.phi(i, l);
    }
    #pragma unroll (SgUnroll<NV>::L)
    for (int l = 0; l < NV; ++l) if (l < neqn) W.v(l) = W.yy(l) + hi * yo[l];
}

// The general path (order k > kSgKM, or propagated-roundoff control on) runs for a few steps in ten thousand: it is kept OUT OF LINE,
// away from the hot loop's instruction-cache footprint, and exchanges p / yp with the caller through the slot (no array
// references cross the call, which would force the caller's register arrays into local memory).
template <int NV> static RD_NOINLINE void sg2_predict_ool(int neqn, const SgSlot<NV> W, int bits) {
    double p[NV];
    sg2_predict<NV>(neqn, W, bits, p);
    for (int l = 0; l < neqn; ++l) W.p(l) = p[l];
}
template <int NV> static RD_NOINLINE int sg2_after_predict_ool(int neqn, const SgSlot<NV> W, double eps, int bits) {
    using HL = SgHot<NV>;
    double p[NV], yp[NV];
    for (int l = 0; l < NV; ++l) { p[l] = l < neqn ? W.p(l) : 0.0; yp[l] = l < neqn ? W.yp(l) : 0.0; }
    const int r = sg2_after_predict<NV>(neqn, W, eps, bits, p, yp);
    W.hs(HL::EPS) = eps; W.hi(HL::I_BITS) = bits;      // handed back through the slot
    return r;
}
template <int NV> static RD_NOINLINE int sg2_after_correct_ool(int neqn, const SgSlot<NV> W, double eps, int bits) {
    double yp[NV];
    for (int l = 0; l < NV; ++l) yp[l] = l < neqn ? W.yp(l) : 0.0;
    sg2_after_correct<NV>(neqn, W, eps, bits, yp);
    return bits;
}

// ---- register-resident fast path: order k <= kSgKM, no propagated-roundoff control (99.95 % of the internal steps of the bench
// fans).  The live state of the slot is loaded in one burst (independent loads: one memory latency instead of one per loop
// iteration), the statements of sg2_predict / sg2_after_predict / sg2_after_correct run on register arrays with loops bounded by
// the compile-time kSgKM and predicated on the run-time k / ns (same operations on the same operands in the same order, so the
// results are the general path's bit for bit; tests compare the two), and the modified state is stored in one burst.
#define SG3_FOR(i, lo, hi) _Pragma("unroll") for (int i = (lo); i <= (hi); ++i)
#define SG3_L(l) _Pragma("unroll") for (int l = 0; l < NV; ++l)

template <int NV> RD_INLINE void sg3_predict(const SgSlot<NV> &W, double (&p)[NV]) {
    using HL = SgHot<NV>;
    constexpr int KM = kSgKM;
    const int k = W.hi(HL::I_K), kold = W.hi(HL::I_KOLD);
    int ns = W.hi(HL::I_NS);
    const double h = W.hs(HL::H), hold = W.hs(HL::HOLD), x = W.hs(HL::X);
    const int kp1 = k + 1, kp2 = k + 2;
    double psi[KM + 2], alpha[KM + 2], beta[KM + 2], sig[KM + 3], vv[KM + 3], ww[KM + 3], g[KM + 3];
    double phi[KM + 3][NV], yy[NV];
    SG3_FOR(i, 1, KM) { psi[i] = W.psi(i); alpha[i] = W.alpha(i); beta[i] = W.beta(i); vv[i] = W.vv(i); }
    SG3_FOR(i, 1, KM + 1) { sig[i] = W.sig(i); g[i] = W.g(i); }
    SG3_FOR(i, 1, KM + 1) SG3_L(l) phi[i][l] = W.phi(i, l);
    SG3_L(l) yy[l] = W.yy(l);
    vv[KM + 1] = 0.0; ww[KM + 1] = 0.0; ww[KM + 2] = 0.0; sig[KM + 2] = 0.0; g[KM + 2] = 0.0; psi[KM + 1] = 0.0; alpha[KM + 1] = 0.0; beta[KM + 1] = 0.0;
    SG3_FOR(i, 1, KM) ww[i] = 0.0;
    if (h != hold) ns = 0;
    if (ns <= kold) ns = ns + 1;
    const int nsp1 = ns + 1;
    if (ns <= k) {
        SG3_FOR(i, 1, KM) if (i == ns) { beta[i] = 1.0; alpha[i] = kSGinv[i]; }
        double temp1 = h * (double)ns;
        SG3_FOR(i, 2, KM + 1) if (i == nsp1) sig[i] = 1.0;
        SG3_FOR(i, 2, KM) if (i >= nsp1 && i <= k) {
            const double temp2 = psi[i - 1];
            psi[i - 1] = temp1;
            beta[i] = sg_div(beta[i - 1] * psi[i - 1], temp2);
            temp1 = temp2 + h;
            alpha[i] = sg_div(h, temp1);
            sig[i + 1] = (double)i * alpha[i] * sig[i];
        }
        SG3_FOR(i, 1, KM) if (i == k) psi[i] = temp1;
        if (ns <= 1) {
            SG3_FOR(iq, 1, KM) if (iq <= k) { vv[iq] = kSGinvTri[iq]; ww[iq] = vv[iq]; }
        } else {
            if (kold < k) {
                SG3_FOR(i, 1, KM) if (i == k) vv[i] = kSGinvTri[i];
                SG3_FOR(j, 1, KM - 1) if (j <= ns - 2) {       // i = k - j: v(i) = v(i) - alpha(j+1)*v(i+1)
                    SG3_FOR(i, 1, KM - 1) if (i == k - j) vv[i] = vv[i] - alpha[j + 1] * vv[i + 1];
                }
            }
            double alns = 0.0;
            SG3_FOR(i, 1, KM) if (i == ns) alns = alpha[i];
            SG3_FOR(iq, 1, KM) if (iq <= kp1 - ns) { vv[iq] = vv[iq] - alns * vv[iq + 1]; ww[iq] = vv[iq]; }
            SG3_FOR(i, 2, KM + 1) if (i == nsp1) g[i] = ww[1];
        }
        SG3_FOR(i, 3, KM + 1) if (i >= ns + 2 && i <= kp1) {
            SG3_FOR(iq, 1, KM) if (iq <= kp2 - i) ww[iq] = ww[iq] - alpha[i - 1] * ww[iq + 1];
            g[i] = ww[1];
        }
    }
    SG3_FOR(i, 2, KM) if (i >= nsp1 && i <= k) SG3_L(l) phi[i][l] = beta[i] * phi[i][l];
    double up[NV];
    SG3_L(l) { p[l] = 0.0; up[l] = 0.0; }
    // phi(kp2) = phi(kp1); phi(kp1) = 0
    SG3_FOR(i, 2, KM + 1) if (i == kp1) SG3_L(l) { phi[i + 1][l] = phi[i][l]; phi[i][l] = 0.0; }
    _Pragma("unroll") for (int i = KM; i >= 1; --i) if (i <= k) {
        SG3_L(l) {
            const double f = phi[i][l];
            p[l] = p[l] + f * g[i];
            up[l] = f + up[l];
            phi[i][l] = up[l];
        }
    }
    SG3_L(l) p[l] = yy[l] + h * p[l];
    // burst store of what changed
    W.hi(HL::I_NS) = ns;
    SG3_FOR(i, 1, KM) if (i <= k) { W.psi(i) = psi[i]; W.alpha(i) = alpha[i]; W.beta(i) = beta[i]; W.vv(i) = vv[i]; }
    SG3_FOR(i, 1, KM + 1) if (i <= kp1) { W.sig(i) = sig[i]; W.g(i) = g[i]; }
    SG3_FOR(i, 1, KM + 2) if (i <= kp2) SG3_L(l) W.phi(i, l) = phi[i][l];
    W.hs(HL::XOLD) = x;
    W.hs(HL::X) = x + h;
    W.hs(HL::ABSH) = fabs(h);
}

template <int NV> RD_INLINE int sg3_after_predict(const SgSlot<NV> &W, double &eps, int &bits, const double (&p)[NV], const double (&yp)[NV]) {
    using HL = SgHot<NV>;
    constexpr int KM = kSgKM;
    const double fouru = 4.0 * DBL_EPSILON;
    const int k = W.hi(HL::I_K), kp1 = k + 1, km1 = k - 1, km2 = k - 2;
    const double absh = W.hs(HL::ABSH), p5eps = 0.5 * eps, h = W.hs(HL::H), xold = W.hs(HL::XOLD);
    const int ifail0 = W.hi(HL::I_IFAIL);
    double sig[KM + 2], g[KM + 2], beta[KM + 1], psi[KM + 1], wt[NV], phi[KM + 2][NV];
    SG3_FOR(i, 1, KM + 1) { sig[i] = W.sig(i); g[i] = W.g(i); }
    SG3_FOR(i, 1, KM) { beta[i] = W.beta(i); psi[i] = W.psi(i); }
    SG3_L(l) wt[l] = W.wt(l);
    SG3_FOR(i, 1, KM + 1) SG3_L(l) phi[i][l] = W.phi(i, l);
    double erkm2 = 0.0, erkm1 = 0.0, erk = 0.0;
    double dl[NV];
    SG3_L(l) {
        const Rcp wl = rcp_of(wt[l]);   // up to three quotients by the same weight
        const double ph1 = phi[1][l];
        if (0 < km2) { const double q = sg_quot(phi[KM - 1][l] + yp[l] - ph1, wl); erkm2 = erkm2 + q * q; }   // k = KM = 3: km1 = 2
        if (0 <= km2) { const double phk = k == 2 ? phi[2][l] : phi[KM][l]; const double q = sg_quot(phk + yp[l] - ph1, wl); erkm1 = erkm1 + q * q; }
        dl[l] = yp[l] - ph1;
        const double q = sg_quot(dl[l], wl);
        erk = erk + q * q;
    }
    double sig_km1 = 0.0, sig_k = 0.0, sig_kp1 = 0.0, g_k = 0.0, g_kp1 = 0.0;
    SG3_FOR(i, 1, KM + 1) { if (i == km1) sig_km1 = sig[i]; if (i == k) { sig_k = sig[i]; g_k = g[i]; } if (i == kp1) { sig_kp1 = sig[i]; g_kp1 = g[i]; } }
    if (0 < km2) erkm2 = absh * sig_km1 * kSGgstr[km2] * sg_sqrt(erkm2);
    if (0 <= km2) erkm1 = absh * sig_k * kSGgstr[km1] * sg_sqrt(erkm1);
    const double rt_erk = sg_sqrt(erk);
    const double err = absh * rt_erk * (g_k - g_kp1);
    erk = absh * rt_erk * sig_kp1 * kSGgstr[k];
    int knew = k;
    if (0 < km2) {
        if (fmax(erkm1, erkm2) <= erk) knew = km1;
    } else if (0 == km2) {
        if (erkm1 <= 0.5 * erk) knew = km1;
    }
    W.hi(HL::I_KNEW) = knew; W.hs(HL::ERK) = erk; W.hs(HL::ERKM1) = erkm1;
    if (err <= eps) {   // accepted: correct (:1123-1141)
        W.hi(HL::I_KOLD) = k;
        W.hs(HL::HOLD) = h;
        SG3_L(l) W.yy(l) = p[l] + h * g_kp1 * dl[l];
        return 0;
    }
    // step failed (:1076-1110): restore x, phi, psi; halve (or more) the step
    bits &= ~B_PHASE1;
    W.hs(HL::X) = xold;
    SG3_FOR(i, 1, KM) if (i <= k) {        // ascending: row i + 1 is still the unrestored one when row i reads it
        const Rcp bi = rcp_of(beta[i]);
        SG3_L(l) phi[i][l] = sg_quot(phi[i][l] - phi[i + 1][l], bi);
    }
    SG3_FOR(i, 2, KM) if (i <= k) psi[i - 1] = psi[i] - h;
    SG3_FOR(i, 1, KM) if (i <= k) SG3_L(l) W.phi(i, l) = phi[i][l];
    SG3_FOR(i, 1, KM - 1) if (i <= k - 1) W.psi(i) = psi[i];
    const int ifail = ifail0 + 1;
    W.hi(HL::I_IFAIL) = ifail;
    double temp2 = 0.5;
    if (3 < ifail) { if (p5eps < 0.25 * erk) temp2 = sg_sqrt(sg_div(p5eps, erk)); }
    if (3 <= ifail) knew = 1;
    const double hn = temp2 * h;
    W.hi(HL::I_K) = knew;
    if (fabs(hn) < fouru * fabs(xold)) {
        W.hs(HL::H) = copysign(fouru * fabs(xold), hn);
        eps = eps + eps;
        return 2;
    }
    W.hs(HL::H) = hn;
    return 1;
}

template <int NV> RD_INLINE void sg3_after_correct(const SgSlot<NV> &W, double eps, int &bits, const double (&yp)[NV]) {
    using HL = SgHot<NV>;
    constexpr int KM = kSgKM;
    const double fouru = 4.0 * DBL_EPSILON;
    int k = W.hi(HL::I_K);
    const int kp1 = k + 1, kp2 = k + 2, km1 = k - 1, knew = W.hi(HL::I_KNEW), ns = W.hi(HL::I_NS);
    const double absh = W.hs(HL::ABSH), p5eps = 0.5 * eps, h = W.hs(HL::H), erkm1 = W.hs(HL::ERKM1), x = W.hs(HL::X);
    double erk = W.hs(HL::ERK);
    double phi[KM + 3][NV], wt[NV];
    SG3_FOR(i, 1, KM + 2) SG3_L(l) phi[i][l] = W.phi(i, l);
    SG3_L(l) wt[l] = W.wt(l);
    if (knew == km1 || k == 12) bits &= ~B_PHASE1;
    const bool phase1 = (bits & B_PHASE1) != 0;
    const bool want_erkp1 = !phase1 && knew != km1 && kp1 <= ns;
    double erkp1 = 0.0;
    SG3_L(l) {
        const double d = yp[l] - phi[1][l];
        double old2 = 0.0;
        SG3_FOR(i, 3, KM + 2) if (i == kp2) old2 = phi[i][l];
        const double e2 = d - old2;
        SG3_FOR(i, 2, KM + 1) if (i == kp1) { phi[i][l] = d; phi[i + 1][l] = e2; }
        SG3_FOR(i, 1, KM) if (i <= k) phi[i][l] = phi[i][l] + d;
        if (want_erkp1) { const double q = sg_div(e2, wt[l]); erkp1 = erkp1 + q * q; }
    }
    SG3_FOR(i, 1, KM + 2) if (i <= kp2) SG3_L(l) W.phi(i, l) = phi[i][l];
    if (phase1) {
        k = kp1; erk = 0.0;
    } else if (knew == km1) {
        k = km1; erk = erkm1;
    } else if (kp1 <= ns) {
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
    if (!phase1) {
        if (p5eps < erk * kSGtwo[k + 1]) {
            hnew = h;
            if (p5eps < erk) {
                const double r = sg_root(sg_div(p5eps, erk), k + 1);
                hnew = absh * fmax(0.5, fmin((double)0.9f, r));
                hnew = copysign(fmax(hnew, fouru * fabs(x)), h);
            }
        }
    }
    W.hi(HL::I_K) = k;
    W.hs(HL::H) = hnew;
}

// intrp for kold <= kSgKM (ki = kold + 1 <= 4) on registers: the statements of sg2_intrp, loops bounded at compile time
template <int NV> RD_INLINE void sg3_intrp(const SgSlot<NV> &W, double xout) {
    using HL = SgHot<NV>;
    constexpr int KI = kSgKM + 1;
    const double hi = xout - W.hs(HL::X);
    const int ki = W.hi(HL::I_KOLD) + 1;
    double w[KI + 2], g[KI + 1], psi[KI];
    SG3_FOR(i, 1, KI) w[i] = kSGinv[i];
    w[KI + 1] = 0.0;
    SG3_FOR(i, 1, KI - 1) psi[i] = W.psi(i);
    g[1] = 1.0;
    SG3_FOR(i, 2, KI) g[i] = 0.0;
    double term = 0.0;
    SG3_FOR(j, 2, KI) if (j <= ki) {
        const double psijm1 = psi[j - 1];
        const Rcp pj = rcp_of(psijm1);
        const double gamma = sg_quot(hi + term, pj);
        const double eta = sg_quot(hi, pj);
        SG3_FOR(i, 1, KI - 1) if (i <= ki + 1 - j) w[i] = gamma * w[i] - eta * w[i + 1];
        g[j] = w[1];
        term = psijm1;
    }
    double yo[NV];
    SG3_L(l) yo[l] = 0.0;
    _Pragma("unroll") for (int i = KI; i >= 1; --i) if (i <= ki) {       // j = 1..ki of the reference: i = ki + 1 - j descends
        SG3_L(l) yo[l] = yo[l] + g[i] * W.phi(i, l);
    }
    SG3_L(l) W.v(l) = W.yy(l) + hi * yo[l];
}

// streaming copy-out of one finished ray by the whole warp: out of line, it runs once per ray
static RD_NOINLINE void sg2_flush_row(const TraceArgs &a, long long ir, int np, int pf, size_t rw, int nv, unsigned lane) {
    if (a.host_ray_vec)
        copy_row_to_host(a.host_ray_vec + ((size_t)(a.host_ray0 + ir * a.host_ray_stride) * a.host_npoints_alloc + pf) * nv, a.ray_vec + rw * a.npoints_alloc * nv, np * nv, lane);
    if (a.host_residual)
        copy_row_to_host(a.host_residual + (size_t)(a.host_ray0 + ir * a.host_ray_stride) * a.host_npoints_alloc + pf, a.residual + rw * a.npoints_alloc, np, lane);
}

// global record bytes of one CTA; shared-memory bytes of one slot; shared-memory bytes besides the slots (rings; the deposition
// bins come on top)
template <int NV> constexpr size_t sg2_state_bytes_per_cta() { return (size_t)kSgStride * (SgLayout<NV>::NDBL * 8 + SgLayout<NV>::NINT * 4); }
template <int NV> constexpr size_t sg2_hot_bytes_per_slot() { return (size_t)SgHot<NV>::NDBL * 8; }
constexpr size_t sg2_ring_bytes() { return (size_t)kSgNQ * kSgRingCap * sizeof(unsigned short); }

// ---- the kernel -------------------------------------------------------------------------------------------------------------
#ifndef RAYS_SG_FAST
#define RAYS_SG_FAST 1
#endif
template <class T>
__global__ void __launch_bounds__(kSgBlock, kSgCtas) trace_sg2_kernel(const TraceArgs a) {
    constexpr int NV = T::NV;
    constexpr int NSM = NSpec<T::NS>::MAX;
    constexpr int NT = kSgBlock;
    constexpr int NW = kSgBlock / 32;
    using L = SgLayout<NV>;
    using HL = SgHot<NV>;
    const int S = a.sg_slots;                   // slots of this CTA
    const int nv = T::nv();
    const rays_cfg &c = g_dc.c;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const int warp = tid >> 5;
    const int maxnum = 500;
    const double fouru = 4.0 * DBL_EPSILON;
    __shared__ unsigned s_tail[kSgNQ];
    __shared__ unsigned s_head[2][kSgNQ];
    __shared__ unsigned char s_kind[kSgRingCap];
    __shared__ int s_exhausted;
    // dynamic shared memory: [deposition bins | hot records | rings]
    double *const hot = reinterpret_cast<double *>(s_dep_bins) + ((a.dep_smem + 15) / 16) * 2;
    unsigned short *const ring = reinterpret_cast<unsigned short *>(hot + (size_t)S * HL::NDBL);
    const size_t gcta = blockIdx.x;
    double *const D = a.sg_state + gcta * L::NDBL * kSgStride;
    int *const I = reinterpret_cast<int *>(a.sg_state + (size_t)gridDim.x * L::NDBL * kSgStride) + gcta * L::NINT * kSgStride;
    const bool binning = a.dep_acc != nullptr && T::damp();
    const DepBins dbins = dep_begin(a, binning);
    const bool streaming = a.host_ray_vec != nullptr || a.host_residual != nullptr;
    unsigned long long my_steps = 0;
    unsigned my_rhs = 0;
    if (tid < kSgNQ) { s_tail[tid] = tid == Q_IDLE ? (unsigned)S : 0u; s_head[0][tid] = 0u; s_head[1][tid] = 0u; }
    if (tid == 0) s_exhausted = 0;
    for (int j = tid; j < S; j += NT) { ring[Q_IDLE * kSgRingCap + j] = (unsigned short)j; s_kind[j] = K_IDLE; }

    // push slot `sl` (or nothing: q < 0) onto ring q: one shared-memory atomic per ring and warp
    auto push = [&](int q, int sl) {
        const unsigned m = __match_any_sync(0xffffffffu, q);
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if ((int)lane == leader && q >= 0) base = atomicAdd(&s_tail[q], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (q >= 0) ring[q * kSgRingCap + ((base + __popc(m & ((1u << lane) - 1u))) & (kSgRingCap - 1))] = (unsigned short)sl;
    };

    for (unsigned iter = 0;; ++iter) {
        __syncthreads();     // THE barrier of the iteration: the pushes of the last one are visible, the heads of this one are written
        const int par = iter & 1;
        unsigned hd[kSgNQ], cnt[kSgNQ];
#pragma unroll
        for (int q = 0; q < kSgNQ; ++q) { hd[q] = s_head[par][q]; cnt[q] = s_tail[q] - hd[q]; }
        const bool exhausted = s_exhausted != 0;
        // ---- the plan every thread comes to (same counters, same arithmetic): finished rays first, then a refill when it pays
        // (or the waiting slots cannot fill the CTA); else a COMPUTE iteration, in which every WARP gets a batch of one kind:
        // full batches of 32 first (predictor, corrector, segment boundary, then the rare kinds), then the largest remainders
        // for the warps still free.  At the barrier every slot sits in a ring, so with more slots than threads all warps are
        // busy in every iteration, and a warp runs one kind: fast / general bookkeeping is not a per-thread choice.
        int q = -1;                  // Q_FIN / Q_IDLE: an iteration of the whole CTA; -2: compute iteration
        int wq = -1;                 // compute iteration: this warp's ring, ...
        unsigned wbase = 0, wn = 0;  // ... its first entry (from the head) and its entries
        unsigned tk[kSgNQ];
#pragma unroll
        for (int k = 0; k < kSgNQ; ++k) tk[k] = 0u;
        const unsigned total = cnt[Q_PRED] + cnt[Q_CORR] + cnt[Q_CHECK] + cnt[Q_START] + cnt[Q_GPRED] + cnt[Q_GCORR];
        if (cnt[Q_FIN] > 0) { q = Q_FIN; tk[Q_FIN] = cnt[Q_FIN]; }
        else if (cnt[Q_IDLE] > 0 && !exhausted && (cnt[Q_IDLE] >= 32u || total < (unsigned)NT)) { q = Q_IDLE; tk[Q_IDLE] = cnt[Q_IDLE] < (unsigned)NT ? cnt[Q_IDLE] : (unsigned)NT; }
        else if (total > 0) {
            q = -2;
            int wnext = 0;
            // sg_mixed = 0: ONE kind per iteration (the one most slots wait for).  The kernels whose right-hand side is short
            // (deriv_cold) spend half their instructions in the bookkeeping blocks, which differ from kind to kind: with
            // several kinds in flight the SM's instruction cache thrashes (ncu: hit rate 87 % -> 61 %, 3.7 of 9 stall cycles
            // per issue on instruction fetch) and an iteration lasts as long as its slowest kind.  deriv_num (14 determinants
            // per right-hand side) is the other way round: there filling every warp wins (+33 %).
            if (a.sg_mixed != 1) {
                int best = Q_PRED;
                unsigned bc = cnt[Q_PRED];
                if (cnt[Q_CORR] > bc) { best = Q_CORR; bc = cnt[Q_CORR]; }
                if (cnt[Q_CHECK] > bc) { best = Q_CHECK; bc = cnt[Q_CHECK]; }
                int rare = Q_START;
                unsigned rc = cnt[Q_START];
                if (cnt[Q_GPRED] > rc) { rare = Q_GPRED; rc = cnt[Q_GPRED]; }
                if (cnt[Q_GCORR] > rc) { rare = Q_GCORR; rc = cnt[Q_GCORR]; }
                if (rc > 0 && (rc >= 16u || bc == 0u || (iter & 255u) == 0u)) best = rare;
                // sg_mixed = 2: predictor and corrector batches (which share the right-hand side, the largest block) may run
                // together, the segment-boundary and the rare kinds run on their own
                const bool pair = a.sg_mixed == 2 && (best == Q_PRED || best == Q_CORR);
                if (!pair) {      // one ring: warp w takes its entries 32 w ... (the general plan below cost 10 % of all instructions)
                    unsigned cb = 0;
#pragma unroll
                    for (int k = 0; k < Q_FIN; ++k) if (k == best) cb = cnt[k];
                    const unsigned tkn = cb < (unsigned)NT ? cb : (unsigned)NT;
#pragma unroll
                    for (int k = 0; k < Q_FIN; ++k) if (k == best) tk[k] = tkn;
                    wbase = 32u * (unsigned)warp;
                    if (wbase < tkn) { wq = best; wn = tkn - wbase < 32u ? tkn - wbase : 32u; }
                    wnext = NW;       // (no warp left for the loops below)
#pragma unroll
                    for (int k = 0; k < Q_FIN; ++k) cnt[k] = tk[k];
                } else {
#pragma unroll
                    for (int k = 0; k < Q_FIN; ++k) if (k != Q_PRED && k != Q_CORR) cnt[k] = 0u;      // (this iteration sees only those rings)
                }
            }
            if (a.sg_mixed == 1 && (iter & 63u) == 0u) {    // aging: a rare kind that never fills a batch gets one warp now and then
                int rare = Q_START;
                unsigned rc = cnt[Q_START];
                if (cnt[Q_GPRED] > rc) { rare = Q_GPRED; rc = cnt[Q_GPRED]; }
                if (cnt[Q_GCORR] > rc) { rare = Q_GCORR; rc = cnt[Q_GCORR]; }
                if (rc > 0) {
                    const unsigned n = rc < 32u ? rc : 32u;
                    if (warp == 0) { wq = rare; wbase = 0; wn = n; }
#pragma unroll
                    for (int k = 0; k < kSgNQ; ++k) if (k == rare) tk[k] = n;
                    wnext = 1;
                }
            }
            if (wnext < NW) {
#pragma unroll
            for (int k = 0; k < Q_FIN; ++k) {          // full batches
                const int full = (int)((cnt[k] - tk[k]) >> 5);
                const int use = full < NW - wnext ? full : NW - wnext;
                if (warp >= wnext && warp < wnext + use) { wq = k; wbase = tk[k] + 32u * (unsigned)(warp - wnext); wn = 32u; }
                tk[k] += 32u * (unsigned)use;
                wnext += use;
            }
            for (; wnext < NW; ++wnext) {               // the largest remainders
                int bk = -1;
                unsigned br = 0;
#pragma unroll
                for (int k = 0; k < Q_FIN; ++k) { const unsigned r = cnt[k] - tk[k]; if (r > br) { br = r; bk = k; } }
                if (bk < 0) break;
                const unsigned n = br < 32u ? br : 32u;
                if (warp == wnext) { wq = bk; wbase = 0; wn = n; }
#pragma unroll
                for (int k = 0; k < Q_FIN; ++k) if (k == bk) { if (warp == wnext) wbase = tk[k]; tk[k] += n; }
            }
            }
        }
        if (q == -1) break;      // every slot idle and the queue empty
        if (tid < kSgNQ) {
            unsigned t = 0;
#pragma unroll
            for (int k = 0; k < kSgNQ; ++k) if (tid == k) t = tk[k];
            s_head[par ^ 1][tid] = hd[tid] + t;
        }
        const unsigned take = q == Q_FIN ? tk[Q_FIN] : tk[Q_IDLE];
        // ---- finished rays: streaming copy-out by the warps, slot idle
        if (q == Q_FIN) {
            for (unsigned e = warp; e < take; e += NW) {
                const int sl = ring[Q_FIN * kSgRingCap + ((hd[Q_FIN] + e) & (kSgRingCap - 1))];
                if (streaming) {
                    const SgSlot<NV> F{hot + (size_t)sl * HL::NDBL, D + sl, I + sl};
                    const long long ir = (long long)F.f(L::IRAY);
                    const int np = F.i(L::FINNP), pf = F.i(L::P0);
                    sg2_flush_row(a, ir, np, pf, gcta * S + sl, nv, lane);
                }
                if (lane == 0) { s_kind[sl] = K_IDLE; const unsigned pos = atomicAdd(&s_tail[Q_IDLE], 1u); ring[Q_IDLE * kSgRingCap + (pos & (kSgRingCap - 1))] = (unsigned short)sl; }
            }
            continue;
        }
        // ---- refill idle slots from the work queue: one global atomic per warp
        if (q == Q_IDLE) {
            const int slot = tid < (int)take ? (int)ring[Q_IDLE * kSgRingCap + ((hd[Q_IDLE] + tid) & (kSgRingCap - 1))] : -1;
            const unsigned want = __ballot_sync(0xffffffffu, slot >= 0);
            int nq = -1;
            if (want) {
                unsigned long long base = 0;
                const int leader = __ffs(want) - 1;
                if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(want));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (slot >= 0) {
                    const long long idx = (long long)(base + __popc(want & ((1u << lane) - 1u)));
                    nq = Q_IDLE;
                    if (idx >= a.nray) s_exhausted = 1;
                    else {
                        const SgSlot<NV> W{hot + (size_t)slot * HL::NDBL, D + slot, I + slot};
                        const long long iray = a.order ? (long long)a.order[idx] : idx;
                        const size_t row = streaming ? gcta * S + slot : (size_t)iray;
                        W.f(L::IRAY) = (double)iray;
                        W.f(L::PWR) = a.ray_pwr_wt ? a.ray_pwr_wt[iray] : 0.0;
                        W.i(L::SLICE) = 0;
                        double v[NV];
                        if (a.resume) {   // a ray suspended by an earlier launch: its last point is still to be checked and saved
                            RayCarry k;
                            resume_ray(a, iray, v, nv, k);
                            W.f(L::S_) = k.s; W.f(L::SOUT) = k.sout; W.i(L::NSTEP) = k.nstep; W.i(L::FLAG) = k.flag;
                            W.f(L::RPREV) = k.resid_prev; W.f(L::RLAST) = k.resid_last; W.f(L::RMAX) = k.resid_max;
                            W.f(L::DEPX) = k.dep_x; W.f(L::DEPQ) = k.dep_Q; W.f(L::REL) = k.rel_err; W.f(L::ABS) = k.abs_err;
                            W.i(L::P0) = streaming ? k.nstep + 1 : 0;
                            W.hi(HL::I_BITS) = 0;
                        } else {
                            W.f(L::S_) = 0.0; W.f(L::SOUT) = 0.0; W.i(L::NSTEP) = 0; W.i(L::FLAG) = 0; W.i(L::P0) = 0;
                            W.f(L::REL) = c.rel_err0; W.f(L::ABS) = c.abs_err0;   // ray_init_SG_ode (SG_ode_m.f90:73-85)
                            W.f(L::RPREV) = 0.0; W.f(L::RLAST) = 0.0; W.f(L::RMAX) = 0.0; W.f(L::DEPX) = 0.0; W.f(L::DEPQ) = 0.0;
                            initialize_ode_vector<T>(a.rvec0 + 3 * iray, a.rindex_vec0 + 3 * iray, v);
                            if (a.ray_vec) {
                                double *dst = a.ray_vec + row * a.npoints_alloc * nv;
                                if (T::GENERIC) store_point(dst, v, nv); else store_point_fixed<NV>(dst, v);
                            }
                            if (a.residual) a.residual[row * a.npoints_alloc] = 0.0;
                            if (a.start_ray_vec) for (int i = 0; i < nv; ++i) a.start_ray_vec[(size_t)iray * nv + i] = v[i];
                            W.hi(HL::I_BITS) = B_FIRST;
                        }
#pragma unroll
                        for (int i = 0; i < NV; ++i) if (i < nv) W.v(i) = v[i];
                        W.hs(HL::EPS) = 0.0;
                        s_kind[slot] = K_CHECK;
                        nq = Q_CHECK;
                    }
                }
            }
            push(nq, slot);
            continue;
        }
        // ---- one macro-step of the slot: bookkeeping -> one right-hand side -> bookkeeping
        int slot = -1;
        if (wq >= 0 && lane < wn) {
            unsigned hq = 0;
#pragma unroll
            for (int k = 0; k < Q_FIN; ++k) if (wq == k) hq = hd[k];
            slot = (int)ring[wq * kSgRingCap + ((hq + wbase + lane) & (kSgRingCap - 1))];
        }
        const bool gen = wq == Q_GPRED || wq == Q_GCORR || T::GENERIC || !RAYS_SG_FAST;
        int next = -1;
        if (slot >= 0) {
            const SgSlot<NV> W{hot + (size_t)slot * HL::NDBL, D + slot, I + slot};
            const int kind = s_kind[slot];
            int bits = W.hi(HL::I_BITS);
            next = kind;
            int req = 0;   // 1: derivative at yy (start), 2: at the predicted p, 3: at the corrected yy,
                           // 4: check_save of the new point v + the start derivative of the next segment there
            bool stop = false, did_not_start = false, crashed = false, enter = false, de_top = false, have_f1 = false;
            int f1_code = 0;
            int flag = INT_MIN;            // the ray's sticky flag lives in the global record: read where it can change
            double eps = W.hs(HL::EPS);
            double uu[NV], ff[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) { uu[i] = 0.0; ff[i] = 0.0; }
            if (kind == K_PRED) {
                if (!gen) sg3_predict<NV>(W, uu);
                else if (T::GENERIC) sg2_predict<NV>(nv, W, bits, uu);
                else {
                    sg2_predict_ool<NV>(nv, W, bits);
#pragma unroll
                    for (int i = 0; i < NV; ++i) if (i < nv) uu[i] = W.p(i);
                }
                req = 2;
            }
            else if (kind == K_CORR) req = 3;
            else if (kind == K_START) req = 1;
            else if (kind == K_BEGIN) enter = true;
            else {   // K_CHECK
                req = 4;
                flag = W.i(L::FLAG);
                if (bits & B_INTRP) {   // past the output point: interpolate, segment done (iflag = 2)
                    bits &= ~B_INTRP;
                    const double sout = W.f(L::SOUT);
                    if (!T::GENERIC && RAYS_SG_FAST && W.hi(HL::I_KOLD) <= kSgKM) sg3_intrp<NV>(W, sout);
                    else sg2_intrp<NV>(nv, W, sout);
                    W.f(L::S_) = sout;
                    const int sn = W.i(L::SLICE) + 1;
                    W.i(L::SLICE) = sn;
                    if (a.slice_steps > 0 && sn >= a.slice_steps) {   // suspend: packed into full warps by the next launch
                        double v[NV];
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[i] = i < nv ? W.v(i) : 0.0;
                        const int nstep = W.i(L::NSTEP);
                        RayCarry k{sout, sout, W.f(L::RPREV), W.f(L::RLAST), W.f(L::RMAX), W.f(L::DEPX), W.f(L::DEPQ), W.f(L::REL), W.f(L::ABS), nstep, flag};
                        suspend_ray(a, (long long)W.f(L::IRAY), v, nv, k);
                        W.i(L::FINNP) = nstep + 1 - W.i(L::P0);
                        req = 0; next = K_FIN;
                    }
                }
            }
            // ---- equilibrium + derivative evaluation: ONE site.  req = 4 also runs check_save on the same equilibrium
            // (the reference evaluates it twice at this point: check_save.f90:38 and eqn_ray.f90:87 of the next segment's
            // first derivative), saves the point and runs the loop-top tests of ray_tracing.f90:118-172
            if (req) {
                if (req != 2) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) if (i < nv) uu[i] = W.yy(i);      // v shares yy's storage
                }
                Eq<NSM> e;
                equilibrium<T::EQ, T::NS, true>(uu[0], uu[1], uu[2], e);
                double dddx[3], dddk[3], dddw = 0.0;
                bool have_derivs = false;
                if (req == 4) {
                    const bool first = (bits & B_FIRST) != 0;
                    double resid = 0.0;
                    if (e.err) {
                        flag = e.err;
                        if (T::damp() && uu[7 < NV ? 7 : 0] > c.total_damping_limit) { stop = true; flag = RAYS_STOP_TOTAL_ABSORPTION; }
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
                        const double pwr = W.f(L::PWR);
                        if (!first) {   // the point passed check_save: save it (ray_tracing.f90:237-243)
                            const int nstep = W.i(L::NSTEP) + 1;
                            W.i(L::NSTEP) = nstep;
                            const int pp = nstep - W.i(L::P0);
                            const size_t row = streaming ? gcta * S + slot : (size_t)(long long)W.f(L::IRAY);
                            if (a.ray_vec) {
                                double *dst = a.ray_vec + (row * a.npoints_alloc + pp) * nv;
                                if (T::GENERIC) store_point(dst, uu, nv); else store_point_fixed<NV>(dst, uu);
                            }
                            if (a.residual) a.residual[row * a.npoints_alloc + pp] = resid;
                            const double rp = W.f(L::RLAST);
                            W.f(L::RPREV) = rp;
                            W.f(L::RLAST) = resid;
                            if (fabs(rp) > W.f(L::RMAX)) W.f(L::RMAX) = fabs(rp);
                            if (binning) {
                                const double xn = dep_abscissa<T::EQ>(uu), Qn = uu[7 < NV ? 7 : 0] * pwr;
                                bin_segment(dbins, W.f(L::DEPX), xn, W.f(L::DEPQ), Qn);
                                W.f(L::DEPX) = xn; W.f(L::DEPQ) = Qn;
                            }
                            ++my_steps;
                        } else if (binning) { W.f(L::DEPX) = dep_abscissa<T::EQ>(uu); W.f(L::DEPQ) = uu[7 < NV ? 7 : 0] * pwr; }
                        bits &= ~B_FIRST;
                        const double sout0 = W.f(L::SOUT);    // top of the trajectory loop (ray_tracing.f90:118-172)
                        const double sout1 = sout0 + c.ds;
                        W.f(L::S_) = sout0;
                        W.f(L::SOUT) = sout1;
                        if (sout1 > c.s_max) { stop = true; flag = RAYS_STOP_SOUT_GT_SMAX; }
                        else if (W.i(L::NSTEP) + 1 > c.nstep_max) { stop = true; flag = RAYS_STOP_NSTEP_MAX; }
                    }
                }
                if (!stop) {
                    int code = e.err;
                    if (!code && !have_derivs) code = ray_derivs<T>(e, uu, dddx, dddk, dddw);
                    if (!code) code = ray_equations<T>(e, uu, dddx, dddk, dddw, ff);
                    if (req == 4) {          // the start derivative of the next segment (consumed after de's entry tests, below)
                        have_f1 = true; f1_code = code;
                        enter = true;
                    } else {
                        my_rhs += 1;
                        if (code) { flag = code; W.f(L::SOUT) = W.f(L::S_); stop = true; }     // SG_ode: stop_ode set in eqn_ray -> sout = s
                        else if (req == 1) { sg2_after_start<NV>(nv, W, eps, bits, ff); next = K_PRED; }
                        else if (req == 2) {
                            int r;
                            if (!gen) r = sg3_after_predict<NV>(W, eps, bits, uu, ff);
                            else if (T::GENERIC) r = sg2_after_predict<NV>(nv, W, eps, bits, uu, ff);
                            else {
#pragma unroll
                                for (int i = 0; i < NV; ++i) if (i < nv) W.yp(i) = ff[i];
                                r = sg2_after_predict_ool<NV>(nv, W, eps, bits);
                                eps = W.hs(HL::EPS); bits = W.hi(HL::I_BITS);
                            }
                            if (r == 0) next = K_CORR;
                            else if (r == 1) next = K_PRED;   // failed: predict again with the reduced step
                            else crashed = true;
                        } else {
                            if (!gen) sg3_after_correct<NV>(W, eps, bits, ff);
                            else if (T::GENERIC) sg2_after_correct<NV>(nv, W, eps, bits, ff);
                            else {
#pragma unroll
                                for (int i = 0; i < NV; ++i) if (i < nv) W.yp(i) = ff[i];
                                bits = sg2_after_correct_ool<NV>(nv, W, eps, bits);
                            }
                            const int nostep = W.hi(HL::I_NOSTEP) + 1;       // de: step counter and stiffness test (ode_RAYS.f90:578-590)
                            W.hi(HL::I_NOSTEP) = nostep;
                            int kle4 = W.hi(HL::I_KLE4) + 1;
                            if (4 < W.hi(HL::I_KOLD)) kle4 = 0;
                            W.hi(HL::I_KLE4) = kle4;
                            if (50 <= kle4) bits |= B_STIFF;
                            de_top = true;
                        }
                    }
                }
            }
            if (enter && !stop) {   // ode/de entry with iflag = 1 (ode_RAYS.f90:425-505)
                const double s = W.f(L::S_), sout = W.f(L::SOUT), rel_err = W.f(L::REL), abs_err = W.f(L::ABS);
                if (s == sout) { stop = true; flag = RAYS_STOP_SG_T_EQ_TOUT; }
                else if (rel_err < 0.0 || abs_err < 0.0) { stop = true; flag = RAYS_STOP_SG_BAD_TOL; }
                else {
                    eps = fmax(rel_err, abs_err);
                    if (eps <= 0.0) { stop = true; flag = RAYS_STOP_SG_EPS_LE_0; }
                    else {
                        const double del = sout - s;
                        W.hs(HL::T0) = s;
                        W.hs(HL::ABSDEL) = fabs(del);
                        W.hs(HL::TEND) = s + 10.0 * del;
                        W.hi(HL::I_NOSTEP) = 0; W.hi(HL::I_KLE4) = 0;
                        W.hs(HL::RELEPS) = rel_err / eps;
                        W.hs(HL::ABSEPS) = abs_err / eps;
                        bits = (bits & B_FIRST) | B_START | B_NORND;
                        W.hs(HL::X) = s;
                        // yy = v: the same storage
                        W.hs(HL::H) = copysign(fmax(fabs(sout - s), fouru * fabs(s)), sout - s);
                        W.hi(HL::I_NS) = 0; W.hi(HL::I_K) = 0; W.hi(HL::I_KOLD) = 0; W.hs(HL::HOLD) = 0.0;
                        de_top = true;
                    }
                }
            }
            if (de_top && !stop) {   // top of de's loop (ode_RAYS.f90:509-562)
                const double x = W.hs(HL::X);
                if (W.hs(HL::ABSDEL) <= fabs(x - W.hs(HL::T0))) {   // past the output point: the segment ends (the interpolation waits for its batch)
                    bits |= B_INTRP;
                    next = K_CHECK;
                } else if (maxnum <= W.hi(HL::I_NOSTEP)) {             // iflag = 4 / 5: error return, the ray stops with y = yy, t = x
                    flag = (bits & B_STIFF) ? RAYS_STOP_SG_STIFF : RAYS_STOP_SG_MAXNUM;
                    W.f(L::S_) = x;      // v = yy: the same storage
                    stop = true;
                } else {
                    const double h = W.hs(HL::H);
                    W.hs(HL::H) = copysign(fmin(fabs(h), fabs(W.hs(HL::TEND) - x)), h);
                    const double releps = W.hs(HL::RELEPS), abseps = W.hs(HL::ABSEPS);
#pragma unroll (SgUnroll<NV>::L)
                    for (int l = 0; l < NV; ++l) if (l < nv) W.wt(l) = releps * fabs(W.yy(l)) + abseps;
                    if (sg2_block0<NV>(nv, W, eps)) crashed = true;   // (a start derivative evaluated above is dropped: the restart evaluates it again)
                    else if (bits & B_START) {
                        if (have_f1) {   // the start derivative was evaluated together with check_save at this very point
                            my_rhs += 1;
                            if (f1_code) { flag = f1_code; W.f(L::SOUT) = W.f(L::S_); stop = true; }
                            else { sg2_after_start<NV>(nv, W, eps, bits, ff); next = K_PRED; }
                        } else next = K_START;
                    } else next = K_PRED;
                }
            }
            if (crashed && !stop) {   // iflag = 3: tolerances raised (ode_RAYS.f90:566-575), then SG_ode's test (SG_ode_m.f90:138-149)
                const double rel_err = eps * W.hs(HL::RELEPS), abs_err = eps * W.hs(HL::ABSEPS);
                W.f(L::REL) = rel_err; W.f(L::ABS) = abs_err;
                W.f(L::S_) = W.hs(HL::X);        // v = yy: the same storage
                const double total_error = fabs(rel_err) + fabs(abs_err);
                if (total_error > c.SG_error_limit) { flag = RAYS_STOP_ODE_TOTAL_ERROR; stop = true; }
                else next = K_BEGIN;           // SG_ode loops: ode again from the current s to sout
            }
            W.hs(HL::EPS) = eps;
            W.hi(HL::I_BITS) = bits;
            if (stop) {
                const long long iray = (long long)W.f(L::IRAY);
                if (flag == INT_MIN) flag = W.i(L::FLAG);
                a.stop_code[iray] = flag;
                const int nstep = W.i(L::NSTEP);
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
                    if (a.initial_ray_power) a.initial_ray_power[iray] = W.f(L::PWR);
                    if (a.end_residuals) a.end_residuals[iray] = nstep >= 1 ? W.f(L::RPREV) : 0.0;
                    if (a.max_residuals) a.max_residuals[iray] = nstep >= 1 ? W.f(L::RMAX) : -DBL_MAX;
                    if (a.end_ray_parameter) a.end_ray_parameter[iray] = W.v(6);
                    if (a.end_ray_vec) for (int i = 0; i < nv; ++i) a.end_ray_vec[(size_t)iray * nv + i] = W.v(i);
                }
                W.i(L::FINNP) = did_not_start ? 1 : nstep + 1 - W.i(L::P0);
                next = K_FIN;
            } else if (flag != INT_MIN) W.i(L::FLAG) = flag;
            s_kind[slot] = (unsigned char)next;
        }
        // ---- the ring of the next macro-step: predictor / corrector of a slot at order > 3 (or with propagated-roundoff
        // control on) go to the rings of the general code
        int nq = -1;
        if (slot >= 0) {
            if (next == K_PRED || next == K_CORR) {
                const int *hi_ints = reinterpret_cast<const int *>(hot + (size_t)slot * HL::NDBL + HL::INTS);
                const bool fast = T::GENERIC || !RAYS_SG_FAST || (hi_ints[HL::I_K] <= kSgKM && (hi_ints[HL::I_BITS] & B_NORND));   // (one code path: no second ring)
                nq = next == K_PRED ? (fast ? Q_PRED : Q_GPRED) : (fast ? Q_CORR : Q_GCORR);
            } else if (next == K_CHECK) nq = Q_CHECK;
            else if (next == K_START || next == K_BEGIN) nq = Q_START;
            else nq = streaming ? Q_FIN : Q_IDLE;      // K_FIN: nothing to copy out unless the rows are staged
            if (next == K_FIN && !streaming) s_kind[slot] = K_IDLE;
        }
        push(nq, slot);
    }
    dep_end(a, binning);
    unsigned long long stt = my_steps, rh = my_rhs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { stt += __shfl_down_sync(0xffffffffu, stt, o); rh += __shfl_down_sync(0xffffffffu, rh, o); }
    if (lane == 0) { atomicAdd(a.counters, stt); atomicAdd(a.counters + 1, rh); }
}

}  // namespace rays_dev
