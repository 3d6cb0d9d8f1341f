// trace_tu.cu — kernel specialisations for ONE (equilibrium model, ODE stepper) pair.
// Compiled 8 times: -DRAYS_TU_EQ=<1..4> -DRAYS_TU_ODE=<1,2>, always with -fmad=false (see the parity
// contract in ray_physics.cuh).  The RK4 translation units also carry the one-point probes and the
// launch-fan kernels of their equilibrium.
#include "trace_tu.cuh"
#if RAYS_TU_ODE == 2
#include "ray_trace_sg2.cuh"
#endif

#ifndef RAYS_TU_EQ
#error "compile with -DRAYS_TU_EQ=<1..4> -DRAYS_TU_ODE=<1,2>"
#endif

namespace rays_dev {
namespace {

constexpr int kEQ = RAYS_TU_EQ;
[[maybe_unused]] constexpr int kODE = RAYS_TU_ODE;

cudaError_t tu_upload(const DevCfg *dc, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(g_dc, dc, sizeof(DevCfg), 0, cudaMemcpyHostToDevice, st);
}

#if RAYS_TU_ODE == 2
// slots of one slot-machine CTA: what its share of the SM's shared memory holds besides the deposition bins and the rings
template <class T> cudaError_t sg2_geometry(const TraceArgs &a, int *slots, size_t *dyn_bytes) {
    static int static_bytes = -1;       // per specialisation: the kernel's static shared memory
    if (static_bytes < 0) {
        cudaFuncAttributes fa{};
        cudaError_t e = cudaFuncGetAttributes(&fa, trace_sg2_kernel<T>);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(trace_sg2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return e;
        static_bytes = (int)fa.sharedSizeBytes;
    }
    const size_t per_cta = (size_t)233472 / kSgCtas - 1024 - (size_t)static_bytes;     // 228 KB per SM, 1 KB reserved per CTA
    const size_t fixed = ((size_t)(a.dep_smem + 15) / 16) * 16 + sg2_ring_bytes();
    const size_t hot = sg2_hot_bytes_per_slot<T::NV>();
    long long s = per_cta > fixed ? (long long)((per_cta - fixed) / hot) : 0;
    if (s > kSgRingCap) s = kSgRingCap;
    if (const char *env = getenv("RAYS_B200_SG_SLOTS")) { const int v = atoi(env); if (v > 0 && v < s) s = v; }   // measurement aid
    if (s < 32) return cudaErrorInvalidConfiguration;
    *slots = (int)s;
    *dyn_bytes = fixed + (size_t)s * hot;
    return cudaSuccess;
}
#endif

template <class T> cudaError_t launch_trace(const KernelSel &s, const TraceArgs &a_in, int grid, cudaStream_t st, int *bps, const char **name, size_t *sgb, int *rpc, const char *nm) {
    if (name) *name = nm;
    if (sgb) *sgb = 0;
#if RAYS_TU_ODE == 1
    if (rpc) *rpc = Rk4Block<T>::value;
#else
    if (rpc) *rpc = kTraceBlock;
#endif
    TraceArgs a = a_in;
#if RAYS_TU_ODE == 2
    size_t dyn = a.dep_smem;
    if (!s.sg_lanes) {
        int slots = 0;
        cudaError_t e = sg2_geometry<T>(a, &slots, &dyn);
        if (e != cudaSuccess) return e;
        a.sg_slots = slots;
        a.sg_mixed = T::DERIV == RAYS_DERIV_NUM ? 1 : 0;
        if (const char *env = getenv("RAYS_B200_SG_MIXED")) { if (env[0]) a.sg_mixed = atoi(env); }   // measurement aid: 0, 1, 2
        if (sgb) *sgb = sg2_state_bytes_per_cta<T::NV>();
        if (rpc) *rpc = slots;
    }
#endif
    if (bps) {
#if RAYS_TU_ODE == 1
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, trace_rk4_kernel<T>, Rk4Block<T>::value, 0);
#else
        cudaError_t e = s.sg_lanes ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, trace_sg_kernel<T>, kTraceBlock, 0)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, trace_sg2_kernel<T>, kSgBlock, dyn);
#endif
        if (e != cudaSuccess) return e;
    }
    if (grid <= 0) return cudaSuccess;
#if RAYS_TU_ODE == 1
    (void)s;
    trace_rk4_kernel<T><<<grid, Rk4Block<T>::value, a.dep_smem, st>>>(a);
#else
    if (s.sg_lanes) trace_sg_kernel<T><<<grid, kTraceBlock, a.dep_smem, st>>>(a);
    else trace_sg2_kernel<T><<<grid, kSgBlock, dyn, st>>>(a);
#endif
    return cudaGetLastError();
}

#define RAYS_SEL(DER, DMP, GRD, NM)                                                                            \
    if (s.ray_deriv == DER && !s.generic && s.damp == (DMP == 1) && s.grads == (GRD == 1))                     \
        return launch_trace<Traits<kEQ, 2, DER, DMP, GRD>>(s, a, grid, st, bps, name, sgb, rpc, NM);
#define RAYS_GEN(DER, NM)                                                                                      \
    if (s.ray_deriv == DER && s.generic) return launch_trace<Traits<kEQ, 0, DER, -1, -1>>(s, a, grid, st, bps, name, sgb, rpc, NM);

cudaError_t tu_trace(const KernelSel &s, const TraceArgs &a, int grid, cudaStream_t st, int *bps, const char **name, size_t *sgb, int *rpc) {
    RAYS_SEL(RAYS_DERIV_COLD, 0, 0, "trace<ns2,cold,nv7>")
    RAYS_SEL(RAYS_DERIV_COLD, 1, 0, "trace<ns2,cold,damp,nv8>")
    RAYS_SEL(RAYS_DERIV_COLD, 0, 1, "trace<ns2,cold,grads,nv12>")
    RAYS_SEL(RAYS_DERIV_COLD, 1, 1, "trace<ns2,cold,damp,grads,nv13>")
    RAYS_SEL(RAYS_DERIV_NUM, 0, 0, "trace<ns2,num,nv7>")
    RAYS_SEL(RAYS_DERIV_NUM, 1, 0, "trace<ns2,num,damp,nv8>")
    RAYS_SEL(RAYS_DERIV_NUM, 0, 1, "trace<ns2,num,grads,nv12>")
    RAYS_SEL(RAYS_DERIV_NUM, 1, 1, "trace<ns2,num,damp,grads,nv13>")
    RAYS_GEN(RAYS_DERIV_COLD, "trace<generic,cold>")
    RAYS_GEN(RAYS_DERIV_NUM, "trace<generic,num>")
    return cudaErrorInvalidValue;
}

#if RAYS_TU_ODE == 1
// ---- probes: always through the generic traits (run-time nv layout), species fixed to 2 when possible
template <class T> cudaError_t run_probe_eq(long long n, const double *r, double *o, int *e, cudaStream_t st) {
    probe_equilibrium_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, r, o, e);
    return cudaGetLastError();
}
cudaError_t tu_probe_equilibrium(const KernelSel &s, long long n, const double *r, double *o, int *e, cudaStream_t st) {
    if (s.generic) return run_probe_eq<Traits<kEQ, 0, RAYS_DERIV_COLD, -1, -1>>(n, r, o, e, st);
    return run_probe_eq<Traits<kEQ, 2, RAYS_DERIV_COLD, 0, 0>>(n, r, o, e, st);
}
template <class T> cudaError_t run_probe_rhs(long long n, const double *v, double *d, int *e, cudaStream_t st) {
    probe_rhs_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, v, d, e);
    return cudaGetLastError();
}
#define RAYS_PSEL(FN, DER, DMP, GRD)                                                                           \
    if (s.ray_deriv == DER && !s.generic && s.damp == (DMP == 1) && s.grads == (GRD == 1))                     \
        return FN<Traits<kEQ, 2, DER, DMP, GRD>>(n, v, d, e, st);
cudaError_t tu_probe_rhs(const KernelSel &s, long long n, const double *v, double *d, int *e, cudaStream_t st) {
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_COLD, 0, 0)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_COLD, 1, 0)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_COLD, 0, 1)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_COLD, 1, 1)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_NUM, 0, 0)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_NUM, 1, 0)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_NUM, 0, 1)
    RAYS_PSEL(run_probe_rhs, RAYS_DERIV_NUM, 1, 1)
    if (s.ray_deriv == RAYS_DERIV_COLD) return run_probe_rhs<Traits<kEQ, 0, RAYS_DERIV_COLD, -1, -1>>(n, v, d, e, st);
    return run_probe_rhs<Traits<kEQ, 0, RAYS_DERIV_NUM, -1, -1>>(n, v, d, e, st);
}
template <class T> cudaError_t run_probe_cs(long long n, const double *v, double *d, int *e, cudaStream_t st) {
    probe_check_save_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, v, d, e);
    return cudaGetLastError();
}
cudaError_t tu_probe_check_save(const KernelSel &s, long long n, const double *v, double *d, int *e, cudaStream_t st) {
    RAYS_PSEL(run_probe_cs, RAYS_DERIV_COLD, 0, 0)
    RAYS_PSEL(run_probe_cs, RAYS_DERIV_COLD, 1, 0)
    RAYS_PSEL(run_probe_cs, RAYS_DERIV_COLD, 0, 1)
    RAYS_PSEL(run_probe_cs, RAYS_DERIV_COLD, 1, 1)
    if (!s.generic) {  // check_save never uses deriv_num: reuse the cold specialisations
        KernelSel c = s; c.ray_deriv = RAYS_DERIV_COLD;
        return tu_probe_check_save(c, n, v, d, e, st);
    }
    return run_probe_cs<Traits<kEQ, 0, RAYS_DERIV_COLD, -1, -1>>(n, v, d, e, st);
}

// ---- launch fans (row f1): one thread = one candidate ray ----------------------------------------------
// simple_slab_ray_init (simple_slab_ray_init_m.f90:58-185), ray_init_solovev_nphi_ntheta
// (solovev_ray_init_nphi_ntheta_m.f90:61-209), ray_init_axisym_toroid_R_Z_nphi_ntheta
// (axisym_toroid_ray_init_R_Z_nphi_ntheta_m.f90:67-244), ray_init_XYZ_k_direction
// (one_ray_init_XYZ_k_direction_m.f90:131-180) as file_input_ray_init drives it.
// Launch POSITIONS come from the host (they involve cos/sin of a handful of angles, evaluated there
// with the host libm like the reference); the per-candidate equilibrium + dispersion root runs here.
template <int NS_> __global__ void launch_fan_kernel(const FanLaunchArgs f) {
    constexpr int NSM = NSpec<NS_>::MAX;
    const int ns = NSpec<NS_>::n();
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= f.ncand) return;
    int ok = 0;
    double rvec[3] = {0.0, 0.0, 0.0}, nout[3] = {0.0, 0.0, 0.0};
    if (f.kind == 4) {
        for (int i = 0; i < 3; ++i) { rvec[i] = f.rvec_in[3 * idx + i]; nout[i] = f.nvec_in[3 * idx + i]; }
        Eq<NSM> e;
        equilibrium<kEQ, NS_, false>(rvec[0], rvec[1], rvec[2], e);
        if (!e.err) {
            const double nn = sqrt(nout[0] * nout[0] + nout[1] * nout[1] + nout[2] * nout[2]);
            for (int i = 0; i < 3; ++i) nout[i] = nout[i] / nn;
            const double cos_theta = e.bunit[0] * nout[0] + e.bunit[1] * nout[1] + e.bunit[2] * nout[2];
            const double theta = acos(cos_theta);
            double n;
            if (solve_n_vs_theta<NSM>(e, ns, theta, n)) {
                for (int i = 0; i < 3; ++i) nout[i] = n * nout[i];
                ok = 1;
            }
        }
    } else {
        const long long per_pos = (long long)f.n_a * f.n_b;
        const long long ipos = idx / per_pos;
        const int rem = (int)(idx - ipos * per_pos);
        const int ia = rem / f.n_b, ib = rem - ia * f.n_b;
        for (int i = 0; i < 3; ++i) rvec[i] = f.rvec_in[3 * ipos + i];
        const double va = f.a0 + (double)ia * f.da;   // rindex_y / rindex_theta
        const double vb = f.b0 + (double)ib * f.db;   // rindex_z / rindex_phi
        Eq<NSM> e;
        equilibrium<kEQ, NS_, false>(rvec[0], rvec[1], rvec[2], e);
        if (!e.err) {
            if (f.kind == 1) {   // solve_nx_vs_ny_nz_by_bz (dispersion_solvers_m.f90:116-153)
                const double n2 = va * e.bunit[2] - vb * e.bunit[1];
                const double n3 = va * e.bunit[1] + vb * e.bunit[2];
                const cplx nx = solve_n1_vs_n2_n3<NSM>(e, ns, n2, n3);
                if (nx.im == 0.0) { nout[0] = nx.re; nout[1] = va; nout[2] = vb; ok = 1; }
            } else {
                // grad(psi) = (x*bz, y*bz, -R*br) (solovev_eq_m.f90:312-316)
                const DevCfg &d = g_dc;
                const double x = rvec[0], y = rvec[1], z = rvec[2];
                const double r = sqrt(x * x + y * y);
                const double br = -d.sv_bp0 * r * z / d.sv_rk2;
                const double zk = z / d.sv_rk, rr = r / d.sv_rmaj;
                const double bz = d.sv_bp0 * ((zk * zk) + .5 * ((rr * rr) - 1.0));
                const double gp[3] = {x * bz, y * bz, -r * br};
                const double gn = sqrt(gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]);
                const double psi_unit[3] = {gp[0] / gn, gp[1] / gn, gp[2] / gn};
                double theta_unit[3] = {-gp[2], 0.0, gp[0]};
                const double tn = sqrt(theta_unit[0] * theta_unit[0] + theta_unit[1] * theta_unit[1] + theta_unit[2] * theta_unit[2]);
                for (int i = 0; i < 3; ++i) theta_unit[i] = theta_unit[i] / tn;
                const double trans[3] = {e.bunit[1] * psi_unit[2] - e.bunit[2] * psi_unit[1],
                                         e.bunit[2] * psi_unit[0] - e.bunit[0] * psi_unit[2],
                                         e.bunit[0] * psi_unit[1] - e.bunit[1] * psi_unit[0]};
                const double phi_unit[3] = {0.0, 1.0, 0.0};
                double rindex_vec[3];
                for (int i = 0; i < 3; ++i) rindex_vec[i] = vb * phi_unit[i] + va * theta_unit[i];
                const double n3 = e.bunit[0] * rindex_vec[0] + e.bunit[1] * rindex_vec[1] + e.bunit[2] * rindex_vec[2];
                const double n2 = trans[0] * rindex_vec[0] + trans[1] * rindex_vec[1] + trans[2] * rindex_vec[2];
                const cplx npsi = solve_n1_vs_n2_n3<NSM>(e, ns, n2, n3);
                const bool evanescent = f.kind == 2 ? (npsi.im != 0.0) : (fabs(npsi.im) > 10.0 * DBL_MIN);
                if (!evanescent) {
                    for (int i = 0; i < 3; ++i) nout[i] = rindex_vec[i] - npsi.re * psi_unit[i];
                    ok = 1;
                }
            }
        }
    }
    f.valid[idx] = ok;
    for (int i = 0; i < 3; ++i) { f.rvec_out[3 * idx + i] = rvec[i]; f.nvec_out[3 * idx + i] = nout[i]; }
}
cudaError_t tu_launch_fan(const KernelSel &s, const FanLaunchArgs &f, cudaStream_t st) {
    if (f.ncand <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((f.ncand + 127) / 128);
    if (s.generic) launch_fan_kernel<0><<<grid, 128, 0, st>>>(f);
    else launch_fan_kernel<2><<<grid, 128, 0, st>>>(f);
    return cudaGetLastError();
}
#endif  // RAYS_TU_ODE == 1

}  // namespace

#define RAYS_CAT3(a, b, c) a##b##_##c
#define RAYS_TU_NAME(eq, ode) RAYS_CAT3(rays_tu_ops_, eq, ode)
extern const TuOps RAYS_TU_NAME(RAYS_TU_EQ, RAYS_TU_ODE) = {
    tu_upload, tu_trace,
#if RAYS_TU_ODE == 1
    tu_probe_equilibrium, tu_probe_rhs, tu_probe_check_save, tu_launch_fan
#else
    nullptr, nullptr, nullptr, nullptr
#endif
};

}  // namespace rays_dev
