// trace_tu.cuh — interface between the C-ABI layer (rays_b200.cu) and the per-equilibrium translation
// units that hold the kernel specialisations (trace_tu.cu compiled once per (equilibrium, stepper)
// with -DRAYS_TU_EQ=<1..4> -DRAYS_TU_ODE=<1,2>; each has its own __constant__ copy of the config).
#pragma once
#include <cuda_runtime.h>

#include "ray_trace.cuh"

namespace rays_dev {

struct KernelSel {
    int ray_deriv;        // RAYS_DERIV_*
    bool generic;         // run-time species count / nv layout
    bool damp, grads;     // used when !generic
    bool sg_lanes;        // Shampine-Gordon: the per-lane state machine of round 1 (measurement aid) instead of the slot machine
};

struct FanLaunchArgs {    // launch-fan kernels (ray_init modules), see trace_tu.cu
    int kind;             // 1 slab (ny,nz grid), 2 solovev, 3 axisym (ntheta,nphi grid), 4 positions+directions
    long long ncand;      // candidates = positions x n_a x n_b (kind 4: = positions)
    const double *rvec_in;   // device: launch positions [npos][3] (kind 4: [ncand][3])
    const double *nvec_in;   // device: kind 4 directions [ncand][3]
    int n_a, n_b;         // inner grids, b fastest: (n_ky, n_kz) or (n_rindex_theta, n_rindex_phi)
    double a0, da, b0, db;
    double *rvec_out, *nvec_out;       // [ncand][3] candidate results (device)
    int *valid;                        // [ncand]
};

struct TuOps {
    cudaError_t (*upload)(const DevCfg *, cudaStream_t);
    // occupancy query + launch of the trace kernel selected by sel; grid < 0 -> only report
    // sg_state_bytes_per_cta: global slot records the Shampine-Gordon slot-machine kernel needs per CTA (0 for the other kernels);
    // rays_per_cta: rays one CTA holds in flight (its threads; the slot machine: its slots, as many as shared memory holds)
    cudaError_t (*trace)(const KernelSel &, const TraceArgs &, int grid, cudaStream_t, int *blocks_per_sm, const char **name, size_t *sg_state_bytes_per_cta,
                         int *rays_per_cta);
    cudaError_t (*probe_equilibrium)(const KernelSel &, long long, const double *, double *, int *, cudaStream_t);
    cudaError_t (*probe_rhs)(const KernelSel &, long long, const double *, double *, int *, cudaStream_t);
    cudaError_t (*probe_check_save)(const KernelSel &, long long, const double *, double *, int *, cudaStream_t);
    cudaError_t (*launch_fan)(const KernelSel &, const FanLaunchArgs &, cudaStream_t);
};

// defined by the translation units; index [equilib_model][ode_solver]; entries may be null for
// (eq, ode) pairs whose TU only carries the trace kernel (probes/launch fans live in the RK4 TUs)
const TuOps *tu_ops(int equilib_model, int ode_solver);

}  // namespace rays_dev
