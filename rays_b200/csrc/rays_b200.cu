// rays_b200.cu — the C ABI of include/rays_b200.h: device/stream life cycle, upload of the marshalled
// module state and spline tables, fan and result buffers in HBM, kernel selection and launch, batched
// trajectory copy-out in the reference's ray_results_m layout, launch-fan compaction, deposition
// binning and the measurement helpers.  No CPU fallback: every entry point fails with RAYS_ERR_CUDA
// when no GPU is present.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "trace_tu.cuh"
#include "coil_field.cuh"

namespace rays_dev {
#define RAYS_TU_DECL(eq, ode) extern const TuOps rays_tu_ops_##eq##_##ode;
RAYS_TU_DECL(1, 1) RAYS_TU_DECL(1, 2) RAYS_TU_DECL(2, 1) RAYS_TU_DECL(2, 2)
RAYS_TU_DECL(3, 1) RAYS_TU_DECL(3, 2) RAYS_TU_DECL(4, 1) RAYS_TU_DECL(4, 2)
const TuOps *tu_ops(int eq, int ode) {
    static const TuOps *tab[5][3] = {{nullptr, nullptr, nullptr},
                                     {nullptr, &rays_tu_ops_1_1, &rays_tu_ops_1_2},
                                     {nullptr, &rays_tu_ops_2_1, &rays_tu_ops_2_2},
                                     {nullptr, &rays_tu_ops_3_1, &rays_tu_ops_3_2},
                                     {nullptr, &rays_tu_ops_4_1, &rays_tu_ops_4_2}};
    if (eq < 1 || eq > 4 || ode < 1 || ode > 2) return nullptr;
    return tab[eq][ode];
}
}  // namespace rays_dev

using namespace rays_dev;

namespace {

thread_local std::string g_err;   // per host thread: rays_b200_trace_multi drives one GPU per thread
int set_err(int code, const std::string &m) { g_err = m; return code; }
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return set_err(RAYS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

template <class Tp> struct DevBuf {
    Tp *p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(Tp));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Ctx {
    bool inited = false;
    int device = -1, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_batch[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    // config
    bool cfg_set = false;
    DevCfg dc{};
    KernelSel sel{};
    DevBuf<double> zx, zf, rgrid, zgrid, br, bz, aphi;
    DevBuf<double> prof_grid[3], prof_fspl[3];   // axisym 1-D profile splines: ne, Te, Ti
    DevBuf<double> eq_rgrid, eq_zgrid, eq_psi, eq_T;   // eqdsk magnetics: Psi_profile, T_profile
    // fan (device)
    long long nray = 0;
    DevBuf<double> rvec0, nvec0, wt;
    // results (device)
    long long res_nray = 0;
    int res_nv = 0, res_npa = 0;
    bool have_traj = false;
    DevBuf<double> ray_vec, residual, pwr, endres, maxres, endpar, startv, endv;
    DevBuf<int> npoints, stop;
    DevBuf<unsigned long long> queue;   // [0] queue, [1] ray-steps, [2] RHS evaluations
    DevBuf<int> done_list;              // copier kernel (copy_out_kernel): rays in the order they ended
    DevBuf<unsigned long long> copy_ctl;   // [0] done count, [1] claim counter, [2] abort flag (int), [3] deferred flag (int), [4] trace-started flag (int)
    bool copier_on = false;             // the trace launches of this call feed done_list
    int last_copier = 0;                // the last rays_b200_trace delivered its trajectories through the copier kernel
    std::string kernel_name_buf;        // "<trace kernel> + copy_out_kernel" for rays_b200_last_trace_info
    int *copier_started = nullptr;      // page-locked, device-visible: CTAs of the copier kernel that are resident
    // time slicing: suspended-ray records and the two resume lists
    DevBuf<double> cont_state;
    DevBuf<int> cont_list[2];
    DevBuf<double> sg_state;       // slot memory of the Shampine-Gordon slot-machine kernel
    size_t cached_sgb = 0;
    int cached_rpc = 0, cached_dep_smem = -1;   // rays in flight per CTA of the selected kernel (for the deposition-bin size it was queried with)
    double last_first_ms = 0, last_resume_ms = 0;
    int last_phases = 0;
    cudaEvent_t ev_m0 = nullptr, ev_m1 = nullptr;
    bool main_pending = false, main_is_resume = false;
    int cached_bps = 0;
    const char *cached_name = "";
    const void *cached_ops = nullptr;
    // deposition
    DevBuf<unsigned long long> dep;   // fixed-point bins (see TraceArgs::dep_acc)
    int dep_bins = 0;
    double dep_min = 0, dep_max = 0;
    bool dep_fused = false;
    double fan_weight = 0.0;          // sum |ray_pwr_wt| over the fan as uploaded / launched (kept by fan_shard): sets the bin unit
    double dep_scale = 1.0;           // units per watt-fraction: 2^(62 - e), 2^e > fan_weight
    // stats
    double last_ms = 0;
    long long last_steps = 0, last_rhs = 0;
    int last_launches = 0;
    const char *last_kernel = "";
    int last_grid = 0, last_bps = 0;
    // pinned staging for small results
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // staged copy-out to PAGEABLE result arrays: packed rows HBM -> pinned ring (copy engine) -> caller's arrays (host threads)
    static constexpr int kRing = 3;
    DevBuf<double> pack_dev;          // kRing packed chunks
    DevBuf<long long> pack_off;       // packed offset (doubles) of every ray of the batch
    double *ring_host = nullptr;      // kRing chunks, page-locked
    size_t ring_chunk_doubles = 0;
    cudaEvent_t ev_ring[kRing] = {nullptr, nullptr, nullptr};
};
// One context per GPU.  Every entry point works on the CURRENT context of the calling host thread: context 0 unless
// rays_b200_*_multi selected another one (each of its worker threads drives one GPU).
constexpr int kMaxGpus = 16;
Ctx g_ctx[kMaxGpus];
thread_local Ctx *g_cur = &g_ctx[0];
inline Ctx &cx() { return *g_cur; }

int need_init() { return cx().inited ? 0 : set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_init has not been called"); }

const char *kStopStrings[RAYS_STOP_CODE_MAX] = {
    "", "sout > s_max", " nstep > nstep_max", "infinite Vg", "ray stalled", "dispersion_residual",
    "infinite_Vg", "total_absorption", "ODE total error", "step number .ge. maxnum", "equations stiff",
    "t == tout", "relerr or abserr < 0", "eps <= 0", "", "", "", "", "", "",
    "x out_of_bounds", "y out_of_bounds", "z out_of_bounds", "R out_of_box", "z out_of_box",
    "R_out_of_box", "Z_out_of_box", "out_of_plasma", "R out_of_bounds", "z out_of_bounds",
    "negative_dens", "negative_temp"};

int upload_table(DevBuf<double> &b, const double *src, size_t n) {
    CK(b.reserve(n));
    CK(cudaMemcpyAsync(b.p, src, n * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
    return 0;
}

int validate_cfg(const rays_cfg &c) {
    if (c.equilib_model < RAYS_EQ_SLAB || c.equilib_model > RAYS_EQ_MULTIPLE_MIRROR) return set_err(RAYS_ERR_INVALID_CONFIG, "equilibrium: improper equilib_model");
    if (c.ode_solver != RAYS_ODE_RK4 && c.ode_solver != RAYS_ODE_SG) return set_err(RAYS_ERR_INVALID_CONFIG, "ode_solver: invalid ode solver");
    if (c.ray_deriv != RAYS_DERIV_COLD && c.ray_deriv != RAYS_DERIV_NUM) return set_err(RAYS_ERR_INVALID_CONFIG, "EQN_RAY: invalid value, ray_deriv_name");
    if (c.ray_param != RAYS_PARAM_ARCL && c.ray_param != RAYS_PARAM_TIME) return set_err(RAYS_ERR_INVALID_CONFIG, "EQN_RAY: invalid ray parameter");
    if (c.nspec < 0 || c.nspec > RAYS_NSPEC0) return set_err(RAYS_ERR_INVALID_CONFIG, "species: nspec out of range");
    if (c.damping_model != RAYS_DAMP_NONE && c.damping_model != RAYS_DAMP_FUND_ECH) return set_err(RAYS_ERR_INVALID_CONFIG, "damping: Unimplemented damping model");
    int nv = 7;
    if (c.damping_model != RAYS_DAMP_NONE) { nv += 1; if (c.multi_spec_damping) nv += 1 + c.nspec; }
    if (c.integrate_eq_gradients) nv += 5;
    if (c.nv != nv) return set_err(RAYS_ERR_INVALID_CONFIG, "ode_m: nv does not match damping/gradient options (ode_m.f90:160-173)");
    if (c.nstep_max < 0) return set_err(RAYS_ERR_INVALID_CONFIG, "ode_m: nstep_max < 0");
    if (c.ode_solver == RAYS_ODE_SG && (c.rel_err0 < (double)1.e-10f || c.abs_err0 < (double)1.e-10f))   // SG_ode_m.f90:63-66
        return set_err(RAYS_ERR_INVALID_CONFIG, "initialize_SG_ode: rel_err0, abs_err0 too small");
    if (c.damping_model == RAYS_DAMP_FUND_ECH && (!c.zfun_re.x_grid || !c.zfun_re.fspl || c.zfun_re.nx < 2))
        return set_err(RAYS_ERR_INVALID_CONFIG, "damping: Z-function spline table missing");
    if (c.equilib_model == RAYS_EQ_MULTIPLE_MIRROR) {
        const rays_mirror_eq &m = c.mirror;
        if (!m.Br_spline.fspl || !m.Bz_spline.fspl || !m.Aphi_spline.fspl || !m.Br_spline.x_grid || !m.Br_spline.y_grid)
            return set_err(RAYS_ERR_INVALID_CONFIG, "multiple_mirror: spline tables missing");
        if (m.Br_spline.nx < 2 || m.Br_spline.ny < 2) return set_err(RAYS_ERR_INVALID_CONFIG, "multiple_mirror: spline grids need nx, ny >= 2");
        if (m.Bz_spline.nx != m.Br_spline.nx || m.Aphi_spline.nx != m.Br_spline.nx || m.Bz_spline.ny != m.Br_spline.ny || m.Aphi_spline.ny != m.Br_spline.ny)
            return set_err(RAYS_ERR_INVALID_CONFIG, "multiple_mirror: Br, Bz, Aphi must share one (r,z) grid");
    }
    if (c.equilib_model == RAYS_EQ_AXISYM_TOROID) {
        const rays_axisym_eq &a = c.axisym;
        auto ok = [](const rays_spline1d &s) { return s.x_grid && s.fspl && s.nx >= 2; };
        if (a.magnetics_model != RAYS_MAG_SOLOVEV && a.magnetics_model != RAYS_MAG_EQDSK_SPLINE)
            return set_err(RAYS_ERR_INVALID_CONFIG, "initialize_axisym_toroid_eq: unknown magnetics model");
        if (a.magnetics_model == RAYS_MAG_EQDSK_SPLINE &&
            (!a.Psi_spline.fspl || !a.Psi_spline.x_grid || !a.Psi_spline.y_grid || a.Psi_spline.nx < 2 || a.Psi_spline.ny < 2 || !ok(a.T_spline) || a.eq_psibound == 0.0))
            return set_err(RAYS_ERR_INVALID_CONFIG, "axisym_toroid: eqdsk_magnetics_spline_interp tables missing");
        if (a.density_prof_model == RAYS_PROF_SPLINE && !ok(a.ne_spline))
            return set_err(RAYS_ERR_INVALID_CONFIG, "axisym_toroid: density_spline_interp table missing");
        for (int s = 0; s <= c.nspec; ++s)
            if (a.temperature_prof_model[s] == RAYS_PROF_SPLINE && !ok(s == 0 ? a.Te_spline : a.Ti_spline))
                return set_err(RAYS_ERR_INVALID_CONFIG, "axisym_toroid: temperature_spline_interp table missing");
    }
    return 0;
}

// ---- small utility kernels -------------------------------------------------------------------------------
__global__ void fill_kernel(double *p, long long n, double v) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// order-preserving compaction of the launch-fan candidates: block-level exclusive scan of the
// validity flags (one pass over per-block counts on a single block, then scatter)
constexpr int kScanBlock = 256;
__global__ void fan_count_kernel(const int *valid, long long n, int *block_counts) {
    __shared__ int sh[kScanBlock / 32];
    const long long i = blockIdx.x * (long long)kScanBlock + threadIdx.x;
    const int f = (i < n) ? valid[i] : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f != 0);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kScanBlock / 32; ++w) s += sh[w];
        block_counts[blockIdx.x] = s;
    }
}
__global__ void fan_offsets_kernel(int *block_counts, int nblocks, long long *total) {
    // single thread block, sequential over chunks: nblocks <= ~ 33k for 8M candidates
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int c = i < nblocks ? block_counts[i] : 0;
        // inclusive scan in the block via warp shuffles
        int x = c;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        __shared__ int wsum[32];
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int wv = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, wv, o); if (threadIdx.x >= o) wv += y; }
            wsum[threadIdx.x] = wv;
        }
        __syncthreads();
        const int woff = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
        const long long excl = carry + woff + x - c;
        if (i < nblocks) block_counts[i] = (int)excl;   // fans are < 2^31 rays per GPU
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void fan_scatter_kernel(const int *valid, long long n, const int *block_offsets, const double *rv, const double *nv,
                                   double *rvec0, double *nvec0) {
    __shared__ int sh[kScanBlock / 32];
    const long long i = blockIdx.x * (long long)kScanBlock + threadIdx.x;
    const int f = (i < n) ? valid[i] : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f != 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = __popc(b);
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += sh[k];
    if (f) {
        const long long dst = (long long)block_offsets[blockIdx.x] + woff + __popc(b & ((1u << lane) - 1u));
        for (int k = 0; k < 3; ++k) { rvec0[3 * dst + k] = rv[3 * i + k]; nvec0[3 * dst + k] = nv[3 * i + k]; }
    }
}
// The trimmed copy-out moves whole 32-point groups: clear the points between npoints and the next multiple of 32
// so that the caller's arrays receive zeros there (the reference zero-fills and writes only 1:npoints).
__global__ void zero_tail_kernel(double *ray_vec, double *residual, const int *npoints, int npa, int nv, long long nray) {
    const long long iray = blockIdx.x;
    if (iray >= nray) return;
    // copy width of this ray's 64-ray group (copy_trajectories): longest ray of the group, rounded up to 32 points
    __shared__ int gmax;
    if (threadIdx.x == 0) gmax = 1;
    __syncthreads();
    const long long g0 = iray / 64 * 64;
    if (g0 + threadIdx.x < nray && threadIdx.x < 64) atomicMax(&gmax, npoints[g0 + threadIdx.x]);
    __syncthreads();
    const int np = npoints[iray];
    const int end = min(npa, (gmax + 31) / 32 * 32);
    if (ray_vec) for (int i = np * nv + threadIdx.x; i < end * nv; i += blockDim.x) ray_vec[(size_t)iray * npa * nv + i] = 0.0;
    if (residual) for (int i = np + threadIdx.x; i < end; i += blockDim.x) residual[(size_t)iray * npa + i] = 0.0;
}
// Packed copy-out (pageable result arrays): the saved points of ray i, [npoints*nv of ray_vec][npoints of residual], go to
// out + off[i] - off0.  One CTA per ray, 16-byte accesses where the row allows it.
__global__ void pack_rows_kernel(const double *__restrict__ ray_vec, const double *__restrict__ residual, const int *__restrict__ npoints,
                                 const long long *__restrict__ off, long long off0, int npa, int nv, double *__restrict__ out) {
    const long long i = blockIdx.x;
    const int np = npoints[i];
    double *dst = out + (off[i] - off0);
    if (ray_vec) {
        const double *src = ray_vec + (size_t)i * npa * nv;
        const int n = np * nv;
        if ((((size_t)dst | (size_t)src) & 15) == 0) {
            const double2 *s2 = reinterpret_cast<const double2 *>(src);
            double2 *d2 = reinterpret_cast<double2 *>(dst);
            for (int k = threadIdx.x; k < n / 2; k += blockDim.x) d2[k] = s2[k];
            if ((n & 1) && threadIdx.x == 0) dst[n - 1] = src[n - 1];
        } else
            for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
        dst += n;
    }
    if (residual) {
        const double *src = residual + (size_t)i * npa;
        for (int k = threadIdx.x; k < np; k += blockDim.x) dst[k] = src[k];
    }
}
// keep rays with iray % world == rank, order preserved
__global__ void fan_shard_kernel(long long n_out, int rank, int world, const double *rv, const double *nv, const double *w,
                                 double *rv_o, double *nv_o, double *w_o) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const long long i = j * world + rank;
    for (int k = 0; k < 3; ++k) { rv_o[3 * j + k] = rv[3 * i + k]; nv_o[3 * j + k] = nv[3 * i + k]; }
    w_o[j] = w[i];
}
// per-ray summaries of the last trace as one row of 6 + 2 nv doubles per ray (the fixed-size record SURVEY.md 8e gathers
// across GPUs): npoints, stop code, initial_ray_power, end_residuals, max_residuals, end_ray_parameter, start_ray_vec, end_ray_vec
__global__ void pack_summaries_kernel(long long nray, int nv, const int *npoints, const int *stop, const double *pwr, const double *endres,
                                      const double *maxres, const double *endpar, const double *startv, const double *endv, double *out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nray) return;
    double *o = out + (size_t)i * (6 + 2 * nv);
    o[0] = (double)npoints[i]; o[1] = (double)stop[i]; o[2] = pwr[i]; o[3] = endres[i]; o[4] = maxres[i]; o[5] = endpar[i];
    for (int k = 0; k < nv; ++k) { o[6 + k] = startv[(size_t)i * nv + k]; o[6 + nv + k] = endv[(size_t)i * nv + k]; }
}
// calculate_deposition_profiles on stored trajectories (deposition_profiles_m.f90:228-292): one thread per ray
template <int EQ_> __global__ void deposition_kernel(long long nray, int nv, int npa, const double *ray_vec, const int *npoints,
                                                      const double *pwr, const DepBins bins) {
    const long long iray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (iray >= nray) return;
    const int np = npoints[iray];
    const double *v = ray_vec + (size_t)iray * npa * nv;
    const double P = pwr[iray];
    double xa = dep_abscissa<EQ_>(v), Qa = v[7] * P;
    for (int ip = 1; ip < np; ++ip) {
        const double *vp = v + (size_t)ip * nv;
        const double xb = dep_abscissa<EQ_>(vp), Qb = vp[7] * P;
        bin_segment(bins, xa, xb, Qa, Qb);
        xa = xb; Qa = Qb;
    }
}
// ---- O-X conversion analysis of stored trajectories (post_process_lib/OX_conv_analysis_m.f90:91-407): one thread per ray
struct OxEq { double alpha0, gamma0, ns0, gradns0[3], bunit[3]; };
// equilibrium() as that module uses it: alpha_e, gamma_e, n_e, grad n_e, b.  The Fortran forms the derived quantities whatever
// equib_err says (equilibrium_m.f90:229-268): a point outside the plasma boundary is evaluated, a point outside the box is not
template <int EQ_> __device__ __forceinline__ void ox_equilibrium(const double *r, OxEq &q) {
    Eq<RAYS_NSPECIES> e;
    equilibrium<EQ_, 0, true>(r[0], r[1], r[2], e);
    const bool hard = e.err != 0 && e.err != RAYS_STOP_OUT_OF_PLASMA && e.err != RAYS_STOP_NEGATIVE_DENS && e.err != RAYS_STOP_NEGATIVE_TEMP;
    q.alpha0 = 0.0; q.gamma0 = 0.0; q.ns0 = 0.0;
    for (int i = 0; i < 3; ++i) { q.gradns0[i] = 0.0; q.bunit[i] = 0.0; }
    if (hard) return;
    const DevCfg &d = g_dc;
    const double bmag = sqrt(e.bvec[0] * e.bvec[0] + e.bvec[1] * e.bvec[1] + e.bvec[2] * e.bvec[2]);
    for (int i = 0; i < 3; ++i) { q.bunit[i] = e.bvec[i] / bmag; q.gradns0[i] = e.gradns[i][0]; }
    q.ns0 = e.ns[0];
    const double omgc = d.c.qs[0] * bmag / d.c.ms[0];
    const double omgp2 = e.ns[0] * (d.c.qs[0] * d.c.qs[0]) / (d.c.eps0 * d.c.ms[0]);
    q.alpha0 = omgp2 / (d.c.omgrf * d.c.omgrf);
    q.gamma0 = omgc / d.c.omgrf;
}
__device__ __forceinline__ double ox_norm2(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
__device__ __forceinline__ double ox_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <int EQ_> __global__ void ox_conv_kernel(long long nray, int nv, int npa, const double *ray_vec, const int *npoints, rays_ox_conv *out) {
    const long long iray = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (iray >= nray) return;
    const double one = 1.0, two = 2.0;
    const double pi = (double)3.1415926535897932385f, conversion_threshold = (double)0.0001f;
    const double *v = ray_vec + (size_t)iray * npa * nv;
    const int np = npoints[iray];
    rays_ox_conv o;
    memset(&o, 0, sizeof(o));
    o.ray_number = (int)iray + 1;
    // find_x_max_ray (:202-252)
    bool found_max = false;
    int step_number = 1;
    double alpha_max = 0.0, x_max[3] = {0, 0, 0}, k_max[3] = {0, 0, 0};
    OxEq q;
    ox_equilibrium<EQ_>(v, q);
    double alpha_low = q.alpha0;
    for (int i = 2; i <= np; ++i) {
        ox_equilibrium<EQ_>(v + (size_t)(i - 1) * nv, q);
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
    if (found_max) {
        for (int k = 0; k < 3; ++k) { o.x_max[k] = x_max[k]; o.k_max[k] = k_max[k]; }
        o.alpha_max = alpha_max;
        o.step_number = step_number;
        // find_x_cutoff_ray (:256-311)
        const double alpha_tolerence = one / (10.0 * 10.0 * 10.0 * 10.0);
        bool found_cutoff = false;
        double x_temp[3] = {x_max[0], x_max[1], x_max[2]};
        ox_equilibrium<EQ_>(x_temp, q);
        double alpha_temp = q.alpha0;
        int iteration;
        for (iteration = 1; iteration <= 10; ++iteration) {
            if (fabs(alpha_temp - one) <= alpha_tolerence) { found_cutoff = true; break; }
            ox_equilibrium<EQ_>(x_temp, q);
            alpha_temp = q.alpha0;
            const double ng = ox_norm2(q.gradns0);
            const double mod_grad_alpha = ng * (alpha_temp / q.ns0);
            const double r = sqrt(x_temp[0] * x_temp[0] + x_temp[1] * x_temp[1]);
            const double delta = (one - q.alpha0) / mod_grad_alpha;
            const double step = fmin(delta, 0.25 * r);
            for (int k = 0; k < 3; ++k) x_temp[k] = x_temp[k] + q.gradns0[k] / ng * step;
        }
        o.iteration = iteration;
        o.found_cutoff = found_cutoff ? 1 : 0;
        if (found_cutoff) {
            for (int k = 0; k < 3; ++k) o.x_cut[k] = x_temp[k];
            // OX_conv_coeff (:315-407)
            ox_equilibrium<EQ_>(x_temp, q);
            const double ng = ox_norm2(q.gradns0);
            double xc[3], yc[3], zc[3], vt[3];
            for (int k = 0; k < 3; ++k) xc[k] = q.gradns0[k] / ng;
            vt[0] = q.bunit[1] * xc[2] - q.bunit[2] * xc[1];
            vt[1] = q.bunit[2] * xc[0] - q.bunit[0] * xc[2];
            vt[2] = q.bunit[0] * xc[1] - q.bunit[1] * xc[0];
            const double nvt = ox_norm2(vt);
            for (int k = 0; k < 3; ++k) yc[k] = vt[k] / nvt;
            zc[0] = xc[1] * yc[2] - xc[2] * yc[1];
            zc[1] = xc[2] * yc[0] - xc[0] * yc[2];
            zc[2] = xc[0] * yc[1] - xc[1] * yc[0];
            const double theta = acos(ox_dot(xc, q.bunit));
            const double gamma = fabs(q.gamma0);
            const double L = q.ns0 / ng;
            const double k0 = g_dc.c.k0;
            const double n_vertical = ox_dot(k_max, xc) / k0;
            const double nz_c = ox_dot(k_max, zc) / k0;
            const double ny_c = ox_dot(k_max, yc) / k0;
            const double n_crit = sin(theta) * sqrt(gamma / (one + gamma));
            const double ct = cos(theta), st = sin(theta);
            const double F = 0.5 * (one + gamma) * sqrt(gamma) / pow((one + gamma) * (ct * ct) + (st * st) / two, 1.5);
            const double G = 0.5 * sqrt(gamma) / sqrt((one + gamma) * (ct * ct) + (st * st) / two);
            const double dz = fabs(nz_c) - n_crit, ay = fabs(ny_c);
            const double conv_coeff = exp(-pi * k0 * L * (F * (dz * dz) + G * (ay * ay)));
            if (conv_coeff > conversion_threshold) {
                o.converted = 1;
                o.conv_coeff = conv_coeff;
                for (int k = 0; k < 3; ++k) { o.nvecx_c[k] = n_vertical * xc[k]; o.nvecy_c[k] = ny_c * yc[k]; o.nvecz_c[k] = nz_c * zc[k]; }
            }
        }
    }
    out[iray] = o;
}

// ---- copier kernel: delivers finished rays to the caller's page-locked arrays WHILE the trace kernels run ----------------
// The trace kernel of the page-locked end-to-end path used to copy each finished ray out itself (flush_finished_rays): its warps
// then sit on PCIe back-pressure (14.2 GB per 1M-ray fan against 147 ms of integration), and while a resume pass integrates the
// longest rays the link is idle.  When the whole fan's trajectories fit in HBM, the trace kernel instead writes them there at
// full speed and appends each ended ray to done_list; this kernel - a few CTAs on a second stream, resident beside the trace
// kernel's CTAs for the whole call - takes the entries in order and streams the rows to the host, so that the link is busy from
// the first finished ray to the last.  It only ever waits for the trace kernels, never the other way round.
struct CopyOutArgs {
    long long nray;
    const int *done_list;           // [nray], preset to -1
    unsigned long long *claim;      // next entry to take
    const int *abort_flag;          // set by the host if a trace launch failed
    const int *trace_started;       // set by the trace kernel (TraceArgs::trace_started)
    int *deferred_flag;             // set here when the kernel stops waiting: the host then delivers the rays after the trace (see below)
    const int *npoints;
    const double *ray_vec, *residual;   // [nray][npoints_alloc][nv], [nray][npoints_alloc] (device) or NULL
    double *host_ray_vec, *host_residual;
    int npoints_alloc, host_npoints_alloc, nv;
    long long host_ray0, host_ray_stride;
    int *started;                   // page-locked host counter: one increment per CTA once it is resident (the host launches the
                                    // trace kernel after all have reported, or the trace CTAs would fill the SMs first)
};
__global__ void __launch_bounds__(256) copy_out_kernel(const CopyOutArgs c) {
    const unsigned lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && c.started) { atomicAdd_system(c.started, 1); __threadfence_system(); }
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(c.claim, 1ULL);
        k = __shfl_sync(0xffffffffu, k, 0);
        if ((long long)k >= c.nray) break;
        int ir = -1;
        if (lane == 0) {
            // Waiting ends (and the host delivers everything after the trace kernels, with this kernel run once more) when the trace
            // kernel has not started within ~0.1 s - kernels are being serialised: a profiler, compute-sanitizer,
            // CUDA_LAUNCH_BLOCKING - or when no ray has ended for ~10 s: the copier must never be what hangs the GPU.
            const long long t0 = clock64();
            while ((ir = reinterpret_cast<const volatile int *>(c.done_list)[k]) < 0) {
                if (*reinterpret_cast<const volatile int *>(c.abort_flag) || *reinterpret_cast<volatile int *>(c.deferred_flag)) { ir = -2; break; }
                const long long dt = clock64() - t0;
                if ((dt > 200000000LL && !*reinterpret_cast<const volatile int *>(c.trace_started)) || dt > 20000000000LL) {
                    atomicExch(c.deferred_flag, 1); ir = -2; break;
                }
                __nanosleep(256);
            }
        }
        ir = __shfl_sync(0xffffffffu, ir, 0);
        if (ir < 0) break;
        __threadfence();
        const int np = __ldcg(c.npoints + ir);
        const size_t hrow = (size_t)(c.host_ray0 + (long long)ir * c.host_ray_stride) * c.host_npoints_alloc;
        if (c.host_ray_vec) copy_row_to_host(c.host_ray_vec + hrow * c.nv, c.ray_vec + (size_t)ir * c.npoints_alloc * c.nv, np * c.nv, lane);
        if (c.host_residual) copy_row_to_host(c.host_residual + hrow, c.residual + (size_t)ir * c.npoints_alloc, np, lane);
    }
}

// DFMA-only microbenchmark (fp64 roofline denominator): 8 independent chains per thread
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int ensure_results(long long nray, int nv, int npa, bool traj) {
    CK(cx().npoints.reserve((size_t)nray));
    CK(cx().stop.reserve((size_t)nray));
    CK(cx().pwr.reserve((size_t)nray));
    CK(cx().endres.reserve((size_t)nray));
    CK(cx().maxres.reserve((size_t)nray));
    CK(cx().endpar.reserve((size_t)nray));
    CK(cx().startv.reserve((size_t)nray * nv));
    CK(cx().endv.reserve((size_t)nray * nv));
    CK(cx().queue.reserve(8));
    if (traj) {
        CK(cx().ray_vec.reserve((size_t)nray * npa * nv));
        CK(cx().residual.reserve((size_t)nray * npa));
    }
    cx().res_nray = nray; cx().res_nv = nv; cx().res_npa = npa; cx().have_traj = traj;
    return 0;
}

// launch the trace kernel over rays [first, first+count) of the device fan into result slots [first..)
// trajectories go to traj_base (indexed from 0 for ray `first`) when non-null
struct HostOut { double *ray_vec = nullptr, *residual = nullptr; int npa = 0; long long row0 = 0, stride = 1; };   // device-accessible pinned host arrays

// prepare_only: make every reservation the launches need and return (device allocations synchronise with running kernels, so the
// copier path makes them before its copier kernel starts)
int launch_trace(long long first, long long count, double *traj_base, double *resid_base, bool binned, const HostOut *host = nullptr, bool prepare_only = false) {
    const rays_cfg &c = cx().dc.c;
    const TuOps *ops = tu_ops(c.equilib_model, c.ode_solver);
    if (!ops) return set_err(RAYS_ERR_INVALID_CONFIG, "no kernel for this equilibrium/ode pair");
    TraceArgs a{};
    a.nray = count;
    a.rvec0 = cx().rvec0.p + 3 * first;
    a.rindex_vec0 = cx().nvec0.p + 3 * first;
    a.ray_pwr_wt = cx().wt.p + first;
    a.ray_vec = traj_base;
    a.residual = resid_base;
    a.npoints_alloc = cx().res_npa;
    a.npoints = cx().npoints.p + first;
    a.stop_code = cx().stop.p + first;
    a.initial_ray_power = cx().pwr.p + first;
    a.end_residuals = cx().endres.p + first;
    a.max_residuals = cx().maxres.p + first;
    a.end_ray_parameter = cx().endpar.p + first;
    a.start_ray_vec = cx().startv.p + (size_t)first * cx().res_nv;
    a.end_ray_vec = cx().endv.p + (size_t)first * cx().res_nv;
    a.queue = cx().queue.p;
    a.counters = cx().queue.p + 1;
    a.done_list = cx().copier_on ? cx().done_list.p : nullptr;
    a.done_count = cx().copier_on ? cx().copy_ctl.p : nullptr;
    a.trace_started = cx().copier_on ? reinterpret_cast<int *>(cx().copy_ctl.p + 4) : nullptr;
    a.dep_acc = binned ? cx().dep.p : nullptr;
    a.n_bins = cx().dep_bins; a.grid_min = cx().dep_min; a.grid_max = cx().dep_max; a.dep_scale = cx().dep_scale;
    a.dep_smem = (binned && (size_t)cx().dep_bins * 8 <= 40 * 1024) ? cx().dep_bins * 8 : 0;
    int bps = cx().cached_bps;
    const char *name = cx().cached_name;
    if (bps <= 0 || cx().cached_ops != ops || cx().cached_dep_smem != a.dep_smem) {   // occupancy of the selected specialisation: queried once per configuration
        size_t sgb = 0;
        int rpc = kTraceBlock;
        CK(ops->trace(cx().sel, a, 0, cx().stream, &bps, &name, &sgb, &rpc));
        if (bps < 1) bps = 1;
        cx().cached_bps = bps; cx().cached_name = name; cx().cached_ops = ops; cx().cached_sgb = sgb; cx().cached_rpc = rpc; cx().cached_dep_smem = a.dep_smem;
    }
    if (cx().cached_sgb) {   // Shampine-Gordon slot machine: slot memory for every CTA of the largest grid
        CK(cx().sg_state.reserve(((size_t)cx().num_sms * bps * cx().cached_sgb + 7) / 8));
        a.sg_state = cx().sg_state.p;
    }
    long long blocks_needed = (count + kTraceBlock - 1) / kTraceBlock;   // (an SG CTA holds twice that many rays; a small fan still spreads over the SMs)
    int grid = (int)std::min<long long>((long long)cx().num_sms * bps, std::max<long long>(blocks_needed, 1));
    const int rays_per_cta = cx().cached_rpc;   // rays in flight per CTA (SG slot machine: its slots)
    if (host) {   // streaming copy-out: per-lane (per-slot) staging rows, finished rays go straight to the caller's arrays
        const size_t lanes = (size_t)grid * rays_per_cta;
        if (host->ray_vec) { CK(cx().ray_vec.reserve(lanes * cx().res_npa * cx().res_nv)); a.ray_vec = cx().ray_vec.p; a.host_ray_vec = host->ray_vec; }
        if (host->residual) { CK(cx().residual.reserve(lanes * cx().res_npa)); a.residual = cx().residual.p; a.host_residual = host->residual; }
        a.host_npoints_alloc = host->npa;
        a.host_ray0 = host->row0 + first * host->stride;
        a.host_ray_stride = host->stride;
    }
    // Time slicing (see TraceArgs): rays are suspended after `slice` steps and the survivors are re-launched packed
    // into full warps, until they fit one per lane.  RAYS_B200_SLICE=0 disables it, =n forces n steps.
    const long long lanes = (long long)cx().num_sms * bps * rays_per_cta;
    int slice = 0;
    if (count >= 2 * lanes) slice = std::max(64, c.nstep_max / 4);
    bool forced = false;   // an explicit RAYS_B200_SLICE applies to fans of any size (tests), every pass
    // An SG ray-step is ~25 right-hand sides and the lanes of a warp drift apart in integrator phase: short slices on
    // every pass keep the warps packed (1M-ray Solov'ev fan, tol 1e-6: 250 -> 64 steps: 5.15e7 -> 5.65e7 ray-steps/s)
    // (the per-lane SG kernel of round 1 needs short slices on every pass to keep its warps packed; the slot machine regroups
    // its slots every iteration and slices like RK4)
    if (c.ode_solver == RAYS_ODE_SG && cx().sel.sg_lanes && count >= 2 * lanes) { slice = 64; forced = true; }
    if (const char *env = getenv("RAYS_B200_SLICE")) { if (env[0]) { slice = atoi(env); forced = true; } }
    a.sg_align = 1;        // measurement aid: RAYS_B200_SG_ALIGN=0 lets the SG lanes run free (results are identical)
    if (const char *env = getenv("RAYS_B200_SG_ALIGN")) { if (env[0]) a.sg_align = atoi(env) != 0; }
    if (slice > 0) {
        CK(cx().cont_state.reserve((size_t)count * kContStride));
        CK(cx().cont_list[0].reserve((size_t)count)); CK(cx().cont_list[1].reserve((size_t)count));
    }
    if (prepare_only) {
        // ... and run the selected kernel once over an empty fan: a kernel's first launch loads its code (lazy module loading), and
        // that load can wait for every running kernel - i.e. for the copier, which waits for this one
        TraceArgs w = a;
        w.nray = 0; w.slice_steps = 0; w.resume = 0; w.order = nullptr; w.done_list = nullptr; w.done_count = nullptr;
        w.cont_state = cx().cont_state.p; w.cont_list = cx().cont_list[0].p; w.cont_count = cx().queue.p + 3;
        CK(ops->trace(cx().sel, w, 1, cx().stream, nullptr, nullptr, nullptr, nullptr));
        CK(cudaStreamSynchronize(cx().stream));
        cx().last_launches += 1;
        return 0;
    }
    long long n_this = count;
    for (int phase = 0;; ++phase) {
        a.nray = n_this;
        a.slice_steps = (slice > 0 && (forced || n_this > lanes)) ? slice : 0;      // survivors that fit one per lane run to the end
        a.resume = phase > 0 ? 1 : 0;
        a.order = phase > 0 ? cx().cont_list[(phase - 1) & 1].p : nullptr;
        a.cont_state = cx().cont_state.p;
        a.cont_list = cx().cont_list[phase & 1].p;
        a.cont_count = cx().queue.p + 3;
        CK(cudaMemsetAsync(cx().queue.p, 0, sizeof(unsigned long long), cx().stream));
        CK(cudaMemsetAsync(cx().queue.p + 3, 0, sizeof(unsigned long long), cx().stream));
        const int grid_p = (int)std::min<long long>((long long)cx().num_sms * bps, std::max<long long>((n_this + kTraceBlock - 1) / kTraceBlock, 1));
        CK(cudaEventRecord(cx().ev_m0, cx().stream));
        CK(ops->trace(cx().sel, a, grid_p, cx().stream, nullptr, nullptr, nullptr, nullptr));
        CK(cudaEventRecord(cx().ev_m1, cx().stream));
        cx().last_launches += 1;
        cx().last_phases = phase + 1;
        if (a.slice_steps == 0) { cx().main_pending = true; cx().main_is_resume = phase > 0; break; }
        unsigned long long n_susp = 0;
        CK(cudaMemcpyAsync(&n_susp, cx().queue.p + 3, sizeof(n_susp), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, cx().ev_m0, cx().ev_m1));
        (phase > 0 ? cx().last_resume_ms : cx().last_first_ms) += ms;
        if (n_susp == 0) break;
        n_this = (long long)n_susp;
    }
    cx().last_kernel = name; cx().last_grid = grid; cx().last_bps = bps;
    return 0;
}

int fetch_counters() {
    unsigned long long h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, cx().queue.p, sizeof(h), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    if (cx().main_pending) {   // duration of the (last) trace kernel itself
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, cx().ev_m0, cx().ev_m1) == cudaSuccess) (cx().main_is_resume ? cx().last_resume_ms : cx().last_first_ms) += ms;
        cx().main_pending = false;
    }
    cx().last_steps = (long long)h[1];
    cx().last_rhs = (long long)h[2];
    return 0;
}

// Where the rays of a device fan live in the caller's result arrays: local ray i is row row0 + i*stride
// (row0 = 0, stride = 1 for a whole fan; row0 = gpu, stride = ngpu for the interleaved shards of rays_b200_trace_multi)
struct RowMap { long long row0 = 0, stride = 1; };

int copy_small_results(rays_results *res, long long first, long long count, RowMap rm = RowMap{}) {
    const int nv = cx().res_nv;
    if (rm.stride == 1) {
        const long long o = rm.row0 + first;
        auto cp = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            if (!dst || bytes == 0) return cudaSuccess;
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cx().stream);
        };
        CK(cp(res->npoints ? res->npoints + o : nullptr, cx().npoints.p + first, count * sizeof(int)));
        CK(cp(res->ray_stop_code ? res->ray_stop_code + o : nullptr, cx().stop.p + first, count * sizeof(int)));
        CK(cp(res->initial_ray_power ? res->initial_ray_power + o : nullptr, cx().pwr.p + first, count * sizeof(double)));
        CK(cp(res->end_residuals ? res->end_residuals + o : nullptr, cx().endres.p + first, count * sizeof(double)));
        CK(cp(res->max_residuals ? res->max_residuals + o : nullptr, cx().maxres.p + first, count * sizeof(double)));
        CK(cp(res->end_ray_parameter ? res->end_ray_parameter + o : nullptr, cx().endpar.p + first, count * sizeof(double)));
        CK(cp(res->start_ray_vec ? res->start_ray_vec + (size_t)o * nv : nullptr, cx().startv.p + (size_t)first * nv, (size_t)count * nv * sizeof(double)));
        CK(cp(res->end_ray_vec ? res->end_ray_vec + (size_t)o * nv : nullptr, cx().endv.p + (size_t)first * nv, (size_t)count * nv * sizeof(double)));
        return 0;
    }
    // interleaved shard: one contiguous D2H per array, scattered to the strided rows on the host
    std::vector<int> hi((size_t)count);
    std::vector<double> hd((size_t)count * nv);
    auto geti = [&](int32_t *dst, const int *src) -> int {
        if (!dst) return 0;
        CK(cudaMemcpyAsync(hi.data(), src + first, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        for (long long i = 0; i < count; ++i) dst[rm.row0 + (first + i) * rm.stride] = hi[(size_t)i];
        return 0;
    };
    auto getd = [&](double *dst, const double *src, int w) -> int {
        if (!dst) return 0;
        CK(cudaMemcpyAsync(hd.data(), src + (size_t)first * w, (size_t)count * w * sizeof(double), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        for (long long i = 0; i < count; ++i)
            for (int k = 0; k < w; ++k) dst[(size_t)(rm.row0 + (first + i) * rm.stride) * w + k] = hd[(size_t)i * w + k];
        return 0;
    };
    int rc;
    if ((rc = geti(res->npoints, cx().npoints.p)) || (rc = geti(res->ray_stop_code, cx().stop.p))) return rc;
    if ((rc = getd(res->initial_ray_power, cx().pwr.p, 1)) || (rc = getd(res->end_residuals, cx().endres.p, 1)) ||
        (rc = getd(res->max_residuals, cx().maxres.p, 1)) || (rc = getd(res->end_ray_parameter, cx().endpar.p, 1)) ||
        (rc = getd(res->start_ray_vec, cx().startv.p, nv)) || (rc = getd(res->end_ray_vec, cx().endv.p, nv))) return rc;
    return 0;
}

void fill_flags(rays_results *res, const std::vector<int> &codes, long long first, RowMap rm = RowMap{}) {
    if (!res->ray_stop_flag) return;
    char table[RAYS_STOP_CODE_MAX][RAYS_FLAG_LEN];
    for (int c = 0; c < RAYS_STOP_CODE_MAX; ++c) rays_b200_stop_string(c, table[c], RAYS_FLAG_LEN);
    for (size_t i = 0; i < codes.size(); ++i) {
        const int c = (codes[i] >= 0 && codes[i] < RAYS_STOP_CODE_MAX) ? codes[i] : 0;
        std::memcpy(res->ray_stop_flag + (size_t)(rm.row0 + (first + (long long)i) * rm.stride) * RAYS_FLAG_LEN, table[c], RAYS_FLAG_LEN);
    }
}

// D2H of trajectories in the reference layout, trimmed to the points actually written: rays are
// taken in groups of 64, the copy width of a group is its longest ray rounded up to 32 points, and
// neighbouring groups of equal width share one cudaMemcpy2DAsync.  (A fan whose longest ray hits
// nstep_max would otherwise move every ray at full length.)
int copy_trajectories(rays_results *res, const std::vector<int> &np, long long first, long long count, const double *tv, const double *tr,
                      int npa, int nv, cudaStream_t st, RowMap rm = RowMap{}) {
    const long long G = 64;
    long long run0 = first;
    int w0 = -1;
    auto flush = [&](long long a, long long b, int w) -> cudaError_t {
        if (b <= a || w <= 0) return cudaSuccess;
        cudaError_t e = cudaSuccess;
        if (res->ray_vec)
            e = cudaMemcpy2DAsync(res->ray_vec + (size_t)(rm.row0 + a * rm.stride) * res->npoints_alloc * nv, (size_t)rm.stride * res->npoints_alloc * nv * 8,
                                  tv + (size_t)(a - first) * npa * nv, (size_t)npa * nv * 8, (size_t)w * nv * 8, (size_t)(b - a), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && res->residual)
            e = cudaMemcpy2DAsync(res->residual + (size_t)(rm.row0 + a * rm.stride) * res->npoints_alloc, (size_t)rm.stride * res->npoints_alloc * 8, tr + (size_t)(a - first) * npa,
                                  (size_t)npa * 8, (size_t)w * 8, (size_t)(b - a), cudaMemcpyDeviceToHost, st);
        return e;
    };
    for (long long gs = first; gs < first + count; gs += G) {
        const long long ge = std::min(gs + G, first + count);
        int m = 1;
        for (long long i = gs; i < ge; ++i) m = std::max(m, np[(size_t)i]);
        const int w = std::min(npa, (m + 31) / 32 * 32);
        if (w != w0) {
            CK(flush(run0, gs, w0));
            run0 = gs; w0 = w;
        }
    }
    CK(flush(run0, first + count, w0));
    return 0;
}

// Is p ordinary pageable host memory (not page-locked, not registered)?
bool is_pageable(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}
// the staged path serves pageable destinations; RAYS_B200_STAGED_COPY=0 keeps the plain cudaMemcpy2D (measurement aid)
bool staged_copy_wanted(const rays_results *res) {
    const char *e = getenv("RAYS_B200_STAGED_COPY");
    if (e && e[0] == '0') return false;
    return (res->ray_vec ? is_pageable(res->ray_vec) : true) && (res->residual ? is_pageable(res->residual) : true) && (res->ray_vec || res->residual);
}
int g_copy_threads_div = 1;   // GPUs driven by this process (rays_b200_init_multi): the host cores are shared between their copy-outs
int copy_threads() {
    if (const char *e = getenv("RAYS_B200_COPY_THREADS")) { const int v = atoi(e); if (v > 0) return std::min(v, 64); }
    const int hw = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(16, (hw > 0 ? hw : 4) / std::max(1, g_copy_threads_div)));
}

// D2H of trajectories into PAGEABLE arrays (what `allocate(ray_vec(nv, nstep_max+1, nray))` of an unmodified host gives,
// ray_results_m.f90:132-142).  A cudaMemcpy2D into pageable memory is staged by the driver through one small bounce buffer and
// reaches ~3 GB/s; here the saved points are packed on the device (exactly the algorithmic bytes), moved by the copy engine into a
// ring of page-locked chunks, and spread into the caller's rows by host threads while the next chunk is in flight.
int copy_trajectories_staged(rays_results *res, const std::vector<int> &np, long long first, long long count, const double *tv, const double *tr,
                             int npa, int nv, RowMap rm = RowMap{}) {
    Ctx &c = cx();
    const bool wv = res->ray_vec != nullptr, wr = res->residual != nullptr;
    if (count <= 0 || (!wv && !wr)) return 0;
    const int per_pt = (wv ? nv : 0) + (wr ? 1 : 0);
    constexpr int R = Ctx::kRing;
    const size_t chunk = (size_t(64) << 20) / 8;   // doubles per chunk
    if (!c.ring_host || c.ring_chunk_doubles != chunk) {
        if (c.ring_host) cudaFreeHost(c.ring_host);
        c.ring_host = nullptr;
        CK(cudaHostAlloc((void **)&c.ring_host, R * chunk * 8, cudaHostAllocDefault));
        c.ring_chunk_doubles = chunk;
        for (int k = 0; k < R; ++k) if (!c.ev_ring[k]) CK(cudaEventCreateWithFlags(&c.ev_ring[k], cudaEventDisableTiming));
    }
    CK(c.pack_dev.reserve(R * chunk));
    std::vector<long long> off((size_t)count + 1);
    off[0] = 0;
    for (long long i = 0; i < count; ++i) {
        const int n = np[(size_t)(first + i)];
        if (n > res->npoints_alloc) return set_err(RAYS_ERR_INVALID_CONFIG, "copy-out: npoints_alloc smaller than the longest ray");
        off[(size_t)i + 1] = off[(size_t)i] + (long long)n * per_pt;
    }
    if ((size_t)npa * per_pt > chunk) return set_err(RAYS_ERR_INVALID_CONFIG, "copy-out: one ray does not fit a staging chunk");
    CK(c.pack_off.reserve((size_t)count + 1));
    CK(cudaMemcpyAsync(c.pack_off.p, off.data(), ((size_t)count + 1) * sizeof(long long), cudaMemcpyHostToDevice, c.copy_stream));
    const int T = copy_threads();
    const int dev = c.device;
    std::vector<std::thread> workers[R];
    std::atomic<int> failed{0};
    auto join = [&](int slot) { for (auto &t : workers[slot]) t.join(); workers[slot].clear(); };
    int rc = 0, ic = 0;
    for (long long a = 0; a < count && !rc; ++ic) {
        long long b = a + 1;   // rays [a, b) of the batch: as many as fit one chunk
        while (b < count && (size_t)(off[(size_t)b + 1] - off[(size_t)a]) <= chunk) ++b;
        const int slot = ic % R;
        join(slot);
        double *dchunk = c.pack_dev.p + (size_t)slot * chunk;
        double *hchunk = c.ring_host + (size_t)slot * chunk;
        const size_t ndbl = (size_t)(off[(size_t)b] - off[(size_t)a]);
        cudaError_t e = cudaSuccess;
        if (ndbl) {
            pack_rows_kernel<<<(unsigned)(b - a), 128, 0, c.copy_stream>>>(wv ? tv + (size_t)a * npa * nv : nullptr, wr ? tr + (size_t)a * npa : nullptr, c.npoints.p + first + a,
                                                                          c.pack_off.p + a, off[(size_t)a], npa, nv, dchunk);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(hchunk, dchunk, ndbl * 8, cudaMemcpyDeviceToHost, c.copy_stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(c.ev_ring[slot], c.copy_stream);
        if (e != cudaSuccess) { rc = set_err(RAYS_ERR_CUDA, std::string("staged copy-out: ") + cudaGetErrorString(e)); break; }
        cudaEvent_t ev = c.ev_ring[slot];
        const long long base = off[(size_t)a];
        for (int t = 0; t < T; ++t) {
            // thread t spreads the rays whose packed data start in its share of the chunk
            const long long lo_off = base + (long long)(ndbl * (size_t)t / T), hi_off = base + (long long)(ndbl * (size_t)(t + 1) / T);
            const long long r0 = std::lower_bound(off.begin() + a, off.begin() + b, lo_off) - off.begin();
            const long long r1 = (t == T - 1) ? b : std::lower_bound(off.begin() + a, off.begin() + b, hi_off) - off.begin();
            if (r1 <= r0) continue;
            workers[slot].emplace_back([=, &off, &np, &failed]() {
                cudaSetDevice(dev);
                if (cudaEventSynchronize(ev) != cudaSuccess) { failed.store(1); return; }
                for (long long i = r0; i < r1; ++i) {
                    const int n = np[(size_t)(first + i)];
                    const double *src = hchunk + (off[(size_t)i] - base);
                    const size_t row = (size_t)(rm.row0 + (first + i) * rm.stride);
                    if (wv) { std::memcpy(res->ray_vec + row * res->npoints_alloc * nv, src, (size_t)n * nv * 8); src += (size_t)n * nv; }
                    if (wr) std::memcpy(res->residual + row * res->npoints_alloc, src, (size_t)n * 8);
                }
            });
        }
        a = b;
    }
    for (int k = 0; k < R; ++k) join(k);
    if (!rc && failed.load()) rc = set_err(RAYS_ERR_CUDA, "staged copy-out: a copy of a staging chunk failed");
    if (!rc) CK(cudaStreamSynchronize(c.copy_stream));
    return rc;
}

}  // namespace

extern "C" {

int rays_b200_version(void) { return 100; }
int rays_b200_struct_sizes(int32_t *out, int n) {
    const int32_t sz[13] = {(int32_t)sizeof(rays_cfg), (int32_t)sizeof(rays_fan), (int32_t)sizeof(rays_results), (int32_t)sizeof(rays_deposition),
                            (int32_t)sizeof(rays_solovev_launch), (int32_t)sizeof(rays_axisym_launch), (int32_t)sizeof(rays_slab_launch),
                            (int32_t)sizeof(rays_spline1d), (int32_t)sizeof(rays_spline2d), (int32_t)sizeof(rays_slab_eq),
                            (int32_t)sizeof(rays_solovev_eq), (int32_t)sizeof(rays_axisym_eq), (int32_t)sizeof(rays_mirror_eq)};
    for (int i = 0; i < n && i < 13; ++i) out[i] = sz[i];
    return 13;
}
const char *rays_b200_last_error(void) { return g_err.c_str(); }

int rays_b200_stop_string(int code, char *buf, int len) {
    const char *s = (code >= 0 && code < RAYS_STOP_CODE_MAX) ? kStopStrings[code] : "";
    const int n = (int)std::strlen(s);
    for (int i = 0; i < len; ++i) buf[i] = i < n ? s[i] : ' ';
    return 0;
}

int rays_b200_init(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(RAYS_ERR_CUDA, std::string("rays_b200_init: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(RAYS_ERR_CUDA, "rays_b200_init: device index out of range");
    if (cx().inited && cx().device == device) return 0;
    if (cx().inited) rays_b200_finalize();
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return set_err(RAYS_ERR_CUDA, std::string("rays_b200_init: kernels are built for sm_100a only; found ") + prop.name);
    cx().device = device;
    cx().num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&cx().stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&cx().copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&cx().ev0));
    CK(cudaEventCreate(&cx().ev1));
    CK(cudaEventCreate(&cx().ev_m0));
    CK(cudaEventCreate(&cx().ev_m1));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&cx().ev_batch[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&cx().ev_copy[i], cudaEventDisableTiming));
    }
    cx().inited = true;
    return 0;
}

int rays_b200_finalize(void) {
    if (!cx().inited) return 0;
    cudaSetDevice(cx().device);
    cudaDeviceSynchronize();
    DevBuf<double> *bufs[] = {&cx().zx, &cx().zf, &cx().rgrid, &cx().zgrid, &cx().br, &cx().bz, &cx().aphi, &cx().prof_grid[0], &cx().prof_grid[1], &cx().prof_grid[2],
                              &cx().prof_fspl[0], &cx().prof_fspl[1], &cx().prof_fspl[2], &cx().eq_rgrid, &cx().eq_zgrid, &cx().eq_psi, &cx().eq_T, &cx().rvec0, &cx().nvec0, &cx().wt, &cx().ray_vec,
                              &cx().residual, &cx().pwr, &cx().endres, &cx().maxres, &cx().endpar, &cx().startv, &cx().endv};
    for (auto *b : bufs) b->release();
    cx().npoints.release(); cx().stop.release(); cx().queue.release(); cx().dep.release();
    cx().cont_state.release(); cx().cont_list[0].release(); cx().cont_list[1].release(); cx().sg_state.release();
    cudaEventDestroy(cx().ev_m0); cudaEventDestroy(cx().ev_m1);
    if (cx().pinned) cudaFreeHost(cx().pinned);
    cx().pinned = nullptr; cx().pinned_bytes = 0;
    cx().pack_dev.release(); cx().pack_off.release(); cx().done_list.release(); cx().copy_ctl.release();
    if (cx().ring_host) cudaFreeHost(cx().ring_host);
    if (cx().copier_started) { cudaFreeHost(cx().copier_started); cx().copier_started = nullptr; }
    for (int k = 0; k < Ctx::kRing; ++k) if (cx().ev_ring[k]) cudaEventDestroy(cx().ev_ring[k]);
    cudaEventDestroy(cx().ev0); cudaEventDestroy(cx().ev1);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(cx().ev_batch[i]); cudaEventDestroy(cx().ev_copy[i]); }
    cudaStreamDestroy(cx().stream); cudaStreamDestroy(cx().copy_stream);
    cx() = Ctx{};
    return 0;
}

void *rays_b200_stream(void) { return cx().inited ? (void *)cx().stream : nullptr; }

int rays_b200_set_config(const rays_cfg *cfg) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cfg) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_set_config: null config");
    int rc = validate_cfg(*cfg);
    if (rc) return rc;
    CK(cudaSetDevice(cx().device));
    cx().cfg_set = false;      // a failed upload below must not leave a half-updated configuration usable
    DevCfg &d = cx().dc;
    d.c = *cfg;
    rays_cfg &c = d.c;
    // tables -> HBM; the struct then carries device pointers
    if (c.damping_model == RAYS_DAMP_FUND_ECH) {
        if ((rc = upload_table(cx().zx, cfg->zfun_re.x_grid, (size_t)cfg->zfun_re.nx))) return rc;
        if ((rc = upload_table(cx().zf, cfg->zfun_re.fspl, (size_t)4 * cfg->zfun_re.nx))) return rc;
        c.zfun_re.x_grid = cx().zx.p; c.zfun_re.fspl = cx().zf.p;
    } else { c.zfun_re.x_grid = nullptr; c.zfun_re.fspl = nullptr; c.zfun_re.nx = 0; }
    if (c.equilib_model == RAYS_EQ_MULTIPLE_MIRROR) {
        const rays_mirror_eq &m = cfg->mirror;
        const size_t nx = (size_t)m.Br_spline.nx, ny = (size_t)m.Br_spline.ny;
        if ((rc = upload_table(cx().rgrid, m.Br_spline.x_grid, nx))) return rc;
        if ((rc = upload_table(cx().zgrid, m.Br_spline.y_grid, ny))) return rc;
        if ((rc = upload_table(cx().br, m.Br_spline.fspl, 16 * nx * ny))) return rc;
        if ((rc = upload_table(cx().bz, m.Bz_spline.fspl, 16 * nx * ny))) return rc;
        if ((rc = upload_table(cx().aphi, m.Aphi_spline.fspl, 16 * nx * ny))) return rc;
        rays_spline2d base = m.Br_spline;
        base.x_grid = cx().rgrid.p; base.y_grid = cx().zgrid.p;
        c.mirror.Br_spline = base; c.mirror.Br_spline.fspl = cx().br.p;
        c.mirror.Bz_spline = base; c.mirror.Bz_spline.fspl = cx().bz.p;
        c.mirror.Aphi_spline = base; c.mirror.Aphi_spline.fspl = cx().aphi.p;
        d.mir_x1[0] = m.Br_spline.x_grid[0]; d.mir_xn[0] = m.Br_spline.x_grid[nx - 1];
        d.mir_x1[1] = m.Br_spline.y_grid[0]; d.mir_xn[1] = m.Br_spline.y_grid[ny - 1];
        for (int k = 0; k < 2; ++k) { d.mir_range[k].d = d.mir_xn[k] - d.mir_x1[k]; d.mir_range[k].r = 1.0 / d.mir_range[k].d; }
    } else {
        rays_spline2d none{};
        c.mirror.Br_spline = none; c.mirror.Bz_spline = none; c.mirror.Aphi_spline = none;
    }
    {   // axisym_toroid 1-D profile splines (density_spline_interp_m, temperature_spline_interp_m)
        rays_spline1d *sp[3] = {&c.axisym.ne_spline, &c.axisym.Te_spline, &c.axisym.Ti_spline};
        const rays_spline1d *src[3] = {&cfg->axisym.ne_spline, &cfg->axisym.Te_spline, &cfg->axisym.Ti_spline};
        bool used[3] = {false, false, false};
        if (c.equilib_model == RAYS_EQ_AXISYM_TOROID) {
            used[0] = c.axisym.density_prof_model == RAYS_PROF_SPLINE;
            for (int s = 0; s <= c.nspec; ++s)
                if (c.axisym.temperature_prof_model[s] == RAYS_PROF_SPLINE) used[s == 0 ? 1 : 2] = true;
        }
        for (int k = 0; k < 3; ++k) {
            *sp[k] = rays_spline1d{};
            if (!used[k]) continue;
            const size_t nx = (size_t)src[k]->nx;
            if ((rc = upload_table(cx().prof_grid[k], src[k]->x_grid, nx))) return rc;
            if ((rc = upload_table(cx().prof_fspl[k], src[k]->fspl, 4 * nx))) return rc;
            sp[k]->nx = src[k]->nx; sp[k]->x_grid = cx().prof_grid[k].p; sp[k]->fspl = cx().prof_fspl[k].p;
        }
    }
    c.axisym.Psi_spline = rays_spline2d{}; c.axisym.T_spline = rays_spline1d{};
    if (c.equilib_model == RAYS_EQ_AXISYM_TOROID && c.axisym.magnetics_model == RAYS_MAG_EQDSK_SPLINE) {
        const rays_axisym_eq &a = cfg->axisym;   // eqdsk_magnetics_spline_interp_m: Psi_profile (bicubic), T_profile (cubic)
        const size_t nx = (size_t)a.Psi_spline.nx, ny = (size_t)a.Psi_spline.ny, nt = (size_t)a.T_spline.nx;
        if ((rc = upload_table(cx().eq_rgrid, a.Psi_spline.x_grid, nx))) return rc;
        if ((rc = upload_table(cx().eq_zgrid, a.Psi_spline.y_grid, ny))) return rc;
        if ((rc = upload_table(cx().eq_psi, a.Psi_spline.fspl, 16 * nx * ny))) return rc;
        if ((rc = upload_table(cx().eq_T, a.T_spline.fspl, 4 * nt))) return rc;
        c.axisym.Psi_spline.nx = a.Psi_spline.nx; c.axisym.Psi_spline.ny = a.Psi_spline.ny;
        c.axisym.Psi_spline.x_grid = cx().eq_rgrid.p; c.axisym.Psi_spline.y_grid = cx().eq_zgrid.p; c.axisym.Psi_spline.fspl = cx().eq_psi.p;
        c.axisym.T_spline.nx = a.T_spline.nx; c.axisym.T_spline.fspl = cx().eq_T.p;
        if (a.T_spline.x_grid == a.Psi_spline.x_grid && nt == nx) c.axisym.T_spline.x_grid = cx().eq_rgrid.p;
        else return set_err(RAYS_ERR_INVALID_CONFIG, "axisym_toroid: T_profile must be splined on the g-file's R grid (eqdsk_magnetics_spline_interp_m.f90:184)");
    }
    // constant products, formed with the same IEEE operations the reference performs per call
    for (int s = 0; s < RAYS_NSPECIES; ++s) { d.qs2[s] = c.qs[s] * c.qs[s]; d.eps0ms[s] = c.eps0 * c.ms[s]; }
    d.omgrf2 = c.omgrf * c.omgrf;
    d.two_over_k0 = 2.0 / c.k0;
    d.two_over_omgrf = 2.0 / c.omgrf;
    d.m2_over_omgrf = -(2.0 / c.omgrf);
    d.one_over_omgrf = 1.0 / c.omgrf;
    if (c.equilib_model == RAYS_EQ_SOLOVEV) {
        d.sv_rmaj = c.solovev.rmaj; d.sv_kappa = c.solovev.kappa; d.sv_bphi0 = c.solovev.bphi0; d.sv_iota0 = c.solovev.iota0; d.sv_psiB = c.solovev.psiB;
    } else {
        d.sv_rmaj = c.axisym.sm_rmaj; d.sv_kappa = c.axisym.sm_kappa; d.sv_bphi0 = c.axisym.sm_bphi0; d.sv_iota0 = c.axisym.sm_iota0; d.sv_psiB = c.axisym.sm_psiB;
    }
    d.sv_bp0 = d.sv_bphi0 * d.sv_iota0;
    d.sv_rk = d.sv_rmaj * d.sv_kappa;
    d.sv_rk2 = d.sv_rk * d.sv_rk;
    d.sv_rmaj2 = d.sv_rmaj * d.sv_rmaj;
    d.sv_bphi0_rmaj = d.sv_bphi0 * d.sv_rmaj;
    d.dn_delta = (double)1.e-6f;
    d.dn_two_delta = 2.0 * d.dn_delta;
    d.dn_omg_p = c.omgrf * (1.0 + d.dn_delta / 2.0);
    d.dn_omg_m = c.omgrf * (1.0 - d.dn_delta / 2.0);
    d.dn_k0_p = d.dn_omg_p / c.clight;
    d.dn_k0_m = d.dn_omg_m / c.clight;
    d.dn_omg_p2 = d.dn_omg_p * d.dn_omg_p;
    d.dn_omg_m2 = d.dn_omg_m * d.dn_omg_m;
    d.dn_omg_delta = c.omgrf * d.dn_delta;
    auto mk = [](double v) { Rcp r; r.d = v; r.r = 1.0 / v; return r; };   // IEEE, correctly rounded
    d.rc_k0 = mk(c.k0); d.rc_omgrf = mk(c.omgrf); d.rc_omgrf2 = mk(d.omgrf2); d.rc_clight = mk(c.clight); d.rc_six = mk(6.0);
    for (int s = 0; s < RAYS_NSPECIES; ++s) { d.rc_ms[s] = mk(c.ms[s]); d.rc_eps0ms[s] = mk(d.eps0ms[s]); }
    d.rc_rk = mk(d.sv_rk); d.rc_rk2 = mk(d.sv_rk2); d.rc_rmaj = mk(d.sv_rmaj); d.rc_rmaj2 = mk(d.sv_rmaj2); d.rc_psiB = mk(d.sv_psiB);
    d.rc_Aphi_LUFS = mk(c.mirror.Aphi_LUFS);
    d.rc_eq_psibound = mk(c.axisym.eq_psibound);
    d.hyp_delta[0] = mk(c.mirror.delta_d); d.hyp_two_delta[0] = mk(2.0 * c.mirror.delta_d); d.hyp_t0[0] = mk(std::tanh(c.mirror.AphiN0_d / c.mirror.delta_d));
    for (int s = 0; s < RAYS_NSPECIES; ++s) {
        d.hyp_delta[1 + s] = mk(c.mirror.delta_t[s]); d.hyp_two_delta[1 + s] = mk(2.0 * c.mirror.delta_t[s]);
        d.hyp_t0[1 + s] = mk(std::tanh(c.mirror.AphiN0_t[s] / c.mirror.delta_t[s]));
    }
    d.rc_two_delta = mk(d.dn_two_delta); d.rc_omg_p = mk(d.dn_omg_p); d.rc_omg_m = mk(d.dn_omg_m); d.rc_omg_p2 = mk(d.dn_omg_p2);
    d.rc_omg_m2 = mk(d.dn_omg_m2); d.rc_k0_p = mk(d.dn_k0_p); d.rc_k0_m = mk(d.dn_k0_m); d.rc_omg_delta = mk(d.dn_omg_delta);
    {   // are the temperatures read by anything on this run's path?
        bool need = c.damping_model != RAYS_DAMP_NONE || c.integrate_eq_gradients != 0;
        for (int s = 0; s <= c.nspec; ++s) {
            if (c.t0s[s] < 0.0) need = true;                                                    // 'negative_temp' must still fire
            if (c.equilib_model == RAYS_EQ_SOLOVEV && c.solovev.t_prof_model[s] == RAYS_PROF_CONSTANT) need = true;   // (R) resets the density
            if (c.equilib_model == RAYS_EQ_AXISYM_TOROID && c.axisym.temperature_prof_model[s] == RAYS_PROF_SPLINE) need = true;   // a splined profile may go negative
        }
        d.need_temp = need ? 1 : 0;
    }
    // kernel selection: the two-species specialisations cover electron + one ion without per-species damping slots
    cx().sel.ray_deriv = c.ray_deriv;
    cx().sel.generic = !(c.nspec == 1 && !c.multi_spec_damping);
    cx().sel.damp = c.damping_model != RAYS_DAMP_NONE;
    cx().sel.grads = c.integrate_eq_gradients != 0;
    cx().sel.sg_lanes = false;   // measurement aid: RAYS_B200_SG_LANES=1 runs the per-lane SG state machine (identical results)
    if (const char *env = getenv("RAYS_B200_SG_LANES")) cx().sel.sg_lanes = env[0] && atoi(env) != 0;
    cx().cached_bps = 0; cx().cached_ops = nullptr;
    for (int ode = 1; ode <= 2; ++ode) {
        const TuOps *ops = tu_ops(c.equilib_model, ode);
        if (ops) CK(ops->upload(&d, cx().stream));
    }
    CK(cudaMemcpyToSymbolAsync(g_dc, &d, sizeof(DevCfg), 0, cudaMemcpyHostToDevice, cx().stream));   // this TU: deposition kernel
    CK(cudaStreamSynchronize(cx().stream));
    cx().cfg_set = true;
    return 0;
}

int rays_b200_fan_upload(const rays_fan *fan) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!fan || fan->nray < 0 || (fan->nray > 0 && (!fan->rvec0 || !fan->rindex_vec0))) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_fan_upload: bad fan");
    CK(cudaSetDevice(cx().device));
    const size_t n = (size_t)fan->nray;
    CK(cx().rvec0.reserve(3 * n)); CK(cx().nvec0.reserve(3 * n)); CK(cx().wt.reserve(n));
    if (n) {
        CK(cudaMemcpyAsync(cx().rvec0.p, fan->rvec0, 3 * n * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
        CK(cudaMemcpyAsync(cx().nvec0.p, fan->rindex_vec0, 3 * n * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
        if (fan->ray_pwr_wt) CK(cudaMemcpyAsync(cx().wt.p, fan->ray_pwr_wt, n * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
        else CK(cudaMemsetAsync(cx().wt.p, 0, n * sizeof(double), cx().stream));
    }
    CK(cudaStreamSynchronize(cx().stream));
    cx().nray = fan->nray;
    double wsum = 0.0;
    if (fan->ray_pwr_wt) for (size_t i = 0; i < n; ++i) wsum += std::fabs(fan->ray_pwr_wt[i]);
    cx().fan_weight = wsum;
    return 0;
}

int rays_b200_fan_download(double *rvec0, double *rindex_vec0, double *ray_pwr_wt) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    const size_t n = (size_t)cx().nray;
    if (n == 0) return 0;
    if (rvec0) CK(cudaMemcpyAsync(rvec0, cx().rvec0.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, cx().stream));
    if (rindex_vec0) CK(cudaMemcpyAsync(rindex_vec0, cx().nvec0.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, cx().stream));
    if (ray_pwr_wt) CK(cudaMemcpyAsync(ray_pwr_wt, cx().wt.p, n * sizeof(double), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    return 0;
}

int rays_b200_fan_shard(int rank, int world) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (world < 1 || rank < 0 || rank >= world) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_fan_shard: bad rank/world");
    if (world == 1 || cx().nray == 0) return 0;
    const long long n_out = (cx().nray - rank + world - 1) / world;
    DevBuf<double> rv, nv, w;
    CK(rv.reserve((size_t)3 * std::max<long long>(n_out, 1))); CK(nv.reserve((size_t)3 * std::max<long long>(n_out, 1))); CK(w.reserve((size_t)std::max<long long>(n_out, 1)));
    if (n_out > 0) {
        fan_shard_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, cx().stream>>>(n_out, rank, world, cx().rvec0.p, cx().nvec0.p, cx().wt.p, rv.p, nv.p, w.p);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(cx().stream));
    std::swap(cx().rvec0, rv); std::swap(cx().nvec0, nv); std::swap(cx().wt, w);
    rv.release(); nv.release(); w.release();
    cx().nray = n_out;
    return 0;
}

// unit of the fixed-point deposition bins: the absorbed power of a fan cannot exceed the sum of its ray weights, so with
// 2^e > sum |ray_pwr_wt| every partial sum stays below 2^62 units of 2^(e-62)
static void set_dep_scale() {
    int e = 0;
    if (cx().fan_weight > 0.0 && std::isfinite(cx().fan_weight)) e = std::ilogb(cx().fan_weight) + 1;
    cx().dep_scale = std::ldexp(1.0, 62 - e);
}

static int trace_device_impl(int store_trajectories, bool binned) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    CK(cudaSetDevice(cx().device));
    const rays_cfg &c = cx().dc.c;
    const int npa = c.nstep_max + 1;
    if (store_trajectories && (cx().ray_vec.n < (size_t)cx().nray * npa * c.nv || cx().residual.n < (size_t)cx().nray * npa)) {   // buffers must grow
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const double need = (double)cx().nray * npa * (c.nv + 1) * 8.0;
        const double have = (double)free_b + (double)(cx().ray_vec.n + cx().residual.n) * 8.0;
        if (need > 0.92 * have)
            return set_err(RAYS_ERR_ALLOC, "rays_b200_trace_device: trajectories of this fan do not fit in HBM; trace without storage (binned) or use rays_b200_trace, which batches");
    }
    int rc = ensure_results(cx().nray, c.nv, npa, store_trajectories != 0);
    if (rc) return rc;
    cx().last_launches = 0; cx().last_first_ms = 0; cx().last_resume_ms = 0; cx().last_phases = 0;
    CK(cudaMemsetAsync(cx().queue.p, 0, 3 * sizeof(unsigned long long), cx().stream));
    if (binned) CK(cudaMemsetAsync(cx().dep.p, 0, (size_t)cx().dep_bins * sizeof(unsigned long long), cx().stream));
    CK(cudaEventRecord(cx().ev0, cx().stream));
    if (cx().nray > 0) {
        rc = launch_trace(0, cx().nray, store_trajectories ? cx().ray_vec.p : nullptr, store_trajectories ? cx().residual.p : nullptr, binned);
        if (rc) return rc;
    }
    CK(cudaEventRecord(cx().ev1, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, cx().ev0, cx().ev1));
    cx().last_ms = ms;
    cx().dep_fused = binned;
    return fetch_counters();
}

int rays_b200_trace_device(int store_trajectories) { return trace_device_impl(store_trajectories, false); }

int rays_b200_trace_device_binned(int n_bins, double grid_min, double grid_max, int store_trajectories) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    if (n_bins < 1 || !(grid_max > grid_min)) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: bad grid");
    const rays_cfg &c = cx().dc.c;
    if (c.damping_model == RAYS_DAMP_NONE) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: needs a damping model (v(8) = absorbed power)");
    if (c.equilib_model != RAYS_EQ_SLAB && c.equilib_model != RAYS_EQ_AXISYM_TOROID)
        return set_err(RAYS_ERR_INVALID_CONFIG, "deposition profiles exist for slab (Ptotal_x) and axisym_toroid (Ptotal_psi) only");
    CK(cx().dep.reserve((size_t)n_bins));
    cx().dep_bins = n_bins; cx().dep_min = grid_min; cx().dep_max = grid_max;
    set_dep_scale();
    return trace_device_impl(store_trajectories, true);
}

int rays_b200_last_trace_stats(double *kernel_ms, int64_t *ray_steps, int32_t *n_launches) {
    if (kernel_ms) *kernel_ms = cx().last_ms;
    if (ray_steps) *ray_steps = cx().last_steps;
    if (n_launches) *n_launches = cx().last_launches;
    return 0;
}
// extra stats for bench/profiles: RHS evaluations, kernel name, grid, resident CTAs per SM
int rays_b200_last_trace_info(int64_t *rhs_evals, char *kernel_name, int name_len, int32_t *grid, int32_t *blocks_per_sm) {
    if (rhs_evals) *rhs_evals = cx().last_rhs;
    if (kernel_name && name_len > 0) { std::strncpy(kernel_name, cx().last_kernel, (size_t)name_len - 1); kernel_name[name_len - 1] = 0; }
    if (grid) *grid = cx().last_grid;
    if (blocks_per_sm) *blocks_per_sm = cx().last_bps;
    return 0;
}

// kernel time of the last trace: first pass over the fan, resume passes over the suspended rays, launch count
int rays_b200_last_trace_breakdown(double *first_pass_ms, double *resume_pass_ms, int32_t *n_passes) {
    if (first_pass_ms) *first_pass_ms = cx().last_first_ms;
    if (resume_pass_ms) *resume_pass_ms = cx().last_resume_ms;
    if (n_passes) *n_passes = cx().last_phases;
    return 0;
}

int rays_b200_results_download(rays_results *res) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!res) return set_err(RAYS_ERR_INVALID_CONFIG, "null results");
    if (res->nray < cx().res_nray || res->nv != cx().res_nv) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_results_download: result arrays do not match the last trace");
    const long long n = cx().res_nray;
    const int nv = cx().res_nv;
    int rc = copy_small_results(res, 0, n);
    if (rc) return rc;
    std::vector<int> np((size_t)n), codes((size_t)n);
    if (n) {
        CK(cudaMemcpyAsync(np.data(), cx().npoints.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaMemcpyAsync(codes.data(), cx().stop.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
    }
    CK(cudaStreamSynchronize(cx().stream));
    if (cx().have_traj && n && (res->ray_vec || res->residual)) {
        if (res->npoints_alloc < 1) return set_err(RAYS_ERR_INVALID_CONFIG, "npoints_alloc < 1");
        for (long long i = 0; i < n; ++i)
            if (np[(size_t)i] > res->npoints_alloc) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_results_download: npoints_alloc smaller than the longest ray");
        if (staged_copy_wanted(res)) {     // pageable arrays: packed rows through the page-locked ring (the stream is idle: synchronised above)
            if ((rc = copy_trajectories_staged(res, np, 0, n, cx().ray_vec.p, cx().residual.p, cx().res_npa, nv))) return rc;
        } else {
            zero_tail_kernel<<<(unsigned)n, 64, 0, cx().stream>>>(res->ray_vec ? cx().ray_vec.p : nullptr, res->residual ? cx().residual.p : nullptr, cx().npoints.p, cx().res_npa, nv, n);
            CK(cudaGetLastError());
            if ((rc = copy_trajectories(res, np, 0, n, cx().ray_vec.p, cx().residual.p, cx().res_npa, nv, cx().stream))) return rc;
            CK(cudaStreamSynchronize(cx().stream));
        }
    }
    fill_flags(res, codes, 0);
    res->total_trace_time = cx().last_ms * 1e-3;
    res->total_ray_steps = cx().last_steps;
    if (res->ray_trace_time) for (long long i = 0; i < n; ++i) res->ray_trace_time[i] = n ? res->total_trace_time / (double)n : 0.0;
    return 0;
}

// The copier path (RAYS_B200_COPIER=0 selects the in-kernel streaming copy-out instead) needs the whole fan's trajectories in HBM
// (nray x (nstep_max+1) x (nv+1) doubles: 67 GB for the 1M-ray bench fan) and is built into the RK4 kernels (an SG trace is
// integration-bound: its in-kernel copy-out already hides behind the arithmetic).
static bool copier_wanted(const rays_cfg &c, long long n, int npa, int nv) {
    if (c.ode_solver != RAYS_ODE_RK4 || n <= 0 || n >= (1LL << 31)) return false;
    if (const char *env = getenv("RAYS_B200_COPIER")) { if (env[0] == '0') return false; }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return false; }
    const double have = (double)free_b + (double)(cx().ray_vec.n + cx().residual.n) * 8.0;
    return (double)n * npa * (nv + 1) * 8.0 + (double)n * 64.0 <= 0.85 * have;
}

// Is [p, p+bytes) page-locked host memory the device can address?  Arrays from rays_b200_host_alloc (or
// registered by the caller) are used as they are; small pageable arrays (<= 1 GiB) are registered for the
// duration of the call; larger pageable arrays take the batched fallback, which only touches the pages
// that receive data.
static double *device_view_of_host(double *p, size_t bytes, std::vector<void *> &temp_registered) {
    if (!p) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeUnregistered) {
        const char *env = getenv("RAYS_B200_REGISTER_HOST");
        if ((env && env[0] == '0') || bytes > (size_t(1) << 30)) return nullptr;
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        temp_registered.push_back((void *)p);
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    cudaPointerAttributes at2{};   // the whole range must be page-locked
    if (cudaPointerGetAttributes(&at2, (char *)p + bytes - 1) != cudaSuccess || at2.type != cudaMemoryTypeHost) { cudaGetLastError(); return nullptr; }
    return (double *)at.devicePointer;
}

// trace_rays with HOST buffers (ray_tracing.f90:1-290): H2D of the fan, trace in batches sized to HBM with
// two trajectory buffers so that the copy-out of batch b overlaps the integration of batch b+1, results
// land in the caller's arrays in the reference layout.
static int trace_host_impl(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, RowMap rm);
int rays_b200_trace(const rays_cfg *cfg, const rays_fan *fan, rays_results *res) { return trace_host_impl(cfg, fan, res, RowMap{}); }
// res describes the WHOLE result arrays (res->nray rows); the rays of `fan` go to the rows rm selects
static int trace_host_impl(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, RowMap rm) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!fan || !res) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace: null argument");
    int rc;
    if (cfg) { if ((rc = rays_b200_set_config(cfg))) return rc; }
    else if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_trace: no config");
    const rays_cfg &c = cx().dc.c;
    if ((fan->nray > 0 && res->nray < rm.row0 + (fan->nray - 1) * rm.stride + 1) || res->nv != c.nv) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace: result arrays do not match fan/nv");
    if (!res->npoints) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace: npoints array is required");
    const bool want_traj = res->ray_vec != nullptr || res->residual != nullptr;
    const int npa = c.nstep_max + 1;
    if (want_traj && res->npoints_alloc < npa) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace: npoints_alloc < nstep_max+1");
    CK(cudaEventRecord(cx().ev0, cx().stream));
    if ((rc = rays_b200_fan_upload(fan))) return rc;
    const long long n = cx().nray;
    const int nv = c.nv;
    // Preferred path: the result arrays are (or can be made) page-locked: finished rays are copied out by
    // the trace kernel itself, overlapped with the integration; HBM holds only one staging row per lane.
    HostOut hv;
    bool streaming = false;
    std::vector<void *> temp_registered;
    struct Unreg { std::vector<void *> &v; ~Unreg() { for (void *q : v) cudaHostUnregister(q); } } unreg{temp_registered};
    if (want_traj && n > 0) {
        hv.npa = res->npoints_alloc; hv.row0 = rm.row0; hv.stride = rm.stride;
        hv.ray_vec = device_view_of_host(res->ray_vec, (size_t)res->nray * res->npoints_alloc * nv * 8, temp_registered);
        hv.residual = device_view_of_host(res->residual, (size_t)res->nray * res->npoints_alloc * 8, temp_registered);
        streaming = (res->ray_vec == nullptr || hv.ray_vec) && (res->residual == nullptr || hv.residual);
    }
    cx().last_copier = 0;
    if (streaming && copier_wanted(c, n, npa, nv)) {
        // Copier path (see copy_out_kernel): trajectories of the whole fan to HBM at full speed, a concurrent kernel on the copy
        // stream delivers every ray to the caller's arrays as soon as it has ended.
        if ((rc = ensure_results(n, nv, npa, true))) return rc;
        cx().have_traj = hv.ray_vec != nullptr && hv.residual != nullptr;   // the device copy is complete only if both arrays were asked for
        CK(cx().done_list.reserve((size_t)n));
        CK(cx().copy_ctl.reserve(8));
        cx().last_launches = 0; cx().last_first_ms = 0; cx().last_resume_ms = 0; cx().last_phases = 0;
        if (!cx().copier_started) CK(cudaHostAlloc((void **)&cx().copier_started, sizeof(int), cudaHostAllocMapped));
        if ((rc = launch_trace(0, n, hv.ray_vec ? cx().ray_vec.p : nullptr, hv.residual ? cx().residual.p : nullptr, false, nullptr, true))) return rc;
        CK(cudaMemsetAsync(cx().queue.p, 0, 3 * sizeof(unsigned long long), cx().stream));
        CK(cudaMemsetAsync(cx().done_list.p, 0xFF, (size_t)n * sizeof(int), cx().stream));
        CK(cudaMemsetAsync(cx().copy_ctl.p, 0, 8 * sizeof(unsigned long long), cx().stream));
        CK(cudaEventRecord(cx().ev_batch[0], cx().stream));
        CK(cudaStreamWaitEvent(cx().copy_stream, cx().ev_batch[0], 0));
        CopyOutArgs ca{};
        ca.nray = n; ca.done_list = cx().done_list.p; ca.claim = cx().copy_ctl.p + 1;
        ca.abort_flag = reinterpret_cast<const int *>(cx().copy_ctl.p + 2); ca.deferred_flag = reinterpret_cast<int *>(cx().copy_ctl.p + 3);
        ca.trace_started = reinterpret_cast<const int *>(cx().copy_ctl.p + 4);
        ca.npoints = cx().npoints.p;
        ca.ray_vec = hv.ray_vec ? cx().ray_vec.p : nullptr; ca.residual = hv.residual ? cx().residual.p : nullptr;
        ca.host_ray_vec = hv.ray_vec; ca.host_residual = hv.residual;
        ca.npoints_alloc = npa; ca.host_npoints_alloc = hv.npa; ca.nv = nv; ca.host_ray0 = hv.row0; ca.host_ray_stride = hv.stride;
        int copy_ctas = 16;
        if (const char *env = getenv("RAYS_B200_COPY_CTAS")) { if (env[0] && atoi(env) > 0) copy_ctas = atoi(env); }
        *reinterpret_cast<volatile int *>(cx().copier_started) = 0;
        CK(cudaHostGetDevicePointer((void **)&ca.started, cx().copier_started, 0));
        copy_out_kernel<<<copy_ctas, 256, 0, cx().copy_stream>>>(ca);
        CK(cudaGetLastError());
        {   // the copier's CTAs must be resident before the trace kernel fills the SMs (they would otherwise start when it ends)
            const auto t0 = std::chrono::steady_clock::now();
            while (*reinterpret_cast<volatile int *>(cx().copier_started) < copy_ctas &&
                   std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(200)) { }
        }
        cx().copier_on = true;
        rc = launch_trace(0, n, hv.ray_vec ? cx().ray_vec.p : nullptr, hv.residual ? cx().residual.p : nullptr, false, nullptr);
        cx().copier_on = false;
        if (rc) {   // release the copier before reporting the failure
            const int one = 1;
            cudaMemcpyAsync(cx().copy_ctl.p + 2, &one, sizeof(int), cudaMemcpyHostToDevice, cx().stream);
            cudaStreamSynchronize(cx().stream); cudaStreamSynchronize(cx().copy_stream);
            return rc;
        }
        cx().last_launches += 1;
        std::vector<int> codes((size_t)n);
        CK(cudaMemcpyAsync(codes.data(), cx().stop.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        if ((rc = copy_small_results(res, 0, n, rm))) return rc;
        CK(cudaEventRecord(cx().ev_copy[0], cx().copy_stream));
        CK(cudaStreamWaitEvent(cx().stream, cx().ev_copy[0], 0));
        int deferred = 0;
        CK(cudaMemcpyAsync(&deferred, cx().copy_ctl.p + 3, sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        if (deferred) {   // the copier stopped waiting (kernels serialised by a tool, or a very long quiet stretch): every ray has ended by now
            const int one = 1;
            CK(cudaMemsetAsync(cx().copy_ctl.p + 1, 0, 3 * sizeof(unsigned long long), cx().stream));   // claim, abort, deferred
            CK(cudaMemcpyAsync(cx().copy_ctl.p + 4, &one, sizeof(int), cudaMemcpyHostToDevice, cx().stream));
            ca.started = nullptr;
            copy_out_kernel<<<4 * cx().num_sms, 256, 0, cx().stream>>>(ca);
            CK(cudaGetLastError());
            cx().last_launches += 1;
            CK(cudaMemcpyAsync(&deferred, cx().copy_ctl.p + 3, sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        }
        CK(cudaEventRecord(cx().ev1, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        if (deferred) return set_err(RAYS_ERR_CUDA, "rays_b200_trace: a ray never reached the copy-out list");
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, cx().ev0, cx().ev1));
        cx().last_ms = ms;
        if ((rc = fetch_counters())) return rc;
        fill_flags(res, codes, 0, rm);
        res->total_trace_time = cx().last_ms * 1e-3;
        res->total_ray_steps = cx().last_steps;
        if (res->ray_trace_time) for (long long i = 0; i < n; ++i) res->ray_trace_time[rm.row0 + i * rm.stride] = res->total_trace_time / (double)n;
        cx().dep_fused = false;
        cx().last_copier = 1;
        cx().kernel_name_buf = std::string(cx().last_kernel ? cx().last_kernel : "") + " + copy_out_kernel";
        cx().last_kernel = cx().kernel_name_buf.c_str();
        return 0;
    }
    if (streaming) {
        if ((rc = ensure_results(n, nv, npa, false))) return rc;
        cx().have_traj = false;
        cx().last_launches = 0; cx().last_first_ms = 0; cx().last_resume_ms = 0; cx().last_phases = 0;
        CK(cudaMemsetAsync(cx().queue.p, 0, 3 * sizeof(unsigned long long), cx().stream));
        if ((rc = launch_trace(0, n, nullptr, nullptr, false, &hv))) return rc;
        std::vector<int> codes((size_t)n);
        CK(cudaMemcpyAsync(codes.data(), cx().stop.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        if ((rc = copy_small_results(res, 0, n, rm))) return rc;
        CK(cudaEventRecord(cx().ev1, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, cx().ev0, cx().ev1));
        cx().last_ms = ms;
        if ((rc = fetch_counters())) return rc;
        fill_flags(res, codes, 0, rm);
        res->total_trace_time = cx().last_ms * 1e-3;
        res->total_ray_steps = cx().last_steps;
        if (res->ray_trace_time) for (long long i = 0; i < n; ++i) res->ray_trace_time[rm.row0 + i * rm.stride] = res->total_trace_time / (double)n;
        cx().dep_fused = false;
        return 0;
    }
    // Fallback (pageable arrays that cannot be registered): batches sized to HBM, two trajectory buffers
    long long batch = std::max<long long>(n, 1);
    if (want_traj && n > 0) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const double avail = 0.70 * ((double)free_b + (double)(cx().ray_vec.n + cx().residual.n) * 8.0);
        const double per_ray = (double)npa * (nv + 1) * 8.0;
        if ((double)n * per_ray > avail) {
            batch = (long long)(avail / (2.0 * per_ray));
            batch = std::max<long long>(batch / kTraceBlock * kTraceBlock, kTraceBlock);
        }
    }
    const int nbuf = (batch < n) ? 2 : 1;
    const bool staged = want_traj && staged_copy_wanted(res);
    if ((rc = ensure_results(n, nv, npa, false))) return rc;
    if (want_traj) {
        CK(cx().ray_vec.reserve((size_t)nbuf * batch * npa * nv));
        CK(cx().residual.reserve((size_t)nbuf * batch * npa));
        cx().have_traj = (nbuf == 1);
    }
    cx().last_launches = 0; cx().last_first_ms = 0; cx().last_resume_ms = 0; cx().last_phases = 0;
    CK(cudaMemsetAsync(cx().queue.p, 0, 3 * sizeof(unsigned long long), cx().stream));
    std::vector<int> np((size_t)std::max<long long>(n, 1)), codes((size_t)std::max<long long>(n, 1));
    int ib = 0;
    for (long long first = 0; first < n; first += batch, ++ib) {
        const long long count = std::min(batch, n - first);
        const int b = ib % nbuf;
        double *tv = want_traj ? cx().ray_vec.p + (size_t)b * batch * npa * nv : nullptr;
        double *tr = want_traj ? cx().residual.p + (size_t)b * batch * npa : nullptr;
        if (ib >= nbuf) CK(cudaStreamWaitEvent(cx().stream, cx().ev_copy[b], 0));   // buffer b free again
        if ((rc = launch_trace(first, count, tv, tr, false))) return rc;
        if (want_traj && !staged) {
            zero_tail_kernel<<<(unsigned)count, 64, 0, cx().stream>>>(res->ray_vec ? tv : nullptr, res->residual ? tr : nullptr, cx().npoints.p + first, npa, nv, count);
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(np.data() + first, cx().npoints.p + first, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaMemcpyAsync(codes.data() + first, cx().stop.p + first, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
        CK(cudaEventRecord(cx().ev_batch[b], cx().stream));
        if (want_traj) {
            CK(cudaEventSynchronize(cx().ev_batch[b]));   // npoints of this batch are on the host: trim the copy
            if (staged) rc = copy_trajectories_staged(res, np, first, count, tv, tr, npa, nv, rm);   // returns with the batch delivered
            else rc = copy_trajectories(res, np, first, count, tv, tr, npa, nv, cx().copy_stream, rm);
            if (rc) return rc;
            CK(cudaEventRecord(cx().ev_copy[b], cx().copy_stream));
        }
    }
    if ((rc = copy_small_results(res, 0, n, rm))) return rc;
    CK(cudaEventRecord(cx().ev1, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    CK(cudaStreamSynchronize(cx().copy_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, cx().ev0, cx().ev1));
    cx().last_ms = ms;
    if ((rc = fetch_counters())) return rc;
    if (res->ray_stop_code == nullptr) { /* codes only needed for the strings */ }
    codes.resize((size_t)n);
    fill_flags(res, codes, 0, rm);
    res->total_trace_time = cx().last_ms * 1e-3;
    res->total_ray_steps = cx().last_steps;
    if (res->ray_trace_time) for (long long i = 0; i < n; ++i) res->ray_trace_time[rm.row0 + i * rm.stride] = n ? res->total_trace_time / (double)n : 0.0;
    cx().dep_fused = false;
    return 0;
}

// ======================= launch fans =========================================================================
static int run_launch_fan(int kind, long long npos, const std::vector<double> &pos_host, const double *dir_host, int n_a, int n_b, double a0,
                          double da, double b0, double db, int64_t *nray_out) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    CK(cudaSetDevice(cx().device));
    const rays_cfg &c = cx().dc.c;
    const TuOps *ops = tu_ops(c.equilib_model, RAYS_ODE_RK4);
    const long long ncand = kind == 4 ? npos : npos * n_a * n_b;
    if (ncand <= 0) { cx().nray = 0; if (nray_out) *nray_out = 0; return 0; }
    if (ncand >= (1LL << 31)) return set_err(RAYS_ERR_INVALID_CONFIG, "launch fan: more than 2^31 candidates on one GPU");
    DevBuf<double> pos, dir, rv, nv;
    DevBuf<int> valid, counts;
    DevBuf<long long> total;
    CK(pos.reserve(pos_host.size()));
    CK(cudaMemcpyAsync(pos.p, pos_host.data(), pos_host.size() * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
    if (kind == 4) {
        CK(dir.reserve((size_t)3 * ncand));
        CK(cudaMemcpyAsync(dir.p, dir_host, (size_t)3 * ncand * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
    }
    const int nblocks = (int)((ncand + kScanBlock - 1) / kScanBlock);
    CK(rv.reserve((size_t)3 * ncand)); CK(nv.reserve((size_t)3 * ncand)); CK(valid.reserve((size_t)ncand));
    CK(counts.reserve((size_t)nblocks)); CK(total.reserve(1));
    FanLaunchArgs f{};
    f.kind = kind; f.ncand = ncand; f.rvec_in = pos.p; f.nvec_in = dir.p; f.n_a = n_a; f.n_b = n_b;
    f.a0 = a0; f.da = da; f.b0 = b0; f.db = db; f.rvec_out = rv.p; f.nvec_out = nv.p; f.valid = valid.p;
    CK(ops->launch_fan(cx().sel, f, cx().stream));
    fan_count_kernel<<<nblocks, kScanBlock, 0, cx().stream>>>(valid.p, ncand, counts.p);
    CK(cudaGetLastError());
    fan_offsets_kernel<<<1, 1024, 0, cx().stream>>>(counts.p, nblocks, total.p);
    CK(cudaGetLastError());
    long long nray = 0;
    CK(cudaMemcpyAsync(&nray, total.p, sizeof(long long), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    CK(cx().rvec0.reserve((size_t)3 * std::max<long long>(nray, 1))); CK(cx().nvec0.reserve((size_t)3 * std::max<long long>(nray, 1)));
    CK(cx().wt.reserve((size_t)std::max<long long>(nray, 1)));
    if (nray > 0) {
        fan_scatter_kernel<<<nblocks, kScanBlock, 0, cx().stream>>>(valid.p, ncand, counts.p, rv.p, nv.p, cx().rvec0.p, cx().nvec0.p);
        CK(cudaGetLastError());
        double w = 1.0 / (double)nray;
        if (kind == 1) w = 1.0 / (double)nray / (double)nray;   // (R) divided by nray twice (simple_slab_ray_init_m.f90:179,182)
        fill_kernel<<<(unsigned)((nray + 255) / 256), 256, 0, cx().stream>>>(cx().wt.p, nray, w);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(cx().stream));
    pos.release(); dir.release(); rv.release(); nv.release(); valid.release(); counts.release(); total.release();
    cx().nray = nray;
    cx().fan_weight = nray > 0 ? (kind == 1 ? 1.0 / (double)nray : 1.0) : 0.0;
    if (nray_out) *nray_out = nray;
    return 0;
}

int rays_b200_launch_fan_slab(const rays_slab_launch *p, int64_t *nray_out) {
    if (!p) return set_err(RAYS_ERR_INVALID_CONFIG, "null launch parameters");
    if (cx().cfg_set && cx().dc.c.equilib_model != RAYS_EQ_SLAB) return set_err(RAYS_ERR_INVALID_CONFIG, "simple_slab ray init needs equilib_model 'slab'");
    std::vector<double> pos;
    for (int iz = 1; iz <= p->n_z_launch; ++iz)
        for (int iy = 1; iy <= p->n_y_launch; ++iy)
            for (int ix = 1; ix <= p->n_x_launch; ++ix) {
                pos.push_back(p->x_launch0 + (ix - 1) * p->dx_launch);
                pos.push_back(p->y_launch0 + (iy - 1) * p->dy_launch);
                pos.push_back(p->z_launch0 + (iz - 1) * p->dy_launch);   // (R) z uses dy_launch (simple_slab_ray_init_m.f90:122)
            }
    return run_launch_fan(1, (long long)pos.size() / 3, pos, nullptr, p->n_ky_launch, p->n_kz_launch, p->rindex_y0, p->delta_rindex_y0, p->rindex_z0,
                          p->delta_rindex_z0, nray_out);
}
int rays_b200_launch_fan_solovev(const rays_solovev_launch *p, int64_t *nray_out) {
    if (!p) return set_err(RAYS_ERR_INVALID_CONFIG, "null launch parameters");
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    if (cx().dc.c.equilib_model != RAYS_EQ_SOLOVEV) return set_err(RAYS_ERR_INVALID_CONFIG, "solovev ray init needs equilib_model 'solovev'");
    std::vector<double> pos;
    const double rmaj = cx().dc.c.solovev.rmaj;
    for (int ir = 1; ir <= p->n_r_launch; ++ir)
        for (int it = 1; it <= p->n_theta_launch; ++it) {
            const double theta = p->theta_launch0 + (it - 1) * p->dtheta_launch;
            const double rmin_launch = p->r_launch0 + (ir - 1) * p->dr_launch;
            pos.push_back(rmaj + rmin_launch * std::cos(theta));
            pos.push_back(0.0);
            pos.push_back(rmin_launch * std::sin(theta));
        }
    return run_launch_fan(2, (long long)pos.size() / 3, pos, nullptr, p->n_rindex_theta, p->n_rindex_phi, p->rindex_theta0, p->delta_rindex_theta,
                          p->rindex_phi0, p->delta_rindex_phi, nray_out);
}
int rays_b200_launch_fan_axisym(const rays_axisym_launch *p, int64_t *nray_out) {
    if (!p) return set_err(RAYS_ERR_INVALID_CONFIG, "null launch parameters");
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    if (cx().dc.c.equilib_model != RAYS_EQ_AXISYM_TOROID) return set_err(RAYS_ERR_INVALID_CONFIG, "axisym_toroid ray init needs equilib_model 'axisym_toroid'");
    std::vector<double> pos;
    for (int iR = 1; iR <= p->n_R_launch; ++iR)
        for (int iZ = 1; iZ <= p->n_Z_launch; ++iZ) {   // (R) every (i_R, i_Z) launches from the same point
            pos.push_back(p->R_launch0); pos.push_back(0.0); pos.push_back(p->Z_launch0);
        }
    return run_launch_fan(3, (long long)pos.size() / 3, pos, nullptr, p->n_rindex_theta, p->n_rindex_phi, p->rindex_theta0, p->delta_rindex_theta,
                          p->rindex_phi0, p->delta_rindex_phi, nray_out);
}
int rays_b200_launch_fan_directions(int64_t n_in, const double *rvec_in, const double *nvec_in, int64_t *nray_out) {
    if (n_in < 0 || (n_in > 0 && (!rvec_in || !nvec_in))) return set_err(RAYS_ERR_INVALID_CONFIG, "launch fan: bad input arrays");
    std::vector<double> pos(rvec_in, rvec_in + 3 * n_in);
    return run_launch_fan(4, n_in, pos, nvec_in, 1, 1, 0, 0, 0, 0, nray_out);
}

// ======================= deposition ==========================================================================
// calculate_deposition_profiles (deposition_profiles_m.f90:228-292).  The device bins into 64-bit fixed-point accumulators
// (TraceArgs::dep_acc); acc_out (host, n_bins int64, optional) and d_acc_out (device, optional) receive the raw bins of THIS
// GPU for an exact integer reduction across GPUs, *unit the value of one count.
static int deposition_impl(rays_deposition *dep, int64_t *acc_out, void *d_acc_out, double *unit) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!dep || dep->n_bins < 1) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: bad arguments");
    const rays_cfg &c = cx().dc.c;
    const int nb = dep->n_bins;
    if (cx().dep_fused) {
        if (nb != cx().dep_bins || dep->grid_min != cx().dep_min || dep->grid_max != cx().dep_max)
            return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: n_bins / grid differ from the binned trace");
    } else {
        if (!cx().have_traj) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: no stored trajectories; trace with storage or use rays_b200_trace_device_binned");
        if (c.damping_model == RAYS_DAMP_NONE) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: needs a damping model");
        if (c.equilib_model != RAYS_EQ_SLAB && c.equilib_model != RAYS_EQ_AXISYM_TOROID)
            return set_err(RAYS_ERR_INVALID_CONFIG, "deposition profiles exist for slab (Ptotal_x) and axisym_toroid (Ptotal_psi) only");
        if (!(dep->grid_max > dep->grid_min)) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: bad grid");
        CK(cx().dep.reserve((size_t)nb));
        cx().dep_bins = nb; cx().dep_min = dep->grid_min; cx().dep_max = dep->grid_max;
        set_dep_scale();
        CK(cudaMemsetAsync(cx().dep.p, 0, (size_t)nb * sizeof(unsigned long long), cx().stream));
        if (cx().res_nray > 0) {
            const unsigned grid = (unsigned)((cx().res_nray + 127) / 128);
            const DepBins bins{cx().dep.p, nb, dep->grid_min, dep->grid_max, cx().dep_scale};
            if (c.equilib_model == RAYS_EQ_SLAB)
                deposition_kernel<RAYS_EQ_SLAB><<<grid, 128, 0, cx().stream>>>(cx().res_nray, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, cx().pwr.p, bins);
            else
                deposition_kernel<RAYS_EQ_AXISYM_TOROID><<<grid, 128, 0, cx().stream>>>(cx().res_nray, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, cx().pwr.p, bins);
            CK(cudaGetLastError());
        }
    }
    std::vector<long long> h((size_t)nb);
    CK(cudaMemcpyAsync(h.data(), cx().dep.p, (size_t)nb * sizeof(long long), cudaMemcpyDeviceToHost, cx().stream));
    if (d_acc_out) CK(cudaMemcpyAsync(d_acc_out, cx().dep.p, (size_t)nb * sizeof(long long), cudaMemcpyDeviceToDevice, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    const double u = 1.0 / cx().dep_scale;   // a power of two: the conversions below are exact scalings
    long long q = 0;
    for (int b = 0; b < nb; ++b) {
        if (dep->profile) dep->profile[b] = (double)h[(size_t)b] * u;
        if (acc_out) acc_out[b] = (int64_t)h[(size_t)b];
        q += h[(size_t)b];
    }
    dep->Q_sum = (double)q * u;
    if (unit) *unit = u;
    return 0;
}
int rays_b200_deposition(rays_deposition *dep, double *d_profile_out) {
    int rc = deposition_impl(dep, nullptr, nullptr, nullptr);
    if (rc || !d_profile_out) return rc;
    std::vector<double> h((size_t)dep->n_bins + 1);   // this GPU's partial profile + Q_sum as doubles (exact: power-of-two unit)
    std::vector<long long> raw((size_t)dep->n_bins);
    CK(cudaMemcpyAsync(raw.data(), cx().dep.p, raw.size() * sizeof(long long), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    for (int b = 0; b < dep->n_bins; ++b) h[(size_t)b] = (double)raw[(size_t)b] / cx().dep_scale;
    h[(size_t)dep->n_bins] = dep->Q_sum;
    CK(cudaMemcpyAsync(d_profile_out, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    return 0;
}
int rays_b200_deposition_fixed(rays_deposition *dep, int64_t *acc_out, void *d_acc_out, double *unit) { return deposition_impl(dep, acc_out, d_acc_out, unit); }
int rays_b200_deposition_set_total_weight(double total_weight) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!(total_weight >= 0.0)) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: total weight must be >= 0");
    cx().fan_weight = total_weight;
    return 0;
}

int rays_b200_summaries_pack(void *d_out, int64_t rows_capacity, int64_t *rows, int32_t *row_doubles) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (rows) *rows = cx().res_nray;
    if (row_doubles) *row_doubles = 6 + 2 * cx().res_nv;
    if (!d_out) return 0;
    if (rows_capacity < cx().res_nray) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_summaries_pack: output smaller than the last trace");
    CK(cudaSetDevice(cx().device));
    if (cx().res_nray > 0) {
        pack_summaries_kernel<<<(unsigned)((cx().res_nray + 255) / 256), 256, 0, cx().stream>>>(cx().res_nray, cx().res_nv, cx().npoints.p, cx().stop.p, cx().pwr.p, cx().endres.p, cx().maxres.p,
                                                                                             cx().endpar.p, cx().startv.p, cx().endv.p, (double *)d_out);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(cx().stream));
    return 0;
}

// ======================= one host process, several GPUs ==========================================================
// program rays is a single process (RAYS_code/RAYS.f90:9-15), so the library drives the GPUs of the box itself: one context and
// one host worker thread per GPU, the fan sharded iray mod ngpu (ray length varies smoothly along the launch loops),
// no traffic during integration; afterwards ONE ncclReduce of the fixed-point deposition bins and ONE ncclAllGather of the packed
// per-ray summaries over an NCCL communicator created with ncclCommInitAll.  NCCL is loaded with dlopen (libnccl.so.2 of the
// image, or the one already mapped into the process), so a single-GPU host needs no NCCL at all.
}  // extern "C"
namespace {
struct NcclApi {
    void *h = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} nccl;
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0;   // nccl.h: ncclDataType_t / ncclRedOp_t
struct MultiCtx {
    int n = 0;
    void *comm[kMaxGpus] = {};
    DevBuf<long long> acc[kMaxGpus];      // per GPU: its fixed-point bins (GPU 0: the reduced profile)
    DevBuf<double> summ[kMaxGpus], gathered[kMaxGpus];
    long long gathered_rows_per_gpu = 0;
    int gathered_row_doubles = 0;
} mg;

int load_nccl() {
    if (nccl.h) return 0;
    const char *cands[] = {getenv("RAYS_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *c : cands) { if (c && c[0] && (h = dlopen(c, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!h) return set_err(RAYS_ERR_CUDA, std::string("rays_b200_init_multi: cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : ""));
    auto sym = [&](const char *n) { return dlsym(h, n); };
    nccl.CommInitAll = (int (*)(void **, int, const int *))sym("ncclCommInitAll");
    nccl.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
    nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    nccl.Reduce = (int (*)(const void *, void *, size_t, int, int, int, void *, cudaStream_t))sym("ncclReduce");
    nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))sym("ncclAllGather");
    nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    if (!nccl.CommInitAll || !nccl.CommDestroy || !nccl.GroupStart || !nccl.GroupEnd || !nccl.Reduce || !nccl.AllGather)
        return set_err(RAYS_ERR_CUDA, "rays_b200_init_multi: NCCL library lacks a required symbol");
    nccl.h = h;
    return 0;
}
#define NCK(call)                                                                                                   \
    do {                                                                                                            \
        int r_ = (call);                                                                                            \
        if (r_ != 0) return set_err(RAYS_ERR_CUDA, std::string(#call) + ": " + (nccl.GetErrorString ? nccl.GetErrorString(r_) : "NCCL error")); \
    } while (0)

// run fn(gpu) on one host thread per GPU, each with its GPU's context current; returns the first non-zero status
template <class F> int on_every_gpu(F fn) {
    std::vector<int> rc((size_t)mg.n, 0);
    std::vector<std::string> err((size_t)mg.n);
    std::vector<std::thread> th;
    for (int i = 0; i < mg.n; ++i)
        th.emplace_back([&, i] {
            g_cur = &g_ctx[i];
            cudaSetDevice(g_ctx[i].device);
            rc[(size_t)i] = fn(i);
            if (rc[(size_t)i]) err[(size_t)i] = g_err;
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < mg.n; ++i)
        if (rc[(size_t)i]) return set_err(rc[(size_t)i], "GPU " + std::to_string(i) + ": " + err[(size_t)i]);
    return 0;
}
}  // namespace
extern "C" {

int rays_b200_init_multi(int ngpu) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(RAYS_ERR_CUDA, std::string("rays_b200_init_multi: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    int n = ngpu <= 0 ? ndev : ngpu;
    if (n > ndev) return set_err(RAYS_ERR_CUDA, "rays_b200_init_multi: more GPUs requested than visible");
    if (n > kMaxGpus) n = kMaxGpus;
    if (mg.n == n && mg.comm[0]) return 0;
    Ctx *keep = g_cur;
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i) { g_cur = &g_ctx[i]; rc = rays_b200_init(i); }
    g_cur = keep;
    if (rc) return rc;
    mg.n = n;
    g_copy_threads_div = n;
    if (n > 1) {
        if ((rc = load_nccl())) return rc;
        int devs[kMaxGpus];
        for (int i = 0; i < n; ++i) devs[i] = i;
        NCK(nccl.CommInitAll(mg.comm, n, devs));
    }
    CK(cudaSetDevice(g_ctx[0].device));
    return 0;
}
int rays_b200_ngpu(void) { return mg.n > 0 ? mg.n : (g_ctx[0].inited ? 1 : 0); }

// trace_rays over every GPU of rays_b200_init_multi: GPU g integrates rays g, g + n, g + 2n, ... of the fan and writes their
// trajectories and summaries straight into the caller's arrays (rows g + i*n); optionally binning the deposition while tracing
static int trace_multi_impl(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, rays_deposition *dep) {
    if (mg.n < 1) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_init_multi has not been called");
    if (!cfg || !fan || !res) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace_multi: null argument");
    const int n = mg.n;
    const long long nray = fan->nray;
    if (res->nray < nray) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace_multi: result arrays smaller than the fan");
    const bool want_traj = res->ray_vec != nullptr || res->residual != nullptr;
    if (dep && want_traj) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace_multi_binned: bins while tracing, without trajectory arrays (ray_vec = residual = NULL)");
    if (dep && (dep->n_bins < 1 || !(dep->grid_max > dep->grid_min))) return set_err(RAYS_ERR_INVALID_CONFIG, "deposition: bad grid");
    // page-lock pageable trajectory arrays ONCE for all GPUs (each worker would otherwise try to register the same range)
    std::vector<void *> registered;
    struct Unreg { std::vector<void *> &v; ~Unreg() { for (void *q : v) cudaHostUnregister(q); } } unreg{registered};
    if (want_traj) {
        const char *env = getenv("RAYS_B200_REGISTER_HOST");
        double *arrs[2] = {res->ray_vec, res->residual};
        const size_t bytes[2] = {(size_t)res->nray * res->npoints_alloc * res->nv * 8, (size_t)res->nray * res->npoints_alloc * 8};
        for (int k = 0; k < 2; ++k) {
            if (!arrs[k]) continue;
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, arrs[k]) != cudaSuccess) { cudaGetLastError(); continue; }
            if (at.type == cudaMemoryTypeUnregistered && !(env && env[0] == '0') && bytes[k] <= (size_t(1) << 30)) {
                if (cudaHostRegister(arrs[k], bytes[k], cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) registered.push_back(arrs[k]);
                else cudaGetLastError();
            }
        }
    }
    double wsum = 0.0;
    if (fan->ray_pwr_wt) for (long long i = 0; i < nray; ++i) wsum += std::fabs(fan->ray_pwr_wt[i]);
    std::vector<rays_results> part((size_t)n, *res);
    std::vector<long long> steps((size_t)n, 0);
    std::vector<double> ms((size_t)n, 0.0);
    int rc = on_every_gpu([&](int g) -> int {
        const long long m = nray > g ? (nray - g + n - 1) / n : 0;
        std::vector<double> r((size_t)3 * std::max<long long>(m, 1)), k((size_t)3 * std::max<long long>(m, 1)), w((size_t)std::max<long long>(m, 1));
        for (long long i = 0; i < m; ++i) {
            const long long j = g + i * n;
            for (int q = 0; q < 3; ++q) { r[(size_t)3 * i + q] = fan->rvec0[3 * j + q]; k[(size_t)3 * i + q] = fan->rindex_vec0[3 * j + q]; }
            w[(size_t)i] = fan->ray_pwr_wt ? fan->ray_pwr_wt[j] : 0.0;
        }
        rays_fan f{m, r.data(), k.data(), w.data()};
        RowMap rm; rm.row0 = g; rm.stride = n;
        int e2;
        if (!dep) {
            if ((e2 = trace_host_impl(cfg, &f, &part[(size_t)g], rm))) return e2;
        } else {   // fused binning: summaries to the host, bins stay on the GPU for the reduce
            if ((e2 = rays_b200_set_config(cfg)) || (e2 = rays_b200_fan_upload(&f))) return e2;
            cx().fan_weight = wsum;                      // the unit of the bins is the whole fan's, the same on every GPU
            if ((e2 = rays_b200_trace_device_binned(dep->n_bins, dep->grid_min, dep->grid_max, 0))) return e2;
            std::vector<int> codes((size_t)m);
            if (m) CK(cudaMemcpyAsync(codes.data(), cx().stop.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
            if ((e2 = copy_small_results(&part[(size_t)g], 0, m, rm))) return e2;
            CK(cudaStreamSynchronize(cx().stream));
            fill_flags(&part[(size_t)g], codes, 0, rm);
            part[(size_t)g].total_ray_steps = cx().last_steps;
            part[(size_t)g].total_trace_time = cx().last_ms * 1e-3;
            CK(mg.acc[g].reserve((size_t)dep->n_bins));
            CK(cudaMemcpyAsync(mg.acc[g].p, cx().dep.p, (size_t)dep->n_bins * sizeof(long long), cudaMemcpyDeviceToDevice, cx().stream));
            // packed summaries for the all-gather (rows padded to the largest shard)
            const long long mmax = (nray + n - 1) / n;
            const int rd = 6 + 2 * cx().res_nv;
            CK(mg.summ[g].reserve((size_t)std::max<long long>(mmax, 1) * rd));
            CK(cudaMemsetAsync(mg.summ[g].p, 0, (size_t)std::max<long long>(mmax, 1) * rd * sizeof(double), cx().stream));
            if ((e2 = rays_b200_summaries_pack(mg.summ[g].p, mmax, nullptr, nullptr))) return e2;
            if (n > 1) CK(mg.gathered[g].reserve((size_t)n * std::max<long long>(mmax, 1) * rd));
            CK(cudaStreamSynchronize(cx().stream));
        }
        steps[(size_t)g] = part[(size_t)g].total_ray_steps;
        ms[(size_t)g] = part[(size_t)g].total_trace_time;
        return 0;
    });
    if (rc) return rc;
    res->total_ray_steps = 0;
    res->total_trace_time = 0.0;
    for (int g = 0; g < n; ++g) { res->total_ray_steps += steps[(size_t)g]; res->total_trace_time = std::max(res->total_trace_time, ms[(size_t)g]); }
    if (dep) {
        const int nb = dep->n_bins;
        const long long mmax = std::max<long long>((nray + n - 1) / n, 1);
        const int rd = 6 + 2 * g_ctx[0].res_nv;
        if (n > 1) {   // the two collectives of the run, one NCCL group over the GPUs of this process
            NCK(nccl.GroupStart());
            for (int g = 0; g < n; ++g) NCK(nccl.Reduce(mg.acc[g].p, mg.acc[g].p, (size_t)nb, kNcclInt64, kNcclSum, 0, mg.comm[g], g_ctx[g].stream));
            NCK(nccl.GroupEnd());
            NCK(nccl.GroupStart());
            for (int g = 0; g < n; ++g) NCK(nccl.AllGather(mg.summ[g].p, mg.gathered[g].p, (size_t)mmax * rd, kNcclFloat64, mg.comm[g], g_ctx[g].stream));
            NCK(nccl.GroupEnd());
            for (int g = 0; g < n; ++g) { CK(cudaSetDevice(g_ctx[g].device)); CK(cudaStreamSynchronize(g_ctx[g].stream)); }
        }
        mg.gathered_rows_per_gpu = mmax; mg.gathered_row_doubles = rd;
        CK(cudaSetDevice(g_ctx[0].device));
        std::vector<long long> h((size_t)nb);
        CK(cudaMemcpy(h.data(), mg.acc[0].p, (size_t)nb * sizeof(long long), cudaMemcpyDeviceToHost));
        const double u = 1.0 / g_ctx[0].dep_scale;
        long long q = 0;
        for (int b = 0; b < nb; ++b) { if (dep->profile) dep->profile[b] = (double)h[(size_t)b] * u; q += h[(size_t)b]; }
        dep->Q_sum = (double)q * u;
    }
    CK(cudaSetDevice(g_ctx[0].device));
    return 0;
}
int rays_b200_trace_multi(const rays_cfg *cfg, const rays_fan *fan, rays_results *res) { return trace_multi_impl(cfg, fan, res, nullptr); }
int rays_b200_trace_multi_binned(const rays_cfg *cfg, const rays_fan *fan, rays_results *res, rays_deposition *dep) {
    if (!dep) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_trace_multi_binned: null deposition");
    return trace_multi_impl(cfg, fan, res, dep);
}
// the all-gathered summaries on GPU `gpu` after rays_b200_trace_multi_binned: ngpu blocks of *rows_per_gpu rows (block g = the rays
// g, g + ngpu, ... in order; rows past a shard's length are zero), *row_doubles doubles each (layout: rays_b200_summaries_pack)
const double *rays_b200_multi_summaries(int gpu, int64_t *rows_per_gpu, int32_t *row_doubles) {
    if (gpu < 0 || gpu >= mg.n) return nullptr;
    if (rows_per_gpu) *rows_per_gpu = mg.gathered_rows_per_gpu;
    if (row_doubles) *row_doubles = mg.gathered_row_doubles;
    return mg.n > 1 ? mg.gathered[gpu].p : mg.summ[gpu].p;
}
int rays_b200_finalize_multi(void) {
    Ctx *keep = g_cur;
    for (int i = 0; i < mg.n; ++i) {
        g_cur = &g_ctx[i];
        if (g_ctx[i].inited) cudaSetDevice(g_ctx[i].device);
        mg.acc[i].release(); mg.summ[i].release(); mg.gathered[i].release();
        if (mg.comm[i] && nccl.CommDestroy) nccl.CommDestroy(mg.comm[i]);
        mg.comm[i] = nullptr;
        rays_b200_finalize();
    }
    g_cur = keep;
    mg.n = 0;
    g_copy_threads_div = 1;
    return 0;
}

// ======================= probes ==============================================================================
static int run_probe(int which, int64_t n, const double *in, size_t in_per, double *out, size_t out_per, int32_t *code) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    if (n <= 0) return 0;
    CK(cudaSetDevice(cx().device));
    const TuOps *ops = tu_ops(cx().dc.c.equilib_model, RAYS_ODE_RK4);
    DevBuf<double> din, dout;
    DevBuf<int> dcode;
    CK(din.reserve((size_t)n * in_per)); CK(dout.reserve((size_t)n * out_per)); CK(dcode.reserve((size_t)n));
    CK(cudaMemcpyAsync(din.p, in, (size_t)n * in_per * sizeof(double), cudaMemcpyHostToDevice, cx().stream));
    cudaError_t e;
    if (which == 0) e = ops->probe_equilibrium(cx().sel, n, din.p, dout.p, dcode.p, cx().stream);
    else if (which == 1) e = ops->probe_rhs(cx().sel, n, din.p, dout.p, dcode.p, cx().stream);
    else e = ops->probe_check_save(cx().sel, n, din.p, dout.p, dcode.p, cx().stream);
    CK(e);
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * out_per * sizeof(double), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaMemcpyAsync(code, dcode.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    din.release(); dout.release(); dcode.release();
    return 0;
}
int rays_b200_probe_equilibrium(int64_t n, const double *rvec, double *out, int32_t *err) { return run_probe(0, n, rvec, 3, out, RAYS_EQ_OUT, err); }
int rays_b200_probe_rhs(int64_t n, const double *v, double *dvds, int32_t *stop) {
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    return run_probe(1, n, v, (size_t)cx().dc.c.nv, dvds, (size_t)cx().dc.c.nv, stop);
}
int rays_b200_probe_check_save(int64_t n, const double *v, double *resid, int32_t *stop) {
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    return run_probe(2, n, v, (size_t)cx().dc.c.nv, resid, 1, stop);
}

// ======================= O-X conversion analysis (row f4) ======================================================
int rays_b200_ox_conv_analysis(rays_ox_conv *out, int64_t *n_converted) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!cx().cfg_set) return set_err(RAYS_ERR_NOT_INITIALIZED, "rays_b200_set_config has not been called");
    if (!out) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_ox_conv_analysis: null output");
    if (!cx().have_traj || cx().res_nray <= 0) return set_err(RAYS_ERR_INVALID_CONFIG, "analyze_OX_conv needs the trajectories of a stored device trace (trace_device with store)");
    CK(cudaSetDevice(cx().device));
    const long long n = cx().res_nray;
    rays_ox_conv *d_out = nullptr;
    CK(cudaMalloc(&d_out, sizeof(rays_ox_conv) * (size_t)n));
    const unsigned grid = (unsigned)((n + 63) / 64);
    switch (cx().dc.c.equilib_model) {
        case RAYS_EQ_SLAB: ox_conv_kernel<RAYS_EQ_SLAB><<<grid, 64, 0, cx().stream>>>(n, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, d_out); break;
        case RAYS_EQ_SOLOVEV: ox_conv_kernel<RAYS_EQ_SOLOVEV><<<grid, 64, 0, cx().stream>>>(n, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, d_out); break;
        case RAYS_EQ_AXISYM_TOROID: ox_conv_kernel<RAYS_EQ_AXISYM_TOROID><<<grid, 64, 0, cx().stream>>>(n, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, d_out); break;
        default: ox_conv_kernel<RAYS_EQ_MULTIPLE_MIRROR><<<grid, 64, 0, cx().stream>>>(n, cx().res_nv, cx().res_npa, cx().ray_vec.p, cx().npoints.p, d_out); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(rays_ox_conv) * (size_t)n, cudaMemcpyDeviceToHost, cx().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cx().stream);
    cudaFree(d_out);
    CK(e);
    int64_t nc = 0;
    for (long long i = 0; i < n; ++i) nc += out[i].converted ? 1 : 0;
    if (n_converted) *n_converted = nc;
    return 0;
}

// ======================= mirror coil fields (row f4) ===========================================================
int rays_b200_mirror_brz_grid(const rays_coil *coils, int32_t n_coils, int32_t n_r, double r_min, double r_max, int32_t n_z, double z_min,
                              double z_max, double *r_grid, double *z_grid, double *Br, double *Bz, double *Aphi) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    if (!coils || n_coils < 1 || n_r < 1 || n_z < 1 || !Br || !Bz || !Aphi) return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_mirror_brz_grid: bad arguments");
    for (int i = 0; i < n_coils; ++i)
        if (coils[i].n_r_layers < 1 || coils[i].n_z_slices < 1 || !(coils[i].inner_radius > 0.0) || !(coils[i].outer_radius >= coils[i].inner_radius))
            return set_err(RAYS_ERR_INVALID_CONFIG, "rays_b200_mirror_brz_grid: coil needs n_r_layers, n_z_slices >= 1 and 0 < inner_radius <= outer_radius");
    CK(cudaSetDevice(cx().device));
    // grids exactly as calculate_B_on_rz_grid forms them (mirror_magnetics_m.f90:340-356)
    std::vector<double> rg((size_t)n_r), zg((size_t)n_z);
    if (n_r == 1) rg[0] = r_min; else for (int i = 1; i <= n_r; ++i) rg[(size_t)i - 1] = r_min + (r_max - r_min) * (i - 1) / (n_r - 1);
    if (n_z == 1) zg[0] = z_min; else for (int j = 1; j <= n_z; ++j) zg[(size_t)j - 1] = z_min + (z_max - z_min) * (j - 1) / (n_z - 1);
    BLoopConst K;   // B_loop_m.f90:28-33: pi is a default-real literal widened to double
    K.pi = (double)3.1415926535897932385f;
    K.mu0 = K.pi * (double)4.e-7f;
    K.c0 = K.mu0 / (2.0 * K.pi);
    const size_t n = (size_t)n_r * n_z;
    DevBuf<double> d_r, d_z, d_out;
    rays_coil *d_coils = nullptr;
    CK(d_r.reserve((size_t)n_r)); CK(d_z.reserve((size_t)n_z)); CK(d_out.reserve(3 * n));
    CK(cudaMalloc(&d_coils, sizeof(rays_coil) * (size_t)n_coils));
    cudaError_t e = cudaMemcpyAsync(d_coils, coils, sizeof(rays_coil) * (size_t)n_coils, cudaMemcpyHostToDevice, cx().stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_r.p, rg.data(), sizeof(double) * (size_t)n_r, cudaMemcpyHostToDevice, cx().stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_z.p, zg.data(), sizeof(double) * (size_t)n_z, cudaMemcpyHostToDevice, cx().stream);
    if (e == cudaSuccess) {
        mirror_Brz_grid_kernel<<<(unsigned)((n + 63) / 64), 64, 0, cx().stream>>>(d_coils, n_coils, K, n_r, n_z, d_r.p, d_z.p, d_out.p, d_out.p + n, d_out.p + 2 * n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(Br, d_out.p, sizeof(double) * n, cudaMemcpyDeviceToHost, cx().stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Bz, d_out.p + n, sizeof(double) * n, cudaMemcpyDeviceToHost, cx().stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Aphi, d_out.p + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, cx().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cx().stream);
    cudaFree(d_coils);
    d_r.release(); d_z.release(); d_out.release();
    CK(e);
    if (r_grid) std::copy(rg.begin(), rg.end(), r_grid);
    if (z_grid) std::copy(zg.begin(), zg.end(), z_grid);
    return 0;
}

// ======================= measurement helpers =================================================================
int rays_b200_fp64_peak(double *tflops, double *sm_mhz) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    CK(cudaSetDevice(cx().device));
    const int threads = 256, blocks = cx().num_sms * 8, iters = 1 << 16;
    DevBuf<double> out;
    CK(out.reserve((size_t)threads * blocks));
    fp64_peak_kernel<<<blocks, threads, 0, cx().stream>>>(out.p, 1024, 1.0000001, 1e-9);   // warm-up
    CK(cudaGetLastError());
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(cx().ev0, cx().stream));
        fp64_peak_kernel<<<blocks, threads, 0, cx().stream>>>(out.p, iters, 1.0000001, 1e-9);
        CK(cudaGetLastError());
        CK(cudaEventRecord(cx().ev1, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, cx().ev0, cx().ev1));
        const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    out.release();
    if (tflops) *tflops = best;
    if (sm_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, cx().device);
        *sm_mhz = khz / 1000.0;   // maximum SM clock; bench.py samples the clock under load with nvidia-smi
    }
    return 0;
}

// rcp_rn / sqrt_rn / qdiv of the kernels against IEEE 1.0/d, sqrt(x), x/d on n pseudo-random operand pairs
// (a quarter of them adversarial); mismatch[0..2] = counts for reciprocal, square root, quotient (must be 0);
// mismatch[3] = reciprocals of all-ones-significand divisors that are 1 ulp off (the documented exception)
int rays_b200_selftest_arith(int64_t n, uint64_t seed, int64_t *mismatch) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    CK(cudaSetDevice(cx().device));
    DevBuf<unsigned long long> d;
    CK(d.reserve(4));
    CK(cudaMemsetAsync(d.p, 0, 4 * sizeof(unsigned long long), cx().stream));
    const int threads = 256, blocks = cx().num_sms * 4;
    const int per_thread = (int)std::max<int64_t>(1, n / ((int64_t)threads * blocks));
    selftest_arith_kernel<<<blocks, threads, 0, cx().stream>>>(seed, per_thread, d.p);
    CK(cudaGetLastError());
    unsigned long long h[4];
    CK(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
    d.release();
    for (int i = 0; i < 4; ++i) mismatch[i] = (int64_t)h[i];
    return 0;
}

// pinned host memory for result arrays (so the trajectory copy-out runs at full PCIe rate)
int rays_b200_host_alloc(void **p, size_t bytes) {
    if (need_init()) return RAYS_ERR_NOT_INITIALIZED;
    CK(cudaHostAlloc(p, bytes, cudaHostAllocPortable | cudaHostAllocMapped));   // usable from every GPU of rays_b200_init_multi
    return 0;
}
int rays_b200_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}

}  // extern "C"
