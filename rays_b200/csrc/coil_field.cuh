// coil_field.cuh — field of a set of circular coils on an (r,z) grid: the generator of the Brz field file that
// mirror_magnetics_spline_interp reads.  Device restatement of
//   mirror_magnetics_lib/B_loop_m.f90:191-246 (Brz_loop_scaled), math_functions_lib/complete_elliptic_int_m.f90
//   (elliptic_Em/Km through Carlson's RF and RD, errtol 1e-3), mirror_magnetics_lib/mirror_magnetics_m.f90:222-368
//   (filament positions, coil_Brz_field_1Amp, mirror_Brz_field, calculate_B_on_rz_grid)
// in the reference's operation order with IEEE division and square root (-fmad=false), so every grid point off
// the axis is bitwise equal to the CPU oracle; on and next to the axis (r < 1e-3 loop radii) the reference's
// series uses real powers -> pow(), CUDA's differs from glibc's in the last bits.
// One thread per grid point; the work is tiny (n_r * n_z * filaments loop fields), latency-bound by design.
#pragma once
#include <cfloat>

#include "../../include/rays_b200.h"

namespace rays_dev {

struct BLoopConst { double pi, mu0, c0; };   // B_loop_m.f90:28-33, formed on the host (single-precision pi literal)

__device__ __forceinline__ double cf_rf(double x, double y, double z, double errtol) {
    const double lolim = 3.e-78, uplim = 1.e+75;
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < lolim || x + z < lolim || y + z < lolim || uplim <= x || uplim <= y || uplim <= z) return 0.0;
    double xn = x, yn = y, zn = z;
    for (;;) {
        const double mu = (xn + yn + zn) / 3.0;
        const double xndev = 2.0 - (mu + xn) / mu;
        const double yndev = 2.0 - (mu + yn) / mu;
        const double zndev = 2.0 - (mu + zn) / mu;
        const double epslon = fmax(fmax(fabs(xndev), fabs(yndev)), fabs(zndev));
        if (epslon < errtol) {
            const double c1 = 1.0 / 24.0, c2 = 3.0 / 44.0, c3 = 1.0 / 14.0;
            const double e2 = xndev * yndev - zndev * zndev;
            const double e3 = xndev * yndev * zndev;
            const double s = 1.0 + (c1 * e2 - 0.1 - c2 * e3) * e2 + c3 * e3;
            return s / sqrt(mu);
        }
        const double xnroot = sqrt(xn), ynroot = sqrt(yn), znroot = sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
    }
}
__device__ __forceinline__ double cf_rd(double x, double y, double z, double errtol) {
    const double lolim = 3.e-78, uplim = 1.e+75;
    if (x < 0.0 || y < 0.0 || x + y < lolim || z < lolim || uplim < x || uplim < y || uplim < z) return 0.0;
    double xn = x, yn = y, zn = z, sigma = 0.0, power4 = 1.0;
    for (;;) {
        const double mu = (xn + yn + 3.0 * zn) * 0.2;
        const double xndev = (mu - xn) / mu;
        const double yndev = (mu - yn) / mu;
        const double zndev = (mu - zn) / mu;
        const double epslon = fmax(fmax(fabs(xndev), fabs(yndev)), fabs(zndev));
        if (epslon < errtol) {
            const double c1 = 3.0 / 14.0, c2 = 1.0 / 6.0, c3 = 9.0 / 22.0, c4 = 3.0 / 26.0;
            const double ea = xndev * yndev;
            const double eb = zndev * zndev;
            const double ec = ea - eb;
            const double ed = ea - 6.0 * eb;
            const double ef = ed + ec + ec;
            const double s1 = ed * (-c1 + 0.25 * c3 * ed - 1.5 * c4 * zndev * ef);
            const double s2 = zndev * (c2 * ef + zndev * (-c3 * ec + zndev * c4 * ea));
            return 3.0 * sigma + power4 * (1.0 + s1 + s2) / (mu * sqrt(mu));
        }
        const double xnroot = sqrt(xn), ynroot = sqrt(yn), znroot = sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        sigma = sigma + power4 / (znroot * (zn + lamda));
        power4 = power4 * 0.25;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
    }
}

// Brz_loop_scaled (B_loop_m.f90:191-246)
__device__ __forceinline__ void cf_loop_scaled(const BLoopConst &K, double r, double z, double &Br, double &Bz, double &Aphi) {
    const double r0 = 0.001;
    const double r2 = r * r, z2 = z * z;
    if (r < 2.0 * DBL_MIN) {
        Br = 0.0;
        Bz = K.mu0 / (2.0 * pow(1.0 + z2, 1.5));
        Aphi = 0.0;
        return;
    }
    if (r < r0) {
        const double r3 = r * r2, r4 = r2 * r2, z4 = z2 * z2, f = 1.0 + z2;
        Br = 3.0 * z * r / (4.0 * pow(f, 2.5));
        Br = Br - 15.0 * z * r3 * (-3.0 + 4.0 * z2) / (32.0 * pow(f, 4.5));
        Br = K.mu0 * Br;
        Bz = 1.0 / 2.0 / pow(f, 1.5) + 3.0 / 8.0 * (1.0 - 4.0 * z2) * r2 / pow(f, 3.5) +
             45.0 / 128.0 * (1.0 - 12.0 * z2 + 8.0 * z4) * r4 / pow(f, 5.5);
        Bz = K.mu0 * Bz;
        Aphi = r2 / (4.0 * K.pi * pow(f, 1.5)) + 3.0 * (1.0 - 4.0 * z2) * r4 / (32.0 * K.pi * pow(f, 3.5));
        return;
    }
    const double m0 = (1.0 + r) * (1.0 + r) + z2;
    const double m = 4.0 * r / m0;
    const double alpha = 1.0 + r2 + z2;
    const double beta = 1.0 - r2 - z2;
    const double gamma = (1.0 - r) * (1.0 - r) + z2;
    const double y = 1.0 - m;
    const double rf = cf_rf(0.0, y, 1.0, 1.0e-3);          // elliptic_Km; elliptic_Em evaluates the same RF again
    const double Em = rf - m * cf_rd(0.0, y, 1.0, 1.0e-3) / 3.0;
    const double Km = rf;
    const double sm0 = sqrt(m0);
    Br = K.c0 * z / (r * sm0) * (alpha / gamma * Em - Km);
    Bz = K.c0 / sm0 * (beta / gamma * Em + Km);
    Aphi = -sm0 * Em + alpha / sm0 * Km;
    Aphi = K.c0 * Aphi;
}

// calculate_B_on_rz_grid: thread t -> grid point (i_r = t % n_r, j_z = t / n_r), outputs [j_z][i_r]
static __global__ void mirror_Brz_grid_kernel(const rays_coil *__restrict__ coils, int n_coils, BLoopConst K, int n_r, int n_z,
                                              const double *__restrict__ r_grid, const double *__restrict__ z_grid,
                                              double *__restrict__ Br, double *__restrict__ Bz, double *__restrict__ Aphi) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n_r * n_z) return;
    const double r = r_grid[t % n_r], z = z_grid[t / n_r];
    double br = 0.0, bz = 0.0, aphi = 0.0;
    for (int ic = 0; ic < n_coils; ++ic) {
        const rays_coil c = coils[ic];
        const double delta_r = (c.outer_radius - c.inner_radius) / (c.n_r_layers + 1);
        const double delta_z = c.z_width / (c.n_z_slices + 1);
        double br_i = 0.0, bz_i = 0.0, aphi_i = 0.0;
        for (int i = 1; i <= c.n_r_layers; ++i) {
            const double a = c.inner_radius + i * delta_r;
            for (int j = 1; j <= c.n_z_slices; ++j) {
                const double z_filament = c.z_center - c.z_width / 2.0 + j * delta_z;
                const double z_relative = z - z_filament;
                double f_r, f_z, f_a;
                cf_loop_scaled(K, r / a, z_relative / a, f_r, f_z, f_a);
                br_i = br_i + f_r / a;
                bz_i = bz_i + f_z / a;
                aphi_i = aphi_i + f_a * a;
            }
        }
        const int nf = c.n_r_layers * c.n_z_slices;
        br_i = br_i / nf; bz_i = bz_i / nf; aphi_i = aphi_i / nf;
        br = br + br_i * (double)c.n_turns * c.I_coil;
        bz = bz + bz_i * (double)c.n_turns * c.I_coil;
        aphi = aphi + aphi_i * (double)c.n_turns * c.I_coil;
    }
    Br[t] = br; Bz[t] = bz; Aphi[t] = aphi;
}

}  // namespace rays_dev
