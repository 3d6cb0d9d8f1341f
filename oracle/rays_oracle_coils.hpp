// rays_oracle_coils.hpp — CPU ORACLE (test infrastructure only, see rays_oracle.hpp) for the mirror coil-field
// generator: the field of circular current loops summed over the filaments of a coil set on an (r,z) grid.
// Restates, statement by statement:
//   MM/B_loop_m.f90:191-246            Brz_loop_scaled (loop of unit radius, 1 A)
//   M/complete_elliptic_int_m.f90:38-148, 150-318, 320-507   elliptic_Em / elliptic_Km via Carlson's RF, RD (errtol 1e-3)
//   MM/mirror_magnetics_m.f90:222-239  filament positions
//   MM/mirror_magnetics_m.f90:246-320  coil_Brz_field_1Amp, mirror_Brz_field
//   MM/mirror_magnetics_m.f90:324-368  calculate_B_on_rz_grid
// (MM/ = RAYS_project/mirror_magnetics_lib/, M/ = RAYS_project/math_functions_lib/).
// PARITY STATUS: unpinned by reference output (the shipped Brz file's coil set is not in the tree); checked against
// the textbook loop formulas with scipy's K(m), E(m) and against B = curl A (tests/test_coils.py).
#pragma once
#include <cfloat>
#include <cmath>
#include <vector>

#include "../include/rays_b200.h"

namespace rays_oracle {

// rf (M/complete_elliptic_int_m.f90:150-318)
inline double carlson_rf(double x, double y, double z, double errtol, int &ierr) {
    const double lolim = 3.e-78, uplim = 1.e+75;
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < lolim || x + z < lolim || y + z < lolim || uplim <= x || uplim <= y || uplim <= z) {
        ierr = 1;
        return 0.0;
    }
    ierr = 0;
    double xn = x, yn = y, zn = z;
    for (;;) {
        const double mu = (xn + yn + zn) / 3.0;
        const double xndev = 2.0 - (mu + xn) / mu;
        const double yndev = 2.0 - (mu + yn) / mu;
        const double zndev = 2.0 - (mu + zn) / mu;
        const double epslon = std::fmax(std::fmax(std::fabs(xndev), std::fabs(yndev)), std::fabs(zndev));
        if (epslon < errtol) {
            const double c1 = 1.0 / 24.0, c2 = 3.0 / 44.0, c3 = 1.0 / 14.0;
            const double e2 = xndev * yndev - zndev * zndev;
            const double e3 = xndev * yndev * zndev;
            const double s = 1.0 + (c1 * e2 - 0.1 - c2 * e3) * e2 + c3 * e3;
            return s / std::sqrt(mu);
        }
        const double xnroot = std::sqrt(xn), ynroot = std::sqrt(yn), znroot = std::sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
    }
}
// rd (M/complete_elliptic_int_m.f90:320-507)
inline double carlson_rd(double x, double y, double z, double errtol, int &ierr) {
    const double lolim = 3.e-78, uplim = 1.e+75;   // the module-level parameters shadow the routine's own (:29-30)
    if (x < 0.0 || y < 0.0 || x + y < lolim || z < lolim || uplim < x || uplim < y || uplim < z) {
        ierr = 1;
        return 0.0;
    }
    ierr = 0;
    double xn = x, yn = y, zn = z, sigma = 0.0, power4 = 1.0;
    for (;;) {
        const double mu = (xn + yn + 3.0 * zn) * 0.2;
        const double xndev = (mu - xn) / mu;
        const double yndev = (mu - yn) / mu;
        const double zndev = (mu - zn) / mu;
        const double epslon = std::fmax(std::fmax(std::fabs(xndev), std::fabs(yndev)), std::fabs(zndev));
        if (epslon < errtol) {
            const double c1 = 3.0 / 14.0, c2 = 1.0 / 6.0, c3 = 9.0 / 22.0, c4 = 3.0 / 26.0;
            const double ea = xndev * yndev;
            const double eb = zndev * zndev;
            const double ec = ea - eb;
            const double ed = ea - 6.0 * eb;
            const double ef = ed + ec + ec;
            const double s1 = ed * (-c1 + 0.25 * c3 * ed - 1.5 * c4 * zndev * ef);
            const double s2 = zndev * (c2 * ef + zndev * (-c3 * ec + zndev * c4 * ea));
            return 3.0 * sigma + power4 * (1.0 + s1 + s2) / (mu * std::sqrt(mu));
        }
        const double xnroot = std::sqrt(xn), ynroot = std::sqrt(yn), znroot = std::sqrt(zn);
        const double lamda = xnroot * (ynroot + znroot) + ynroot * znroot;
        sigma = sigma + power4 / (znroot * (zn + lamda));
        power4 = power4 * 0.25;
        xn = (xn + lamda) * 0.25;
        yn = (yn + lamda) * 0.25;
        zn = (zn + lamda) * 0.25;
    }
}
// elliptic_Em, elliptic_Km (M/complete_elliptic_int_m.f90:38-148): parameter m = k^2
inline double elliptic_Em(double m) {
    int ierr;
    const double x = 0.0, y = 1.0 - m, z = 1.0, errtol = 1.0e-3;
    return carlson_rf(x, y, z, errtol, ierr) - m * carlson_rd(x, y, z, errtol, ierr) / 3.0;
}
inline double elliptic_Km(double m) {
    int ierr;
    return carlson_rf(0.0, 1.0 - m, 1.0, 1.0e-3, ierr);
}

// module constants of B_loop_m (MM/B_loop_m.f90:28-33): pi is a default-real literal
struct BLoopConst {
    double pi, mu0, c0;
    BLoopConst() {
        pi = (double)3.1415926535897932385f;
        mu0 = pi * (double)4.e-7f;
        c0 = mu0 / (2.0 * pi);
    }
};

// Brz_loop_scaled (MM/B_loop_m.f90:191-246): r, z in units of the loop radius
inline void Brz_loop_scaled(double r, double z, double &Br, double &Bz, double &Aphi) {
    static const BLoopConst K;
    const double r0 = 0.001;
    const double r2 = r * r, z2 = z * z;
    if (r < 2.0 * DBL_MIN) {   // on axis
        Br = 0.0;
        Bz = K.mu0 / (2.0 * std::pow(1.0 + z2, 1.5));
        Aphi = 0.0;
        return;
    }
    if (r < r0) {   // near axis
        const double r3 = r * r2, r4 = r2 * r2, z4 = z2 * z2, f = 1.0 + z2;
        Br = 3.0 * z * r / (4.0 * std::pow(f, 2.5));
        Br = Br - 15.0 * z * r3 * (-3.0 + 4.0 * z2) / (32.0 * std::pow(f, 4.5));
        Br = K.mu0 * Br;
        Bz = 1.0 / 2.0 / std::pow(f, 1.5) + 3.0 / 8.0 * (1.0 - 4.0 * z2) * r2 / std::pow(f, 3.5) +
             45.0 / 128.0 * (1.0 - 12.0 * z2 + 8.0 * z4) * r4 / std::pow(f, 5.5);
        Bz = K.mu0 * Bz;
        Aphi = r2 / (4.0 * K.pi * std::pow(f, 1.5)) + 3.0 * (1.0 - 4.0 * z2) * r4 / (32.0 * K.pi * std::pow(f, 3.5));
        return;
    }
    const double m0 = (1.0 + r) * (1.0 + r) + z2;
    const double m = 4.0 * r / m0;
    const double alpha = 1.0 + r2 + z2;
    const double beta = 1.0 - r2 - z2;
    const double gamma = (1.0 - r) * (1.0 - r) + z2;
    const double Em = elliptic_Em(m);
    const double Km = elliptic_Km(m);
    Br = K.c0 * z / (r * std::sqrt(m0)) * (alpha / gamma * Em - Km);
    Bz = K.c0 / std::sqrt(m0) * (beta / gamma * Em + Km);
    Aphi = -std::sqrt(m0) * Em + alpha / std::sqrt(m0) * Km;
    Aphi = K.c0 * Aphi;
}

// coil_Brz_field_1Amp + mirror_Brz_field (MM/mirror_magnetics_m.f90:246-320) at one point
inline void mirror_Brz_field(const rays_coil *coils, int n_coils, double r, double z, double &Br, double &Bz, double &Aphi) {
    Br = 0.0; Bz = 0.0; Aphi = 0.0;
    for (int ic = 0; ic < n_coils; ++ic) {
        const rays_coil &c = coils[ic];
        // filament positions (:222-239)
        const double delta_r = (c.outer_radius - c.inner_radius) / (c.n_r_layers + 1);
        const double delta_z = c.z_width / (c.n_z_slices + 1);
        double Br_i = 0.0, Bz_i = 0.0, Aphi_i = 0.0;
        for (int i = 1; i <= c.n_r_layers; ++i) {
            const double a = c.inner_radius + i * delta_r;
            for (int j = 1; j <= c.n_z_slices; ++j) {
                const double z_filament = c.z_center - c.z_width / 2.0 + j * delta_z;
                const double z_relative = z - z_filament;
                double Br_f, Bz_f, Aphi_f;
                Brz_loop_scaled(r / a, z_relative / a, Br_f, Bz_f, Aphi_f);
                Br_i = Br_i + Br_f / a;
                Bz_i = Bz_i + Bz_f / a;
                Aphi_i = Aphi_i + Aphi_f * a;
            }
        }
        const int nf = c.n_r_layers * c.n_z_slices;
        Br_i = Br_i / nf; Bz_i = Bz_i / nf; Aphi_i = Aphi_i / nf;
        Br = Br + Br_i * (double)c.n_turns * c.I_coil;
        Bz = Bz + Bz_i * (double)c.n_turns * c.I_coil;
        Aphi = Aphi + Aphi_i * (double)c.n_turns * c.I_coil;
    }
}

// calculate_B_on_rz_grid (MM/mirror_magnetics_m.f90:324-368); output in the file's order [j_z][i_r] (= Fortran (n_r, n_z))
inline void mirror_Brz_grid(const rays_coil *coils, int n_coils, int n_r, double r_min, double r_max, int n_z, double z_min, double z_max,
                            double *r_grid, double *z_grid, double *Br, double *Bz, double *Aphi) {
    if (n_r == 1) r_grid[0] = r_min;
    else for (int i = 1; i <= n_r; ++i) r_grid[i - 1] = r_min + (r_max - r_min) * (i - 1) / (n_r - 1);
    if (n_z == 1) z_grid[0] = z_min;
    else for (int j = 1; j <= n_z; ++j) z_grid[j - 1] = z_min + (z_max - z_min) * (j - 1) / (n_z - 1);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n_z; ++j)
        for (int i = 0; i < n_r; ++i)
            mirror_Brz_field(coils, n_coils, r_grid[i], z_grid[j], Br[(size_t)j * n_r + i], Bz[(size_t)j * n_r + i], Aphi[(size_t)j * n_r + i]);
}

}  // namespace rays_oracle
