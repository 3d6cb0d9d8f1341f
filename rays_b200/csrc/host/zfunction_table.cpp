// zfunction_table.cpp — host-side construction of the plasma-dispersion-function spline table.
// Mirrors RAYS_project/math_functions_lib/zfunctions_m.f90:
//   wzdisp_D :283-347 (Gautschi's algorithm for w(z)), zzdisp_D :231-255, zfun_D :205-214,
//   initialize_spline_coeffs :436-466 (2001 points on [-10,10], not-a-knot cubic spline).
// Single-precision literals in the Fortran (4.29, 5.33, 28.41, 1., .5 ...) are reproduced
// (SURVEY.md A.1): comparisons against them use the float32-rounded value.
#include <cmath>

#include "splines_setup.hpp"

namespace rays_host {

static inline double f32(double x) { return (double)(float)x; }

// __builtin_powi as libgcc's __powidf2 (Fortran h2**capn with integer capn)
static double powi(double x, int m) {
    unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y *= x;
    }
    return m < 0 ? 1.0 / y : y;
}

static void wzdisp_D(double x, double y, double *re, double *im) {
    const double eps = 1.e-12;  // epsh = epsl = epsy (equivalence)
    double h, h2 = 0.0, lambda = 0.0;
    int capn, nu;
    if (y < f32(4.29) && x < f32(5.33)) {
        double s = (1. - y / f32(4.29)) * std::sqrt(1. - x * x / f32(28.41));
        h = 1.6 * s;
        h2 = 2.0 * h;
        capn = (int)(6.0 + 23.0 * s + .5);
        nu = (int)(9.0 + 21.0 * s + .5);
        lambda = powi(h2, capn);
    } else {
        h = 0.0;
        capn = 0;
        nu = 8;
    }
    bool b = (h == 0.0 || lambda < eps);
    double rr = 0.0, ri = 0.0, sr = 0.0, si = 0.0;
    int nup = nu + 1;
    for (int i = 1; i <= nup; ++i) {
        int n = nup - i;
        int np1 = n + 1;
        double tr = y + h + np1 * rr;
        double ti = x - np1 * ri;
        double c = .5 / (tr * tr + ti * ti);
        rr = c * tr;
        ri = c * ti;
        if (!(h > 0.0 && n <= capn)) continue;
        tr = lambda + sr;
        sr = rr * tr - ri * si;
        si = ri * tr + rr * si;
        lambda = lambda / h2;
    }
    const double cc = 1.12837916709551;
    if (y < eps) *re = std::exp(-x * x);
    else if (b) *re = rr * cc;
    else *re = sr * cc;
    if (b) *im = ri * cc;
    else *im = si * cc;
}

// zfun_D -> zzdisp_D (zfunctions_m.f90:205-255): Z(x + i y)
void zfun_D(double x, double y, double *zzr, double *zzi) {
    const double pi = std::atan2(0.0, -1.0);  // atan(zero,-one) (:16)
    const double sqrt_pi = std::sqrt(pi);
    double x1 = std::fabs(x), y1 = std::fabs(y), wzr1, wzi1;
    wzdisp_D(x1, y1, &wzr1, &wzi1);
    if (x >= 0.0 && y >= 0.0) {
    } else if (x <= 0.0 && y >= 0.0) {
        wzi1 = -wzi1;
    } else {
        double a = 2.0 * x1 * y1;
        double b = -(x1 * x1 - y1 * y1);
        double abr = 2.0 * std::exp(b) * std::cos(a);
        double abi = -2.0 * std::exp(b) * std::sin(a);
        wzr1 = abr - wzr1;
        wzi1 = abi - wzi1;
        if (!(x <= 0.0 && y <= 0.0)) wzi1 = -wzi1;
    }
    *zzr = -sqrt_pi * wzi1;
    *zzi = sqrt_pi * wzr1;
}

int zfun_table(double *x_grid, double *fsplRe, double *fsplIm) {
    const int nx = ZFUN_NX;
    const double x_grid_min = -10.0, x_grid_max = 10.0;
    for (int i = 1; i <= nx; ++i) {
        x_grid[i - 1] = x_grid_min + (i - 1) * (x_grid_max - x_grid_min) / (nx - 1);
        double re, im;
        zfun_D(x_grid[i - 1], 0.0, &re, &im);
        for (int c = 0; c < 4; ++c) { fsplRe[4 * (i - 1) + c] = 0.0; if (fsplIm) fsplIm[4 * (i - 1) + c] = 0.0; }
        fsplRe[4 * (i - 1)] = re;
        if (fsplIm) fsplIm[4 * (i - 1)] = im;
    }
    int ilinx;
    int rc = cspline(x_grid, nx, fsplRe, &ilinx);
    if (rc) return rc;
    if (fsplIm) rc = cspline(x_grid, nx, fsplIm, &ilinx);
    return rc;
}

}  // namespace rays_host
