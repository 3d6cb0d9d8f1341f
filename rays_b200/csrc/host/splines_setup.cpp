// splines_setup.cpp — spline COEFFICIENT SETUP on the host (once per run), producing the tables
// the device evaluates.  Mirrors the reference's splines_lib (S/ = RAYS_project/splines_lib/):
//   v_spline   S/v_spline.f90:1-637   (general n > 3 tridiagonal path, BC types 0 (not-a-knot), 1, 2)
//   cspline    S/cspline.f90:7-169
//   bcspline   S/bcspline.f90:23-434  (homogeneous/not-a-knot BCs as used by
//                                      quick_cube_splines_m.f90:180-250: ibc* = 0)
//   splinck    S/splinck.f90
// The reference only ever calls these with not-a-knot conditions on all edges
// (S/quick_cube_splines_m.f90:131-139,214-221, M/zfunctions_m.f90:449-458).
#include "splines_setup.hpp"

#include <algorithm>
#include <cmath>
#include <vector>

namespace rays_host {

// f is Fortran f(4,n): f[4*i + c]; x is x(n).  wk unused for non-periodic BCs.
// Follows v_spline's "ELSE IF(n.gt.2)" branch (S/v_spline.f90:398-560) for i_bc in {0,1,2}.
int v_spline(int k_bc1, int k_bcn, int n, const double *x, double *f) {
    if (n < 4) return 1;  // the small-n special cases are never reached by the reference's callers
    int i_bc1 = k_bc1, i_bcn = k_bcn;
    if (i_bc1 < 0 || i_bc1 > 2) i_bc1 = 0;
    if (i_bcn < 0 || i_bcn > 2) i_bcn = 0;
#define F(c, i) f[4 * ((i)-1) + ((c)-1)]
#define X(i) x[(i)-1]
    int imin = 1, imax = n;
    double a1 = 0.0, b1 = 0.0, an = 0.0, bn = 0.0;
    if (i_bc1 == 1) a1 = F(2, 1);
    else if (i_bc1 == 2) b1 = F(3, 1);
    if (i_bcn == 1) an = F(2, n);
    else if (i_bcn == 2) bn = F(3, n);
    F(2, n) = 0.0; F(3, n) = 0.0; F(4, n) = 0.0;
    // set up the tridiagonal system: F(4,*) = h, F(2,*) = diagonal, F(3,*) = rhs
    F(4, 1) = X(2) - X(1);
    F(3, 2) = (F(1, 2) - F(1, 1)) / F(4, 1);
    for (int i = 2; i <= n - 1; ++i) {
        F(4, i) = X(i + 1) - X(i);
        F(2, i) = 2.0 * (F(4, i - 1) + F(4, i));
        F(3, i + 1) = (F(1, i + 1) - F(1, i)) / F(4, i);
        F(3, i) = F(3, i + 1) - F(3, i);
    }
    double elem21 = F(4, 1);
    double elemnn1 = F(4, n - 1);
    if (i_bc1 == 1) {
        F(2, 1) = 2.0 * F(4, 1);
        F(3, 1) = (F(1, 2) - F(1, 1)) / F(4, 1) - a1;
    } else if (i_bc1 == 2) {
        F(2, 1) = 2.0 * F(4, 1);
        F(3, 1) = F(4, 1) * b1 / 3.0;
        F(4, 1) = 0.0;
    } else {  // not a knot
        imin = 2;
        F(2, 2) = F(4, 1) + 2.0 * F(4, 2);
        F(3, 2) = F(3, 2) * F(4, 2) / (F(4, 1) + F(4, 2));
    }
    if (i_bcn == 1) {
        F(2, n) = 2.0 * F(4, n - 1);
        F(3, n) = -(F(1, n) - F(1, n - 1)) / F(4, n - 1) + an;
    } else if (i_bcn == 2) {
        F(2, n) = 2.0 * F(4, n - 1);
        F(3, n) = F(4, n - 1) * bn / 3.0;
        elemnn1 = 0.0;
    } else {  // not a knot
        imax = n - 1;
        F(2, n - 1) = 2.0 * F(4, n - 2) + F(4, n - 1);
        F(3, n - 1) = F(3, n - 1) * F(4, n - 2) / (F(4, n - 1) + F(4, n - 2));
    }
    // forward elimination
    for (int i = imin + 1; i <= imax; ++i) {
        double t;
        if (i == n - 1 && imax == n - 1) t = (F(4, i - 1) - F(4, i)) / F(2, i - 1);
        else if (i == 2) t = elem21 / F(2, i - 1);
        else if (i == n) t = elemnn1 / F(2, i - 1);
        else t = F(4, i - 1) / F(2, i - 1);
        if (i == imin + 1 && imin == 2) F(2, i) = F(2, i) - t * (F(4, i - 1) - F(4, i - 2));
        else F(2, i) = F(2, i) - t * F(4, i - 1);
        F(3, i) = F(3, i) - t * F(3, i - 1);
    }
    // back substitution
    F(3, imax) = F(3, imax) / F(2, imax);
    for (int ib = 1; ib <= imax - imin; ++ib) {
        int i = imax - ib;
        if (i == 2 && imin == 2) F(3, i) = (F(3, i) - (F(4, i) - F(4, i - 1)) * F(3, i + 1)) / F(2, i);
        else F(3, i) = (F(3, i) - F(4, i) * F(3, i + 1)) / F(2, i);
    }
    F(4, 1) = X(2) - X(1);
    F(4, n - 1) = X(n) - X(n - 1);
    if (i_bc1 <= 0) F(3, 1) = (F(3, 2) * (F(4, 1) + F(4, 2)) - F(3, 3) * F(4, 1)) / F(4, 2);
    if (i_bcn <= 0) F(3, n) = F(3, n - 1) + (F(3, n - 1) - F(3, n - 2)) * F(4, n - 1) / F(4, n - 2);
    for (int i = 1; i <= n - 1; ++i) {
        F(2, i) = (F(1, i + 1) - F(1, i)) / F(4, i) - F(4, i) * (F(3, i + 1) + 2.0 * F(3, i));
        F(4, i) = (F(3, i + 1) - F(3, i)) / F(4, i);
        F(3, i) = 6.0 * F(3, i);
        F(4, i) = 6.0 * F(4, i);
    }
    double hn = X(n) - X(n - 1);
    F(2, n) = F(2, n - 1) + hn * (F(3, n - 1) + 0.5 * hn * F(4, n - 1));
    F(3, n) = F(3, n - 1) + hn * F(4, n - 1);
    F(4, n) = F(4, n - 1);
    if (i_bcn == 1) F(2, n) = an;
    else if (i_bcn == 2) F(3, n) = bn;
#undef F
#undef X
    return 0;
}

// splinck (S/splinck.f90): ilinx = 1 if evenly spaced within ztol, else 2; ier = 2 if not ascending
int splinck(const double *x, int inx, double ztol, int *ier) {
    *ier = 0;
    int ilinx = 1;
    if (inx <= 1) return ilinx;
    double dxavg = (x[inx - 1] - x[0]) / (inx - 1);
    double zeps = std::fabs(ztol * dxavg);
    for (int ix = 1; ix < inx; ++ix) {
        double zdiffx = x[ix] - x[ix - 1];
        if (zdiffx <= 0.0) *ier = 2;
        if (std::fabs(zdiffx - dxavg) > zeps) ilinx = 2;
    }
    return ilinx;
}

// cspline (S/cspline.f90:7-169) with ibcxmin = ibcxmax = 0
int cspline(const double *x, int nx, double *fspl, int *ilinx) {
    if (nx < 2) return 1;
    int ierx;
    *ilinx = splinck(x, nx, 1.0e-3, &ierx);
    if (ierx != 0) return 2;
    int rc = v_spline(0, 0, nx, x, fspl);
    if (rc) return rc;
    const double half = 0.5, sixth = 0.166666666666666667;
    for (int i = 0; i < nx; ++i) {
        fspl[4 * i + 2] = half * fspl[4 * i + 2];
        fspl[4 * i + 3] = sixth * fspl[4 * i + 3];
    }
    return 0;
}

// bcspline (S/bcspline.f90:23-434) with all four BC flags = 0 (iflg2 = 0: no BC correction pass).
// fspl is Fortran fspl(4,4,nx,ny): fspl[((j*nx+i)*4+cy)*4+cx], fspl(1,1,i,j) = data on entry.
int bcspline(const double *x, int inx, const double *th, int inth, double *fspl, int *ilinx, int *ilinth) {
    if (inx < 2 || inth < 2) return 1;
    int ierx, ierth;
    *ilinx = splinck(x, inx, 1.0e-3, &ierx);
    if (ierx != 0) return 2;
    *ilinth = splinck(th, inth, 1.0e-3, &ierth);
    if (ierth != 0) return 3;
    const double xo2 = 0.5, xo6 = 1.0 / 6.0;
#define FS(cx, cy, ix, ith) fspl[((size_t)(((ith)-1) * inx + ((ix)-1)) * 4 + ((cy)-1)) * 4 + ((cx)-1)]
    std::vector<double> wk(4 * (size_t)std::max(inx, inth));
    // spline in x for each theta
    for (int ith = 1; ith <= inth; ++ith) {
        for (int ix = 1; ix <= inx; ++ix) { wk[4 * (ix - 1)] = FS(1, 1, ix, ith); wk[4 * (ix - 1) + 1] = 0; wk[4 * (ix - 1) + 2] = 0; wk[4 * (ix - 1) + 3] = 0; }
        int rc = v_spline(0, 0, inx, x, wk.data());
        if (rc) return 10 + rc;
        for (int ix = 1; ix <= inx; ++ix) {
            FS(2, 1, ix, ith) = wk[4 * (ix - 1) + 1];
            FS(3, 1, ix, ith) = wk[4 * (ix - 1) + 2] * xo2;
            FS(4, 1, ix, ith) = wk[4 * (ix - 1) + 3] * xo6;
        }
    }
    // spline each x-coefficient in theta
    for (int ix = 1; ix <= inx; ++ix)
        for (int ic = 1; ic <= 4; ++ic) {
            for (int ith = 1; ith <= inth; ++ith) { wk[4 * (ith - 1)] = FS(ic, 1, ix, ith); wk[4 * (ith - 1) + 1] = 0; wk[4 * (ith - 1) + 2] = 0; wk[4 * (ith - 1) + 3] = 0; }
            int rc = v_spline(0, 0, inth, th, wk.data());
            if (rc) return 20 + rc;
            for (int ith = 1; ith <= inth; ++ith) {
                FS(ic, 2, ix, ith) = wk[4 * (ith - 1) + 1];
                FS(ic, 3, ix, ith) = wk[4 * (ith - 1) + 2] * xo2;
                FS(ic, 4, ix, ith) = wk[4 * (ith - 1) + 3] * xo6;
            }
        }
#undef FS
    return 0;
}

}  // namespace rays_host
