// splines_setup.hpp — host-side spline coefficient setup (see splines_setup.cpp)
#pragma once
namespace rays_host {
int v_spline(int k_bc1, int k_bcn, int n, const double *x, double *f);
int splinck(const double *x, int inx, double ztol, int *ier);
int cspline(const double *x, int nx, double *fspl, int *ilinx);
int bcspline(const double *x, int inx, const double *th, int inth, double *fspl, int *ilinx, int *ilinth);
// plasma Z function table (math_functions_lib/zfunctions_m.f90:436-466): x_grid[2001], fsplRe[4*2001]
void zfun_D(double x, double y, double *re, double *im);
int zfun_table(double *x_grid, double *fsplRe, double *fsplIm);
constexpr int ZFUN_NX = 2001;
}  // namespace rays_host
