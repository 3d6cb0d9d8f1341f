// host check of the double-double integer power against long double / exact: same algorithm in C
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
typedef struct { double hi, lo; } DD;
static DD dd_mul(DD a, DD b) { double p = a.hi*b.hi; double e = fma(a.hi,b.hi,-p); e = fma(a.hi,b.lo,e); e = fma(a.lo,b.hi,e); double s = p+e; DD r = {s, e-(s-p)}; return r; }
static double powi(double x, int n) {
    double p2 = x*x; DD x2 = {p2, fma(x,x,-p2)};
    if (n <= 5) { DD xx = {x,0}; if (n==3) { DD r = dd_mul(x2,xx); return r.hi+r.lo; } DD x4 = dd_mul(x2,x2); if (n==4) return x4.hi+x4.lo; DD r = dd_mul(x4,xx); return r.hi+r.lo; }
    DD r = {1,0}, b = x2; int first = 1;
    if (n & 1) { r.hi = x; r.lo = 0; first = 0; }
    for (int bit = 1; bit < 5; ++bit) { if (n & (1<<bit)) { r = first ? b : dd_mul(r,b); first = 0; } if ((n >> (bit+1)) != 0) b = dd_mul(b,b); }
    return r.hi + r.lo;
}
int main() {
    long bad = 0, tot = 0, badlibm = 0;
    srand(7);
    for (int n = 3; n <= 16; ++n) for (int i = 0; i < 200000; ++i) {
        double x = (double)rand()/RAND_MAX; if (i & 1) x = x*3.0 - 1.5;
        if (fabs(x) < 1e-3) continue;
        __float128 xq = x, rq = 1; for (int k = 0; k < n; ++k) rq *= xq;   // 113-bit: exact enough for n*53 <= ... rounding check below
        double ref = (double)rq; double got = powi(x, n); double lm = pow(x, (double)n);
        ++tot; if (got != ref) ++bad; if (lm != ref) ++badlibm;
    }
    printf("dd power: %ld of %ld differ from the float128 result rounded to double; glibc pow: %ld differ\n", bad, tot, badlibm);
    return 0;
}
