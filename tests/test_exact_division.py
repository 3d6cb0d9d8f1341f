"""The kernels replace IEEE division x/d by q = x*r; q' = fma(fma(-d, q, x), r, q) with r = RN(1/d)
(ray_physics.cuh: qdiv).  This test checks that identity against true division on 10^8 operand pairs,
half of them adversarial (divisor just below a power of two with a dividend just below the divisor, where
the first product is NOT a faithful rounding; all-ones and near-power-of-two mantissas)."""
import os
import numpy as np
import subprocess
import tempfile

SRC = r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s = 0x9E3779B97F4A7C15ULL;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double mk(uint64_t m, int e) { uint64_t b = ((uint64_t)(1023 + e) << 52) | (m & 0xFFFFFFFFFFFFFULL); double d; memcpy(&d, &b, 8); return d; }
int main(void) {
    long bad = 0, unfaithful = 0;
    for (long it = 0; it < 100000000L; ++it) {
        uint64_t ma = rnd(), mb = rnd();
        int mode = it & 7, ea = 0, eb = 0;
        if (mode == 0) { mb |= 0xFFF0000000000ULL; ma |= 0xFF00000000000ULL; }
        else if (mode == 1) { mb |= 0xFFFFFFF000000ULL; ma |= 0xFFFFFF0000000ULL; }
        else if (mode == 2) { mb = 0xFFFFFFFFFFFFFULL - (mb & 0xFFFFF); ma = 0xFFFFFFFFFFFFFULL - (ma & 0xFFFFFF); }
        else if (mode == 3) { mb &= 0xFFFULL; }
        else { ea = (int)(rnd() % 120) - 60; eb = (int)(rnd() % 120) - 60; }
        double a = mk(ma, ea), b = mk(mb, eb);
        if (rnd() & 1) a = -a;
        if (rnd() & 1) b = -b;
        double r = 1.0 / b;
        double q = a * r;
        double q2 = fma(fma(-b, q, a), r, q);
        double t = a / b;
        if (q2 != t) bad++;
        if (q != t && q != nextafter(t, fma(-b, t, a) * b > 0 ? INFINITY : -INFINITY)) unfaithful++;
    }
    printf("%ld %ld\n", bad, unfaithful);
    return 0;
}
"""


def test_reciprocal_fma_division_is_exact():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(SRC)
        subprocess.run(["gcc", "-O2", "-mfma", "-o", exe, src, "-lm"], check=True)
        bad, unfaithful = map(int, subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split())
    assert bad == 0
    assert unfaithful > 1000000   # the adversarial cases really do hit the hard region


def test_tanh_cosh_from_one_exponential():
    """hyperbolic_prof on the device (ray_physics.cuh: tanh_cosh) forms tanh(a) = 1 - 2/(e^2a + 1) and cosh(a) = (e^a + 1/e^a)/2 from ONE
    exponential.  The same formulas in IEEE double against extended precision: absolute error of tanh and relative error of cosh and of
    1/cosh^2 (what the profile and its derivative use) stay at rounding level over the argument range of the profiles, limits as IEEE's."""
    a = np.concatenate([np.linspace(-2.0, 45.0, 400001), np.linspace(-1e-3, 1e-3, 20001)])
    e = np.exp(a)
    ch = 0.5 * (e + 1.0 / e)
    th = 1.0 - 2.0 / (e * e + 1.0)
    al = a.astype(np.longdouble)
    assert float(np.max(np.abs(th - np.tanh(al)))) < 5e-16
    assert float(np.max(np.abs((ch - np.cosh(al)) / np.cosh(al)))) < 5e-16
    s2, s2r = 1.0 / (ch * ch), 1.0 / (np.cosh(al) * np.cosh(al))
    assert float(np.max(np.abs((s2 - s2r) / s2r))) < 1e-15
    with np.errstate(over="ignore", divide="ignore"):
        big = np.array([800.0, -800.0])
        eb = np.exp(big)
        assert np.array_equal(1.0 - 2.0 / (eb * eb + 1.0), [1.0, -1.0]) and np.all(np.isinf(0.5 * (eb + 1.0 / eb)))


def test_integer_powers_in_double_double_are_the_correctly_rounded_powers():
    """pow_ool (ray_physics.cuh) forms x**n, 3 <= n <= 16, in double-double arithmetic.  tests/dd_pow_check.c is the same algorithm on the
    host: against the power accumulated in 113-bit arithmetic and rounded to double it agrees in every one of 2.8 M cases (glibc's pow,
    which the reference calls, differs from that in about 0.1 % of them by one ulp; CUDA's pow by up to two)."""
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "dd_pow_check")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(here, "dd_pow_check.c"), "-lm", "-lquadmath"], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("dd power: 0 of "), out
