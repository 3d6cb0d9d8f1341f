// oracle_real.hpp — scalar types for the CPU oracle (TEST INFRASTRUCTURE, not product code).
//
// The oracle is templated on its real type R:
//   * double      : the reference arithmetic (compile with -O2 -ffp-contract=off, no fast-math)
//   * CountReal   : a double that counts every +,-,*,/,sqrt and libm call as one flop, which gives
//                   the exact ALGORITHMIC flop count per ray-step used by bench.py's roofline
//                   ("replace with the exact count from the oracle's counting type").
#pragma once
#include <cmath>
#include <cstdint>

namespace rays_oracle {

struct FlopCounter {
    static uint64_t &n() {
        static thread_local uint64_t c = 0;
        return c;
    }
};

struct CountReal {
    double v;
    CountReal() : v(0.0) {}
    CountReal(double x) : v(x) {}
    explicit operator double() const { return v; }
};
#define RAYS_ORACLE_BINOP(op)                                                                  \
    inline CountReal operator op(CountReal a, CountReal b) {                                   \
        ++FlopCounter::n();                                                                    \
        return CountReal(a.v op b.v);                                                          \
    }                                                                                          \
    inline CountReal operator op(CountReal a, double b) {                                      \
        ++FlopCounter::n();                                                                    \
        return CountReal(a.v op b);                                                            \
    }                                                                                          \
    inline CountReal operator op(double a, CountReal b) {                                      \
        ++FlopCounter::n();                                                                    \
        return CountReal(a op b.v);                                                            \
    }
RAYS_ORACLE_BINOP(+)
RAYS_ORACLE_BINOP(-)
RAYS_ORACLE_BINOP(*)
RAYS_ORACLE_BINOP(/)
#undef RAYS_ORACLE_BINOP
inline CountReal operator-(CountReal a) { return CountReal(-a.v); }
inline CountReal &operator+=(CountReal &a, CountReal b) { a = a + b; return a; }
inline CountReal &operator-=(CountReal &a, CountReal b) { a = a - b; return a; }
inline CountReal &operator*=(CountReal &a, CountReal b) { a = a * b; return a; }
#define RAYS_ORACLE_CMP(op)                                                                    \
    inline bool operator op(CountReal a, CountReal b) { return a.v op b.v; }                   \
    inline bool operator op(CountReal a, double b) { return a.v op b; }                        \
    inline bool operator op(double a, CountReal b) { return a op b.v; }
RAYS_ORACLE_CMP(<)
RAYS_ORACLE_CMP(>)
RAYS_ORACLE_CMP(<=)
RAYS_ORACLE_CMP(>=)
RAYS_ORACLE_CMP(==)
RAYS_ORACLE_CMP(!=)
#undef RAYS_ORACLE_CMP

// value extraction / libm wrappers, overloaded for both scalar types
inline double val(double x) { return x; }
inline double val(CountReal x) { return x.v; }

inline double Sqrt(double x) { return std::sqrt(x); }
inline double Pow(double x, double y) { return std::pow(x, y); }
inline double Exp(double x) { return std::exp(x); }
inline double Tanh(double x) { return std::tanh(x); }
inline double Cosh(double x) { return std::cosh(x); }
inline double Cos(double x) { return std::cos(x); }
inline double Acos(double x) { return std::acos(x); }
inline double Fabs(double x) { return std::fabs(x); }
inline double Copysign(double a, double b) { return std::copysign(a, b); }
inline double Fmax(double a, double b) { return a > b ? a : b; }  // Fortran max(a,b)
inline double Fmin(double a, double b) { return a < b ? a : b; }  // Fortran min(a,b)

#define RAYS_ORACLE_F1(name, fn)                                                               \
    inline CountReal name(CountReal x) {                                                       \
        ++FlopCounter::n();                                                                    \
        return CountReal(fn(x.v));                                                             \
    }
RAYS_ORACLE_F1(Sqrt, std::sqrt)
RAYS_ORACLE_F1(Exp, std::exp)
RAYS_ORACLE_F1(Tanh, std::tanh)
RAYS_ORACLE_F1(Cosh, std::cosh)
RAYS_ORACLE_F1(Cos, std::cos)
RAYS_ORACLE_F1(Acos, std::acos)
#undef RAYS_ORACLE_F1
inline CountReal Pow(CountReal x, CountReal y) {
    ++FlopCounter::n();
    return CountReal(std::pow(x.v, y.v));
}
inline CountReal Fabs(CountReal x) { return CountReal(std::fabs(x.v)); }
inline CountReal Copysign(CountReal a, CountReal b) { return CountReal(std::copysign(a.v, b.v)); }
inline CountReal Fmax(CountReal a, CountReal b) { return a.v > b.v ? a : b; }
inline CountReal Fmin(CountReal a, CountReal b) { return a.v < b.v ? a : b; }

// The Fortran source is compiled without -fdefault-real-8: a default-kind
// literal such as 1.e-6 is SINGLE precision and is widened to double on use.
inline double f32lit(double x) { return (double)(float)x; }

// ---- complex arithmetic the way gfortran expands it -------------------------
template <class R> struct Cx {
    R re, im;
    Cx() : re(0.0), im(0.0) {}
    Cx(R r, R i) : re(r), im(i) {}
};
template <class R> inline Cx<R> operator+(Cx<R> a, Cx<R> b) { return Cx<R>(a.re + b.re, a.im + b.im); }
template <class R> inline Cx<R> operator-(Cx<R> a, Cx<R> b) { return Cx<R>(a.re - b.re, a.im - b.im); }
template <class R> inline Cx<R> operator-(Cx<R> a) { return Cx<R>(-a.re, -a.im); }
// (a+ib)(c+id) = (ac - bd) + i(ad + bc): -fcx-fortran-rules, no NaN recovery
template <class R> inline Cx<R> operator*(Cx<R> a, Cx<R> b) {
    return Cx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> inline Cx<R> cscale(R s, Cx<R> a) { return Cx<R>(s * a.re, s * a.im); }  // real*complex
template <class R> inline Cx<R> conjg(Cx<R> a) { return Cx<R>(a.re, -a.im); }
// complex division with Smith's range reduction (GCC expand_complex_div_wide)
template <class R> inline Cx<R> operator/(Cx<R> x, Cx<R> y) {
    R a = x.re, b = x.im, c = y.re, d = y.im;
    if (Fabs(c) < Fabs(d)) {
        R ratio = c / d;
        R denom = (c * ratio) + d;
        return Cx<R>(((a * ratio) + b) / denom, ((b * ratio) - a) / denom);
    } else {
        R ratio = d / c;
        R denom = (d * ratio) + c;
        return Cx<R>(((b * ratio) + a) / denom, (b - (a * ratio)) / denom);
    }
}
template <class R> inline R cabs(Cx<R> a) {  // abs(complex) = hypot
    double r = std::hypot(val(a.re), val(a.im));
    // count as mul,mul,add,sqrt
    R t = Sqrt(a.re * a.re + a.im * a.im);
    (void)t;
    return R(r);
}
template <class R> inline Cx<R> csqrt_(Cx<R> z) {  // principal square root, csqrt semantics for signed zero
    double x = val(z.re), y = val(z.im);
    double re, im;
    if (y == 0.0) {
        if (x >= 0.0) {
            re = std::sqrt(x);
            im = y;  // keeps the sign of the zero
        } else {
            re = 0.0;
            im = std::copysign(std::sqrt(-x), y);
        }
    } else {
        double m = std::hypot(x, y);
        if (x >= 0.0) {
            re = std::sqrt(0.5 * (m + x));
            im = y / (2.0 * re);
        } else {
            im = std::copysign(std::sqrt(0.5 * (m - x)), y);
            re = y / (2.0 * im);
        }
    }
    R cnt = Sqrt(z.re * z.re + z.im * z.im);  // flop accounting only
    (void)cnt;
    return Cx<R>(R(re), R(im));
}

}  // namespace rays_oracle
