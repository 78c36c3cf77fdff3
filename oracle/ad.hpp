// oracle/ad.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Forward-mode dual numbers restating the semantics of the reference's vendored Eigen::AutoDiffScalar
// (reference: src/autodiff/AutoDiffScalar.h).  AD<S,n> carries a value of type S and n derivatives of type S.
//   first order : AD<double,n>              (reference ad_scalar_t,  continuous_ocp.hpp:124-125)
//   second order: AD<AD<double,n>,n>        (reference ad2_scalar_t, continuous_ocp.hpp:126,137) — nested, so the
//                 Hessian entry (i,j) is  x.d[i].d[j]  and the gradient is  x.v.d[j].
// Operand orders of every chain rule follow the reference line by line:
//   a*b  : (a.v*b.v , a.d*b.v + b.d*a.v)                                   AutoDiffScalar.h:341-351
//   a/b  : (a.v/b.v , (a.d*b.v - b.d*a.v) * (1/(b.v*b.v)))                 AutoDiffScalar.h:327-339
//   a*s  : (a.v*s   , a.d*s)          s*a : (a.v*s , a.d*s)                AutoDiffScalar.h:269-279
//   a/s  : (a.v/s   , a.d*(1/s))      s/a : (s/a.v , a.d*((-s)/(a.v*a.v))) AutoDiffScalar.h:295-305
//   unary functions                                                         AutoDiffScalar.h:592-684
// Transcendentals bottom out in pmb::dm (deterministic, shared with the device code).
#pragma once
#include "canon.hpp"
#include <type_traits>

namespace orc {

// scalar leaf functions (T = double)
inline double sin(double x) { return dm::sin(x); }
inline double cos(double x) { return dm::cos(x); }
inline double tan(double x) { return dm::tan(x); }
inline double exp(double x) { return dm::exp(x); }
inline double log(double x) { return dm::log(x); }
inline double sqrt(double x) { return dm::sqrt(x); }
inline double atan2(double y, double x) { return dm::atan2(y, x); }
inline double asin(double x) { return dm::asin(x); }
inline double acos(double x) { return dm::acos(x); }
inline double sinh(double x) { return dm::sinh(x); }
inline double cosh(double x) { return dm::cosh(x); }
inline double tanh(double x) { return dm::tanh(x); }
inline double pow(double x, double y) { return dm::pow(x, y); }
inline double abs(double x) { return dm::fabs(x); }
inline double abs2(double x) { return x * x; }
inline double value_of(double x) { return x; }

template <class S, int n>
struct AD {
    S v;
    S d[n > 0 ? n : 1];

    AD() {}
    /** AutoDiffScalar(const Real& value): derivatives zeroed (AutoDiffScalar.h:97-103) */
    AD(const S& value) : v(value) { for (int i = 0; i < n; ++i) d[i] = S(0.0); }
    /** nested type from a plain double (two-step conversion double -> S -> AD<S,n>) */
    template <class U = S, class = typename std::enable_if<!std::is_same<U, double>::value>::type>
    AD(double value) : v(S(value)) { for (int i = 0; i < n; ++i) d[i] = S(0.0); }
};

template <class T> struct is_ad : std::false_type {};
template <class S, int n> struct is_ad<AD<S, n>> : std::true_type {};

// ---- AD (+,-) AD
template <class S, int n> inline AD<S, n> operator+(const AD<S, n>& a, const AD<S, n>& b)
{ AD<S, n> r; r.v = a.v + b.v; for (int i = 0; i < n; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <class S, int n> inline AD<S, n> operator-(const AD<S, n>& a, const AD<S, n>& b)
{ AD<S, n> r; r.v = a.v - b.v; for (int i = 0; i < n; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <class S, int n> inline AD<S, n> operator-(const AD<S, n>& a)
{ AD<S, n> r; r.v = -a.v; for (int i = 0; i < n; ++i) r.d[i] = -a.d[i]; return r; }

// ---- AD (+,-) Scalar   (AutoDiffScalar.h:186-237): derivatives untouched / negated
template <class S, int n> inline AD<S, n> add_s(const AD<S, n>& a, const S& s)
{ AD<S, n> r; r.v = a.v + s; for (int i = 0; i < n; ++i) r.d[i] = a.d[i]; return r; }
template <class S, int n> inline AD<S, n> s_add(const S& s, const AD<S, n>& a)
{ AD<S, n> r; r.v = s + a.v; for (int i = 0; i < n; ++i) r.d[i] = a.d[i]; return r; }
template <class S, int n> inline AD<S, n> sub_s(const AD<S, n>& a, const S& s)
{ AD<S, n> r; r.v = a.v - s; for (int i = 0; i < n; ++i) r.d[i] = a.d[i]; return r; }
template <class S, int n> inline AD<S, n> s_sub(const S& s, const AD<S, n>& a)
{ AD<S, n> r; r.v = s - a.v; for (int i = 0; i < n; ++i) r.d[i] = -a.d[i]; return r; }
// ---- AD (*,/) Scalar
template <class S, int n> inline AD<S, n> mul_s(const AD<S, n>& a, const S& s)
{ AD<S, n> r; r.v = a.v * s; for (int i = 0; i < n; ++i) r.d[i] = a.d[i] * s; return r; }
template <class S, int n> inline AD<S, n> div_s(const AD<S, n>& a, const S& s)
{ AD<S, n> r; r.v = a.v / s; const S inv = S(1.0) / s; for (int i = 0; i < n; ++i) r.d[i] = a.d[i] * inv; return r; }
template <class S, int n> inline AD<S, n> s_div(const S& s, const AD<S, n>& a)
{ AD<S, n> r; r.v = s / a.v; const S f = S(-s) / (a.v * a.v); for (int i = 0; i < n; ++i) r.d[i] = a.d[i] * f; return r; }

#define ORC_AD_SCALAR_OPS(SCALAR_T)                                                                                   \
    template <class S, int n> inline AD<S, n> operator+(const AD<S, n>& a, SCALAR_T s) { return add_s(a, S(s)); }      \
    template <class S, int n> inline AD<S, n> operator+(SCALAR_T s, const AD<S, n>& a) { return s_add(S(s), a); }      \
    template <class S, int n> inline AD<S, n> operator-(const AD<S, n>& a, SCALAR_T s) { return sub_s(a, S(s)); }      \
    template <class S, int n> inline AD<S, n> operator-(SCALAR_T s, const AD<S, n>& a) { return s_sub(S(s), a); }      \
    template <class S, int n> inline AD<S, n> operator*(const AD<S, n>& a, SCALAR_T s) { return mul_s(a, S(s)); }      \
    template <class S, int n> inline AD<S, n> operator*(SCALAR_T s, const AD<S, n>& a) { return mul_s(a, S(s)); }      \
    template <class S, int n> inline AD<S, n> operator/(const AD<S, n>& a, SCALAR_T s) { return div_s(a, S(s)); }      \
    template <class S, int n> inline AD<S, n> operator/(SCALAR_T s, const AD<S, n>& a) { return s_div(S(s), a); }
ORC_AD_SCALAR_OPS(double)
ORC_AD_SCALAR_OPS(int)
#undef ORC_AD_SCALAR_OPS

// nested: AD<AD<double,n>,n> (op) AD<double,n>   — "Scalar" of the outer type is the inner AD
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator+(const AD<S, n>& a, const S& s) { return add_s(a, s); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator+(const S& s, const AD<S, n>& a) { return s_add(s, a); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator-(const AD<S, n>& a, const S& s) { return sub_s(a, s); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator-(const S& s, const AD<S, n>& a) { return s_sub(s, a); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator*(const AD<S, n>& a, const S& s) { return mul_s(a, s); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator*(const S& s, const AD<S, n>& a) { return mul_s(a, s); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator/(const AD<S, n>& a, const S& s) { return div_s(a, s); }
template <class S, int n, class = typename std::enable_if<is_ad<S>::value>::type>
inline AD<S, n> operator/(const S& s, const AD<S, n>& a) { return s_div(s, a); }

// ---- AD (*,/) AD
template <class S, int n> inline AD<S, n> operator*(const AD<S, n>& a, const AD<S, n>& b)
{
    AD<S, n> r;
    r.v = a.v * b.v;
    for (int i = 0; i < n; ++i) r.d[i] = (a.d[i] * b.v) + (b.d[i] * a.v);
    return r;
}
template <class S, int n> inline AD<S, n> operator/(const AD<S, n>& a, const AD<S, n>& b)
{
    AD<S, n> r;
    r.v = a.v / b.v;
    const S f = S(1.0) / (b.v * b.v);
    for (int i = 0; i < n; ++i) r.d[i] = ((a.d[i] * b.v) - (b.d[i] * a.v)) * f;
    return r;
}

#define ORC_AD_COMPOUND(OP)                                                                                           \
    template <class S, int n, class U> inline AD<S, n>& operator OP##=(AD<S, n>& a, const U& b) { a = a OP b; return a; }
ORC_AD_COMPOUND(+)
ORC_AD_COMPOUND(-)
ORC_AD_COMPOUND(*)
ORC_AD_COMPOUND(/)
#undef ORC_AD_COMPOUND

// comparisons act on values (AutoDiffScalar.h:162-184)
template <class S, int n> inline double value_of(const AD<S, n>& a) { return value_of(a.v); }
#define ORC_AD_CMP(OP)                                                                                                \
    template <class S, int n> inline bool operator OP(const AD<S, n>& a, const AD<S, n>& b) { return value_of(a) OP value_of(b); } \
    template <class S, int n> inline bool operator OP(const AD<S, n>& a, double b) { return value_of(a) OP b; }        \
    template <class S, int n> inline bool operator OP(double a, const AD<S, n>& b) { return a OP value_of(b); }
ORC_AD_CMP(<)
ORC_AD_CMP(<=)
ORC_AD_CMP(>)
ORC_AD_CMP(>=)
ORC_AD_CMP(==)
ORC_AD_CMP(!=)
#undef ORC_AD_CMP

// ---- unary functions: (f(x.v), x.d * f'(x.v))   AutoDiffScalar.h:592-684
template <class S, int n> inline AD<S, n> scale_d(const S& val, const AD<S, n>& x, const S& f)
{ AD<S, n> r; r.v = val; for (int i = 0; i < n; ++i) r.d[i] = x.d[i] * f; return r; }

template <class S, int n> inline AD<S, n> cos(const AD<S, n>& x) { return scale_d(cos(x.v), x, S(-sin(x.v))); }
template <class S, int n> inline AD<S, n> sin(const AD<S, n>& x) { return scale_d(sin(x.v), x, S(cos(x.v))); }
template <class S, int n> inline AD<S, n> exp(const AD<S, n>& x) { const S e = exp(x.v); return scale_d(e, x, e); }
template <class S, int n> inline AD<S, n> log(const AD<S, n>& x) { return scale_d(log(x.v), x, S(S(1.0) / x.v)); }
template <class S, int n> inline AD<S, n> sqrt(const AD<S, n>& x) { const S s = sqrt(x.v); return scale_d(s, x, S(S(0.5) / s)); }
template <class S, int n> inline AD<S, n> abs2(const AD<S, n>& x) { return scale_d(abs2(x.v), x, S(S(2.0) * x.v)); }
template <class S, int n> inline AD<S, n> abs(const AD<S, n>& x)
{ return scale_d(abs(x.v), x, (value_of(x.v) < 0.0) ? S(-1.0) : S(1.0)); }
template <class S, int n> inline AD<S, n> tan(const AD<S, n>& x) { return scale_d(tan(x.v), x, S(S(1.0) / abs2(cos(x.v)))); }
template <class S, int n> inline AD<S, n> asin(const AD<S, n>& x)
{ return scale_d(asin(x.v), x, S(S(1.0) / sqrt(S(1.0) - abs2(x.v)))); }
template <class S, int n> inline AD<S, n> acos(const AD<S, n>& x)
{ return scale_d(acos(x.v), x, S(S(-1.0) / sqrt(S(1.0) - abs2(x.v)))); }
template <class S, int n> inline AD<S, n> tanh(const AD<S, n>& x) { return scale_d(tanh(x.v), x, S(S(1.0) / abs2(cosh(x.v)))); }
template <class S, int n> inline AD<S, n> sinh(const AD<S, n>& x) { return scale_d(sinh(x.v), x, S(cosh(x.v))); }
template <class S, int n> inline AD<S, n> cosh(const AD<S, n>& x) { return scale_d(cosh(x.v), x, S(sinh(x.v))); }
/** pow with a plain exponent (AutoDiffScalar.h:621-629): x.d * (y * pow(x.v, y-1)) */
template <class S, int n> inline AD<S, n> pow(const AD<S, n>& x, double y)
{ return scale_d(pow(x.v, y), x, S(y * pow(x.v, y - 1.0))); }
/** atan2 (AutoDiffScalar.h:631-647) */
template <class S, int n> inline AD<S, n> atan2(const AD<S, n>& a, const AD<S, n>& b)
{
    AD<S, n> r;
    r.v = atan2(a.v, b.v);
    const S sq = a.v * a.v + b.v * b.v;
    for (int i = 0; i < n; ++i) r.d[i] = (a.d[i] * b.v - a.v * b.d[i]) / sq;
    return r;
}

} // namespace orc
