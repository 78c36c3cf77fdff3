// pmb_dual.hpp — forward-mode dual numbers for host and device (product code).
//
// Replaces, on the GPU path, the reference's vendored Eigen::AutoDiffScalar (reference: src/autodiff/AutoDiffScalar.h).
//   Dual<double,n>            <->  ad_scalar_t   (reference src/control/continuous_ocp.hpp:124-125): value + n partials
//   Dual<Dual<double,n>,n>    <->  ad2_scalar_t  (continuous_ocp.hpp:126,137): nested; Hessian(i,j) = x.d[i].d[j]
// Every rule keeps the operand order of the reference so that results are reproducible term by term:
//   product   AutoDiffScalar.h:341-351   (a.d*b.v) + (b.d*a.v)
//   quotient  AutoDiffScalar.h:327-339   ((a.d*b.v) - (b.d*a.v)) * (1/(b.v*b.v))
//   scalar    AutoDiffScalar.h:186-305   a/s -> a.d*(1/s);  s/a -> a.d*((-s)/(a.v*a.v))
//   unary     AutoDiffScalar.h:592-684   f(x) -> (f(x.v), x.d * f'(x.v))
// Transcendentals bottom out in pmb::dm (pmb_detmath.h): bit-identical on host and device.
#pragma once
#include "pmb_detmath.h"
#include <type_traits>

namespace pmb {

// ---- leaves ----------------------------------------------------------------------------------------------------
PMB_HD double sin(double x) { return dm::sin(x); }
PMB_HD double cos(double x) { return dm::cos(x); }
PMB_HD double tan(double x) { return dm::tan(x); }
PMB_HD double exp(double x) { return dm::exp(x); }
PMB_HD double log(double x) { return dm::log(x); }
PMB_HD double sqrt(double x) { return dm::sqrt(x); }
PMB_HD double atan2(double y, double x) { return dm::atan2(y, x); }
PMB_HD double asin(double x) { return dm::asin(x); }
PMB_HD double acos(double x) { return dm::acos(x); }
PMB_HD double sinh(double x) { return dm::sinh(x); }
PMB_HD double cosh(double x) { return dm::cosh(x); }
PMB_HD double tanh(double x) { return dm::tanh(x); }
PMB_HD double pow(double x, double y) { return dm::pow(x, y); }
PMB_HD double abs(double x) { return dm::fabs(x); }
PMB_HD double abs2(double x) { return x * x; }
PMB_HD double value_of(double x) { return x; }

template <class S, int n>
struct Dual {
    S v;
    S d[n > 0 ? n : 1];

    PMB_HD Dual() {}
    /** value constructor: partials zeroed (AutoDiffScalar.h:97-103) */
    PMB_HD Dual(const S& value) : v(value)
    {
#pragma unroll
        for (int i = 0; i < n; ++i) d[i] = S(0.0);
    }
    /** nested type from a plain double */
    template <class U = S, class = typename std::enable_if<!std::is_same<U, double>::value>::type>
    PMB_HD Dual(double value) : v(S(value))
    {
#pragma unroll
        for (int i = 0; i < n; ++i) d[i] = S(0.0);
    }
};

template <class T> struct is_dual : std::false_type {};
template <class S, int n> struct is_dual<Dual<S, n>> : std::true_type {};

#define PMB_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < n; ++i)

template <class S, int n> PMB_HD Dual<S, n> operator+(const Dual<S, n>& a, const Dual<S, n>& b)
{ Dual<S, n> r; r.v = a.v + b.v; PMB_DUAL_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> operator-(const Dual<S, n>& a, const Dual<S, n>& b)
{ Dual<S, n> r; r.v = a.v - b.v; PMB_DUAL_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> operator-(const Dual<S, n>& a)
{ Dual<S, n> r; r.v = -a.v; PMB_DUAL_LOOP r.d[i] = -a.d[i]; return r; }

template <class S, int n> PMB_HD Dual<S, n> dual_add_s(const Dual<S, n>& a, const S& s)
{ Dual<S, n> r; r.v = a.v + s; PMB_DUAL_LOOP r.d[i] = a.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_s_add(const S& s, const Dual<S, n>& a)
{ Dual<S, n> r; r.v = s + a.v; PMB_DUAL_LOOP r.d[i] = a.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_sub_s(const Dual<S, n>& a, const S& s)
{ Dual<S, n> r; r.v = a.v - s; PMB_DUAL_LOOP r.d[i] = a.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_s_sub(const S& s, const Dual<S, n>& a)
{ Dual<S, n> r; r.v = s - a.v; PMB_DUAL_LOOP r.d[i] = -a.d[i]; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_mul_s(const Dual<S, n>& a, const S& s)
{ Dual<S, n> r; r.v = a.v * s; PMB_DUAL_LOOP r.d[i] = a.d[i] * s; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_div_s(const Dual<S, n>& a, const S& s)
{ Dual<S, n> r; r.v = a.v / s; const S inv = S(1.0) / s; PMB_DUAL_LOOP r.d[i] = a.d[i] * inv; return r; }
template <class S, int n> PMB_HD Dual<S, n> dual_s_div(const S& s, const Dual<S, n>& a)
{ Dual<S, n> r; r.v = s / a.v; const S f = S(-s) / (a.v * a.v); PMB_DUAL_LOOP r.d[i] = a.d[i] * f; return r; }

#define PMB_DUAL_SCALAR_OPS(SCALAR_T)                                                                                  \
    template <class S, int n> PMB_HD Dual<S, n> operator+(const Dual<S, n>& a, SCALAR_T s) { return dual_add_s(a, S(s)); } \
    template <class S, int n> PMB_HD Dual<S, n> operator+(SCALAR_T s, const Dual<S, n>& a) { return dual_s_add(S(s), a); } \
    template <class S, int n> PMB_HD Dual<S, n> operator-(const Dual<S, n>& a, SCALAR_T s) { return dual_sub_s(a, S(s)); } \
    template <class S, int n> PMB_HD Dual<S, n> operator-(SCALAR_T s, const Dual<S, n>& a) { return dual_s_sub(S(s), a); } \
    template <class S, int n> PMB_HD Dual<S, n> operator*(const Dual<S, n>& a, SCALAR_T s) { return dual_mul_s(a, S(s)); } \
    template <class S, int n> PMB_HD Dual<S, n> operator*(SCALAR_T s, const Dual<S, n>& a) { return dual_mul_s(a, S(s)); } \
    template <class S, int n> PMB_HD Dual<S, n> operator/(const Dual<S, n>& a, SCALAR_T s) { return dual_div_s(a, S(s)); } \
    template <class S, int n> PMB_HD Dual<S, n> operator/(SCALAR_T s, const Dual<S, n>& a) { return dual_s_div(S(s), a); }
PMB_DUAL_SCALAR_OPS(double)
PMB_DUAL_SCALAR_OPS(int)
#undef PMB_DUAL_SCALAR_OPS

// nested: Dual<Dual<double,n>,n> (op) Dual<double,n> — the "Scalar" of the outer type is the inner dual
#define PMB_DUAL_NESTED(OP, FN_AS, FN_SA)                                                                              \
    template <class S, int n, class = typename std::enable_if<is_dual<S>::value>::type>                                \
    PMB_HD Dual<S, n> operator OP(const Dual<S, n>& a, const S& s) { return FN_AS(a, s); }                             \
    template <class S, int n, class = typename std::enable_if<is_dual<S>::value>::type>                                \
    PMB_HD Dual<S, n> operator OP(const S& s, const Dual<S, n>& a) { return FN_SA(s, a); }
template <class S, int n> PMB_HD Dual<S, n> dual_s_mul(const S& s, const Dual<S, n>& a) { return dual_mul_s(a, s); }
PMB_DUAL_NESTED(+, dual_add_s, dual_s_add)
PMB_DUAL_NESTED(-, dual_sub_s, dual_s_sub)
PMB_DUAL_NESTED(*, dual_mul_s, dual_s_mul)
PMB_DUAL_NESTED(/, dual_div_s, dual_s_div)
#undef PMB_DUAL_NESTED

template <class S, int n> PMB_HD Dual<S, n> operator*(const Dual<S, n>& a, const Dual<S, n>& b)
{
    Dual<S, n> r;
    r.v = a.v * b.v;
    PMB_DUAL_LOOP r.d[i] = (a.d[i] * b.v) + (b.d[i] * a.v);
    return r;
}
template <class S, int n> PMB_HD Dual<S, n> operator/(const Dual<S, n>& a, const Dual<S, n>& b)
{
    Dual<S, n> r;
    r.v = a.v / b.v;
    const S f = S(1.0) / (b.v * b.v);
    PMB_DUAL_LOOP r.d[i] = ((a.d[i] * b.v) - (b.d[i] * a.v)) * f;
    return r;
}

#define PMB_DUAL_COMPOUND(OP)                                                                                          \
    template <class S, int n, class U> PMB_HD Dual<S, n>& operator OP##=(Dual<S, n>& a, const U& b) { a = a OP b; return a; }
PMB_DUAL_COMPOUND(+)
PMB_DUAL_COMPOUND(-)
PMB_DUAL_COMPOUND(*)
PMB_DUAL_COMPOUND(/)
#undef PMB_DUAL_COMPOUND

// comparisons act on values (AutoDiffScalar.h:162-184)
template <class S, int n> PMB_HD double value_of(const Dual<S, n>& a) { return value_of(a.v); }
#define PMB_DUAL_CMP(OP)                                                                                               \
    template <class S, int n> PMB_HD bool operator OP(const Dual<S, n>& a, const Dual<S, n>& b) { return value_of(a) OP value_of(b); } \
    template <class S, int n> PMB_HD bool operator OP(const Dual<S, n>& a, double b) { return value_of(a) OP b; }      \
    template <class S, int n> PMB_HD bool operator OP(double a, const Dual<S, n>& b) { return a OP value_of(b); }
PMB_DUAL_CMP(<)
PMB_DUAL_CMP(<=)
PMB_DUAL_CMP(>)
PMB_DUAL_CMP(>=)
PMB_DUAL_CMP(==)
PMB_DUAL_CMP(!=)
#undef PMB_DUAL_CMP

// ---- unary chain rules ------------------------------------------------------------------------------------------
template <class S, int n> PMB_HD Dual<S, n> dual_chain(const S& val, const Dual<S, n>& x, const S& f)
{ Dual<S, n> r; r.v = val; PMB_DUAL_LOOP r.d[i] = x.d[i] * f; return r; }

template <class S, int n> PMB_HD Dual<S, n> cos(const Dual<S, n>& x) { return dual_chain(cos(x.v), x, S(-sin(x.v))); }
template <class S, int n> PMB_HD Dual<S, n> sin(const Dual<S, n>& x) { return dual_chain(sin(x.v), x, S(cos(x.v))); }
template <class S, int n> PMB_HD Dual<S, n> exp(const Dual<S, n>& x) { const S e = exp(x.v); return dual_chain(e, x, e); }
template <class S, int n> PMB_HD Dual<S, n> log(const Dual<S, n>& x) { return dual_chain(log(x.v), x, S(S(1.0) / x.v)); }
template <class S, int n> PMB_HD Dual<S, n> sqrt(const Dual<S, n>& x) { const S s = sqrt(x.v); return dual_chain(s, x, S(S(0.5) / s)); }
template <class S, int n> PMB_HD Dual<S, n> abs2(const Dual<S, n>& x) { return dual_chain(abs2(x.v), x, S(S(2.0) * x.v)); }
template <class S, int n> PMB_HD Dual<S, n> abs(const Dual<S, n>& x)
{ return dual_chain(abs(x.v), x, (value_of(x.v) < 0.0) ? S(-1.0) : S(1.0)); }
template <class S, int n> PMB_HD Dual<S, n> tan(const Dual<S, n>& x) { return dual_chain(tan(x.v), x, S(S(1.0) / abs2(cos(x.v)))); }
template <class S, int n> PMB_HD Dual<S, n> asin(const Dual<S, n>& x)
{ return dual_chain(asin(x.v), x, S(S(1.0) / sqrt(S(1.0) - abs2(x.v)))); }
template <class S, int n> PMB_HD Dual<S, n> acos(const Dual<S, n>& x)
{ return dual_chain(acos(x.v), x, S(S(-1.0) / sqrt(S(1.0) - abs2(x.v)))); }
template <class S, int n> PMB_HD Dual<S, n> tanh(const Dual<S, n>& x) { return dual_chain(tanh(x.v), x, S(S(1.0) / abs2(cosh(x.v)))); }
template <class S, int n> PMB_HD Dual<S, n> sinh(const Dual<S, n>& x) { return dual_chain(sinh(x.v), x, S(cosh(x.v))); }
template <class S, int n> PMB_HD Dual<S, n> cosh(const Dual<S, n>& x) { return dual_chain(cosh(x.v), x, S(sinh(x.v))); }
/** pow with a plain exponent (AutoDiffScalar.h:621-629) */
template <class S, int n> PMB_HD Dual<S, n> pow(const Dual<S, n>& x, double y)
{ return dual_chain(pow(x.v, y), x, S(y * pow(x.v, y - 1.0))); }
/** atan2 (AutoDiffScalar.h:631-647) */
template <class S, int n> PMB_HD Dual<S, n> atan2(const Dual<S, n>& a, const Dual<S, n>& b)
{
    Dual<S, n> r;
    r.v = atan2(a.v, b.v);
    const S sq = a.v * a.v + b.v * b.v;
    PMB_DUAL_LOOP r.d[i] = (a.d[i] * b.v - a.v * b.d[i]) / sq;
    return r;
}
#undef PMB_DUAL_LOOP

/** Eigen's unrolled reduction order for small fixed sizes: sum(start,len) = sum(first half) + sum(second half) */
template <class T, int START, int LEN, class F>
struct HalvingSum {
    PMB_HD static T run(F& f) { return HalvingSum<T, START, LEN / 2, F>::run(f) + HalvingSum<T, START + LEN / 2, LEN - LEN / 2, F>::run(f); }
};
template <class T, int START, class F>
struct HalvingSum<T, START, 1, F> {
    PMB_HD static T run(F& f) { return f(START); }
};
template <class T, int n, class F> PMB_HD T hsum(F f) { return HalvingSum<T, 0, n, F>::run(f); }

} // namespace pmb
