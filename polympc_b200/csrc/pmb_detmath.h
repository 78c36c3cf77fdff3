// pmb_detmath.h — deterministic fp64 elementary functions, identical on host (g++) and device (nvcc).
//
// Why this exists: the SQP path takes data-dependent decisions (line-search acceptance, ADMM termination
// trip, BFGS damping branch, LDLT pivot order).  glibc and the CUDA math library round sin/cos/exp/... differently
// (both are within ~1 ulp but not bit-identical), so a kernel calling ::sin could flip a decision relative to the CPU
// path.  Every transcendental that the engine or a problem functor evaluates therefore goes through this header, which
// uses only IEEE-754 +,-,*,/,sqrt,fma and integer bit moves — operations that are correctly rounded on both sides.
// Compile host code with -ffp-contract=off and device code with -fmad=false so that no implicit contraction happens;
// fused multiply-adds are always spelled pmb::dm::fma().
//
// Algorithms: argument reduction + minimax polynomials in the style of Sun's freely distributable fdlibm
// (coefficients S1..S6, C1..C6, P1..P5, Lg1..Lg7, aT0..aT10 are the published fdlibm minimax coefficients);
// accuracy is checked against libm in tests/test_detmath.py (<= 2 ulp for sin/cos/exp/log/atan).
//
// Replaces, for the purposes of the hot path, the std:: calls made by the reference's AutoDiffScalar chain rules
// (reference: src/autodiff/AutoDiffScalar.h:592-684) and by the problem functors
// (reference: tests/control/mpc_wrapper_test.cpp:46-55, tests/control/cstr_control_test.cpp:63-100).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#if defined(__CUDACC__)
#define PMB_HD __host__ __device__ __forceinline__
#define PMB_DM_NOINLINE __host__ __device__ __noinline__ inline
#else
#define PMB_HD inline
#define PMB_DM_NOINLINE __attribute__((noinline)) inline
#endif

namespace pmb {
namespace dm {

PMB_HD double from_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; std::memcpy(&d, &u, sizeof d); return d;
#endif
}
PMB_HD uint64_t to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; std::memcpy(&u, &d, sizeof u); return u;
#endif
}

/** fused multiply-add, single rounding on both sides */
PMB_HD double fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

PMB_HD double fabs(double x) { return from_bits(to_bits(x) & 0x7fffffffffffffffULL); }
PMB_HD bool   isnan(double x) { return x != x; }
PMB_HD double inf() { return from_bits(0x7ff0000000000000ULL); }
PMB_HD double nan() { return from_bits(0x7ff8000000000000ULL); }

PMB_HD double sqrt(double x)
{
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return std::sqrt(x);
#endif
}

PMB_HD double floor(double x)
{
#if defined(__CUDA_ARCH__)
    return ::floor(x);
#else
    return std::floor(x);
#endif
}

/** comparisons with the exact semantics of Eigen's cwiseMax/cwiseMin (std::max/std::min): (a<b)?b:a */
PMB_HD double max(double a, double b) { return (a < b) ? b : a; }
PMB_HD double min(double a, double b) { return (b < a) ? b : a; }

/** 2^k for any integer k (handles subnormal results by two-step scaling) */
PMB_HD double pow2i(int k)
{
    if (k > 1023) return inf();
    if (k >= -1022) return from_bits((uint64_t)(k + 1023) << 52);
    if (k < -1074) return 0.0;
    return from_bits((uint64_t)(k + 1023 + 200) << 52) * from_bits((uint64_t)(1023 - 200) << 52);
}

// ---------------------------------------------------------------------------------------------------------
// sin / cos
// ---------------------------------------------------------------------------------------------------------
namespace detail {

PMB_HD double k_sin(double x, double y, int iy)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

PMB_HD double k_cos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double ax = fabs(x);
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    // qx ~ |x|/4 with few significant bits so that 1 - qx is exact
    double qx;
    if (ax < 0.3) qx = 0.0;
    else if (ax > 0.78125) qx = 0.28125;
    else qx = floor(ax * 262144.0) * (1.0 / 1048576.0);
    const double hz = 0.5 * z - qx;
    const double a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}

/** x = n*(pi/2) + (y0 + y1), |y0+y1| <= pi/4 (+eps); returns n mod 4. Valid for |x| < ~1.6e6. */
PMB_HD int rem_pio2(double x, double& y0, double& y1)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1  = 1.57079632673412561417e+00;  // first 33 bits of pi/2
    const double pio2_1t = 6.07710050650619224932e-11;  // pi/2 - pio2_1
    const double pio2_2  = 6.07710050630396597660e-11;  // second 33 bits
    const double pio2_2t = 2.02226624879595063154e-21;
    const double pio2_3  = 2.02226624871116645580e-21;  // third 33 bits
    const double pio2_3t = 8.47842766036889956997e-32;
    (void)pio2_1t; (void)pio2_2t;
    const double fn = floor(x * invpio2 + 0.5);
    // three unconditional Cody-Waite rounds (fn*pio2_k are exact: 33-bit constants, |fn| < 2^20)
    const double r1 = x - fn * pio2_1;
    const double w2 = fn * pio2_2;
    const double r2 = r1 - w2;
    const double e2 = (r1 - r2) - w2;             // r1 - w2 == r2 + e2 exactly
    const double w3 = fn * pio2_3;
    const double r3 = r2 - w3;
    const double e3 = (r2 - r3) - w3;             // r2 - w3 == r3 + e3 exactly
    const double tail = (fn * pio2_3t - e3) - e2; // what is still to be subtracted from r3
    y0 = r3 - tail;
    y1 = (r3 - y0) - tail;
    // n mod 4 without 64-bit conversion issues: fn is an exact integer in double
    const double q = fn - 4.0 * floor(fn * 0.25);
    return (int)q;
}

/** exact remainder ax mod C for finite ax >= 0 (every subtraction is exact by Sterbenz' lemma: y <= ax < 2y) */
PMB_HD double fmod_pos(double ax, double C)
{
    if (ax < C) return ax;
    const int ex = (int)((to_bits(ax) >> 52) & 0x7ff) - (int)((to_bits(C) >> 52) & 0x7ff);
    double y = C * pow2i(ex);
    if (y > ax) y *= 0.5;
    while (y >= C) {
        if (ax >= y) ax -= y;
        y *= 0.5;
    }
    return ax;
}

/** arguments beyond the range of the Cody-Waite reduction are first folded, exactly, modulo C = fl(2^18 * 2 pi): the result
 *  stays a sine/cosine of a nearby angle (|value| <= 1, phase error <= |x| * 2^-53) instead of degenerating — at such
 *  magnitudes the spacing of doubles exceeds 2 pi anyway.  Same operations on host and device. */
constexpr double TRIG_FOLD = 1647099.3291652855;   // fl(2^20 * pi / 2)
PMB_HD double fold_large(double x)
{
    const double C = TRIG_FOLD;
    if (fabs(x) < C) return x;
    const double r = fmod_pos(fabs(x), C);
    return x < 0.0 ? -r : r;
}

} // namespace detail

namespace detail {
/** sine / cosine of a finite |x| < TRIG_FOLD (the range of the Cody-Waite reduction) */
PMB_HD double sin_core(double x)
{
    if (fabs(x) <= 0.78539816339744830962) {
        if (fabs(x) < 7.450580596923828125e-09) return x;  // 2^-27
        return k_sin(x, 0.0, 0);
    }
    double y0, y1;
    const int n = rem_pio2(x, y0, y1);
    // quadrant 0: sin, 1: cos, 2: -sin, 3: -cos — one copy of each polynomial instead of two (code size)
    const double v = (n & 1) ? k_cos(y0, y1) : k_sin(y0, y1, 1);
    return (n & 2) ? -v : v;
}
PMB_HD double cos_core(double x)
{
    if (fabs(x) <= 0.78539816339744830962) {
        if (fabs(x) < 7.450580596923828125e-09) return 1.0;
        return k_cos(x, 0.0);
    }
    double y0, y1;
    const int n = rem_pio2(x, y0, y1);
    // quadrant 0: cos, 1: -sin, 2: -cos, 3: sin
    const double v = (n & 1) ? k_sin(y0, y1, 1) : k_cos(y0, y1);
    return ((n + 1) & 2) ? -v : v;
}
/** everything that is not a finite |x| < TRIG_FOLD: NaN, +-inf, huge arguments.  Out of line on purpose: the hot kernels
 *  inline sin / cos at every functor call site and never take this path (instruction-cache footprint). */
PMB_DM_NOINLINE double sin_rare(double x)
{
    if (isnan(x) || fabs(x) == inf()) return nan();
    return sin_core(fold_large(x));
}
PMB_DM_NOINLINE double cos_rare(double x)
{
    if (isnan(x) || fabs(x) == inf()) return nan();
    return cos_core(fold_large(x));
}
} // namespace detail

#ifdef PMB_DM_LIBM   /* oracle 'native' flavour only (oracle/Makefile): the C library instead of the deterministic algorithms */
PMB_HD double sin(double x) { return std::sin(x); }
PMB_HD double cos(double x) { return std::cos(x); }
#else
PMB_HD double sin(double x) { return (fabs(x) < detail::TRIG_FOLD) ? detail::sin_core(x) : detail::sin_rare(x); }
PMB_HD double cos(double x) { return (fabs(x) < detail::TRIG_FOLD) ? detail::cos_core(x) : detail::cos_rare(x); }
#endif

PMB_HD double tan(double x) { return sin(x) / cos(x); }

// ---------------------------------------------------------------------------------------------------------
// exp / log
// ---------------------------------------------------------------------------------------------------------
PMB_HD double exp(double x)
{
#ifdef PMB_DM_LIBM
    return std::exp(x);
#endif
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    if (isnan(x)) return x;
    if (x > 7.09782712893383973096e+02) return inf();
    if (x < -7.45133219101941108420e+02) return 0.0;
    const double ax = fabs(x);
    if (ax < 3.725290298461914e-09) return 1.0 + x;  // 2^-28
    int k = 0;
    double hi = x, lo = 0.0;
    if (ax > 0.34657359027997264) {  // 0.5 ln2
        const double fk = floor(invln2 * x + 0.5);
        k = (int)fk;
        hi = x - fk * ln2HI;
        lo = fk * ln2LO;
    }
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((r * c) / (c - 2.0) - r);
    const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    if (k >= -1021 && k <= 1022) return y * pow2i(k);
    if (k > 1022) return (y * pow2i(k - 2)) * 4.0;
    return (y * pow2i(k + 100)) * pow2i(-100);
}

PMB_HD double log(double x)
{
#ifdef PMB_DM_LIBM
    return std::log(x);
#endif
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    if (isnan(x)) return x;
    if (x < 0.0) return nan();
    if (x == 0.0) return -inf();
    if (x == inf()) return x;
    int k = 0;
    uint64_t u = to_bits(x);
    if ((u >> 52) == 0) {  // subnormal: scale up by 2^54
        x *= 18014398509481984.0;
        u = to_bits(x);
        k -= 54;
    }
    k += (int)(u >> 52) - 1023;
    // mantissa m in [1,2)
    double m = from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m *= 0.5; k += 1; }
    const double f = m - 1.0;
    const double dk = (double)k;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

// ---------------------------------------------------------------------------------------------------------
// atan / atan2 / asin / acos
// ---------------------------------------------------------------------------------------------------------
PMB_HD double atan(double x)
{
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    if (isnan(x)) return x;
    const bool neg = x < 0.0 || (x == 0.0 && (to_bits(x) >> 63));
    double ax = fabs(x);
    int id;
    double hi = 0.0, lo = 0.0;
    if (ax >= 7.3786976294838206464e+19) {  // 2^66
        const double r = 1.57079632679489655800e+00 + 6.12323399573676603587e-17;
        return neg ? -r : r;
    }
    if (ax < 0.4375) {
        if (ax < 1.862645149230957e-09) return x;  // 2^-29
        id = -1;
    } else if (ax < 0.6875) {
        id = 0; ax = (2.0 * ax - 1.0) / (2.0 + ax);
        hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17;
    } else if (ax < 1.1875) {
        id = 1; ax = (ax - 1.0) / (ax + 1.0);
        hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17;
    } else if (ax < 2.4375) {
        id = 2; ax = (ax - 1.5) / (1.0 + 1.5 * ax);
        hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17;
    } else {
        id = 3; ax = -1.0 / ax;
        hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17;
    }
    const double z = ax * ax;
    const double w = z * z;
    const double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    double r;
    if (id < 0) r = ax - ax * (s1 + s2);
    else        r = hi - ((ax * (s1 + s2) - lo) - ax);
    return neg ? -r : r;
}

PMB_HD double atan2(double y, double x)
{
#ifdef PMB_DM_LIBM
    return std::atan2(y, x);
#endif
    const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
    if (isnan(x) || isnan(y)) return nan();
    const bool yneg = (to_bits(y) >> 63) != 0;
    const bool xneg = (to_bits(x) >> 63) != 0;
    if (y == 0.0) {
        if (!xneg) return y;
        return yneg ? -pi : pi;
    }
    if (x == 0.0) return yneg ? -0.5 * pi : 0.5 * pi;
    if (fabs(x) == inf()) {
        if (fabs(y) == inf()) {
            const double r = xneg ? 0.75 * pi : 0.25 * pi;
            return yneg ? -r : r;
        }
        const double r = xneg ? pi : 0.0;
        return yneg ? -r : r;
    }
    if (fabs(y) == inf()) return yneg ? -0.5 * pi : 0.5 * pi;
    const double z = atan(fabs(y / x));
    if (!xneg) return yneg ? -z : z;
    const double r = pi - (z - pi_lo);
    return yneg ? -r : r;
}

PMB_HD double asin(double x)
{
#ifdef PMB_DM_LIBM
    return std::asin(x);
#endif
    if (fabs(x) > 1.0) return nan();
    return atan2(x, sqrt((1.0 - x) * (1.0 + x)));
}
PMB_HD double acos(double x)
{
#ifdef PMB_DM_LIBM
    return std::acos(x);
#endif
    if (fabs(x) > 1.0) return nan();
    return atan2(sqrt((1.0 - x) * (1.0 + x)), x);
}

// ---------------------------------------------------------------------------------------------------------
// hyperbolics, pow
// ---------------------------------------------------------------------------------------------------------
PMB_HD double sinh(double x)
{
#ifdef PMB_DM_LIBM
    return std::sinh(x);
#endif
    const double ax = fabs(x);
    if (ax < 0.5) {
        const double z = x * x;
        // x * (1 + z/6 + z^2/120 + ... + z^6/13!)
        const double p = 1.0 + z * (1.0 / 6.0 + z * (1.0 / 120.0 + z * (1.0 / 5040.0 + z * (1.0 / 362880.0 +
                         z * (1.0 / 39916800.0 + z * (1.0 / 6227020800.0))))));
        return x * p;
    }
    if (ax > 709.0) {
        const double e = exp(0.5 * ax);
        const double r = (0.5 * e) * e;
        return x < 0.0 ? -r : r;
    }
    const double e = exp(ax);
    const double r = 0.5 * (e - 1.0 / e);
    return x < 0.0 ? -r : r;
}
PMB_HD double cosh(double x)
{
#ifdef PMB_DM_LIBM
    return std::cosh(x);
#endif
    const double ax = fabs(x);
    if (ax > 709.0) {
        const double e = exp(0.5 * ax);
        return (0.5 * e) * e;
    }
    const double e = exp(ax);
    return 0.5 * (e + 1.0 / e);
}
PMB_HD double tanh(double x)
{
#ifdef PMB_DM_LIBM
    return std::tanh(x);
#endif
    const double ax = fabs(x);
    if (ax < 0.5) return sinh(x) / cosh(x);
    if (ax > 22.0) return x < 0.0 ? -1.0 : 1.0;
    const double r = 1.0 - 2.0 / (exp(2.0 * ax) + 1.0);
    return x < 0.0 ? -r : r;
}

PMB_HD double pow(double x, double y)
{
#ifdef PMB_DM_LIBM
    return std::pow(x, y);
#endif
    if (y == 0.0) return 1.0;
    if (isnan(x) || isnan(y)) return nan();
    const double fy = floor(y);
    if (fy == y && fabs(y) <= 1024.0) {
        // integer exponent: square-and-multiply (exact operation order, no log/exp)
        int n = (int)fabs(y);
        double base = x, r = 1.0;
        while (n) {
            if (n & 1) r *= base;
            base *= base;
            n >>= 1;
        }
        return y < 0.0 ? 1.0 / r : r;
    }
    if (x < 0.0) return nan();
    if (x == 0.0) return y > 0.0 ? 0.0 : inf();
    return exp(y * log(x));
}

/** function table of pmb_dm_eval (include/polympc_b200.h, enum pmb_dm_fn) */
PMB_HD double dm_dispatch(int fn, double x, double y)
{
    switch (fn) {
    case 0: return sin(x);
    case 1: return cos(x);
    case 2: return tan(x);
    case 3: return exp(x);
    case 4: return log(x);
    case 5: return atan2(x, y);
    case 6: return asin(x);
    case 7: return acos(x);
    case 8: return sinh(x);
    case 9: return cosh(x);
    case 10: return tanh(x);
    case 11: return pow(x, y);
    case 12: return sqrt(x);
    }
    return nan();
}

} // namespace dm
} // namespace pmb
