// oracle/canon.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Canonical floating-point evaluation orders.  Eigen (absent here) does not specify its summation orders (they depend
// on SIMD width, unrolling limits and version), so the oracle fixes one order per reduction shape and the CUDA kernels
// reproduce exactly that order; this is what makes oracle-vs-kernel parity bit-exact instead of "within rounding".
//
//  * fixed-size reductions of length <= ~16 (Eigen::redux on small fixed vectors, row sums, 3x3 products inside problem
//    functors): Eigen's redux_novec_unroller binary-halving tree  [Eigen-ext: Eigen/src/Core/Redux.h]
//        sum(start,len) = len==1 ? x[start] : sum(start,len/2) + sum(start+len/2, len-len/2)
//  * matrix-vector rows, LDLT inner products, triangular solves: sequential ascending fused chain
//        acc = 0; for j ascending: acc = fma(a_j, b_j, acc)
//  * long dot products / 1-norms (BFGS s'Bs, s'y; line-search |c|_1, h'p): "tree32" —
//        partial[l] = sequential sum over i = l, l+32, l+64, ...   (l = 0..31)
//        then butterfly partial[l] += partial[l ^ off] for off = 16,8,4,2,1 ; result = partial[0]
//    (this is what a warp computes with lanes striding the vector and __shfl_xor reductions)
//  * max / inf-norm reductions: exact, order free.
#pragma once
#include "../polympc_b200/csrc/pmb_detmath.h"

namespace orc {

namespace dm = pmb::dm;

/** Eigen redux_novec_unroller order for a generic indexable term generator */
template <class T, class F>
inline T sum_halving(int start, int len, F&& term)
{
    if (len == 1) return term(start);
    const int half = len / 2;
    T a = sum_halving<T>(start, half, term);
    T b = sum_halving<T>(start + half, len - half, term);
    return a + b;
}

/** sequential fused chain: sum_i a[i*sa]*b[i*sb], ascending */
inline double dot_seq(const double* a, int sa, const double* b, int sb, int n)
{
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = dm::fma(a[i * sa], b[i * sb], acc);
    return acc;
}

/** tree32 sum of term(i), i = 0..n-1 */
template <class F>
inline double sum_tree32(int n, F&& term)
{
    double partial[32];
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int i = l; i < n; i += 32) acc = acc + term(i);
        partial[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double next[32];
        for (int l = 0; l < 32; ++l) next[l] = partial[l] + partial[l ^ off];
        for (int l = 0; l < 32; ++l) partial[l] = next[l];
    }
    return partial[0];
}

/** tree32 dot product with fused per-lane accumulation */
inline double dot_tree32(const double* a, const double* b, int n)
{
    double partial[32];
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int i = l; i < n; i += 32) acc = dm::fma(a[i], b[i], acc);
        partial[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double next[32];
        for (int l = 0; l < 32; ++l) next[l] = partial[l] + partial[l ^ off];
        for (int l = 0; l < 32; ++l) partial[l] = next[l];
    }
    return partial[0];
}

/** lpNorm<Infinity>() = cwiseAbs().maxCoeff(): a chain of std::max-like steps  m = (m < v) ? v : m  that starts from the
 *  first coefficient [Eigen-ext: Redux.h + scalar_max_op].  A NaN in the first coefficient therefore sticks (nothing
 *  compares greater than it), a NaN anywhere else is skipped; an all-NaN vector has norm NaN, so `norm <= eps` is false.
 *  (Eigen's packet path keeps one such chain per SIMD lane; which coefficients count as "first" then depends on the vector
 *  ISA — parity unpinned for partially-NaN vectors, the all-NaN / all-finite cases do not depend on it.) */
inline double norm_inf(const double* a, int n)
{
    if (n <= 0) return 0.0;
    double m = dm::fabs(a[0]);
    for (int i = 1; i < n; ++i) { const double v = dm::fabs(a[i]); if (m < v) m = v; }
    return m;
}

/** libm fmax/fmin semantics (NaN ignored) spelled out so host and device agree */
inline double fmax_(double a, double b) { if (a != a) return b; if (b != b) return a; return (a < b) ? b : a; }
inline double fmin_(double a, double b) { if (a != a) return b; if (b != b) return a; return (b < a) ? b : a; }

} // namespace orc
