// pmb_cheb.hpp — Chebyshev-Gauss-Lobatto tables computed once on the host at problem creation and replicated to every
// device inside the problem descriptor (reference: src/polynomials/ebyshev.hpp — compute_nodes 111-117,
// compute_int_weights 120-159, compute_diff_matrix 198-214).  cos() is pmb::dm::cos so that the tables do not depend on
// the host libm.  Row sums of the differentiation matrix use pairwise halving (Eigen's small fixed-size reduction order).
#pragma once
#include "pmb_detmath.h"

namespace pmb {

namespace detail {
inline double halving_sum(const double* v, int stride, int start, int len)
{
    if (len == 1) return v[start * stride];
    const int half = len / 2;
    return halving_sum(v, stride, start, half) + halving_sum(v, stride, start + half, len - half);
}
} // namespace detail

/** nodes[P+1] (descending from +1), D[(P+1)^2] column-major, w[P+1] Clenshaw-Curtis.  P <= 64. */
inline void cheb_tables(int P, double* nodes, double* D, double* w)
{
    const double PI = 3.14159265358979323846;
    const int n = P + 1;
    const double step = PI / P;
    double theta[65], v[65], c[65], Dn[65 * 65];
    for (int k = 0; k < n; ++k) { theta[k] = (double)k * step; nodes[k] = dm::cos(theta[k]); w[k] = 0.0; }

    for (int j = 0; j < P - 1; ++j) v[j] = 1.0;
    const int kmax = (P % 2 == 0) ? P / 2 - 1 : (P - 1) / 2;
    w[0] = (P % 2 == 0) ? 1.0 / ((double)P * (double)P - 1.0) : 1.0 / ((double)P * (double)P);
    w[P] = w[0];
    for (int k = 1; k <= kmax; ++k) {
        const double coef = 2.0 / (4.0 * (double)k * (double)k - 1.0);
        for (int j = 0; j < P - 1; ++j) v[j] -= coef * dm::cos((double)(2 * k) * theta[j + 1]);
    }
    if (P % 2 == 0) {
        const double den = (double)P * (double)P - 1.0;
        for (int j = 0; j < P - 1; ++j) v[j] -= dm::cos((double)P * theta[j + 1]) / den;
    }
    for (int j = 0; j < P - 1; ++j) w[j + 1] = (2.0 / (double)P) * v[j];

    for (int k = 0; k < n; ++k) c[k] = ((k % 2) ? -1.0 : 1.0) * ((k == 0 || k == P) ? 2.0 : 1.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double dx = nodes[i] - nodes[j];
            Dn[i + j * n] = (c[i] * (1.0 / c[j])) * (1.0 / (dx + (i == j ? 1.0 : 0.0)));
        }
    for (int i = 0; i < n; ++i) {
        const double rs = detail::halving_sum(Dn + i, n, 0, n);
        for (int j = 0; j < n; ++j) D[i + j * n] = Dn[i + j * n] - (i == j ? rs : 0.0);
    }
}

} // namespace pmb
