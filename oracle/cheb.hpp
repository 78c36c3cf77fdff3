// oracle/cheb.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Chebyshev-Gauss-Lobatto tables, restating reference src/polynomials/ebyshev.hpp:
//   compute_nodes        111-117   x_k = cos(k * (pi/P))
//   compute_int_weights  120-159   Clenshaw-Curtis weights (even / odd P branches)
//   compute_diff_matrix  198-214   D = Dn - diag(rowsum(Dn)),  Dn_ij = (c_i * (1/c_j)) * (1/(x_i - x_j + delta_ij))
// cos() is pmb::dm::cos (deterministic).  Row sums use Eigen's unrolled halving order (canon.hpp).
#pragma once
#include "canon.hpp"
#include <vector>

namespace orc {

struct ChebTables {
    int P = 0;
    std::vector<double> nodes;  // P+1, descending from +1
    std::vector<double> D;      // (P+1)x(P+1) column-major: D[i + j*(P+1)]
    std::vector<double> w;      // P+1 Clenshaw-Curtis weights
    double d(int i, int j) const { return D[i + j * (P + 1)]; }
};

inline ChebTables cheb_tables(int P)
{
    const double PI = 3.14159265358979323846;
    ChebTables t;
    t.P = P;
    const int n = P + 1;
    t.nodes.resize(n);
    t.w.assign(n, 0.0);
    t.D.assign(n * n, 0.0);

    // ebyshev.hpp:114-116 : (grid * (M_PI / P)).cos()
    const double step = PI / P;
    for (int k = 0; k < n; ++k) t.nodes[k] = dm::cos((double)k * step);

    // ebyshev.hpp:124-158
    std::vector<double> theta(n), v(P - 1, 1.0);
    for (int k = 0; k < n; ++k) theta[k] = (double)k * step;
    if (P % 2 == 0) {
        t.w[0] = 1.0 / ((double)P * (double)P - 1.0);
        t.w[P] = t.w[0];
        for (int k = 1; k <= P / 2 - 1; ++k) {
            const double coef = 2.0 / (4.0 * (double)k * (double)k - 1.0);
            for (int j = 0; j < P - 1; ++j) v[j] -= coef * dm::cos((double)(2 * k) * theta[j + 1]);
        }
        const double den = (double)P * (double)P - 1.0;
        for (int j = 0; j < P - 1; ++j) v[j] -= dm::cos((double)P * theta[j + 1]) / den;
    } else {
        t.w[0] = 1.0 / ((double)P * (double)P);
        t.w[P] = t.w[0];
        for (int k = 1; k <= (P - 1) / 2; ++k) {
            const double coef = 2.0 / (4.0 * (double)k * (double)k - 1.0);
            for (int j = 0; j < P - 1; ++j) v[j] -= coef * dm::cos((double)(2 * k) * theta[j + 1]);
        }
    }
    for (int j = 0; j < P - 1; ++j) t.w[j + 1] = (2.0 / (double)P) * v[j];

    // ebyshev.hpp:202-213
    std::vector<double> c(n, 1.0);
    c[0] = 2.0; c[P] = 2.0;
    for (int k = 0; k < n; ++k) c[k] = ((k % 2) ? -1.0 : 1.0) * c[k];
    std::vector<double> Dn(n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double dx = t.nodes[i] - t.nodes[j];
            Dn[i + j * n] = (c[i] * (1.0 / c[j])) * (1.0 / (dx + (i == j ? 1.0 : 0.0)));
        }
    for (int i = 0; i < n; ++i) {
        const double rs = sum_halving<double>(0, n, [&](int j) { return Dn[i + j * n]; });
        for (int j = 0; j < n; ++j) t.D[i + j * n] = Dn[i + j * n] - (i == j ? rs : 0.0);
    }
    return t;
}

} // namespace orc
