// oracle/precond.hpp — TEST INFRASTRUCTURE (CPU restatement; see oracle/README.md).  Never used by the product.
//
// RuizEquilibration of the reference (src/solvers/qp_preconditioners.hpp:113-230 dense compute, 236-300 sparse compute,
// 356-385 scale/unscale of the solution, 389-404 unscale of the QP data) restated on plain column-major arrays.
//
// [Eigen-ext] what Eigen's expression templates evaluate, coefficient by coefficient (parity unpinned: Eigen is absent):
//   diag(d) * H * diag(d)           -> (d_i * H_ij) * d_j                  (left-associated lazy diagonal products)
//   diag(e) * A * diag(d)           -> (e_i * A_ij) * d_j                  (DENSE compute, both unscale variants)
//   diag(e) * (A * diag(d))         -> e_i * (A_ij * d_j)                  (SPARSE compute: the parentheses are in the source)
//   (1/c) * diag(1/d) * H * diag(1/d) -> (((1/c) * (1/d_i)) * H_ij) * (1/d_j)   (scalar * DiagonalWrapper scales the diagonal)
//   v.mean()                        -> sequential ascending sum / size     (Eigen's order depends on the vector ISA)
//   lpNorm<Infinity> / maxCoeff     -> chain from the first coefficient (canon.hpp::norm_inf)
#pragma once
#include <vector>
#include <limits>
#include "canon.hpp"

namespace orc {

enum { PRECOND_IDENTITY = 0, PRECOND_RUIZ_DENSE = 1, PRECOND_RUIZ_SPARSE = 2 };

struct Ruiz {
    int N, M, variant;
    std::vector<double> D, E, mD, mE;
    double c = 1.0;
    Ruiz(int n, int m, int v) : N(n), M(m), variant(v), D(n, 1.0), E(m, 1.0), mD(n, 1.0), mE(m, 1.0) {}

    static double max_chain(double m, double v) { return m < v ? v : m; }   // std::max / cwiseMax

    /** qp_preconditioners.hpp:151-230 (variant DENSE) / 236-300 (variant SPARSE); everything is scaled in place */
    void compute(double* H, double* h, double* A, double* Al, double* Au, double* l, double* u)
    {
        const bool sparse = variant == PRECOND_RUIZ_SPARSE;
        const int max_iter = 4;
        c = 1.0;
        for (int i = 0; i < N; ++i) { mD[i] = 1.0; D[i] = 1.0; }
        for (int i = 0; i < M; ++i) { mE[i] = 1.0; E[i] = 1.0; }
        const double approx_zero = sparse ? 1e-4 : std::numeric_limits<double>::epsilon();
        const double tolerance = 1e-3;
        double scaling_norm = 10 * tolerance;
        std::vector<double> xa(N);
        for (int iter = 0; iter < max_iter && (1.0 - scaling_norm) >= tolerance; ++iter) {
            for (int i = 0; i < M; ++i) { double m = dm::fabs(A[i]); for (int j = 1; j < N; ++j) m = max_chain(m, dm::fabs(A[i + (size_t)j * M])); mE[i] = m; }
            for (int j = 0; j < N; ++j) { double m = dm::fabs(H[(size_t)j * N]); for (int i = 1; i < N; ++i) m = max_chain(m, dm::fabs(H[i + (size_t)j * N])); mD[j] = m; }
            for (int j = 0; j < N; ++j) {
                double m = M > 0 ? dm::fabs(A[(size_t)j * M]) : 0.0;
                for (int i = 1; i < M; ++i) m = max_chain(m, dm::fabs(A[i + (size_t)j * M]));
                xa[j] = m;
            }
            for (int j = 0; j < N; ++j) mD[j] = max_chain(mD[j], xa[j]);
            double mxD = mD[0], mnD = mD[0];
            for (int j = 1; j < N; ++j) { mxD = max_chain(mxD, mD[j]); if (mD[j] < mnD) mnD = mD[j]; }
            double mxE = M > 0 ? mE[0] : -std::numeric_limits<double>::infinity(), mnE = M > 0 ? mE[0] : std::numeric_limits<double>::infinity();
            for (int i = 1; i < M; ++i) { mxE = max_chain(mxE, mE[i]); if (mE[i] < mnE) mnE = mE[i]; }
            scaling_norm = max_chain(mxD, mxE);
            if (mnD < approx_zero) for (int k = 0; k < N; ++k) if (mD[k] < approx_zero) mD[k] = 1.0;
            if (mnE < approx_zero) for (int k = 0; k < M; ++k) if (mE[k] < approx_zero) mE[k] = 1.0;
            for (int k = 0; k < N; ++k) mD[k] = 1.0 / dm::sqrt(mD[k]);
            for (int k = 0; k < M; ++k) mE[k] = 1.0 / dm::sqrt(mE[k]);
            for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) H[i + (size_t)j * N] = (mD[i] * H[i + (size_t)j * N]) * mD[j];
            if (sparse) { for (int j = 0; j < N; ++j) for (int i = 0; i < M; ++i) A[i + (size_t)j * M] = mE[i] * (A[i + (size_t)j * M] * mD[j]); }
            else        { for (int j = 0; j < N; ++j) for (int i = 0; i < M; ++i) A[i + (size_t)j * M] = (mE[i] * A[i + (size_t)j * M]) * mD[j]; }
            for (int k = 0; k < N; ++k) h[k] = h[k] * mD[k];
            for (int k = 0; k < N; ++k) D[k] = D[k] * mD[k];
            for (int k = 0; k < M; ++k) E[k] = E[k] * mE[k];
            // gamma
            double sum = 0.0;
            for (int j = 0; j < N; ++j) { double m = dm::fabs(H[(size_t)j * N]); for (int i = 1; i < N; ++i) m = max_chain(m, dm::fabs(H[i + (size_t)j * N])); mD[j] = m; sum += m; }
            const double mean = sum / (double)N;
            double h_inf = norm_inf(h, N);
            if (!sparse) h_inf = h_inf > approx_zero ? h_inf : 1.0;
            const double gamma = 1.0 / max_chain(mean, h_inf);
            for (size_t k = 0; k < (size_t)N * N; ++k) H[k] *= gamma;
            for (int k = 0; k < N; ++k) h[k] *= gamma;
            c *= gamma;
        }
        for (int k = 0; k < M; ++k) { Au[k] = Au[k] * E[k]; Al[k] = Al[k] * E[k]; }
        for (int k = 0; k < N; ++k) { const double di = 1.0 / D[k]; l[k] = l[k] * di; u[k] = u[k] * di; }
    }

    /** qp_preconditioners.hpp:364-369 */
    void unscale_solution(double* x, double* y) const
    {
        const double ic = 1 / c;
        for (int k = 0; k < N; ++k) x[k] = x[k] * D[k];
        for (int k = 0; k < M; ++k) y[k] = ic * (y[k] * E[k]);
        for (int k = 0; k < N; ++k) y[M + k] = ic * (y[M + k] * (1.0 / D[k]));
    }

    /** qp_preconditioners.hpp:372-404 (the DENSE and SPARSE overloads evaluate the same coefficients) */
    void unscale_data(double* H, double* h, double* A, double* Al, double* Au, double* l, double* u) const
    {
        const double ic = 1 / c;
        std::vector<double> Di(N), Ei(M);
        for (int k = 0; k < N; ++k) Di[k] = 1.0 / D[k];
        for (int k = 0; k < M; ++k) Ei[k] = 1.0 / E[k];
        for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) H[i + (size_t)j * N] = ((ic * Di[i]) * H[i + (size_t)j * N]) * Di[j];
        for (int j = 0; j < N; ++j) for (int i = 0; i < M; ++i) A[i + (size_t)j * M] = (Ei[i] * A[i + (size_t)j * M]) * Di[j];
        for (int k = 0; k < N; ++k) h[k] = ic * (h[k] * Di[k]);
        for (int k = 0; k < M; ++k) { Au[k] = Au[k] * Ei[k]; Al[k] = Al[k] * Ei[k]; }
        for (int k = 0; k < N; ++k) { l[k] = l[k] * D[k]; u[k] = u[k] * D[k]; }
    }
};

/** LSFilter of the reference (src/solvers/line_search.hpp:30-98) on a fixed array; entry 0 is the front of the std::list */
struct LsFilter {
    static constexpr int CAP = 16;
    double cost[CAP], constr[CAP];
    int size = 0, max_depth = 10;
    double beta = 1e-5;
    void clear() { size = 0; }
    bool is_acceptable(double c0, double v0) const      // :66-75
    {
        for (int k = 0; k < size; ++k)
            if (((cost[k] - beta * constr[k]) <= c0) && ((constr[k] - beta * constr[k]) <= v0)) return false;
        return true;
    }
    void push_front(double c0, double v0)
    {
        for (int k = size; k > 0; --k) { cost[k] = cost[k - 1]; constr[k] = constr[k - 1]; }
        cost[0] = c0; constr[0] = v0; ++size;
    }
    void add(double c0, double v0)                       // :77-93
    {
        if (size < max_depth) {
            int w = 0;
            for (int k = 0; k < size; ++k)
                if (!((cost[k] >= c0) && (constr[k] >= v0))) { cost[w] = cost[k]; constr[w] = constr[k]; ++w; }
            size = w;
            push_front(c0, v0);
        } else {
            --size;
            push_front(c0, v0);
        }
    }
};

} // namespace orc
