// experiment only
#pragma once
#include <vector>
#include <cmath>
namespace orc {
inline int& ldlt_variant() { static int v = 0; return v; }
struct LdltFast {
    int n = 0;
    std::vector<double> L, Linv, dinv; // L: n x n col-major unit lower + D on diagonal (of permuted matrix)
    // factor the already permuted matrix Kp (lower triangle valid)
    void compute(const double* Kp, int n_, int variant)
    {
        n = n_;
        L.assign(Kp, Kp + (size_t)n * n);
        auto m = [&](int r, int c) -> double& { return L[r + (size_t)c * n]; };
        // right-looking, reciprocal multiply, non-fused ops
        dinv.assign(n, 0.0);
        for (int k = 0; k < n; ++k) {
            const double dk = m(k, k);
            const double r = (std::fabs(dk) > 0.0) ? 1.0 / dk : 1.0;
            dinv[k] = (std::fabs(dk) > DBL_MIN) ? 1.0 / dk : 0.0;
            for (int i = k + 1; i < n; ++i) m(i, k) = m(i, k) * r;
            for (int j = k + 1; j < n; ++j) {
                const double u = m(j, k) * dk;
                for (int i = j; i < n; ++i) m(i, j) = m(i, j) - m(i, k) * u;   // compiler may fuse (contract on)
            }
        }
        if (variant >= 2) {
            const int B = (variant == 2) ? 32 : n;   // block size of explicit inverses
            Linv.assign((size_t)n * n, 0.0);
            for (int b0 = 0; b0 < n; b0 += B) {
                const int b1 = std::min(n, b0 + B);
                // invert unit lower block [b0,b1)
                for (int j = b0; j < b1; ++j) {
                    Linv[j + (size_t)j * n] = 1.0;
                    for (int i = j + 1; i < b1; ++i) {
                        double acc = 0.0;
                        for (int k = j; k < i; ++k) acc += m(i, k) * Linv[k + (size_t)j * n];
                        Linv[i + (size_t)j * n] = -acc;
                    }
                }
            }
        }
    }
    void solve(double* y, int variant) const
    {
        auto m = [&](int r, int c) -> double { return L[r + (size_t)c * n]; };
        if (variant == 1) {
            for (int i = 0; i < n; ++i) { double a0 = 0, a1 = 0, a2 = 0, a3 = 0; int j = 0;
                for (; j + 4 <= i; j += 4) { a0 += m(i, j) * y[j]; a1 += m(i, j + 1) * y[j + 1]; a2 += m(i, j + 2) * y[j + 2]; a3 += m(i, j + 3) * y[j + 3]; }
                for (; j < i; ++j) a0 += m(i, j) * y[j];
                y[i] -= (a0 + a1) + (a2 + a3); }
            for (int i = 0; i < n; ++i) y[i] *= dinv[i];
            for (int i = n - 1; i >= 0; --i) { double a0 = 0, a1 = 0; int j = i + 1;
                for (; j + 2 <= n; j += 2) { a0 += m(j, i) * y[j]; a1 += m(j + 1, i) * y[j + 1]; }
                for (; j < n; ++j) a0 += m(j, i) * y[j];
                y[i] -= a0 + a1; }
            return;
        }
        const int B = (variant == 2) ? 32 : n;
        std::vector<double> t(n);
        // forward: block substitution with explicit diagonal-block inverses
        for (int b0 = 0; b0 < n; b0 += B) {
            const int b1 = std::min(n, b0 + B);
            for (int i = b0; i < b1; ++i) { double acc = 0; for (int j = b0; j <= i; ++j) acc += Linv[i + (size_t)j * n] * y[j]; t[i] = acc; }
            for (int i = b0; i < b1; ++i) y[i] = t[i];
            for (int i = b1; i < n; ++i) { double acc = 0; for (int j = b0; j < b1; ++j) acc += m(i, j) * y[j]; y[i] -= acc; }
        }
        for (int i = 0; i < n; ++i) y[i] *= dinv[i];
        for (int b1 = n; b1 > 0; ) {
            const int b0 = (variant == 2) ? ((b1 - 1) / 32) * 32 : 0;
            for (int i = b0; i < b1; ++i) { double acc = 0; for (int j = i; j < b1; ++j) acc += Linv[j + (size_t)i * n] * y[j]; t[i] = acc; }
            for (int i = b0; i < b1; ++i) y[i] = t[i];
            for (int i = 0; i < b0; ++i) { double acc = 0; for (int j = b0; j < b1; ++j) acc += m(j, i) * y[j]; y[i] -= acc; }
            b1 = b0;
        }
    }
};
}
