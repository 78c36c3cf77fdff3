// oracle/qp.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// boxADMM restatement: reference src/solvers/box_admm.hpp (solve_impl 88-205, KKT 207-223, rhs 351-355,
// rho_vec_update 357-396, residuals 398-415, estimate_rho 433-445, update_kkt_rho 447-452) and
// src/solvers/qp_base.hpp (settings 17-53, status 55-62, info 64-72, bound classification 195-222).
// The dense linear solver is Eigen::LDLT<K, Lower> (reference src/utils/helpers.hpp:38-43); Eigen is not vendored, so
// its algorithm is restated from Eigen 3.3.7 Eigen/src/Cholesky/LDLT.h  [Eigen-ext]:
//   * left-looking in-place LDL^T; at step k the pivot is the largest |diagonal| of the *trailing, not yet updated*
//     diagonal (first maximum wins), applied as a symmetric transposition on the lower triangle;
//   * column scaling A21 /= d_k only if |d_k| > 0;
//   * solve: P, unit-lower solve, x_i /= d_i if |d_i| > DBL_MIN else 0, unit-upper solve, P^T.
// Inner products follow canon.hpp (sequential ascending fused chains).
#pragma once
#include "canon.hpp"
#include <vector>
#include <cfloat>
#include "ldlt_variants.hpp"

namespace orc {

struct QpSettings {          // qp_base.hpp:17-53 (ADMM related subset)
    double eps_rel = 1e-3, eps_abs = 1e-3;
    int max_iter = 1000;
    int warm_start = 0, reuse_pattern = 0, verbose = 0;
    double rho = 1e-1, sigma = 1e-6, alpha = 1.0;
    int check_termination = 25;
    int adaptive_rho = 0;
    double adaptive_rho_tolerance = 5;
    int adaptive_rho_interval = 25;
};

enum QpStatus { QP_SOLVED = 0, QP_MAX_ITER_EXCEEDED = 1, QP_UNSOLVED = 2, QP_UNINITIALIZED = 3, QP_INFEASIBLE = 4, QP_INCONSISTENT = 5 };
enum CType { INEQUALITY_CONSTRAINT = 0, EQUALITY_CONSTRAINT = 1, LOOSE_BOUNDS = 2 };  // qp_base.hpp:132-136

struct QpInfo {              // qp_base.hpp:64-72
    int status = QP_UNINITIALIZED;
    int iter = 0;
    int rho_updates = 0;
    double rho_estimate = 0;
    double res_prim = 1;
    double res_dual = 1;
};

/** Eigen::LDLT<Matrix, Lower> restatement on a dense column-major K (only the lower triangle is read). */
struct Ldlt {
    int n = 0;
    std::vector<double> L;     // n x n column-major, unit lower factor below the diagonal, D on the diagonal
    std::vector<int> transp;   // transpositions
    std::vector<int> perm;     // composed permutation: (P v)[a] = v[perm[a]]
    std::vector<double> temp;

    void compute(const double* K, int n_)
    {
        n = n_;
        L.assign(K, K + (size_t)n * n);
        transp.assign(n, 0); perm.resize(n); temp.assign(n, 0.0);
        auto m = [&](int r, int c) -> double& { return L[r + (size_t)c * n]; };
        for (int k = 0; k < n; ++k) {
            // pivot: first maximum of |diag| over the trailing part (not yet updated: left-looking)
            int big = k; double best = dm::fabs(m(k, k));
            for (int i = k + 1; i < n; ++i) { const double v = dm::fabs(m(i, i)); if (v > best) { best = v; big = i; } }
            transp[k] = big;
            if (big != k) {
                const int s = n - big - 1;
                for (int j = 0; j < k; ++j) { const double t = m(k, j); m(k, j) = m(big, j); m(big, j) = t; }
                for (int i = 0; i < s; ++i) { const double t = m(big + 1 + i, k); m(big + 1 + i, k) = m(big + 1 + i, big); m(big + 1 + i, big) = t; }
                { const double t = m(k, k); m(k, k) = m(big, big); m(big, big) = t; }
                for (int i = k + 1; i < big; ++i) { const double t = m(i, k); m(i, k) = m(big, i); m(big, i) = t; }
            }
            if (k > 0) {
                for (int j = 0; j < k; ++j) temp[j] = m(j, j) * m(k, j);
                double acc = m(k, k);
                for (int j = 0; j < k; ++j) acc = dm::fma(-m(k, j), temp[j], acc);
                m(k, k) = acc;
                for (int i = k + 1; i < n; ++i) {
                    double v = m(i, k);
                    for (int j = 0; j < k; ++j) v = dm::fma(-m(i, j), temp[j], v);
                    m(i, k) = v;
                }
            }
            const double dk = m(k, k);
            if (dm::fabs(dk) > 0.0)
                for (int i = k + 1; i < n; ++i) m(i, k) = m(i, k) / dk;
        }
        for (int i = 0; i < n; ++i) perm[i] = i;
        for (int k = 0; k < n; ++k) { const int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }
        if (ldlt_variant() > 0) {
            std::vector<double> Kp((size_t)n * n, 0.0);
            for (int b = 0; b < n; ++b) for (int a = b; a < n; ++a) {
                const int r = perm[a], c = perm[b];
                Kp[a + (size_t)b * n] = r >= c ? K[r + (size_t)c * n] : K[c + (size_t)r * n];
            }
            fast.compute(Kp.data(), n, ldlt_variant());
        }
    }
    LdltFast fast;

    void solve(const double* rhs, double* x) const
    {
        std::vector<double> y(n);
        auto m = [&](int r, int c) -> double { return L[r + (size_t)c * n]; };
        for (int a = 0; a < n; ++a) y[a] = rhs[perm[a]];
        if (ldlt_variant() > 0) { fast.solve(y.data(), ldlt_variant()); for (int a = 0; a < n; ++a) x[perm[a]] = y[a]; return; }
        for (int i = 0; i < n; ++i) {            // unit lower: ascending columns
            double acc = y[i];
            for (int j = 0; j < i; ++j) acc = dm::fma(-m(i, j), y[j], acc);
            y[i] = acc;
        }
        for (int i = 0; i < n; ++i) {
            const double di = m(i, i);
            if (dm::fabs(di) > DBL_MIN) y[i] = y[i] / di; else y[i] = 0.0;
        }
        for (int i = n - 1; i >= 0; --i) {       // unit upper (L^T): descending columns
            double acc = y[i];
            for (int j = n - 1; j > i; --j) acc = dm::fma(-m(j, i), y[j], acc);
            y[i] = acc;
        }
        for (int a = 0; a < n; ++a) x[perm[a]] = y[a];
    }
};

/** boxADMM<N, M, double, DENSE, Eigen::LDLT, Lower> */
struct BoxAdmm {
    int N = 0, M = 0;
    QpSettings settings;
    QpInfo info;
    std::vector<double> x, y;              // m_x (N), m_y (M+N) = [y_A ; y_box]
    std::vector<double> x_tilde, q, z, z_tilde, z_prev;
    std::vector<double> rho_vec, rho_inv_vec, rho_box, rho_box_inv, rho_box_prev;
    std::vector<int> constr_type, box_constr_type;
    std::vector<double> K;                 // (N+M)^2 column-major; upper-right block never written (stays 0)
    Ldlt ldlt;
    std::vector<int> first_perm;           // permutation of the first factorisation of the last solve (decision trace)
    double rho = 0, max_Ax_z_norm = 0, max_Hx_ATy_h_norm = 0;
    int iter = 0, n_factor = 0;

    static constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e+6, RHO_EQ_FACTOR = 1e+3;
    static constexpr double LOOSE_BOUNDS_THRESH = 1e+10, EQ_TOL = 1e-4, DIV_BY_ZERO_REGUL = 10e-10;

    BoxAdmm(int n, int m) : N(n), M(m)
    {
        x.assign(N, 0); y.assign(N + M, 0); x_tilde.assign(N, 0); q.assign(N, 0);
        z.assign(M, 0); z_tilde.assign(M, 0); z_prev.assign(M, 0);
        rho_vec.assign(M, settings.rho); rho_inv_vec.assign(M, 1 / settings.rho);
        rho_box.assign(N, 0); rho_box_inv.assign(N, 0); rho_box_prev.assign(N, 0);
        constr_type.assign(M, 0); box_constr_type.assign(N, 0);
        K.assign((size_t)(N + M) * (N + M), 0.0);
    }

    void parse_constraints_bounds(const double* Alb, const double* Aub, const double* xlb, const double* xub)
    {
        for (int i = 0; i < M; ++i) {
            if (Alb[i] < -LOOSE_BOUNDS_THRESH && Aub[i] > LOOSE_BOUNDS_THRESH) constr_type[i] = LOOSE_BOUNDS;
            else if (Aub[i] - Alb[i] < EQ_TOL) constr_type[i] = EQUALITY_CONSTRAINT;
            else constr_type[i] = INEQUALITY_CONSTRAINT;
        }
        for (int i = 0; i < N; ++i) {
            if (xlb[i] < -LOOSE_BOUNDS_THRESH && xub[i] > LOOSE_BOUNDS_THRESH) box_constr_type[i] = LOOSE_BOUNDS;
            else if (xub[i] - xlb[i] < EQ_TOL) box_constr_type[i] = EQUALITY_CONSTRAINT;
            else box_constr_type[i] = INEQUALITY_CONSTRAINT;
        }
    }

    void rho_vec_update(double rho0)
    {
        for (int i = 0; i < M; ++i) {
            switch (constr_type[i]) {
            case LOOSE_BOUNDS: rho_vec[i] = RHO_MIN; break;
            case EQUALITY_CONSTRAINT: rho_vec[i] = RHO_EQ_FACTOR * rho0; break;
            default: rho_vec[i] = rho0;
            }
            rho_inv_vec[i] = 1.0 / rho_vec[i];
        }
        rho = rho0;
        for (int i = 0; i < N; ++i) {
            switch (box_constr_type[i]) {
            case LOOSE_BOUNDS: rho_box[i] = RHO_MIN; break;
            case EQUALITY_CONSTRAINT: rho_box[i] = RHO_EQ_FACTOR * rho0; break;
            default: rho_box[i] = rho0;
            }
            rho_box_inv[i] = 1.0 / rho_box[i];
        }
        info.rho_updates += 1;
    }

    void construct_kkt_matrix(const double* H, const double* A)
    {
        const int Kd = N + M;
        for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) K[i + (size_t)j * Kd] = H[i + j * N];
        for (int i = 0; i < N; ++i) K[i + (size_t)i * Kd] += settings.sigma;
        for (int i = 0; i < N; ++i) K[i + (size_t)i * Kd] += rho_box[i];
        for (int j = 0; j < N; ++j) for (int i = 0; i < M; ++i) K[(N + i) + (size_t)j * Kd] = A[i + j * M];
        for (int i = 0; i < M; ++i) K[(N + i) + (size_t)(N + i) * Kd] = -rho_inv_vec[i];
    }
    void update_kkt_rho()
    {
        const int Kd = N + M;
        for (int i = 0; i < N; ++i) K[i + (size_t)i * Kd] += (rho_box[i] - rho_box_prev[i]);
        for (int i = 0; i < M; ++i) K[(N + i) + (size_t)(N + i) * Kd] = -rho_inv_vec[i];
    }
    void factorise() { ldlt.compute(K.data(), N + M); if (n_factor == 0) first_perm = ldlt.perm; ++n_factor; }

    // canonical mat-vecs: sequential ascending fused chains
    void mv_A(const double* A, const double* v, double* out) const   // A v
    { for (int i = 0; i < M; ++i) out[i] = dot_seq(A + i, M, v, 1, N); }
    void mv_AT(const double* A, const double* v, double* out) const  // A^T v
    { for (int j = 0; j < N; ++j) out[j] = dot_seq(A + (size_t)j * M, 1, v, 1, M); }
    void mv_H(const double* H, const double* v, double* out) const   // H v (full matrix, row i)
    { for (int i = 0; i < N; ++i) out[i] = dot_seq(H + i, N, v, 1, N); }

    void residuals_update(const double* H, const double* h, const double* A)
    {
        std::vector<double> Ax(M), Hx(N), ATy(N);
        mv_A(A, x.data(), Ax.data());
        mv_H(H, x.data(), Hx.data());
        mv_AT(A, y.data(), ATy.data());
        const double norm_Ax = norm_inf(Ax.data(), M), norm_z = norm_inf(z.data(), M);
        max_Ax_z_norm = fmax_(norm_Ax, fmax_(norm_z, norm_inf(x.data(), N)));
        const double norm_Hx = norm_inf(Hx.data(), N), norm_ATy = norm_inf(ATy.data(), N), norm_h = norm_inf(h, N),
                     norm_y_box = norm_inf(y.data() + M, N);
        max_Hx_ATy_h_norm = fmax_(norm_Hx, fmax_(norm_ATy, fmax_(norm_h, norm_y_box)));
        // primal_residual / dual_residual (qp_base.hpp) are lpNorm<Infinity>() of the residual vectors
        std::vector<double> rpv(M), rqv(N), rdv(N);
        for (int i = 0; i < M; ++i) rpv[i] = Ax[i] - z[i];
        for (int i = 0; i < N; ++i) rqv[i] = x[i] - q[i];
        for (int i = 0; i < N; ++i) rdv[i] = ((Hx[i] + h[i]) + ATy[i]) + y[M + i];
        info.res_prim = norm_inf(rpv.data(), M) + norm_inf(rqv.data(), N);
        info.res_dual = norm_inf(rdv.data(), N);
    }
    bool termination_criteria() const
    {
        const double eps_prim = settings.eps_abs + settings.eps_rel * max_Ax_z_norm;
        const double eps_dual = settings.eps_abs + settings.eps_rel * max_Hx_ATy_h_norm;
        return info.res_prim <= eps_prim && info.res_dual <= eps_dual;
    }
    double estimate_rho(double rho0) const
    {
        const double rp_norm = info.res_prim / (max_Ax_z_norm + DIV_BY_ZERO_REGUL);
        const double rd_norm = info.res_dual / (max_Hx_ATy_h_norm + DIV_BY_ZERO_REGUL);
        return rho0 * dm::sqrt(rp_norm / (rd_norm + DIV_BY_ZERO_REGUL));
    }

    /** solve_impl, 9-argument form (box_admm.hpp:88-205) */
    int solve(const double* H, const double* h, const double* A, const double* Alb, const double* Aub, const double* xlb,
              const double* xub, const double* x_guess, const double* y_guess)
    {
        const int Kd = N + M;
        std::vector<double> rhs(Kd), sol(Kd);
        bool check_termination = false;
        n_factor = 0;
        for (int i = 0; i < N; ++i) x[i] = x_guess ? x_guess[i] : 0.0;
        for (int i = 0; i < N + M; ++i) y[i] = y_guess ? y_guess[i] : 0.0;
        mv_A(A, x.data(), z.data());
        q = x;

        parse_constraints_bounds(Alb, Aub, xlb, xub);
        rho_vec_update(settings.rho);
        construct_kkt_matrix(H, A);
        factorise();
        info.status = QP_UNSOLVED;

        const double alpha = settings.alpha, sigma = settings.sigma;
        for (iter = 1; iter <= settings.max_iter; iter++) {
            z_prev = z;
            // compute_kkt_rhs (351-355)
            for (int i = 0; i < N; ++i) rhs[i] = ((sigma * x[i] - h[i]) + rho_box[i] * q[i]) - y[M + i];
            for (int i = 0; i < M; ++i) rhs[N + i] = z[i] - rho_inv_vec[i] * y[i];
            ldlt.solve(rhs.data(), sol.data());
            for (int i = 0; i < N; ++i) x_tilde[i] = sol[i];
            for (int i = 0; i < M; ++i) z_tilde[i] = z_prev[i] + rho_inv_vec[i] * (sol[N + i] - y[i]);
            // update x (129-130): note the second statement reads the already overwritten m_x
            for (int i = 0; i < N; ++i) { x[i] = alpha * x_tilde[i]; x[i] += (1 - alpha) * x[i]; }
            // update z (133-135)
            for (int i = 0; i < M; ++i) {
                double v = alpha * z_tilde[i];
                v += ((1 - alpha) * z_prev[i]) + (rho_inv_vec[i] * y[i]);
                z[i] = dm::min(dm::max(v, Alb[i]), Aub[i]);
            }
            // update q (138-139)
            for (int i = 0; i < N; ++i) {
                const double v = x[i] + rho_box_inv[i] * y[M + i];
                q[i] = dm::min(dm::max(v, xlb[i]), xub[i]);
            }
            // dual updates (142-147)
            for (int i = 0; i < M; ++i) y[i] += rho_vec[i] * (((alpha * z_tilde[i]) + ((1 - alpha) * z_prev[i])) - z[i]);
            for (int i = 0; i < N; ++i) y[M + i] += rho_box[i] * (x[i] - q[i]);

            check_termination = (settings.check_termination != 0 && iter % settings.check_termination == 0);
            if (check_termination) {
                residuals_update(H, h, A);
                if (termination_criteria()) { info.status = QP_SOLVED; break; }
            }
            if (settings.adaptive_rho && iter % settings.adaptive_rho_interval == 0) {
                if (!check_termination) residuals_update(H, h, A);
                double new_rho = estimate_rho(rho);
                new_rho = fmax_(RHO_MIN, fmin_(new_rho, RHO_MAX));
                info.rho_estimate = new_rho;
                if (new_rho < rho / settings.adaptive_rho_tolerance || new_rho > rho * settings.adaptive_rho_tolerance) {
                    rho_box_prev = rho_box;
                    rho_vec_update(new_rho);
                    update_kkt_rho();
                    factorise();
                }
            }
        }
        if (iter > settings.max_iter) info.status = QP_MAX_ITER_EXCEEDED;
        info.iter = iter;
        return info.status;
    }
};

} // namespace orc
