// oracle/admm_qp.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// ADMM<N, M, double, DENSE, Eigen::LDLT, Lower> restatement: reference src/solvers/admm.hpp (solve_impl 112-213, construct_A
// 215-222, construct_kkt_matrix 245-259, compute_kkt_rhs 389-393, box_projection 397-403, rho_vec_update 405-444,
// residuals_update 446-468, termination 470-486, estimate_rho 488-496, update_kkt_rho 498-502) — the OSQP-style splitting:
// the box constraints are appended to A as identity rows, Ae = [A; I] ((M + N) x N), one multiplier / auxiliary vector of
// size M + N, KKT system [[H + sigma I, Ae'], [Ae, -diag(1 / rho)]] of size 2N + M.  Differences from boxADMM that matter
// for bit parity: x = alpha x~ + (1 - alpha) x is ONE expression here (box_admm.hpp:129-130 reads the overwritten x),
// res_prim = max(|A x - z_A|, |x - z_box|) (a sum there), |z| runs over all M + N entries.
// Linear solver, inner products and norms as in qp.hpp / canon.hpp.
#pragma once
#include "qp.hpp"

namespace orc {

struct OsqpAdmm {
    int N = 0, M = 0, Me = 0;
    QpSettings settings;
    QpInfo info;
    std::vector<double> x, y;              // m_x (N), m_y (M + N) = [y_A ; y_box]
    std::vector<double> x_tilde, z, z_tilde, z_prev, rho_vec, rho_inv_vec;
    std::vector<int> constr_type, box_constr_type;
    std::vector<double> Ae;                // (M + N) x N column-major
    std::vector<double> K;                 // (2N + M)^2 column-major, lower part
    Ldlt ldlt;
    std::vector<int> first_perm;
    double rho = 0, max_Ax_z_norm = 0, max_Hx_ATy_h_norm = 0;
    int iter = 0, n_factor = 0;

    static constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e+6, RHO_EQ_FACTOR = 1e+3;
    static constexpr double LOOSE_BOUNDS_THRESH = 1e+10, EQ_TOL = 1e-4, DIV_BY_ZERO_REGUL = 10e-10;

    OsqpAdmm(int n, int m) : N(n), M(m), Me(n + m)
    {
        x.assign(N, 0); y.assign(Me, 0); x_tilde.assign(N, 0);
        z.assign(Me, 0); z_tilde.assign(Me, 0); z_prev.assign(Me, 0);
        rho_vec.assign(Me, settings.rho); rho_inv_vec.assign(Me, 1 / settings.rho);
        constr_type.assign(M, 0); box_constr_type.assign(N, 0);
        Ae.assign((size_t)Me * N, 0.0);
        K.assign((size_t)(N + Me) * (N + Me), 0.0);
    }

    static int classify(double lb, double ub)      // qp_base.hpp:195-222
    {
        if (lb < -LOOSE_BOUNDS_THRESH && ub > LOOSE_BOUNDS_THRESH) return LOOSE_BOUNDS;
        if (ub - lb < EQ_TOL) return EQUALITY_CONSTRAINT;
        return INEQUALITY_CONSTRAINT;
    }
    static double rho_of(int type, double rho0)
    { return type == LOOSE_BOUNDS ? RHO_MIN : (type == EQUALITY_CONSTRAINT ? RHO_EQ_FACTOR * rho0 : rho0); }

    void rho_vec_update(double rho0)               // admm.hpp:405-444
    {
        for (int i = 0; i < M; ++i) rho_vec[i] = rho_of(constr_type[i], rho0);
        for (int i = 0; i < N; ++i) rho_vec[M + i] = rho_of(box_constr_type[i], rho0);
        for (int i = 0; i < Me; ++i) rho_inv_vec[i] = 1.0 / rho_vec[i];
        rho = rho0;
        info.rho_updates += 1;
    }
    void construct_kkt_matrix(const double* H)     // admm.hpp:245-259
    {
        const int Kd = N + Me;
        for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) K[i + (size_t)j * Kd] = H[i + j * N];
        for (int i = 0; i < N; ++i) K[i + (size_t)i * Kd] += settings.sigma;
        for (int j = 0; j < N; ++j) for (int i = 0; i < Me; ++i) K[(N + i) + (size_t)j * Kd] = Ae[i + (size_t)j * Me];
        for (int i = 0; i < Me; ++i) K[(N + i) + (size_t)(N + i) * Kd] = -1.0 * rho_inv_vec[i];
    }
    void update_kkt_rho() { const int Kd = N + Me; for (int i = 0; i < Me; ++i) K[(N + i) + (size_t)(N + i) * Kd] = -rho_inv_vec[i]; }
    void factorise() { ldlt.compute(K.data(), N + Me); if (n_factor == 0) first_perm = ldlt.perm; ++n_factor; }

    void residuals_update(const double* H, const double* h)      // admm.hpp:446-468; A = the first M rows of Ae
    {
        std::vector<double> Ax(M > 0 ? M : 1), Hx(N), ATy(N), rp(M > 0 ? M : 1), rb(N), rd(N);
        for (int i = 0; i < M; ++i) Ax[i] = dot_seq(Ae.data() + i, Me, x.data(), 1, N);
        for (int i = 0; i < N; ++i) Hx[i] = dot_seq(H + i, N, x.data(), 1, N);
        for (int j = 0; j < N; ++j) ATy[j] = dot_seq(Ae.data() + (size_t)j * Me, 1, y.data(), 1, M);
        double norm_Ax = M > 0 ? norm_inf(Ax.data(), M) : 0.0;
        norm_Ax = fmax_(norm_Ax, norm_inf(x.data(), N));
        const double norm_z = norm_inf(z.data(), Me);
        max_Ax_z_norm = fmax_(norm_Ax, norm_z);
        const double norm_Hx = norm_inf(Hx.data(), N), norm_ATy = norm_inf(ATy.data(), N), norm_h = norm_inf(h, N),
                     norm_y_box = norm_inf(y.data() + M, N);
        max_Hx_ATy_h_norm = fmax_(norm_Hx, fmax_(norm_ATy, fmax_(norm_h, norm_y_box)));
        for (int i = 0; i < M; ++i) rp[i] = Ax[i] - z[i];
        for (int i = 0; i < N; ++i) rb[i] = x[i] - z[M + i];
        for (int i = 0; i < N; ++i) rd[i] = ((Hx[i] + h[i]) + ATy[i]) + y[M + i];
        info.res_prim = M > 0 ? norm_inf(rp.data(), M) : 0.0;
        info.res_prim = fmax_(info.res_prim, norm_inf(rb.data(), N));
        info.res_dual = norm_inf(rd.data(), N);
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

    /** solve_impl, 9-argument form (admm.hpp:112-213) */
    int solve(const double* H, const double* h, const double* A, const double* Alb, const double* Aub, const double* xlb,
              const double* xub, const double* x_guess, const double* y_guess)
    {
        const int Kd = N + Me;
        std::vector<double> rhs(Kd), sol(Kd);
        n_factor = 0;
        for (int i = 0; i < N; ++i) x[i] = x_guess ? x_guess[i] : 0.0;
        for (int i = 0; i < Me; ++i) y[i] = y_guess ? y_guess[i] : 0.0;
        // construct_A (215-222): Ae = [A; I]
        for (int j = 0; j < N; ++j) {
            for (int i = 0; i < M; ++i) Ae[i + (size_t)j * Me] = A[i + (size_t)j * M];
            for (int i = 0; i < N; ++i) Ae[(M + i) + (size_t)j * Me] = i == j ? 1.0 : 0.0;
        }
        for (int i = 0; i < Me; ++i) z[i] = dot_seq(Ae.data() + i, Me, x.data(), 1, N);       // m_z = m_A * x_guess
        for (int i = 0; i < M; ++i) constr_type[i] = classify(Alb[i], Aub[i]);
        for (int i = 0; i < N; ++i) box_constr_type[i] = classify(xlb[i], xub[i]);
        rho_vec_update(settings.rho);
        construct_kkt_matrix(H);
        factorise();
        info.status = QP_UNSOLVED;

        const double alpha = settings.alpha, sigma = settings.sigma;
        for (iter = 1; iter <= settings.max_iter; iter++) {
            z_prev = z;
            for (int i = 0; i < N; ++i) rhs[i] = sigma * x[i] - h[i];
            for (int i = 0; i < Me; ++i) rhs[N + i] = z[i] - rho_inv_vec[i] * y[i];
            ldlt.solve(rhs.data(), sol.data());
            for (int i = 0; i < N; ++i) x_tilde[i] = sol[i];
            for (int i = 0; i < Me; ++i) z_tilde[i] = z_prev[i] + rho_inv_vec[i] * (sol[N + i] - y[i]);
            for (int i = 0; i < N; ++i) x[i] = (alpha * x_tilde[i]) + ((1 - alpha) * x[i]);
            for (int i = 0; i < Me; ++i) {
                double v = alpha * z_tilde[i];
                v += ((1 - alpha) * z_prev[i]) + (rho_inv_vec[i] * y[i]);
                const double lb = i < M ? Alb[i] : xlb[i - M], ub = i < M ? Aub[i] : xub[i - M];
                z[i] = dm::min(dm::max(v, lb), ub);
            }
            for (int i = 0; i < Me; ++i) y[i] += rho_vec[i] * (((alpha * z_tilde[i]) + ((1 - alpha) * z_prev[i])) - z[i]);

            const bool check_termination = (settings.check_termination != 0 && iter % settings.check_termination == 0);
            if (check_termination) {
                residuals_update(H, h);
                if (termination_criteria()) { info.status = QP_SOLVED; break; }
            }
            if (settings.adaptive_rho && iter % settings.adaptive_rho_interval == 0) {
                if (!check_termination) residuals_update(H, h);
                double new_rho = estimate_rho(rho);
                new_rho = fmax_(RHO_MIN, fmin_(new_rho, RHO_MAX));
                info.rho_estimate = new_rho;
                if (new_rho < rho / settings.adaptive_rho_tolerance || new_rho > rho * settings.adaptive_rho_tolerance) {
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
