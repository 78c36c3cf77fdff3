// oracle/sqp.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// SQP restatement: reference src/solvers/sqp_base.hpp (settings 24-47, ctor QP defaults 83-90, line search 378-419,
// constraint violation 421-474, BFGS linearisation update 489-504, termination 523-529, solve_qp 532-565, solve 568-696)
// and src/solvers/bfgs.hpp:23-52 (damped BFGS).
#pragma once
#include "ocp.hpp"
#include "qp.hpp"
#include "precond.hpp"
#include "admm_qp.hpp"
#include <limits>

namespace orc {

struct SqpSettings {  // sqp_base.hpp:24-34
    double tau = 0.5, eta = 0.25, rho = 0.5, eps_prim = 1e-3, eps_dual = 1e-3;
    int max_iter = 100, line_search_max_iter = 100;
};
enum SqpStatus { SQP_SOLVED = 0, SQP_MAX_ITER_EXCEEDED = 1, SQP_INVALID_SETTINGS = 2 };
struct SqpInfo { int iter = 0, qp_solver_iter = 0, status = SQP_MAX_ITER_EXCEEDED; };

/** bfgs.hpp:23-52. B is n x n column-major. Returns the branch taken: 0 plain, 1 damped, 2 skipped. */
inline int bfgs_update(double* B, const double* s, const double* y, int n)
{
    std::vector<double> Bs(n), r(n);
    for (int i = 0; i < n; ++i) Bs[i] = dot_seq(B + i, n, s, 1, n);
    const double sBs = dot_tree32(s, Bs.data(), n);
    const double sy = dot_tree32(s, y, n);
    double sr;
    int branch;
    if (sy < 0.2 * sBs) {
        const double theta = 0.8 * sBs / (sBs - sy);
        for (int i = 0; i < n; ++i) r[i] = theta * y[i] + (1 - theta) * Bs[i];
        sr = theta * sy + (1 - theta) * sBs;
        branch = 1;
    } else {
        for (int i = 0; i < n; ++i) r[i] = y[i];
        sr = sy;
        branch = 0;
    }
    if (sr < std::numeric_limits<double>::epsilon()) return 2;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double b = B[i + j * n];
            b += ((-Bs[i]) * Bs[j]) / sBs;
            b += (r[i] * r[j]) / sr;
            B[i + j * n] = b;
        }
    return branch;
}

/** ContinuousOCP<..., SPARSE>::hessian_update_impl (continuous_ocp.hpp:2303-2431): the "sparsity preserving block BFGS" every
 *  reference control test installs through `this->problem.hessian_update_impl(hessian, x_step, grad_step)`.  Only the
 *  per-node (x_k,u_k) diagonal blocks and the parameter rows / columns of H are stored (and updated); everything else of H
 *  is structurally zero.  Differences from bfgs.hpp: the product v = H s runs over the stored pattern, the update is never
 *  skipped, the coefficients are reciprocals (-1/s'v, 1/s'y or 1/s'r) multiplied into the FIRST factor of every outer
 *  product, and the increments of one block are summed before they are added to H.
 *  Canonical orders: v_i = fused chain over the stored columns of row i in ascending column index; dots are tree32.
 *  Returns the branch: 0 plain, 1 damped. */
template <class OcpT>
inline int block_bfgs_update(double* H, const double* s, const double* y)
{
    constexpr int NX = OcpT::NX, NU = OcpT::NU, NP = OcpT::NP, NN = OcpT::NN, VARX = OcpT::VARX, VARU = OcpT::VARU, N = OcpT::N;
    auto Hm = [&](int r, int c) -> double& { return H[r + (size_t)c * N]; };
    std::vector<double> v(N), r(N);
    for (int k = 0; k < NN; ++k) {
        for (int a = 0; a < NX + NU; ++a) {
            const int row = a < NX ? k * NX + a : VARX + k * NU + (a - NX);
            double acc = 0.0;
            for (int j = 0; j < NX; ++j) acc = dm::fma(Hm(row, k * NX + j), s[k * NX + j], acc);
            for (int j = 0; j < NU; ++j) acc = dm::fma(Hm(row, VARX + k * NU + j), s[VARX + k * NU + j], acc);
            for (int j = 0; j < NP; ++j) acc = dm::fma(Hm(row, VARX + VARU + j), s[VARX + VARU + j], acc);
            v[row] = acc;
        }
    }
    for (int a = 0; a < NP; ++a) v[VARX + VARU + a] = dot_seq(H + VARX + VARU + a, N, s, 1, N);
    const double scaling = dot_tree32(s, v.data(), N);
    const double scaling_inv = 1.0 / scaling;
    const double sy = dot_tree32(s, y, N);
    const double sy_inv = 1.0 / sy;
    const bool plain = sy >= 0.2 * scaling;
    double c2 = sy_inv;
    const double* w = y;                        // second rank-one term: c2 * w w'
    if (!plain) {
        const double theta = 0.8 * scaling / (scaling - sy);
        for (int i = 0; i < N; ++i) r[i] = theta * y[i] + (1 - theta) * v[i];
        c2 = 1.0 / dot_tree32(s, r.data(), N);
        w = r.data();
    }
    const double c1 = -scaling_inv;
    auto inc = [&](int i, int j) { double h = (c1 * v[i]) * v[j]; h += (c2 * w[i]) * w[j]; return h; };   // (i, j) of a block
    for (int k = 0; k < NN; ++k) {
        const int x0 = k * NX, u0 = VARX + k * NU;
        for (int j = 0; j < NX; ++j) for (int i = 0; i < NX; ++i) Hm(x0 + i, x0 + j) += inc(x0 + i, x0 + j);   // hes_xx
        for (int j = 0; j < NU; ++j) for (int i = 0; i < NU; ++i) Hm(u0 + i, u0 + j) += inc(u0 + i, u0 + j);   // hes_uu
        for (int j = 0; j < NX; ++j) for (int i = 0; i < NU; ++i) {                                            // hes_ux and its transpose
            const double h = inc(u0 + i, x0 + j);
            Hm(u0 + i, x0 + j) += h;
            Hm(x0 + j, u0 + i) += h;
        }
    }
    if (NP > 0) {
        const int p0 = VARX + VARU;
        for (int j = 0; j < NP; ++j) for (int i = 0; i < NP; ++i) Hm(p0 + i, p0 + j) += inc(p0 + i, p0 + j);   // hes_pp
        for (int j = 0; j < NP; ++j) for (int i = 0; i < p0; ++i) {                                            // hes_ap: column j and row j
            const double h = inc(i, p0 + j);
            Hm(i, p0 + j) += h;
            Hm(p0 + j, i) += h;
        }
    }
    return plain ? 0 : 1;
}

/** SQPBase<Derived, ContinuousOCP<...,DENSE>, boxADMM<...>, IdentityPreconditioner> with all default hooks */
template <class OcpT>
struct Sqp {
    static constexpr int N = OcpT::N, M = OcpT::M, NUM_EQ = OcpT::NUM_EQ, NUM_INEQ = OcpT::NUM_INEQ, DUAL = OcpT::DUAL, ND = OcpT::ND;
    OcpT problem;
    SqpSettings settings;
    SqpInfo info;
    BoxAdmm qp;
    std::vector<double> H, h, x, lam, lam_k, A, al, au, p_static, lbx, ubx, lx, ux, lbg, ubg, lag_gradient, step_prev;
    double cost_val = 0, primal_norm = 0, dual_norm = 0, max_violation = 0;
    // decision trace (per SQP iteration)
    std::vector<int> tr_qp_iter, tr_bfgs, tr_ls_trials, tr_qp_factor;
    std::vector<double> tr_alpha;

    Sqp() : qp(N, M)
    {
        const double INF = std::numeric_limits<double>::infinity();
        H.assign((size_t)N * N, 0); h.assign(N, 0); x.assign(N, 0); lam.assign(DUAL, 0); lam_k.assign(DUAL, 0);
        A.assign((size_t)M * N, 0); al.assign(M, 0); au.assign(M, 0); p_static.assign(ND > 0 ? ND : 1, 0);
        lbx.assign(N, -INF); ubx.assign(N, INF); lx.assign(N, 0); ux.assign(N, 0);
        lbg.assign(NUM_INEQ, -INF); ubg.assign(NUM_INEQ, INF);
        lag_gradient.assign(N, 0); step_prev.assign(N, 0);
        // sqp_base.hpp:83-90
        qp.settings.warm_start = 0;
        qp.settings.check_termination = 10;
        qp.settings.eps_abs = 1e-4;
        qp.settings.eps_rel = 1e-4;
        qp.settings.max_iter = 100;
        qp.settings.adaptive_rho = 1;
        qp.settings.adaptive_rho_interval = 50;
        qp.settings.alpha = 1.0;
    }

    /** sqp_base.hpp:421-444 */
    double constraints_violation(const double* xv) const
    {
        double cl1 = std::numeric_limits<double>::epsilon();
        std::vector<double> c(NUM_EQ), g(NUM_INEQ > 0 ? NUM_INEQ : 1);
        problem.equalities(xv, p_static.data(), c.data());
        cl1 += sum_tree32(NUM_EQ, [&](int i) { return dm::fabs(c[i]); });
        problem.inequalities(xv, p_static.data(), g.data());
        cl1 += sum_tree32(NUM_INEQ, [&](int i) { return dm::max(lbg[i] - g[i], 0.0); });
        cl1 += sum_tree32(NUM_INEQ, [&](int i) { return dm::max(g[i] - ubg[i], 0.0); });
        cl1 += sum_tree32(N, [&](int i) { return dm::max(lbx[i] - xv[i], 0.0); });
        cl1 += sum_tree32(N, [&](int i) { return dm::max(xv[i] - ubx[i], 0.0); });
        return cl1;
    }
    /** sqp_base.hpp:446-474 */
    double max_constraints_violation(const double* xv) const
    {
        double c = 0.0;
        if (NUM_EQ > 0) {
            std::vector<double> ce(NUM_EQ);
            problem.equalities(xv, p_static.data(), ce.data());
            c = norm_inf(ce.data(), NUM_EQ);
        }
        if (NUM_INEQ > 0) {
            std::vector<double> g(NUM_INEQ);
            problem.inequalities(xv, p_static.data(), g.data());
            double m1 = lbg[0] - g[0], m2 = g[0] - ubg[0];
            for (int i = 1; i < NUM_INEQ; ++i) { if (lbg[i] - g[i] > m1) m1 = lbg[i] - g[i]; if (g[i] - ubg[i] > m2) m2 = g[i] - ubg[i]; }
            c = fmax_(c, m1); c = fmax_(c, m2);
        }
        double m1 = lbx[0] - xv[0], m2 = xv[0] - ubx[0];
        for (int i = 1; i < N; ++i) { if (lbx[i] - xv[i] > m1) m1 = lbx[i] - xv[i]; if (xv[i] - ubx[i] > m2) m2 = xv[i] - ubx[i]; }
        c = fmax_(c, m1); c = fmax_(c, m2);
        return c;
    }

    /** sqp_base.hpp:378-419 */
    double step_size_selection(const double* p, int& trials)
    {
        const double tau = settings.tau;
        const double constr_l1 = constraints_violation(x.data());
        const double mu = norm_inf(lam_k.data(), DUAL);
        double cost_1;
        problem.cost(x.data(), p_static.data(), cost_1);
        const double phi_l1 = cost_1 + mu * constr_l1;
        const double Dp_phi_l1 = dot_tree32(h.data(), p, N) - mu * constr_l1;
        double alpha = 1.0, cost_step;
        std::vector<double> x_step(N);
        trials = 0;
        for (int i = 1; i < settings.line_search_max_iter; i++) {
            for (int j = 0; j < N; ++j) { x_step[j] = alpha * p[j]; x_step[j] += x[j]; }
            problem.cost(x_step.data(), p_static.data(), cost_step);
            cost_val = cost_step;
            ++trials;
            const double phi_l1_step = cost_step + mu * constraints_violation(x_step.data());
            if (phi_l1_step <= (phi_l1 + alpha * settings.eta * Dp_phi_l1)) return alpha;
            else alpha = tau * alpha;
        }
        return alpha;
    }

    /** tests/control/valet_parking_mpc_test.cpp:110-155: step_size_selection_impl with LSFilter (line_search.hpp:30-98) */
    double step_size_selection_filter(const double* p, int& trials)
    {
        const double tau = settings.tau;
        const double constr_l1 = constraints_violation(x.data());
        double cost_1;
        problem.cost(x.data(), p_static.data(), cost_1);
        if (filter.is_acceptable(cost_1, constr_l1)) filter.add(cost_1, constr_l1);
        double alpha = 1.0, cost_step;
        std::vector<double> x_step(N);
        trials = 0;
        for (int i = 1; i < settings.line_search_max_iter; i++) {
            for (int j = 0; j < N; ++j) { x_step[j] = alpha * p[j]; x_step[j] += x[j]; }
            problem.cost(x_step.data(), p_static.data(), cost_step);
            const double constr_step = constraints_violation(x_step.data());
            cost_val = cost_step;
            ++trials;
            if (filter.is_acceptable(cost_step, constr_step)) { filter.add(cost_step, constr_step); return alpha; }
            else alpha *= tau;
        }
        return alpha;
    }

    // The fixed menu of SQPBase CRTP overrides the engine can honour (pmb_sqp_set_hessian_options): both are what the
    // reference's own solvers install, tests/control/minimal_time_test.cpp:90-135
    int opt_exact_hessian = 0;   // update_linearisation_dense_impl := linearisation_dense_impl (exact Hessian at every iteration)
    int opt_gershgorin = 0;      // hessian_regularisation_dense_impl := Gershgorin shift of the diagonal
    int opt_block_bfgs = 0;      // hessian_update_impl := the OCP's block BFGS (ContinuousOCP<..., SPARSE>::hessian_update_impl)
    int opt_qp_solver = 0;       // QPSolver template argument: 0 boxADMM<>, 1 ADMM<> (the OSQP-style splitting, admm.hpp)
    OsqpAdmm qp_admm{N, M};
    int opt_precond = PRECOND_IDENTITY;   // Preconditioner template argument: RuizEquilibration<..., DENSE | SPARSE> (sqp_base.hpp:605-611, 662-667)
    int opt_line_search = 0;     // 0: l1 merit (default), 1: the filter line search of tests/control/valet_parking_mpc_test.cpp:110-155
    LsFilter filter;             // the solver member `filter` of that test: it lives as long as the solver object
    Ruiz ruiz{N, M, PRECOND_RUIZ_DENSE};

    /** minimal_time_test.cpp:90-104; called from linearisation_dense_impl (sqp_base.hpp:316-317).  cwiseAbs().sum() of a
     *  column is taken in sequential ascending order ("parity unpinned": Eigen's order depends on the vector ISA). */
    void regularise_hessian()
    {
        if (!opt_gershgorin) return;
        for (int i = 0; i < N; ++i) {
            const double aii = H[i + (size_t)i * N];
            double sum = 0.0;
            for (int j = 0; j < N; ++j) sum += dm::fabs(H[j + (size_t)i * N]);
            const double ri = sum - dm::fabs(aii);
            if (aii - ri <= 0) H[i + (size_t)i * N] += (ri - aii) + 0.01;
        }
    }

    void prepare_qp_bounds()  // sqp_base.hpp:588-593
    {
        for (int i = 0; i < M; ++i) { al[i] = -al[i]; au[i] = al[i]; }
        for (int i = 0; i < NUM_INEQ; ++i) { al[NUM_EQ + i] += lbg[i]; au[NUM_EQ + i] += ubg[i]; }
        for (int i = 0; i < N; ++i) { lx[i] = lbx[i] - x[i]; ux[i] = ubx[i] - x[i]; }
    }

    bool iterate_tail(std::vector<double>& p, std::vector<double>& p_lambda)
    {
        // m_preconditioner.compute (605 / 662), solve_qp (532-565, status ignored), unscale of the solution and of the data (609-611)
        if (opt_precond != PRECOND_IDENTITY) {
            ruiz.variant = opt_precond;
            ruiz.compute(H.data(), h.data(), A.data(), al.data(), au.data(), lx.data(), ux.data());
        }
        int qp_iter, qp_nfac;
        if (opt_qp_solver == 1) {
            qp_admm.settings = qp.settings;
            qp_admm.solve(H.data(), h.data(), A.data(), al.data(), au.data(), lx.data(), ux.data(), nullptr, nullptr);
            p = qp_admm.x; p_lambda = qp_admm.y; qp_iter = qp_admm.info.iter; qp_nfac = qp_admm.n_factor;
        } else {
            qp.solve(H.data(), h.data(), A.data(), al.data(), au.data(), lx.data(), ux.data(), nullptr, nullptr);
            p = qp.x; p_lambda = qp.y; qp_iter = qp.info.iter; qp_nfac = qp.n_factor;
        }
        info.qp_solver_iter += qp_iter;
        if (opt_precond != PRECOND_IDENTITY) {
            ruiz.unscale_solution(p.data(), p_lambda.data());
            ruiz.unscale_data(H.data(), h.data(), A.data(), al.data(), au.data(), lx.data(), ux.data());
        }
        tr_qp_iter.push_back(qp_iter); tr_qp_factor.push_back(qp_nfac);
        lam_k = p_lambda;
        for (int i = 0; i < DUAL; ++i) p_lambda[i] -= lam[i];
        int trials = 0;
        const double alpha = opt_line_search == 1 ? step_size_selection_filter(p.data(), trials) : step_size_selection(p.data(), trials);
        tr_alpha.push_back(alpha); tr_ls_trials.push_back(trials);
        for (int i = 0; i < N; ++i) x[i] += alpha * p[i];
        for (int i = 0; i < DUAL; ++i) lam[i] += alpha * p_lambda[i];
        for (int i = 0; i < N; ++i) step_prev[i] = alpha * p[i];
        primal_norm = alpha * norm_inf(p.data(), N);
        dual_norm = alpha * norm_inf(p_lambda.data(), DUAL);
        // termination_criteria (523-529)
        max_violation = max_constraints_violation(x.data());
        return (primal_norm <= settings.eps_prim) && (dual_norm <= settings.eps_dual) && (max_violation <= settings.eps_prim);
    }

    /** sqp_base.hpp:568-696 */
    void solve()
    {
        info.status = SQP_MAX_ITER_EXCEEDED;
        std::vector<double> p(N), p_lambda(DUAL, 0.0);
        info.qp_solver_iter = 0;
        info.iter = 1;
        tr_qp_iter.clear(); tr_bfgs.clear(); tr_ls_trials.clear(); tr_alpha.clear(); tr_qp_factor.clear();
        double lagv;
        // linearisation (309-318): exact Hessian
        problem.lagrangian_gradient_hessian(x.data(), p_static.data(), lam.data(), lagv, lag_gradient.data(), H.data(), h.data(),
                                            al.data(), A.data());
        regularise_hessian();
        tr_bfgs.push_back(-1);
        prepare_qp_bounds();
        if (iterate_tail(p, p_lambda)) { info.status = SQP_SOLVED; return; }
        while (info.iter < settings.max_iter) {
            info.iter++;
            if (opt_exact_hessian) {
                problem.lagrangian_gradient_hessian(x.data(), p_static.data(), lam.data(), lagv, lag_gradient.data(), H.data(), h.data(),
                                                    al.data(), A.data());
                regularise_hessian();
                tr_bfgs.push_back(-1);
                prepare_qp_bounds();
                if (iterate_tail(p, p_lambda)) { info.status = SQP_SOLVED; break; }
                continue;
            }
            // update_linearisation_dense_impl (489-504)
            std::vector<double> lg(N), yv(N);
            problem.lagrangian_gradient(x.data(), p_static.data(), lam.data(), lagv, lg.data(), h.data(), al.data(), A.data());
            for (int i = 0; i < N; ++i) yv[i] = lg[i] - lag_gradient[i];
            tr_bfgs.push_back(opt_block_bfgs ? block_bfgs_update<OcpT>(H.data(), step_prev.data(), yv.data())
                                             : bfgs_update(H.data(), step_prev.data(), yv.data(), N));
            lag_gradient = lg;
            prepare_qp_bounds();
            if (iterate_tail(p, p_lambda)) { info.status = SQP_SOLVED; break; }
        }
    }
};

} // namespace orc
