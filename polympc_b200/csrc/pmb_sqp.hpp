// pmb_sqp.hpp — per-instance pieces of the SQP iteration around the QP: damped BFGS, QP data preparation, l1-merit
// backtracking line search, step and termination test.  One warp per instance.
//
// Reference: src/solvers/sqp_base.hpp — step_size_selection_impl 378-419, constraints_violation_impl 421-444,
// max_constraints_violation_impl 446-474, update_linearisation_dense_impl 489-504, termination_criteria_impl 523-529,
// solve 568-696 (QP bounds 588-593, step 617-632); src/solvers/bfgs.hpp:23-52.
//
// Reductions: long dot products / 1-norms are "tree32" sums — lane l accumulates elements l, l+32, ... sequentially and
// the 32 partials are combined with an xor butterfly (16,8,4,2,1); infinity norms are exact maxima.
#pragma once
#include "pmb_ocp.hpp"

namespace pmb {

PMB_DEV double warp_sum_butterfly(const Warp& w, double p)
{
    for (int off = 16; off >= 1; off >>= 1) p = p + w.shfl_xor(p, off);
    return p;
}
PMB_DEV double dot_tree32(const Warp& w, const double* a, const double* b, int n)
{
    double acc = 0.0;
    for (int i = w.lane(); i < n; i += 32) acc = dm::fma(a[i], b[i], acc);
    return warp_sum_butterfly(w, acc);
}
PMB_DEV double norm_inf_warp(const Warp& w, const double* a, int n)
{
    double m = 0.0;
    for (int i = w.lane(); i < n; i += 32) { const double v = dm::fabs(a[i]); if (v > m) m = v; }
    for (int off = 16; off >= 1; off >>= 1) { const double o = w.shfl_xor(m, off); if (o > m) m = o; }
    return m;
}

/** bfgs.hpp:23-52.  B (n x n, column-major) is updated in place; Bs, r: scratch of n doubles each (shared memory).
 *  Returns 0 plain, 1 damped, 2 skipped. */
PMB_DEV int bfgs_update_warp(const Warp& w, int n, double* B, const double* s, const double* y, double* Bs, double* r)
{
    const int lane = w.lane();
    for (int i = lane; i < n; i += 32) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc = dm::fma(B[i + (size_t)j * n], s[j], acc);
        Bs[i] = acc;
    }
    w.sync();
    const double sBs = dot_tree32(w, s, Bs, n);
    const double sy = dot_tree32(w, s, y, n);
    double sr;
    int branch;
    if (sy < 0.2 * sBs) {
        const double theta = 0.8 * sBs / (sBs - sy);
        for (int i = lane; i < n; i += 32) r[i] = theta * y[i] + (1 - theta) * Bs[i];
        sr = theta * sy + (1 - theta) * sBs;
        branch = 1;
    } else {
        for (int i = lane; i < n; i += 32) r[i] = y[i];
        sr = sy;
        branch = 0;
    }
    w.sync();
    if (sr < DBL_EPSILON) return 2;
    for (int j = 0; j < n; ++j) {
        const double bsj = Bs[j], rj = r[j];
        for (int i = lane; i < n; i += 32) {
            double b = B[i + (size_t)j * n];
            b += ((-Bs[i]) * bsj) / sBs;
            b += (r[i] * rj) / sr;
            B[i + (size_t)j * n] = b;
        }
    }
    w.sync();
    return branch;
}

/** per-instance views of the SQP state in global memory */
struct SqpInst {
    double *x, *lam, *lam_k, *H, *A, *h, *al, *au, *lx, *ux, *lag_grad, *step_prev, *p, *plam, *stats;
    const double *lbx, *ubx, *lbg, *ubg, *d;
    pmb_sqp_info_t* info;
    pmb_qp_info_t* qp_info;
    int* qp_nfac;
    // decision trace rows (may be null)
    int *tr_qp_iter, *tr_bfgs, *tr_ls, *tr_qp_factor;
    double* tr_alpha;
};

template <class O>
struct SqpDev {
    using E = OcpEval<O>;
    static constexpr int N = O::N, M = O::M, NUM_EQ = O::NUM_EQ, NUM_INEQ = O::NUM_INEQ, DUAL = O::DUAL;

    /** shared scratch (doubles) needed by linearise / step kernels */
    static constexpr int SCRATCH_DOUBLES = 4 * N + M + 8;

    /** sqp_base.hpp:421-444; c/g scratch of M doubles */
    PMB_DEV static double constraints_violation(const Warp& w, const O& o, const double* xv, const SqpInst& s, double* cg)
    {
        E::equalities(w, o, xv, s.d, cg);
        E::inequalities(w, o, xv, s.d, cg + NUM_EQ);
        w.sync();
        const int lane = w.lane();
        double cl1 = DBL_EPSILON;
        double acc = 0.0;
        for (int i = lane; i < NUM_EQ; i += 32) acc = acc + dm::fabs(cg[i]);
        cl1 += warp_sum_butterfly(w, acc);
        if (NUM_INEQ > 0) {
            acc = 0.0;
            for (int i = lane; i < NUM_INEQ; i += 32) acc = acc + dm::max(s.lbg[i] - cg[NUM_EQ + i], 0.0);
            cl1 += warp_sum_butterfly(w, acc);
            acc = 0.0;
            for (int i = lane; i < NUM_INEQ; i += 32) acc = acc + dm::max(cg[NUM_EQ + i] - s.ubg[i], 0.0);
            cl1 += warp_sum_butterfly(w, acc);
        }
        acc = 0.0;
        for (int i = lane; i < N; i += 32) acc = acc + dm::max(s.lbx[i] - xv[i], 0.0);
        cl1 += warp_sum_butterfly(w, acc);
        acc = 0.0;
        for (int i = lane; i < N; i += 32) acc = acc + dm::max(xv[i] - s.ubx[i], 0.0);
        cl1 += warp_sum_butterfly(w, acc);
        w.sync();
        return cl1;
    }

    /** sqp_base.hpp:446-474 */
    PMB_DEV static double max_constraints_violation(const Warp& w, const O& o, const double* xv, const SqpInst& s, double* cg)
    {
        const int lane = w.lane();
        double c = 0.0;
        if (NUM_EQ > 0) {
            E::equalities(w, o, xv, s.d, cg);
            w.sync();
            c = norm_inf_warp(w, cg, NUM_EQ);
        }
        const double NEG = -dm::inf();
        if (NUM_INEQ > 0) {
            E::inequalities(w, o, xv, s.d, cg + NUM_EQ);
            w.sync();
            double m1 = NEG, m2 = NEG;
            for (int i = lane; i < NUM_INEQ; i += 32) {
                const double a = s.lbg[i] - cg[NUM_EQ + i], b = cg[NUM_EQ + i] - s.ubg[i];
                if (a > m1) m1 = a;
                if (b > m2) m2 = b;
            }
            m1 = warp_max(w, m1); m2 = warp_max(w, m2);
            c = fmax_nan(c, m1); c = fmax_nan(c, m2);
        }
        double m1 = NEG, m2 = NEG;
        for (int i = lane; i < N; i += 32) {
            const double a = s.lbx[i] - xv[i], b = xv[i] - s.ubx[i];
            if (a > m1) m1 = a;
            if (b > m2) m2 = b;
        }
        m1 = warp_max(w, m1); m2 = warp_max(w, m2);
        c = fmax_nan(c, m1); c = fmax_nan(c, m2);
        w.sync();
        return c;
    }

    /** first (exact Hessian) or later (BFGS) linearisation + QP bounds (sqp_base.hpp:583-593, 649-657, 489-504) */
    PMB_DEV static void linearise(const Warp& w, const O& o, const SqpInst& s, bool first, int trace_row, double* scratch)
    {
        const int lane = w.lane();
        if (first) {
            E::lagrangian_gradient_hessian(w, o, s.x, s.d, s.lam, s.lag_grad, s.H, s.h, s.al, s.A);
            if (s.tr_bfgs && lane == 0) s.tr_bfgs[trace_row] = -1;
        } else {
            double* lg = scratch;          // N
            double* yv = lg + N;           // N
            double* Bs = yv + N;           // N
            double* r = Bs + N;            // N
            E::lagrangian_gradient(w, o, s.x, s.d, s.lam, lg, s.h, s.al, s.A);
            for (int i = lane; i < N; i += 32) yv[i] = lg[i] - s.lag_grad[i];
            w.sync();
            const int br = bfgs_update_warp(w, N, s.H, s.step_prev, yv, Bs, r);
            if (s.tr_bfgs && lane == 0) s.tr_bfgs[trace_row] = br;
            for (int i = lane; i < N; i += 32) s.lag_grad[i] = lg[i];
        }
        // sqp_base.hpp:588-593
        for (int i = lane; i < M; i += 32) {
            double a = -s.al[i];
            double b = a;
            if (i >= NUM_EQ) { a += s.lbg[i - NUM_EQ]; b += s.ubg[i - NUM_EQ]; }
            s.al[i] = a; s.au[i] = b;
        }
        for (int i = lane; i < N; i += 32) { s.lx[i] = s.lbx[i] - s.x[i]; s.ux[i] = s.ubx[i] - s.x[i]; }
        w.sync();
    }

    /** everything after the QP: multipliers, line search, step, norms, termination.  Returns true when converged. */
    PMB_DEV static bool step(const Warp& w, const O& o, const SqpInst& s, const pmb_sqp_settings_t& st, int trace_row, double* scratch)
    {
        const int lane = w.lane();
        double* x_step = scratch;       // N
        double* cg = x_step + N;        // M
        // solve_qp bookkeeping (sqp_base.hpp:532-565) and lam_k / p_lambda (617-619)
        if (lane == 0) {
            s.info->qp_solver_iter += s.qp_info->iter;
            if (s.tr_qp_iter) s.tr_qp_iter[trace_row] = s.qp_info->iter;
            if (s.tr_qp_factor) s.tr_qp_factor[trace_row] = *s.qp_nfac;
        }
        for (int i = lane; i < DUAL; i += 32) { const double v = s.plam[i]; s.lam_k[i] = v; s.plam[i] = v - s.lam[i]; }
        w.sync();

        // ---- step_size_selection_impl (378-419)
        const double constr_l1 = constraints_violation(w, o, s.x, s, cg);
        const double mu = norm_inf_warp(w, s.lam_k, DUAL);
        const double cost_1 = E::cost(w, o, s.x, s.d);
        const double phi_l1 = cost_1 + mu * constr_l1;
        const double Dp_phi_l1 = dot_tree32(w, s.h, s.p, N) - mu * constr_l1;
        double alpha = 1.0, cost_step = 0.0;
        int trials = 0;
        bool accepted = false;
        for (int it = 1; it < st.line_search_max_iter; ++it) {
            for (int j = lane; j < N; j += 32) { double v = alpha * s.p[j]; v += s.x[j]; x_step[j] = v; }
            w.sync();
            cost_step = E::cost(w, o, x_step, s.d);
            ++trials;
            const double phi_l1_step = cost_step + mu * constraints_violation(w, o, x_step, s, cg);
            if (phi_l1_step <= (phi_l1 + alpha * st.eta * Dp_phi_l1)) { accepted = true; break; }
            alpha = st.tau * alpha;
        }
        (void)accepted;

        // ---- take the step (626-632)
        for (int i = lane; i < N; i += 32) { const double sp = alpha * s.p[i]; s.x[i] += sp; s.step_prev[i] = sp; }
        for (int i = lane; i < DUAL; i += 32) s.lam[i] += alpha * s.plam[i];
        const double primal_norm = alpha * norm_inf_warp(w, s.p, N);
        const double dual_norm = alpha * norm_inf_warp(w, s.plam, DUAL);
        w.sync();
        // ---- termination_criteria_impl (523-529)
        const double max_viol = max_constraints_violation(w, o, s.x, s, cg);
        const bool done = (primal_norm <= st.eps_prim) && (dual_norm <= st.eps_dual) && (max_viol <= st.eps_prim);
        if (lane == 0) {
            s.stats[0] = cost_step; s.stats[1] = primal_norm; s.stats[2] = dual_norm; s.stats[3] = max_viol;
            if (s.tr_alpha) s.tr_alpha[trace_row] = alpha;
            if (s.tr_ls) s.tr_ls[trace_row] = trials;
        }
        return done;
    }
};

} // namespace pmb
