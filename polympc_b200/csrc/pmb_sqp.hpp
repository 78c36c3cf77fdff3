// pmb_sqp.hpp — the SQP iteration around the QP: damped BFGS, QP data preparation, l1-merit backtracking line search,
// step and termination test, and the per-instance SQP loop.  One CTA per instance.
//
// Reference: src/solvers/sqp_base.hpp — step_size_selection_impl 378-419, constraints_violation_impl 421-444,
// max_constraints_violation_impl 446-474, update_linearisation_dense_impl 489-504, termination_criteria_impl 523-529,
// solve 568-696 (QP bounds 588-593, step 617-632); src/solvers/bfgs.hpp:23-52.
//
// Reductions: long dot products / 1-norms are "tree32" sums — lane l of warp 0 accumulates elements l, l+32, ...
// sequentially and the 32 partials are combined with an xor butterfly (16,8,4,2,1); infinity norms are exact maxima taken
// over the whole block.
#pragma once
#include "pmb_ocp.hpp"
#include "pmb_qp.hpp"

namespace pmb {

PMB_DEV double dot_tree32(Cta& c, const double* a, const double* b, int n)
{ return sum_tree32(c, n, [&](int i, double acc) { return dm::fma(a[i], b[i], acc); }); }

PMB_DEV double norm_inf_cta(Cta& c, const double* a, int n)
{
    double m = 0.0;
    for (int i = c.tid(); i < n; i += c.nthreads()) { const double v = dm::fabs(a[i]); if (v > m) m = v; }
    return c.max_all1(m);
}

/** bfgs.hpp:23-52.  B (n x n, column-major) is updated in place; Bs, r: scratch of n doubles each (shared memory).
 *  Returns 0 plain, 1 damped, 2 skipped (uniform over the block). */
PMB_DEV int bfgs_update_cta(Cta& c, int n, double* B, const double* s, const double* y, double* Bs, double* r)
{
    const int tid = c.tid(), nt = c.nthreads();
    for (int i = tid; i < n; i += nt) Bs[i] = dot_chain(B + i, (size_t)n, s, n);
    c.sync();
    const double sBs = dot_tree32(c, s, Bs, n);
    const double sy = dot_tree32(c, s, y, n);
    double sr;
    int branch;
    if (sy < 0.2 * sBs) {
        const double theta = 0.8 * sBs / (sBs - sy);
        for (int i = tid; i < n; i += nt) r[i] = theta * y[i] + (1 - theta) * Bs[i];
        sr = theta * sy + (1 - theta) * sBs;
        branch = 1;
    } else {
        for (int i = tid; i < n; i += nt) r[i] = y[i];
        sr = sy;
        branch = 0;
    }
    c.sync();
    if (sr < DBL_EPSILON) return 2;
    for (int e = tid; e < n * n; e += nt) {
        const int j = e / n, i = e - j * n;
        double b = B[e];
        b += ((-Bs[i]) * Bs[j]) / sBs;
        b += (r[i] * r[j]) / sr;
        B[e] = b;
    }
    c.sync();
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
    unsigned long long* phase;   // profiling counters (may be null)
};

template <class O>
struct SqpDev {
    using E = OcpEval<O>;
    static constexpr int N = O::N, M = O::M, NUM_EQ = O::NUM_EQ, NUM_INEQ = O::NUM_INEQ, DUAL = O::DUAL;

    /** shared scratch (doubles) needed by linearise / step */
    static constexpr int SCRATCH_DOUBLES = 4 * N + M + 8 + E::NV_DOUBLES;

    /** sqp_base.hpp:421-444; cg: scratch of M doubles */
    PMB_DEV static double constraints_violation(Cta& c, const O& o, const double* xv, const SqpInst& s, double* cg)
    {
        E::equalities(c, o, xv, s.d, cg);
        E::inequalities(c, o, xv, s.d, cg + NUM_EQ);
        c.sync();
        double cl1 = DBL_EPSILON;
        cl1 += sum_tree32(c, NUM_EQ, [&](int i, double acc) { return acc + dm::fabs(cg[i]); });
        if (NUM_INEQ > 0) {
            cl1 += sum_tree32(c, NUM_INEQ, [&](int i, double acc) { return acc + dm::max(s.lbg[i] - cg[NUM_EQ + i], 0.0); });
            cl1 += sum_tree32(c, NUM_INEQ, [&](int i, double acc) { return acc + dm::max(cg[NUM_EQ + i] - s.ubg[i], 0.0); });
        }
        cl1 += sum_tree32(c, N, [&](int i, double acc) { return acc + dm::max(s.lbx[i] - xv[i], 0.0); });
        cl1 += sum_tree32(c, N, [&](int i, double acc) { return acc + dm::max(xv[i] - s.ubx[i], 0.0); });
        return cl1;
    }

    /** sqp_base.hpp:446-474 */
    PMB_DEV static double max_constraints_violation(Cta& c, const O& o, const double* xv, const SqpInst& s, double* cg)
    {
        const int tid = c.tid(), nt = c.nthreads();
        double cv = 0.0;
        if (NUM_EQ > 0) {
            E::equalities(c, o, xv, s.d, cg);
            c.sync();
            cv = norm_inf_cta(c, cg, NUM_EQ);
        }
        const double NEG = -dm::inf();
        if (NUM_INEQ > 0) {
            E::inequalities(c, o, xv, s.d, cg + NUM_EQ);
            c.sync();
            double m[2] = {NEG, NEG};
            for (int i = tid; i < NUM_INEQ; i += nt) {
                const double a = s.lbg[i] - cg[NUM_EQ + i], b = cg[NUM_EQ + i] - s.ubg[i];
                if (a > m[0]) m[0] = a;
                if (b > m[1]) m[1] = b;
            }
            c.max_all<2>(m);
            cv = fmax_nan(cv, m[0]); cv = fmax_nan(cv, m[1]);
        }
        double m[2] = {NEG, NEG};
        for (int i = tid; i < N; i += nt) {
            const double a = s.lbx[i] - xv[i], b = xv[i] - s.ubx[i];
            if (a > m[0]) m[0] = a;
            if (b > m[1]) m[1] = b;
        }
        c.max_all<2>(m);
        cv = fmax_nan(cv, m[0]); cv = fmax_nan(cv, m[1]);
        return cv;
    }

    /** first (exact Hessian) or later (BFGS) linearisation + QP bounds (sqp_base.hpp:583-593, 649-657, 489-504) */
    PMB_DEV static void linearise(Cta& c, const O& o, const SqpInst& s, bool first, int trace_row, double* scratch)
    {
        const int tid = c.tid(), nt = c.nthreads();
        if (first) {
            E::lagrangian_gradient_hessian(c, o, s.x, s.d, s.lam, s.lag_grad, s.H, s.h, s.al, s.A, scratch);
            if (s.tr_bfgs && tid == 0) s.tr_bfgs[trace_row] = -1;
        } else {
            double* lg = scratch;          // N
            double* yv = lg + N;           // N
            double* Bs = yv + N;           // N
            double* r = Bs + N;            // N
            E::lagrangian_gradient(c, o, s.x, s.d, s.lam, lg, s.h, s.al, s.A);
            for (int i = tid; i < N; i += nt) yv[i] = lg[i] - s.lag_grad[i];
            c.sync();
            const int br = bfgs_update_cta(c, N, s.H, s.step_prev, yv, Bs, r);
            if (s.tr_bfgs && tid == 0) s.tr_bfgs[trace_row] = br;
            for (int i = tid; i < N; i += nt) s.lag_grad[i] = lg[i];
        }
        // sqp_base.hpp:588-593
        for (int i = tid; i < M; i += nt) {
            double a = -s.al[i];
            double b = a;
            if (i >= NUM_EQ) { a += s.lbg[i - NUM_EQ]; b += s.ubg[i - NUM_EQ]; }
            s.al[i] = a; s.au[i] = b;
        }
        for (int i = tid; i < N; i += nt) { s.lx[i] = s.lbx[i] - s.x[i]; s.ux[i] = s.ubx[i] - s.x[i]; }
        c.sync();
    }

    /** everything after the QP: multipliers, line search, step, norms, termination.  Returns true when converged. */
    PMB_DEV static bool step(Cta& c, const O& o, const SqpInst& s, const pmb_sqp_settings_t& st, int trace_row, double* scratch)
    {
        const int tid = c.tid(), nt = c.nthreads();
        double* x_step = scratch;       // N
        double* cg = x_step + N;        // M
        // solve_qp bookkeeping (sqp_base.hpp:532-565) and lam_k / p_lambda (617-619)
        if (tid == 0) {
            s.info->qp_solver_iter += s.qp_info->iter;
            if (s.tr_qp_iter) s.tr_qp_iter[trace_row] = s.qp_info->iter;
            if (s.tr_qp_factor) s.tr_qp_factor[trace_row] = *s.qp_nfac;
        }
        for (int i = tid; i < DUAL; i += nt) { const double v = s.plam[i]; s.lam_k[i] = v; s.plam[i] = v - s.lam[i]; }
        c.sync();

        // ---- step_size_selection_impl (378-419)
        const double constr_l1 = constraints_violation(c, o, s.x, s, cg);
        const double mu = norm_inf_cta(c, s.lam_k, DUAL);
        const double cost_1 = E::cost(c, o, s.x, s.d);
        const double phi_l1 = cost_1 + mu * constr_l1;
        const double Dp_phi_l1 = dot_tree32(c, s.h, s.p, N) - mu * constr_l1;
        double alpha = 1.0, cost_step = 0.0;
        int trials = 0;
        for (int it = 1; it < st.line_search_max_iter; ++it) {
            for (int j = tid; j < N; j += nt) { double v = alpha * s.p[j]; v += s.x[j]; x_step[j] = v; }
            c.sync();
            cost_step = E::cost(c, o, x_step, s.d);
            ++trials;
            const double phi_l1_step = cost_step + mu * constraints_violation(c, o, x_step, s, cg);
            if (phi_l1_step <= (phi_l1 + alpha * st.eta * Dp_phi_l1)) break;
            alpha = st.tau * alpha;
        }

        // ---- take the step (626-632)
        const double primal_norm = alpha * norm_inf_cta(c, s.p, N);
        const double dual_norm = alpha * norm_inf_cta(c, s.plam, DUAL);
        for (int i = tid; i < N; i += nt) { const double sp = alpha * s.p[i]; s.x[i] += sp; s.step_prev[i] = sp; }
        for (int i = tid; i < DUAL; i += nt) s.lam[i] += alpha * s.plam[i];
        c.sync();
        // ---- termination_criteria_impl (523-529)
        const double max_viol = max_constraints_violation(c, o, s.x, s, cg);
        const bool done = (primal_norm <= st.eps_prim) && (dual_norm <= st.eps_dual) && (max_viol <= st.eps_prim);
        if (tid == 0) {
            s.stats[0] = cost_step; s.stats[1] = primal_norm; s.stats[2] = dual_norm; s.stats[3] = max_viol;
            if (s.tr_alpha) s.tr_alpha[trace_row] = alpha;
            if (s.tr_ls) s.tr_ls[trace_row] = trials;
        }
        return done;
    }

    /** SQPBase::solve (sqp_base.hpp:568-696) of one instance: iterate linearise -> QP -> line search / step until the
     *  termination test holds or max_iter QPs were solved.  Lp / vec: QP workspaces (pmb_qp.hpp), scratch: SCRATCH_DOUBLES. */
    template <int R>
    PMB_DEV static void solve(Cta& c, const O& o, const SqpInst& s, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst,
                              double* Lp, unsigned char* vec, double* scratch)
    {
        if (c.tid() == 0) { s.info->iter = 1; s.info->qp_solver_iter = 0; s.info->status = PMB_SQP_MAX_ITER_EXCEEDED; }
        QpArgs qa;
        qa.N = N; qa.M = M; qa.H = s.H; qa.h = s.h; qa.A = s.A; qa.Alb = s.al; qa.Aub = s.au; qa.xlb = s.lx; qa.xub = s.ux;
        qa.xg = nullptr; qa.yg = nullptr; qa.x = s.p; qa.y = s.plam; qa.info = s.qp_info; qa.z = nullptr; qa.q = nullptr;
        qa.perm = nullptr; qa.ctype = nullptr; qa.nfac = s.qp_nfac;
        QpProf qprof;
        qa.prof = s.phase ? &qprof : nullptr;
        c.sync();
        unsigned long long t_lin = 0, t_qp = 0, t_step = 0, n_it = 0;
        for (int it = 1; it <= st.max_iter; ++it) {
            const int row = it - 1;
            const unsigned long long t0 = c.w.clock();
            linearise(c, o, s, it == 1, row, scratch);
            const unsigned long long t1 = c.w.clock();
            qp_solve_cta<R>(c, qst, qa, Lp, vec);
            const unsigned long long t2 = c.w.clock();
            const bool done = step(c, o, s, st, row, scratch);
            const unsigned long long t3 = c.w.clock();
            t_lin += t1 - t0; t_qp += t2 - t1; t_step += t3 - t2; ++n_it;
            if (done) { if (c.tid() == 0) s.info->status = PMB_SQP_SOLVED; break; }
            if (it < st.max_iter && c.tid() == 0) s.info->iter = it + 1;
        }
        if (s.phase && c.tid() == 0) {
            atomic_add_u64(s.phase + 0, t_lin); atomic_add_u64(s.phase + 1, t_qp); atomic_add_u64(s.phase + 2, t_step); atomic_add_u64(s.phase + 3, n_it);
            atomic_add_u64(s.phase + 4, qprof.pivot); atomic_add_u64(s.phase + 5, qprof.gather); atomic_add_u64(s.phase + 6, qprof.factor);
            atomic_add_u64(s.phase + 7, qprof.solve); atomic_add_u64(s.phase + 8, qprof.update); atomic_add_u64(s.phase + 9, qprof.resid);
            atomic_add_u64(s.phase + 10, (unsigned long long)s.info->qp_solver_iter);
        }
        c.sync();
    }
};

} // namespace pmb
