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
#include "pmb_precond.hpp"
#include "pmb_qp_admm.hpp"

namespace pmb {

PMB_DEV double dot_tree32(Cta& c, const double* a, const double* b, int n)
{ return sum_tree32(c, n, [&](int i, double acc) { return dm::fma(a[i], b[i], acc); }); }

/** lpNorm<Infinity>() = cwiseAbs().maxCoeff(): a std::max-like chain that starts from the first coefficient, so a NaN in
 *  a[0] sticks and a NaN anywhere else is skipped (oracle/canon.hpp::norm_inf).  The maximum over the non-NaN entries is
 *  exact and order free; the first-coefficient rule is applied afterwards. */
PMB_DEV double norm_inf_cta(Cta& c, const double* a, int n)
{
    double m[2] = {0.0, 0.0};                                   // [1]: flag "the first coefficient is NaN"
    for (int i = c.tid(); i < n; i += c.nthreads()) {
        const double v = dm::fabs(a[i]);
        if (v > m[0]) m[0] = v;
        if (i == 0 && v != v) m[1] = 1.0;
    }
    c.max_all<2>(m);
    return m[1] > 0.0 ? dm::fabs(a[0]) : m[0];                  // (the load only happens on the NaN path)
}

/** bfgs.hpp:23-52.  B (n x n, column-major) is updated in place; Bs, r: scratch of n doubles each (shared memory).
 *  Returns 0 plain, 1 damped, 2 skipped (uniform over the block). */
template <bool FAST = false>
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
    // columns over the warps, rows over the lanes (no integer division per element).  Exact arithmetic divides per element like
    // bfgs.hpp:48-49 (`-Bs Bs^T / sBs`, `r r^T / sr`); fast arithmetic multiplies by the two reciprocals instead.
    const double isBs = FAST ? 1.0 / sBs : 0.0, isr = FAST ? 1.0 / sr : 0.0;
    for (int j = c.warp_id(); j < n; j += c.nwarps()) {
        const double bsj = Bs[j], rj = r[j];
        double* col = B + (size_t)j * n;
        for (int i = c.lane(); i < n; i += 32) {
            double b = col[i];
            if (FAST) { b += ((-Bs[i]) * bsj) * isBs; b += (r[i] * rj) * isr; }
            else { b += ((-Bs[i]) * bsj) / sBs; b += (r[i] * rj) / sr; }
            col[i] = b;
        }
    }
    c.sync();
    return branch;
}

/** ContinuousOCP<..., SPARSE>::hessian_update_impl (continuous_ocp.hpp:2303-2431), the block BFGS the reference's control
 *  tests install: only the per-node (x_k, u_k) blocks and the parameter rows / columns of H are stored and updated.
 *  v = H s over the stored pattern (ascending column index, fused chain), coefficients -1/s'v, 1/s'y (or 1/s'r, damped)
 *  multiplied into the first factor of every outer product, the two increments of an entry summed before they are added;
 *  the (x, u) and (., p) blocks are the transposes of the (u, x) and (p, .) ... see oracle/sqp.hpp::block_bfgs_update.
 *  v, r: scratch of N doubles each (shared memory).  Returns 0 plain, 1 damped (uniform over the block). */
template <class O>
PMB_DEV int block_bfgs_update_cta(Cta& c, double* H, const double* s, const double* y, double* v, double* r)
{
    constexpr int NX = O::NX, NU = O::NU, NP = O::NP, NN = O::NN, VARX = O::VARX, VARU = O::VARU, N = O::N, NB = NX + NU, P0 = VARX + VARU;
    const int tid = c.tid(), nt = c.nthreads();
    for (int row = tid; row < N; row += nt) {
        double acc = 0.0;
        if (row < P0) {
            const int k = row < VARX ? row / NX : (row - VARX) / NU;
            const double* Hr = H + row;
            for (int j = 0; j < NX; ++j) acc = dm::fma(Hr[(size_t)(k * NX + j) * N], s[k * NX + j], acc);
            for (int j = 0; j < NU; ++j) acc = dm::fma(Hr[(size_t)(VARX + k * NU + j) * N], s[VARX + k * NU + j], acc);
            for (int j = 0; j < NP; ++j) acc = dm::fma(Hr[(size_t)(P0 + j) * N], s[P0 + j], acc);
        } else {
            acc = dot_chain(H + row, (size_t)N, s, N);
        }
        v[row] = acc;
    }
    c.sync();
    const double scaling = dot_tree32(c, s, v, N);
    const double c1 = -(1.0 / scaling);
    const double sy = dot_tree32(c, s, y, N);
    const bool plain = sy >= 0.2 * scaling;
    double c2 = 1.0 / sy;
    if (!plain) {
        const double theta = 0.8 * scaling / (scaling - sy);
        for (int i = tid; i < N; i += nt) r[i] = theta * y[i] + (1 - theta) * v[i];
        c.sync();
        c2 = 1.0 / dot_tree32(c, s, r, N);
    } else {
        for (int i = tid; i < N; i += nt) r[i] = y[i];
        c.sync();
    }
    auto inc = [&](int i, int j) { double h = (c1 * v[i]) * v[j]; h += (c2 * r[i]) * r[j]; return h; };
    // node blocks: entry (a, b) of the (NX+NU)^2 block of node k; the (x, u) part is the transpose of the (u, x) part
    for (int e = tid; e < NN * NB * NB; e += nt) {
        const int k = e / (NB * NB), q = e - k * NB * NB, b = q / NB, a = q - b * NB;
        const int gi = a < NX ? k * NX + a : VARX + k * NU + (a - NX);
        const int gj = b < NX ? k * NX + b : VARX + k * NU + (b - NX);
        const double h = (a < NX && b >= NX) ? inc(gj, gi) : inc(gi, gj);
        H[gi + (size_t)gj * N] += h;
    }
    if (NP > 0) {
        constexpr int NPd = NP > 0 ? NP : 1;
        for (int e = tid; e < NP * NP; e += nt) { const int j = e / NPd, i = e - j * NPd; H[(P0 + i) + (size_t)(P0 + j) * N] += inc(P0 + i, P0 + j); }
        for (int e = tid; e < NP * P0; e += nt) {
            const int j = e / P0, i = e - j * P0;
            const double h = inc(i, P0 + j);
            H[i + (size_t)(P0 + j) * N] += h;
            H[(P0 + j) + (size_t)i * N] += h;
        }
    }
    c.sync();
    return plain ? 0 : 1;
}

/** batch-wide SQP state in global memory (instance-major arrays) */
struct SqpWs {
    double *x, *lam, *lam_k, *H, *A, *h, *al, *au, *lx, *ux, *lbx, *ubx, *lbg, *ubg, *d, *lag_grad, *step_prev, *p, *plam, *stats;
    pmb_sqp_info_t* info;
    pmb_qp_info_t* qp_info;
    int* qp_nfac;
    int *tr_qp_iter, *tr_bfgs, *tr_ls, *tr_qp_factor;
    double* tr_alpha;
    int trace_rows;
    int opt_exact_hessian;       // pmb_sqp_set_hessian_options: exact Lagrangian Hessian at every iteration (no BFGS)
    int opt_gershgorin;          //                              Gershgorin shift after every exact Hessian
    int opt_block_bfgs;          // pmb_sqp_set_hessian_update: the OCP's block BFGS instead of the dense damped BFGS
    int opt_precond;             // pmb_sqp_set_preconditioner: PRECOND_IDENTITY / RUIZ_DENSE / RUIZ_SPARSE
    int opt_line_search;         // pmb_sqp_set_line_search: LS_L1_MERIT / LS_FILTER
    int filter_depth;            //                          LSFilter::max_depth
    double filter_beta;          //                          LSFilter::beta
    double* Ae;                  // batch x (M + N) x N: [A; I] of ADMM<> (only with pmb_sqp_set_qp_solver(PMB_QP_OSQP_ADMM))
    double* ruiz;                // batch x (N + M + 1): D, E, c of the current QP (only with a preconditioner)
    double* filter;              // batch x FILTER_DOUBLES: the filter of each solver object (only with the filter line search)
    const int* order;            // work-queue order (pmb_sqp_set_schedule): ticket q solves instance order[q]; nullptr = identity
    unsigned long long* phase;   // profiling, cycles of thread 0 summed over CTAs: {linearise, qp, step}, [3] = instance-iterations,
                                 // [4..9] = QP {pivot, gather, factor, solve, update, resid}, [10] = ADMM trips, [11] = line-search trials
};

/** per-instance view of the SQP state.  Only (workspace, instance index) are held; every pointer is recomputed from the
 *  kernel parameters where it is used — 25 live 64-bit pointers across the QP were the main cause of register spills in the
 *  fused kernel. */
template <class O>
struct SqpInst {
    const SqpWs& ws;
    int b;
#define PMB_INST_F(T, name, len) PMB_DEV T* name() const { return ws.name + (size_t)b * (size_t)(len); }
    PMB_INST_F(double, x, O::N) PMB_INST_F(double, lam, O::DUAL) PMB_INST_F(double, lam_k, O::DUAL) PMB_INST_F(double, H, O::N * O::N)
    PMB_INST_F(double, A, O::M * O::N) PMB_INST_F(double, h, O::N) PMB_INST_F(double, al, O::M) PMB_INST_F(double, au, O::M)
    PMB_INST_F(double, lx, O::N) PMB_INST_F(double, ux, O::N) PMB_INST_F(double, lag_grad, O::N) PMB_INST_F(double, step_prev, O::N)
    PMB_INST_F(double, p, O::N) PMB_INST_F(double, plam, O::DUAL) PMB_INST_F(double, stats, 4)
    PMB_INST_F(const double, lbx, O::N) PMB_INST_F(const double, ubx, O::N) PMB_INST_F(const double, lbg, O::NUM_INEQ)
    PMB_INST_F(const double, ubg, O::NUM_INEQ) PMB_INST_F(const double, d, O::ND)
    PMB_INST_F(double, Ae, (O::N + O::M) * O::N) PMB_INST_F(double, ruiz, O::N + O::M + 1) PMB_INST_F(double, filter, FILTER_DOUBLES)
    PMB_INST_F(pmb_sqp_info_t, info, 1) PMB_INST_F(pmb_qp_info_t, qp_info, 1) PMB_INST_F(int, qp_nfac, 1)
#undef PMB_INST_F
    // decision trace rows (arrays may be null)
#define PMB_INST_T(T, name) PMB_DEV T* name() const { return ws.name ? ws.name + (size_t)b * (size_t)ws.trace_rows : nullptr; }
    PMB_INST_T(int, tr_qp_iter) PMB_INST_T(int, tr_bfgs) PMB_INST_T(int, tr_ls) PMB_INST_T(int, tr_qp_factor) PMB_INST_T(double, tr_alpha)
#undef PMB_INST_T
    PMB_DEV unsigned long long* phase() const { return ws.phase; }
};

template <class O>
struct SqpDev {
    using E = OcpEval<O>;
    static constexpr int N = O::N, M = O::M, NUM_EQ = O::NUM_EQ, NUM_INEQ = O::NUM_INEQ, DUAL = O::DUAL;

    /** shared scratch (doubles) needed by linearise / step */
    static constexpr int SCRATCH_DOUBLES = 4 * N + M + 8 + E::NV_DOUBLES;

    /** sqp_base.hpp:446-474 */
    /** have_cg: cg already holds c(xv) (and g(xv)) */
    PMB_DEV static double max_constraints_violation(Cta& c, const O& o, const double* xv, const SqpInst<O>& s, double* cg, bool have_cg)
    {
        const int tid = c.tid(), nt = c.nthreads();
        double cv = 0.0;
        if (!have_cg) {
            E::equalities(c, o, xv, s.d(), cg, tid);
            E::inequalities(c, o, xv, s.d(), cg + NUM_EQ, tid);
            c.sync();
        }
        if (NUM_EQ > 0) cv = norm_inf_cta(c, cg, NUM_EQ);
        const double NEG = -dm::inf();
        // maxCoeff() is the same kind of chain as lpNorm<Infinity>: it starts from the first coefficient, a NaN there sticks
        if (NUM_INEQ > 0) {
            double m[2] = {NEG, NEG};
            for (int i = tid; i < NUM_INEQ; i += nt) {
                const double a = s.lbg()[i] - cg[NUM_EQ + i], b = cg[NUM_EQ + i] - s.ubg()[i];
                if (a > m[0]) m[0] = a;
                if (b > m[1]) m[1] = b;
            }
            c.max_all<2>(m);
            const double a0 = s.lbg()[0] - cg[NUM_EQ], b0 = cg[NUM_EQ] - s.ubg()[0];
            if (a0 != a0) m[0] = a0;
            if (b0 != b0) m[1] = b0;
            cv = fmax_nan(cv, m[0]); cv = fmax_nan(cv, m[1]);
        }
        double m[2] = {NEG, NEG};
        for (int i = tid; i < N; i += nt) {
            const double a = s.lbx()[i] - xv[i], b = xv[i] - s.ubx()[i];
            if (a > m[0]) m[0] = a;
            if (b > m[1]) m[1] = b;
        }
        c.max_all<2>(m);
        const double a0 = s.lbx()[0] - xv[0], b0 = xv[0] - s.ubx()[0];
        if (a0 != a0) m[0] = a0;
        if (b0 != b0) m[1] = b0;
        cv = fmax_nan(cv, m[0]); cv = fmax_nan(cv, m[1]);
        return cv;
    }

    /** first (exact Hessian) or later (BFGS) linearisation + QP bounds (sqp_base.hpp:583-593, 649-657, 489-504) */
    /** returns the cost at x (the value every linearisation computes on the way) */
    template <bool FAST = false>
    PMB_DEV static double linearise(Cta& c, const O& o, const SqpInst<O>& s, bool first, int trace_row, double* scratch)
    {
        const int tid = c.tid(), nt = c.nthreads();
        double cost_x;
        if (first) {
            cost_x = E::lagrangian_gradient_hessian(c, o, s.x(), s.d(), s.lam(), s.lag_grad(), s.H(), s.h(), s.al(), s.A(), scratch);
            if (s.tr_bfgs() && tid == 0) s.tr_bfgs()[trace_row] = -1;
            if (s.ws.opt_gershgorin) {
                // hessian_regularisation_dense_impl of reference tests/control/minimal_time_test.cpp:90-104 (cold path): one
                // thread per column, |column| summed in sequential ascending order like the oracle; column i only changes (i, i)
                c.sync();
                double* Hm = s.H();
                for (int i = tid; i < N; i += nt) {
                    const double aii = Hm[i + (size_t)i * N];
                    double sum = 0.0;
                    for (int j = 0; j < N; ++j) sum += dm::fabs(Hm[j + (size_t)i * N]);
                    const double ri = sum - dm::fabs(aii);
                    if (aii - ri <= 0) Hm[i + (size_t)i * N] += (ri - aii) + 0.01;
                }
                c.sync();
            }
        } else {
            double* lg = scratch;          // N
            double* yv = lg + N;           // N
            double* Bs = yv + N;           // N
            double* r = Bs + N;            // N
            cost_x = E::lagrangian_gradient(c, o, s.x(), s.d(), s.lam(), lg, s.h(), s.al(), s.A(), r + N);
            for (int i = tid; i < N; i += nt) yv[i] = lg[i] - s.lag_grad()[i];
            c.sync();
            const int br = s.ws.opt_block_bfgs ? block_bfgs_update_cta<O>(c, s.H(), s.step_prev(), yv, Bs, r)
                                               : bfgs_update_cta<FAST>(c, N, s.H(), s.step_prev(), yv, Bs, r);
            if (s.tr_bfgs() && tid == 0) s.tr_bfgs()[trace_row] = br;
            for (int i = tid; i < N; i += nt) s.lag_grad()[i] = lg[i];
        }
        // sqp_base.hpp:588-593
        for (int i = tid; i < M; i += nt) {
            double a = -s.al()[i];
            double b = a;
            if (i >= NUM_EQ) { a += s.lbg()[i - NUM_EQ]; b += s.ubg()[i - NUM_EQ]; }
            s.al()[i] = a; s.au()[i] = b;
        }
        for (int i = tid; i < N; i += nt) { s.lx()[i] = s.lbx()[i] - s.x()[i]; s.ux()[i] = s.ubx()[i] - s.x()[i]; }
        c.sync();
        return cost_x;
    }

    /** merit ingredients at the point xv for the line search (sqp_base.hpp:395-399, 421-444): cost(xv) and
     *  viol_1(xv) = eps + |c|_1 + sum max(0, lbg - g) + sum max(0, g - ubg) + sum max(0, lbx - xv) + sum max(0, xv - ubx).
     *  The four independent pieces run on different warps (cost on warp 0, constraints on warp 1, the two box sums on
     *  warp 2) and meet at ONE barrier; every sum keeps its canonical tree32 order.  cg[M] receives c(xv) (and g(xv)). */
    PMB_DEV static void merit_terms(Cta& c, const O& o, const double* xv, const SqpInst<O>& s, double* cg, double& cost_out, double& viol_out)
    {
        const int nw = c.nwarps(), wid = c.warp_id(), lane = c.lane();
        double* slot = c.bc + Cta::BC_SLOTS;       // reduction scratch doubles as the meeting point: [cost, eq, ilb, iub, lbx, ubx]
        const Warp& w = c.w;
        auto tree = [&](int n, auto term) {        // tree32 sum inside the calling warp; result in every lane
            double acc = 0.0;
            for (int i = lane; i < n; i += 32) acc = term(i, acc);
            for (int off = 16; off >= 1; off >>= 1) acc = acc + w.shfl_xor(acc, off);
            return acc;
        };
        if (wid == 0) {
            const double cv = E::cost_warp0(c, o, xv, s.d());
            if (lane == 0) slot[0] = cv;
        }
        if (wid == 1 % nw) {
            E::equalities(c, o, xv, s.d(), cg, lane);
            E::inequalities(c, o, xv, s.d(), cg + NUM_EQ, lane);
            w.sync();
            const double se = tree(NUM_EQ, [&](int i, double acc) { return acc + dm::fabs(cg[i]); });
            double sl = 0.0, su = 0.0;
            if (NUM_INEQ > 0) {
                sl = tree(NUM_INEQ, [&](int i, double acc) { return acc + dm::max(s.lbg()[i] - cg[NUM_EQ + i], 0.0); });
                su = tree(NUM_INEQ, [&](int i, double acc) { return acc + dm::max(cg[NUM_EQ + i] - s.ubg()[i], 0.0); });
            }
            if (lane == 0) { slot[1] = se; slot[2] = sl; slot[3] = su; }
        }
        if (wid == 2 % nw) {
            const double sl = tree(N, [&](int i, double acc) { return acc + dm::max(s.lbx()[i] - xv[i], 0.0); });
            const double su = tree(N, [&](int i, double acc) { return acc + dm::max(xv[i] - s.ubx()[i], 0.0); });
            if (lane == 0) { slot[4] = sl; slot[5] = su; }
        }
        c.sync();
        cost_out = slot[0];
        double cl1 = DBL_EPSILON;
        cl1 += slot[1];
        if (NUM_INEQ > 0) { cl1 += slot[2]; cl1 += slot[3]; }
        cl1 += slot[4];
        cl1 += slot[5];
        viol_out = cl1;
        c.sync();                                   // the slots may be rewritten after this point
    }

    /** everything after the QP: multipliers, line search, step, norms, termination.  Returns true when converged.
     *  cost_x = cost at the current x (known from the linearisation); the constraint values at x are known as well:
     *  s.al() still holds -c(x) (the QP bounds) when there are no inequality rows. */
    PMB_DEV static bool step(Cta& c, const O& o, const SqpInst<O>& s, const pmb_sqp_settings_t& st, int trace_row, double* scratch, double cost_x)
    {
        const int tid = c.tid(), nt = c.nthreads();
        double* x_step = scratch;       // N
        double* cg = x_step + N;        // M
        // solve_qp bookkeeping (sqp_base.hpp:532-565) and lam_k / p_lambda (617-619)
        if (tid == 0) {
            s.info()->qp_solver_iter += s.qp_info()->iter;
            if (s.tr_qp_iter()) s.tr_qp_iter()[trace_row] = s.qp_info()->iter;
            if (s.tr_qp_factor()) s.tr_qp_factor()[trace_row] = *s.qp_nfac();
        }
        for (int i = tid; i < DUAL; i += nt) { const double v = s.plam()[i]; s.lam_k()[i] = v; s.plam()[i] = v - s.lam()[i]; }
        c.sync();

        // ---- step_size_selection_impl (378-419)
        double constr_l1, cost_1;
        const bool filter_ls = s.ws.opt_line_search == LS_FILTER;
        if (NUM_INEQ == 0 && s.ws.opt_precond == PRECOND_IDENTITY && !filter_ls) {
            // |c(x)|_1 from the QP bounds al = -c(x) (|-c| == |c| bit for bit), cost(x) from the linearisation
            double cl1 = DBL_EPSILON;
            cl1 += sum_tree32(c, NUM_EQ, [&](int i, double acc) { return acc + dm::fabs(s.al()[i]); });
            cl1 += sum_tree32(c, N, [&](int i, double acc) { return acc + dm::max(s.lbx()[i] - s.x()[i], 0.0); });
            cl1 += sum_tree32(c, N, [&](int i, double acc) { return acc + dm::max(s.x()[i] - s.ubx()[i], 0.0); });
            constr_l1 = cl1;
            cost_1 = cost_x;
        } else {
            merit_terms(c, o, s.x(), s, cg, cost_1, constr_l1);
        }
        double alpha = 1.0, cost_step = 0.0;
        int trials = 0;
        bool accepted = false;          // cg holds c(x + alpha p) of the accepted trial
        if (!filter_ls) {
            const double mu = norm_inf_cta(c, s.lam_k(), DUAL);
            const double phi_l1 = cost_1 + mu * constr_l1;
            const double Dp_phi_l1 = dot_tree32(c, s.h(), s.p(), N) - mu * constr_l1;
            for (int it = 1; it < st.line_search_max_iter; ++it) {
                for (int j = tid; j < N; j += nt) { double v = alpha * s.p()[j]; v += s.x()[j]; x_step[j] = v; }
                c.sync();
                double viol_step;
                merit_terms(c, o, x_step, s, cg, cost_step, viol_step);
                ++trials;
                const double phi_l1_step = cost_step + mu * viol_step;
                if (phi_l1_step <= (phi_l1 + alpha * st.eta * Dp_phi_l1)) { accepted = true; break; }
                alpha = st.tau * alpha;
            }
        } else {
            // the filter line search of reference tests/control/valet_parking_mpc_test.cpp:110-155 (LSFilter, line_search.hpp:30-98)
            double* flt = s.filter();
            const double beta = s.ws.filter_beta;
            if (filter_is_acceptable(flt, beta, cost_1, constr_l1)) filter_add(c, flt, s.ws.filter_depth, cost_1, constr_l1);
            for (int it = 1; it < st.line_search_max_iter; ++it) {
                for (int j = tid; j < N; j += nt) { double v = alpha * s.p()[j]; v += s.x()[j]; x_step[j] = v; }
                c.sync();
                double viol_step;
                merit_terms(c, o, x_step, s, cg, cost_step, viol_step);
                ++trials;
                if (filter_is_acceptable(flt, beta, cost_step, viol_step)) { filter_add(c, flt, s.ws.filter_depth, cost_step, viol_step); accepted = true; break; }
                alpha *= st.tau;
            }
        }

        // ---- take the step (626-632)
        const double primal_norm = alpha * norm_inf_cta(c, s.p(), N);
        const double dual_norm = alpha * norm_inf_cta(c, s.plam(), DUAL);
        for (int i = tid; i < N; i += nt) { const double sp = alpha * s.p()[i]; s.x()[i] += sp; s.step_prev()[i] = sp; }
        for (int i = tid; i < DUAL; i += nt) s.lam()[i] += alpha * s.plam()[i];
        c.sync();
        // ---- termination_criteria_impl (523-529).  x + (alpha p) == (alpha p) + x bit for bit, so after an accepted trial
        // cg already holds the constraint values at the new x
        const double max_viol = max_constraints_violation(c, o, s.x(), s, cg, accepted);
        const bool done = (primal_norm <= st.eps_prim) && (dual_norm <= st.eps_dual) && (max_viol <= st.eps_prim);
        if (tid == 0) {
            s.stats()[0] = cost_step; s.stats()[1] = primal_norm; s.stats()[2] = dual_norm; s.stats()[3] = max_viol;
            if (s.tr_alpha()) s.tr_alpha()[trace_row] = alpha;
            if (s.tr_ls()) s.tr_ls()[trace_row] = trials;
        }
        return done;
    }

    /** SQPBase::solve (sqp_base.hpp:568-696) of one instance: iterate linearise -> QP -> line search / step until the
     *  termination test holds or max_iter QPs were solved.  Lp / vec: QP workspaces (pmb_qp.hpp), scratch: SCRATCH_DOUBLES. */
    /** QPK: 0 = boxADMM (pmb_qp.hpp), 1 = the OSQP-style ADMM<> (pmb_qp_admm.hpp; R then counts the chunks of 2N + M) */
    template <int R, int NW, bool FAST = false, bool GLOBAL_SLOT = false, int QPK = 0>
    PMB_DEV static void solve(Cta& c, const O& o, const SqpInst<O>& s, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst,
                              double* Lp, unsigned char* vec, double* scratch, double* Ld = nullptr)
    {
        if (c.tid() == 0) { s.info()->iter = 1; s.info()->qp_solver_iter = 0; s.info()->status = PMB_SQP_MAX_ITER_EXCEEDED; }
        QpArgs qa;
        qa.N = N; qa.M = M; qa.H = s.H(); qa.h = s.h(); qa.A = s.A(); qa.Alb = s.al(); qa.Aub = s.au(); qa.xlb = s.lx(); qa.xub = s.ux();
        qa.xg = nullptr; qa.yg = nullptr; qa.x = s.p(); qa.y = s.plam(); qa.info = s.qp_info(); qa.z = nullptr; qa.q = nullptr;
        qa.perm = nullptr; qa.ctype = nullptr; qa.nfac = s.qp_nfac();
        QpProf qprof;
        qa.prof = s.phase() ? &qprof : nullptr;
        AdmmArgs aa;
        aa.N = N; aa.M = M; aa.H = s.H(); aa.h = s.h(); aa.Ae = QPK == 1 ? s.Ae() : nullptr; aa.Alb = s.al(); aa.Aub = s.au(); aa.xlb = s.lx(); aa.xub = s.ux();
        aa.xg = nullptr; aa.yg = nullptr; aa.x = s.p(); aa.y = s.plam(); aa.info = s.qp_info(); aa.z = nullptr; aa.perm = nullptr; aa.ctype = nullptr;
        aa.nfac = s.qp_nfac();
        c.sync();
        unsigned long long t_lin = 0, t_qp = 0, t_step = 0, n_it = 0;
        for (int it = 1; ; ++it) {                   // the first iteration always runs (sqp_base.hpp:583-637 precede the loop)
            const int row = it - 1;
            const unsigned long long t0 = c.w.clock();
            const double cost_x = linearise<FAST>(c, o, s, it == 1 || s.ws.opt_exact_hessian != 0, row, scratch);
            const unsigned long long t1 = c.w.clock();
            if (s.ws.opt_precond != PRECOND_IDENTITY)      // m_preconditioner.compute (sqp_base.hpp:605, 662)
                RuizCta<N, M>::compute(c, N, M, s.ws.opt_precond, s.H(), s.h(), s.A(), s.al(), s.au(), s.lx(), s.ux(), s.ruiz(), scratch);
            if (QPK == 1) {
                // ADMM::construct_A (admm.hpp:215-222) on the (scaled) Jacobian of this iteration: Ae = [A; I]
                double* Ae = s.Ae();
                const double* Am = s.A();
                constexpr int Me = M + N;
                for (int e = c.tid(); e < Me * N; e += c.nthreads()) { const int j = e / Me, i = e - j * Me; Ae[e] = i < M ? Am[i + (size_t)j * M] : (i - M == j ? 1.0 : 0.0); }
                c.sync();
                admm_solve_cta<R, NW>(c, qst, aa, Lp, vec);
            } else {
                qp_solve_cta<R, N, M, NW, FAST, GLOBAL_SLOT>(c, qst, qa, Lp, vec, Ld);
            }
            if (s.ws.opt_precond != PRECOND_IDENTITY)      // unscale(p, p_lambda), unscale(H, h, A, ...) (609-611, 666-667)
                RuizCta<N, M>::unscale(c, N, M, s.ruiz(), s.H(), s.h(), s.A(), s.al(), s.au(), s.lx(), s.ux(), s.p(), s.plam());
            const unsigned long long t2 = c.w.clock();
            const bool done = step(c, o, s, st, row, scratch, cost_x);
            const unsigned long long t3 = c.w.clock();
            t_lin += t1 - t0; t_qp += t2 - t1; t_step += t3 - t2; ++n_it;
            if (done) { if (c.tid() == 0) s.info()->status = PMB_SQP_SOLVED; break; }
            if (it >= st.max_iter) break;
            if (c.tid() == 0) s.info()->iter = it + 1;
        }
        if (s.phase() && c.tid() == 0) {
            atomic_add_u64(s.phase() + 0, t_lin); atomic_add_u64(s.phase() + 1, t_qp); atomic_add_u64(s.phase() + 2, t_step); atomic_add_u64(s.phase() + 3, n_it);
            atomic_add_u64(s.phase() + 4, qprof.pivot); atomic_add_u64(s.phase() + 5, qprof.gather); atomic_add_u64(s.phase() + 6, qprof.factor);
            atomic_add_u64(s.phase() + 7, qprof.solve); atomic_add_u64(s.phase() + 8, qprof.update); atomic_add_u64(s.phase() + 9, qprof.resid);
            atomic_add_u64(s.phase() + 10, (unsigned long long)s.info()->qp_solver_iter);
            atomic_add_u64(s.phase() + 12, qprof.f_diag); atomic_add_u64(s.phase() + 13, qprof.f_panel); atomic_add_u64(s.phase() + 14, qprof.f_trail); atomic_add_u64(s.phase() + 15, qprof.f_inv);
        }
        c.sync();
    }
};

} // namespace pmb
