// pmb_qp_admm.hpp — the reference's OSQP-style ADMM<> QP solver on one CTA per instance (stand-alone operator
// pmb_qp_solve_admm; inside the fused SQP loop the QP solver is boxADMM, pmb_qp.hpp).
//
// Reference: src/solvers/admm.hpp — solve_impl 112-213, construct_A 215-222 (Ae = [A; I], done by the host wrapper),
// construct_kkt_matrix 245-259, compute_kkt_rhs 389-393, box_projection 397-403, rho_vec_update 405-444, residuals_update
// 446-468, termination 470-486, estimate_rho 488-496, update_kkt_rho 498-502; bound classification qp_base.hpp:195-222.
//
// The KKT system [[H + sigma I, Ae'], [Ae, -diag(1 / rho)]] has size n = 2N + M; pivot order, gather, factorisation and the
// substitutions are the routines of pmb_qp.hpp (exact arithmetic: every fp64 operation in the order of oracle/admm_qp.hpp).
#pragma once
#include "pmb_qp.hpp"

namespace pmb {

struct AdmmArgs {
    int N, M;
    const double *H, *h, *Ae /* (M + N) x N column-major: [A; I] */, *Alb, *Aub, *xlb, *xub, *xg, *yg;
    double *x, *y;
    pmb_qp_info_t* info;
    double* z;            // M + N, may be null
    int *perm, *ctype, *nfac;
};

/** shared-memory bytes of the vector workspace of one ADMM instance (everything except the packed factor) */
PMB_HD constexpr size_t admm_vec_bytes(int N, int M)
{
    // n = 2N + M: dK, tmp, sol (3 n); x, h (2 N); z, y, rho, 1/rho, lb, ub (6 (M + N)); 12 first coefficients; ints: perm (n), ctype (M + N)
    return (3 * (2 * (size_t)N + M) + 2 * (size_t)N + 6 * ((size_t)N + M) + 12) * sizeof(double) + ((2 * (size_t)N + M) + ((size_t)N + M)) * sizeof(int) + 16;
}

template <int R, int NW = 4>
PMB_DEV void admm_solve_cta(Cta& c, const pmb_qp_settings_t& st, const AdmmArgs& a, double* Lp, unsigned char* vec)
{
    const int N = a.N, M = a.M, Me = N + M, n = N + Me, tid = c.tid(), nt = c.nthreads();
    double* dK = reinterpret_cast<double*>(vec);
    double* tmp = dK + n;
    double* sol = tmp + n;
    double* x = sol + n;
    double* h = x + N;
    double* z = h + N;
    double* y = z + Me;
    double* rv = y + Me;
    double* rvi = rv + Me;
    double* lb = rvi + Me;
    double* ub = lb + Me;
    double* first = ub + Me;
    int* perm = reinterpret_cast<int*>(first + 12);
    int* ctype = perm + n;

    for (int i = tid; i < N; i += nt) { h[i] = a.h[i]; x[i] = a.xg ? a.xg[i] : 0.0; }
    for (int i = tid; i < Me; i += nt) {
        y[i] = a.yg ? a.yg[i] : 0.0;
        const double l = i < M ? a.Alb[i] : a.xlb[i - M], u = i < M ? a.Aub[i] : a.xub[i - M];
        lb[i] = l; ub[i] = u;
        ctype[i] = (l < -qpc::LOOSE_BOUNDS_THRESH && u > qpc::LOOSE_BOUNDS_THRESH) ? PMB_LOOSE_BOUNDS
                   : ((u - l < qpc::EQ_TOL) ? PMB_EQUALITY_CONSTRAINT : PMB_INEQUALITY_CONSTRAINT);
    }
    c.sync();
    for (int i = tid; i < Me; i += nt) z[i] = dot_chain(a.Ae + i, (size_t)Me, x, N);      // m_z = m_A * x_guess (identity rows included)

    int rho_updates = 0, n_factor = 0;
    double rho = 0.0;
    auto rho_vec_update = [&](double rho0) {
        for (int i = tid; i < Me; i += nt) {
            const int t = ctype[i];
            const double r = t == PMB_LOOSE_BOUNDS ? qpc::RHO_MIN : (t == PMB_EQUALITY_CONSTRAINT ? qpc::RHO_EQ_FACTOR * rho0 : rho0);
            rv[i] = r; rvi[i] = 1.0 / r;
        }
        rho = rho0;
        rho_updates += 1;
        c.sync();
    };
    auto factorise = [&]() {
        ldlt_pivot_order<R>(c, n, dK, perm, reinterpret_cast<int*>(tmp));
        if (n_factor == 0 && a.perm) { for (int i = tid; i < n; i += nt) a.perm[i] = perm[i]; }
        kkt_gather_permuted<R>(c, N, Me, a.H, a.Ae, dK, perm, Lp);
        ldlt_factor_packed<R>(c, n, Lp);
        ++n_factor;
    };

    rho_vec_update(st.rho);
    for (int i = tid; i < N; i += nt) { double v = a.H[i + (size_t)i * N]; v += st.sigma; dK[i] = v; }
    for (int i = tid; i < Me; i += nt) dK[N + i] = -1.0 * rvi[i];
    c.sync();

    int status = PMB_QP_UNSOLVED;
    double res_prim = 1.0, res_dual = 1.0, rho_estimate = 0.0, max_Ax_z = 0.0, max_Hx_ATy_h = 0.0;
    const double alpha = st.alpha, sigma = st.sigma;

    auto residuals_update = [&]() {   // admm.hpp:446-468; task t < M: row t of A x, task M + i: row i of H x and A^T y_A
        enum { nAx = 0, nx, nz, nHx, nATy, nh, nyb, rp, rbx, rd, NRED };
        double m[NRED];
        PMB_UNROLL
        for (int k = 0; k < NRED; ++k) m[k] = 0.0;
        if (tid < NRED) first[tid] = 0.0;
        c.sync();
        for (int t = tid; t < Me; t += nt) {
            const double vz = dm::fabs(z[t]);
            if (vz > m[nz]) m[nz] = vz;
            if (t == 0) first[nz] = vz;
            if (t < M) {
                const double acc = dot_chain(a.Ae + t, (size_t)Me, x, N);
                const double v0 = dm::fabs(acc), v2 = dm::fabs(acc - z[t]);
                if (v0 > m[nAx]) m[nAx] = v0;
                if (v2 > m[rp]) m[rp] = v2;
                if (t == 0) { first[nAx] = v0; first[rp] = v2; }
            } else {
                const int i = t - M;
                const double hx = dot_chain(a.H + i, (size_t)N, x, N);
                const double aty = dot_chain(a.Ae + (size_t)i * Me, 1, y, M);
                const double v0 = dm::fabs(x[i]), v1 = dm::fabs(hx), v2 = dm::fabs(aty), v3 = dm::fabs(h[i]), v4 = dm::fabs(y[M + i]);
                const double v5 = dm::fabs(x[i] - z[M + i]), v6 = dm::fabs(((hx + h[i]) + aty) + y[M + i]);
                if (v0 > m[nx]) m[nx] = v0;
                if (v1 > m[nHx]) m[nHx] = v1;
                if (v2 > m[nATy]) m[nATy] = v2;
                if (v3 > m[nh]) m[nh] = v3;
                if (v4 > m[nyb]) m[nyb] = v4;
                if (v5 > m[rbx]) m[rbx] = v5;
                if (v6 > m[rd]) m[rd] = v6;
                if (i == 0) { first[nx] = v0; first[nHx] = v1; first[nATy] = v2; first[nh] = v3; first[nyb] = v4; first[rbx] = v5; first[rd] = v6; }
            }
        }
        c.sync();
        c.max_all<NRED>(m);
        PMB_UNROLL
        for (int k = 0; k < NRED; ++k) { const double f = first[k]; if (f != f) m[k] = f; }
        const double norm_Ax = fmax_nan(M > 0 ? m[nAx] : 0.0, m[nx]);
        max_Ax_z = fmax_nan(norm_Ax, m[nz]);
        max_Hx_ATy_h = fmax_nan(m[nHx], fmax_nan(m[nATy], fmax_nan(m[nh], m[nyb])));
        res_prim = fmax_nan(M > 0 ? m[rp] : 0.0, m[rbx]);
        res_dual = m[rd];
    };

    bool need_factor = true;
    int iter;
    for (iter = 1; ; ++iter) {
        if (need_factor) { factorise(); need_factor = false; }
        if (iter > st.max_iter) break;
        for (int i = tid; i < N; i += nt) sol[i] = sigma * x[i] - h[i];
        for (int i = tid; i < Me; i += nt) sol[N + i] = z[i] - rvi[i] * y[i];
        c.sync();
        ldlt_solve_packed<R, NW>(c, n, Lp, perm, sol, tmp);
        for (int i = tid; i < N; i += nt) x[i] = (alpha * sol[i]) + ((1 - alpha) * x[i]);
        for (int i = tid; i < Me; i += nt) {
            const double zp = z[i];
            const double zt = zp + rvi[i] * (sol[N + i] - y[i]);
            double v = alpha * zt;
            v += ((1 - alpha) * zp) + (rvi[i] * y[i]);
            const double zn = dm::min(dm::max(v, lb[i]), ub[i]);
            y[i] += rv[i] * (((alpha * zt) + ((1 - alpha) * zp)) - zn);
            z[i] = zn;
        }
        c.sync();

        const bool check = (st.check_termination != 0) && (iter % st.check_termination == 0);
        const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0);
        if (check || adapt) residuals_update();
        if (check) {
            const double eps_prim = st.eps_abs + st.eps_rel * max_Ax_z;
            const double eps_dual = st.eps_abs + st.eps_rel * max_Hx_ATy_h;
            if (res_prim <= eps_prim && res_dual <= eps_dual) { status = PMB_QP_SOLVED; break; }
        }
        if (adapt) {
            const double rp_norm = res_prim / (max_Ax_z + qpc::DIV_BY_ZERO_REGUL);
            const double rd_norm = res_dual / (max_Hx_ATy_h + qpc::DIV_BY_ZERO_REGUL);
            double new_rho = rho * dm::sqrt(rp_norm / (rd_norm + qpc::DIV_BY_ZERO_REGUL));
            new_rho = fmax_nan(qpc::RHO_MIN, fmin_nan(new_rho, qpc::RHO_MAX));
            rho_estimate = new_rho;
            if (new_rho < rho / st.adaptive_rho_tolerance || new_rho > rho * st.adaptive_rho_tolerance) {
                rho_vec_update(new_rho);
                for (int i = tid; i < Me; i += nt) dK[N + i] = -rvi[i];     // update_kkt_rho (498-502)
                c.sync();
                need_factor = true;
            }
        }
    }
    if (iter > st.max_iter) status = PMB_QP_MAX_ITER_EXCEEDED;

    for (int i = tid; i < N; i += nt) a.x[i] = x[i];
    for (int i = tid; i < Me; i += nt) { a.y[i] = y[i]; if (a.z) a.z[i] = z[i]; if (a.ctype) a.ctype[i] = ctype[i]; }
    if (tid == 0) {
        if (a.info) {
            a.info->status = status; a.info->iter = iter; a.info->rho_updates = rho_updates; a.info->_pad = 0;
            a.info->rho_estimate = rho_estimate; a.info->res_prim = res_prim; a.info->res_dual = res_dual;
        }
        if (a.nfac) *a.nfac = n_factor;
    }
    c.sync();
}

} // namespace pmb
