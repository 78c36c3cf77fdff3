// pmb_qp.hpp — batched boxADMM with a dense pivoted LDL^T: one CTA per QP instance, KKT factor in shared memory.
//
// Reference: src/solvers/box_admm.hpp (solve_impl 88-205, construct_kkt_matrix 207-223, factorise 335-341, compute_kkt_rhs
// 351-355, rho_vec_update 357-396, residuals_update 398-415, termination 417-431, estimate_rho 433-445, update_kkt_rho
// 447-452), src/solvers/qp_base.hpp (settings 17-53, parse_constraints_bounds 195-222) and Eigen::LDLT<Lower>
// (src/utils/helpers.hpp:38-43 selects it; Eigen itself is not in the reference tree).
//
// B200 mapping.  K = [[H + sigma I + diag(rho_box), .],[A, -diag(1/rho_A)]] is (N+M)^2; only its lower triangle is
// ever read, so the CTA keeps the packed lower triangle (n(n+1)/2 doubles: 43.7 KB for the mobile robot) in shared
// memory and factors it in place; K is never written to HBM.  Problems whose factor does not fit (kite 12x1: 570 KB)
// keep it in a per-CTA global scratch slot that stays L2 resident.
//   * Pivoting.  Eigen's LDLT picks, at step k, the largest |diagonal| of the *not yet updated* trailing diagonal (its
//     in-place algorithm is left-looking), i.e. the pivot order depends only on diag(K): warp 0 replays that selection on
//     the diagonal alone, the block gathers P K P^T straight into the packed layout, and then runs an unpivoted
//     right-looking LDL^T whose per-element update order (ascending elimination step, fused multiply-add) is identical
//     to the left-looking inner products.
//   * Factorisation: columns of the trailing matrix are dealt round-robin to the warps, rows to the lanes.
//   * Triangular solves are a dependent chain of n steps: warp 0 keeps the right-hand side in registers (lane l owns
//     rows l, l+32, ...) and broadcasts each finished component with one shuffle; the other warps wait at the barrier.
//   * ADMM vector updates are strided over the block; the residual mat-vecs (every check_termination trips) give every
//     thread one row of A x, H x or A^T y and stream H and A from global memory (L2 resident).
#pragma once
#include "pmb_cta.hpp"
#include "../../include/polympc_b200.h"
#include <cfloat>

// micro-benchmark hook (tools/ubench): expands to nothing in the product build
#ifndef PMB_TICK
#define PMB_TICK(k)
#endif

namespace pmb {

/** optional cycle counters of one QP solve (thread 0's clock): pivot order, gather, factorisation, triangular solves,
 *  ADMM vector updates, residual checks */
struct QpProf { unsigned long long pivot = 0, gather = 0, factor = 0, solve = 0, update = 0, resid = 0; };

struct QpArgs {
    int N, M;
    const double *H, *h, *A, *Alb, *Aub, *xlb, *xub, *xg, *yg;   // this instance
    double *x, *y;
    pmb_qp_info_t* info;
    double *z, *q;
    int *perm, *ctype, *nfac;
    QpProf* prof;   // thread-local accumulator or nullptr
};

namespace qpc {
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e+6, RHO_EQ_FACTOR = 1e+3;
constexpr double LOOSE_BOUNDS_THRESH = 1e+10, EQ_TOL = 1e-4, DIV_BY_ZERO_REGUL = 10e-10;
}

/** doubles of the packed factor */
inline size_t qp_factor_doubles(int N, int M) { const size_t n = (size_t)N + M; return n * (n + 1) / 2; }
/** shared-memory bytes of the vector workspace of one QP instance (everything except the packed factor) */
inline size_t qp_vec_bytes(int N, int M)
{
    const size_t n = (size_t)N + M;
    const size_t doubles = 3 * n + 6 * (size_t)N + 4 * (size_t)M;
    return doubles * sizeof(double) + 2 * n * sizeof(int) + 16;
}

PMB_DEV double fmax_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return (a < b) ? b : a; }
PMB_DEV double fmin_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return (b < a) ? b : a; }
PMB_DEV double warp_max(const Warp& w, double m)
{
    for (int off = 16; off >= 1; off >>= 1) { const double o = w.shfl_xor(m, off); if (o > m) m = o; }
    return m;
}

PMB_DEV int packed_off(int j, int n) { return j * n - (j * (j - 1)) / 2; }

/** sum_j a[j*ld] * x[j], sequential ascending fused chain; loads are issued eight at a time ahead of the chain */
PMB_DEV double dot_chain(const double* a, size_t ld, const double* x, int n)
{
    double acc = 0.0;
    int j = 0;
    for (; j + 8 <= n; j += 8) {
        double v[8];
        PMB_UNROLL
        for (int u = 0; u < 8; ++u) v[u] = a[(size_t)(j + u) * ld];
        PMB_UNROLL
        for (int u = 0; u < 8; ++u) acc = dm::fma(v[u], x[j + u], acc);
    }
    for (; j < n; ++j) acc = dm::fma(a[(size_t)j * ld], x[j], acc);
    return acc;
}

/** replay Eigen's diagonal pivot selection (LDLT.h, ldlt_inplace<Lower>::unblocked: `mat.diagonal().tail(size-k).cwiseAbs()
 *  .maxCoeff(&idx)` then a symmetric swap k <-> idx) on dd, a scratch copy of diag(K); perm[a] = original index at position a.
 *  maxCoeff semantics: the candidate at position k is the initial best (even when it is NaN — then nothing replaces it),
 *  a later position wins only with a strictly larger value (NaN never does).
 *  Executed by warp 0 on registers (lane l owns positions l, l+32, ...): one step = 3 REDUX + 4 SHFL.  Ends with a block
 *  barrier. */
template <int R>
PMB_DEV void ldlt_pivot_order(Cta& c, int n, const double* dd, int* perm)
{
    if (c.warp_id() == 0) {
        const Warp& w = c.w;
        const int lane = w.lane();
        // key: 0 = not a candidate (NaN or out of range); otherwise bits(|v|) + 1, order preserving for |v| in [0, inf]
        unsigned hi[R], lo[R];
        int pr[R];
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            uint64_t key = 0;
            if (i < n) { const double v = dm::fabs(dd[i]); key = (v != v) ? 0 : dm::to_bits(v) + 1; }
            hi[r] = (unsigned)(key >> 32); lo[r] = (unsigned)key; pr[r] = i;
        }
        for (int k = 0; k < n - 1; ++k) {
            const int kr = k >> 5, kl = k & 31;
            // local best over owned positions >= k (ascending positions: first maximum wins)
            unsigned bh = 0, bl = 0; int bp = 0x7fffffff;
            PMB_UNROLL
            for (int r = 0; r < R; ++r) {
                const int i = lane + 32 * r;
                const bool cand = i >= k && (hi[r] | lo[r]) != 0;
                if (cand && (hi[r] > bh || (hi[r] == bh && lo[r] > bl))) { bh = hi[r]; bl = lo[r]; bp = i; }
            }
            const unsigned mh = w.reduce_max(bh);
            const unsigned ml = w.reduce_max(bh == mh ? bl : 0u);
            const bool mine = (bh == mh) && (bl == ml) && bp != 0x7fffffff;
            int bi = (int)w.reduce_min(mine ? (unsigned)bp : 0x7fffffffu);
            // key / perm currently at position k
            unsigned kh = 0, kw = 0; int kp = 0;
            PMB_UNROLL
            for (int r = 0; r < R; ++r) if (r == kr) { kh = hi[r]; kw = lo[r]; kp = pr[r]; }
            kh = (unsigned)w.shfl((int)kh, kl); kw = (unsigned)w.shfl((int)kw, kl); kp = w.shfl(kp, kl);
            if ((kh | kw) == 0 || (mh | ml) == 0) bi = k;          // NaN at k stays; nothing selectable: no swap
            if (bi != k) {
                const int br = bi >> 5, bl2 = bi & 31;
                int bperm = 0;
                PMB_UNROLL
                for (int r = 0; r < R; ++r) if (r == br) bperm = pr[r];
                bperm = w.shfl(bperm, bl2);
                PMB_UNROLL
                for (int r = 0; r < R; ++r) {
                    if (r == kr && lane == kl) { hi[r] = mh; lo[r] = ml; pr[r] = bperm; }
                    if (r == br && lane == bl2) { hi[r] = kh; lo[r] = kw; pr[r] = kp; }
                }
            }
        }
        PMB_UNROLL
        for (int r = 0; r < R; ++r) { const int i = lane + 32 * r; if (i < n) perm[i] = pr[r]; }
    }
    c.sync();
}

/** gather P K P^T into the packed lower triangle: columns round-robin over warps, rows over lanes */
PMB_DEV void kkt_gather_permuted(Cta& c, int N, int M, const double* H, const double* A, const double* dK, const int* perm, double* Lp)
{
    const int n = N + M, lane = c.lane(), nw = c.nwarps();
    for (int b = c.warp_id(); b < n; b += nw) {
        const int cc = perm[b];
        double* col = Lp + packed_off(b, n) - b;
        for (int a = b + lane; a < n; a += 32) {
            const int r = perm[a];
            const int hi = r > cc ? r : cc, lo = r > cc ? cc : r;
            double v;
            if (hi == lo) v = dK[hi];
            else if (hi < N) v = H[hi + (size_t)lo * N];
            else if (lo < N) v = A[(hi - N) + (size_t)lo * M];
            else v = 0.0;
            col[a] = v;
        }
    }
    c.sync();
}

/** unpivoted right-looking LDL^T on the packed lower triangle; tmp[n] scratch.  R = ceil(n / 32).
 *  Lane l owns rows l, l+32, ... (its multipliers L(i,j) stay in registers for the whole step); the columns k > j of the
 *  trailing matrix are dealt round-robin to the warps. */
template <int R>
PMB_DEV void ldlt_factor_packed(Cta& c, int n, double* Lp, double* tmp)
{
    const int tid = c.tid(), nt = c.nthreads(), lane = c.lane(), wid = c.warp_id(), nw = c.nwarps();
    int bj = 0;                                    // packed_off(j, n) - j
    for (int j = 0; j < n; ++j) {
        double* cj = Lp + bj;                      // cj[i] = L(i,j), i >= j
        const double dj = cj[j];
        const bool scale = dm::fabs(dj) > 0.0;
        for (int i = j + 1 + tid; i < n; i += nt) {
            double v = cj[i];
            if (scale) { v = v / dj; cj[i] = v; }
            tmp[i] = dj * v;
        }
        c.sync();
        // a(i,k) = fma(-L(i,j), tmp[k], a(i,k)) for j < k <= i < n
        double nl[R];
        PMB_UNROLL
        for (int r = 0; r < R; ++r) { const int i = lane + 32 * r; nl[r] = (i > j && i < n) ? -cj[i] : 0.0; }
        int k = j + 1 + wid;
        int bk = k * n - ((k * (k + 1)) >> 1);     // packed_off(k, n) - k
        for (; k < n; k += nw) {
            const double tk = tmp[k];
            double* ck = Lp + bk + lane;
            PMB_UNROLL
            for (int r = 0; r < R; ++r) {
                if (32 * r + 31 >= k) {            // warp-uniform: chunk has rows >= k
                    const int i = lane + 32 * r;
                    if (i >= k && i < n) ck[32 * r] = dm::fma(nl[r], tk, ck[32 * r]);
                }
            }
            // advance nw columns: off(k+1) - off(k) = n - k - 1
            PMB_UNROLL
            for (int q = 0; q < 8; ++q) if (q < nw) bk += n - (k + q) - 1;
        }
        c.sync();
        bj += n - j - 1;
    }
}

/** solve (P^T L D L^T P) s = rhs; `sol` holds rhs on entry and the solution on exit (unpermuted indexing).
 *  Executed by warp 0: lane l keeps components l, l+32, ... in registers; step j broadcasts the finished component with
 *  one shuffle and applies column j (forward) / row j (backward) of L.  Ends with a block barrier. */
template <int R>
PMB_DEV void ldlt_solve_packed(Cta& c, int n, const double* Lp, const int* perm, double* sol)
{
    if (c.warp_id() == 0) {
        const Warp& w = c.w;
        const int lane = w.lane();
        double y[R];
        int offr[R];      // packed_off(i, n) - i of the lane's rows
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            y[r] = i < n ? sol[perm[i]] : 0.0;
            offr[r] = i < n ? i * n - ((i * (i + 1)) >> 1) : 0;
        }
        PMB_TICK(0)
        // unit lower: ascending columns.  col_j[i] = Lp[bj + i], bj = packed_off(j) - j advances by n - j - 1
        {
            const double* colp = Lp + lane;
            PMB_UNROLL
            for (int jb = 0; jb < R; ++jb) {
                const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
                PMB_NOUNROLL
                for (int jj = 0; jj < jend; ++jj) {
                    const int j = jb * 32 + jj;
                    const double yj = w.shfl(y[jb], jj);
                    if (lane > jj && 32 * jb + lane < n) y[jb] = dm::fma(-colp[32 * jb], yj, y[jb]);
                    PMB_UNROLL
                    for (int r = jb + 1; r < R; ++r) {
                        if (r < R - 1 || lane + 32 * r < n) y[r] = dm::fma(-colp[32 * r], yj, y[r]);
                    }
                    colp += n - j - 1;
                }
            }
        }
        PMB_TICK(1)
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            if (i < n) {
                const double di = Lp[offr[r] + i];
                y[r] = (dm::fabs(di) > DBL_MIN) ? (y[r] / di) : 0.0;
            }
        }
        PMB_TICK(2)
        // unit upper (L^T): descending columns; lane's row i reads L(j,i) = Lp[offr + j]
        PMB_UNROLL
        for (int jb = R - 1; jb >= 0; --jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
            PMB_NOUNROLL
            for (int jj = jend - 1; jj >= 0; --jj) {
                const int j = jb * 32 + jj;
                const double yj = w.shfl(y[jb], jj);
                PMB_UNROLL
                for (int r = 0; r < jb; ++r) y[r] = dm::fma(-Lp[offr[r] + j], yj, y[r]);
                if (lane < jj) y[jb] = dm::fma(-Lp[offr[jb] + j], yj, y[jb]);
            }
        }
        PMB_TICK(3)
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            if (i < n) sol[perm[i]] = y[r];
        }
    }
    c.sync();
    PMB_TICK(4)
}

/** the whole boxADMM solve of one instance by one CTA.  Lp: n(n+1)/2 doubles (shared or global), vec: qp_vec_bytes() of
 *  shared memory. */
template <int R>
PMB_DEV void qp_solve_cta(Cta& c, const pmb_qp_settings_t& st, const QpArgs& a, double* Lp, unsigned char* vec)
{
    const int N = a.N, M = a.M, n = N + M, tid = c.tid(), nt = c.nthreads();
    double* dK = reinterpret_cast<double*>(vec);
    double* tmp = dK + n;
    double* sol = tmp + n;
    double* x = sol + n;
    double* q = x + N;
    double* yb = q + N;
    double* h = yb + N;
    double* rb = h + N;
    double* rbi = rb + N;
    double* z = rbi + N;
    double* ya = z + M;
    double* rv = ya + M;
    double* rvi = rv + M;
    int* perm = reinterpret_cast<int*>(rvi + M);
    int* ctype = perm + n;     // [constr_type (M) ; box_constr_type (N)]
    const double *alb = a.Alb, *aub = a.Aub, *xlb = a.xlb, *xub = a.xub;   // bounds stay in global memory (read-only here)

    // ---- load, initial iterates (box_admm.hpp:97-100) -----------------------------------------------------------
    for (int i = tid; i < N; i += nt) {
        h[i] = a.h[i];
        x[i] = a.xg ? a.xg[i] : 0.0;
        q[i] = x[i];
        yb[i] = a.yg ? a.yg[M + i] : 0.0;
    }
    for (int i = tid; i < M; i += nt) ya[i] = a.yg ? a.yg[i] : 0.0;
    c.sync();
    for (int i = tid; i < M; i += nt) z[i] = a.xg ? dot_chain(a.A + i, (size_t)M, x, N) : 0.0;

    // ---- parse_constraints_bounds (qp_base.hpp:195-222) ------------------------------------------------------------
    for (int i = tid; i < M; i += nt)
        ctype[i] = (alb[i] < -qpc::LOOSE_BOUNDS_THRESH && aub[i] > qpc::LOOSE_BOUNDS_THRESH) ? PMB_LOOSE_BOUNDS
                   : ((aub[i] - alb[i] < qpc::EQ_TOL) ? PMB_EQUALITY_CONSTRAINT : PMB_INEQUALITY_CONSTRAINT);
    for (int i = tid; i < N; i += nt)
        ctype[M + i] = (xlb[i] < -qpc::LOOSE_BOUNDS_THRESH && xub[i] > qpc::LOOSE_BOUNDS_THRESH) ? PMB_LOOSE_BOUNDS
                       : ((xub[i] - xlb[i] < qpc::EQ_TOL) ? PMB_EQUALITY_CONSTRAINT : PMB_INEQUALITY_CONSTRAINT);
    c.sync();

    int rho_updates = 0, n_factor = 0;
    double rho = 0.0;
    auto rho_vec_update = [&](double rho0) {   // box_admm.hpp:357-396
        for (int i = tid; i < M; i += nt) {
            const int t = ctype[i];
            const double r = t == PMB_LOOSE_BOUNDS ? qpc::RHO_MIN : (t == PMB_EQUALITY_CONSTRAINT ? qpc::RHO_EQ_FACTOR * rho0 : rho0);
            rv[i] = r; rvi[i] = 1.0 / r;
        }
        for (int i = tid; i < N; i += nt) {
            const int t = ctype[M + i];
            const double r = t == PMB_LOOSE_BOUNDS ? qpc::RHO_MIN : (t == PMB_EQUALITY_CONSTRAINT ? qpc::RHO_EQ_FACTOR * rho0 : rho0);
            rb[i] = r; rbi[i] = 1.0 / r;
        }
        rho = rho0;
        rho_updates += 1;
        c.sync();
    };
    QpProf* const prof = a.prof;
    auto factorise = [&]() {
        const unsigned long long t0 = prof ? c.w.clock() : 0;
        ldlt_pivot_order<R>(c, n, dK, perm);
        if (n_factor == 0 && a.perm) { for (int i = tid; i < n; i += nt) a.perm[i] = perm[i]; }
        const unsigned long long t1 = prof ? c.w.clock() : 0;
        kkt_gather_permuted(c, N, M, a.H, a.A, dK, perm, Lp);
        const unsigned long long t2 = prof ? c.w.clock() : 0;
        ldlt_factor_packed<R>(c, n, Lp, tmp);
        if (prof) { const unsigned long long t3 = c.w.clock(); prof->pivot += t1 - t0; prof->gather += t2 - t1; prof->factor += t3 - t2; }
        ++n_factor;
    };

    rho_vec_update(st.rho);
    // diag of construct_kkt_matrix (box_admm.hpp:207-223): (H_ii + sigma) + rho_box_i ; -1/rho_A
    for (int i = tid; i < N; i += nt) { double v = a.H[i + (size_t)i * N]; v += st.sigma; v += rb[i]; dK[i] = v; }
    for (int i = tid; i < M; i += nt) dK[N + i] = -rvi[i];
    c.sync();
    factorise();

    int status = PMB_QP_UNSOLVED;
    double res_prim = 1.0, res_dual = 1.0, rho_estimate = 0.0, max_Ax_z = 0.0, max_Hx_ATy_h = 0.0;
    const double alpha = st.alpha, sigma = st.sigma;

    auto residuals_update = [&]() {   // box_admm.hpp:398-415; task t < M: row t of A x, task M + i: row i of H x and A^T y
        enum { nAx = 0, nz, nx, nHx, nATy, nh, nyb, rp, rq, rd, NRED };
        double m[NRED];
        PMB_UNROLL
        for (int k = 0; k < NRED; ++k) m[k] = 0.0;
        for (int t = tid; t < M + N; t += nt) {
            if (t < M) {
                const int i = t;
                const double acc = dot_chain(a.A + i, (size_t)M, x, N);
                double v = dm::fabs(acc); if (v > m[nAx]) m[nAx] = v;
                v = dm::fabs(z[i]); if (v > m[nz]) m[nz] = v;
                v = dm::fabs(acc - z[i]); if (v > m[rp]) m[rp] = v;
            } else {
                const int i = t - M;
                const double hx = dot_chain(a.H + i, (size_t)N, x, N);
                const double aty = dot_chain(a.A + (size_t)i * M, 1, ya, M);
                double v = dm::fabs(x[i]); if (v > m[nx]) m[nx] = v;
                v = dm::fabs(hx); if (v > m[nHx]) m[nHx] = v;
                v = dm::fabs(aty); if (v > m[nATy]) m[nATy] = v;
                v = dm::fabs(h[i]); if (v > m[nh]) m[nh] = v;
                v = dm::fabs(yb[i]); if (v > m[nyb]) m[nyb] = v;
                v = dm::fabs(x[i] - q[i]); if (v > m[rq]) m[rq] = v;
                v = dm::fabs(((hx + h[i]) + aty) + yb[i]); if (v > m[rd]) m[rd] = v;
            }
        }
        c.max_all<NRED>(m);
        max_Ax_z = fmax_nan(m[nAx], fmax_nan(m[nz], m[nx]));
        max_Hx_ATy_h = fmax_nan(m[nHx], fmax_nan(m[nATy], fmax_nan(m[nh], m[nyb])));
        res_prim = m[rp] + m[rq];
        res_dual = m[rd];
    };

    int iter;
    for (iter = 1; iter <= st.max_iter; ++iter) {
        const unsigned long long ta = prof ? c.w.clock() : 0;
        // compute_kkt_rhs (351-355)
        for (int i = tid; i < N; i += nt) sol[i] = ((sigma * x[i] - h[i]) + rb[i] * q[i]) - yb[i];
        for (int i = tid; i < M; i += nt) sol[N + i] = z[i] - rvi[i] * ya[i];
        c.sync();
        const unsigned long long tb = prof ? c.w.clock() : 0;
        ldlt_solve_packed<R>(c, n, Lp, perm, sol);
        const unsigned long long tc = prof ? c.w.clock() : 0;
        // z, y_A (126, 133-135, 142-144)
        for (int i = tid; i < M; i += nt) {
            const double zp = z[i];
            const double zt = zp + rvi[i] * (sol[N + i] - ya[i]);
            double v = alpha * zt;
            v += ((1 - alpha) * zp) + (rvi[i] * ya[i]);
            const double zn = dm::min(dm::max(v, alb[i]), aub[i]);
            ya[i] += rv[i] * (((alpha * zt) + ((1 - alpha) * zp)) - zn);
            z[i] = zn;
        }
        // x, q, y_box (129-130, 138-139, 146-147)
        for (int i = tid; i < N; i += nt) {
            double xv = alpha * sol[i];
            xv += (1 - alpha) * xv;
            x[i] = xv;
            const double qv = dm::min(dm::max(xv + rbi[i] * yb[i], xlb[i]), xub[i]);
            q[i] = qv;
            yb[i] += rb[i] * (xv - qv);
        }
        c.sync();
        if (prof) { const unsigned long long td = c.w.clock(); prof->update += (tb - ta) + (td - tc); prof->solve += tc - tb; }

        const bool check = (st.check_termination != 0) && (iter % st.check_termination == 0);
        if (check) {
            const unsigned long long te = prof ? c.w.clock() : 0;
            residuals_update();
            if (prof) prof->resid += c.w.clock() - te;
            const double eps_prim = st.eps_abs + st.eps_rel * max_Ax_z;
            const double eps_dual = st.eps_abs + st.eps_rel * max_Hx_ATy_h;
            if (res_prim <= eps_prim && res_dual <= eps_dual) { status = PMB_QP_SOLVED; break; }
        }
        if (st.adaptive_rho && st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0)) {
            if (!check) residuals_update();
            const double rp_norm = res_prim / (max_Ax_z + qpc::DIV_BY_ZERO_REGUL);
            const double rd_norm = res_dual / (max_Hx_ATy_h + qpc::DIV_BY_ZERO_REGUL);
            double new_rho = rho * dm::sqrt(rp_norm / (rd_norm + qpc::DIV_BY_ZERO_REGUL));
            new_rho = fmax_nan(qpc::RHO_MIN, fmin_nan(new_rho, qpc::RHO_MAX));
            rho_estimate = new_rho;
            if (new_rho < rho / st.adaptive_rho_tolerance || new_rho > rho * st.adaptive_rho_tolerance) {
                for (int i = tid; i < N; i += nt) tmp[i] = rb[i];   // rho_box_prev
                c.sync();
                rho_vec_update(new_rho);
                // update_kkt_rho (447-452)
                for (int i = tid; i < N; i += nt) dK[i] += (rb[i] - tmp[i]);
                for (int i = tid; i < M; i += nt) dK[N + i] = -rvi[i];
                c.sync();
                factorise();
            }
        }
    }
    if (iter > st.max_iter) status = PMB_QP_MAX_ITER_EXCEEDED;

    for (int i = tid; i < N; i += nt) { a.x[i] = x[i]; a.y[M + i] = yb[i]; if (a.q) a.q[i] = q[i]; }
    for (int i = tid; i < M; i += nt) { a.y[i] = ya[i]; if (a.z) a.z[i] = z[i]; }
    if (a.ctype) for (int i = tid; i < n; i += nt) a.ctype[i] = ctype[i];
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
