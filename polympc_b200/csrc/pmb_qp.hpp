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
struct QpProf { unsigned long long pivot = 0, gather = 0, factor = 0, solve = 0, update = 0, resid = 0, f_diag = 0, f_panel = 0, f_trail = 0, f_inv = 0; };

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
PMB_HD constexpr size_t qp_vec_bytes(int N, int M)
{
    // 3 n + 6 N + 4 M doubles, + 12: first coefficients of the 10 residual norms; 3 n ints: perm, ctype, inverse perm
    return (3 * ((size_t)N + M) + 6 * (size_t)N + 4 * (size_t)M + 12) * sizeof(double) + 3 * ((size_t)N + M) * sizeof(int) + 16;
}

PMB_DEV double fmax_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return (a < b) ? b : a; }
PMB_DEV double fmin_nan(double a, double b) { if (a != a) return b; if (b != b) return a; return (b < a) ? b : a; }
PMB_DEV double warp_max(const Warp& w, double m)
{
    for (int off = 16; off >= 1; off >>= 1) { const double o = w.shfl_xor(m, off); if (o > m) m = o; }
    return m;
}

PMB_DEV int packed_off(int j, int n) { return j * n - (j * (j - 1)) / 2; }

/** sum_j a[j*ld] * 0.0 accumulated from +0.0: what `A * x_guess` (box_admm.hpp:99) gives for the default zero guess — +0.0
 *  for a row of finite entries (every product is +-0 and (+0) + (-0) = +0 in any order), NaN as soon as the row holds a NaN
 *  or an infinity.  Order-free, so the dependent chain of dot_chain() is not needed; loads go out eight at a time. */
PMB_DEV double zero_guess_row(const double* a, size_t ld, int n)
{
    double acc0 = 0.0, acc1 = 0.0;
    int j = 0;
    for (; j + 8 <= n; j += 8) {
        double v[8];
        PMB_UNROLL
        for (int u = 0; u < 8; ++u) v[u] = a[(size_t)(j + u) * ld];
        PMB_UNROLL
        for (int u = 0; u < 8; u += 2) { acc0 = dm::fma(v[u], 0.0, acc0); acc1 = dm::fma(v[u + 1], 0.0, acc1); }
    }
    for (; j < n; ++j) acc0 = dm::fma(a[(size_t)j * ld], 0.0, acc0);
    return acc0 + acc1;
}

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
 *  .maxCoeff(&idx)` then a symmetric swap k <-> idx) on dd = diag(K); perm[a] = original index at position a.
 *  maxCoeff semantics: the candidate at position k is the initial best (even when it is NaN — then nothing replaces it),
 *  a later position wins only with a strictly larger value (NaN never does).
 *
 *  Only the ORDER of the |dd| matters, so the block first ranks them (thread i counts the entries smaller than |dd_i|: ties
 *  share a rank, NaN gets rank 0 = never a candidate) and warp 0 then replays the selection on one 32-bit word per position,
 *  (rank << 18) | ((511 - position) << 9) | original index: a single REDUX.max per step yields the largest value, the first
 *  position holding it AND the original index sitting there; one shuffle moves the element displaced from position k.
 *  ~50 cycles per step instead of ~660 (3 REDUX on 64-bit keys + 4 SHFL).  rk: scratch of n ints.  Ends with a block barrier. */
template <int R>
PMB_DEV void ldlt_pivot_order(Cta& c, int n, const double* dd, int* perm, int* rk)
{
    static_assert(R <= 16, "positions and indices are packed in 9 bits");
    for (int i = c.tid(); i < n; i += c.nthreads()) {
        const double v = dm::fabs(dd[i]);
        int cnt = 1;
        if (v != v) cnt = 0;
        else for (int j = 0; j < n; ++j) { const double u = dm::fabs(dd[j]); cnt += (u < v) ? 1 : 0; }
        rk[i] = cnt;
    }
    c.sync();
    if (c.warp_id() == 0) {
        const Warp& w = c.w;
        const int lane = w.lane();
        unsigned pk[R];                                        // packed word of the element currently at position lane + 32 r
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            pk[r] = i < n ? (((unsigned)rk[i] << 18) | ((unsigned)(511 - i) << 9) | (unsigned)i) : 0u;
        }
        for (int k = 0; k < n - 1; ++k) {
            const int kr = k >> 5, kl = k & 31;
            unsigned best = 0;
            PMB_UNROLL
            for (int r = 0; r < R; ++r) {
                const bool cand = (lane + 32 * r) >= k && (pk[r] >> 18) != 0;
                const unsigned v = cand ? pk[r] : 0u;
                best = v > best ? v : best;
            }
            unsigned atk = 0;
            PMB_UNROLL
            for (int r = 0; r < R; ++r) if (r == kr) atk = pk[r];
            atk = (unsigned)w.shfl((int)atk, kl);              // the element at position k (independent of the reduction)
            const unsigned m = w.reduce_max(best);
            const bool stay = (atk >> 18) == 0 || m == 0;      // NaN at k stays; nothing selectable: no swap
            const int bi = stay ? k : 511 - (int)((m >> 9) & 511u);
            const unsigned chosen = stay ? atk : m;
            if (bi != k) {
                const int br = bi >> 5, bl = bi & 31;
                PMB_UNROLL
                for (int r = 0; r < R; ++r) {
                    if (r == kr && lane == kl) pk[r] = (chosen & ~(511u << 9)) | ((unsigned)(511 - k) << 9);
                    if (r == br && lane == bl) pk[r] = (atk & ~(511u << 9)) | ((unsigned)(511 - bi) << 9);
                }
            }
        }
        PMB_UNROLL
        for (int r = 0; r < R; ++r) { const int i = lane + 32 * r; if (i < n) perm[i] = (int)(pk[r] & 511u); }
    }
    c.sync();
}

/** gather P K P^T into the packed lower triangle: columns round-robin over warps, rows over lanes; the R loads of a lane are
 *  independent (memory-level parallelism: H and A come from L2) */
template <int R>
PMB_DEV void kkt_gather_permuted(Cta& c, int N, int M, const double* H, const double* A, const double* dK, const int* perm, double* Lp)
{
    const int n = N + M, lane = c.lane(), nw = c.nwarps();
    int pa[R];
    PMB_UNROLL
    for (int r = 0; r < R; ++r) { const int a = lane + 32 * r; pa[r] = a < n ? perm[a] : 0; }
    PMB_NOUNROLL
    for (int b = c.warp_id(); b < n; b += nw) {
        const int cc = perm[b];
        double* col = Lp + packed_off(b, n) - b;
        double v[R];
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int a = lane + 32 * r;
            double x = 0.0;
            if (a >= b && a < n) {
                const int rr = pa[r];
                const int hi = rr > cc ? rr : cc, lo = rr > cc ? cc : rr;
                if (hi == lo) x = dK[hi];
                else if (hi < N) x = H[hi + (size_t)lo * N];
                else if (lo < N) x = A[(hi - N) + (size_t)lo * M];
            }
            v[r] = x;
        }
        PMB_UNROLL
        for (int r = 0; r < R; ++r) { const int a = lane + 32 * r; if (a >= b && a < n) col[a] = v[r]; }
    }
    c.sync();
}

/** unpivoted right-looking LDL^T on the packed lower triangle, four columns (a panel) at a time.  R = ceil(n / 32).
 *
 *  Element-wise the arithmetic is exactly the column-by-column algorithm (and Eigen's left-looking inner products):
 *  a(i,k) receives fma(-L(i,j), d_j L(k,j), .) for j ascending, L(i,j) = a(i,j) / d_j (only when |d_j| > 0).  The panel
 *  form only changes the schedule: (P1) every thread factors the 4x4 diagonal block of the panel redundantly in
 *  registers (no communication) and then eliminates its own row of the panel — 4 dependent divisions per row instead of
 *  a barrier pair per column; (P2) the trailing matrix is updated once per panel: lane l keeps the 4 multipliers of its
 *  rows l, l+32, ... in registers, the columns are dealt round-robin to the warps, and every element is read and written
 *  once per panel (4 chained FMAs in registers) instead of once per column.  Two block barriers per panel. */
template <int R>
PMB_DEV void ldlt_factor_packed(Cta& c, int n, double* Lp)
{
    const int tid = c.tid(), nt = c.nthreads(), lane = c.lane(), wid = c.warp_id(), nw = c.nwarps();
    PMB_NOUNROLL
    for (int j0 = 0; j0 < n; j0 += 4) {
        const int nb = (n - j0) < 4 ? (n - j0) : 4;            // columns in this panel (uniform)
        int bc[4];                                             // packed_off(j0 + c, n) - (j0 + c)
        bc[0] = j0 * n - ((j0 * (j0 + 1)) >> 1);
        PMB_UNROLL
        for (int q = 1; q < 4; ++q) bc[q] = bc[q - 1] + n - (j0 + q - 1) - 1;
        // ---- P1: the 4x4 diagonal block (redundantly in every thread, b[r][q] = a(j0 + r, j0 + q), q <= r) and the thread's
        // own row below the panel are eliminated together, straight-line and branch-free, so that the divisions of one
        // elimination step (3 + 1, 2 + 1, 1 + 1, 1) are independent and overlap: 4 division latencies per panel.
        double b[4][4], d[4], ub[4][4];
        bool sc[4];
        PMB_UNROLL
        for (int q = 0; q < 4; ++q)
            PMB_UNROLL
            for (int r = q; r < 4; ++r) b[r][q] = (r < nb) ? Lp[bc[q] + j0 + r] : 0.0;
        const int i0 = j0 + nb + tid;                          // the thread's first row below the panel
        const bool own = i0 < n;
        double av[4];                                          // threads without a row read the (read-only in P1) diagonal entry instead
        PMB_UNROLL
        for (int q = 0; q < 4; ++q) av[q] = (q < nb) ? Lp[bc[q] + (own ? i0 : j0 + q)] : 0.0;
        PMB_UNROLL
        for (int q = 0; q < 4; ++q) {
            d[q] = b[q][q];
            sc[q] = dm::fabs(d[q]) > 0.0;
            PMB_UNROLL
            for (int r = q + 1; r < 4; ++r) {
                const double v0 = b[r][q];
#ifdef PMB_UB_SKIP
                const double v = (PMB_UB_SKIP & 32) ? v0 * d[q] : (sc[q] ? v0 / d[q] : v0);
#else
                const double v = sc[q] ? v0 / d[q] : v0;       // division issued unconditionally, result selected
#endif
                b[r][q] = v;                                   // L(j0 + r, j0 + q)
                ub[r][q] = d[q] * v;                           // d_q L(j0 + r, j0 + q)
            }
            {
#ifdef PMB_UB_SKIP
                av[q] = (PMB_UB_SKIP & 16) ? av[q] * d[q] : (sc[q] ? av[q] / d[q] : av[q]);
#else
                av[q] = sc[q] ? av[q] / d[q] : av[q];
#endif
            }
            PMB_UNROLL
            for (int q2 = q + 1; q2 < 4; ++q2) {
                PMB_UNROLL
                for (int r = q2; r < 4; ++r) b[r][q2] = dm::fma(-b[r][q], ub[q2][q], b[r][q2]);
                av[q2] = dm::fma(-av[q], ub[q2][q], av[q2]);
            }
        }
        PMB_UNROLL
        for (int q = 0; q < 4; ++q) if (q < nb && own) Lp[bc[q] + i0] = av[q];
        // further rows of this thread (only when n - panel > number of threads)
        for (int i = i0 + nt; i < n; i += nt) {
            double aw[4];
            PMB_UNROLL
            for (int q = 0; q < 4; ++q) aw[q] = (q < nb) ? Lp[bc[q] + i] : 0.0;
            PMB_UNROLL
            for (int q = 0; q < 4; ++q) {
                aw[q] = sc[q] ? aw[q] / d[q] : aw[q];
                PMB_UNROLL
                for (int q2 = q + 1; q2 < 4; ++q2) aw[q2] = dm::fma(-aw[q], ub[q2][q], aw[q2]);
            }
            PMB_UNROLL
            for (int q = 0; q < 4; ++q) if (q < nb) Lp[bc[q] + i] = aw[q];
        }
        c.sync();
        // rows inside the panel: thread r < nb stores row j0 + r of the factored diagonal block (after the barrier: every
        // thread has read the unfactored block by now; nothing below reads these entries before the next barrier)
        if (tid < nb) {
            PMB_UNROLL
            for (int r = 0; r < 4; ++r)
                if (r == tid) {
                    PMB_UNROLL
                    for (int q = 0; q <= r; ++q) Lp[bc[q] + j0 + r] = (q == r) ? d[q] : b[r][q];
                }
        }
        // ---- P2: trailing update, a(i,k) = fma(-L(i,j0+q), d_q L(k,j0+q), a(i,k)) for q ascending, k > panel, i >= k
        const int kfirst = j0 + nb;
#ifdef PMB_UB_SKIP
        if (!(PMB_UB_SKIP & 8))
#endif
        if (kfirst < n) {
            double nl[R][4];
            PMB_UNROLL
            for (int r = 0; r < R; ++r) {
                const int i = lane + 32 * r;
                PMB_UNROLL
                for (int q = 0; q < 4; ++q) nl[r][q] = (q < nb && i >= kfirst && i < n) ? -Lp[bc[q] + i] : 0.0;
            }
            // two columns (k and k + nw) per trip, so that two independent load -> 4 FMA -> store chains are in flight
            PMB_NOUNROLL
            for (int k = kfirst + wid; k < n; k += 2 * nw) {
                const int k2 = k + nw;
                const bool has2 = k2 < n;
                const int k2c = has2 ? k2 : k;                 // clamped: loads stay inside the matrix, stores are predicated
                const int bk = k * n - ((k * (k + 1)) >> 1);   // packed_off(k, n) - k
                const int bk2 = k2c * n - ((k2c * (k2c + 1)) >> 1);
                double uk[4], uk2[4];
                PMB_UNROLL
                for (int q = 0; q < 4; ++q) {
                    uk[q] = (q < nb) ? d[q] * Lp[bc[q] + k] : 0.0;
                    uk2[q] = (q < nb) ? d[q] * Lp[bc[q] + k2c] : 0.0;
                }
                double* ck = Lp + bk;
                double* ck2 = Lp + bk2;
                const int c0 = k >> 5;                         // first chunk with rows >= k (warp-uniform)
                // one uniform dispatch per column pair, then straight-line code: unconditional loads (lanes whose row
                // lies above the diagonal read an entry of the finished panel instead — nobody writes it in P2 — and
                // discard it), 4 chained FMAs, predicated stores
                const double* dummy = Lp + bc[0] + k;
                PMB_UNROLL
                for (int r0 = 0; r0 < R; ++r0) {
                    if (c0 == r0) {
                        double acc[R], acc2[R];
                        PMB_UNROLL
                        for (int r = r0; r < R; ++r) {
                            const int i = lane + 32 * r;
                            acc[r] = *((i >= k && i < n) ? ck + i : dummy);
                            acc2[r] = *((i >= k2c && i < n) ? ck2 + i : dummy);
                        }
                        PMB_UNROLL
                        for (int q = 0; q < 4; ++q) {
                            if (q < nb) {
                                PMB_UNROLL
                                for (int r = r0; r < R; ++r) { acc[r] = dm::fma(nl[r][q], uk[q], acc[r]); acc2[r] = dm::fma(nl[r][q], uk2[q], acc2[r]); }
                            }
                        }
                        PMB_UNROLL
                        for (int r = r0; r < R; ++r) {
                            const int i = lane + 32 * r;
                            if (i >= k && i < n) ck[i] = acc[r];
                            if (has2 && i >= k2 && i < n) ck2[i] = acc2[r];
                        }
                    }
                }
            }
        }
        c.sync();
    }
}

/** solve (P^T L D L^T P) s = rhs; `sol` holds rhs on entry and the solution on exit (unpermuted indexing).
 *
 *  The substitutions are a dependent chain of n steps (one shuffle + one FMA per step, ~45-60 cycles measured on B200), so
 *  the schedule keeps that chain as short as possible and moves everything else off it.  Rows are split in chunks of 32;
 *  chunk r belongs to warp r % NW and every thread keeps its rows in registers.  Phase p (ascending for L, descending
 *  for L^T): (A) the owner warp substitutes through the 32x32 diagonal block of chunk p — the only serial part — and
 *  publishes the finished components in shared memory; one block barrier; (B) every warp applies the 32 finished
 *  columns (rows for L^T) to the chunks it owns further down (up) — 32 independent-of-other-warps FMAs per row, which
 *  overlap with the next owner's diagonal block.  Every component still receives exactly the same fused multiply-adds
 *  in exactly the same order as the plain column-by-column substitution: results are bit-identical to the oracle.
 *  The division by D is spread over the whole block (an fp64 division is ~125 cycles).  Ends with a block barrier. */
/** shared-memory copy of the 32 x 32 diagonal blocks of the packed factor: Ld[(p * 32 + jl) * DIAG_LD + il] = L(32 p + il, 32 p + jl),
 *  il >= jl.  Only used when the factor itself lives in a global (L2) slot: the serial substitution through a diagonal block is
 *  a chain of dependent steps and every step loads its multipliers first — from L2 that is ~700 cycles per step (measured on the
 *  kite, n = 377: 190 k cycles per forward + backward solve, 12 blocks x 8 steps x 2 sweeps), from shared memory ~30. */
constexpr int DIAG_LD = 33;
PMB_HD constexpr size_t diag_block_doubles(int n) { return (size_t)((n + 31) / 32) * 32 * DIAG_LD; }
/** shared memory (doubles) behind the Ld pointer of qp_solve_cta: the diagonal blocks + one transposition tile per warp */
PMB_HD constexpr size_t staged_solve_doubles(int n, int /*nwarps*/) { return diag_block_doubles(n); }

template <int R>
PMB_DEV void ldlt_stage_diag_blocks(Cta& c, int n, const double* Lp, double* Ld)
{
    for (int e = c.tid(); e < R * 32 * 32; e += c.nthreads()) {
        const int il = e & 31, jl = (e >> 5) & 31, p = e >> 10;
        const int i = 32 * p + il, j = 32 * p + jl;
        if (il >= jl && i < n) Ld[(p * 32 + jl) * DIAG_LD + il] = Lp[j * n - ((j * (j + 1)) >> 1) + i];
    }
    c.sync();
}

/** GSLOT: the factor lives in a global (L2) slot.  Two further schedules were tried for that case and measured on the kite
 *  (n = 377, 92 ADMM trips per QP) — prefetching the critical-path block of the next phase into registers (slower: the 255-register
 *  kernel spills more) and reading the backward sweep's blocks with the lanes along the contiguous direction through a
 *  shared-memory transposition tile (no gain with a small tile; with a full 32 x 33 tile per warp the lost L1 slows the spilled
 *  code by 2x).  What is left is the L2 streaming of the factor (2 x 570 KB per trip) plus the dependent chain. */
template <int R, int NW = 4, bool GSLOT = false>
PMB_DEV void ldlt_solve_packed(Cta& c, int n, const double* Lp, const int* perm, double* sol, double* ybuf /* n doubles, shared */,
                               const double* Ld = nullptr /* staged diagonal blocks (ldlt_stage_diag_blocks) or null */)
{
    constexpr int RW = (R + NW - 1) / NW;        // chunks per warp
    const Warp& w = c.w;
    const int lane = w.lane(), wid = c.warp_id();
    double y[RW];
    int offr[RW];      // packed_off(i, n) - i of the thread's rows
    PMB_UNROLL
    for (int s = 0; s < RW; ++s) {
        const int i = lane + 32 * (wid + NW * s);
        y[s] = i < n ? sol[perm[i]] : 0.0;
        offr[s] = i < n ? i * n - ((i * (i + 1)) >> 1) : 0;
    }
    PMB_TICK(0)
    // ---- unit lower, ascending.  L(i,j) = Lp[bc(j) + i], bc(j) = packed_off(j) - j, bc(j+1) = bc(j) + n - j - 1
    PMB_UNROLL
    for (int p = 0; p < R; ++p) {
        const int j0 = 32 * p;
        const int jend = (n - j0) < 32 ? (n - j0) : 32;
        const int bc0 = j0 * n - ((j0 * (j0 + 1)) >> 1);
#ifdef PMB_UB_SKIP
        if (!(PMB_UB_SKIP & 1))
#endif
        if (wid == p % NW) {                                  // (A) diagonal block of chunk p
            const int s = p / NW;   // compile-time after unrolling
            double yy = y[s];
            // four columns at a time: the 4 partial components are broadcast with independent shuffles, every lane
            // finishes the 4x4 triangle in registers (same FMAs, same order as the owner lane would do), then the
            // rows below apply the four columns in ascending order — one shuffle latency per 4 steps instead of per step
            int cb = bc0 + j0;                                // bc(j) + j0 for the block's first column j
            int step = n - j0 - 1;                            // bc(j+1) - bc(j)
            int l0 = 0;
            PMB_NOUNROLL
            for (; l0 + 4 <= jend; l0 += 4) {
                const int c0 = cb, c1 = c0 + step, c2 = c1 + step - 1, c3 = c2 + step - 2;
                double y0 = w.shfl(yy, l0), y1 = w.shfl(yy, l0 + 1), y2 = w.shfl(yy, l0 + 2), y3 = w.shfl(yy, l0 + 3);
                // branch-free: lanes that are not below the block read a valid dummy row and discard the result
                const bool below = lane >= l0 + 4 && lane < jend;
                const int il = below ? lane : l0 + 3;
                double d10, d20, d30, d21, d31, d32, e0, e1, e2, e3;
                if (Ld) {                                     // column j0 + l0 + t of the block is row (p * 32 + l0 + t) of Ld
                    const double* b0 = Ld + (p * 32 + l0) * DIAG_LD;
                    const double *b1 = b0 + DIAG_LD, *b2 = b1 + DIAG_LD, *b3 = b2 + DIAG_LD;
                    d10 = b0[l0 + 1]; d20 = b0[l0 + 2]; d30 = b0[l0 + 3]; d21 = b1[l0 + 2]; d31 = b1[l0 + 3]; d32 = b2[l0 + 3];
                    e0 = b0[il]; e1 = b1[il]; e2 = b2[il]; e3 = b3[il];
                } else {
                    d10 = Lp[c0 + l0 + 1]; d20 = Lp[c0 + l0 + 2]; d30 = Lp[c0 + l0 + 3];
                    d21 = Lp[c1 + l0 + 2]; d31 = Lp[c1 + l0 + 3]; d32 = Lp[c2 + l0 + 3];
                    e0 = Lp[c0 + il]; e1 = Lp[c1 + il]; e2 = Lp[c2 + il]; e3 = Lp[c3 + il];
                }
                y1 = dm::fma(-d10, y0, y1);
                y2 = dm::fma(-d20, y0, y2); y2 = dm::fma(-d21, y1, y2);
                y3 = dm::fma(-d30, y0, y3); y3 = dm::fma(-d31, y1, y3); y3 = dm::fma(-d32, y2, y3);
                double t = dm::fma(-e0, y0, yy); t = dm::fma(-e1, y1, t); t = dm::fma(-e2, y2, t); t = dm::fma(-e3, y3, t);
                const double own = lane == l0 + 1 ? y1 : (lane == l0 + 2 ? y2 : (lane == l0 + 3 ? y3 : yy));
                yy = below ? t : own;
                cb = c3 + step - 3; step -= 4;
            }
            const double* colp = Lp + cb + lane;              // remaining (< 4) columns one by one
            PMB_NOUNROLL
            for (int jj = l0; jj < jend; ++jj) {
                const double yj = w.shfl(yy, jj);
                if (lane > jj && lane < jend) yy = dm::fma(-(Ld ? Ld[(p * 32 + jj) * DIAG_LD + lane] : colp[0]), yj, yy);
                colp += step; --step;
            }
            y[s] = yy;
            if (lane < jend) ybuf[j0 + lane] = yy;
        }
        c.sync();
        PMB_UNROLL
#ifdef PMB_UB_SKIP
        if (!(PMB_UB_SKIP & 2))
#endif
        for (int s = 0; s < RW; ++s) {                        // (B) columns of chunk p -> owned chunks below
            const int r = wid + NW * s;
            const int i = lane + 32 * r;
            if (r > p && r < R && i < n) {                    // chunk p has rows below it, so it is a full chunk: 32 columns
                const double* colp = Lp + bc0 + i;
                const double* yp = ybuf + j0;
                double yy = y[s];
                int step = n - j0 - 1;
                PMB_UNROLL
                for (int jj = 0; jj < 32; ++jj) { yy = dm::fma(-colp[0], yp[jj], yy); colp += step; --step; }
                y[s] = yy;
            }
        }
    }
    PMB_TICK(1)
    // ---- D^-1 (component zeroed when |d| <= DBL_MIN, like Eigen's LDLT::solve): every thread divides its own rows
    PMB_UNROLL
    for (int s = 0; s < RW; ++s) {
        const int i = lane + 32 * (wid + NW * s);
#ifdef PMB_UB_SKIP
        if (PMB_UB_SKIP & 4) { if (i < n) y[s] = y[s] * Lp[offr[s] + i]; } else
#endif
        if (i < n) { const double di = Lp[offr[s] + i]; y[s] = (dm::fabs(di) > DBL_MIN) ? (y[s] / di) : 0.0; }
    }
    PMB_TICK(2)
    // ---- unit upper (L^T), descending.  The thread's row i reads L(j,i) = Lp[offr + j]
    PMB_UNROLL
    for (int p = R - 1; p >= 0; --p) {
        const int j0 = 32 * p;
        const int jend = (n - j0) < 32 ? (n - j0) : 32;
        if (wid == p % NW) {
            const int s = p / NW;   // compile-time after unrolling
            // L(j0 + jj, i) for the thread's i = j0 + lane: row `lane` of the staged block, or the packed column i
            const double* rowp = Ld ? Ld + (p * 32 + (lane < jend ? lane : 0)) * DIAG_LD : Lp + offr[s] + j0;
            double yy = y[s];
            const int full = jend & ~3;
            PMB_NOUNROLL
            for (int jj = jend - 1; jj >= full; --jj) {       // trailing (< 4) columns one by one
                const double yj = w.shfl(yy, jj);
                if (lane < jj) yy = dm::fma(-rowp[jj], yj, yy);
            }
            PMB_NOUNROLL
            for (int l0 = full - 4; l0 >= 0; l0 -= 4) {       // four columns at a time, descending (see the forward sweep)
                const int j = j0 + l0;
                const int c0 = j * n - ((j * (j + 1)) >> 1), c1 = c0 + (n - j - 1), c2 = c1 + (n - j - 2);
                double y0 = w.shfl(yy, l0), y1 = w.shfl(yy, l0 + 1), y2 = w.shfl(yy, l0 + 2), y3 = w.shfl(yy, l0 + 3);
                const bool above = lane < l0;
                double d10, d20, d30, d21, d31, d32;
                const double* rp;                                     // dummy for the other lanes: L(j0 + l0 + t, j), valid entries of column j
                if (Ld) {
                    const double* b0 = Ld + (p * 32 + l0) * DIAG_LD;
                    const double *b1 = b0 + DIAG_LD, *b2 = b1 + DIAG_LD;
                    d10 = b0[l0 + 1]; d20 = b0[l0 + 2]; d30 = b0[l0 + 3]; d21 = b1[l0 + 2]; d31 = b1[l0 + 3]; d32 = b2[l0 + 3];
                    rp = above ? rowp : b0;
                } else {
                    d10 = Lp[c0 + j + 1]; d20 = Lp[c0 + j + 2]; d30 = Lp[c0 + j + 3];
                    d21 = Lp[c1 + j + 2]; d31 = Lp[c1 + j + 3]; d32 = Lp[c2 + j + 3];
                    rp = above ? rowp : Lp + c0 + j0;
                }
                const double e0 = rp[l0], e1 = rp[l0 + 1], e2 = rp[l0 + 2], e3 = rp[l0 + 3];
                y2 = dm::fma(-d32, y3, y2);
                y1 = dm::fma(-d31, y3, y1); y1 = dm::fma(-d21, y2, y1);
                y0 = dm::fma(-d30, y3, y0); y0 = dm::fma(-d20, y2, y0); y0 = dm::fma(-d10, y1, y0);
                double t = dm::fma(-e3, y3, yy); t = dm::fma(-e2, y2, t); t = dm::fma(-e1, y1, t); t = dm::fma(-e0, y0, t);
                const double own = lane == l0 ? y0 : (lane == l0 + 1 ? y1 : (lane == l0 + 2 ? y2 : yy));
                yy = above ? t : own;
            }
            y[s] = yy;
            if (lane < jend) ybuf[j0 + lane] = yy;
        }
        c.sync();
        PMB_UNROLL
        for (int s = 0; s < RW; ++s) {
            const int r = wid + NW * s;
            if (r < p) {
                const double* rowp = Lp + offr[s] + j0;
                const double* yp = ybuf + j0;
                double yy = y[s];
                if (p == R - 1) {                             // only the last chunk can be partial
                    PMB_UNROLL
                    for (int jj = 31; jj >= 0; --jj) if (jj < jend) yy = dm::fma(-rowp[jj], yp[jj], yy);
                } else {
                    PMB_UNROLL
                    for (int jj = 31; jj >= 0; --jj) yy = dm::fma(-rowp[jj], yp[jj], yy);
                }
                y[s] = yy;
            }
        }
    }
    PMB_TICK(3)
    // every chunk's final values are in ybuf (published in its phase, visible after that phase's barrier)
    for (int i = c.tid(); i < n; i += c.nthreads()) sol[perm[i]] = ybuf[i];
    c.sync();
    PMB_TICK(4)
}

/** the whole boxADMM solve of one instance by one CTA.  Lp: n(n+1)/2 doubles (shared or global), vec: qp_vec_bytes() of
 *  shared memory. */
}  // namespace pmb
#include "pmb_qp_fast.hpp"
namespace pmb {

/** FAST: the factor workspace Lp is laid out by fast::Ws (fast::workspace_doubles) and the linear algebra runs on the fp64
 *  tensor cores (pmb_qp_fast.hpp); everything else — classification, rho, ADMM updates, residuals, termination — is shared */
template <int R, int NC = 0, int MC = 0, int NW = 4, bool FAST = false, bool GLOBAL_SLOT = false>   // NC, MC: problem size when known at compile time (fused SQP kernel), 0 = a.N, a.M; NW: warps per CTA; GLOBAL_SLOT: Lp is global memory (prefetching solves)
PMB_DEV void qp_solve_cta(Cta& c, const pmb_qp_settings_t& st, const QpArgs& a, double* Lp, unsigned char* vec,
                          double* Ld = nullptr /* exact arithmetic, factor in a global slot: diag_block_doubles(n) of shared memory */)
{
    const int N = NC > 0 ? NC : a.N, M = (NC > 0) ? MC : a.M, n = N + M, tid = c.tid(), nt = c.nthreads();
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
    double* first = rvi + M;   // [12]
    int* perm = reinterpret_cast<int*>(first + 12);
    int* ctype = perm + n;     // [constr_type (M) ; box_constr_type (N)]
    int* iperm = ctype + n;    // inverse pivot permutation (fast arithmetic: the right-hand side is kept in pivot order)
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
    for (int i = tid; i < M; i += nt) z[i] = a.xg ? dot_chain(a.A + i, (size_t)M, x, N) : zero_guess_row(a.A + i, (size_t)M, N);

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
        ldlt_pivot_order<R>(c, n, dK, perm, reinterpret_cast<int*>(tmp));
        if (n_factor == 0 && a.perm) { for (int i = tid; i < n; i += nt) a.perm[i] = perm[i]; }
        if (FAST) { for (int i = tid; i < n; i += nt) iperm[perm[i]] = i; }
        const unsigned long long t1 = prof ? c.w.clock() : 0;
        if (FAST) fast::gather<R>(c, N, M, a.H, a.A, dK, perm, fast::Ws(Lp, n));
        else kkt_gather_permuted<R>(c, N, M, a.H, a.A, dK, perm, Lp);
        const unsigned long long t2 = prof ? c.w.clock() : 0;
        if (FAST) {
            const fast::Ws fw(Lp, n);
            fast::FactorProf fp;
            fast::factor(c, fw, prof ? &fp : nullptr);
            const unsigned long long ti = prof ? c.w.clock() : 0;
            fast::invert(c, fw);
            if (prof) { prof->f_diag += fp.diag; prof->f_panel += fp.panel; prof->f_trail += fp.trail; prof->f_inv += c.w.clock() - ti; }
        }
        else {
            ldlt_factor_packed<R>(c, n, Lp);
            if (Ld) ldlt_stage_diag_blocks<R>(c, n, Lp, Ld);
        }
        if (prof) { const unsigned long long t3 = c.w.clock(); prof->pivot += t1 - t0; prof->gather += t2 - t1; prof->factor += t3 - t2; }
        ++n_factor;
    };

    rho_vec_update(st.rho);
    // diag of construct_kkt_matrix (box_admm.hpp:207-223): (H_ii + sigma) + rho_box_i ; -1/rho_A
    for (int i = tid; i < N; i += nt) { double v = a.H[i + (size_t)i * N]; v += st.sigma; v += rb[i]; dK[i] = v; }
    for (int i = tid; i < M; i += nt) dK[N + i] = -rvi[i];
    c.sync();

    int status = PMB_QP_UNSOLVED;
    double res_prim = 1.0, res_dual = 1.0, rho_estimate = 0.0, max_Ax_z = 0.0, max_Hx_ATy_h = 0.0;
    const double alpha = st.alpha, sigma = st.sigma;

    auto residuals_update = [&]() {   // box_admm.hpp:398-415; task t < M: row t of A x, task M + i: row i of H x and A^T y
        enum { nAx = 0, nz, nx, nHx, nATy, nh, nyb, rp, rq, rd, NRED };
        double m[NRED];
        PMB_UNROLL
        for (int k = 0; k < NRED; ++k) m[k] = 0.0;
        // every norm is an lpNorm<Infinity>() = a max chain that starts from the FIRST coefficient: a NaN there sticks, a NaN
        // anywhere else is skipped (oracle/canon.hpp::norm_inf).  The owners of row 0 publish |first coefficient| in first[].
        if (tid < NRED) first[tid] = 0.0;
        c.sync();
        for (int t = tid; t < M + N; t += nt) {
            if (t < M) {
                const int i = t;
                const double acc = dot_chain(a.A + i, (size_t)M, x, N);
                const double v0 = dm::fabs(acc), v1 = dm::fabs(z[i]), v2 = dm::fabs(acc - z[i]);
                if (v0 > m[nAx]) m[nAx] = v0;
                if (v1 > m[nz]) m[nz] = v1;
                if (v2 > m[rp]) m[rp] = v2;
                if (i == 0) { first[nAx] = v0; first[nz] = v1; first[rp] = v2; }
            } else {
                const int i = t - M;
                const double hx = dot_chain(a.H + i, (size_t)N, x, N);
                const double aty = dot_chain(a.A + (size_t)i * M, 1, ya, M);
                const double v0 = dm::fabs(x[i]), v1 = dm::fabs(hx), v2 = dm::fabs(aty), v3 = dm::fabs(h[i]), v4 = dm::fabs(yb[i]);
                const double v5 = dm::fabs(x[i] - q[i]), v6 = dm::fabs(((hx + h[i]) + aty) + yb[i]);
                if (v0 > m[nx]) m[nx] = v0;
                if (v1 > m[nHx]) m[nHx] = v1;
                if (v2 > m[nATy]) m[nATy] = v2;
                if (v3 > m[nh]) m[nh] = v3;
                if (v4 > m[nyb]) m[nyb] = v4;
                if (v5 > m[rq]) m[rq] = v5;
                if (v6 > m[rd]) m[rd] = v6;
                if (i == 0) { first[nx] = v0; first[nHx] = v1; first[nATy] = v2; first[nh] = v3; first[nyb] = v4; first[rq] = v5; first[rd] = v6; }
            }
        }
        c.sync();
        c.max_all<NRED>(m);
        PMB_UNROLL
        for (int k = 0; k < NRED; ++k) { const double f = first[k]; if (f != f) m[k] = f; }
        max_Ax_z = fmax_nan(m[nAx], fmax_nan(m[nz], m[nx]));
        max_Hx_ATy_h = fmax_nan(m[nHx], fmax_nan(m[nATy], fmax_nan(m[nh], m[nyb])));
        res_prim = m[rp] + m[rq];
        res_dual = m[rd];
    };

    // The loop has ONE call site for the factorisation and one for the residuals (code size: the kernel's hot loops must
    // stay resident in the instruction cache).  The reference factorises before the loop and again right after a rho
    // update; doing it at the top of the next trip (also when that trip is not executed any more) is the same sequence of
    // operations.
    bool need_factor = true;
    int iter;
    for (iter = 1; ; ++iter) {
        if (need_factor) {
            factorise(); need_factor = false;
            if (FAST) {   // compute_kkt_rhs (351-355) in pivot order, padded to whole tiles; later trips refresh it inside the update phase
                const fast::Ws fw(Lp, n);
                for (int e = tid; e < fw.T * 8; e += nt) {
                    double v = 0.0;
                    if (e < n) { const int i = perm[e]; v = i < N ? ((sigma * x[i] - h[i]) + rb[i] * q[i]) - yb[i] : z[i - N] - rvi[i - N] * ya[i - N]; }
                    fw.tb[e] = v;
                }
                c.sync();
            }
        }
        if (iter > st.max_iter) break;
        const unsigned long long ta = prof ? c.w.clock() : 0;
        if (!FAST) {
            // compute_kkt_rhs (351-355)
            for (int i = tid; i < N; i += nt) sol[i] = ((sigma * x[i] - h[i]) + rb[i] * q[i]) - yb[i];
            for (int i = tid; i < M; i += nt) sol[N + i] = z[i] - rvi[i] * ya[i];
            c.sync();
        }
        const unsigned long long tb = prof ? c.w.clock() : 0;
        if (FAST) fast::solve_rows<(4 * R + NW - 1) / NW, NW>(c, fast::Ws(Lp, n), perm, sol);   // T <= 4 R tile rows, dealt round-robin
        else ldlt_solve_packed<R, NW, GLOBAL_SLOT>(c, n, Lp, perm, sol, tmp, Ld);
        const unsigned long long tc = prof ? c.w.clock() : 0;
        double* const tbn = FAST ? fast::Ws(Lp, n).tb : nullptr;      // fast arithmetic: the next trip's right-hand side, pivot order
        // z, y_A (126, 133-135, 142-144)
        for (int i = tid; i < M; i += nt) {
            const double zp = z[i];
            const double zt = zp + rvi[i] * (sol[N + i] - ya[i]);
            double v = alpha * zt;
            v += ((1 - alpha) * zp) + (rvi[i] * ya[i]);
            const double zn = dm::min(dm::max(v, alb[i]), aub[i]);
            const double yn = ya[i] + rv[i] * (((alpha * zt) + ((1 - alpha) * zp)) - zn);
            ya[i] = yn;
            z[i] = zn;
            if (FAST) tbn[iperm[N + i]] = zn - rvi[i] * yn;
        }
        // x, q, y_box (129-130, 138-139, 146-147)
        for (int i = tid; i < N; i += nt) {
            double xv = alpha * sol[i];
            xv += (1 - alpha) * xv;
            x[i] = xv;
            const double qv = dm::min(dm::max(xv + rbi[i] * yb[i], xlb[i]), xub[i]);
            q[i] = qv;
            const double ybn = yb[i] + rb[i] * (xv - qv);
            yb[i] = ybn;
            if (FAST) tbn[iperm[i]] = ((sigma * xv - h[i]) + rb[i] * qv) - ybn;
        }
        c.sync();
        if (prof) { const unsigned long long td = c.w.clock(); prof->update += (tb - ta) + (td - tc); prof->solve += tc - tb; }

        const bool check = (st.check_termination != 0) && (iter % st.check_termination == 0);
        const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0);
        if (check || adapt) {
            const unsigned long long te = prof ? c.w.clock() : 0;
            residuals_update();
            if (prof) prof->resid += c.w.clock() - te;
        }
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
                for (int i = tid; i < N; i += nt) tmp[i] = rb[i];   // rho_box_prev
                c.sync();
                rho_vec_update(new_rho);
                // update_kkt_rho (447-452)
                for (int i = tid; i < N; i += nt) dK[i] += (rb[i] - tmp[i]);
                for (int i = tid; i < M; i += nt) dK[N + i] = -rvi[i];
                c.sync();
                need_factor = true;
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
