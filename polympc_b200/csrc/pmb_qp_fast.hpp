// pmb_qp_fast.hpp — the KKT linear algebra of boxADMM in "fast arithmetic": tile-blocked LDL^T on the fp64 tensor cores, an
// explicit inverse of the unit-lower factor, and triangular solves as streaming mat-vecs.
//
// Same mathematics as pmb_qp.hpp (reference src/solvers/box_admm.hpp:207-223, 335-355 + Eigen::LDLT<Lower>): the SAME pivot
// permutation (Eigen's diagonal rule, replayed exactly), K = P^T L D L^T P, and every ADMM trip solves with that factor.
// What changes is the order of the floating-point operations — so results agree with the oracle to rounding (1e-12 per SQP
// iteration, tests/test_gpu_fast.py), not bit for bit.  The bit-exact path stays available (pmb_sqp_set_arithmetic).
//
// Why: on the exact path one SQP iteration spends 312 k of its 740 k cycles in 2 x 104 dependent substitution steps per
// ADMM trip and 179 k in a right-looking LDL^T whose inner loops are fp64-issue and division bound (profiles/README.md).
//   * Layout.  P K P^T lives in shared memory as 8 x 8 tiles of the lower block triangle, tile (I, J) at
//     (I (I + 1) / 2 + J) * 64 doubles; inside a tile element (r, c) sits at (c >> 2) * 32 + r * 4 + (c & 3): two 8 x 4 halves,
//     each row-major.  In this order BOTH tensor-core fragment shapes are bank-conflict free: the A / B^T operand of
//     mma.m8n8k4 (lane l <-> element (l >> 2, l & 3) of a half) is 32 consecutive doubles, the accumulator (lane l <-> row
//     l >> 2, columns 2 (l & 3) + {0, 1}) is 32 consecutive double2 — and so is the streaming mat-vec, where lane l owns
//     double2 number l of every tile it visits.
//   * Factorisation, block size 8: (a) one warp factors the 8 x 8 diagonal tile in the accumulator layout with shuffles —
//     the only serial part, 8 dependent (shuffle, reciprocal, FMA) steps — and builds L_kk^-1 on the way; (b) the panel
//     below is two DMMAs per tile, U = A_Ik L_kk^-T, L = U D^-1; (c) the trailing update C_IJ -= L_Ik U_Jk^T is two DMMAs per
//     tile.  A 104 x 104 factorisation is ~900 DMMAs instead of ~190 k scalar FMAs.
//   * The unit-lower factor is then inverted in place (row by row of tiles, three block barriers per tile row), so that a
//     triangular solve becomes y = X t: one pass over the 46 KB of tiles with 16-byte loads at full shared-memory bandwidth and
//     NO dependent chain.  An ADMM trip costs two such passes.
#pragma once
#include "pmb_cta.hpp"
#include <cfloat>

namespace pmb {
namespace fast {

PMB_DEV constexpr int tiles_of(int n) { return (n + 7) >> 3; }
/** doubles of the fast workspace: tiles | two tile rows (negated U panel / M rows of the inversion) | W fragments (2 x 64) |
 *  rfac, mask (2 x 16) | dinv, tb, yb (n_p each) */
PMB_HD constexpr size_t workspace_doubles(int n)
{
    return ((size_t)((n + 7) >> 3) * (size_t)(((n + 7) >> 3) + 1) / 2) * 64 + 2 * (size_t)((n + 7) >> 3) * 64 + 128 + 32 + 3 * (size_t)((n + 7) >> 3) * 8;
}

struct Ws {
    int T, n;
    double *tiles, *upanel, *wfrag, *rfac, *dinv, *tb, *yb;
    PMB_DEV Ws(double* base, int n_) : T((n_ + 7) >> 3), n(n_)
    {
        tiles = base;
        upanel = tiles + (size_t)(T * (T + 1) / 2) * 64;     // [2][T tiles]
        wfrag = upanel + 2 * T * 64;                          // [2][64]
        rfac = wfrag + 128;                                   // [2][16]
        dinv = rfac + 32;
        tb = dinv + T * 8;
        yb = tb + T * 8;
    }
    PMB_DEV double* tile(int I, int J) const { return tiles + ((I * (I + 1)) / 2 + J) * 64; }
};

PMB_DEV int pos(int r, int c) { return ((c >> 2) << 5) + (r << 2) + (c & 3); }

/** 16-byte pair: one LDS.128 / STS.128 */
struct alignas(16) d2 { double x, y; };

/** optional cycle counters of the factorisation (thread-local): diagonal tiles, panels, trailing updates, inversion */
struct FactorProf { unsigned long long diag = 0, panel = 0, trail = 0, inv = 0; };

/** 1 / x for the pivots: hardware seed + two Newton steps (error ~1 ulp; the exact path divides) */
PMB_DEV double rcp(double x)
{
#if defined(__CUDACC__) && !defined(PMB_EMU)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = __fma_rn(-x, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-x, y, 1.0);
    y = __fma_rn(y, e, y);
    return y;
#else
    return 1.0 / x;
#endif
}

/** gather P K P^T into the tiles (diagonal tiles are filled symmetrically; rows / columns beyond n are identity).  Columns are
 *  dealt round-robin to the warps, rows to the lanes; the R loads of a lane are independent (memory-level parallelism: H and
 *  A come from L2). */
template <int R>
PMB_DEV void gather(Cta& c, int N, int M, const double* H, const double* A, const double* dK, const int* perm, const Ws& w)
{
    const int n = N + M, np = w.T * 8, lane = c.lane(), nw = c.nwarps();
    int pa[R];
    PMB_UNROLL
    for (int r = 0; r < R; ++r) { const int a = lane + 32 * r; pa[r] = a < n ? perm[a] : -1; }
    PMB_NOUNROLL
    for (int b = c.warp_id(); b < np; b += nw) {
        const int cc = b < n ? perm[b] : -1;
        const int J = b >> 3, cb = b & 7;
        double v[R];
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int a = lane + 32 * r;
            double x = (a == b) ? 1.0 : 0.0;
            if (a >= b && a < n && b < n) {
                const int rr = pa[r];
                const int hi = rr > cc ? rr : cc, lo = rr > cc ? cc : rr;
                if (hi == lo) x = dK[hi];
                else if (hi < N) x = H[hi + (size_t)lo * N];
                else if (lo < N) x = A[(hi - N) + (size_t)lo * M];
                else x = 0.0;
            }
            v[r] = x;
        }
        PMB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int a = lane + 32 * r;
            if (a >= b && a < np) {
                const int I = a >> 3, ra = a & 7;
                double* t = w.tile(I, J);
                t[pos(ra, cb)] = v[r];
                if (I == J) t[pos(cb, ra)] = v[r];
            }
        }
    }
    c.sync();
}

/** (a) of the factorisation: LDL^T of the symmetric 8 x 8 diagonal tile, executed by ONE warp in the accumulator layout
 *  (lane l: row l >> 2, columns 2 (l & 3), 2 (l & 3) + 1).  Leaves X = L_kk^-1 (unit lower, zeros above) in the tile, 1 / d
 *  in dinv / rfac and the B fragments of X^T (for U = A X^T) in wfrag. */
PMB_DEV void factor_diag_tile(const Warp& w, double* tile, double* dinv8, double* rfac, double* wfrag)
{
    const int l = w.lane(), r = l >> 2, q = l & 3, quad = l & ~3;
    double a0 = tile[pos(r, 2 * q)], a1 = tile[pos(r, 2 * q + 1)];
    double x0 = (r == 2 * q) ? 1.0 : 0.0, x1 = (r == 2 * q + 1) ? 1.0 : 0.0;
    double my_dinv = 0.0, my_rfac = 1.0, my_mask = -1.0;
    PMB_UNROLL
    for (int j = 0; j < 8; ++j) {
        const int sq = j >> 1;
        const double sel = (j & 1) ? a1 : a0;
        const double piv = w.shfl(sel, j * 4 + sq);                  // T[j][j]
        const double arj = w.shfl(sel, quad | sq);                   // T[r][j]
        const double ac0 = w.shfl(sel, (2 * q) * 4 + sq);            // T[2q][j]
        const double ac1 = w.shfl(sel, (2 * q + 1) * 4 + sq);        // T[2q+1][j]
        const double xj0 = w.shfl(x0, j * 4 + q), xj1 = w.shfl(x1, j * 4 + q);   // X[j][2q], X[j][2q+1]
        const double ap = dm::fabs(piv);
        const double ri = rcp(piv);
        const double rf = ap > 0.0 ? ri : 1.0;                       // column scaling (Eigen: only when |d| > 0)
        if (l == j) { my_dinv = ap > DBL_MIN ? ri : 0.0; my_rfac = rf; my_mask = ap > 0.0 ? -1.0 : 0.0; }
        const double lr = r > j ? arj * rf : 0.0;                   // L(r, j)
        const double lu = ap > 0.0 ? lr : 0.0;                      // a zero pivot contributes nothing (L d L^T with d = 0)
        if (2 * q > j) a0 = dm::fma(-lu, ac0, a0);
        if (2 * q + 1 > j) a1 = dm::fma(-lu, ac1, a1);
        x0 = dm::fma(-lr, xj0, x0);
        x1 = dm::fma(-lr, xj1, x1);
    }
    if (l < 8) { dinv8[l] = my_dinv; rfac[l] = my_rfac; rfac[8 + l] = my_mask; }
    // X into the tile (accumulator layout == double2 number l)
    tile[pos(r, 2 * q)] = x0; tile[pos(r, 2 * q + 1)] = x1;
    // B fragments of X^T: lane (n = r, kk = q [+4]) needs X[r][q] and X[r][q + 4]
    {
        const double v0 = w.shfl(x0, quad | (q >> 1)), v1 = w.shfl(x1, quad | (q >> 1));
        const double u0 = w.shfl(x0, quad | (2 + (q >> 1))), u1 = w.shfl(x1, quad | (2 + (q >> 1)));
        wfrag[l] = (q & 1) ? v1 : v0;
        wfrag[32 + l] = (q & 1) ? u1 : u0;
    }
}

/** blocked right-looking LDL^T on the tiles; afterwards tile (I, J), J < I, holds L_IJ, tile (I, I) holds L_II^-1 and dinv
 *  holds 1 / d (0 where |d| <= DBL_MIN, like Eigen's LDLT::solve).
 *
 *  Schedule per block column k (two block barriers):
 *    panel     every warp: U = A_Ik X_kk^T (two DMMAs per tile), L = U D^-1 -> tile (I, k), -U -> upanel[I]
 *    trailing  C_IJ += L_Ik (-U_Jk)^T for k < J <= I, two tiles in flight per warp.  LOOK-AHEAD: one warp updates tile
 *              (k+1, k+1) first and factors it right away (the 8-step serial chain of factor_diag_tile) while the other warps
 *              stream through the remaining tiles, so the chain is hidden behind tensor-core work. */
PMB_DEV void factor(Cta& c, const Ws& w, FactorProf* prof = nullptr)
{
    const Warp& wp = c.w;
    const int T = w.T, lane = c.lane(), wid = c.warp_id(), nw = c.nwarps();
    const int r = lane >> 2, q = lane & 3;
    const int cl = ((q >> 1) << 5) + (r << 2) + ((q & 1) << 1);     // accumulator layout: pos(r, 2q), pos(r, 2q) + 1
    // iteration k = -1 only factors tile (0, 0): the serial diagonal-tile routine then has a single (inlined) call site
    PMB_NOUNROLL
    for (int k = -1; k + 1 < T; ++k) {
        const unsigned long long t1 = prof ? wp.clock() : 0;
        if (k >= 0) {
        {   // panel
            const double* wf = w.wfrag + (k & 1) * 64;
            const double* rf = w.rfac + (k & 1) * 16;
            const double b0 = wf[lane], b1 = wf[32 + lane];
            const double f0 = rf[2 * q], f1 = rf[2 * q + 1], g0 = rf[8 + 2 * q], g1 = rf[8 + 2 * q + 1];   // g = -1, or 0 for a zero pivot
            for (int I = k + 1 + wid; I < T; I += nw) {
                double* t = w.tile(I, k);
                const double a0 = t[lane], a1 = t[32 + lane];
                double u0 = 0.0, u1 = 0.0;
                wp.dmma(u0, u1, a0, b0);
                wp.dmma(u0, u1, a1, b1);
                wp.sync();                                           // every lane has read its A fragments of this tile
                d2* tp = reinterpret_cast<d2*>(t + cl);
                d2* up = reinterpret_cast<d2*>(w.upanel + I * 64 + cl);
                *tp = d2{u0 * f0, u1 * f1};
                *up = d2{u0 * g0, u1 * g1};
            }
        }
        c.sync();
        }
        const unsigned long long t2 = prof ? wp.clock() : 0;
        if (prof) prof->panel += t2 - t1;
        {   // trailing update with look-ahead
            const int dw = (k + 1) % nw;                              // the warp that factors the next diagonal tile
            const int workers = nw > 1 ? nw - 1 : 1;
            const int me = nw > 1 ? (wid + nw - dw - 1) % nw : 0;     // 0 .. nw-2 for the workers, nw-1 for the warp of the diagonal tile
            if (wid == dw) {
                if (k >= 0) {                                         // tile (k+1, k+1) first ...
                    const double* la = w.tile(k + 1, k);
                    const double* ub = w.upanel + (k + 1) * 64;
                    d2* ct = reinterpret_cast<d2*>(w.tile(k + 1, k + 1) + cl);
                    d2 cv = *ct;
                    wp.dmma(cv.x, cv.y, la[lane], ub[lane]);
                    wp.dmma(cv.x, cv.y, la[32 + lane], ub[32 + lane]);
                    *ct = cv;
                }
                wp.sync();                                            // ... then its factorisation, while the others stream the rest
                factor_diag_tile(wp, w.tile(k + 1, k + 1), w.dinv + 8 * (k + 1), w.rfac + ((k + 1) & 1) * 16, w.wfrag + ((k + 1) & 1) * 64);
            }
            if (k >= 0 && (me < workers || nw == 1)) {
                // tile columns J = k+1 .. T-1; within a column the rows I >= J (I > J for the first column) are dealt to the
                // workers; the B fragments (-U_Jk) are loaded once per column, two tiles are in flight per step
                for (int J = k + 1; J < T; ++J) {
                    const double* ub = w.upanel + J * 64;
                    const double b0 = ub[lane], b1 = ub[32 + lane];
                    const int first = (J == k + 1 ? J + 1 : J) + me;
                    for (int I = first; I < T; I += 2 * workers) {
                        const int I2 = I + workers < T ? I + workers : I;
                        const double* la = w.tile(I, k);
                        const double* la2 = w.tile(I2, k);
                        d2* ct = reinterpret_cast<d2*>(w.tile(I, J) + cl);
                        d2* ct2 = reinterpret_cast<d2*>(w.tile(I2, J) + cl);
                        const double a0 = la[lane], a1 = la[32 + lane], e0 = la2[lane], e1 = la2[32 + lane];
                        d2 cv = *ct, cw = *ct2;
                        wp.dmma(cv.x, cv.y, a0, b0);
                        wp.dmma(cw.x, cw.y, e0, b0);
                        wp.dmma(cv.x, cv.y, a1, b1);
                        wp.dmma(cw.x, cw.y, e1, b1);
                        *ct = cv;
                        if (I2 != I) *ct2 = cw;
                    }
                }
            }
        }
        c.sync();
        if (prof) prof->trail += wp.clock() - t2;
    }
}

/** in-place inverse of the unit-lower block factor: tile (I, J), J < I, becomes X_IJ = (L^-1)_IJ, row of tiles by row of tiles:
 *      M_IK = X_II L_IK (K < I)            -> a spare tile row (double buffered)
 *      X_IJ = -sum_{K = J}^{I-1} M_IK X_KJ  -> tile (I, J)
 *  The M row of I + 1 only needs L and X_(I+1)(I+1), so it is computed in the same phase as the X row of I: ONE block barrier per
 *  tile row.  Within a warp up to MAXC columns advance together (independent accumulator chains). */
PMB_DEV void invert(Cta& c, const Ws& w)
{
    const Warp& wp = c.w;
    const int T = w.T, lane = c.lane(), wid = c.warp_id(), nw = c.nwarps();
    const int r = lane >> 2, q = lane & 3;
    const int cl = ((q >> 1) << 5) + (r << 2) + ((q & 1) << 1);
    // B operand taken NON-transposed from a tile: element (kk = q [+4], n = r)
    const int bl0 = ((r >> 2) << 5) + (q << 2) + (r & 3), bl1 = bl0 + 16;
    auto m_row = [&](int I) {                                         // M_IK = X_II L_IK for K < I into buffer I & 1
        const double* xd = w.tile(I, I);
        const double xa0 = xd[lane], xa1 = xd[32 + lane];
        double* mb = w.upanel + (I & 1) * T * 64;
        for (int K = wid; K < I; K += nw) {
            const double* t = w.tile(I, K);
            double m0 = 0.0, m1 = 0.0;
            wp.dmma(m0, m1, xa0, t[bl0]);
            wp.dmma(m0, m1, xa1, t[bl1]);
            *reinterpret_cast<d2*>(mb + K * 64 + cl) = d2{m0, m1};
        }
    };
    if (T > 1) m_row(1);
    c.sync();
    PMB_NOUNROLL
    for (int I = 1; I < T; ++I) {
        const double* mb = w.upanel + (I & 1) * T * 64;
        // the warp's columns J = wid, wid + nw, ...: X_IJ = -sum_{K = J}^{I-1} M_IK X_KJ as two independent accumulator chains
        // (even / odd K); nobody reads tile row I in this phase, so the result is stored right away
        for (int J = wid; J < I; J += nw) {
            double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
            int K = J;
            for (; K + 1 < I; K += 2) {
                const double* xt = w.tile(K, J);
                const double* xu = w.tile(K + 1, J);
                const double* ma = mb + K * 64;
                wp.dmma(e0, e1, ma[lane], xt[bl0]);
                wp.dmma(o0, o1, ma[64 + lane], xu[bl0]);
                wp.dmma(e0, e1, ma[32 + lane], xt[bl1]);
                wp.dmma(o0, o1, ma[96 + lane], xu[bl1]);
            }
            if (K < I) {
                const double* xt = w.tile(K, J);
                const double* ma = mb + K * 64;
                wp.dmma(e0, e1, ma[lane], xt[bl0]);
                wp.dmma(e0, e1, ma[32 + lane], xt[bl1]);
            }
            *reinterpret_cast<d2*>(w.tile(I, J) + cl) = d2{-(e0 + o0), -(e1 + o1)};
        }
        if (I + 1 < T) m_row(I + 1);
        c.sync();
    }
}

/** sol <- K^-1 rhs with the inverted factor.  The caller has filled w.tb with the PERMUTED right-hand side (tb[a] = rhs[perm[a]],
 *  zeros beyond n) and synchronised; y = D^-1 X t, x = X^T y, sol[perm[a]] = x[a].  Lane l owns pair number l of every tile it
 *  streams (one LDS.128): rows of tiles (forward) and columns of tiles (backward) are dealt round-robin to the warps.  Ends with
 *  a block barrier.
 *  Tile-column outer / the warp's tile rows inner: the right-hand side pair of tile column J (forward) and the y
 *  value of tile row I (backward) are loaded ONCE per warp and step and shared by the warp's RS tile rows (columns), which
 *  advance together as independent accumulator chains.  The loops are rolled on purpose: this code runs once per ADMM trip and
 *  must stay resident in the 32 KB instruction cache next to the other phases of the three co-resident CTAs (a fully unrolled
 *  version was measured with 45 % of its stall samples on instruction fetch). */
template <int RS, int NW>
PMB_DEV void solve_rows(Cta& c, const Ws& w, const int* perm, double* sol)
{
    const Warp& wp = c.w;
    const int T = w.T, n = w.n, lane = c.lane(), wid = c.warp_id();
    const int h = lane >> 4, r8 = (lane >> 1) & 7, jj = lane & 1;
    const d2* tl = reinterpret_cast<const d2*>(w.tiles) + lane;          // pair `lane` of tile number t: tl[32 t]
    // The warp's tile rows are I_s = wid + NW s.  Row I_s has tiles J = 0 .. I_s, so over the J loop the set of active rows only
    // shrinks: segment `seg` covers the J for which exactly the rows s >= seg are active.  Inside a segment there is no
    // predicate at all: the loads of a step are independent and pipeline (with a per-row `if` the compiler emitted branches
    // and serialised the shared-memory latencies: 9.4 k instead of 4.7 k cycles per solve pair).  Rows beyond T (last warps)
    // are clamped to row T - 1: they compute garbage that is never stored.
    {
        const d2* tv = reinterpret_cast<const d2*>(w.tb) + 2 * h + jj;
        double a0[RS], a1[RS];
        int base[RS];
        PMB_UNROLL
        for (int s = 0; s < RS; ++s) {
            a0[s] = 0.0; a1[s] = 0.0;
            int I = wid + NW * s; I = I < T ? I : T - 1;
            base[s] = 32 * ((I * (I + 1)) / 2);
        }
        PMB_UNROLL
        for (int seg = 0; seg < RS; ++seg) {
            const int jlo = seg == 0 ? 0 : wid + NW * (seg - 1) + 1;
            int jhi = wid + NW * seg; jhi = jhi < T ? jhi : T - 1;
            PMB_NOUNROLL
            for (int J = jlo; J <= jhi; ++J) {
                const d2 t = tv[4 * J];
                PMB_UNROLL
                for (int s = seg; s < RS; ++s) {
                    const d2 x = tl[base[s] + 32 * J];
                    a0[s] = dm::fma(x.x, t.x, a0[s]);
                    a1[s] = dm::fma(x.y, t.y, a1[s]);
                }
            }
        }
        PMB_UNROLL
        for (int s = 0; s < RS; ++s) {
            const int I = wid + NW * s;
            double acc = a0[s] + a1[s];
            acc += wp.shfl_xor(acc, 1);
            acc += wp.shfl_xor(acc, 16);
            if (I < T && jj == 0 && h == 0) w.yb[8 * I + r8] = acc * w.dinv[8 * I + r8];
        }
    }
    c.sync();
    // backward: the warp's tile columns J_s = wid + NW s; over the I loop the set of active columns only grows
    {
        double b0[RS], b1[RS];
        PMB_UNROLL
        for (int s = 0; s < RS; ++s) { b0[s] = 0.0; b1[s] = 0.0; }
        PMB_UNROLL
        for (int seg = 0; seg < RS; ++seg) {
            const int ilo = wid + NW * seg;
            int ihi = ilo + NW - 1; ihi = (seg == RS - 1 || ihi >= T) ? T - 1 : ihi;
            PMB_NOUNROLL
            for (int I = ilo; I <= ihi; ++I) {
                const double yv = w.yb[8 * I + r8];
                const d2* row = tl + 32 * ((I * (I + 1)) / 2 + wid);
                PMB_UNROLL
                for (int s = 0; s <= seg; ++s) {
                    const d2 t = row[32 * NW * s];
                    b0[s] = dm::fma(t.x, yv, b0[s]);
                    b1[s] = dm::fma(t.y, yv, b1[s]);
                }
            }
        }
        PMB_UNROLL
        for (int s = 0; s < RS; ++s) {
            const int J = wid + NW * s;
            double acc0 = b0[s], acc1 = b1[s];
            PMB_UNROLL
            for (int off = 2; off <= 8; off <<= 1) { acc0 += wp.shfl_xor(acc0, off); acc1 += wp.shfl_xor(acc1, off); }
            if (J < T && r8 == 0) {
                const int a = 8 * J + 4 * h + 2 * jj;
                if (a < n) sol[perm[a]] = acc0;
                if (a + 1 < n) sol[perm[a + 1]] = acc1;
            }
        }
    }
    c.sync();
}

} // namespace fast
} // namespace pmb
