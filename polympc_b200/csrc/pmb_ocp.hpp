// pmb_ocp.hpp — batched transcription of a Chebyshev-collocated OCP: one warp per instance, lane k <-> collocation node k.
//
// Reference (dense path of src/control/continuous_ocp.hpp; line numbers of the reference):
//   sizes / layout 69-98, time nodes 45-66 & 147-159, seeding 690-735, equalities 738-766, inequalities 769-782,
//   equalities_linearised 794-878, inequalities_linearised 546-575, cost 1180-1207, cost_gradient 1209-1249,
//   cost_gradient_hessian 1253-1367, lagrangian_gradient 1957-1975, lagrangian_gradient_hessian 2097-2174.
// Layout: var = [X (NX*NN) | U (NU*NN) | P (NP)], node k at k*NX / VARX + k*NU; node 0 is the FINAL time (time nodes
// descend); lam = [lam_eq | lam_ineq | lam_box].  Matrices are column-major.
//
// Mapping: every node's functor evaluation (plain, first-order dual, second-order dual) is independent, so thread k of the
// CTA owns node k (NN <= 32: they all sit in warp 0): its NX collocation rows of the Jacobian and its gradient entries;
// Hessian columns are spread over all threads as (node, column) work items.  Zero fills, A^T*lam and other element-wise
// work are strided over the whole block.  Junction nodes of the spline are counted twice in the quadrature, in segment
// order, exactly like the reference loop; the cost sum keeps that fixed sequential order (shuffles inside warp 0) and is
// then broadcast, so every routine returns its scalar in every thread.
#pragma once
#include "pmb_cta.hpp"
#include "pmb_dual.hpp"
#include "pmb_cheb.hpp"

namespace pmb {

/** horizon a model wants at construction; the default leaves [0, 1].  Overloads for other model types are found by
 *  argument-dependent lookup (include/polympc_compat/polympc_compat.hpp: the horizon a reference-style class set in its
 *  constructor through ContinuousOCP::set_time_limits, continuous_ocp.hpp:147-159). */
template <class M> inline void model_time_limits(const M&, double&, double&) {}

template <class Model_, int P_, int S_>
struct Ocp {
    using Model = Model_;
    static constexpr int NX = Model::NX, NU = Model::NU, NP = Model::NP, ND = Model::ND, NG = Model::NG;
    static constexpr int P = P_, S = S_, NN = P_ * S_ + 1;
    static constexpr int VARX = NX * NN, VARU = NU * NN, N = VARX + VARU + NP;
    static constexpr int NUM_EQ = VARX, NUM_INEQ = NG * NN, M = NUM_EQ + NUM_INEQ, DUAL = M + N;
    static constexpr int NDIR = NX + NU + NP;
    static_assert(NN <= 32, "one warp per instance: at most 32 collocation nodes");
    using ad1 = Dual<double, NDIR>;
    using ad2 = Dual<ad1, NDIR>;   // the reference's ad2_scalar_t (full nesting)
    using ad2d = Dual<ad1, 1>;     // one Hessian column at a time (OcpEval::cost_gradient_hessian)

    Model model;
    double D[(P + 1) * (P + 1)];   // column-major
    double w[P + 1];
    double nodes[P + 1];
    double time_nodes[NN];
    double t_start, t_stop, ts;

    void init()
    {
        model.defaults();
        cheb_tables(P, nodes, D, w);
        double t0 = 0.0, tf = 1.0;
        model_time_limits(model, t0, tf);
        set_time_limits(t0, tf);
    }
    /** continuous_ocp.hpp:45-55, 147-159 */
    void set_time_limits(double t0, double tf)
    {
        t_start = t0; t_stop = tf;
        const double t_length = (t_stop - t_start) / (double)S;
        const double t_shift = t_length / 2;
        for (int i = 0; i < S; ++i)
            for (int k = 0; k <= P; ++k)
                time_nodes[i * P + k] = (t_length / 2) * nodes[P - k] + (t_start + t_shift + (double)i * t_length) * 1.0;
        for (int a = 0, b = NN - 1; a < b; ++a, --b) { const double t = time_nodes[a]; time_nodes[a] = time_nodes[b]; time_nodes[b] = t; }
        ts = (t_stop - t_start) / (double)(2 * S);
    }
    PMB_HD double d_(int i, int j) const { return D[i + j * (P + 1)]; }
};

/** index of direction a of node k inside var */
template <class O> PMB_HD int var_index(int k, int a)
{ return a < O::NX ? k * O::NX + a : (a < O::NX + O::NU ? O::VARX + k * O::NU + (a - O::NX) : O::VARX + O::VARU + (a - O::NX - O::NU)); }

/** Everything is inlined into the fused kernel on purpose: out-lining the evaluators (or only the cold exact-Hessian path)
 *  was measured slower — passing the problem descriptor by reference moves it from the constant bank to local memory. */
template <class O>
struct OcpEval {
    static constexpr int NX = O::NX, NU = O::NU, NP = O::NP, NG = O::NG, P = O::P, S = O::S, NN = O::NN;
    static constexpr int VARX = O::VARX, VARU = O::VARU, N = O::N, NUM_EQ = O::NUM_EQ, NUM_INEQ = O::NUM_INEQ, M = O::M, NDIR = O::NDIR;
    using ad1 = typename O::ad1;
    using ad2 = typename O::ad2;
    using ad2d = typename O::ad2d;

    /** segment and row of node k for the collocation rows: the later segment owns a junction node */
    PMB_DEV static void seg_of(int k, int& s, int& i) { s = k / P; if (s > S - 1) s = S - 1; i = k - s * P; }

    /** (D * X_seg)(row of node k): sequential ascending fused chain (continuous_ocp.hpp:747-751) */
    PMB_DEV static void diff_row(const O& o, const double* var, int k, double* DXk)
    {
        int s, i; seg_of(k, s, i);
        for (int n = 0; n < NX; ++n) {
            double acc = 0.0;
            for (int j = 0; j <= P; ++j) acc = dm::fma(o.d_(i, j), var[(s * P + j) * NX + n], acc);
            DXk[n] = acc;
        }
    }

    /** sequential quadrature sum in the reference's loop order; val = thread k's node value (k < NN, warp 0).  Every thread
     *  of the block returns the sum. */
    PMB_DEV static double quadrature_sum(Cta& c, const O& o, double val, double mayer_val)
    {
        double acc = 0.0;
        if (c.warp_id() == 0) {
            for (int s = 0; s < S; ++s)
                for (int k = 0; k <= P; ++k) {
                    const double v = c.w.shfl(val, s * P + k);
                    acc += (o.ts * o.w[k]) * v;
                }
            acc += c.w.shfl(mayer_val, 0);
        }
        return c.bcast(acc, 0);
    }

    /** quadrature coefficients of node k in loop order: c1 always, c2 only for spline junctions */
    PMB_DEV static void node_coeffs(const O& o, int k, double& c1, double& c2, bool& two)
    {
        two = false; c2 = 0.0;
        if (k == 0) c1 = o.ts * o.w[0];
        else if (k == NN - 1) c1 = o.ts * o.w[P];
        else if (k % P == 0) { c1 = o.ts * o.w[P]; c2 = o.ts * o.w[0]; two = true; }
        else c1 = o.ts * o.w[k % P];
    }

    static constexpr int NPA = NP > 0 ? NP : 1;   // array length of the parameter block
    template <class T>
    PMB_DEV static void load_plain(const double* var, int k, T* x, T* u, T* p)
    {
        for (int i = 0; i < NX; ++i) x[i] = T(var[k * NX + i]);
        for (int i = 0; i < NU; ++i) u[i] = T(var[VARX + k * NU + i]);
        for (int i = 0; i < NP; ++i) p[i] = T(var[VARX + VARU + i]);
    }
    PMB_DEV static void seed1(const double* var, int k, ad1* x, ad1* u, ad1* p)
    {
        for (int i = 0; i < NX; ++i) { x[i] = ad1(var[k * NX + i]); x[i].d[i] = 1.0; }
        for (int i = 0; i < NU; ++i) { u[i] = ad1(var[VARX + k * NU + i]); u[i].d[NX + i] = 1.0; }
        for (int i = 0; i < NP; ++i) { p[i] = ad1(var[VARX + VARU + i]); p[i].d[NX + NU + i] = 1.0; }
    }
    // ---- a3: cost (continuous_ocp.hpp:1180-1207) -----------------------------------------------------------------
    PMB_DEV static double cost(Cta& c, const O& o, const double* var, const double* d)
    {
        double acc = 0.0;
        if (c.warp_id() == 0) acc = cost_warp0(c, o, var, d);
        return c.bcast(acc, 0);
    }
    /** the same, called by (virtual) warp 0 only: no block barrier, the sum is returned in every lane of that warp */
    PMB_DEV static double cost_warp0(Cta& c, const O& o, const double* var, const double* d)
    {
        const int k = c.lane();
        double ci = 0.0, mv = 0.0;
        if (k < NN) {
            double x[NX], u[NU > 0 ? NU : 1], p[NPA] = {0.0};
            load_plain<double>(var, k, x, u, p);
            o.model.template lagrange<double>(x, u, p, d, o.time_nodes[k], ci);
            if (k == 0) o.model.template mayer<double>(x, u, p, d, o.time_nodes[0], mv);
        }
        double acc = 0.0;
        for (int s = 0; s < S; ++s)
            for (int kk = 0; kk <= P; ++kk) {
                const double v = c.w.shfl(ci, s * P + kk);
                acc += (o.ts * o.w[kk]) * v;
            }
        acc += c.w.shfl(mv, 0);
        return acc;
    }

    // ---- a4: equalities (continuous_ocp.hpp:738-766); lane k writes c[k*NX .. k*NX+NX) -----------------------------
    /** node k is evaluated by the calling thread (callers pass c.tid(), or the lane when one warp does the whole job) */
    PMB_DEV static void equalities(Cta& c, const O& o, const double* var, const double* d, double* ce, int k)
    {
        if (k < NN) {
            double x[NX], u[NU > 0 ? NU : 1], p[NPA] = {0.0}, f[NX], DXk[NX];
            load_plain<double>(var, k, x, u, p);
            for (int i = 0; i < NX; ++i) f[i] = 0.0;
            const double tk = o.time_nodes[k];
            o.model.template dynamics<double>(x, u, p, d, tk, f);
            diff_row(o, var, k, DXk);
            for (int i = 0; i < NX; ++i) ce[k * NX + i] = DXk[i] - o.ts * f[i];
        }
    }

    // ---- a5: inequalities (continuous_ocp.hpp:769-782) ------------------------------------------------------------
    PMB_DEV static void inequalities(Cta& c, const O& o, const double* var, const double* d, double* g, int k)
    {
        if (NG == 0) return;
        if (k < NN) {
            double x[NX], u[NU > 0 ? NU : 1], p[NPA] = {0.0}, gr[NG > 0 ? NG : 1];
            load_plain<double>(var, k, x, u, p);
            for (int i = 0; i < NG; ++i) gr[i] = 0.0;
            o.model.template ineq<double>(x, u, p, d, o.time_nodes[k], gr);
            for (int i = 0; i < NG; ++i) g[k * NG + i] = gr[i];
        }
    }

    // ---- a6: constraint linearisation (continuous_ocp.hpp:794-878, 546-575) ----------------------------------------
    /** writes c[NUM_EQ] (+ g[NUM_INEQ] behind it when NG > 0) and the rows x N Jacobian A (column-major, leading
     *  dimension ldA; rows = NUM_EQ or M).  A is zero-filled first like the reference (802). */
    PMB_DEV static void constraints_linearised(Cta& cta, const O& o, const double* var, const double* d, double* c, double* A,
                                               int ldA, bool with_ineq)
    {
        const int rows = with_ineq ? M : NUM_EQ;
        if (rows == ldA) { for (int e = cta.tid(); e < rows * N; e += cta.nthreads()) A[e] = 0.0; }
        else { for (int e = cta.tid(); e < rows * N; e += cta.nthreads()) { const int j = e / rows; A[(e - j * rows) + j * ldA] = 0.0; } }
        cta.sync();
        const int k = cta.tid();
        if (k < NN) {
            // D (x) I block row of this node
            if (k < NN - 1) {
                int s, i; seg_of(k, s, i);
                const int shift = s * P * NX;
                for (int j = 0; j <= P; ++j) {
                    const double dij = o.d_(i, j);
                    for (int r = 0; r < NX; ++r)
                        for (int q = 0; q < NX; ++q)
                            A[(shift + i * NX + r) + (shift + j * NX + q) * ldA] = dij * (r == q ? 1.0 : 0.0);
                }
            } else {
                // last block row = -reverse(first block row) (845-846)
                const int W = NX * (P + 1);
                for (int jp = 0; jp <= P; ++jp) {
                    const double d0 = o.d_(0, P - jp);
                    for (int r = 0; r < NX; ++r)
                        for (int q = 0; q < NX; ++q)
                            A[(VARX - NX + r) + (VARX - W + jp * NX + q) * ldA] = -(d0 * (r == q ? 1.0 : 0.0));
                }
            }
            ad1 x[NX], u[NU > 0 ? NU : 1], p[NPA], y[NX];
            seed1(var, k, x, u, p);
            for (int i = 0; i < NX; ++i) y[i] = ad1(0.0);
            const ad1 tk = ad1(o.time_nodes[k]);
            o.model.template dynamics<ad1>(x, u, p, d, tk, y);
            double DXk[NX];
            diff_row(o, var, k, DXk);
            for (int i = 0; i < NX; ++i) {
                double cv = -o.ts * y[i].v;
                cv += DXk[i];
                c[k * NX + i] = cv;
            }
            for (int i = 0; i < NX; ++i) {
                for (int j = 0; j < NX; ++j) A[(k * NX + i) + (k * NX + j) * ldA] -= o.ts * y[i].d[j];
                for (int j = 0; j < NU; ++j) A[(k * NX + i) + (VARX + k * NU + j) * ldA] -= o.ts * y[i].d[NX + j];
                for (int j = 0; j < NP; ++j) A[(k * NX + i) + (VARX + VARU + j) * ldA] -= o.ts * y[i].d[NX + NU + j];   // 870-872
            }
            if (NG > 0 && with_ineq) {
                ad1 gv[NG > 0 ? NG : 1];
                for (int i = 0; i < NG; ++i) gv[i] = ad1(0.0);
                o.model.template ineq<ad1>(x, u, p, d, o.time_nodes[k], gv);
                for (int i = 0; i < NG; ++i) {
                    c[NUM_EQ + k * NG + i] = gv[i].v;
                    const int row = NUM_EQ + k * NG + i;
                    for (int j = 0; j < NX; ++j) A[row + (k * NX + j) * ldA] = gv[i].d[j];
                    for (int j = 0; j < NU; ++j) A[row + (VARX + k * NU + j) * ldA] = gv[i].d[NX + j];
                    for (int j = 0; j < NP; ++j) A[row + (VARX + VARU + j) * ldA] = gv[i].d[NX + NU + j];
                }
            }
        }
        cta.sync();
    }

    // ---- a7: cost gradient (continuous_ocp.hpp:1209-1249) ----------------------------------------------------------
    /** shared scratch (doubles) of the gradient / Hessian evaluators: node values for the quadrature [NN + 1], and for
     *  NP > 0 the per-node parameter pieces that the reference accumulates ACROSS nodes in loop order: parameter gradient
     *  [(NN + 1) NP] and the (p, p) Hessian blocks of the cost and of the constraint curvature [2 NN NP^2]. */
    static constexpr int NV_GP = NN + 1, NV_PPC = NV_GP + (NN + 1) * NP, NV_PPL = NV_PPC + NN * NP * NP;
    static constexpr int NV_DOUBLES = NV_PPL + NN * NP * NP;

    /** grad[p_j] = sum over (segment, node) in loop order of (ts w_k) dL/dp_j, then + dMayer/dp_j (1228-1246); raw[k*NP + j]
     *  holds dL/dp_j of node k, raw[NN*NP + j] the Mayer derivative.  Thread j < NP does the chain. */
    PMB_DEV static void param_gradient_sum(Cta& c, const O& o, const double* raw, double* grad)
    {
        if (NP == 0) return;
        const int j = c.tid();
        if (j < NP) {
            double acc = 0.0;
            for (int s = 0; s < S; ++s)
                for (int k = 0; k <= P; ++k) acc += (o.ts * o.w[k]) * raw[(s * P + k) * NP + j];
            acc += raw[NN * NP + j];
            grad[VARX + VARU + j] = acc;
        }
    }

    PMB_DEV static double cost_gradient(Cta& c, const O& o, const double* var, const double* d, double* grad, double* nv)
    {
        const int k = c.tid();
        double lv = 0.0, mv = 0.0;
        if (k < NN) {
            ad1 x[NX], u[NU > 0 ? NU : 1], p[NPA], L;
            seed1(var, k, x, u, p);
            o.model.template lagrange<ad1>(x, u, p, d, o.time_nodes[k], L);
            lv = L.v;
            double c1, c2; bool two;
            node_coeffs(o, k, c1, c2, two);
            double g[NDIR];
            for (int i = 0; i < NX + NU; ++i) { g[i] = 0.0; g[i] += c1 * L.d[i]; if (two) g[i] += c2 * L.d[i]; }
            for (int j = 0; j < NP; ++j) nv[NV_GP + k * NP + j] = L.d[NX + NU + j];
            if (k == 0) {
                L = ad1(0.0);
                o.model.template mayer<ad1>(x, u, p, d, o.time_nodes[0], L);
                mv = L.v;
                for (int i = 0; i < NX + NU; ++i) g[i] += L.d[i];
                for (int j = 0; j < NP; ++j) nv[NV_GP + NN * NP + j] = L.d[NX + NU + j];
            }
            for (int i = 0; i < NX + NU; ++i) grad[var_index<O>(k, i)] = g[i];
        }
        if (NP > 0) { c.sync(); param_gradient_sum(c, o, nv + NV_GP, grad); }
        return quadrature_sum(c, o, lv, mv);
    }

    // ---- a8 / a10: cost gradient + Hessian, optionally plus the constraint curvature of the Lagrangian ---------------
    /** seed node k for ONE outer direction j: value part = first-order dual seeded in all NDIR directions, the single
     *  outer partial = derivative along direction j (continuous_ocp.hpp:690-735 restricted to one Hessian column). */
    PMB_DEV static void seed2_dir(const double* var, int k, int j, ad2d* x, ad2d* u, ad2d* p)
    {
        for (int i = 0; i < NX; ++i) {
            x[i].v = ad1(var[k * NX + i]); x[i].v.d[i] = 1.0;
            x[i].d[0] = ad1(0.0); if (j == i) x[i].d[0].v = 1.0;
        }
        for (int i = 0; i < NU; ++i) {
            u[i].v = ad1(var[VARX + k * NU + i]); u[i].v.d[NX + i] = 1.0;
            u[i].d[0] = ad1(0.0); if (j == NX + i) u[i].d[0].v = 1.0;
        }
        for (int i = 0; i < NP; ++i) {
            p[i].v = ad1(var[VARX + VARU + i]); p[i].v.d[NX + NU + i] = 1.0;
            p[i].d[0] = ad1(0.0); if (j == NX + NU + i) p[i].d[0].v = 1.0;
        }
    }

    /** cost_gradient_hessian (1253-1367); with lam != nullptr also adds, per node,
     *  sum_n (-lam_eq[k*NX+n]*ts) * Hess f_n + sum_n lam_ineq[k*NG+n] * Hess g_n  (2128-2173).  H is N x N, zero-filled first.
     *
     *  Mapping: the reference evaluates a nested dual with NDIR outer partials per node.  Every outer partial of a
     *  nested dual is computed independently of the others, so the CTA instead spreads the NN*NDIR (node, Hessian
     *  column) pairs over all its threads and evaluates the functors on Dual<ad1,1>: identical arithmetic per entry,
     *  (NX+NU)x less live state per thread, and every thread busy even when NN < 32.
     *
     *  NP > 0: entries of H that couple node k with the parameters have a single contributor (node k) and are stored
     *  directly; the parameter gradient and the (p, p) block are sums over all nodes in the reference's loop order — the
     *  per-node pieces go to shared scratch and NP (NP^2) threads run the sequential chains afterwards.  The Mayer term
     *  reproduces the reference's block bookkeeping (1352-1366): its (p, p) second derivatives are
     *  NOT added to the (p, p) block; instead hes(p_i, j), j < NP, is added a second time to H(p_i, column j). */
    PMB_DEV static double cost_gradient_hessian(Cta& cta, const O& o, const double* var, const double* d, const double* lam,
                                                double* grad, double* H, double* nv /* shared scratch, NV_DOUBLES */)
    {
        const int tid = cta.tid(), nt = cta.nthreads();
        constexpr int NB = NX + NU;
        for (int i = tid; i < N * N; i += nt) H[i] = 0.0;
        cta.sync();
        constexpr int ITEMS = NN * NDIR;
        PMB_NOUNROLL
        for (int item = tid; item < ITEMS; item += nt) {
            {   // one (node k, Hessian column c) work item
                const int k = item / NDIR, c = item - k * NDIR;
                ad2d x[NX], u[NU > 0 ? NU : 1], p[NPA];
                double Hc[NDIR];   // column c of the node's (NDIR x NDIR) block
                double g[NDIR];
                seed2_dir(var, k, c, x, u, p);
                {
                    ad2d L;
                    o.model.template lagrange<ad2d>(x, u, p, d, o.time_nodes[k], L);
                    if (c == 0) nv[k] = L.v.v;
                    double c1, c2; bool two;
                    node_coeffs(o, k, c1, c2, two);
                    for (int i = 0; i < NB; ++i) { g[i] = 0.0; g[i] += c1 * L.v.d[i]; if (two) g[i] += c2 * L.v.d[i]; }
                    if (c == 0) for (int j = 0; j < NP; ++j) nv[NV_GP + k * NP + j] = L.v.d[NB + j];
                    for (int r = 0; r < NDIR; ++r) {
                        if (r >= NB && c >= NB) { nv[NV_PPC + (k * NP + (c - NB)) * NP + (r - NB)] = L.d[0].d[r]; Hc[r] = 0.0; continue; }
                        double h = 0.0;
                        h += c1 * L.d[0].d[r];
                        if (two) h += c2 * L.d[0].d[r];
                        Hc[r] = h;
                    }
                }
                if (k == 0) {
                    ad2d Mv(0.0);
                    o.model.template mayer<ad2d>(x, u, p, d, o.time_nodes[0], Mv);
                    if (c == 0) nv[NN] = Mv.v.v;
                    for (int i = 0; i < NB; ++i) g[i] += Mv.v.d[i];
                    if (c == 0) for (int j = 0; j < NP; ++j) nv[NV_GP + NN * NP + j] = Mv.v.d[NB + j];
                    for (int r = 0; r < NDIR; ++r) {
                        if (r >= NB && c >= NB) continue;                       // no (p, p) contribution (quirk)
                        Hc[r] += Mv.d[0].d[r];
                        if (r >= NB && c < NP) Hc[r] += Mv.d[0].d[r];           // bottomLeftCorner<NP,NP> added on top of dxdp (quirk)
                    }
                }
                if (lam != nullptr) {
                    double hes[NDIR];
                    for (int i = 0; i < NDIR; ++i) hes[i] = 0.0;
                    {
                        ad2d xdot[NX];
                        const ad2d tk(o.time_nodes[k]);
                        o.model.template dynamics<ad2d>(x, u, p, d, tk, xdot);
                        for (int n = 0; n < NX; ++n) {
                            const double coeff = -lam[n + k * NX] * o.ts;
                            for (int r = 0; r < NDIR; ++r) hes[r] += coeff * xdot[n].d[0].d[r];
                        }
                    }
                    if (NG > 0) {
                        ad2d gv[NG > 0 ? NG : 1];
                        o.model.template ineq<ad2d>(x, u, p, d, o.time_nodes[k], gv);
                        for (int n = 0; n < NG; ++n) {
                            const double coeff = lam[n + k * NG + NUM_EQ];
                            for (int r = 0; r < NDIR; ++r) hes[r] += coeff * gv[n].d[0].d[r];
                        }
                    }
                    for (int i = 0; i < NDIR; ++i) {
                        if (i >= NB && c >= NB) nv[NV_PPL + (k * NP + (c - NB)) * NP + (i - NB)] = hes[i];
                        else Hc[i] += hes[i];
                    }
                }
                if (c == 0) for (int i = 0; i < NB; ++i) grad[var_index<O>(k, i)] = g[i];
                for (int r = 0; r < NDIR; ++r)
                    if (!(r >= NB && c >= NB)) H[var_index<O>(k, r) + var_index<O>(k, c) * N] = Hc[r];
            }
        }
        cta.sync();
        if (NP > 0) {
            param_gradient_sum(cta, o, nv + NV_GP, grad);
            // (p, p) block: cost part in loop order (junction nodes twice), then the constraint curvature node by node (2141-2160)
            for (int e = tid; e < NP * NP; e += nt) {
                const int cp = e / NPA, rp = e - cp * NPA;
                double acc = 0.0;
                for (int s = 0; s < S; ++s)
                    for (int k = 0; k <= P; ++k) acc += (o.ts * o.w[k]) * nv[NV_PPC + ((s * P + k) * NP + cp) * NP + rp];
                if (lam != nullptr) for (int k = 0; k < NN; ++k) acc += nv[NV_PPL + (k * NP + cp) * NP + rp];
                H[(VARX + VARU + rp) + (size_t)(VARX + VARU + cp) * N] = acc;
            }
            cta.sync();
        }
        // node values -> thread k (quadrature_sum expects thread k <-> node k)
        const double lv = tid < NN ? nv[tid] : 0.0;
        const double mv = tid == 0 ? nv[NN] : 0.0;
        return quadrature_sum(cta, o, lv, mv);
    }

    /** lag_grad = A^T lam_head + cost_grad + lam_box (continuous_ocp.hpp:1970-1974, 2112-2114); A is M x N, ld M */
    PMB_DEV static void lag_grad_from(Cta& c, const double* A, const double* lam, const double* cost_grad, double* lag_grad)
    {
        for (int j = c.tid(); j < N; j += c.nthreads()) {
            double acc = 0.0;
            const double* col = A + (size_t)j * M;
            for (int i = 0; i < M; ++i) acc = dm::fma(col[i], lam[i], acc);
            double v = acc;
            v += cost_grad[j];
            v += lam[M + j];
            lag_grad[j] = v;
        }
        c.sync();
    }

    /** a9: lagrangian_gradient (1957-1975) */
    PMB_DEV static double lagrangian_gradient(Cta& cta, const O& o, const double* var, const double* d, const double* lam,
                                              double* lag_grad, double* cost_grad, double* g, double* A, double* nv)
    {
        const double c = cost_gradient(cta, o, var, d, cost_grad, nv);
        constraints_linearised(cta, o, var, d, g, A, M, true);
        lag_grad_from(cta, A, lam, cost_grad, lag_grad);
        return c;
    }
    /** a10: lagrangian_gradient_hessian (2097-2174) */
    PMB_DEV static double lagrangian_gradient_hessian(Cta& cta, const O& o, const double* var, const double* d, const double* lam,
                                                      double* lag_grad, double* H, double* cost_grad, double* g, double* A, double* nv)
    {
        const double c = cost_gradient_hessian(cta, o, var, d, lam, cost_grad, H, nv);
        constraints_linearised(cta, o, var, d, g, A, M, true);
        lag_grad_from(cta, A, lam, cost_grad, lag_grad);
        return c;
    }
};

} // namespace pmb
