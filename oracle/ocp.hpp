// oracle/ocp.hpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Dense-path restatement of reference src/control/continuous_ocp.hpp (ContinuousOCP<OCP, Spline<Chebyshev<P>,S>, DENSE>).
// All matrices are column-major like Eigen.  Layout (continuous_ocp.hpp:69-98):
//   var = [X (NX*nn) | U (NU*nn) | P (NP)],   node k at k*NX / VARX + k*NU,  node 0 = final time (time_nodes descend)
//   lam = [lam_eq (NX*nn) | lam_ineq (NG*nn) | lam_box (N)]
#pragma once
#include "cheb.hpp"
#include "models.hpp"
#include <vector>

namespace orc {

template <class Model, int P_, int S_>
struct Ocp {
    static constexpr int NX = Model::NX, NU = Model::NU, NP = Model::NP, ND = Model::ND, NG = Model::NG;
    static constexpr int P = P_, S = S_, NN = P_ * S_ + 1;
    static constexpr int VARX = NX * NN, VARU = NU * NN, N = VARX + VARU + NP;
    static constexpr int NUM_EQ = VARX, NUM_INEQ = NG * NN, M = NUM_EQ + NUM_INEQ, DUAL = M + N;
    static constexpr int ND_ = NX + NU + NP;  // number of AD directions
    using ad1 = AD<double, ND_>;
    using ad2 = AD<ad1, ND_>;

    Model model;
    ChebTables tab;
    double t_start = 0.0, t_stop = 1.0;
    double time_nodes[NN];

    Ocp() : tab(cheb_tables(P)) { set_time_limits(0.0, 1.0); }

    /** continuous_ocp.hpp:45-55, 147-159 */
    void set_time_limits(double t0, double tf)
    {
        t_start = t0; t_stop = tf;
        const double t_length = (t_stop - t_start) / (double)S;
        const double t_shift = t_length / 2;
        for (int i = 0; i < S; ++i)
            for (int k = 0; k <= P; ++k)  // m_nodes.reverse()
                time_nodes[i * P + k] = (t_length / 2) * tab.nodes[P - k] + (t_start + t_shift + (double)i * t_length) * 1.0;
        for (int a = 0, b = NN - 1; a < b; ++a, --b) { const double t = time_nodes[a]; time_nodes[a] = time_nodes[b]; time_nodes[b] = t; }
    }
    double t_scale() const { return (t_stop - t_start) / (double)(2 * S); }

    /** D * X_seg^T for every segment, later segment overwrites the junction row (continuous_ocp.hpp:747-751) */
    void diff_states(const double* var, double* DX /*NN x NX, DX[k*NX+n]*/) const
    {
        for (int s = 0; s < S; ++s)
            for (int i = 0; i <= P; ++i)
                for (int n = 0; n < NX; ++n) {
                    double acc = 0.0;
                    for (int j = 0; j <= P; ++j) acc = dm::fma(tab.d(i, j), var[(s * P + j) * NX + n], acc);
                    DX[(s * P + i) * NX + n] = acc;
                }
    }

    /** a3: continuous_ocp.hpp:1180-1207 */
    void cost(const double* var, const double* d, double& cost_out) const
    {
        double c = 0.0, ci = 0.0;
        const double ts = t_scale();
        for (int s = 0; s < S; ++s) {
            const int shift = s * P;
            for (int k = 0; k <= P; ++k) {
                model.template lagrange<double>(var + (k + shift) * NX, var + (k + shift) * NU + VARX, var + VARX + VARU, d,
                                                time_nodes[k + shift], ci);
                c += ts * tab.w[k] * ci;
            }
        }
        ci = 0.0;
        model.template mayer<double>(var, var + VARX, var + VARX + VARU, d, time_nodes[0], ci);
        c += ci;
        cost_out = c;
    }

    /** a4: continuous_ocp.hpp:738-766 */
    void equalities(const double* var, const double* d, double* c) const
    {
        double DX[NN * NX];
        diff_states(var, DX);
        const double ts = t_scale();
        for (int k = 0; k < NN; ++k) {
            double f[NX];
            for (int i = 0; i < NX; ++i) f[i] = 0.0;
            const double tk = time_nodes[k];
            model.template dynamics<double>(var + k * NX, var + VARX + k * NU, var + VARX + VARU, d, tk, f);
            for (int i = 0; i < NX; ++i) c[k * NX + i] = DX[k * NX + i] - ts * f[i];
        }
    }

    /** a5: continuous_ocp.hpp:769-782 */
    void inequalities(const double* var, const double* d, double* g) const
    {
        for (int k = 0; k < NN; ++k) {
            double gr[NG > 0 ? NG : 1];
            for (int i = 0; i < NG; ++i) gr[i] = 0.0;
            model.template ineq<double>(var + k * NX, var + VARX + k * NU, var + VARX + VARU, d, time_nodes[k], gr);
            for (int i = 0; i < NG; ++i) g[k * NG + i] = gr[i];
        }
    }

    void seed1(const double* var, int k, ad1* x, ad1* u, ad1* p) const
    {
        for (int i = 0; i < NX; ++i) { x[i] = ad1(var[k * NX + i]); x[i].d[i] = 1.0; }
        for (int i = 0; i < NU; ++i) { u[i] = ad1(var[VARX + k * NU + i]); u[i].d[NX + i] = 1.0; }
        for (int i = 0; i < NP; ++i) { p[i] = ad1(var[VARX + VARU + i]); p[i].d[NX + NU + i] = 1.0; }
    }
    /** seeding of the nested type: continuous_ocp.hpp:690-735 */
    void seed2(const double* var, int k, ad2* x, ad2* u, ad2* p) const
    {
        auto mk = [](double val, int idx) {
            ad2 a;
            a.v = ad1(val); a.v.d[idx] = 1.0;                     // value().derivatives() = Unit(idx)
            for (int j = 0; j < ND_; ++j) a.d[j] = ad1(0.0);        // derivatives()(j).derivatives() = Zero
            a.d[idx].v = 1.0;                                      // derivatives() = Unit(idx)
            return a;
        };
        for (int i = 0; i < NX; ++i) x[i] = mk(var[k * NX + i], i);
        for (int i = 0; i < NU; ++i) u[i] = mk(var[VARX + k * NU + i], NX + i);
        for (int i = 0; i < NP; ++i) p[i] = mk(var[VARX + VARU + i], NX + NU + i);
    }

    /** a6: continuous_ocp.hpp:794-878.  A is rows x N column-major with leading dimension ldA (rows of the full
     *  Jacobian), only the first NUM_EQ rows are written. */
    void equalities_linearised(const double* var, const double* d, double* c, double* A, int ldA) const
    {
        for (int j = 0; j < N; ++j) for (int i = 0; i < NUM_EQ; ++i) A[i + j * ldA] = 0.0;
        double DX[NN * NX];
        diff_states(var, DX);
        const double ts = t_scale();
        // D (x) I blocks, rows i < P of every segment (817-827)
        for (int s = 0; s < S; ++s)
            for (int i = 0; i < P; ++i)
                for (int j = 0; j <= P; ++j) {
                    const int shift = s * P * NX;
                    for (int r = 0; r < NX; ++r)
                        for (int q = 0; q < NX; ++q)
                            A[(shift + i * NX + r) + (shift + j * NX + q) * ldA] = tab.d(i, j) * (r == q ? 1.0 : 0.0);
                }
        // last block row = -reverse(first block row) (845-846)
        {
            const int W = NX * (P + 1);
            double blk[NX * NX * (P_ + 1)];
            for (int r = 0; r < NX; ++r) for (int q = 0; q < W; ++q) blk[r + q * NX] = A[r + q * ldA];
            for (int r = 0; r < NX; ++r)
                for (int q = 0; q < W; ++q)
                    A[(VARX - NX + r) + (VARX - W + q) * ldA] = -blk[(NX - 1 - r) + (W - 1 - q) * NX];
        }
        ad1 x[NX], u[NU > 0 ? NU : 1], p[NP > 0 ? NP : 1], y[NX];
        for (int k = 0; k < NN; ++k) {
            seed1(var, k, x, u, p);
            for (int i = 0; i < NX; ++i) y[i] = ad1(0.0);
            const ad1 tk = ad1(time_nodes[k]);
            model.template dynamics<ad1>(x, u, p, d, tk, y);
            for (int i = 0; i < NX; ++i) {
                c[k * NX + i] = -ts * y[i].v;
                c[k * NX + i] += DX[k * NX + i];
            }
            for (int i = 0; i < NX; ++i) {
                for (int j = 0; j < NX; ++j) A[(k * NX + i) + (k * NX + j) * ldA] -= ts * y[i].d[j];
                for (int j = 0; j < NU; ++j) A[(k * NX + i) + (VARX + k * NU + j) * ldA] -= ts * y[i].d[NX + j];
                for (int j = 0; j < NP; ++j) A[(k * NX + i) + (VARX + VARU + j) * ldA] -= ts * y[i].d[NX + NU + j];
            }
        }
    }

    /** continuous_ocp.hpp:546-575; writes rows NUM_EQ.. of the full Jacobian */
    void inequalities_linearised(const double* var, const double* d, double* g, double* A, int ldA) const
    {
        if (NG == 0) return;
        for (int j = 0; j < N; ++j) for (int i = 0; i < NUM_INEQ; ++i) A[NUM_EQ + i + j * ldA] = 0.0;
        ad1 x[NX], u[NU > 0 ? NU : 1], p[NP > 0 ? NP : 1], gv[NG > 0 ? NG : 1];
        for (int k = 0; k < NN; ++k) {
            seed1(var, k, x, u, p);
            for (int i = 0; i < NG; ++i) gv[i] = ad1(0.0);
            model.template ineq<ad1>(x, u, p, d, time_nodes[k], gv);
            for (int i = 0; i < NG; ++i) {
                g[k * NG + i] = gv[i].v;
                const int row = NUM_EQ + k * NG + i;
                for (int j = 0; j < NX; ++j) A[row + (k * NX + j) * ldA] = gv[i].d[j];
                for (int j = 0; j < NU; ++j) A[row + (VARX + k * NU + j) * ldA] = gv[i].d[NX + j];
                for (int j = 0; j < NP; ++j) A[row + (VARX + VARU + j) * ldA] = gv[i].d[NX + NU + j];
            }
        }
    }

    /** a7: continuous_ocp.hpp:1209-1249 */
    void cost_gradient(const double* var, const double* d, double& cost_out, double* grad) const
    {
        double c = 0.0;
        for (int i = 0; i < N; ++i) grad[i] = 0.0;
        const double ts = t_scale();
        ad1 x[NX], u[NU > 0 ? NU : 1], p[NP > 0 ? NP : 1], L;
        for (int s = 0; s < S; ++s) {
            const int shift = s * P;
            for (int k = 0; k <= P; ++k) {
                const int nd = k + shift;
                seed1(var, nd, x, u, p);
                model.template lagrange<ad1>(x, u, p, d, time_nodes[nd], L);
                const double coeff = ts * tab.w[k];
                c += coeff * L.v;
                for (int i = 0; i < NX; ++i) grad[nd * NX + i] += coeff * L.d[i];
                for (int i = 0; i < NU; ++i) grad[VARX + nd * NU + i] += coeff * L.d[NX + i];
                for (int i = 0; i < NP; ++i) grad[VARX + VARU + i] += coeff * L.d[NX + NU + i];
            }
        }
        seed1(var, 0, x, u, p);
        L = ad1(0.0);
        model.template mayer<ad1>(x, u, p, d, time_nodes[0], L);
        c += L.v;
        for (int i = 0; i < NX; ++i) grad[i] += L.d[i];
        for (int i = 0; i < NU; ++i) grad[VARX + i] += L.d[NX + i];
        for (int i = 0; i < NP; ++i) grad[VARX + VARU + i] += L.d[NX + NU + i];
        cost_out = c;
    }

    /** scatter a (NX+NU+NP)^2 node Hessian `hes` (column-major, hes[r + c*ND_]) scaled by coeff into H at node nd.
     *  continuous_ocp.hpp:1305-1325 / 2161-2172 */
    void scatter_hes(double* H, int nd, const double* hes, double coeff, bool scaled) const
    {
        auto idx = [&](int a) { return a < NX ? nd * NX + a : (a < NX + NU ? VARX + nd * NU + (a - NX) : VARX + VARU + (a - NX - NU)); };
        for (int cc = 0; cc < ND_; ++cc)
            for (int r = 0; r < ND_; ++r) {
                const double v = scaled ? coeff * hes[r + cc * ND_] : hes[r + cc * ND_];
                H[idx(r) + idx(cc) * N] += v;
            }
    }

    /** a8: continuous_ocp.hpp:1253-1367 */
    void cost_gradient_hessian(const double* var, const double* d, double& cost_out, double* grad, double* H) const
    {
        double c = 0.0;
        for (int i = 0; i < N; ++i) grad[i] = 0.0;
        for (int i = 0; i < N * N; ++i) H[i] = 0.0;
        const double ts = t_scale();
        ad2 x[NX], u[NU > 0 ? NU : 1], p[NP > 0 ? NP : 1], L;
        double hes[ND_ * ND_];
        for (int s = 0; s < S; ++s) {
            const int shift = s * P;
            for (int k = 0; k <= P; ++k) {
                const int nd = k + shift;
                seed2(var, nd, x, u, p);
                model.template lagrange<ad2>(x, u, p, d, time_nodes[nd], L);
                const double coeff = ts * tab.w[k];
                c += coeff * L.v.v;
                for (int i = 0; i < NX; ++i) grad[nd * NX + i] += coeff * L.v.d[i];
                for (int i = 0; i < NU; ++i) grad[VARX + nd * NU + i] += coeff * L.v.d[NX + i];
                for (int i = 0; i < NP; ++i) grad[VARX + VARU + i] += coeff * L.v.d[NX + NU + i];
                for (int i = 0; i < ND_; ++i) for (int r = 0; r < ND_; ++r) hes[r + i * ND_] = L.d[i].d[r];  // hes.col(i)
                scatter_hes(H, nd, hes, coeff, true);
            }
        }
        // Mayer term at node 0 (1345-1366)
        seed2(var, 0, x, u, p);
        ad2 Mv(0.0);
        model.template mayer<ad2>(x, u, p, d, time_nodes[0], Mv);
        c += Mv.v.v;
        for (int i = 0; i < NX; ++i) grad[i] += Mv.v.d[i];
        for (int i = 0; i < NU; ++i) grad[VARX + i] += Mv.v.d[NX + i];
        for (int i = 0; i < NP; ++i) grad[VARX + VARU + i] += Mv.v.d[NX + NU + i];
        for (int i = 0; i < ND_; ++i) for (int r = 0; r < ND_; ++r) hes[r + i * ND_] = Mv.d[i].d[r];
        // Mayer blocks exactly as the reference books them (1352-1366): xx, uu, xu, ux, xp, px, up, pu are
        // added; the (p, p) second derivatives are NOT (the reference writes `bottomLeftCorner<NP,NP>() +=
        // hes.bottomLeftCorner<NP,NP>()`), instead hes(p_i, j), j < NP, lands a second time on H(p_i, column j).
        {
            constexpr int NB = NX + NU;
            auto idx = [&](int a) { return a < NX ? a : (a < NB ? VARX + (a - NX) : VARX + VARU + (a - NB)); };   // node 0
            for (int cc = 0; cc < ND_; ++cc)
                for (int r = 0; r < ND_; ++r) {
                    if (r >= NB && cc >= NB) continue;
                    H[idx(r) + idx(cc) * N] += hes[r + cc * ND_];
                }
            for (int j = 0; j < NP; ++j)
                for (int i = 0; i < NP; ++i) H[(N - NP + i) + j * N] += hes[(ND_ - NP + i) + j * ND_];
        }
        cost_out = c;
    }

    /** lag_grad = A^T lam_head + cost_grad + lam_box  (continuous_ocp.hpp:1970-1974, 2112-2114) */
    void lag_grad_from(const double* A, const double* lam, const double* cost_grad, double* lag_grad) const
    {
        for (int j = 0; j < N; ++j) {
            double acc = 0.0;
            for (int i = 0; i < M; ++i) acc = dm::fma(A[i + j * M], lam[i], acc);
            double v = acc;
            v += cost_grad[j];
            v += lam[M + j];
            lag_grad[j] = v;
        }
    }

    /** a9: continuous_ocp.hpp:1957-1975 */
    void lagrangian_gradient(const double* var, const double* d, const double* lam, double& cost_out, double* lag_grad,
                             double* cost_grad, double* g, double* A) const
    {
        cost_gradient(var, d, cost_out, cost_grad);
        equalities_linearised(var, d, g, A, M);
        inequalities_linearised(var, d, g + NUM_EQ, A, M);
        lag_grad_from(A, lam, cost_grad, lag_grad);
    }

    /** a10: continuous_ocp.hpp:2097-2174 (cost_scale == 1) */
    void lagrangian_gradient_hessian(const double* var, const double* d, const double* lam, double& cost_out, double* lag_grad,
                                     double* H, double* cost_grad, double* g, double* A) const
    {
        cost_gradient_hessian(var, d, cost_out, cost_grad, H);
        equalities_linearised(var, d, g, A, M);
        inequalities_linearised(var, d, g + NUM_EQ, A, M);
        lag_grad_from(A, lam, cost_grad, lag_grad);

        const double ts = t_scale();
        ad2 x[NX], u[NU > 0 ? NU : 1], p[NP > 0 ? NP : 1], xdot[NX], gv[NG > 0 ? NG : 1];
        double hes[ND_ * ND_];
        for (int k = 0; k < NN; ++k) {
            seed2(var, k, x, u, p);
            for (int i = 0; i < ND_ * ND_; ++i) hes[i] = 0.0;
            const ad2 tk(time_nodes[k]);
            model.template dynamics<ad2>(x, u, p, d, tk, xdot);
            for (int n = 0; n < NX; ++n) {
                const double coeff = -lam[n + k * NX] * ts;
                for (int i = 0; i < ND_; ++i)
                    for (int r = 0; r < ND_; ++r) hes[r + i * ND_] += coeff * xdot[n].d[i].d[r];
            }
            if (NG > 0) {
                model.template ineq<ad2>(x, u, p, d, time_nodes[k], gv);
                for (int n = 0; n < NG; ++n) {
                    const double coeff = lam[n + k * NG + NUM_EQ];
                    for (int i = 0; i < ND_; ++i)
                        for (int r = 0; r < ND_; ++r) hes[r + i * ND_] += coeff * gv[n].d[i].d[r];
                }
            }
            scatter_hes(H, k, hes, 1.0, false);
        }
    }
};

} // namespace orc
