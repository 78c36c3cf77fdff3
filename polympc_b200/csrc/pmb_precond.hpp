// pmb_precond.hpp — RuizEquilibration of the reference on one CTA per instance, and the LSFilter acceptance test of its
// filter line search.
//
// Reference: src/solvers/qp_preconditioners.hpp — RuizEquilibration::compute 151-220 (DENSE) / 236-300 (SPARSE), unscale of
// the solution 364-369, unscale of the QP data 372-404; src/solvers/line_search.hpp:30-98 (LSFilter).
//
// Arithmetic: every coefficient is produced by the same fp64 operations in the same order as oracle/precond.hpp (which
// restates what Eigen's lazy diagonal products evaluate); maxima are exact and order free, the first-coefficient NaN rule of
// lpNorm<Infinity> / maxCoeff is applied afterwards (cf. norm_inf_cta); the one sum (the mean of the column norms) is taken
// in ascending order by every thread.  H (N x N) and A (M x N) are column-major and live in global memory (L2-resident
// between the passes), D / E / c as well: the SQP scratch in shared memory is aliased by the KKT factor during the QP.
#pragma once
#include "pmb_cta.hpp"
#include "pmb_detmath.h"

namespace pmb {

enum { PRECOND_IDENTITY = 0, PRECOND_RUIZ_DENSE = 1, PRECOND_RUIZ_SPARSE = 2 };
enum { LS_L1_MERIT = 0, LS_FILTER = 1 };
constexpr int FILTER_CAP = 16;
constexpr int FILTER_DOUBLES = 1 + 2 * FILTER_CAP;

/** NC, MC: compile-time sizes (the fused SQP kernel), or NC = 0: sizes n, m given at run time (the stand-alone kernel).
 *  Per instance: state st[N + M + 1] = D, E, c in global memory; scratch sc[N + M] in shared memory during compute. */
template <int NC, int MC>
struct RuizCta {

    PMB_DEV static double mx(double m, double v) { return m < v ? v : m; }

    /** |.|_inf of column j of a column-major rows x cols matrix, by the calling warp (lanes stride over the rows: coalesced);
     *  result in every lane */
    PMB_DEV static double col_absmax(const Warp& w, const double* X, int rows, int j)
    {
        const double* col = X + (size_t)j * rows;
        double m = 0.0;
        for (int i = w.lane(); i < rows; i += 32) { const double v = dm::fabs(col[i]); if (v > m) m = v; }
        for (int off = 16; off >= 1; off >>= 1) { const double o = w.shfl_xor(m, off); if (o > m) m = o; }
        const double a0 = dm::fabs(col[0]);
        return a0 != a0 ? a0 : m;
    }

    /** compute(): H, h, A, Al, Au, l, u are scaled in place; st = {D, E, c} (global), sc = SCRATCH_DOUBLES of shared memory */
    PMB_DEV static void compute(Cta& c, int n, int m, int variant, double* H, double* h, double* A, double* Al, double* Au, double* l, double* u,
                                double* st, double* sc)
    {
        const int tid = c.tid(), nt = c.nthreads(), nw = c.nwarps(), wid = c.warp_id();
        const int N = NC > 0 ? NC : n, M = NC > 0 ? MC : m;
        const bool sparse = variant == PRECOND_RUIZ_SPARSE;
        double* D = st;
        double* E = st + N;
        double* mD = sc;
        double* mE = sc + N;
        const double approx_zero = sparse ? 1e-4 : DBL_EPSILON;
        const double tolerance = 1e-3;
        double cs = 1.0;
        for (int k = tid; k < N; k += nt) D[k] = 1.0;
        for (int k = tid; k < M; k += nt) E[k] = 1.0;
        double scaling_norm = 10 * tolerance;
        for (int iter = 0; iter < 4 && (1.0 - scaling_norm) >= tolerance; ++iter) {
            // row norms of A (one thread per row: neighbouring threads read neighbouring rows), column norms of H and A
            for (int i = tid; i < M; i += nt) {
                double m = dm::fabs(A[i]);
                for (int j = 1; j < N; ++j) m = mx(m, dm::fabs(A[i + (size_t)j * M]));
                mE[i] = m;
            }
            for (int j = wid; j < N; j += nw) {
                const double mh = col_absmax(c.w, H, N, j);
                const double ma = M > 0 ? col_absmax(c.w, A, M, j) : 0.0;
                if (c.lane() == 0) mD[j] = mx(mh, ma);
            }
            c.sync();
            // maxCoeff / minCoeff of the two norm vectors
            const double NEG = -dm::inf();
            double r[4] = {NEG, NEG, NEG, NEG};        // max D, -min D, max E, -min E
            for (int k = tid; k < N; k += nt) { const double v = mD[k]; if (v > r[0]) r[0] = v; if (-v > r[1]) r[1] = -v; }
            for (int k = tid; k < M; k += nt) { const double v = mE[k]; if (v > r[2]) r[2] = v; if (-v > r[3]) r[3] = -v; }
            c.max_all<4>(r);
            double mxD = r[0], mnD = -r[1], mxE = r[2], mnE = -r[3];
            if (mD[0] != mD[0]) { mxD = mD[0]; mnD = mD[0]; }
            if (M > 0) { if (mE[0] != mE[0]) { mxE = mE[0]; mnE = mE[0]; } }
            else { mxE = NEG; mnE = dm::inf(); }
            scaling_norm = mx(mxD, mxE);
            c.sync();                                   // every thread has read mD[0] / mE[0]
            for (int k = tid; k < N; k += nt) {
                double v = mD[k];
                if (mnD < approx_zero && v < approx_zero) v = 1.0;
                v = 1.0 / dm::sqrt(v);
                mD[k] = v;
                D[k] = D[k] * v;
                h[k] = h[k] * v;
            }
            for (int k = tid; k < M; k += nt) {
                double v = mE[k];
                if (mnE < approx_zero && v < approx_zero) v = 1.0;
                v = 1.0 / dm::sqrt(v);
                mE[k] = v;
                E[k] = E[k] * v;
            }
            c.sync();
            for (int e = tid; e < N * N; e += nt) { const int j = e / N, i = e - j * N; H[e] = (mD[i] * H[e]) * mD[j]; }
            if (sparse) for (int e = tid; e < M * N; e += nt) { const int j = e / M, i = e - j * M; A[e] = mE[i] * (A[e] * mD[j]); }
            else        for (int e = tid; e < M * N; e += nt) { const int j = e / M, i = e - j * M; A[e] = (mE[i] * A[e]) * mD[j]; }
            c.sync();
            // gamma = 1 / max(mean of the column norms of H, |h|_inf)
            for (int j = wid; j < N; j += nw) { const double mh = col_absmax(c.w, H, N, j); if (c.lane() == 0) mD[j] = mh; }
            double hm[2] = {0.0, 0.0};
            for (int i = tid; i < N; i += nt) { const double v = dm::fabs(h[i]); if (v > hm[0]) hm[0] = v; if (i == 0 && v != v) hm[1] = 1.0; }
            c.sync();                                   // mD is published
            c.max_all<2>(hm);
            double h_inf = hm[1] > 0.0 ? dm::fabs(h[0]) : hm[0];
            double sum = 0.0;
            for (int j = 0; j < N; ++j) sum += mD[j];
            const double mean = sum / (double)N;
            if (!sparse) h_inf = h_inf > approx_zero ? h_inf : 1.0;
            const double gamma = 1.0 / mx(mean, h_inf);
            c.sync();                                   // mD is rewritten by the next pass
            for (int e = tid; e < N * N; e += nt) H[e] *= gamma;
            for (int k = tid; k < N; k += nt) h[k] *= gamma;
            cs *= gamma;
            c.sync();
        }
        for (int k = tid; k < M; k += nt) { const double e = E[k]; Au[k] = Au[k] * e; Al[k] = Al[k] * e; }
        for (int k = tid; k < N; k += nt) { const double di = 1.0 / D[k]; l[k] = l[k] * di; u[k] = u[k] * di; }
        if (tid == 0) st[N + M] = cs;
        c.sync();
    }

    /** unscale(x, y) then unscale(H, h, A, Al, Au, l, u); x / y may be null */
    PMB_DEV static void unscale(Cta& c, int n, int m, const double* st, double* H, double* h, double* A, double* Al, double* Au, double* l, double* u,
                                double* x, double* y)
    {
        const int tid = c.tid(), nt = c.nthreads();
        const int N = NC > 0 ? NC : n, M = NC > 0 ? MC : m;
        const double* D = st;
        const double* E = st + N;
        const double ic = 1 / st[N + M];
        if (x && y) {
            for (int k = tid; k < N; k += nt) x[k] = x[k] * D[k];
            for (int k = tid; k < M; k += nt) y[k] = ic * (y[k] * E[k]);
            for (int k = tid; k < N; k += nt) y[M + k] = ic * (y[M + k] * (1.0 / D[k]));
        }
        for (int e = tid; e < N * N; e += nt) { const int j = e / N, i = e - j * N; H[e] = ((ic * (1.0 / D[i])) * H[e]) * (1.0 / D[j]); }
        for (int e = tid; e < M * N; e += nt) { const int j = e / M, i = e - j * M; A[e] = ((1.0 / E[i]) * A[e]) * (1.0 / D[j]); }
        for (int k = tid; k < N; k += nt) { h[k] = ic * (h[k] * (1.0 / D[k])); l[k] = l[k] * D[k]; u[k] = u[k] * D[k]; }
        for (int k = tid; k < M; k += nt) { const double ei = 1.0 / E[k]; Au[k] = Au[k] * ei; Al[k] = Al[k] * ei; }
        c.sync();
    }
};

/** LSFilter (line_search.hpp:30-98) on the per-instance state f = [size, cost[CAP], violation[CAP]] (entry 0 = newest) */
PMB_DEV bool filter_is_acceptable(const double* f, double beta, double cost, double viol)
{
    const int n = (int)f[0];
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        const double fc = f[1 + k], fv = f[1 + FILTER_CAP + k];
        if (((fc - beta * fv) <= cost) && ((fv - beta * fv) <= viol)) ok = false;
    }
    return ok;
}

/** LSFilter::add by thread 0, between two block syncs (the other threads may still be reading the filter) */
PMB_DEV void filter_add(Cta& c, double* f, int max_depth, double cost, double viol)
{
    c.sync();
    if (c.tid() == 0) {
        int n = (int)f[0];
        if (n < max_depth) {
            int w = 0;
            for (int k = 0; k < n; ++k) {
                const double fc = f[1 + k], fv = f[1 + FILTER_CAP + k];
                if (!((fc >= cost) && (fv >= viol))) { f[1 + w] = fc; f[1 + FILTER_CAP + w] = fv; ++w; }
            }
            n = w;
        } else {
            --n;
        }
        for (int k = n; k > 0; --k) { f[1 + k] = f[k]; f[1 + FILTER_CAP + k] = f[FILTER_CAP + k]; }
        f[1] = cost; f[1 + FILTER_CAP] = viol;
        f[0] = (double)(n + 1);
    }
    c.sync();
}

} // namespace pmb
