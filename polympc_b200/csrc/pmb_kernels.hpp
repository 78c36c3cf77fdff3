// pmb_kernels.hpp — kernel bodies (one warp = one block = one OCP/QP instance unless noted) launched by pmb_capi.cu.
// Every body is a struct with THREADS, NAME, EMU_STACK_BYTES and a static run(); see pmb_warp.hpp.
#pragma once
#include "pmb_ocp.hpp"
#include "pmb_qp.hpp"
#include "pmb_sqp.hpp"

namespace pmb {

// ---- transcription operators (C ABI pmb_ocp_*) ----------------------------------------------------------------------
enum OcpMode { OCP_COST = 0, OCP_EQ, OCP_INEQ, OCP_EQ_LIN, OCP_COST_GRAD, OCP_COST_GRAD_HESS, OCP_LAG_GRAD, OCP_LAG_GRAD_HESS };

struct OcpIo {
    const double *var, *d, *lam;
    double *cost, *c, *g, *jac, *grad, *hess, *lag_grad;
};

template <class O, int MODE>
struct OcpEvalBody {
    static constexpr int THREADS = 64;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr const char* NAME = "ocp_eval";
    static constexpr size_t EMU_STACK_BYTES = 4u << 20;
    static constexpr size_t SMEM = (Cta::SCRATCH_DOUBLES + OcpEval<O>::NV_DOUBLES) * sizeof(double);
    PMB_DEV static void run(const Warp& w, int b, unsigned char* smem, O o, OcpIo io)
    {
        using E = OcpEval<O>;
        Cta c(w, reinterpret_cast<double*>(smem));
        double* nv = reinterpret_cast<double*>(smem) + Cta::SCRATCH_DOUBLES;
        const double* var = io.var + (size_t)b * O::N;
        const double* d = io.d ? io.d + (size_t)b * O::ND : nullptr;
        const double* lam = io.lam ? io.lam + (size_t)b * O::DUAL : nullptr;
        double cost = 0.0;
        if (MODE == OCP_COST) cost = E::cost(c, o, var, d);
        else if (MODE == OCP_EQ) E::equalities(c, o, var, d, io.c + (size_t)b * O::NUM_EQ, c.tid());
        else if (MODE == OCP_INEQ) E::inequalities(c, o, var, d, io.g + (size_t)b * O::NUM_INEQ, c.tid());
        else if (MODE == OCP_EQ_LIN)
            E::constraints_linearised(c, o, var, d, io.c + (size_t)b * O::NUM_EQ, io.jac + (size_t)b * O::NUM_EQ * O::N, O::NUM_EQ, false);
        else if (MODE == OCP_COST_GRAD) cost = E::cost_gradient(c, o, var, d, io.grad + (size_t)b * O::N, nv);
        else if (MODE == OCP_COST_GRAD_HESS)
            cost = E::cost_gradient_hessian(c, o, var, d, nullptr, io.grad + (size_t)b * O::N, io.hess + (size_t)b * O::N * O::N, nv);
        else if (MODE == OCP_LAG_GRAD)
            cost = E::lagrangian_gradient(c, o, var, d, lam, io.lag_grad + (size_t)b * O::N, io.grad + (size_t)b * O::N,
                                          io.c + (size_t)b * O::M, io.jac + (size_t)b * O::M * O::N, nv);
        else if (MODE == OCP_LAG_GRAD_HESS)
            cost = E::lagrangian_gradient_hessian(c, o, var, d, lam, io.lag_grad + (size_t)b * O::N, io.hess + (size_t)b * O::N * O::N,
                                                  io.grad + (size_t)b * O::N, io.c + (size_t)b * O::M, io.jac + (size_t)b * O::M * O::N, nv);
        if (io.cost && c.tid() == 0 && MODE != OCP_EQ && MODE != OCP_INEQ && MODE != OCP_EQ_LIN) io.cost[b] = cost;
    }
};

// ---- QP -------------------------------------------------------------------------------------------------------------
struct QpBatch {
    int N, M;
    const double *H, *h, *A, *Alb, *Aub, *xlb, *xub, *xg, *yg;
    double *x, *y;
    pmb_qp_info_t* info;
    double *z, *q;
    int *perm, *ctype, *nfac;
};

PMB_DEV QpArgs qp_instance(const QpBatch& q, int b)
{
    const size_t N = q.N, M = q.M, n = N + M;
    QpArgs a;
    a.N = q.N; a.M = q.M;
    a.H = q.H + b * N * N; a.h = q.h + b * N; a.A = q.A + b * M * N; a.Alb = q.Alb + b * M; a.Aub = q.Aub + b * M;
    a.xlb = q.xlb + b * N; a.xub = q.xub + b * N;
    a.xg = q.xg ? q.xg + b * N : nullptr; a.yg = q.yg ? q.yg + b * n : nullptr;
    a.x = q.x + b * N; a.y = q.y + b * n; a.info = q.info ? q.info + b : nullptr;
    a.z = q.z ? q.z + b * M : nullptr; a.q = q.q ? q.q + b * N : nullptr;
    a.perm = q.perm ? q.perm + b * n : nullptr; a.ctype = q.ctype ? q.ctype + b * n : nullptr; a.nfac = q.nfac ? q.nfac + b : nullptr;
    a.prof = nullptr;
    return a;
}

/** where the packed factor of a CTA lives: in shared memory when it fits, else in a per-CTA global scratch slot */
struct FactorStore {
    double* global;      // grid * factor_doubles (nullptr: shared memory)
    size_t doubles;      // n (n + 1) / 2
    int sm_count;        // SMs of the device (0: unknown) — used to rotate the serial warp between co-resident CTAs
};

/** persistent CTA-per-instance boxADMM: CTAs draw instances from an atomic queue */
template <int R, bool IN_SMEM, bool FAST = false>
struct QpBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = R <= 4 ? (FAST ? 3 : 4) : (R <= 6 ? 2 : 1);
    static constexpr const char* NAME = FAST ? "qp_box_admm_fast" : "qp_box_admm";
    static constexpr size_t EMU_STACK_BYTES = 1u << 20;
    PMB_DEV static void run(const Warp& w, int blk, unsigned char* smem, pmb_qp_settings_t st, QpBatch qb, FactorStore fs, int batch, int* queue)
    {
        Cta c(w, reinterpret_cast<double*>(smem), fs.sm_count > 0 ? blk + blk / fs.sm_count : 0);
        unsigned char* ws = smem + Cta::SCRATCH_DOUBLES * sizeof(double);
        // IN_SMEM is a template parameter so that the factor pointer is provably a shared-memory address (LDS/STS)
        double* Lp = IN_SMEM ? reinterpret_cast<double*>(ws) : fs.global + (size_t)blk * fs.doubles;
        unsigned char* vec = IN_SMEM ? ws + fs.doubles * sizeof(double) : ws;
        for (;;) {
            const int b = c.bcast_int(c.tid() == 0 ? atomic_add(queue, 1) : 0);
            if (b >= batch) break;
            const QpArgs a = qp_instance(qb, b);
            qp_solve_cta<R, 0, 0, 4, FAST>(c, st, a, Lp, vec);
        }
    }
};

// ---- the reference's OSQP-style ADMM<> as a stand-alone batched operator (C ABI pmb_qp_solve_admm) ----------------------------
struct AdmmBatch {
    int N, M;
    const double *H, *h, *Ae, *Alb, *Aub, *xlb, *xub, *xg, *yg;
    double *x, *y;
    pmb_qp_info_t* info;
    double* z;
    int *perm, *ctype, *nfac;
};
template <int R, bool IN_SMEM>
struct QpAdmmBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = R <= 4 ? 4 : (R <= 6 ? 2 : 1);
    static constexpr const char* NAME = "qp_osqp_admm";
    static constexpr size_t EMU_STACK_BYTES = 1u << 20;
    PMB_DEV static void run(const Warp& w, int blk, unsigned char* smem, pmb_qp_settings_t st, AdmmBatch qb, FactorStore fs, int batch, int* queue)
    {
        Cta c(w, reinterpret_cast<double*>(smem), fs.sm_count > 0 ? blk + blk / fs.sm_count : 0);
        unsigned char* ws = smem + Cta::SCRATCH_DOUBLES * sizeof(double);
        double* Lp = IN_SMEM ? reinterpret_cast<double*>(ws) : fs.global + (size_t)blk * fs.doubles;
        unsigned char* vec = IN_SMEM ? ws + fs.doubles * sizeof(double) : ws;
        for (;;) {
            const int b = c.bcast_int(c.tid() == 0 ? atomic_add(queue, 1) : 0);
            if (b >= batch) break;
            const size_t N = qb.N, M = qb.M, Me = N + M, n = N + Me, sb = b;
            AdmmArgs a;
            a.N = qb.N; a.M = qb.M;
            a.H = qb.H + sb * N * N; a.h = qb.h + sb * N; a.Ae = qb.Ae + sb * Me * N; a.Alb = qb.Alb + sb * M; a.Aub = qb.Aub + sb * M;
            a.xlb = qb.xlb + sb * N; a.xub = qb.xub + sb * N;
            a.xg = qb.xg ? qb.xg + sb * N : nullptr; a.yg = qb.yg ? qb.yg + sb * Me : nullptr;
            a.x = qb.x + sb * N; a.y = qb.y + sb * Me; a.info = qb.info ? qb.info + sb : nullptr;
            a.z = qb.z ? qb.z + sb * Me : nullptr; a.perm = qb.perm ? qb.perm + sb * n : nullptr;
            a.ctype = qb.ctype ? qb.ctype + sb * Me : nullptr; a.nfac = qb.nfac ? qb.nfac + sb : nullptr;
            admm_solve_cta<R, 4>(c, st, a, Lp, vec);
        }
    }
};

// ---- a17: KKT assembly, reference layout (dense (N+M)^2, lower part + diagonal blocks written, rest zero) ------------
/** The metric's "KKT kernel": K = [[H + sigma I + diag(rho_box), 0], [A, -diag(1 / rho_A)]] materialised in HBM, one CTA per
 *  instance.  Pure data movement (read H and A once, write K once): columns are dealt to the warps two at a time, a lane owns
 *  rows lane, lane + 32, ... of both columns and issues all its loads before the first store (eight independent 8-byte loads in
 *  flight per lane for n <= 128), every warp-level access is a contiguous run of a column, loads are read-only streaming
 *  (ld.global.cs) and stores are streaming (st.global.cs: K is not re-read here), no integer division anywhere. */
struct KktDenseBody {
    static constexpr int THREADS = 256;
#ifdef PMB_KKT_MINB
    static constexpr int MIN_BLOCKS = PMB_KKT_MINB;
#else
    static constexpr int MIN_BLOCKS = 6;     // 42 registers per thread: measured 5.11 / 5.32 / 5.78 TB/s at 4 / 5 / 6 CTAs per SM (8: spills, 4.35)
#endif
    static constexpr const char* NAME = "kkt_assemble_dense";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static double ld(const double* p)
    {
#if defined(__CUDACC__) && !defined(PMB_EMU)
        return __ldcs(p);
#else
        return *p;
#endif
    }
    PMB_DEV static void st(double* p, double v)
    {
#if defined(__CUDACC__) && !defined(PMB_EMU)
        __stcs(p, v);
#else
        *p = v;
#endif
    }
    PMB_DEV static void run(const Warp& w, int b, unsigned char*, int N, int M, const double* H, const double* A, const double* rho_box,
                            const double* rho_inv, double sigma, double* K)
    {
        const int n = N + M, lane = w.lane(), wid = w.warp_id(), nw = w.nthreads() >> 5;
        const double* Hb = H + (size_t)b * N * N;
        const double* Ab = A + (size_t)b * M * N;
        const double* rb = rho_box + (size_t)b * N;
        const double* ri = rho_inv + (size_t)b * M;
        double* Kb = K + (size_t)b * n * n;
        auto entry = [&](int i, int j) -> double {
            if (j < N) {                                            // warp-uniform
                if (i < N) { double v = ld(Hb + i + (size_t)j * N); if (i == j) { v += sigma; v += rb[i]; } return v; }
                return ld(Ab + (i - N) + (size_t)j * M);
            }
            return i == j ? -ri[i - N] : 0.0;
        };
#if defined(__CUDACC__) && !defined(PMB_EMU)
        if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(Kb) & 15) == 0 && n <= 128) {
            // a lane owns the row pairs (2 lane, 2 lane + 1) and (64 + 2 lane, 65 + 2 lane) of two columns: eight loads in flight,
            // then four 16-byte streaming stores (columns of K start 16-byte aligned when n is even)
            for (int j = 2 * wid; j < n; j += 2 * nw) {
                const int j2 = j + 1 < n ? j + 1 : j;
                const int i0 = 2 * lane, i1 = 64 + 2 * lane;
                const bool hi = i1 < n;
                const double a0 = entry(i0, j), a1 = entry(i0 + 1, j), b0 = entry(i0, j2), b1 = entry(i0 + 1, j2);
                const double c0 = hi ? entry(i1, j) : 0.0, c1 = hi ? entry(i1 + 1, j) : 0.0, d0 = hi ? entry(i1, j2) : 0.0, d1 = hi ? entry(i1 + 1, j2) : 0.0;
                if (i0 < n) {
                    __stcs(reinterpret_cast<double2*>(Kb + i0 + (size_t)j * n), make_double2(a0, a1));
                    if (j2 != j) __stcs(reinterpret_cast<double2*>(Kb + i0 + (size_t)j2 * n), make_double2(b0, b1));
                }
                if (hi) {
                    __stcs(reinterpret_cast<double2*>(Kb + i1 + (size_t)j * n), make_double2(c0, c1));
                    if (j2 != j) __stcs(reinterpret_cast<double2*>(Kb + i1 + (size_t)j2 * n), make_double2(d0, d1));
                }
            }
            return;
        }
#endif
        constexpr int U = 4;                                        // rows per lane and column held in registers
        for (int j = 2 * wid; j < n; j += 2 * nw) {
            const int j2 = j + 1 < n ? j + 1 : j;
            for (int i0 = 0; i0 < n; i0 += 32 * U) {
                double v[U], u[U];
                PMB_UNROLL
                for (int k = 0; k < U; ++k) { const int i = i0 + lane + 32 * k; v[k] = i < n ? entry(i, j) : 0.0; u[k] = i < n ? entry(i, j2) : 0.0; }
                PMB_UNROLL
                for (int k = 0; k < U; ++k) {
                    const int i = i0 + lane + 32 * k;
                    if (i < n) { st(Kb + i + (size_t)j * n, v[k]); if (j2 != j) st(Kb + i + (size_t)j2 * n, u[k]); }
                }
            }
        }
    }
};

// ---- BFGS operator (C ABI pmb_bfgs_update) -----------------------------------------------------------------------------
struct BfgsBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static size_t smem_bytes(int N) { return (Cta::SCRATCH_DOUBLES + 2 * (size_t)N) * sizeof(double); }
    static constexpr const char* NAME = "bfgs_update";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static void run(const Warp& w, int b, unsigned char* smem, int N, double* B, const double* s, const double* y, int* branch)
    {
        Cta c(w, reinterpret_cast<double*>(smem));
        double* Bs = reinterpret_cast<double*>(smem) + Cta::SCRATCH_DOUBLES;
        double* r = Bs + N;
        const int br = bfgs_update_cta<false>(c, N, B + (size_t)b * N * N, s + (size_t)b * N, y + (size_t)b * N, Bs, r);
        if (branch && c.tid() == 0) branch[b] = br;
    }
};

// ---- RuizEquilibration operators (C ABI pmb_ruiz_equilibrate / pmb_ruiz_unscale), sizes at run time ----------------------------
struct RuizArgs { double *H, *h, *A, *Al, *Au, *l, *u, *st, *x, *y; };
struct RuizComputeBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static size_t smem_bytes(int N, int M) { return (Cta::SCRATCH_DOUBLES + (size_t)N + M) * sizeof(double); }
    static constexpr const char* NAME = "ruiz_equilibrate";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static void run(const Warp& w, int b, unsigned char* smem, int N, int M, int variant, RuizArgs a)
    {
        Cta c(w, reinterpret_cast<double*>(smem));
        const size_t sb = b;
        RuizCta<0, 0>::compute(c, N, M, variant, a.H + sb * N * N, a.h + sb * N, a.A + sb * M * N, a.Al + sb * M, a.Au + sb * M, a.l + sb * N,
                               a.u + sb * N, a.st + sb * (N + M + 1), reinterpret_cast<double*>(smem) + Cta::SCRATCH_DOUBLES);
    }
};
struct RuizUnscaleBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = Cta::SCRATCH_DOUBLES * sizeof(double);
    static constexpr const char* NAME = "ruiz_unscale";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static void run(const Warp& w, int b, unsigned char* smem, int N, int M, RuizArgs a)
    {
        Cta c(w, reinterpret_cast<double*>(smem));
        const size_t sb = b;
        RuizCta<0, 0>::unscale(c, N, M, a.st + sb * (N + M + 1), a.H + sb * N * N, a.h + sb * N, a.A + sb * M * N, a.Al + sb * M, a.Au + sb * M,
                               a.l + sb * N, a.u + sb * N, a.x ? a.x + sb * N : nullptr, a.y ? a.y + sb * (N + M) : nullptr);
    }
};

// ---- block BFGS operator (C ABI pmb_ocp_block_bfgs_update): ContinuousOCP<..., SPARSE>::hessian_update_impl ---------------
template <class O>
struct BlockBfgsBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = (Cta::SCRATCH_DOUBLES + 2 * (size_t)O::N) * sizeof(double);
    static constexpr const char* NAME = "block_bfgs_update";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static void run(const Warp& w, int b, unsigned char* smem, double* B, const double* s, const double* y, int* branch)
    {
        Cta c(w, reinterpret_cast<double*>(smem));
        double* v = reinterpret_cast<double*>(smem) + Cta::SCRATCH_DOUBLES;
        double* r = v + O::N;
        const int br = block_bfgs_update_cta<O>(c, B + (size_t)b * O::N * O::N, s + (size_t)b * O::N, y + (size_t)b * O::N, v, r);
        if (branch && c.tid() == 0) branch[b] = br;
    }
};

// ---- work-queue order: longest processing time first, from the iteration counts of the previous solve -------------------
/** The persistent sqp_solve kernel is a list scheduler: CTAs take instances off a queue.  With a few instances that need 10x
 *  the iterations of the rest (0.4 % of the robot sweep run all 100), the makespan of a batch is bulk + the longest chain that
 *  happened to start late.  Re-solving the same fleet (closed-loop MPC; every timed step of bench.py) the previous solve's
 *  iteration counts are an excellent predictor, so the queue is served in descending order of them (LPT rule): the long chains
 *  start at t = 0 and overlap with the bulk.  Counting sort by iteration count, one block.  Results do not depend on the order. */
struct LptOrderBody {
    static constexpr int THREADS = 1024;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int MAX_KEY = 1023;
    static constexpr size_t SMEM = (MAX_KEY + 2) * sizeof(int);
    static constexpr const char* NAME = "lpt_order";
    static constexpr size_t EMU_STACK_BYTES = 64u << 10;
    PMB_DEV static void run(const Warp& w, int, unsigned char* smem, int batch, const pmb_sqp_info_t* info, int* order)
    {
        int* bin = reinterpret_cast<int*>(smem);                   // bin[k]: instances with key k, later the write cursor
        const int tid = w.tid(), nt = w.nthreads();
        for (int k = tid; k <= MAX_KEY + 1; k += nt) bin[k] = 0;
        w.block_sync();
        for (int b = tid; b < batch; b += nt) { int k = info[b].iter; k = k < 0 ? 0 : (k > MAX_KEY ? MAX_KEY : k); atomic_add(bin + k, 1); }
        w.block_sync();
        if (tid == 0) { int run = 0; for (int k = MAX_KEY; k >= 0; --k) { const int cnt = bin[k]; bin[k] = run; run += cnt; } }   // descending keys first
        w.block_sync();
        for (int b = tid; b < batch; b += nt) { int k = info[b].iter; k = k < 0 ? 0 : (k > MAX_KEY ? MAX_KEY : k); order[atomic_add(bin + k, 1)] = b; }
    }
};

// ---- SQP pipeline ---------------------------------------------------------------------------------------------------
/** the whole SQPBase::solve of the batch in ONE persistent launch: each CTA draws an instance from the atomic queue and
 *  iterates linearise -> boxADMM -> line search / step on it until it converges (no host round trip per iteration, no
 *  wave quantisation: a slow instance only occupies its own CTA). */
/** QPK: the QP solver of the loop — 0 boxADMM, 1 the OSQP-style ADMM<> (exact arithmetic; KKT dimension 2N + M) */
template <class O, bool FAST = false, int QPK = 0>
struct SqpSolveBody {
    static_assert(!(FAST && QPK != 0), "the OSQP-style ADMM runs in exact arithmetic only");
    static constexpr const char* NAME = QPK ? "sqp_solve_osqp_admm" : (FAST ? "sqp_solve_fast" : "sqp_solve");
    static constexpr size_t EMU_STACK_BYTES = 4u << 20;
    static constexpr int KN = QPK ? 2 * O::N + O::M : O::N + O::M;       // dimension of the KKT system
    static constexpr int R = (KN + 31) / 32;
    /** doubles of the LDL^T workspace: packed lower triangle (exact arithmetic) or the tile workspace of pmb_qp_fast.hpp */
    static constexpr size_t FACTOR_DOUBLES = FAST ? fast::workspace_doubles(O::N + O::M) : (size_t)KN * (KN + 1) / 2;
    static constexpr size_t SCRATCH_BYTES = SqpDev<O>::SCRATCH_DOUBLES * sizeof(double);
    static constexpr size_t VEC_BYTES = QPK ? admm_vec_bytes(O::N, O::M) : qp_vec_bytes(O::N, O::M);
    static constexpr size_t SMEM_IN = QPK ? Cta::SCRATCH_DOUBLES * sizeof(double) + FACTOR_DOUBLES * sizeof(double) + VEC_BYTES + 1024
                                          : Cta::SCRATCH_DOUBLES * sizeof(double) + FACTOR_DOUBLES * sizeof(double) + 3 * (O::N + O::M) * 8 +
                                            (6 * O::N + 4 * O::M) * 8 + 2 * (O::N + O::M) * 4 + 16 + 1024;
    static constexpr bool IN_SMEM = SMEM_IN <= 227 * 1024;    // placement of the factor, fixed per problem at compile time
    /** Threads per CTA and resident CTAs per SM the register allocation is asked to allow.
     *  Factor in shared memory: 128 threads, at most 3 CTAs per SM (168 registers per thread; measured on the mobile robot:
     *  4 CTAs x 128 registers spill inside the QP loops and are slower end to end — 127 ms vs 109 ms per batch of 8192 — although
     *  shared memory would admit 4).
     *  Factor in a global (L2) slot — large problems with heavy AD code (kite 12 x 1): ONE CTA of 256 threads per SM.  Per-instance
     *  latency is what counts there (one SQP iteration is 10-25 M cycles) and the 255-register budget halves the spills:
     *  measured on kite 12 x 1, batch 1024: 128 threads x 2 CTAs 4.1 k / 8.4 k it/s (exact / fast), 256 x 2 4.8 k / 10.9 k,
     *  256 x 1 6.5 k / 16.1 k.  The same holds for a factor in shared memory that is so large that only one CTA fits per SM
     *  (the OSQP-style ADMM variant of the robot: 115 KB; 253 k -> 294 k it/s with 256 threads). */
#ifdef PMB_SQP_THREADS
    static constexpr int THREADS = PMB_SQP_THREADS;
#else
    static constexpr bool ALONE_ON_SM = !IN_SMEM || (228 * 1024) / SMEM_IN < 2;   // one CTA per SM either way: give it 8 warps
    static constexpr int THREADS = ALONE_ON_SM ? 256 : 128;
#endif
#ifdef PMB_MINB
    static constexpr int MIN_BLOCKS = PMB_MINB;
#else
    static constexpr int MIN_BLOCKS = THREADS == 256 ? 1 : ((228 * 1024) / SMEM_IN > 3 ? 3 : (int)((228 * 1024) / SMEM_IN));
#endif
    /** exact arithmetic with the factor in a global slot: the 32 x 32 diagonal blocks of the factor are staged in shared memory
     *  for the substitutions (pmb_qp.hpp::ldlt_stage_diag_blocks) */
    static constexpr size_t DIAG_WANT = staged_solve_doubles(O::N + O::M, THREADS / 32) * sizeof(double) + 16;
    static constexpr size_t DIAG_BYTES = (!IN_SMEM && !FAST && QPK == 0 && Cta::SCRATCH_DOUBLES * sizeof(double) + SCRATCH_BYTES + qp_vec_bytes(O::N, O::M) +
                                          DIAG_WANT <= 227 * 1024) ? DIAG_WANT : 0;
    /** shared memory: Cta scratch | factor (aliased by the SQP scratch) | QP vectors                    (factor in shared memory)
     *                 Cta scratch | SQP scratch | QP vectors | diagonal blocks of the factor (exact)    (factor in global scratch) */
    static size_t smem_bytes()
    {
        const size_t fac = FACTOR_DOUBLES * sizeof(double);
        const size_t first = IN_SMEM ? (fac > SCRATCH_BYTES ? fac : SCRATCH_BYTES) : SCRATCH_BYTES;
        return Cta::SCRATCH_DOUBLES * sizeof(double) + first + VEC_BYTES + DIAG_BYTES;
    }
    PMB_DEV static void run(const Warp& w, int blk, unsigned char* smem, O o, SqpWs ws, pmb_sqp_settings_t st, pmb_qp_settings_t qst,
                            FactorStore fs, int batch, int* queue)
    {
        // CTAs blk, blk + #SMs, blk + 2 #SMs, ... share an SM (breadth-first placement of a one-wave grid): give them
        // different serial warps (fs.sm_count is the number of SMs of the device)
        Cta c(w, reinterpret_cast<double*>(smem), fs.sm_count > 0 ? blk + blk / fs.sm_count : 0);
        unsigned char* base = smem + Cta::SCRATCH_DOUBLES * sizeof(double);
        double* scratch = reinterpret_cast<double*>(base);
        double* Lp;
        unsigned char* vec;
        if (!IN_SMEM) { Lp = fs.global + (size_t)blk * fs.doubles; vec = base + SCRATCH_BYTES; }
        else {
            Lp = reinterpret_cast<double*>(base);
            const size_t fac = FACTOR_DOUBLES * sizeof(double);
            vec = base + (fac > SCRATCH_BYTES ? fac : SCRATCH_BYTES);
        }
        double* Ld = DIAG_BYTES ? reinterpret_cast<double*>(vec + ((VEC_BYTES + 15) & ~(size_t)15)) : nullptr;
        for (;;) {
            const int ticket = c.bcast_int(c.tid() == 0 ? atomic_add(queue, 1) : 0);
            if (ticket >= batch) break;
            const int b = ws.order ? ws.order[ticket] : ticket;
            const SqpInst<O> s{ws, b};
            SqpDev<O>::template solve<R, THREADS / 32, FAST, !IN_SMEM, QPK>(c, o, s, st, qst, Lp, vec, scratch, Ld);
        }
    }
};

} // namespace pmb
