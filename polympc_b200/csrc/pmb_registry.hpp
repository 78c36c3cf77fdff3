// pmb_registry.hpp — type-erased problem interface between the C ABI (pmb_capi.cu) and the per-problem translation
// units (problems/*.cu).  Each problem class x collocation grid is compiled by nvcc in its own TU (the user functors are
// templates, so every kernel that touches them is stamped out per problem; separate TUs keep ptxas time bounded and let
// the build run in parallel).  A TU exports one factory through PMB_DEFINE_PROBLEM.
#pragma once
#include "../../include/polympc_b200.h"
#include "pmb_rt.hpp"
#include "pmb_problems.hpp"
#include "pmb_kernels.hpp"

namespace pmb {

struct IProblem {
    pmb_dims_t dims{};
    int device = 0;
    virtual ~IProblem() {}
    virtual void set_params(const double*) = 0;
    virtual void get_params(double*) const = 0;
    virtual void set_time_limits(double, double) = 0;
    virtual void time_nodes(double*) const = 0;
    virtual bool launch_eval(int mode, int batch, const OcpIo& io, stream_t s) const = 0;
    virtual bool launch_block_bfgs(int batch, double* B, const double* sv, const double* y, int* branch, stream_t s) const = 0;
    /** placement of the LDL^T factor (shared memory, or a per-CTA global scratch slot), shared-memory bytes and resident CTAs
     *  of the fused SQP kernel */
    virtual bool factor_in_smem() const = 0;
    virtual size_t solve_smem_bytes() const = 0;
    virtual size_t factor_doubles() const = 0;
    virtual int solve_resident_ctas() const = 0;
    /** one persistent launch that solves `batch` instances (grid CTAs draw them from `queue`) */
    virtual bool launch_solve(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst,
                              double* factor_scratch, int batch, int* queue, stream_t s) const = 0;
    /** the same kernel in fast arithmetic (pmb_qp_fast.hpp); the tile workspace lives in shared memory when it fits, else in a
     *  per-CTA global slot of fast_factor_doubles() that stays L2 resident */
    virtual bool fast_in_smem() const = 0;
    virtual size_t fast_smem_bytes() const = 0;
    virtual size_t fast_factor_doubles() const = 0;
    virtual int fast_resident_ctas() const = 0;
    virtual bool launch_solve_fast(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst,
                                   double* factor_scratch, int batch, int* queue, stream_t s) const = 0;
    /** the same loop with the OSQP-style ADMM<> as its QP solver (exact arithmetic; instantiated for 2N + M <= 256) */
    virtual bool admm_supported() const = 0;
    virtual bool admm_in_smem() const = 0;
    virtual size_t admm_smem_bytes() const = 0;
    virtual size_t admm_factor_doubles() const = 0;
    virtual int admm_resident_ctas() const = 0;
    virtual bool launch_solve_admm(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst,
                                   double* factor_scratch, int batch, int* queue, stream_t s) const = 0;
};

template <class O>
struct ProblemImpl : IProblem {
    O o;
    ProblemImpl()
    {
        o.init();
        dims.NX = O::NX; dims.NU = O::NU; dims.NP = O::NP; dims.ND = O::ND; dims.NG = O::NG; dims.P = O::P; dims.S = O::S; dims.NN = O::NN;
        dims.N = O::N; dims.M = O::M; dims.DUAL = O::DUAL; dims.NPARAM = O::Model::NPARAM;
    }
    void set_params(const double* v) override { o.model.set_params(v); }
    void get_params(double* v) const override { o.model.get_params(v); }
    void set_time_limits(double a, double b) override { o.set_time_limits(a, b); }
    void time_nodes(double* t) const override { for (int i = 0; i < O::NN; ++i) t[i] = o.time_nodes[i]; }
    template <int MODE> bool ev(int batch, const OcpIo& io, stream_t s) const { return rt_launch<OcpEvalBody<O, MODE>>(batch, OcpEvalBody<O, MODE>::SMEM, s, o, io); }
    bool launch_eval(int mode, int batch, const OcpIo& io, stream_t s) const override
    {
        switch (mode) {
        case OCP_COST: return ev<OCP_COST>(batch, io, s);
        case OCP_EQ: return ev<OCP_EQ>(batch, io, s);
        case OCP_INEQ: return ev<OCP_INEQ>(batch, io, s);
        case OCP_EQ_LIN: return ev<OCP_EQ_LIN>(batch, io, s);
        case OCP_COST_GRAD: return ev<OCP_COST_GRAD>(batch, io, s);
        case OCP_COST_GRAD_HESS: return ev<OCP_COST_GRAD_HESS>(batch, io, s);
        case OCP_LAG_GRAD: return ev<OCP_LAG_GRAD>(batch, io, s);
        case OCP_LAG_GRAD_HESS: return ev<OCP_LAG_GRAD_HESS>(batch, io, s);
        }
        return false;
    }
    bool launch_block_bfgs(int batch, double* B, const double* sv, const double* y, int* branch, stream_t s) const override
    { return rt_launch<BlockBfgsBody<O>>(batch, BlockBfgsBody<O>::SMEM, s, B, sv, y, branch); }
    using Solve = SqpSolveBody<O>;
    bool factor_in_smem() const override { return Solve::IN_SMEM; }
    size_t solve_smem_bytes() const override { return Solve::smem_bytes(); }
    size_t factor_doubles() const override { return Solve::FACTOR_DOUBLES; }
    int solve_resident_ctas() const override
    {
        return resident_ctas<Solve, O, SqpWs, pmb_sqp_settings_t, pmb_qp_settings_t, FactorStore, int, int*>(
            Solve::smem_bytes(), o, SqpWs{}, pmb_sqp_settings_t{}, pmb_qp_settings_t{}, FactorStore{}, 0, (int*)nullptr);
    }
    bool launch_solve(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst, double* factor_scratch,
                      int batch, int* queue, stream_t s) const override
    {
        FactorStore fs{Solve::IN_SMEM ? nullptr : factor_scratch, Solve::FACTOR_DOUBLES, rt_sm_count()};
        return rt_launch<Solve>(grid, Solve::smem_bytes(), s, o, ws, st, qst, fs, batch, queue);
    }
    using SolveFast = SqpSolveBody<O, true>;
    bool fast_in_smem() const override { return SolveFast::IN_SMEM; }
    size_t fast_smem_bytes() const override { return SolveFast::smem_bytes(); }
    size_t fast_factor_doubles() const override { return SolveFast::FACTOR_DOUBLES; }
    int fast_resident_ctas() const override
    {
        return resident_ctas<SolveFast, O, SqpWs, pmb_sqp_settings_t, pmb_qp_settings_t, FactorStore, int, int*>(
            SolveFast::smem_bytes(), o, SqpWs{}, pmb_sqp_settings_t{}, pmb_qp_settings_t{}, FactorStore{}, 0, (int*)nullptr);
    }
    bool launch_solve_fast(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst, double* factor_scratch,
                           int batch, int* queue, stream_t s) const override
    {
        FactorStore fs{SolveFast::IN_SMEM ? nullptr : factor_scratch, SolveFast::FACTOR_DOUBLES, rt_sm_count()};
        return rt_launch<SolveFast>(grid, SolveFast::smem_bytes(), s, o, ws, st, qst, fs, batch, queue);
    }
    static constexpr bool ADMM_OK = 2 * O::N + O::M <= 256;
    using SolveAdmm = SqpSolveBody<O, false, ADMM_OK ? 1 : 0>;     // (too large: an alias of the boxADMM kernel that is never launched)
    bool admm_supported() const override { return ADMM_OK; }
    bool admm_in_smem() const override { return SolveAdmm::IN_SMEM; }
    size_t admm_smem_bytes() const override { return SolveAdmm::smem_bytes(); }
    size_t admm_factor_doubles() const override { return SolveAdmm::FACTOR_DOUBLES; }
    int admm_resident_ctas() const override
    {
        return resident_ctas<SolveAdmm, O, SqpWs, pmb_sqp_settings_t, pmb_qp_settings_t, FactorStore, int, int*>(
            SolveAdmm::smem_bytes(), o, SqpWs{}, pmb_sqp_settings_t{}, pmb_qp_settings_t{}, FactorStore{}, 0, (int*)nullptr);
    }
    bool launch_solve_admm(int grid, const SqpWs& ws, const pmb_sqp_settings_t& st, const pmb_qp_settings_t& qst, double* factor_scratch,
                           int batch, int* queue, stream_t s) const override
    {
        if (!ADMM_OK) return false;
        FactorStore fs{SolveAdmm::IN_SMEM ? nullptr : factor_scratch, SolveAdmm::FACTOR_DOUBLES, rt_sm_count()};
        return rt_launch<SolveAdmm>(grid, SolveAdmm::smem_bytes(), s, o, ws, st, qst, fs, batch, queue);
    }
};


} // namespace pmb

/** defines the factory `pmb::IProblem* pmb_make_<ID>()` for Ocp<MODEL, P, S> */
#define PMB_DEFINE_PROBLEM(ID, MODEL, P, S) \
    namespace pmb { IProblem* pmb_make_##ID() { return new ProblemImpl<Ocp<MODEL, P, S>>(); } }
#define PMB_DECLARE_PROBLEM(ID) namespace pmb { IProblem* pmb_make_##ID(); }
