// polympc_compat.hpp — source-compatibility layer: PolyMPC problem classes and host code, written against the reference's
// header-only API, compile unchanged on top of the B200 engine.
//
//   reference header / concept (file:line)                                    what this header provides
//   ------------------------------------------------------------------------------------------------------------------
//   polympc::Chebyshev<P, GAUSS_LOBATTO, double>  (src/polynomials/ebyshev.hpp:27-95)   tag carrying POLY_ORDER
//   polympc::Spline<Polynomial, S>                (src/polynomials/splines.hpp:22-46)   tag carrying NUM_SEGMENTS / NUM_NODES
//   POLYMPC_FORWARD_DECLARATION, polympc_traits   (src/control/continuous_ocp.hpp:22-37) same macro, same traits
//   ContinuousOCP<OCP, Approximation, FMT>        (continuous_ocp.hpp:39-182, 191-288)   sizes, state_t<T> ..., default *_impl
//                                                                                        functors, set_time_limits; NO
//                                                                                        transcription — that is the kernels
//   SQPBase<Derived, Problem, QPSolver>           (src/solvers/sqp_base.hpp:64-197, 568) host object, one instance, same
//                                                                                        getters / setters, solve() runs the
//                                                                                        fused sm_100a kernel
//   boxADMM<...>, ADMM<...>, qp_solver_settings_t (src/solvers/box_admm.hpp:15, admm.hpp:15, qp_base.hpp:17-175)  QPBase objects
//                                                                                        (solve 7 / 9 arguments on the device) and the
//                                                                                        QPSolver argument of SQPBase
//   MPC<OCP, Solver, Args...>                     (src/control/mpc_wrapper.hpp:17-298)   same methods
//
// The GPU side: pmb::compat::Model<OCP> adapts the Eigen-style functors of a problem class to the engine's pointer-style
// functor concept (pmb_problems.hpp); plain-double evaluations are routed through Dual<double,0> so that user code calling
// an unqualified cos()/sin()/exp() always lands in the deterministic pmb::dm implementations (argument-dependent lookup),
// never in libm / the CUDA math library.  A problem class must be trivially copyable (it is passed to the kernel by value)
// and its functors must be callable on the device: either mark them POLYMPC_HD, or rely on the default below which makes
// `inline` mean `__host__ __device__ inline` for the code that FOLLOWS this header in an nvcc translation unit (the
// reference's functors are all declared `inline` / EIGEN_STRONG_INLINE).  Define POLYMPC_B200_NO_INLINE_HD to opt out.
//
// CRTP hooks of SQPBase (sqp_base.hpp:198-350).  A fused device loop cannot call host code, so solve() finds out what a
// Derived solver's overrides DO by running each of them once on the host against a recording problem object ("hook probe"),
// and maps the result onto the engine's menu:
//   hessian_update_impl              default (dense damped BFGS)  |  forwarded to problem.hessian_update_impl — the OCP's block
//                                    BFGS when MATRIXFMT == SPARSE (continuous_ocp.hpp:2303-2431), dense BFGS when DENSE
//   update_linearisation_*_impl      default (gradient + quasi-Newton update)  |  forwarded to linearisation_*_impl (exact
//                                    Hessian at every iteration, minimal_time_test.cpp:124-143)
//   hessian_regularisation_*_impl    default (none)  |  the Gershgorin shift of minimal_time_test.cpp:90-122
//   step_size_selection_impl         default (l1 merit)  |  the LSFilter line search of valet_parking_mpc_test.cpp:110-155 (recognised by
//                                    running it against scripted cost / violation values; the solver's `filter` member is kept in sync)
//   Preconditioner                   IdentityPreconditioner  |  RuizEquilibration<..., DENSE | SPARSE>
//   QPSolver                         boxADMM<...>  |  ADMM<...> (the OSQP-style splitting, exact arithmetic, 2N + M <= 256)
// Anything else — an override that does something the menu does not have, another step_size_selection_impl, an overridden
// constraints_violation_impl / max_constraints_violation_impl / termination_criteria_impl, a non-null iteration_callback,
// a QP solver type other than boxADMM<> / ADMM<> — is REFUSED: solve() prints what it found to stderr and
// returns with status INVALID_SETTINGS instead of silently running a different algorithm.
// MATRIXFMT == SPARSE selects the SPARSE *semantics* (block-diagonal quasi-Newton update); storage on the device is dense.
#pragma once
#include "pmb_eigen_shim.hpp"
#include "../polympc_b200.h"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <list>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#if defined(__CUDACC__)
#define POLYMPC_HD __host__ __device__
#else
#define POLYMPC_HD
#endif

namespace pmb { namespace compat {
template <class U> const char* problem_name();
/** recording of what a Derived solver's hooks call while solve() probes them on the host (thread local) */
struct HookProbe {
    enum Event { LAG_GRAD = 1, LAG_GRAD_HESS = 2, HESS_UPDATE_DEFAULT = 3, HESS_UPDATE_OCP = 4, COST = 5, VIOLATION = 6 };
    int events[16]; int n = 0;
    double cost_script[8], viol_script[8]; int n_script = 0, cost_cursor = 0, viol_cursor = 0;   // values handed out by the recorded calls
    void push(int e) { if (n < 16) events[n++] = e; }
    double next_cost() { return cost_cursor < n_script ? cost_script[cost_cursor++] : 0.0; }
    double next_viol() { return viol_cursor < n_script ? viol_script[viol_cursor++] : 0.0; }
    bool is(std::initializer_list<int> want) const { if ((int)want.size() != n) return false; int k = 0; for (int e : want) if (events[k++] != e) return false; return true; }
};
inline HookProbe*& active_probe() { static thread_local HookProbe* p = nullptr; return p; }
} }

// ---- polynomial / spline tags ------------------------------------------------------------------------------------------
namespace polympc {
enum collocation_scheme { GAUSS, GAUSS_RADAU, GAUSS_LOBATTO };
template <int PolyOrder, collocation_scheme Qtype = GAUSS_LOBATTO, typename _Scalar = double>
class Chebyshev {
public:
    static_assert(Qtype == GAUSS_LOBATTO, "the engine transcribes on Chebyshev-Gauss-Lobatto nodes");
    enum { POLY_ORDER = PolyOrder, NUM_NODES = PolyOrder + 1 };
    using scalar_t = _Scalar;
    using nodes_t = Eigen::Matrix<_Scalar, PolyOrder + 1, 1>;
    using q_weights_t = Eigen::Matrix<_Scalar, PolyOrder + 1, 1>;
    using diff_mat_t = Eigen::Matrix<_Scalar, PolyOrder + 1, PolyOrder + 1>;
    /** tables of the engine (pmb_cheb_tables), same formulas as ebyshev.hpp:111-214 */
    static nodes_t compute_nodes() { nodes_t n; diff_mat_t D; q_weights_t w; pmb_cheb_tables(PolyOrder, n.data(), D.data(), w.data()); return n; }
    static diff_mat_t compute_diff_matrix() { nodes_t n; diff_mat_t D; q_weights_t w; pmb_cheb_tables(PolyOrder, n.data(), D.data(), w.data()); return D; }
    static q_weights_t compute_int_weights() { nodes_t n; diff_mat_t D; q_weights_t w; pmb_cheb_tables(PolyOrder, n.data(), D.data(), w.data()); return w; }
};
template <typename Polynomial, int NumSegments>
class Spline {
public:
    enum { POLY_ORDER = Polynomial::POLY_ORDER, NUM_SEGMENTS = NumSegments, NUM_NODES = POLY_ORDER * NUM_SEGMENTS + 1 };
    using scalar_t = typename Polynomial::scalar_t;
    using nodes_t = typename Polynomial::nodes_t;
    using q_weights_t = typename Polynomial::q_weights_t;
    using diff_mat_t = typename Polynomial::diff_mat_t;
    static diff_mat_t compute_diff_matrix() { return Polynomial::compute_diff_matrix(); }
    static q_weights_t compute_int_weights() { return Polynomial::compute_int_weights(); }
    static nodes_t compute_nodes() { return Polynomial::compute_nodes(); }
};
typedef std::chrono::time_point<std::chrono::system_clock> time_point;
inline time_point get_time() { return std::chrono::system_clock::now(); }
struct IdentityPreconditioner { static constexpr int pmb_engine_preconditioner = 0; };   // PMB_PRECOND_IDENTITY
template <typename T> POLYMPC_HD inline void ignore_unused_var(const T&) noexcept {}   // src/utils/helpers.hpp
/** qp_preconditioners.hpp:113-: a tag that selects the engine's device twin (pmb_sqp_set_preconditioner); FMT = DENSE (0) / SPARSE (1) */
template <typename Scalar, int N, int M, int FMT = 0> struct RuizEquilibration { static constexpr int pmb_engine_preconditioner = FMT == 1 ? 2 : 1; };
} // namespace polympc

enum MEMORY { DENSE = 0, SPARSE = 1 };
template <int FMT> struct linear_solver_traits { template <typename Type, int Flags, typename... Args> struct default_solver {}; };

template <typename Derived> struct polympc_traits;
template <typename T> struct polympc_traits<const T> : polympc_traits<T> {};
#define POLYMPC_FORWARD_DECLARATION(cNAME, cNX, cNU, cNP, cND, cNG, TYPE) \
    class cNAME;                                                           \
    template <> struct polympc_traits<cNAME> {                             \
        using Scalar = TYPE;                                               \
        enum { NX = cNX, NU = cNU, NP = cNP, ND = cND, NG = cNG };         \
    };

// ---- ContinuousOCP: sizes, types and the functor defaults (continuous_ocp.hpp:39-288) ----------------------------------
template <typename OCP, typename Approximation, int MatrixFormat = DENSE>
class ContinuousOCP {
public:
    enum {
        NX = polympc_traits<OCP>::NX, NU = polympc_traits<OCP>::NU, NP = polympc_traits<OCP>::NP,
        ND = polympc_traits<OCP>::ND, NG = polympc_traits<OCP>::NG,
        NUM_NODES = Approximation::NUM_NODES, POLY_ORDER = Approximation::POLY_ORDER, NUM_SEGMENTS = Approximation::NUM_SEGMENTS,
        VARX_SIZE = NX * NUM_NODES, VARU_SIZE = NU * NUM_NODES, VARP_SIZE = NP, VARD_SIZE = ND,
        VAR_SIZE = VARX_SIZE + VARU_SIZE + VARP_SIZE, NUM_EQ = VARX_SIZE, NUM_INEQ = NG * NUM_NODES, NUM_BOX = VAR_SIZE,
        DUAL_SIZE = NUM_EQ + NUM_INEQ + NUM_BOX,
        is_sparse = (MatrixFormat == SPARSE) ? 1 : 0, is_dense = is_sparse ? 0 : 1, MATRIXFMT = MatrixFormat
    };
    template <typename scalar_t> using state_t = Eigen::Matrix<scalar_t, NX, 1>;
    template <typename scalar_t> using control_t = Eigen::Matrix<scalar_t, NU, 1>;
    template <typename scalar_t> using parameter_t = Eigen::Matrix<scalar_t, NP, 1>;
    template <typename scalar_t> using constraint_t = Eigen::Matrix<scalar_t, NG, 1>;
    using scalar_t = typename polympc_traits<OCP>::Scalar;
    static_assert(std::is_same<scalar_t, double>::value, "the engine computes in fp64");
    using static_parameter_t = Eigen::Matrix<scalar_t, ND, 1>;
    using time_t = Eigen::Matrix<scalar_t, NUM_NODES, 1>;
    using nodes_t = typename Approximation::nodes_t;

    using nlp_variable_t = Eigen::Matrix<scalar_t, VAR_SIZE, 1>;
    using nlp_constraints_t = Eigen::Matrix<scalar_t, NUM_EQ + NUM_INEQ, 1>;
    using nlp_eq_constraints_t = Eigen::Matrix<scalar_t, NUM_EQ, 1>;
    using nlp_ineq_constraints_t = Eigen::Matrix<scalar_t, NUM_INEQ, 1>;
    using nlp_dual_t = Eigen::Matrix<scalar_t, DUAL_SIZE, 1>;
    using nlp_hessian_t = Eigen::Matrix<scalar_t, VAR_SIZE, VAR_SIZE>;                 // never materialised on the host
    using nlp_jacobian_t = Eigen::Matrix<scalar_t, NUM_EQ + NUM_INEQ, VAR_SIZE>;
    using nlp_eq_jacobian_t = Eigen::Matrix<scalar_t, NUM_EQ, VAR_SIZE>;

    scalar_t t_start{0};
    scalar_t t_stop{1};
    /** continuous_ocp.hpp:147-159 (the time grid itself lives in the engine's problem descriptor) */
    void set_time_limits(const scalar_t& t0, const scalar_t& tf) noexcept { t_start = t0; t_stop = tf; }

    // ---- the Problem concept SQPBase is written against (continuous_ocp.hpp:430-435, 579, 618-647; sqp_base.hpp:110-120):
    // host methods, one instance, evaluated by the transcription kernels through pmb_ocp_* (each call uploads this object's
    // data members and horizon first, so `ocp.Q.diagonal() << ...; ocp.cost(...)` behaves as in the reference).  `_lagrangian`
    // receives what the reference's dense overloads leave in it: the cost value.
    void cost(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p, scalar_t& cost_) const
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::COST); cost_ = pr->next_cost(); return; }
        chk(pmb_ocp_cost(engine(), 1, var.data(), p.data(), &cost_), "cost");
    }
    void cost_gradient(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p, scalar_t& cost_,
                       Eigen::Ref<nlp_variable_t> cost_grad) const
    { chk(pmb_ocp_cost_gradient(engine(), 1, var.data(), p.data(), &cost_, cost_grad.data()), "cost_gradient"); }
    void cost_gradient_hessian(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p, scalar_t& cost_,
                               Eigen::Ref<nlp_variable_t> cost_grad, Eigen::Ref<nlp_hessian_t> cost_hess) const
    { chk(pmb_ocp_cost_gradient_hessian(engine(), 1, var.data(), p.data(), &cost_, cost_grad.data(), cost_hess.data()), "cost_gradient_hessian"); }
    void equalities(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p,
                    Eigen::Ref<nlp_eq_constraints_t> c) const
    { chk(pmb_ocp_equalities(engine(), 1, var.data(), p.data(), c.data()), "equalities"); }
    void inequalities(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p,
                      Eigen::Ref<nlp_ineq_constraints_t> g) const
    { if (NUM_INEQ > 0) chk(pmb_ocp_inequalities(engine(), 1, var.data(), p.data(), g.data()), "inequalities"); }
    void equalities_linearised(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p,
                               Eigen::Ref<nlp_eq_constraints_t> c, Eigen::Ref<nlp_eq_jacobian_t> jac) const
    { chk(pmb_ocp_equalities_linearised(engine(), 1, var.data(), p.data(), c.data(), jac.data()), "equalities_linearised"); }
    void lagrangian_gradient(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p,
                             const Eigen::Ref<const nlp_dual_t>& lam, scalar_t& _lagrangian, Eigen::Ref<nlp_variable_t> lag_gradient,
                             Eigen::Ref<nlp_variable_t> cost_gradient, Eigen::Ref<nlp_constraints_t> g, Eigen::Ref<nlp_jacobian_t> jac_g) const
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::LAG_GRAD); return; }
        chk(pmb_ocp_lagrangian_gradient(engine(), 1, var.data(), p.data(), lam.data(), &_lagrangian, lag_gradient.data(), cost_gradient.data(),
                                        g.data(), jac_g.data()), "lagrangian_gradient");
    }
    void lagrangian_gradient_hessian(const Eigen::Ref<const nlp_variable_t>& var, const Eigen::Ref<const static_parameter_t>& p,
                                     const Eigen::Ref<const nlp_dual_t>& lam, scalar_t& _lagrangian, Eigen::Ref<nlp_variable_t> lag_gradient,
                                     Eigen::Ref<nlp_hessian_t> lag_hessian, Eigen::Ref<nlp_variable_t> cost_gradient,
                                     Eigen::Ref<nlp_constraints_t> g, Eigen::Ref<nlp_jacobian_t> jac_g) const
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::LAG_GRAD_HESS); return; }
        chk(pmb_ocp_lagrangian_gradient_hessian(engine(), 1, var.data(), p.data(), lam.data(), &_lagrangian, lag_gradient.data(),
                                                lag_hessian.data(), cost_gradient.data(), g.data(), jac_g.data()), "lagrangian_gradient_hessian");
    }
    /** hessian_update_impl (continuous_ocp.hpp:676-686): DENSE = BFGS_update, SPARSE = the block BFGS of 2303-2431; evaluated by
     *  the engine's operators (pmb_bfgs_update / pmb_ocp_block_bfgs_update) */
    void hessian_update_impl(Eigen::Ref<nlp_hessian_t> hessian, const Eigen::Ref<const nlp_variable_t> s,
                             const Eigen::Ref<const nlp_variable_t> y) const
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::HESS_UPDATE_OCP); return; }
        if (MatrixFormat == SPARSE) chk(pmb_ocp_block_bfgs_update(engine(), 1, hessian.data(), s.data(), y.data(), nullptr), "hessian_update_impl");
        else chk(pmb_bfgs_update(VAR_SIZE, 1, hessian.data(), s.data(), y.data(), nullptr), "hessian_update_impl");
    }
    /** time grid of the NLP variable, final time first (continuous_ocp.hpp:45-66) */
    time_t time_nodes_now() const { time_t t; chk(pmb_ocp_time_nodes(engine(), t.data()), "time_nodes"); return t; }

private:
    static void chk(int rc, const char* what)
    { if (rc != PMB_OK) throw std::runtime_error(std::string("ContinuousOCP::") + what + " failed (" + std::to_string(rc) + "): " + pmb_last_error()); }
    /** one engine-side descriptor per problem class (this object stays trivially copyable); refreshed from *this on every call */
    pmb_ocp_t* engine() const
    {
        static pmb_ocp_t* h = pmb_ocp_create(pmb::compat::problem_name<OCP>(), 0);
        if (!h) throw std::runtime_error(std::string("pmb_ocp_create: ") + pmb_last_error());
        const OCP& self = *static_cast<const OCP*>(this);
        double blob[(sizeof(OCP) + 7) / 8] = {};
        std::memcpy(blob, (const void*)&self, sizeof(OCP));
        chk(pmb_ocp_set_params(h, blob, (int)((sizeof(OCP) + 7) / 8)), "set_params");
        chk(pmb_ocp_set_time_limits(h, t_start, t_stop), "set_time_limits");
        return h;
    }

public:
    /** defaults of the functor concept: no-ops, like continuous_ocp.hpp:206-216, 247-257, 278-288 */
    template <typename T>
    POLYMPC_HD void inequality_constraints_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                                const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                                const scalar_t& t, Eigen::Ref<constraint_t<T>> g) const noexcept {}
    template <typename T>
    POLYMPC_HD void mayer_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                    const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                    const scalar_t& t, T& mayer) noexcept { mayer = T(0.0); }
    template <typename T>
    POLYMPC_HD void lagrange_term_impl(const Eigen::Ref<const state_t<T>> x, const Eigen::Ref<const control_t<T>> u,
                                       const Eigen::Ref<const parameter_t<T>> p, const Eigen::Ref<const static_parameter_t> d,
                                       const scalar_t& t, T& lagrange) noexcept { lagrange = T(0.0); }
};

// ---- settings / info structs, field for field (sqp_base.hpp:24-61, qp_base.hpp:17-72) -----------------------------------
template <typename Scalar>
struct sqp_settings_t {
    Scalar tau = 0.5, eta = 0.25, rho = 0.5, eps_prim = 1e-3, eps_dual = 1e-3;
    int max_iter = 100;
    int line_search_max_iter = 100;
    void (*iteration_callback)(void* solver) = nullptr;   // host callback: cannot be honoured by a fused device loop
    bool validate() const
    { return 0.0 < tau && tau < 1.0 && 0.0 < eta && eta < 1.0 && 0.0 < rho && rho < 1.0 && eps_prim > 0.0 && eps_dual > 0.0 && max_iter > 0 && line_search_max_iter > 0; }
};
struct sqp_status_t { enum { SOLVED, MAX_ITER_EXCEEDED, INVALID_SETTINGS } value; };
struct sqp_info_t { int iter; int qp_solver_iter; sqp_status_t status; };

template <typename Scalar>
struct qp_solver_settings_t {
    Scalar rho = 1e-1, sigma = 1e-6, alpha = 1.0, eps_rel = 1e-3, eps_abs = 1e-3;
    int max_iter = 1000, check_termination = 25;
    bool warm_start = false, adaptive_rho = false, reuse_pattern = false, verbose = false;
    Scalar adaptive_rho_tolerance = 5;
    int adaptive_rho_interval = 25;
};
/** QP solver types.  boxADMM<> is the engine's inner solver (boxADMM + dense pivoted LDL^T, csrc/pmb_qp.hpp); the template
 *  arguments are accepted so that `boxADMM<VAR_SIZE, NUM_EQ, scalar_t, MATRIXFMT, linear_solver_traits<FMT>::default_solver>`
 *  spells the same.  Used stand-alone it is the QPBase object concept (qp_base.hpp:148-175): solve() with 7 or 9 arguments runs
 *  ONE instance through pmb_qp_solve on the device; primal_solution(), dual_solution() = [y_A ; y_box], info(), settings(). */
typedef enum { SOLVED, MAX_ITER_EXCEEDED, UNSOLVED, UNINITIALIZED, INFEASIBLE, INCONSISTENT } status_t;   // qp_base.hpp:55-62
template <typename Scalar>
struct qp_solver_info_t { status_t status = UNINITIALIZED; int iter = 0; int rho_updates = 0; Scalar rho_estimate = 0, res_prim = 1, res_dual = 1; };

namespace pmb { namespace compat {
/** the QPBase object concept (qp_base.hpp:148-175) shared by boxADMM<> and ADMM<>: one instance through pmb_qp_solve /
 *  pmb_qp_solve_admm on the device */
template <int N, int M, typename Scalar, bool OSQP_SPLITTING>
struct QpObject {
    using scalar_t = Scalar;
    using settings_t = qp_solver_settings_t<Scalar>;
    using info_t = qp_solver_info_t<Scalar>;
    using qp_var_t = Eigen::Matrix<Scalar, N, 1>;
    using qp_dual_t = Eigen::Matrix<Scalar, N + M, 1>;
    using qp_dual_a_t = Eigen::Matrix<Scalar, M, 1>;
    using qp_hessian_t = Eigen::Matrix<Scalar, N, N>;
    using qp_constraint_t = Eigen::Matrix<Scalar, M, N>;
    settings_t m_settings;
    info_t m_info;
    qp_var_t m_x;
    qp_dual_t m_y;
    int iter{0};                                  // box_admm.hpp:44 (public in the reference, read by its tests)
    QpObject() { m_x.setZero(); m_y.setZero(); }
    settings_t& settings() noexcept { return m_settings; }
    const settings_t& settings() const noexcept { return m_settings; }
    const info_t& info() const noexcept { return m_info; }
    const qp_var_t& primal_solution() const noexcept { return m_x; }
    qp_var_t& primal_solution() noexcept { return m_x; }
    const qp_dual_t& dual_solution() const noexcept { return m_y; }
    qp_dual_t& dual_solution() noexcept { return m_y; }

    /** QPBase::solve, 7 arguments (cold start, qp_base.hpp:161-166) and 9 arguments (with guesses, 168-175) */
    status_t solve(const Eigen::Ref<const qp_hessian_t>& H, const Eigen::Ref<const qp_var_t>& h, const Eigen::Ref<const qp_constraint_t>& A,
                   const Eigen::Ref<const qp_dual_a_t>& Alb, const Eigen::Ref<const qp_dual_a_t>& Aub,
                   const Eigen::Ref<const qp_var_t>& xlb, const Eigen::Ref<const qp_var_t>& xub)
    { return run(H.data(), h.data(), A.data(), Alb.data(), Aub.data(), xlb.data(), xub.data(), nullptr, nullptr); }
    status_t solve(const Eigen::Ref<const qp_hessian_t>& H, const Eigen::Ref<const qp_var_t>& h, const Eigen::Ref<const qp_constraint_t>& A,
                   const Eigen::Ref<const qp_dual_a_t>& Alb, const Eigen::Ref<const qp_dual_a_t>& Aub,
                   const Eigen::Ref<const qp_var_t>& xlb, const Eigen::Ref<const qp_var_t>& xub,
                   const Eigen::Ref<const qp_var_t>& x_guess, const Eigen::Ref<const qp_dual_t>& y_guess)
    { return run(H.data(), h.data(), A.data(), Alb.data(), Aub.data(), xlb.data(), xub.data(), x_guess.data(), y_guess.data()); }

private:
    status_t run(const Scalar* H, const Scalar* h, const Scalar* A, const Scalar* Alb, const Scalar* Aub, const Scalar* xlb,
                 const Scalar* xub, const Scalar* xg, const Scalar* yg)
    {
        pmb_qp_settings_t q; pmb_qp_default_settings(&q);
        const settings_t& s = m_settings;
        q.rho = s.rho; q.sigma = s.sigma; q.alpha = s.alpha; q.eps_rel = s.eps_rel; q.eps_abs = s.eps_abs; q.max_iter = s.max_iter;
        q.check_termination = s.check_termination; q.warm_start = s.warm_start; q.adaptive_rho = s.adaptive_rho;
        q.adaptive_rho_tolerance = s.adaptive_rho_tolerance; q.adaptive_rho_interval = s.adaptive_rho_interval;
        q.reuse_pattern = s.reuse_pattern; q.verbose = s.verbose;
        pmb_qp_info_t inf;
        const int rc = OSQP_SPLITTING
            ? pmb_qp_solve_admm(N, M, 1, H, h, A, Alb, Aub, xlb, xub, xg, yg, &q, m_x.data(), m_y.data(), &inf, nullptr, nullptr, nullptr, nullptr)
            : pmb_qp_solve(N, M, 1, H, h, A, Alb, Aub, xlb, xub, xg, yg, &q, m_x.data(), m_y.data(), &inf, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc != PMB_OK) throw std::runtime_error(std::string(OSQP_SPLITTING ? "ADMM::solve failed (" : "boxADMM::solve failed (") + std::to_string(rc) + "): " + pmb_last_error());
        const int prev_updates = m_info.rho_updates;           // accumulates across solves (box_admm.hpp:395)
        m_info.status = (status_t)inf.status; m_info.iter = inf.iter; m_info.rho_updates = prev_updates + inf.rho_updates;
        iter = inf.iter;
        m_info.rho_estimate = inf.rho_estimate; m_info.res_prim = inf.res_prim; m_info.res_dual = inf.res_dual;
        return m_info.status;
    }
};
} }
template <int N, int M, typename Scalar = double, int MatrixType = DENSE,
          template <typename, int, typename...> class LinearSolver = linear_solver_traits<DENSE>::template default_solver, int LinearSolver_UpLo = 1>
struct boxADMM : pmb::compat::QpObject<N, M, Scalar, false> {
    static constexpr bool pmb_engine_inner_solver = true;   // SQPBase::solve() refuses QP solver types without this tag
    static constexpr int pmb_engine_qp_solver = 0;          // PMB_QP_BOX_ADMM
};
/** the OSQP-style ADMM of src/solvers/admm.hpp (box constraints stacked under A, a (2N + M)-dimensional KKT system): a QPBase
 *  object stand-alone (pmb_qp_solve_admm) and a selectable QP solver of the fused SQP loop (pmb_sqp_set_qp_solver; exact
 *  arithmetic, 2N + M <= 256) */
template <int N, int M, typename Scalar = double, int MatrixType = DENSE,
          template <typename, int, typename...> class LinearSolver = linear_solver_traits<DENSE>::template default_solver, int LinearSolver_UpLo = 1>
struct ADMM : pmb::compat::QpObject<N, M, Scalar, true> {
    static constexpr bool pmb_engine_inner_solver = true;
    static constexpr int pmb_engine_qp_solver = 1;          // PMB_QP_OSQP_ADMM
};

// ---- device side: adapter from the Eigen-style functors to the engine's functor concept ---------------------------------
#if defined(__CUDACC__) || defined(PMB_EMU)
#include "pmb_registry.hpp"
namespace pmb {
namespace compat {

template <class U>
struct Model {
    static constexpr int NX = U::NX, NU = U::NU, NP = U::NP, ND = U::ND, NG = U::NG;
    static constexpr int NPARAM = (int)((sizeof(U) + 7) / 8);   // the problem object travels as an opaque blob
    U ocp;
    void defaults() { ocp = U(); }
    void set_params(const double* v) { std::memcpy((void*)&ocp, v, sizeof(U)); }
    void get_params(double* v) const { std::memset(v, 0, sizeof(double) * NPARAM); std::memcpy(v, (const void*)&ocp, sizeof(U)); }

    template <class T> using W = typename std::conditional<std::is_same<T, double>::value, Dual<double, 0>, T>::type;
    template <class T> using St = Eigen::Matrix<T, NX, 1>;
    template <class T> using Ct = Eigen::Matrix<T, NU, 1>;
    template <class T> using Pt = Eigen::Matrix<T, NP, 1>;
    template <class T> using Gt = Eigen::Matrix<T, NG, 1>;
    using Dt = Eigen::Matrix<double, ND, 1>;
    PMB_HD U& self() const { return const_cast<U&>(ocp); }

    template <class T>
    PMB_HD void dynamics(const T* x, const T* u, const T* p, const double* d, const T& t, T* xdot) const
    {
        if constexpr (std::is_same<T, double>::value) {
            using R = Dual<double, 0>;
            St<R> xs; Ct<R> us; Pt<R> ps; St<R> out;
            for (int i = 0; i < NX; ++i) xs.m[i].v = x[i];
            for (int i = 0; i < NU; ++i) us.m[i].v = u[i];
            for (int i = 0; i < NP; ++i) ps.m[i].v = p[i];
            R tt; tt.v = t;
            self().template dynamics_impl<R>(Eigen::Ref<const St<R>>(xs.m), Eigen::Ref<const Ct<R>>(us.m), Eigen::Ref<const Pt<R>>(ps.m),
                                             Eigen::Ref<const Dt>(d), tt, Eigen::Ref<St<R>>(out.m));
            for (int i = 0; i < NX; ++i) xdot[i] = out.m[i].v;
        } else {
            self().template dynamics_impl<T>(Eigen::Ref<const St<T>>(x), Eigen::Ref<const Ct<T>>(u), Eigen::Ref<const Pt<T>>(p),
                                             Eigen::Ref<const Dt>(d), t, Eigen::Ref<St<T>>(xdot));
        }
    }
    template <class T>
    PMB_HD void ineq(const T* x, const T* u, const T* p, const double* d, double t, T* g) const
    {
        if constexpr (NG == 0) { return; }
        else if constexpr (std::is_same<T, double>::value) {
            using R = Dual<double, 0>;
            St<R> xs; Ct<R> us; Pt<R> ps; Gt<R> out;
            for (int i = 0; i < NX; ++i) xs.m[i].v = x[i];
            for (int i = 0; i < NU; ++i) us.m[i].v = u[i];
            for (int i = 0; i < NP; ++i) ps.m[i].v = p[i];
            self().template inequality_constraints_impl<R>(Eigen::Ref<const St<R>>(xs.m), Eigen::Ref<const Ct<R>>(us.m), Eigen::Ref<const Pt<R>>(ps.m),
                                                           Eigen::Ref<const Dt>(d), t, Eigen::Ref<Gt<R>>(out.m));
            for (int i = 0; i < NG; ++i) g[i] = out.m[i].v;
        } else {
            self().template inequality_constraints_impl<T>(Eigen::Ref<const St<T>>(x), Eigen::Ref<const Ct<T>>(u), Eigen::Ref<const Pt<T>>(p),
                                                           Eigen::Ref<const Dt>(d), t, Eigen::Ref<Gt<T>>(g));
        }
    }
#define PMB_COMPAT_SCALAR_FUNCTOR(NAME, IMPL)                                                                                  \
    template <class T>                                                                                                         \
    PMB_HD void NAME(const T* x, const T* u, const T* p, const double* d, double t, T& out) const                              \
    {                                                                                                                          \
        if constexpr (std::is_same<T, double>::value) {                                                                        \
            using R = Dual<double, 0>;                                                                                         \
            St<R> xs; Ct<R> us; Pt<R> ps; R o;                                                                                 \
            for (int i = 0; i < NX; ++i) xs.m[i].v = x[i];                                                                     \
            for (int i = 0; i < NU; ++i) us.m[i].v = u[i];                                                                     \
            for (int i = 0; i < NP; ++i) ps.m[i].v = p[i];                                                                     \
            self().template IMPL<R>(Eigen::Ref<const St<R>>(xs.m), Eigen::Ref<const Ct<R>>(us.m), Eigen::Ref<const Pt<R>>(ps.m), \
                                    Eigen::Ref<const Dt>(d), t, o);                                                            \
            out = o.v;                                                                                                         \
        } else {                                                                                                               \
            self().template IMPL<T>(Eigen::Ref<const St<T>>(x), Eigen::Ref<const Ct<T>>(u), Eigen::Ref<const Pt<T>>(p),        \
                                    Eigen::Ref<const Dt>(d), t, out);                                                          \
        }                                                                                                                      \
    }
    PMB_COMPAT_SCALAR_FUNCTOR(lagrange, lagrange_term_impl)
    PMB_COMPAT_SCALAR_FUNCTOR(mayer, mayer_term_impl)
#undef PMB_COMPAT_SCALAR_FUNCTOR
};

/** found by argument-dependent lookup from Ocp<>::init(): the horizon a problem class set in its constructor */
template <class U> inline void model_time_limits(const Model<U>& m, double& t0, double& tf) { t0 = m.ocp.t_start; tf = m.ocp.t_stop; }

template <class U> using OcpOf = Ocp<Model<U>, U::POLY_ORDER, U::NUM_SEGMENTS>;
template <class U> IProblem* make_problem() { return new ProblemImpl<OcpOf<U>>(); }

/** registers the kernels of a problem class compiled in THIS translation unit with the engine (once) and returns the name
 *  to give to pmb_sqp_create / polympc::b200::BatchedMPC */
template <class U>
const char* problem_name()
{
    static const std::string name = [] {
        std::string n = std::string("compat:") + typeid(U).name();
        const int rc = pmb_register_problem(n.c_str(), (void* (*)())(&make_problem<U>));
        if (rc != PMB_OK) throw std::runtime_error(std::string("pmb_register_problem: ") + pmb_last_error());
        return n;
    }();
    return name.c_str();
}

} // namespace compat
} // namespace pmb
/** in-library registration of a reference-style problem class (problems/*.cu) */
#define PMB_DEFINE_COMPAT_PROBLEM(ID, USER_OCP) \
    namespace pmb { IProblem* pmb_make_##ID() { return compat::make_problem<USER_OCP>(); } }
#endif // __CUDACC__ || PMB_EMU

// ---- LSFilter (src/solvers/line_search.hpp:30-98): the host-side filter object of a solver ----------------------------------
// The engine keeps one filter per instance on the device (pmb_sqp_set_line_search); SQPBase::solve() moves the contents of the
// solver's `filter` member there before the solve and back afterwards, so `solver.filter.beta = ...`, `.clear()` and reading
// `.m_filter` behave as with the reference.  Front of the list = newest entry.
template <typename Scalar>
class LSFilter {
public:
    using scalar_t = Scalar;
    using filter_pair_t = std::pair<Scalar, Scalar>;     // (cost, constraint violation)
    std::list<filter_pair_t> m_filter;
    int max_depth{10};
    scalar_t beta{1e-5};

    void print() const noexcept { std::cout << "Filter: \n"; for (const auto& e : m_filter) std::cout << "( " << e.first << " , " << e.second << " )\n"; }
    void clear() noexcept { m_filter.clear(); }
    bool is_dominated(const scalar_t& cost, const scalar_t& constraint) const noexcept
    { for (const auto& e : m_filter) if (e.first <= cost && e.second <= constraint) return true; return false; }
    /** no entry may be better in both coordinates, with the margin beta * (its violation) */
    bool is_acceptable(const scalar_t& cost, const scalar_t& constraint) const noexcept
    {
        for (const auto& e : m_filter) { const scalar_t margin = beta * e.second; if (e.first - margin <= cost && e.second - margin <= constraint) return false; }
        return true;
    }
    /** below max_depth: entries the new point dominates leave; at max_depth: the oldest leaves */
    void add(const scalar_t& cost, const scalar_t& constraint) noexcept
    {
        if ((int)m_filter.size() < max_depth) m_filter.remove_if([&](const filter_pair_t& e) { return e.first >= cost && e.second >= constraint; });
        else m_filter.pop_back();
        m_filter.emplace_front(cost, constraint);
    }
    void remove_one() noexcept { m_filter.pop_back(); }
};

// ---- SQPBase: one NLP instance on the host, solved by the engine (sqp_base.hpp:64-197, 568-696) --------------------------
template <typename Derived, typename Problem,
          typename QPSolver = boxADMM<Problem::VAR_SIZE, Problem::NUM_EQ + Problem::NUM_INEQ, typename Problem::scalar_t>,
          typename Preconditioner = polympc::IdentityPreconditioner>
class SQPBase {
public:
    enum { VAR_SIZE = Problem::VAR_SIZE, NUM_EQ = Problem::NUM_EQ, NUM_INEQ = Problem::NUM_INEQ, NUM_CONSTR = Problem::DUAL_SIZE };
    using nlp_variable_t = typename Problem::nlp_variable_t;
    using nlp_constraints_t = typename Problem::nlp_constraints_t;
    using nlp_eq_constraints_t = typename Problem::nlp_eq_constraints_t;
    using nlp_ineq_constraints_t = typename Problem::nlp_ineq_constraints_t;
    using nlp_eq_jacobian_t = typename Problem::nlp_eq_jacobian_t;
    using nlp_jacobian_t = typename Problem::nlp_jacobian_t;
    using nlp_hessian_t = typename Problem::nlp_hessian_t;
    using nlp_cost_t = typename Problem::scalar_t;
    using nlp_dual_t = typename Problem::nlp_dual_t;
    using scalar_t = typename Problem::scalar_t;
    using parameter_t = typename Problem::static_parameter_t;
    using qp_solver_t = QPSolver;
    using nlp_settings_t = sqp_settings_t<scalar_t>;
    using nlp_info_t = sqp_info_t;

    /** what the hook probe found (see the header comment); filled by the first solve(), readable afterwards */
    struct engine_options_t {
        bool exact_hessian_every_iteration = false;   // update_linearisation_*_impl forwards to linearisation_*_impl
        bool gershgorin_regularisation = false;       // hessian_regularisation_*_impl is the Gershgorin shift
        bool block_bfgs = false;                      // hessian_update_impl forwards to the SPARSE problem's block BFGS
        bool filter_line_search = false;              // step_size_selection_impl is the LSFilter line search of valet_parking_mpc_test.cpp
        int preconditioner = 0;                       // pmb_preconditioner_t of the Preconditioner template argument
        bool probed = false;
        std::string refused;                          // non-empty: why solve() refuses to run (status INVALID_SETTINGS)
    };
    engine_options_t m_engine_options;
    engine_options_t& engine_options() noexcept { return m_engine_options; }
    const engine_options_t& engine_options() const noexcept { return m_engine_options; }

    Problem problem;
    nlp_settings_t m_settings;
    nlp_info_t m_info;
    qp_solver_t m_qp_solver;
    nlp_variable_t m_x, m_lbx, m_ubx;
    nlp_dual_t m_lam;
    nlp_ineq_constraints_t m_lbg, m_ubg;
    parameter_t m_p;
    scalar_t m_cost = scalar_t(0);  // written by user line searches (sqp_base.hpp:139); cost() reports the engine's value
    nlp_hessian_t m_H;              // host mirror used by the hook probe only (user overrides read this->m_H.rows())
    nlp_variable_t m_lag_gradient;

    SQPBase()
    {
        const scalar_t INF = std::numeric_limits<scalar_t>::infinity();                       // sqp_base.hpp:72-99
        m_lbx.setConstant(-INF); m_ubx.setConstant(INF); m_lbg.setConstant(-INF); m_ubg.setConstant(INF);
        m_x.setZero(); m_lam.setZero(); m_p.setZero();
        m_info.iter = 0; m_info.qp_solver_iter = 0; m_info.status.value = sqp_status_t::MAX_ITER_EXCEEDED;
        pmb_qp_settings_t q; pmb_sqp_default_qp_settings(&q);
        auto& s = m_qp_solver.m_settings;
        s.warm_start = q.warm_start; s.check_termination = q.check_termination; s.eps_abs = q.eps_abs; s.eps_rel = q.eps_rel;
        s.max_iter = q.max_iter; s.adaptive_rho = q.adaptive_rho; s.adaptive_rho_interval = q.adaptive_rho_interval; s.alpha = q.alpha;
        s.rho = q.rho; s.sigma = q.sigma; s.adaptive_rho_tolerance = q.adaptive_rho_tolerance;
    }
    ~SQPBase() { if (m_handle) pmb_sqp_destroy(m_handle); }
    SQPBase(const SQPBase&) = delete;
    SQPBase& operator=(const SQPBase&) = delete;

    const Problem& get_problem() const noexcept { return problem; }
    Problem& get_problem() noexcept { return problem; }
    const nlp_variable_t& primal_solution() const noexcept { return m_x; }
    nlp_variable_t& primal_solution() noexcept { return m_x; }
    const nlp_dual_t& dual_solution() const noexcept { return m_lam; }
    nlp_dual_t& dual_solution() noexcept { return m_lam; }
    const nlp_settings_t& settings() const noexcept { return m_settings; }
    nlp_settings_t& settings() noexcept { return m_settings; }
    const sqp_info_t& info() const noexcept { return m_info; }
    sqp_info_t& info() noexcept { return m_info; }
    const nlp_variable_t& lower_bound_x() const noexcept { return m_lbx; }
    nlp_variable_t& lower_bound_x() noexcept { return m_lbx; }
    const nlp_variable_t& upper_bound_x() const noexcept { return m_ubx; }
    nlp_variable_t& upper_bound_x() noexcept { return m_ubx; }
    const nlp_ineq_constraints_t& lower_bound_g() const noexcept { return m_lbg; }
    nlp_ineq_constraints_t& lower_bound_g() noexcept { return m_lbg; }
    const nlp_ineq_constraints_t& upper_bound_g() const noexcept { return m_ubg; }
    nlp_ineq_constraints_t& upper_bound_g() noexcept { return m_ubg; }
    const parameter_t& parameters() const noexcept { return m_p; }
    parameter_t& parameters() noexcept { return m_p; }
    const typename qp_solver_t::settings_t& qp_settings() const noexcept { return m_qp_solver.m_settings; }
    typename qp_solver_t::settings_t& qp_settings() noexcept { return m_qp_solver.m_settings; }
    scalar_t primal_norm() const noexcept { return m_stats[1]; }
    scalar_t dual_norm() const noexcept { return m_stats[2]; }
    scalar_t constr_violation() const noexcept { return m_stats[3]; }
    scalar_t cost() const noexcept { return m_stats[0]; }
    /** device time of the last solve, ms (not in the reference) */
    double last_solve_ms() const { return m_handle ? pmb_sqp_last_solve_ms(m_handle) : 0.0; }

    void solve() { solve_impl(); }
    void solve(const Eigen::Ref<const nlp_variable_t>& x_guess, const Eigen::Ref<const nlp_dual_t>& lam_guess)   // sqp_base.hpp:558-566
    { m_x = x_guess; m_lam = lam_guess; solve_impl(); }

    // ---- CRTP hooks, default implementations (sqp_base.hpp:263-350).  The fused kernel runs their device twins; on the host
    // they exist so that user solvers which override or call them compile, and so that solve() can probe the overrides.
    Derived& derived() noexcept { return *static_cast<Derived*>(this); }
    void hessian_update(Eigen::Ref<nlp_hessian_t> hessian, const Eigen::Ref<const nlp_variable_t>& x_step,
                        const Eigen::Ref<const nlp_variable_t>& grad_step) noexcept
    { derived().hessian_update_impl(hessian, x_step, grad_step); }
    /** default: dense damped BFGS (bfgs.hpp:23-52) */
    void hessian_update_impl(Eigen::Ref<nlp_hessian_t> hessian, const Eigen::Ref<const nlp_variable_t>& x_step,
                             const Eigen::Ref<const nlp_variable_t>& grad_step) noexcept
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::HESS_UPDATE_DEFAULT); return; }
        pmb_bfgs_update(VAR_SIZE, 1, hessian.data(), x_step.data(), grad_step.data(), nullptr);
    }
    void hessian_regularisation_dense_impl(Eigen::Ref<nlp_hessian_t>) noexcept {}
    void hessian_regularisation_sparse_impl(Eigen::Ref<nlp_hessian_t>) noexcept {}
    void linearisation_dense_impl(const Eigen::Ref<const nlp_variable_t>& x, const Eigen::Ref<const parameter_t>& p,
                                  const Eigen::Ref<const nlp_dual_t>& lam, Eigen::Ref<nlp_variable_t> cost_grad,
                                  Eigen::Ref<nlp_hessian_t> lag_hessian, Eigen::Ref<nlp_jacobian_t> A, Eigen::Ref<nlp_constraints_t> b) noexcept
    {
        scalar_t lag(0.0);
        problem.lagrangian_gradient_hessian(x, p, lam, lag, m_lag_gradient, lag_hessian, cost_grad, b, A);
        derived().hessian_regularisation_dense_impl(m_H);          // the member, like sqp_base.hpp:316-317
    }
    void linearisation_sparse_impl(const Eigen::Ref<const nlp_variable_t>& x, const Eigen::Ref<const parameter_t>& p,
                                   const Eigen::Ref<const nlp_dual_t>& lam, Eigen::Ref<nlp_variable_t> cost_grad,
                                   nlp_hessian_t& lag_hessian, nlp_jacobian_t& A, Eigen::Ref<nlp_constraints_t> b) noexcept
    {
        scalar_t lag(0.0);
        problem.lagrangian_gradient_hessian(x, p, lam, lag, m_lag_gradient, lag_hessian, cost_grad, b, A);
        derived().hessian_regularisation_sparse_impl(m_H);
    }
    void update_linearisation_dense_impl(const Eigen::Ref<const nlp_variable_t>& x, const Eigen::Ref<const parameter_t>& p,
                                         const Eigen::Ref<const nlp_variable_t>& x_step, const Eigen::Ref<const nlp_dual_t>& lam,
                                         Eigen::Ref<nlp_variable_t> cost_grad, Eigen::Ref<nlp_hessian_t> lag_hessian,
                                         Eigen::Ref<nlp_jacobian_t> A, Eigen::Ref<nlp_constraints_t> b) noexcept
    {
        scalar_t lag(0.0);
        nlp_variable_t lag_grad;
        problem.lagrangian_gradient(x, p, lam, lag, lag_grad, cost_grad, b, A);
        hessian_update(lag_hessian, x_step, nlp_variable_t(lag_grad - m_lag_gradient));
        m_lag_gradient = lag_grad;
    }
    void update_linearisation_sparse_impl(const Eigen::Ref<const nlp_variable_t>& x, const Eigen::Ref<const parameter_t>& p,
                                          const Eigen::Ref<const nlp_variable_t>& x_step, const Eigen::Ref<const nlp_dual_t>& lam,
                                          Eigen::Ref<nlp_variable_t> cost_grad, nlp_hessian_t& lag_hessian, nlp_jacobian_t& A,
                                          Eigen::Ref<nlp_constraints_t> b) noexcept
    {
        scalar_t lag(0.0);
        nlp_variable_t lag_grad;
        problem.lagrangian_gradient(x, p, lam, lag, lag_grad, cost_grad, b, A);
        hessian_update(lag_hessian, x_step, nlp_variable_t(lag_grad - m_lag_gradient));
        m_lag_gradient = lag_grad;
    }
    /** constraints_violation (sqp_base.hpp:233-236, 421-444) for user hooks that call it: eps + |c|_1 + the bound violations,
     *  evaluated with the engine's operators; under the hook probe it records the call and returns the scripted value */
    scalar_t constraints_violation(const Eigen::Ref<const nlp_variable_t>& x) noexcept
    {
        if (pmb::compat::HookProbe* pr = pmb::compat::active_probe()) { pr->push(pmb::compat::HookProbe::VIOLATION); return scalar_t(pr->next_viol()); }
        scalar_t v = std::numeric_limits<scalar_t>::epsilon();
        nlp_eq_constraints_t c; problem.equalities(x, m_p, c);
        for (int i = 0; i < NUM_EQ; ++i) v += std::fabs(c.data()[i]);
        if (NUM_INEQ > 0) {
            nlp_ineq_constraints_t g; problem.inequalities(x, m_p, g);
            for (int i = 0; i < NUM_INEQ; ++i) v += std::fmax(m_lbg.data()[i] - g.data()[i], scalar_t(0)) + std::fmax(g.data()[i] - m_ubg.data()[i], scalar_t(0));
        }
        for (int i = 0; i < VAR_SIZE; ++i) v += std::fmax(m_lbx.data()[i] - x.data()[i], scalar_t(0)) + std::fmax(x.data()[i] - m_ubx.data()[i], scalar_t(0));
        return v;
    }
    /** defaults of the hooks that have no engine-side alternative: overriding any of them is refused — except a
     *  step_size_selection_impl that is the filter line search (see probe_hooks) */
    scalar_t step_size_selection_impl(const Eigen::Ref<const nlp_variable_t>&) noexcept { return scalar_t(1); }
    scalar_t constraints_violation_impl(const Eigen::Ref<const nlp_variable_t>&) const noexcept { return scalar_t(0); }
    scalar_t max_constraints_violation_impl(const Eigen::Ref<const nlp_variable_t>&) const noexcept { return scalar_t(0); }
    bool termination_criteria_impl(const Eigen::Ref<const nlp_variable_t>&) noexcept { return false; }

private:
    // `&Derived::name` names the base's member (type `R (SQPBase::*)(...)`) unless Derived declares its own: comparing the
    // member-pointer types detects an override without calling it; an overloaded / templated override makes the expression
    // ill-formed and lands in the `true` fallback
#define PMB_COMPAT_OVERRIDES(NAME)                                                                                            \
    template <class D, class = void> struct overrides_##NAME : std::true_type {};                                             \
    template <class D> struct overrides_##NAME<D, decltype(void(&D::NAME))>                                                   \
        : std::integral_constant<bool, !std::is_same<decltype(&D::NAME), decltype(&SQPBase::NAME)>::value> {};
    PMB_COMPAT_OVERRIDES(step_size_selection_impl)
    PMB_COMPAT_OVERRIDES(constraints_violation_impl)
    PMB_COMPAT_OVERRIDES(max_constraints_violation_impl)
    PMB_COMPAT_OVERRIDES(termination_criteria_impl)
    PMB_COMPAT_OVERRIDES(linearisation_dense_impl)
    PMB_COMPAT_OVERRIDES(linearisation_sparse_impl)
#undef PMB_COMPAT_OVERRIDES

    /** the solver's public `filter` member when it is an LSFilter<scalar_t> (valet_parking_mpc_test.cpp:112), else nullptr */
    template <class D> static auto filter_of(D& d, int)
        -> typename std::enable_if<std::is_same<decltype(d.filter), LSFilter<scalar_t>>::value, LSFilter<scalar_t>*>::type { return &d.filter; }
    template <class D> static LSFilter<scalar_t>* filter_of(D&, long) { return nullptr; }

    /** host model of the engine's filter line search (csrc/pmb_sqp.hpp::step, LS_FILTER) on scripted values: cost[0], viol[0] are
     *  the values at x, cost[k], viol[k] those of trial k.  Returns the step length; evals = points evaluated. */
    static scalar_t filter_search_model(LSFilter<scalar_t>& f, const double* cost, const double* viol, scalar_t tau, int ls_max, int& evals)
    {
        evals = 1;
        if (f.is_acceptable(cost[0], viol[0])) f.add(cost[0], viol[0]);
        scalar_t alpha = 1;
        for (int i = 1; i < ls_max; ++i) {
            const double c = cost[evals], v = viol[evals];
            ++evals;
            if (f.is_acceptable(c, v)) { f.add(c, v); return alpha; }
            alpha *= tau;
        }
        return alpha;
    }

    /** Gershgorin shift of the reference's minimal_time_test.cpp:90-104 on the host (probe comparison only) */
    static void host_gershgorin(nlp_hessian_t& H)
    {
        for (int i = 0; i < VAR_SIZE; ++i) {
            const scalar_t aii = H(i, i);
            scalar_t sum = 0;
            for (int j = 0; j < VAR_SIZE; ++j) sum += std::fabs(H(j, i));
            const scalar_t ri = sum - std::fabs(aii);
            if (aii - ri <= 0) H(i, i) += (ri - aii) + scalar_t(0.01);
        }
    }
    static bool same(const nlp_hessian_t& a, const nlp_hessian_t& b, scalar_t tol)
    { for (int i = 0; i < VAR_SIZE * VAR_SIZE; ++i) if (!(std::fabs(a.data()[i] - b.data()[i]) <= tol * (1 + std::fabs(b.data()[i])))) return false; return true; }

    /** run the overridable hooks once on the host against the recording problem and map what they do onto the engine menu */
    void probe_hooks()
    {
        using P = pmb::compat::HookProbe;
        engine_options_t& eo = m_engine_options;
        eo = engine_options_t();
        eo.probed = true;
        auto refuse = [&](const std::string& why) { if (eo.refused.empty()) eo.refused = why; };
        if (overrides_step_size_selection_impl<Derived>::value) {
            // the one recognised override: the filter line search of reference tests/control/valet_parking_mpc_test.cpp:110-155 —
            // violation and cost at x enter the filter if acceptable, then backtracking until the filter accepts a trial point.
            // The hook is run against scripted (cost, violation) sequences and must agree with the host model of the device twin
            // (filter_search_model) in the step length, the number of evaluations and the filter it leaves behind.
            LSFilter<scalar_t>* flt = filter_of(derived(), 0);
            if (!flt) refuse("Derived::step_size_selection_impl is overridden and the solver has no public LSFilter member `filter` (only the l1-merit "
                             "and the filter line search have device twins)");
            else {
                const LSFilter<scalar_t> saved = *flt;
                const nlp_settings_t saved_settings = m_settings;
                const scalar_t saved_cost = m_cost;
                struct Scenario { int depth, ls_max, n_seed; double seed[2][2]; int n; double cost[5], viol[5]; };
                static const Scenario scenarios[4] = {
                    // empty filter; x accepted; trial 1 rejected; trial 2 dominates everything
                    {10, 10, 0, {{0, 0}, {0, 0}}, 3, {10, 20, -1000}, {10, 20, 0}},
                    // trial 2 has a HIGHER cost than x but a much smaller violation: the filter takes it, a merit-like extra test would not
                    {10, 10, 0, {{0, 0}, {0, 0}}, 3, {10, 20, 15}, {10, 20, 1}},
                    // a full filter (max_depth 2): x is not acceptable, trial 2 is and pushes the oldest entry out
                    {2, 10, 2, {{5, 5}, {6, 4}}, 3, {100, 4, 3}, {100, 4.5, 10}},
                    // nothing is ever accepted: line_search_max_iter - 1 trials, alpha = tau^3
                    {10, 4, 1, {{-1e9, -1e9}, {0, 0}}, 4, {1, 2, 3, 4}, {1, 2, 3, 4}},
                };
                bool ok = true;
                for (const Scenario& sc : scenarios) {
                    LSFilter<scalar_t> f0; f0.beta = scalar_t(0.25); f0.max_depth = sc.depth;
                    for (int k = 0; k < sc.n_seed; ++k) f0.m_filter.emplace_back(sc.seed[k][0], sc.seed[k][1]);
                    LSFilter<scalar_t> model = f0;
                    int evals = 0;
                    const scalar_t alpha_model = filter_search_model(model, sc.cost, sc.viol, m_settings.tau, sc.ls_max, evals);
                    *flt = f0;
                    m_settings.line_search_max_iter = sc.ls_max;
                    P rec;
                    for (int k = 0; k < sc.n; ++k) { rec.cost_script[k] = sc.cost[k]; rec.viol_script[k] = sc.viol[k]; }
                    rec.n_script = sc.n;
                    nlp_variable_t pz; pz.setZero();
                    pmb::compat::active_probe() = &rec;
                    const scalar_t alpha = derived().step_size_selection_impl(pz);
                    pmb::compat::active_probe() = nullptr;
                    ok = ok && alpha == alpha_model && rec.cost_cursor == evals && rec.viol_cursor == evals && rec.n == 2 * evals &&
                         flt->m_filter == model.m_filter;
                }
                *flt = saved; m_settings = saved_settings; m_cost = saved_cost;
                if (!ok) refuse("Derived::step_size_selection_impl is overridden and is not the LSFilter line search of the reference's "
                                "valet_parking_mpc_test.cpp (the only line-search override with a device twin)");
                else if (flt->max_depth < 1 || flt->max_depth > PMB_FILTER_CAP) refuse("filter.max_depth must be in 1..16 (PMB_FILTER_CAP)");
                else if ((int)flt->m_filter.size() > PMB_FILTER_CAP) refuse("the filter holds more than PMB_FILTER_CAP entries");
                else eo.filter_line_search = true;
            }
        }
        if (overrides_constraints_violation_impl<Derived>::value) refuse("Derived::constraints_violation_impl is overridden");
        if (overrides_max_constraints_violation_impl<Derived>::value) refuse("Derived::max_constraints_violation_impl is overridden");
        if (overrides_termination_criteria_impl<Derived>::value) refuse("Derived::termination_criteria_impl is overridden");
        if (overrides_linearisation_dense_impl<Derived>::value || overrides_linearisation_sparse_impl<Derived>::value)
            refuse("Derived::linearisation_*_impl is overridden (only the exact AD linearisation exists on the device)");
        eo.preconditioner = Preconditioner::pmb_engine_preconditioner;     // IdentityPreconditioner or RuizEquilibration<DENSE | SPARSE>
        if (!QPSolver::pmb_engine_inner_solver) refuse("the QP solver type is neither boxADMM<> nor ADMM<> (no device twin)");
        if (m_settings.iteration_callback != nullptr) refuse("settings().iteration_callback is set: a host callback cannot fire inside the fused device loop");

        // test data: a symmetric matrix whose Gershgorin discs partly reach into the negative half plane, so that a Gershgorin
        // regulariser has something to do on some columns and nothing on others.  The diagonal is integer valued on purpose:
        // the reference's own regulariser (minimal_time_test.cpp:98, 113) calls an unqualified `abs(aii)`, which is the C
        // library's int abs(int) on some tool chains — on integers both readings agree.
        nlp_hessian_t Hp;
        for (int j = 0; j < VAR_SIZE; ++j)
            for (int i = 0; i < VAR_SIZE; ++i)
                Hp(i, j) = (i == j) ? ((i % 4 == 0) ? scalar_t(1000) : ((i % 4 == 1) ? scalar_t(-2) : ((i % 4 == 2) ? scalar_t(0) : scalar_t(1))))
                                    : scalar_t(0.5) / scalar_t(1 + ((i + j) % 7));
        nlp_variable_t xs, cg; xs.setConstant(scalar_t(0.1)); cg.setZero();
        nlp_dual_t lam; lam.setZero();
        nlp_constraints_t b; b.setZero();
        std::vector<scalar_t> Abuf((size_t)(NUM_EQ + NUM_INEQ) * VAR_SIZE, scalar_t(0));
        nlp_jacobian_t& A = *reinterpret_cast<nlp_jacobian_t*>(Abuf.data());

        // (1) what does the update of the linearisation do?
        P rec;
        pmb::compat::active_probe() = &rec;
        m_H = Hp; m_lag_gradient.setZero();
        if (Problem::is_sparse) derived().update_linearisation_sparse_impl(xs, m_p, xs, lam, cg, m_H, A, b);
        else derived().update_linearisation_dense_impl(xs, m_p, xs, lam, cg, m_H, A, b);
        pmb::compat::active_probe() = nullptr;
        nlp_hessian_t Hafter = m_H;
        bool regularised_in_update = false;
        if (rec.is({P::LAG_GRAD, P::HESS_UPDATE_DEFAULT})) eo.block_bfgs = false;
        else if (rec.is({P::LAG_GRAD, P::HESS_UPDATE_OCP})) eo.block_bfgs = Problem::is_sparse;     // DENSE problem: plain BFGS (continuous_ocp.hpp:681-686)
        else if (rec.is({P::LAG_GRAD_HESS})) { eo.exact_hessian_every_iteration = true; regularised_in_update = true; }
        else refuse("Derived::update_linearisation_*_impl / hessian_update_impl does something the engine has no device twin for "
                    "(expected: gradient + BFGS, gradient + problem.hessian_update_impl, or the exact linearisation)");
        if (!regularised_in_update && !same(Hafter, Hp, 0)) refuse("Derived::hessian_update_impl modifies the Hessian on its own");

        // (2) what does the regularisation hook do?  (called by linearisation_*_impl after every exact Hessian)
        m_H = Hp;
        if (Problem::is_sparse) derived().hessian_regularisation_sparse_impl(m_H);
        else derived().hessian_regularisation_dense_impl(m_H);
        nlp_hessian_t Hg = Hp;
        host_gershgorin(Hg);
        if (same(m_H, Hp, 0)) eo.gershgorin_regularisation = false;
        else if (same(m_H, Hg, 1e-12)) eo.gershgorin_regularisation = true;
        else refuse("Derived::hessian_regularisation_*_impl is neither the default (none) nor the Gershgorin shift");
        if (regularised_in_update && !eo.gershgorin_regularisation && !same(Hafter, Hp, 0)) refuse("the exact linearisation override also modifies the Hessian");
    }

    pmb_sqp_t* m_handle = nullptr;
    double m_stats[4] = {0, 0, 0, 0};
    static void check(int rc, const char* what)
    { if (rc != PMB_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + pmb_last_error()); }
    void solve_impl()
    {
        probe_hooks();
        if (!m_engine_options.refused.empty()) {
            std::cerr << "polympc_b200: SQPBase::solve() REFUSED — " << m_engine_options.refused
                      << ".  The fused sm_100a SQP loop runs the reference's default hooks plus a fixed menu (block BFGS, exact "
                         "Hessian at every iteration, Gershgorin regularisation, Ruiz equilibration, the LSFilter line search); it will "
                         "not silently run a different algorithm.\n";
            m_info.iter = 0; m_info.qp_solver_iter = 0; m_info.status.value = sqp_status_t::INVALID_SETTINGS;
            return;
        }
#if defined(__CUDACC__) || defined(PMB_EMU)
        if (!m_handle) {
            m_handle = pmb_sqp_create(pmb::compat::problem_name<Problem>(), 1, 0);
            if (!m_handle) throw std::runtime_error(std::string("pmb_sqp_create: ") + pmb_last_error());
        }
#else
        static_assert(sizeof(Problem) == 0, "compile this translation unit with nvcc: the kernels of the problem class are instantiated here");
#endif
        pmb_ocp_t* ocp = pmb_sqp_problem(m_handle);
        // the problem object (Q, R, ... members) travels as a blob; then the horizon
        std::vector<double> blob((sizeof(Problem) + 7) / 8, 0.0);
        std::memcpy(blob.data(), (const void*)&problem, sizeof(Problem));
        check(pmb_ocp_set_params(ocp, blob.data(), (int)blob.size()), "set_params");
        check(pmb_ocp_set_time_limits(ocp, problem.t_start, problem.t_stop), "set_time_limits");
        pmb_sqp_settings_t st; pmb_sqp_default_settings(&st);
        st.tau = m_settings.tau; st.eta = m_settings.eta; st.rho = m_settings.rho; st.eps_prim = m_settings.eps_prim; st.eps_dual = m_settings.eps_dual;
        st.max_iter = m_settings.max_iter; st.line_search_max_iter = m_settings.line_search_max_iter;
        check(pmb_sqp_set_settings(m_handle, &st), "set_settings");
        pmb_qp_settings_t q; pmb_sqp_default_qp_settings(&q);
        const auto& s = m_qp_solver.m_settings;
        q.rho = s.rho; q.sigma = s.sigma; q.alpha = s.alpha; q.eps_rel = s.eps_rel; q.eps_abs = s.eps_abs; q.max_iter = s.max_iter;
        q.check_termination = s.check_termination; q.warm_start = s.warm_start; q.adaptive_rho = s.adaptive_rho;
        q.adaptive_rho_tolerance = s.adaptive_rho_tolerance; q.adaptive_rho_interval = s.adaptive_rho_interval;
        q.reuse_pattern = s.reuse_pattern; q.verbose = s.verbose;
        check(pmb_sqp_set_qp_settings(m_handle, &q), "set_qp_settings");
        check(pmb_sqp_set_hessian_options(m_handle, m_engine_options.exact_hessian_every_iteration, m_engine_options.gershgorin_regularisation),
              "set_hessian_options");
        check(pmb_sqp_set_hessian_update(m_handle, m_engine_options.block_bfgs ? PMB_HESSIAN_BFGS_BLOCK : PMB_HESSIAN_BFGS_DENSE), "set_hessian_update");
        check(pmb_sqp_set_qp_solver(m_handle, QPSolver::pmb_engine_qp_solver), "set_qp_solver");
        check(pmb_sqp_set_preconditioner(m_handle, m_engine_options.preconditioner), "set_preconditioner");
        LSFilter<scalar_t>* flt = m_engine_options.filter_line_search ? filter_of(derived(), 0) : nullptr;
        double fstate[PMB_FILTER_DOUBLES];
        if (flt) {                                          // the solver's filter travels to the device and back
            check(pmb_sqp_set_line_search(m_handle, PMB_LS_FILTER, flt->beta, flt->max_depth), "set_line_search");
            for (double& v : fstate) v = 0.0;
            int k = 0;
            for (const auto& e : flt->m_filter) { fstate[1 + k] = e.first; fstate[1 + PMB_FILTER_CAP + k] = e.second; ++k; }
            fstate[0] = k;
            check(pmb_sqp_set_filter(m_handle, fstate, PMB_FILTER_DOUBLES), "set_filter");
        } else {
            check(pmb_sqp_set_line_search(m_handle, PMB_LS_L1_MERIT, 0.0, 10), "set_line_search");
        }
        check(pmb_sqp_set_bounds_x(m_handle, m_lbx.data(), m_ubx.data(), VAR_SIZE), "set_bounds_x");
        if (NUM_INEQ > 0) check(pmb_sqp_set_bounds_g(m_handle, m_lbg.data(), m_ubg.data(), NUM_INEQ), "set_bounds_g");
        if (Problem::ND > 0) check(pmb_sqp_set_parameters(m_handle, m_p.data(), Problem::ND), "set_parameters");
        check(pmb_sqp_set_primal(m_handle, m_x.data(), VAR_SIZE), "set_primal");          // warm start from the kept iterate,
        check(pmb_sqp_set_dual(m_handle, m_lam.data(), NUM_CONSTR), "set_dual");          // like a second reference solve()
        check(pmb_sqp_solve(m_handle), "solve");
        check(pmb_sqp_get_primal(m_handle, m_x.data()), "get_primal");
        check(pmb_sqp_get_dual(m_handle, m_lam.data()), "get_dual");
        pmb_sqp_info_t inf;
        check(pmb_sqp_get_info(m_handle, &inf), "get_info");
        m_info.iter = inf.iter; m_info.qp_solver_iter = inf.qp_solver_iter;
        m_info.status.value = inf.status == PMB_SQP_SOLVED ? sqp_status_t::SOLVED
                            : (inf.status == PMB_SQP_MAX_ITER_EXCEEDED ? sqp_status_t::MAX_ITER_EXCEEDED : sqp_status_t::INVALID_SETTINGS);
        check(pmb_sqp_get_stats(m_handle, m_stats), "get_stats");
        m_cost = m_stats[0];
        if (flt) {
            check(pmb_sqp_get_filter(m_handle, fstate), "get_filter");
            flt->m_filter.clear();
            for (int k = 0; k < (int)fstate[0]; ++k) flt->m_filter.emplace_back(fstate[1 + k], fstate[1 + PMB_FILTER_CAP + k]);
        }
        if (std::getenv("POLYMPC_B200_REPORT")) {     // one line per solve for harnesses that cannot change the calling code
            bool finite = true;
            for (int i = 0; i < VAR_SIZE; ++i) finite = finite && std::isfinite(m_x(i));
            for (int i = 0; i < NUM_CONSTR; ++i) finite = finite && std::isfinite(m_lam(i));
            std::cerr << "polympc_b200: solve status=" << (int)m_info.status.value << " iter=" << m_info.iter << " qp_iter=" << m_info.qp_solver_iter
                      << " finite=" << (finite ? 1 : 0) << " block_bfgs=" << (m_engine_options.block_bfgs ? 1 : 0)
                      << " exact_hessian=" << (m_engine_options.exact_hessian_every_iteration ? 1 : 0)
                      << " gershgorin=" << (m_engine_options.gershgorin_regularisation ? 1 : 0)
                      << " qp_solver=" << (int)QPSolver::pmb_engine_qp_solver << " preconditioner=" << m_engine_options.preconditioner << " filter_ls=" << (m_engine_options.filter_line_search ? 1 : 0) << "\n";
        }
    }
};

// ---- MPC facade (mpc_wrapper.hpp:17-298) -----------------------------------------------------------------------------------
template <typename OCP, template <typename, typename...> class Solver, typename... Args>
class MPC {
private:
    using nlp_solver_t = Solver<OCP, Args...>;
    nlp_solver_t m_solver;
public:
    static constexpr int nx = OCP::NX, nu = OCP::NU, np = OCP::NP, nd = OCP::ND, ng = OCP::NG;
    static constexpr int var_size = OCP::VAR_SIZE, varx_size = OCP::VARX_SIZE, varu_size = OCP::VARU_SIZE, dual_size = OCP::DUAL_SIZE;
    static constexpr int num_nodes = OCP::NUM_NODES, num_segms = OCP::NUM_SEGMENTS, num_ineq = OCP::NUM_INEQ, poly_order = OCP::POLY_ORDER;

    using scalar_t = typename OCP::scalar_t;
    using state_t = Eigen::Matrix<scalar_t, nx, 1>;
    using control_t = Eigen::Matrix<scalar_t, nu, 1>;
    using parameter_t = Eigen::Matrix<scalar_t, np, 1>;
    using static_param = Eigen::Matrix<scalar_t, nd, 1>;
    using constraint_t = Eigen::Matrix<scalar_t, ng, 1>;
    using traj_state_t = Eigen::Matrix<scalar_t, varx_size, 1>;
    using traj_control_t = Eigen::Matrix<scalar_t, varu_size, 1>;
    using dual_var_t = Eigen::Matrix<scalar_t, dual_size, 1>;
    using constraints_t = Eigen::Matrix<scalar_t, num_ineq, 1>;

    MPC() = default;

    void set_time_limits(const scalar_t& t0, const scalar_t& tf) noexcept { m_solver.get_problem().set_time_limits(t0, tf); }
    void initial_conditions(const Eigen::Ref<const state_t>& x0) noexcept { initial_conditions(x0, x0); }
    void initial_conditions(const Eigen::Ref<const state_t>& x0_lb, const Eigen::Ref<const state_t>& x0_ub) noexcept
    { put(m_solver.upper_bound_x(), varx_size - nx, x0_ub); put(m_solver.lower_bound_x(), varx_size - nx, x0_lb); }
    void x_lower_bound(const Eigen::Ref<const state_t>& xlb) noexcept { for (int k = 0; k < num_nodes - 1; ++k) put(m_solver.lower_bound_x(), k * nx, xlb); }
    void x_upper_bound(const Eigen::Ref<const state_t>& xub) noexcept { for (int k = 0; k < num_nodes - 1; ++k) put(m_solver.upper_bound_x(), k * nx, xub); }
    void state_bounds(const Eigen::Ref<const state_t>& xlb, const Eigen::Ref<const state_t>& xub) noexcept { x_lower_bound(xlb); x_upper_bound(xub); }
    void state_trajectory_bounds(const Eigen::Ref<const traj_state_t>& xlb, const Eigen::Ref<const traj_state_t>& xub) noexcept
    { put(m_solver.lower_bound_x(), 0, xlb); put(m_solver.upper_bound_x(), 0, xub); }
    void x_final_lower_bound(const Eigen::Ref<const state_t>& xlb) noexcept { put(m_solver.lower_bound_x(), 0, xlb); }
    void x_final_upper_bound(const Eigen::Ref<const state_t>& xub) noexcept { put(m_solver.upper_bound_x(), 0, xub); }
    void final_state_bounds(const Eigen::Ref<const state_t>& xlb, const Eigen::Ref<const state_t>& xub) noexcept { x_final_lower_bound(xlb); x_final_upper_bound(xub); }
    void u_lower_bound(const Eigen::Ref<const control_t>& lb) noexcept { for (int k = 0; k < num_nodes; ++k) put(m_solver.lower_bound_x(), varx_size + k * nu, lb); }
    void u_upper_bound(const Eigen::Ref<const control_t>& ub) noexcept { for (int k = 0; k < num_nodes; ++k) put(m_solver.upper_bound_x(), varx_size + k * nu, ub); }
    void control_trajecotry_bounds(const Eigen::Ref<const traj_control_t>& lb, const Eigen::Ref<const traj_control_t>& ub) noexcept
    { put(m_solver.lower_bound_x(), varx_size, lb); put(m_solver.upper_bound_x(), varx_size, ub); }
    void control_bounds(const Eigen::Ref<const control_t>& lb, const Eigen::Ref<const control_t>& ub) noexcept { u_lower_bound(lb); u_upper_bound(ub); }
    void constraints_trajectory_bounds(const Eigen::Ref<const constraints_t>& lbg, const Eigen::Ref<const constraints_t>& ubg) noexcept
    { put(m_solver.lower_bound_g(), 0, lbg); put(m_solver.upper_bound_g(), 0, ubg); }
    void constraints_bounds(const Eigen::Ref<const constraint_t>& lbg, const Eigen::Ref<const constraint_t>& ubg) noexcept
    { for (int k = 0; k < num_nodes; ++k) { put(m_solver.lower_bound_g(), k * ng, lbg); put(m_solver.upper_bound_g(), k * ng, ubg); } }
    void set_static_parameters(const Eigen::Ref<const static_param>& param) noexcept { put(m_solver.parameters(), 0, param); }
    /** optimised parameters (mpc_wrapper.hpp:137-160): box bounds and guess of the tail of the NLP variable */
    void parameters_bounds(const Eigen::Ref<const parameter_t>& lbp, const Eigen::Ref<const parameter_t>& ubp) noexcept
    { put(m_solver.lower_bound_x(), varx_size + varu_size, lbp); put(m_solver.upper_bound_x(), varx_size + varu_size, ubp); }
    void p_guess(const Eigen::Ref<const parameter_t>& g) noexcept { put(m_solver.primal_solution(), varx_size + varu_size, g); }
    void x_guess(const Eigen::VecDyn<scalar_t>& g) noexcept { for (int i = 0; i < varx_size; ++i) m_solver.primal_solution()(i) = g.v[i]; }
    void u_guess(const Eigen::VecDyn<scalar_t>& g) noexcept { for (int i = 0; i < varu_size; ++i) m_solver.primal_solution()(varx_size + i) = g.v[i]; }
    void x_guess(const Eigen::Ref<const traj_state_t>& g) noexcept { put(m_solver.primal_solution(), 0, g); }
    void u_guess(const Eigen::Ref<const traj_control_t>& g) noexcept { put(m_solver.primal_solution(), varx_size, g); }
    void lam_guess(const Eigen::Ref<const dual_var_t>& g) noexcept { put(m_solver.dual_solution(), 0, g); }

    const typename nlp_solver_t::nlp_settings_t& settings() const noexcept { return m_solver.m_settings; }
    typename nlp_solver_t::nlp_settings_t& settings() noexcept { return m_solver.m_settings; }
    const typename nlp_solver_t::qp_solver_t::settings_t& qp_settings() const noexcept { return m_solver.m_qp_solver.m_settings; }
    typename nlp_solver_t::qp_solver_t::settings_t& qp_settings() noexcept { return m_solver.m_qp_solver.m_settings; }
    const typename nlp_solver_t::nlp_info_t& info() const noexcept { return m_solver.m_info; }
    typename nlp_solver_t::nlp_info_t& info() noexcept { return m_solver.m_info; }
    const nlp_solver_t& solver() const noexcept { return m_solver; }
    nlp_solver_t& solver() noexcept { return m_solver; }
    const OCP& ocp() const noexcept { return m_solver.get_problem(); }
    OCP& ocp() noexcept { return m_solver.get_problem(); }
    scalar_t primal_norm() const noexcept { return m_solver.primal_norm(); }
    scalar_t dual_norm() const noexcept { return m_solver.dual_norm(); }
    scalar_t constr_violation() const noexcept { return m_solver.constr_violation(); }
    scalar_t cost() const noexcept { return m_solver.cost(); }

    traj_state_t solution_x() const noexcept { return get<varx_size>(m_solver.primal_solution(), 0); }
    Eigen::Matrix<scalar_t, nx, num_nodes> solution_x_reshaped() const noexcept
    { Eigen::Matrix<scalar_t, nx, num_nodes> r; for (int i = 0; i < varx_size; ++i) r.m[i] = m_solver.primal_solution()(i); return r; }
    /** k-th node counted from the initial time (the NLP variable stores the final time first), mpc_wrapper.hpp:241-243 */
    state_t solution_x_at(const int& k) const noexcept { return get<nx>(m_solver.primal_solution(), varx_size - (k + 1) * nx); }
    state_t solution_x_at(const scalar_t& t) const noexcept { return interpolate<nx>(t, varx_size); }
    traj_control_t solution_u() const noexcept { return get<varu_size>(m_solver.primal_solution(), varx_size); }
    Eigen::Matrix<scalar_t, nu, num_nodes> solution_u_reshaped() const noexcept
    { Eigen::Matrix<scalar_t, nu, num_nodes> r; for (int i = 0; i < varu_size; ++i) r.m[i] = m_solver.primal_solution()(varx_size + i); return r; }
    control_t solution_u_at(const int& k) const noexcept { return get<nu>(m_solver.primal_solution(), varx_size + varu_size - (k + 1) * nu); }
    control_t solution_u_at(const scalar_t& t) const noexcept { return interpolate<nu>(t, varx_size + varu_size); }
    parameter_t solution_p() const noexcept { return get<np>(m_solver.primal_solution(), varx_size + varu_size); }
    dual_var_t solution_dual() const noexcept { return m_solver.dual_solution(); }

    void solve() { m_solver.solve(); }   // not noexcept: an engine failure (CUDA error, missing device) throws instead of std::terminate

private:
    template <class Dst, class Src> static void put(Dst& dst, int off, const Src& src) { for (int i = 0; i < (int)Src::Size; ++i) dst(off + i) = src(i); }
    template <int LEN, class Src> static Eigen::Matrix<scalar_t, LEN, 1> get(const Src& src, int off)
    { Eigen::Matrix<scalar_t, LEN, 1> r; for (int i = 0; i < LEN; ++i) r.m[i] = src(off + i); return r; }
    /** Lagrange interpolation on the Chebyshev nodes of the segment containing t (t relative to t_start), as
     *  polympc::LagrangeSpline::eval does (src/polynomials/splines.hpp:101-139, mpc_wrapper.hpp:245-281) */
    template <int W> Eigen::Matrix<scalar_t, W, 1> interpolate(const scalar_t& t, int block_end) const
    {
        const OCP& o = m_solver.get_problem();
        const scalar_t seg = (o.t_stop - o.t_start) / num_segms;
        const auto nodes = polympc::Chebyshev<poly_order>::compute_nodes();     // descending: cos(k pi / P)
        int idx = (int)std::floor(t / seg);
        idx = idx < 0 ? 0 : (idx > num_segms - 1 ? num_segms - 1 : idx);
        scalar_t tg[poly_order + 1];                                            // ascending time grid of the segment
        for (int a = 0; a <= poly_order; ++a) tg[a] = (seg / 2) * nodes(poly_order - a) + (o.t_start + seg / 2 + idx * seg);
        const scalar_t tq = o.t_start + t;
        Eigen::Matrix<scalar_t, W, 1> out; out.setZero();
        for (int a = 0; a <= poly_order; ++a) {
            scalar_t l = 1.0;
            for (int c = 0; c <= poly_order; ++c) if (c != a) l *= (tq - tg[c]) / (tg[a] - tg[c]);
            const int ka = idx * poly_order + a;
            for (int i = 0; i < W; ++i) out(i) += l * m_solver.primal_solution()(block_end - (ka + 1) * W + i);
        }
        return out;
    }
};

// ---- functor annotation for the code that follows (see the header comment) ----------------------------------------------
#if defined(__CUDACC__) && !defined(POLYMPC_B200_NO_INLINE_HD)
#define inline __host__ __device__ inline
#endif
