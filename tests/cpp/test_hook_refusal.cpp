// tests/cpp/test_hook_refusal.cpp — TEST INFRASTRUCTURE.  The source-compatibility layer must never run an algorithm the
// user did not write: SQPBase::solve() probes the CRTP hooks of the Derived solver (include/polympc_compat/polympc_compat.hpp)
// and either maps them onto the engine's menu or refuses loudly (stderr + status INVALID_SETTINGS).  Each case below is a
// solver a PolyMPC user could write against sqp_base.hpp:198-350; none of the refused ones ever reaches the device.
#include "../dropin/robot_ocp.hpp"
#include "solvers/sqp_base.hpp"
#include "solvers/box_admm.hpp"
#include "solvers/admm.hpp"
#include "solvers/qp_preconditioners.hpp"
#include "control/mpc_wrapper.hpp"

#include <cstdio>

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { ++g_fail; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)

using OCP = dropin::RobotOCP;
using BoxQP = boxADMM<OCP::VAR_SIZE, OCP::NUM_EQ + OCP::NUM_INEQ, double>;
#define SOLVER_TYPES(S, Q)                                   \
    using Base = SQPBase<S<Problem, QPSolver>, Problem, Q>;   \
    using typename Base::scalar_t; using typename Base::nlp_variable_t; using typename Base::nlp_hessian_t;

// (a) all defaults: accepted, dense BFGS
template <typename Problem, typename QPSolver = BoxQP> class Plain : public SQPBase<Plain<Problem, QPSolver>, Problem, QPSolver> {};

// (b) hessian_update_impl forwarded to the problem: accepted (DENSE problem => dense BFGS, continuous_ocp.hpp:681-686)
template <typename Problem, typename QPSolver = BoxQP>
class Forwarding : public SQPBase<Forwarding<Problem, QPSolver>, Problem, QPSolver> {
public:
    SOLVER_TYPES(Forwarding, QPSolver)
    void hessian_update_impl(Eigen::Ref<nlp_hessian_t> H, const Eigen::Ref<const nlp_variable_t>& s, const Eigen::Ref<const nlp_variable_t>& y) noexcept
    { this->problem.hessian_update_impl(H, s, y); }
};

// (c) a home-made SR1 update: refused
template <typename Problem, typename QPSolver = BoxQP>
class Sr1 : public SQPBase<Sr1<Problem, QPSolver>, Problem, QPSolver> {
public:
    SOLVER_TYPES(Sr1, QPSolver)
    void hessian_update_impl(Eigen::Ref<nlp_hessian_t> H, const Eigen::Ref<const nlp_variable_t>& s, const Eigen::Ref<const nlp_variable_t>& y) noexcept
    { for (int i = 0; i < Problem::VAR_SIZE; ++i) H(i, i) += y(i) * y(i) / (1.0 + s(i) * s(i)); }
};

// (d) a custom line search: refused
template <typename Problem, typename QPSolver = BoxQP>
class FullStep : public SQPBase<FullStep<Problem, QPSolver>, Problem, QPSolver> {
public:
    SOLVER_TYPES(FullStep, QPSolver)
    scalar_t step_size_selection_impl(const Eigen::Ref<const nlp_variable_t>&) noexcept { return scalar_t(1); }
};

// (e) a regulariser that is not the Gershgorin shift: refused
template <typename Problem, typename QPSolver = BoxQP>
class ShiftAll : public SQPBase<ShiftAll<Problem, QPSolver>, Problem, QPSolver> {
public:
    SOLVER_TYPES(ShiftAll, QPSolver)
    void hessian_regularisation_dense_impl(Eigen::Ref<nlp_hessian_t> H) noexcept { for (int i = 0; i < Problem::VAR_SIZE; ++i) H(i, i) += 1.0; }
};

// (f) Ruiz preconditioner requested: accepted (pmb_sqp_set_preconditioner); (g) OSQP-style ADMM requested: accepted
// (pmb_sqp_set_qp_solver); (g') a QP solver type the engine does not know: refused
template <int N, int M> struct HomeMadeQP : boxADMM<N, M, double> { static constexpr bool pmb_engine_inner_solver = false; };
template <typename Problem, typename QPSolver = HomeMadeQP<OCP::VAR_SIZE, OCP::NUM_EQ + OCP::NUM_INEQ>> class WithOwnQP : public SQPBase<WithOwnQP<Problem, QPSolver>, Problem, QPSolver> {};
using Ruiz = polympc::RuizEquilibration<double, OCP::VAR_SIZE, OCP::NUM_EQ, DENSE>;
template <typename Problem, typename QPSolver = BoxQP> class WithRuiz : public SQPBase<WithRuiz<Problem, QPSolver>, Problem, QPSolver, Ruiz> {};

// (h) a filter line search written like the reference's tests/control/valet_parking_mpc_test.cpp:110-155: accepted;
// (i) the same with an extra sufficient-decrease test the engine does not have: refused
template <typename Problem, bool EXTRA_TEST, typename QPSolver = BoxQP>
class FilterSearch : public SQPBase<FilterSearch<Problem, EXTRA_TEST, QPSolver>, Problem, QPSolver> {
public:
    using Base = SQPBase<FilterSearch<Problem, EXTRA_TEST, QPSolver>, Problem, QPSolver>;
    using typename Base::scalar_t; using typename Base::nlp_variable_t; using typename Base::nlp_hessian_t;
    LSFilter<scalar_t> filter;
    scalar_t step_size_selection_impl(const Eigen::Ref<const nlp_variable_t>& p) noexcept
    {
        scalar_t v0 = this->constraints_violation(this->m_x), f0;
        this->problem.cost(this->m_x, this->m_p, f0);
        if (filter.is_acceptable(f0, v0)) filter.add(f0, v0);
        scalar_t alpha = 1;
        for (int i = 1; i < this->m_settings.line_search_max_iter; i++) {
            nlp_variable_t xs = alpha * p; xs += this->m_x;
            scalar_t f, v;
            this->problem.cost(xs, this->m_p, f);
            v = this->constraints_violation(xs);
            this->m_cost = f;
            if (filter.is_acceptable(f, v) && (!EXTRA_TEST || f < f0)) { filter.add(f, v); return alpha; }
            alpha *= this->m_settings.tau;
        }
        return alpha;
    }
};
using OsqpAdmm = ADMM<OCP::VAR_SIZE, OCP::NUM_EQ + OCP::NUM_INEQ, double>;
template <typename Problem, typename QPSolver = OsqpAdmm> class WithAdmm : public SQPBase<WithAdmm<Problem, QPSolver>, Problem, QPSolver> {};

static void on_iteration(void*) {}

/** the QPBase object concept stand-alone: the reference's tests/solvers/qp/admm_solver_test.cpp:16-45 (admmSimpleQP) and
 *  box_admm_test.cpp:15-45 through ADMM<2, 1> and boxADMM<2, 1> */
template <class QP> static void simple_qp(const char* what)
{
    Eigen::Matrix<double, 2, 2> H; Eigen::Matrix<double, 2, 1> h, xl, xu, solution; Eigen::Matrix<double, 1, 2> A; Eigen::Matrix<double, 1, 1> al, au;
    H << 4, 1, 1, 2; h << 1, 1; A << 1, 1; al << 1; au << 1; xl << 0, 0; xu << 0.7, 0.7; solution << 0.3, 0.7;
    QP prob;
    prob.settings().max_iter = 1000;
    prob.solve(H, h, A, al, au, xl, xu);
    std::printf("%-28s x = (%.6f, %.6f) status=%d iter=%d\n", what, prob.primal_solution()(0), prob.primal_solution()(1), (int)prob.info().status, prob.iter);
    EXPECT(prob.primal_solution().isApprox(solution, 1e-2));
    EXPECT(prob.iter < prob.settings().max_iter);
    EXPECT(prob.info().status == SOLVED);
}

template <class S> static int run(S& s, const char* what)
{
    s.settings().max_iter = 3; s.settings().line_search_max_iter = 3;
    s.get_problem().set_time_limits(0, 2);
    s.parameters()(0) = 2.0;
    s.solve();
    std::printf("%-28s status=%d refused='%s'\n", what, (int)s.info().status.value, s.engine_options().refused.c_str());
    return (int)s.info().status.value;
}

int main(int argc, char** argv)
{
    const bool have_engine = argc > 1;      // the accepted cases need the engine (emulator or GPU); the refused ones never touch it
    if (have_engine) {
        Plain<OCP> a; EXPECT(run(a, "defaults") != sqp_status_t::INVALID_SETTINGS); EXPECT(!a.engine_options().block_bfgs);
        Forwarding<OCP> b; EXPECT(run(b, "forward to problem (DENSE)") != sqp_status_t::INVALID_SETTINGS); EXPECT(!b.engine_options().block_bfgs);
        EXPECT(a.primal_solution().isApprox(b.primal_solution(), 0.0) || true);
        bool same = true; for (int i = 0; i < OCP::VAR_SIZE; ++i) same = same && a.primal_solution()(i) == b.primal_solution()(i);
        EXPECT(same);                         // DENSE problem: the forwarded update IS the default BFGS
    }
    { Sr1<OCP> s; EXPECT(run(s, "home-made SR1") == sqp_status_t::INVALID_SETTINGS); }
    { FullStep<OCP> s; EXPECT(run(s, "custom line search") == sqp_status_t::INVALID_SETTINGS); }
    { ShiftAll<OCP> s; EXPECT(run(s, "custom regulariser") == sqp_status_t::INVALID_SETTINGS); }
    { FilterSearch<OCP, true> s; EXPECT(run(s, "filter search + extra test") == sqp_status_t::INVALID_SETTINGS); }
    if (have_engine) {
        simple_qp<ADMM<2, 1, double>>("ADMM<2,1> stand-alone");
        simple_qp<boxADMM<2, 1, double>>("boxADMM<2,1> stand-alone");
        WithRuiz<OCP> r; EXPECT(run(r, "RuizEquilibration") != sqp_status_t::INVALID_SETTINGS); EXPECT(r.engine_options().preconditioner == 1);
        FilterSearch<OCP, false> f; f.filter.beta = 0.1; f.filter.add(1e9, 1e9);
        EXPECT(run(f, "filter line search") != sqp_status_t::INVALID_SETTINGS); EXPECT(f.engine_options().filter_line_search);
        EXPECT(f.filter.beta == 0.1 && f.filter.m_filter.size() >= 1 && f.filter.m_filter.back().first != 1e9);   // probed without a trace; synced back (the dominated seed left)
    }
    { WithOwnQP<OCP> s; EXPECT(run(s, "unknown QP solver type") == sqp_status_t::INVALID_SETTINGS); }
    if (have_engine) {
        WithAdmm<OCP> s; Plain<OCP> p;
        for (int k = 0; k < OCP::NX; ++k) {        // a non-trivial problem: the initial state pinned at (0.5, 0.5, 0.5)
            s.lower_bound_x()(OCP::VARX_SIZE - OCP::NX + k) = s.upper_bound_x()(OCP::VARX_SIZE - OCP::NX + k) = 0.5;
            p.lower_bound_x()(OCP::VARX_SIZE - OCP::NX + k) = p.upper_bound_x()(OCP::VARX_SIZE - OCP::NX + k) = 0.5;
        }
        EXPECT(run(s, "OSQP-style ADMM") != sqp_status_t::INVALID_SETTINGS); run(p, "boxADMM (again)");
        double d = 0; for (int i = 0; i < OCP::VAR_SIZE; ++i) d = std::fmax(d, std::fabs(s.primal_solution()(i) - p.primal_solution()(i)));
        std::printf("ADMM vs boxADMM after 3 SQP iterations: max |dx| = %.3e\n", d);
        EXPECT(d < 1e-2 && d > 0.0);              // a different QP solver: close, not identical
    }
    { Plain<OCP> s; s.settings().iteration_callback = &on_iteration; EXPECT(run(s, "iteration_callback") == sqp_status_t::INVALID_SETTINGS); }
    std::printf("%d failures\n", g_fail);
    return g_fail ? 1 : 0;
}
