// tests/cpp/test_problem_concept.cpp — the source-compatibility layer seen from a USER translation unit (TEST).
//
// A reference-style problem class on a grid the library does not ship (robot, 5 x 2: the grid of the reference's CasADi
// fixtures) is compiled HERE — by nvcc for the GPU suite, by g++ against the warp emulator for the CPU suite — so its kernels
// are instantiated in this file and reach the engine through pmb_register_problem.  Checked, all bit for bit:
//   1. the Problem concept of SQPBase (ContinuousOCP::cost / cost_gradient / cost_gradient_hessian / equalities /
//      equalities_linearised / lagrangian_gradient / lagrangian_gradient_hessian, reference continuous_ocp.hpp:430-647)
//      against the library's built-in "mobile_robot_5x2", which tests/test_oracle_golden.py and tests/test_gpu_parity.py pin
//      to the reference's CasADi fixtures;
//   2. a single-instance SQPBase solve against instance 0 of a batched solve of the same registered class
//      (polympc::b200::BatchedMPC with pmb::compat::problem_name<>()), and against the built-in twin.
#define DROPIN_ROBOT_SEGMENTS 2
#include "../dropin/robot_ocp.hpp"
#include "solvers/sqp_base.hpp"
#include "control/mpc_wrapper.hpp"
#undef inline                      // the functor annotation is only wanted for the problem class above
#include "../../include/polympc_b200.hpp"

#include <cstdio>
#include <cstring>

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { ++g_fail; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)

template <typename Problem, typename QPSolver = boxADMM<Problem::VAR_SIZE, Problem::NUM_EQ + Problem::NUM_INEQ, typename Problem::scalar_t>>
class Solver : public SQPBase<Solver<Problem, QPSolver>, Problem, QPSolver> {};

static bool same(const double* a, const double* b, size_t n) { return std::memcmp(a, b, n * sizeof(double)) == 0; }

int main()
{
    using OCP = dropin::RobotOCP;
    constexpr int N = OCP::VAR_SIZE, M = OCP::NUM_EQ + OCP::NUM_INEQ, DUAL = OCP::DUAL_SIZE;
    static_assert(N == 55 && M == 33 && DUAL == 88, "5 x 2 robot grid (SURVEY.md §8 table)");
    const int before = pmb_problem_count();
    const char* name = pmb::compat::problem_name<OCP>();
    EXPECT(pmb_problem_count() == before + 1);
    EXPECT(pmb_register_problem(name, nullptr) != PMB_OK);                       // null factory
    std::printf("registered '%s' (%s)\n", name, pmb_version());

    // ---- 1. Problem concept vs the built-in twin
    OCP ocp;
    ocp.set_time_limits(0, 1);
    OCP::nlp_variable_t var; OCP::nlp_dual_t lam; OCP::static_parameter_t p; p(0) = 1.0;
    unsigned long long z = 88172645463325252ull;
    auto rnd = [&] { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return (double)(z >> 11) / 9007199254740992.0 * 2.0 - 1.0; };
    for (int i = 0; i < N; ++i) var(i) = rnd();
    for (int i = 0; i < DUAL; ++i) lam(i) = 2.0 * rnd();
    pmb_ocp_t* twin = pmb_ocp_create("mobile_robot_5x2", 0);
    EXPECT(twin != nullptr);
    if (!twin) { std::printf("%s\n", pmb_last_error()); return 1; }
    pmb_ocp_set_time_limits(twin, 0, 1);
    {
        double c0, c1; ocp.cost(var, p, c0); pmb_ocp_cost(twin, 1, var.data(), p.data(), &c1); EXPECT(same(&c0, &c1, 1));
        OCP::nlp_eq_constraints_t e0, e1; OCP::nlp_eq_jacobian_t j0, j1;
        ocp.equalities(var, p, e0); pmb_ocp_equalities(twin, 1, var.data(), p.data(), e1.data()); EXPECT(same(e0.data(), e1.data(), M));
        ocp.equalities_linearised(var, p, e0, j0); pmb_ocp_equalities_linearised(twin, 1, var.data(), p.data(), e1.data(), j1.data());
        EXPECT(same(e0.data(), e1.data(), M) && same(j0.data(), j1.data(), (size_t)M * N));
        OCP::nlp_variable_t g0, g1, lg0, lg1; static OCP::nlp_hessian_t H0, H1;
        ocp.cost_gradient(var, p, c0, g0); pmb_ocp_cost_gradient(twin, 1, var.data(), p.data(), &c1, g1.data());
        EXPECT(same(&c0, &c1, 1) && same(g0.data(), g1.data(), N));
        ocp.cost_gradient_hessian(var, p, c0, g0, H0); pmb_ocp_cost_gradient_hessian(twin, 1, var.data(), p.data(), &c1, g1.data(), H1.data());
        EXPECT(same(g0.data(), g1.data(), N) && same(H0.data(), H1.data(), (size_t)N * N));
        OCP::nlp_constraints_t cc0, cc1; static OCP::nlp_jacobian_t J0, J1;
        ocp.lagrangian_gradient(var, p, lam, c0, lg0, g0, cc0, J0);
        pmb_ocp_lagrangian_gradient(twin, 1, var.data(), p.data(), lam.data(), &c1, lg1.data(), g1.data(), cc1.data(), J1.data());
        EXPECT(same(&c0, &c1, 1) && same(lg0.data(), lg1.data(), N) && same(g0.data(), g1.data(), N) && same(cc0.data(), cc1.data(), M) &&
               same(J0.data(), J1.data(), (size_t)M * N));
        ocp.lagrangian_gradient_hessian(var, p, lam, c0, lg0, H0, g0, cc0, J0);
        pmb_ocp_lagrangian_gradient_hessian(twin, 1, var.data(), p.data(), lam.data(), &c1, lg1.data(), H1.data(), g1.data(), cc1.data(), J1.data());
        EXPECT(same(lg0.data(), lg1.data(), N) && same(H0.data(), H1.data(), (size_t)N * N) && same(J0.data(), J1.data(), (size_t)M * N));
        // data members of the class reach the kernels: Q = 2 I doubles the state part of the cost Hessian
        ocp.set_Q_coeff(2.0);
        double c2; ocp.cost(var, p, c2); EXPECT(c2 > c0 || c2 < c0);
        const double tw[8] = {2, 2, 2, 1, 1, 1, 1, 1};
        pmb_ocp_set_params(twin, tw, 8); pmb_ocp_cost(twin, 1, var.data(), p.data(), &c1); EXPECT(same(&c2, &c1, 1));
    }
    pmb_ocp_destroy(twin);

    // ---- 2. SQPBase (one instance) == BatchedMPC over the registered class == BatchedMPC over the built-in twin
    const int batch = 3;
    MPC<OCP, Solver> mpc;
    mpc.settings().max_iter = 10; mpc.settings().line_search_max_iter = 10;
    mpc.set_time_limits(0, 2);
    MPC<OCP, Solver>::static_param d; d << 2.0;
    MPC<OCP, Solver>::control_t lbu, ubu; lbu << -1.5, -0.75; ubu << 1.5, 0.75;
    MPC<OCP, Solver>::state_t x0; x0 << 0.5, 0.5, 0.5;
    mpc.set_static_parameters(d); mpc.control_bounds(lbu, ubu); mpc.initial_conditions(x0);
    mpc.solve();
    EXPECT(mpc.info().status.value == sqp_status_t::SOLVED);
    const char* names[2] = {name, "mobile_robot_5x2"};
    for (int k = 0; k < 2; ++k) {
        polympc::b200::BatchedMPC bm(names[k], batch);
        pmb_sqp_settings_t st = bm.settings(); st.max_iter = 10; st.line_search_max_iter = 10; bm.settings(st);
        bm.set_time_limits(0, 2); bm.set_static_parameters({2.0}); bm.control_bounds({-1.5, -0.75}, {1.5, 0.75});
        polympc::b200::vec x0b; for (int b = 0; b < batch; ++b) { x0b.push_back(0.5 - 0.1 * b); x0b.push_back(0.5); x0b.push_back(0.5); }
        bm.initial_conditions(x0b);
        bm.solve_async(); bm.wait();
        const polympc::b200::vec xs = bm.solution_x(0), us = bm.solution_u(0);
        EXPECT(bm.info(0).iter == mpc.info().iter && bm.info(0).status == PMB_SQP_SOLVED);
        EXPECT(same(xs.data(), mpc.solution_x().data(), xs.size()) && same(us.data(), mpc.solution_u().data(), us.size()));
        for (int b = 0; b < batch; ++b) EXPECT(bm.info(b).status == PMB_SQP_SOLVED);
    }
    std::printf("%d failures\n", g_fail);
    return g_fail == 0 ? 0 : 1;
}
