// tests/cpp/test_facade.cpp — the reference's MPCWrapperTest (tests/control/mpc_wrapper_test.cpp:120-199) written against the
// batched C++ facade include/polympc_b200.hpp.  Built by tests/test_cpp_facade.py
//   * on the CPU suite against the warp-emulator build of the kernels (-include tests/warp_emu/emu_names.h), batch of 2;
//   * on the GPU suite against libpolympc_b200.so, batch of 64.
#include "polympc_b200.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)
static int failures = 0;

static bool approx(const polympc::b200::vec& a, const polympc::b200::vec& b, double prec)
{   // Eigen isApprox: ||a - b|| <= prec * min(||a||, ||b||)
    double d = 0, na = 0, nb = 0;
    for (size_t i = 0; i < a.size(); ++i) { d += (a[i] - b[i]) * (a[i] - b[i]); na += a[i] * a[i]; nb += b[i] * b[i]; }
    return std::sqrt(d) <= prec * std::sqrt(na < nb ? na : nb);
}

int main(int argc, char** argv)
{
    using namespace polympc::b200;
    const int batch = argc > 1 ? std::atoi(argv[1]) : 2;
    BatchedMPC mpc("mobile_robot_5x3", batch);
    {   // mpc.ocp().set_Q_coeff(2.0): Q = diag(2, 2, 2), R = I, QN = I
        vec data = {2, 2, 2, 1, 1, 1, 1, 1};
        mpc.set_problem_data(data);
    }
    pmb_sqp_settings_t st = mpc.settings();
    st.max_iter = 10; st.line_search_max_iter = 10;
    mpc.settings(st);
    mpc.set_time_limits(0, 2);
    mpc.set_static_parameters({2.0});
    mpc.control_bounds({-1.5, -0.75}, {1.5, 0.75});
    vec x0((size_t)batch * 3);
    for (int b = 0; b < batch; ++b) { x0[3 * b] = 0.5; x0[3 * b + 1] = 0.5; x0[3 * b + 2] = 0.5; }   // every instance = the reference test
    mpc.initial_conditions(x0);
    mpc.solve();
    std::vector<int> first(batch);
    for (int b = 0; b < batch; ++b) { first[b] = mpc.info(b).iter; EXPECT(mpc.info(b).status == PMB_SQP_SOLVED); }

    // warm started iteration
    for (int b = 0; b < batch; ++b) { x0[3 * b] = 0.3; x0[3 * b + 1] = 0.4; x0[3 * b + 2] = 0.5; }
    mpc.initial_conditions(x0, x0);
    mpc.solve();
    for (int b = 0; b < batch; ++b) {
        EXPECT(mpc.info(b).iter < first[b]);
        EXPECT(mpc.info(b).status == PMB_SQP_SOLVED);
        // initial condition is met at node 0 (counted from the initial time)
        const vec xs = mpc.solution_x_at(b, 0);
        EXPECT(std::fabs(xs[0] - x0[3 * b]) < 1e-3 && std::fabs(xs[1] - 0.4) < 1e-3 && std::fabs(xs[2] - 0.5) < 1e-3);
        // collocation points [0, 5, 10] vs polynomial interpolation at the corresponding instants
        EXPECT(approx(mpc.solution_x_at(b, 0), mpc.solution_x_at(b, 0.0), 1e-3));
        EXPECT(approx(mpc.solution_x_at(b, 5), mpc.solution_x_at(b, 0.666), 1e-3));
        EXPECT(approx(mpc.solution_x_at(b, 10), mpc.solution_x_at(b, 1.333), 1e-3));
        EXPECT(approx(mpc.solution_u_at(b, 0), mpc.solution_u_at(b, 0.0), 1e-3));
        EXPECT(approx(mpc.solution_u_at(b, 1), mpc.solution_u_at(b, 0.063), 1e-3));
        const vec u = mpc.solution_u(b);
        for (int k = 0; k < mpc.num_nodes(); ++k) EXPECT(std::fabs(u[2 * k]) <= 1.5 + 1e-3 && std::fabs(u[2 * k + 1]) <= 0.75 + 1e-3);
    }
    // replicated instances are solved independently and deterministically: bit-identical results
    for (int b = 1; b < batch; ++b) { EXPECT(mpc.solution_x(b) == mpc.solution_x(0)); EXPECT(mpc.solution_dual(b) == mpc.solution_dual(0)); }
    // exact interpolation at every node of the grid
    const vec tg = mpc.time_grid();
    for (int k = 0; k < mpc.num_nodes(); ++k) EXPECT(approx(mpc.solution_x_at(0, k), mpc.solution_x_at(0, tg[k]), 1e-9));
    // error behaviour: wrong sizes throw, unknown problems throw
    bool threw = false;
    try { mpc.initial_conditions(vec{1.0}); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { BatchedMPC bad("no_such_problem", 1); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
    std::printf("%s: first solve %d iterations, warm start %d iterations, %d failures\n", pmb_version(), first[0], mpc.info(0).iter, failures);
    return failures == 0 ? 0 : 1;
}
