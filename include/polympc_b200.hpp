// polympc_b200.hpp — C++ host facade over the C ABI (polympc_b200.h): the batched counterpart of PolyMPC's `MPC<OCP, Solver>`
// wrapper (reference src/control/mpc_wrapper.hpp:19-298).  Method names, argument meaning and the variable layout are the
// reference's; every setter that takes one state / control vector in the reference takes either one vector (broadcast to
// the whole batch) or a batch of vectors here, and every getter takes the instance index first.
//
//   reference (one instance)                               this header (batch of independent instances)
//   MPC<RobotOCP, Solver> mpc;                             polympc::b200::BatchedMPC mpc("mobile_robot_5x3", batch);
//   mpc.set_time_limits(0, 2);                             mpc.set_time_limits(0, 2);
//   mpc.set_static_parameters(p);                          mpc.set_static_parameters(p);
//   mpc.control_bounds(lbu, ubu);                          mpc.control_bounds(lbu, ubu);
//   mpc.initial_conditions(x0);                            mpc.initial_conditions(x0_batch);         // batch x NX
//   mpc.solve();                                           mpc.solve();
//   mpc.solution_u_at(0); mpc.info().iter                  mpc.solution_u_at(b, 0); mpc.info(b).iter
//
// Header-only, no Eigen needed (std::vector<double> in / out, column-major like Eigen).  Errors of the C layer are thrown
// as std::runtime_error carrying pmb_last_error(); per-instance solver outcomes are reported through info(b), like the
// reference reports them through sqp_info_t (src/solvers/sqp_base.hpp:49-61).
#pragma once
#include "polympc_b200.h"

#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace polympc {
namespace b200 {

using vec = std::vector<double>;

class BatchedMPC {
public:
    BatchedMPC(const std::string& problem, int batch, int device = 0) : m_batch(batch)
    {
        m_solver = pmb_sqp_create(problem.c_str(), batch, device);
        if (!m_solver) throw std::runtime_error(std::string("pmb_sqp_create: ") + pmb_last_error());
        check(pmb_ocp_dims(pmb_sqp_problem(m_solver), &m_d), "pmb_ocp_dims");
        const double inf = std::numeric_limits<double>::infinity();
        m_lbx.assign((size_t)batch * m_d.N, -inf); m_ubx.assign((size_t)batch * m_d.N, inf);          // sqp_base.hpp:75-78
        m_lbg.assign((size_t)batch * m_d.NG * m_d.NN, -inf); m_ubg.assign((size_t)batch * m_d.NG * m_d.NN, inf);
        m_x.assign((size_t)batch * m_d.N, 0.0); m_lam.assign((size_t)batch * m_d.DUAL, 0.0);           // sqp_base.hpp:80-81
        m_p.assign((size_t)batch * (m_d.ND > 0 ? m_d.ND : 1), 0.0);
        m_t0 = 0.0; m_tf = 1.0;
    }
    ~BatchedMPC() { pmb_sqp_destroy(m_solver); }
    BatchedMPC(const BatchedMPC&) = delete;
    BatchedMPC& operator=(const BatchedMPC&) = delete;

    // ---- sizes (mpc_wrapper.hpp:27-42)
    int batch() const { return m_batch; }
    int nx() const { return m_d.NX; }
    int nu() const { return m_d.NU; }
    int nd() const { return m_d.ND; }
    int var_size() const { return m_d.N; }
    int varx_size() const { return m_d.NX * m_d.NN; }
    int varu_size() const { return m_d.NU * m_d.NN; }
    int dual_size() const { return m_d.DUAL; }
    int num_nodes() const { return m_d.NN; }
    int num_segms() const { return m_d.S; }
    const pmb_dims_t& dims() const { return m_d; }

    // ---- problem data
    void set_time_limits(double t0, double tf) { m_t0 = t0; m_tf = tf; check(pmb_ocp_set_time_limits(pmb_sqp_problem(m_solver), t0, tf), "set_time_limits"); }
    /** data members of the problem class (Q, R, ...), flattened */
    void set_problem_data(const vec& v) { check(pmb_ocp_set_params(pmb_sqp_problem(m_solver), v.data(), (int)v.size()), "set_problem_data"); }
    void set_static_parameters(const vec& p) { spread(p, m_p, 0, m_d.ND, m_d.ND); }                          // :184-187

    // ---- bounds (mpc_wrapper.hpp:89-182); a vector of one instance is broadcast, batch x len sets every instance
    void initial_conditions(const vec& x0) { initial_conditions(x0, x0); }
    void initial_conditions(const vec& x0_lb, const vec& x0_ub)
    {
        spread(x0_lb, m_lbx, varx_size() - nx(), nx(), m_d.N);
        spread(x0_ub, m_ubx, varx_size() - nx(), nx(), m_d.N);
    }
    void x_lower_bound(const vec& xlb) { replicate(xlb, m_lbx, 0, nx(), num_nodes() - 1); }
    void x_upper_bound(const vec& xub) { replicate(xub, m_ubx, 0, nx(), num_nodes() - 1); }
    void state_bounds(const vec& xlb, const vec& xub) { x_lower_bound(xlb); x_upper_bound(xub); }
    void x_final_lower_bound(const vec& xlb) { spread(xlb, m_lbx, 0, nx(), m_d.N); }
    void x_final_upper_bound(const vec& xub) { spread(xub, m_ubx, 0, nx(), m_d.N); }
    void final_state_bounds(const vec& xlb, const vec& xub) { x_final_lower_bound(xlb); x_final_upper_bound(xub); }
    void u_lower_bound(const vec& lb) { replicate(lb, m_lbx, varx_size(), nu(), num_nodes()); }
    void u_upper_bound(const vec& ub) { replicate(ub, m_ubx, varx_size(), nu(), num_nodes()); }
    void control_bounds(const vec& lb, const vec& ub) { u_lower_bound(lb); u_upper_bound(ub); }
    void constraints_bounds(const vec& lbg, const vec& ubg)
    {
        if (m_d.NG == 0) return;
        for (int b = 0; b < m_batch; ++b)
            for (int k = 0; k < m_d.NN; ++k)
                for (int i = 0; i < m_d.NG; ++i) {
                    m_lbg[((size_t)b * m_d.NN + k) * m_d.NG + i] = lbg[i];
                    m_ubg[((size_t)b * m_d.NN + k) * m_d.NG + i] = ubg[i];
                }
    }

    // ---- initial guess (mpc_wrapper.hpp:190-205)
    void x_guess(const vec& xg) { spread(xg, m_x, 0, varx_size(), m_d.N); m_guess_dirty = true; }
    void u_guess(const vec& ug) { spread(ug, m_x, varx_size(), varu_size(), m_d.N); m_guess_dirty = true; }
    void lam_guess(const vec& lg) { spread(lg, m_lam, 0, m_d.DUAL, m_d.DUAL); m_guess_dirty = true; }

    // ---- settings (mpc_wrapper.hpp:207-212)
    pmb_sqp_settings_t settings() const { pmb_sqp_settings_t s; check(pmb_sqp_get_settings(m_solver, &s), "settings"); return s; }
    void settings(const pmb_sqp_settings_t& s) { check(pmb_sqp_set_settings(m_solver, &s), "settings"); }
    pmb_qp_settings_t qp_settings() const { pmb_qp_settings_t s; check(pmb_sqp_get_qp_settings(m_solver, &s), "qp_settings"); return s; }
    void qp_settings(const pmb_qp_settings_t& s) { check(pmb_sqp_set_qp_settings(m_solver, &s), "qp_settings"); }

    /** SQPBase::solve() for every instance of the batch (mpc_wrapper.hpp:298).  Like the reference, a second solve()
     *  warm-starts from the previous solution unless a new guess was given. */
    void solve() { upload(); check(pmb_sqp_solve(m_solver), "solve"); fetch(); }

private:
    void upload()
    {
        check(pmb_sqp_set_bounds_x(m_solver, m_lbx.data(), m_ubx.data(), m_d.N), "set_bounds_x");
        if (m_d.NG > 0) check(pmb_sqp_set_bounds_g(m_solver, m_lbg.data(), m_ubg.data(), m_d.NG * m_d.NN), "set_bounds_g");
        if (m_d.ND > 0) check(pmb_sqp_set_parameters(m_solver, m_p.data(), m_d.ND), "set_parameters");
        if (m_guess_dirty) {
            check(pmb_sqp_set_primal(m_solver, m_x.data(), m_d.N), "set_primal");
            check(pmb_sqp_set_dual(m_solver, m_lam.data(), m_d.DUAL), "set_dual");
            m_guess_dirty = false;
        }
    }
    void fetch()
    {
        check(pmb_sqp_get_primal(m_solver, m_x.data()), "get_primal");
        check(pmb_sqp_get_dual(m_solver, m_lam.data()), "get_dual");
        m_info.resize(m_batch);
        check(pmb_sqp_get_info(m_solver, m_info.data()), "get_info");
        m_stats.resize((size_t)m_batch * 4);
        check(pmb_sqp_get_stats(m_solver, m_stats.data()), "get_stats");
    }

public:

    /** solve() split in two (not in the reference): solve_async() uploads bounds / parameters / guess and enqueues the batch
     *  solve on this object's stream, wait() blocks and fetches the results.  Two BatchedMPC objects alternating
     *  solve_async() / wait() keep two batches in flight. */
    void solve_async() { upload(); check(pmb_sqp_solve_async(m_solver), "solve_async"); }
    void wait() { check(pmb_sqp_wait(m_solver), "wait"); fetch(); }

    // ---- results (mpc_wrapper.hpp:214-296)
    const pmb_sqp_info_t& info(int b) const { return m_info.at(b); }
    double cost(int b) const { return m_stats.at((size_t)b * 4 + 0); }
    double primal_norm(int b) const { return m_stats.at((size_t)b * 4 + 1); }
    double dual_norm(int b) const { return m_stats.at((size_t)b * 4 + 2); }
    double constr_violation(int b) const { return m_stats.at((size_t)b * 4 + 3); }
    vec solution_x(int b) const { return slice(m_x, b, m_d.N, 0, varx_size()); }
    vec solution_u(int b) const { return slice(m_x, b, m_d.N, varx_size(), varu_size()); }
    vec solution_dual(int b) const { return slice(m_lam, b, m_d.DUAL, 0, m_d.DUAL); }
    /** k-th node counted from the initial time (node 0 of the NLP variable is the FINAL time) */
    vec solution_x_at(int b, int k) const { return slice(m_x, b, m_d.N, varx_size() - (k + 1) * nx(), nx()); }
    vec solution_u_at(int b, int k) const { return slice(m_x, b, m_d.N, varx_size() + varu_size() - (k + 1) * nu(), nu()); }
    /** state / control at time t (relative to t_start): Lagrange interpolation on the Chebyshev nodes of the segment that
     *  contains t, as polympc::LagrangeSpline::eval does (src/polynomials/splines.hpp:101-139, mpc_wrapper.hpp:245-281) */
    vec solution_x_at(int b, double t) const { return interpolate(b, t, varx_size(), nx()); }
    vec solution_u_at(int b, double t) const { return interpolate(b, t, varx_size() + varu_size(), nu()); }
    /** time grid, ascending */
    vec time_grid() const
    {
        vec t(m_d.NN);
        check(pmb_ocp_time_nodes(pmb_sqp_problem(m_solver), t.data()), "time_nodes");
        for (int a = 0, z = m_d.NN - 1; a < z; ++a, --z) std::swap(t[a], t[z]);
        return t;
    }
    double last_solve_ms() const { return pmb_sqp_last_solve_ms(m_solver); }
    pmb_sqp_t* handle() { return m_solver; }

private:
    static void check(int rc, const char* what)
    { if (rc != PMB_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + pmb_last_error()); }
    /** src holds `len` values (broadcast) or batch*len values; written at [b*ld + off, +len) */
    void spread(const vec& src, vec& dst, int off, int len, int ld) const
    {
        if (len == 0) return;
        const bool per_instance = src.size() == (size_t)m_batch * len && m_batch > 1;
        if (!per_instance && src.size() != (size_t)len) throw std::invalid_argument("polympc::b200: vector of wrong length");
        for (int b = 0; b < m_batch; ++b)
            for (int i = 0; i < len; ++i) dst[(size_t)b * ld + off + i] = src[per_instance ? (size_t)b * len + i : i];
    }
    /** v.replicate(count, 1) written at offset off of every instance (one vector for the whole batch) */
    void replicate(const vec& v, vec& dst, int off, int len, int count) const
    {
        if ((int)v.size() != len) throw std::invalid_argument("polympc::b200: vector of wrong length");
        for (int b = 0; b < m_batch; ++b)
            for (int k = 0; k < count; ++k)
                for (int i = 0; i < len; ++i) dst[(size_t)b * m_d.N + off + k * len + i] = v[i];
    }
    static vec slice(const vec& a, int b, int ld, int off, int len) { return vec(a.begin() + (size_t)b * ld + off, a.begin() + (size_t)b * ld + off + len); }
    vec interpolate(int b, double t, int block_end, int width) const
    {
        const vec tg = time_grid();
        const int P = m_d.P, S = m_d.S;
        const double seg = (m_tf - m_t0) / S;
        int idx = (int)std::floor(t / seg);
        idx = idx < 0 ? 0 : (idx > S - 1 ? S - 1 : idx);
        const double tq = m_t0 + t;
        vec out(width, 0.0);
        for (int a = 0; a <= P; ++a) {
            const int ka = idx * P + a;                     // node index counted from the initial time
            double l = 1.0;
            for (int c = 0; c <= P; ++c) if (c != a) l *= (tq - tg[idx * P + c]) / (tg[ka] - tg[idx * P + c]);
            const size_t base = (size_t)b * m_d.N + block_end - (size_t)(ka + 1) * width;
            for (int i = 0; i < width; ++i) out[i] += l * m_x[base + i];
        }
        return out;
    }

    pmb_sqp_t* m_solver = nullptr;
    pmb_dims_t m_d{};
    int m_batch = 0;
    double m_t0 = 0.0, m_tf = 1.0;
    bool m_guess_dirty = false;
    vec m_lbx, m_ubx, m_lbg, m_ubg, m_x, m_lam, m_p, m_stats;
    std::vector<pmb_sqp_info_t> m_info;
};

} // namespace b200
} // namespace polympc
