// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE (CPU oracle).  Not part of the product; see oracle/README.md.
//
// Exports the C ABI of include/polympc_b200.h under the prefix orc_ (see orc_names.h), implemented by the plain-C++
// restatement of the reference algorithm in this directory.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
#include "orc_names.h"
#include "../include/polympc_b200.h"
#include "sqp.hpp"
#include "admm_qp.hpp"

#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace orc;

namespace {

int g_threads = 1;
std::string g_err;

template <class F>
void parallel_for(int n, F&& body)
{
    const int T = std::max(1, std::min(g_threads, n));
    if (T == 1) { for (int i = 0; i < n; ++i) body(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() { for (int i = t; i < n; i += T) body(i); });
    for (auto& x : th) x.join();
}

// ---- type-erased OCP ------------------------------------------------------------------------------------------
struct IOcp {
    pmb_dims_t dims{};
    virtual ~IOcp() {}
    virtual void set_params(const double*) = 0;
    virtual void get_params(double*) const = 0;
    virtual void set_time_limits(double, double) = 0;
    virtual void time_nodes(double*) const = 0;
    virtual void cost(const double*, const double*, double*) const = 0;
    virtual void equalities(const double*, const double*, double*) const = 0;
    virtual void inequalities(const double*, const double*, double*) const = 0;
    virtual void equalities_linearised(const double*, const double*, double*, double*) const = 0;
    virtual void cost_gradient(const double*, const double*, double*, double*) const = 0;
    virtual void cost_gradient_hessian(const double*, const double*, double*, double*, double*) const = 0;
    virtual void lagrangian_gradient(const double*, const double*, const double*, double*, double*, double*, double*, double*) const = 0;
    virtual void lagrangian_gradient_hessian(const double*, const double*, const double*, double*, double*, double*, double*, double*, double*) const = 0;
    virtual int block_bfgs(double* H, const double* s, const double* y) const = 0;
};

template <class O>
void fill_dims(pmb_dims_t& d)
{
    d.NX = O::NX; d.NU = O::NU; d.NP = O::NP; d.ND = O::ND; d.NG = O::NG; d.P = O::P; d.S = O::S; d.NN = O::NN;
    d.N = O::N; d.M = O::M; d.DUAL = O::DUAL; d.NPARAM = decltype(O::model)::NPARAM;
}

template <class O>
struct OcpImpl : IOcp {
    O o;
    OcpImpl() { fill_dims<O>(dims); }
    void set_params(const double* v) override { o.model.set_params(v); }
    void get_params(double* v) const override { o.model.get_params(v); }
    void set_time_limits(double a, double b) override { o.set_time_limits(a, b); }
    void time_nodes(double* t) const override { for (int i = 0; i < O::NN; ++i) t[i] = o.time_nodes[i]; }
    void cost(const double* v, const double* d, double* c) const override { o.cost(v, d, *c); }
    void equalities(const double* v, const double* d, double* c) const override { o.equalities(v, d, c); }
    void inequalities(const double* v, const double* d, double* g) const override { o.inequalities(v, d, g); }
    void equalities_linearised(const double* v, const double* d, double* c, double* A) const override { o.equalities_linearised(v, d, c, A, O::NUM_EQ); }
    void cost_gradient(const double* v, const double* d, double* c, double* g) const override { o.cost_gradient(v, d, *c, g); }
    void cost_gradient_hessian(const double* v, const double* d, double* c, double* g, double* H) const override { o.cost_gradient_hessian(v, d, *c, g, H); }
    void lagrangian_gradient(const double* v, const double* d, const double* l, double* c, double* lg, double* cg, double* g, double* A) const override
    { o.lagrangian_gradient(v, d, l, *c, lg, cg, g, A); }
    void lagrangian_gradient_hessian(const double* v, const double* d, const double* l, double* c, double* lg, double* H, double* cg, double* g, double* A) const override
    { o.lagrangian_gradient_hessian(v, d, l, *c, lg, H, cg, g, A); }
    int block_bfgs(double* H, const double* s, const double* y) const override { return block_bfgs_update<O>(H, s, y); }
};

// ---- type-erased SQP (one solver object per instance, like the reference) ---------------------------------------
struct ISqpInst {
    virtual ~ISqpInst() {}
    virtual void solve() = 0;
    virtual SqpSettings& settings() = 0;
    virtual QpSettings& qp_settings() = 0;
    virtual void hessian_options(int exact, int gershgorin) = 0;
    virtual void hessian_update(int block) = 0;
    virtual void preconditioner(int kind) = 0;
    virtual void qp_solver(int kind) = 0;
    virtual void line_search(int kind, double beta, int depth) = 0;
    virtual LsFilter& filter() = 0;
    virtual SqpInfo& info() = 0;
    virtual std::vector<double>& x() = 0;
    virtual std::vector<double>& lam() = 0;
    virtual std::vector<double>& lbx() = 0;
    virtual std::vector<double>& ubx() = 0;
    virtual std::vector<double>& lbg() = 0;
    virtual std::vector<double>& ubg() = 0;
    virtual std::vector<double>& d() = 0;
    virtual void stats(double*) const = 0;
    virtual void trace(int rows, int*, double*, int*, int*, int*) const = 0;
    virtual void sync_problem(const IOcp& src) = 0;
};

template <class O>
struct SqpInst : ISqpInst {
    Sqp<O> s;
    void solve() override { s.solve(); }
    SqpSettings& settings() override { return s.settings; }
    QpSettings& qp_settings() override { return s.qp.settings; }
    void hessian_options(int exact, int gershgorin) override { s.opt_exact_hessian = exact; s.opt_gershgorin = gershgorin; }
    void hessian_update(int block) override { s.opt_block_bfgs = block; }
    void preconditioner(int kind) override { s.opt_precond = kind; }
    void qp_solver(int kind) override { s.opt_qp_solver = kind; }
    void line_search(int kind, double beta, int depth) override { s.opt_line_search = kind; s.filter.clear(); s.filter.beta = beta; s.filter.max_depth = depth; }
    LsFilter& filter() override { return s.filter; }
    SqpInfo& info() override { return s.info; }
    std::vector<double>& x() override { return s.x; }
    std::vector<double>& lam() override { return s.lam; }
    std::vector<double>& lbx() override { return s.lbx; }
    std::vector<double>& ubx() override { return s.ubx; }
    std::vector<double>& lbg() override { return s.lbg; }
    std::vector<double>& ubg() override { return s.ubg; }
    std::vector<double>& d() override { return s.p_static; }
    void stats(double* o) const override { o[0] = s.cost_val; o[1] = s.primal_norm; o[2] = s.dual_norm; o[3] = s.max_violation; }
    void trace(int rows, int* qi, double* al, int* bf, int* ls, int* qf) const override
    {
        for (int r = 0; r < rows; ++r) {
            const bool have = r < (int)s.tr_qp_iter.size();
            if (qi) qi[r] = have ? s.tr_qp_iter[r] : -1;
            if (al) al[r] = have ? s.tr_alpha[r] : std::nan("");
            if (bf) bf[r] = have ? s.tr_bfgs[r] : -1;
            if (ls) ls[r] = have ? s.tr_ls_trials[r] : -1;
            if (qf) qf[r] = have ? s.tr_qp_factor[r] : -1;
        }
    }
    void sync_problem(const IOcp& src) override
    {
        const auto& so = static_cast<const OcpImpl<O>&>(src).o;
        s.problem.model = so.model;
        s.problem.set_time_limits(so.t_start, so.t_stop);
    }
};

struct Registry {
    const char* name;
    IOcp* (*make_ocp)();
    ISqpInst* (*make_sqp)();
};
template <class O> IOcp* mk_ocp() { return new OcpImpl<O>(); }
template <class O> ISqpInst* mk_sqp() { return new SqpInst<O>(); }
#define REG(NAME, MODEL, P, S) { NAME, &mk_ocp<Ocp<MODEL, P, S>>, &mk_sqp<Ocp<MODEL, P, S>> }
const Registry g_registry[] = {
    REG("mobile_robot_6x2", RobotModel, 6, 2),   // BASELINE.json configs 1,2,5
    REG("mobile_robot_5x2", RobotModel, 5, 2),   // CasADi fixture / continuous_ocp_test.cpp
    REG("mobile_robot_5x3", RobotModel, 5, 3),   // mpc_wrapper_test.cpp
    REG("cstr_5x2", CstrModel, 5, 2),            // cstr_control_test.cpp, BASELINE.json config 3
    REG("kite_12x1", KiteModel, 12, 1),          // BASELINE.json config 4 (our model)
    REG("kite_4x2", KiteModel, 4, 2),            // small kite variant for fast parity tests
    REG("robot_obstacle_5x2", RobotObstacleModel, 5, 2),   // NG = 1: generic inequality constraints
    REG("parking_5x2", ParkingModel, 5, 2),      // NP = 1: minimal_time_test.cpp / dense_sparse_compare.cpp
};
const int g_nreg = sizeof(g_registry) / sizeof(g_registry[0]);
const Registry* find(const char* name)
{
    if (!name) return nullptr;
    for (int i = 0; i < g_nreg; ++i) if (std::strcmp(g_registry[i].name, name) == 0) return &g_registry[i];
    return nullptr;
}

void to_qp(const pmb_qp_settings_t& a, QpSettings& b)
{
    b.eps_rel = a.eps_rel; b.eps_abs = a.eps_abs; b.max_iter = a.max_iter; b.warm_start = a.warm_start;
    b.reuse_pattern = a.reuse_pattern; b.verbose = a.verbose; b.rho = a.rho; b.sigma = a.sigma; b.alpha = a.alpha;
    b.check_termination = a.check_termination; b.adaptive_rho = a.adaptive_rho;
    b.adaptive_rho_tolerance = a.adaptive_rho_tolerance; b.adaptive_rho_interval = a.adaptive_rho_interval;
}
void from_qp(const QpSettings& b, pmb_qp_settings_t& a)
{
    a.eps_rel = b.eps_rel; a.eps_abs = b.eps_abs; a.max_iter = b.max_iter; a.warm_start = b.warm_start;
    a.reuse_pattern = b.reuse_pattern; a.verbose = b.verbose; a.rho = b.rho; a.sigma = b.sigma; a.alpha = b.alpha;
    a.check_termination = b.check_termination; a.adaptive_rho = b.adaptive_rho;
    a.adaptive_rho_tolerance = b.adaptive_rho_tolerance; a.adaptive_rho_interval = b.adaptive_rho_interval; a._pad = 0;
}

} // namespace

struct pmb_ocp { std::unique_ptr<IOcp> impl; };
struct pmb_sqp {
    pmb_ocp ocp;
    const Registry* reg = nullptr;
    int batch = 0;
    std::vector<std::unique_ptr<ISqpInst>> inst;
    std::vector<std::vector<double>> x_guess, lam_guess;   // values last given to set_primal / set_dual (reset_guess)
    double last_ms = 0;
};

extern "C" {

/** oracle-only knob: number of host threads used to sweep the batch (CPU baseline timing) */
void orc_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }
int orc_get_num_threads(void) { return g_threads; }
void orc_set_ldlt_variant(int v) { orc::ldlt_variant() = v; }

const char* pmb_version(void) { return "polympc-oracle 0.1 (CPU restatement, test infrastructure)"; }
const char* pmb_last_error(void) { return g_err.c_str(); }
int pmb_device_count(void) { return 0; }
int pmb_problem_count(void) { return g_nreg; }
const char* pmb_problem_name(int i) { return (i >= 0 && i < g_nreg) ? g_registry[i].name : nullptr; }
int pmb_problem_dims(const char* name, pmb_dims_t* out)
{
    const Registry* r = find(name);
    if (!r) return PMB_ERR_UNKNOWN_PROBLEM;
    if (!out) return PMB_ERR_BAD_ARGUMENT;
    std::unique_ptr<IOcp> o(r->make_ocp());
    *out = o->dims;
    return PMB_OK;
}

void pmb_qp_default_settings(pmb_qp_settings_t* s) { if (s) from_qp(QpSettings(), *s); }
void pmb_sqp_default_settings(pmb_sqp_settings_t* s)
{
    if (!s) return;
    SqpSettings d;
    s->tau = d.tau; s->eta = d.eta; s->rho = d.rho; s->eps_prim = d.eps_prim; s->eps_dual = d.eps_dual;
    s->max_iter = d.max_iter; s->line_search_max_iter = d.line_search_max_iter;
}
void pmb_sqp_default_qp_settings(pmb_qp_settings_t* s)
{
    if (!s) return;
    QpSettings q;
    q.warm_start = 0; q.check_termination = 10; q.eps_abs = 1e-4; q.eps_rel = 1e-4; q.max_iter = 100;
    q.adaptive_rho = 1; q.adaptive_rho_interval = 50; q.alpha = 1.0;
    from_qp(q, *s);
}

int pmb_cheb_tables(int P, double* nodes, double* D, double* w)
{
    if (P < 2 || !nodes || !D || !w) return PMB_ERR_BAD_ARGUMENT;
    const ChebTables t = cheb_tables(P);
    for (int i = 0; i <= P; ++i) { nodes[i] = t.nodes[i]; w[i] = t.w[i]; }
    for (int i = 0; i < (P + 1) * (P + 1); ++i) D[i] = t.D[i];
    return PMB_OK;
}

pmb_ocp_t* pmb_ocp_create(const char* name, int)
{
    const Registry* r = find(name);
    if (!r) { g_err = "unknown problem"; return nullptr; }
    pmb_ocp_t* h = new pmb_ocp_t();
    h->impl.reset(r->make_ocp());
    return h;
}
void pmb_ocp_destroy(pmb_ocp_t* h) { delete h; }
int pmb_ocp_dims(const pmb_ocp_t* h, pmb_dims_t* out) { if (!h || !out) return PMB_ERR_BAD_ARGUMENT; *out = h->impl->dims; return PMB_OK; }
int pmb_ocp_set_params(pmb_ocp_t* h, const double* v, int n)
{ if (!h || !v || n != h->impl->dims.NPARAM) return PMB_ERR_BAD_ARGUMENT; h->impl->set_params(v); return PMB_OK; }
int pmb_ocp_get_params(const pmb_ocp_t* h, double* v, int n)
{ if (!h || !v || n != h->impl->dims.NPARAM) return PMB_ERR_BAD_ARGUMENT; h->impl->get_params(v); return PMB_OK; }
int pmb_ocp_set_time_limits(pmb_ocp_t* h, double t0, double tf) { if (!h) return PMB_ERR_BAD_ARGUMENT; h->impl->set_time_limits(t0, tf); return PMB_OK; }
int pmb_ocp_time_nodes(const pmb_ocp_t* h, double* t) { if (!h || !t) return PMB_ERR_BAD_ARGUMENT; h->impl->time_nodes(t); return PMB_OK; }

#define OCP_PRE                                                                                                      \
    if (!h || batch < 0 || !var) return PMB_ERR_BAD_ARGUMENT;                                                         \
    const pmb_dims_t& D = h->impl->dims;                                                                              \
    if (D.ND > 0 && !d) return PMB_ERR_BAD_ARGUMENT;                                                                  \
    const IOcp& o = *h->impl;                                                                                         \
    static const double dzero[1] = {0.0};                                                                             \
    auto dp = [&](int b) { return D.ND > 0 ? d + (size_t)b * D.ND : dzero; };                                         \
    (void)o; (void)dp;

int pmb_ocp_cost(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost)
{ OCP_PRE parallel_for(batch, [&](int b) { o.cost(var + (size_t)b * D.N, dp(b), cost + b); }); return PMB_OK; }
int pmb_ocp_equalities(pmb_ocp_t* h, int batch, const double* var, const double* d, double* c)
{ OCP_PRE parallel_for(batch, [&](int b) { o.equalities(var + (size_t)b * D.N, dp(b), c + (size_t)b * D.NX * D.NN); }); return PMB_OK; }
int pmb_ocp_inequalities(pmb_ocp_t* h, int batch, const double* var, const double* d, double* g)
{ OCP_PRE parallel_for(batch, [&](int b) { o.inequalities(var + (size_t)b * D.N, dp(b), g + (size_t)b * D.NG * D.NN); }); return PMB_OK; }
int pmb_ocp_equalities_linearised(pmb_ocp_t* h, int batch, const double* var, const double* d, double* c, double* jac)
{
    OCP_PRE
    const size_t ne = (size_t)D.NX * D.NN;
    parallel_for(batch, [&](int b) { o.equalities_linearised(var + (size_t)b * D.N, dp(b), c + b * ne, jac + b * ne * D.N); });
    return PMB_OK;
}
int pmb_ocp_cost_gradient(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost, double* grad)
{ OCP_PRE parallel_for(batch, [&](int b) { o.cost_gradient(var + (size_t)b * D.N, dp(b), cost + b, grad + (size_t)b * D.N); }); return PMB_OK; }
int pmb_ocp_cost_gradient_hessian(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost, double* grad, double* hess)
{
    OCP_PRE
    parallel_for(batch, [&](int b) { o.cost_gradient_hessian(var + (size_t)b * D.N, dp(b), cost + b, grad + (size_t)b * D.N, hess + (size_t)b * D.N * D.N); });
    return PMB_OK;
}
int pmb_ocp_lagrangian_gradient(pmb_ocp_t* h, int batch, const double* var, const double* d, const double* lam, double* cost,
                                double* lag_grad, double* cost_grad, double* g, double* jac)
{
    OCP_PRE
    parallel_for(batch, [&](int b) {
        o.lagrangian_gradient(var + (size_t)b * D.N, dp(b), lam + (size_t)b * D.DUAL, cost + b, lag_grad + (size_t)b * D.N,
                              cost_grad + (size_t)b * D.N, g + (size_t)b * D.M, jac + (size_t)b * D.M * D.N);
    });
    return PMB_OK;
}
int pmb_ocp_lagrangian_gradient_hessian(pmb_ocp_t* h, int batch, const double* var, const double* d, const double* lam, double* cost,
                                        double* lag_grad, double* lag_hess, double* cost_grad, double* g, double* jac)
{
    OCP_PRE
    parallel_for(batch, [&](int b) {
        o.lagrangian_gradient_hessian(var + (size_t)b * D.N, dp(b), lam + (size_t)b * D.DUAL, cost + b, lag_grad + (size_t)b * D.N,
                                      lag_hess + (size_t)b * D.N * D.N, cost_grad + (size_t)b * D.N, g + (size_t)b * D.M,
                                      jac + (size_t)b * D.M * D.N);
    });
    return PMB_OK;
}

int pmb_qp_solve(int N, int M, int batch, const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                 const double* xlb, const double* xub, const double* x_guess, const double* y_guess, const pmb_qp_settings_t* st,
                 double* x, double* y, pmb_qp_info_t* info, double* z, double* q, int* perm, int* ctype, int* n_factor)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || (M > 0 && (!A || !Alb || !Aub)) || !xlb || !xub || !st || !x || !y || !info)
        return PMB_ERR_BAD_ARGUMENT;
    parallel_for(batch, [&](int b) {
        BoxAdmm s(N, M);
        to_qp(*st, s.settings);
        s.solve(H + (size_t)b * N * N, h + (size_t)b * N, A + (size_t)b * M * N, Alb + (size_t)b * M, Aub + (size_t)b * M,
                xlb + (size_t)b * N, xub + (size_t)b * N, x_guess ? x_guess + (size_t)b * N : nullptr,
                y_guess ? y_guess + (size_t)b * (N + M) : nullptr);
        for (int i = 0; i < N; ++i) x[(size_t)b * N + i] = s.x[i];
        for (int i = 0; i < N + M; ++i) y[(size_t)b * (N + M) + i] = s.y[i];
        info[b].status = s.info.status; info[b].iter = s.info.iter; info[b].rho_updates = s.info.rho_updates; info[b]._pad = 0;
        info[b].rho_estimate = s.info.rho_estimate; info[b].res_prim = s.info.res_prim; info[b].res_dual = s.info.res_dual;
        if (z) for (int i = 0; i < M; ++i) z[(size_t)b * M + i] = s.z[i];
        if (q) for (int i = 0; i < N; ++i) q[(size_t)b * N + i] = s.q[i];
        if (perm) for (int i = 0; i < N + M; ++i) perm[(size_t)b * (N + M) + i] = s.first_perm[i];
        if (ctype) {
            for (int i = 0; i < M; ++i) ctype[(size_t)b * (N + M) + i] = s.constr_type[i];
            for (int i = 0; i < N; ++i) ctype[(size_t)b * (N + M) + M + i] = s.box_constr_type[i];
        }
        if (n_factor) n_factor[b] = s.n_factor;
    });
    return PMB_OK;
}

int pmb_qp_solve_admm(int N, int M, int batch, const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                      const double* xlb, const double* xub, const double* x_guess, const double* y_guess, const pmb_qp_settings_t* st,
                      double* x, double* y, pmb_qp_info_t* info, double* z, int* perm, int* ctype, int* n_factor)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || (M > 0 && (!A || !Alb || !Aub)) || !xlb || !xub || !st || !x || !y || !info)
        return PMB_ERR_BAD_ARGUMENT;
    parallel_for(batch, [&](int b) {
        OsqpAdmm s(N, M);
        to_qp(*st, s.settings);
        s.solve(H + (size_t)b * N * N, h + (size_t)b * N, A + (size_t)b * M * N, Alb + (size_t)b * M, Aub + (size_t)b * M,
                xlb + (size_t)b * N, xub + (size_t)b * N, x_guess ? x_guess + (size_t)b * N : nullptr,
                y_guess ? y_guess + (size_t)b * (N + M) : nullptr);
        const size_t Me = (size_t)N + M, Kd = 2 * (size_t)N + M;
        for (int i = 0; i < N; ++i) x[(size_t)b * N + i] = s.x[i];
        for (size_t i = 0; i < Me; ++i) y[b * Me + i] = s.y[i];
        info[b].status = s.info.status; info[b].iter = s.info.iter; info[b].rho_updates = s.info.rho_updates; info[b]._pad = 0;
        info[b].rho_estimate = s.info.rho_estimate; info[b].res_prim = s.info.res_prim; info[b].res_dual = s.info.res_dual;
        if (z) for (size_t i = 0; i < Me; ++i) z[b * Me + i] = s.z[i];
        if (perm) for (size_t i = 0; i < Kd; ++i) perm[b * Kd + i] = s.first_perm[i];
        if (ctype) {
            for (int i = 0; i < M; ++i) ctype[b * Me + i] = s.constr_type[i];
            for (int i = 0; i < N; ++i) ctype[b * Me + M + i] = s.box_constr_type[i];
        }
        if (n_factor) n_factor[b] = s.n_factor;
    });
    return PMB_OK;
}

int pmb_kkt_assemble(int N, int M, int batch, const double* H, const double* A, const double* rho_box, const double* rho_inv,
                     double sigma, double* K)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !A || !rho_box || !rho_inv || !K) return PMB_ERR_BAD_ARGUMENT;
    const size_t Kd = N + M;
    parallel_for(batch, [&](int b) {
        BoxAdmm s(N, M);
        s.settings.sigma = sigma;
        for (int i = 0; i < N; ++i) s.rho_box[i] = rho_box[(size_t)b * N + i];
        for (int i = 0; i < M; ++i) s.rho_inv_vec[i] = rho_inv[(size_t)b * M + i];
        s.construct_kkt_matrix(H + (size_t)b * N * N, A + (size_t)b * M * N);
        std::memcpy(K + b * Kd * Kd, s.K.data(), Kd * Kd * sizeof(double));
    });
    return PMB_OK;
}

int pmb_bfgs_update(int N, int batch, double* B, const double* s, const double* y, int* branch)
{
    if (N <= 0 || batch < 0 || !B || !s || !y) return PMB_ERR_BAD_ARGUMENT;
    parallel_for(batch, [&](int b) {
        const int br = bfgs_update(B + (size_t)b * N * N, s + (size_t)b * N, y + (size_t)b * N, N);
        if (branch) branch[b] = br;
    });
    return PMB_OK;
}

int pmb_ocp_block_bfgs_update(pmb_ocp_t* h, int batch, double* B, const double* s, const double* y, int* branch)
{
    if (!h || batch < 0 || !B || !s || !y) return PMB_ERR_BAD_ARGUMENT;
    const size_t N = h->impl->dims.N;
    parallel_for(batch, [&](int b) {
        const int br = h->impl->block_bfgs(B + (size_t)b * N * N, s + (size_t)b * N, y + (size_t)b * N);
        if (branch) branch[b] = br;
    });
    return PMB_OK;
}

// ---- SQP ---------------------------------------------------------------------------------------------------------
pmb_sqp_t* pmb_sqp_create(const char* name, int batch, int)
{
    const Registry* r = find(name);
    if (!r || batch <= 0) { g_err = "unknown problem or bad batch"; return nullptr; }
    pmb_sqp_t* s = new pmb_sqp_t();
    s->reg = r; s->batch = batch;
    s->ocp.impl.reset(r->make_ocp());
    s->inst.resize(batch);
    for (int b = 0; b < batch; ++b) s->inst[b].reset(r->make_sqp());
    return s;
}
void pmb_sqp_destroy(pmb_sqp_t* s) { delete s; }
pmb_ocp_t* pmb_sqp_problem(pmb_sqp_t* s) { return s ? &s->ocp : nullptr; }
int pmb_sqp_batch(const pmb_sqp_t* s) { return s ? s->batch : PMB_ERR_BAD_ARGUMENT; }
int pmb_sqp_set_settings(pmb_sqp_t* s, const pmb_sqp_settings_t* st)
{
    if (!s || !st) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) {
        SqpSettings& d = i->settings();
        d.tau = st->tau; d.eta = st->eta; d.rho = st->rho; d.eps_prim = st->eps_prim; d.eps_dual = st->eps_dual;
        d.max_iter = st->max_iter; d.line_search_max_iter = st->line_search_max_iter;
    }
    return PMB_OK;
}
int pmb_sqp_get_settings(const pmb_sqp_t* s, pmb_sqp_settings_t* st)
{
    if (!s || !st) return PMB_ERR_BAD_ARGUMENT;
    const SqpSettings& d = s->inst[0]->settings();
    st->tau = d.tau; st->eta = d.eta; st->rho = d.rho; st->eps_prim = d.eps_prim; st->eps_dual = d.eps_dual;
    st->max_iter = d.max_iter; st->line_search_max_iter = d.line_search_max_iter;
    return PMB_OK;
}
int pmb_sqp_set_qp_settings(pmb_sqp_t* s, const pmb_qp_settings_t* st)
{ if (!s || !st) return PMB_ERR_BAD_ARGUMENT; for (auto& i : s->inst) to_qp(*st, i->qp_settings()); return PMB_OK; }
int pmb_sqp_get_qp_settings(const pmb_sqp_t* s, pmb_qp_settings_t* st)
{ if (!s || !st) return PMB_ERR_BAD_ARGUMENT; from_qp(s->inst[0]->qp_settings(), *st); return PMB_OK; }
int pmb_sqp_set_hessian_options(pmb_sqp_t* s, int exact_every_iteration, int gershgorin_regularisation)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->hessian_options(exact_every_iteration != 0, gershgorin_regularisation != 0);
    return PMB_OK;
}
int pmb_sqp_set_hessian_update(pmb_sqp_t* s, int mode)
{
    if (!s || (mode != PMB_HESSIAN_BFGS_DENSE && mode != PMB_HESSIAN_BFGS_BLOCK)) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->hessian_update(mode == PMB_HESSIAN_BFGS_BLOCK);
    return PMB_OK;
}
int pmb_sqp_set_qp_solver(pmb_sqp_t* s, int kind)
{
    if (!s || (kind != PMB_QP_BOX_ADMM && kind != PMB_QP_OSQP_ADMM)) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->qp_solver(kind);
    return PMB_OK;
}
int pmb_sqp_set_preconditioner(pmb_sqp_t* s, int kind)
{
    if (!s || kind < PMB_PRECOND_IDENTITY || kind > PMB_PRECOND_RUIZ_SPARSE) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->preconditioner(kind);
    return PMB_OK;
}
int pmb_sqp_set_line_search(pmb_sqp_t* s, int kind, double beta, int depth)
{
    if (!s || (kind != PMB_LS_L1_MERIT && kind != PMB_LS_FILTER)) return PMB_ERR_BAD_ARGUMENT;
    if (kind == PMB_LS_FILTER && (depth < 1 || depth > PMB_FILTER_CAP || !(beta == beta))) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->line_search(kind, beta, kind == PMB_LS_FILTER ? depth : 10);
    return PMB_OK;
}
int pmb_sqp_set_filter(pmb_sqp_t* s, const double* st, int stride)
{
    if (!s || !st || (stride != 0 && stride != PMB_FILTER_DOUBLES)) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        const double* f = st + (size_t)b * stride;
        const int n = (int)f[0];
        if (n < 0 || n > PMB_FILTER_CAP) return PMB_ERR_BAD_ARGUMENT;
        LsFilter& F = s->inst[b]->filter();
        F.size = n;
        for (int k = 0; k < PMB_FILTER_CAP; ++k) { F.cost[k] = f[1 + k]; F.constr[k] = f[1 + PMB_FILTER_CAP + k]; }
    }
    return PMB_OK;
}
int pmb_sqp_get_filter(const pmb_sqp_t* s, double* st)
{
    if (!s || !st) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        double* f = st + (size_t)b * PMB_FILTER_DOUBLES;
        const LsFilter& F = s->inst[b]->filter();
        f[0] = F.size;
        for (int k = 0; k < PMB_FILTER_CAP; ++k) { f[1 + k] = k < F.size ? F.cost[k] : 0.0; f[1 + PMB_FILTER_CAP + k] = k < F.size ? F.constr[k] : 0.0; }
    }
    return PMB_OK;
}
int pmb_ruiz_equilibrate(int N, int M, int batch, int variant, double* H, double* h, double* A, double* Al, double* Au, double* l, double* u,
                         double* D, double* E, double* c)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || !A || !Al || !Au || !l || !u || !D || !E || !c) return PMB_ERR_BAD_ARGUMENT;
    if (variant != PMB_PRECOND_RUIZ_DENSE && variant != PMB_PRECOND_RUIZ_SPARSE) return PMB_ERR_BAD_ARGUMENT;
    parallel_for(batch, [&](int b) {
        Ruiz r(N, M, variant);
        r.compute(H + (size_t)b * N * N, h + (size_t)b * N, A + (size_t)b * M * N, Al + (size_t)b * M, Au + (size_t)b * M, l + (size_t)b * N, u + (size_t)b * N);
        for (int k = 0; k < N; ++k) D[(size_t)b * N + k] = r.D[k];
        for (int k = 0; k < M; ++k) E[(size_t)b * M + k] = r.E[k];
        c[b] = r.c;
    });
    return PMB_OK;
}
int pmb_ruiz_unscale(int N, int M, int batch, const double* D, const double* E, const double* c, double* H, double* h, double* A, double* Al,
                     double* Au, double* l, double* u, double* x, double* y)
{
    if (N <= 0 || M < 0 || batch < 0 || !D || !E || !c || !H || !h || !A || !Al || !Au || !l || !u) return PMB_ERR_BAD_ARGUMENT;
    parallel_for(batch, [&](int b) {
        Ruiz r(N, M, PMB_PRECOND_RUIZ_DENSE);
        for (int k = 0; k < N; ++k) r.D[k] = D[(size_t)b * N + k];
        for (int k = 0; k < M; ++k) r.E[k] = E[(size_t)b * M + k];
        r.c = c[b];
        if (x && y) r.unscale_solution(x + (size_t)b * N, y + (size_t)b * (M + N));
        r.unscale_data(H + (size_t)b * N * N, h + (size_t)b * N, A + (size_t)b * M * N, Al + (size_t)b * M, Au + (size_t)b * M, l + (size_t)b * N, u + (size_t)b * N);
    });
    return PMB_OK;
}
/* the oracle has one arithmetic: the canonical orders of canon.hpp */
int pmb_set_default_arithmetic(int mode) { return mode == PMB_ARITH_EXACT ? PMB_OK : PMB_ERR_UNSUPPORTED; }
int pmb_get_default_arithmetic(void) { return PMB_ARITH_EXACT; }
int pmb_sqp_set_arithmetic(pmb_sqp_t* s, int mode) { return !s ? PMB_ERR_BAD_ARGUMENT : (mode == PMB_ARITH_EXACT ? PMB_OK : PMB_ERR_UNSUPPORTED); }
int pmb_sqp_get_arithmetic(const pmb_sqp_t* s) { return s ? (int)PMB_ARITH_EXACT : (int)PMB_ERR_BAD_ARGUMENT; }
int pmb_sqp_set_schedule(pmb_sqp_t* s, int) { return s ? PMB_OK : PMB_ERR_BAD_ARGUMENT; }   /* no queue in the oracle */
int pmb_sqp_set_trace(pmb_sqp_t* s, int) { return s ? PMB_OK : PMB_ERR_BAD_ARGUMENT; }   /* the oracle always records its traces */
/* the oracle has no kernels to register: problem classes are added to its own table (REG above) */
int pmb_register_problem(const char*, void* (*)(void)) { g_err = "the oracle does not register external problems"; return PMB_ERR_BAD_ARGUMENT; }

static int set_vec(pmb_sqp_t* s, const double* v, int stride, int len, std::vector<double>& (ISqpInst::*acc)())
{
    if (!s || (!v && len > 0) || (stride != 0 && stride != len)) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        std::vector<double>& dst = ((*s->inst[b]).*acc)();
        for (int i = 0; i < len; ++i) dst[i] = v[(size_t)b * stride + i];
    }
    return PMB_OK;
}
int pmb_sqp_set_bounds_x(pmb_sqp_t* s, const double* lb, const double* ub, int stride)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    const int N = s->ocp.impl->dims.N;
    int r = set_vec(s, lb, stride, N, &ISqpInst::lbx);
    return r ? r : set_vec(s, ub, stride, N, &ISqpInst::ubx);
}
int pmb_sqp_set_bounds_g(pmb_sqp_t* s, const double* lb, const double* ub, int stride)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    const int n = s->ocp.impl->dims.NG * s->ocp.impl->dims.NN;
    if (n == 0) return PMB_OK;
    int r = set_vec(s, lb, stride, n, &ISqpInst::lbg);
    return r ? r : set_vec(s, ub, stride, n, &ISqpInst::ubg);
}
int pmb_sqp_set_parameters(pmb_sqp_t* s, const double* d, int stride)
{ if (!s) return PMB_ERR_BAD_ARGUMENT; const int n = s->ocp.impl->dims.ND; return n == 0 ? PMB_OK : set_vec(s, d, stride, n, &ISqpInst::d); }
int pmb_sqp_set_primal(pmb_sqp_t* s, const double* x, int stride)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    const int r = set_vec(s, x, stride, s->ocp.impl->dims.N, &ISqpInst::x);
    if (r == PMB_OK) { s->x_guess.resize(s->batch); for (int b = 0; b < s->batch; ++b) s->x_guess[b] = s->inst[b]->x(); }
    return r;
}
int pmb_sqp_set_dual(pmb_sqp_t* s, const double* l, int stride)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    const int r = set_vec(s, l, stride, s->ocp.impl->dims.DUAL, &ISqpInst::lam);
    if (r == PMB_OK) { s->lam_guess.resize(s->batch); for (int b = 0; b < s->batch; ++b) s->lam_guess[b] = s->inst[b]->lam(); }
    return r;
}
int pmb_sqp_set_initial_conditions(pmb_sqp_t* s, const double* x0_lb, const double* x0_ub)
{
    if (!s || !x0_lb || !x0_ub) return PMB_ERR_BAD_ARGUMENT;
    const pmb_dims_t& D = s->ocp.impl->dims;
    const int off = D.NX * D.NN - D.NX;
    for (int b = 0; b < s->batch; ++b)
        for (int i = 0; i < D.NX; ++i) {
            s->inst[b]->lbx()[off + i] = x0_lb[(size_t)b * D.NX + i];
            s->inst[b]->ubx()[off + i] = x0_ub[(size_t)b * D.NX + i];
        }
    return PMB_OK;
}
int pmb_sqp_solve(pmb_sqp_t* s);
/* the CPU oracle is synchronous: _async does the work, _wait has nothing to wait for */
int pmb_sqp_solve_async(pmb_sqp_t* s) { return pmb_sqp_solve(s); }
int pmb_sqp_wait(pmb_sqp_t* s) { return s ? PMB_OK : PMB_ERR_BAD_ARGUMENT; }

int pmb_sqp_solve(pmb_sqp_t* s)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    for (auto& i : s->inst) i->sync_problem(*s->ocp.impl);
    parallel_for(s->batch, [&](int b) { s->inst[b]->solve(); });
    return PMB_OK;
}
int pmb_sqp_get_primal(const pmb_sqp_t* s, double* x)
{
    if (!s || !x) return PMB_ERR_BAD_ARGUMENT;
    const int N = s->ocp.impl->dims.N;
    for (int b = 0; b < s->batch; ++b) std::memcpy(x + (size_t)b * N, s->inst[b]->x().data(), N * sizeof(double));
    return PMB_OK;
}
int pmb_sqp_get_dual(const pmb_sqp_t* s, double* l)
{
    if (!s || !l) return PMB_ERR_BAD_ARGUMENT;
    const int n = s->ocp.impl->dims.DUAL;
    for (int b = 0; b < s->batch; ++b) std::memcpy(l + (size_t)b * n, s->inst[b]->lam().data(), n * sizeof(double));
    return PMB_OK;
}
int pmb_sqp_get_info(const pmb_sqp_t* s, pmb_sqp_info_t* info)
{
    if (!s || !info) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        const SqpInfo& i = s->inst[b]->info();
        info[b].iter = i.iter; info[b].qp_solver_iter = i.qp_solver_iter; info[b].status = i.status;
    }
    return PMB_OK;
}
int pmb_sqp_get_stats(const pmb_sqp_t* s, double* st)
{ if (!s || !st) return PMB_ERR_BAD_ARGUMENT; for (int b = 0; b < s->batch; ++b) s->inst[b]->stats(st + 4 * (size_t)b); return PMB_OK; }
int pmb_sqp_get_trace(const pmb_sqp_t* s, int rows, int* qi, double* al, int* bf, int* ls, int* qf)
{
    if (!s || rows <= 0) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        const size_t o = (size_t)b * rows;
        s->inst[b]->trace(rows, qi ? qi + o : nullptr, al ? al + o : nullptr, bf ? bf + o : nullptr, ls ? ls + o : nullptr, qf ? qf + o : nullptr);
    }
    return PMB_OK;
}
double pmb_sqp_last_solve_ms(const pmb_sqp_t*) { return 0.0; }
double pmb_sqp_last_kernel_ms(const pmb_sqp_t*) { return 0.0; }
long long pmb_sqp_last_solve_launches(const pmb_sqp_t*) { return 0; }
int pmb_sqp_set_stream(pmb_sqp_t*, void*) { return PMB_OK; }
int pmb_dm_eval(int fn, int n, const double* x, const double* y, double* out)
{
    if (fn < 0 || fn >= PMB_DM_COUNT || n < 0 || !x || !out) return PMB_ERR_BAD_ARGUMENT;
    for (int i = 0; i < n; ++i) out[i] = pmb::dm::dm_dispatch(fn, x[i], y ? y[i] : 0.0);   // the host instantiation of the same header
    return PMB_OK;
}
int pmb_sqp_set_profiling(pmb_sqp_t*, int) { return PMB_OK; }
int pmb_sqp_get_phase_cycles(const pmb_sqp_t*, unsigned long long* c) { if (c) for (int k = 0; k < 16; ++k) c[k] = 0; return PMB_OK; }
int pmb_sqp_get_kernel_times(const pmb_sqp_t*, double* ms, long long* n) { for (int k = 0; k < 3; ++k) { if (ms) ms[k] = 0; if (n) n[k] = 0; } return PMB_OK; }
int pmb_sqp_reset_guess(pmb_sqp_t* s)
{
    if (!s) return PMB_ERR_BAD_ARGUMENT;
    for (int b = 0; b < s->batch; ++b) {
        if (!s->x_guess.empty()) s->inst[b]->x() = s->x_guess[b]; else std::fill(s->inst[b]->x().begin(), s->inst[b]->x().end(), 0.0);
        if (!s->lam_guess.empty()) s->inst[b]->lam() = s->lam_guess[b]; else std::fill(s->inst[b]->lam().begin(), s->inst[b]->lam().end(), 0.0);
    }
    return PMB_OK;
}
int pmb_kkt_assemble_dev(int N, int M, int batch, const double* H, const double* A, const double* rho_box, const double* rho_inv, double sigma,
                         double* K, void*)
{ return pmb_kkt_assemble(N, M, batch, H, A, rho_box, rho_inv, sigma, K); }   /* the oracle's "device" is the host */

} // extern "C"
