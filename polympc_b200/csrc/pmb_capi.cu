// pmb_capi.cu — the extern "C" boundary of include/polympc_b200.h: problem registry, device workspaces, kernel launches.
// Compiled by nvcc for sm_100a into polympc_b200/libpolympc_b200.so.  There is no CPU execution path in that library:
// every compute entry point returns PMB_ERR_NO_DEVICE when no CUDA device is present.
// (tests/warp_emu compiles this same file with g++ -DPMB_EMU into a test-only library; see pmb_warp.hpp.)
#ifdef PMB_EMU
#include "emu_names.h"
#endif
#include "../../include/polympc_b200.h"
#include "pmb_registry.hpp"

#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

PMB_DECLARE_PROBLEM(mobile_robot_6x2)
PMB_DECLARE_PROBLEM(mobile_robot_5x2)
PMB_DECLARE_PROBLEM(mobile_robot_5x3)
PMB_DECLARE_PROBLEM(cstr_5x2)
PMB_DECLARE_PROBLEM(kite_12x1)
PMB_DECLARE_PROBLEM(kite_4x2)
PMB_DECLARE_PROBLEM(robot_obstacle_5x2)
PMB_DECLARE_PROBLEM(parking_5x2)

namespace pmb {
namespace {

#define PMB_FAIL(code, msg) do { last_error_string() = (msg); return (code); } while (0)

struct Registry { const char* name; IProblem* (*make)(); };
#define PMB_REG(NAME, ID) { NAME, &pmb_make_##ID }
const Registry g_builtin[] = {
    PMB_REG("mobile_robot_6x2", mobile_robot_6x2),   // BASELINE.json configs 1, 2, 5
    PMB_REG("mobile_robot_5x2", mobile_robot_5x2),   // reference CasADi fixture / continuous_ocp_test.cpp
    PMB_REG("mobile_robot_5x3", mobile_robot_5x3),   // reference mpc_wrapper_test.cpp
    PMB_REG("cstr_5x2", cstr_5x2),                   // reference cstr_control_test.cpp, BASELINE.json config 3
    PMB_REG("kite_12x1", kite_12x1),                 // BASELINE.json config 4 (our model)
    PMB_REG("kite_4x2", kite_4x2),                   // small kite variant for fast parity tests
    PMB_REG("robot_obstacle_5x2", robot_obstacle_5x2),   // NG = 1: generic inequality constraints
    PMB_REG("parking_5x2", parking_5x2),             // NP = 1: reference minimal_time_test.cpp / dense_sparse_compare.cpp
};
/** built-in problems followed by the ones user translation units registered (pmb_register_problem); entries are never
 *  removed, names are owned by the registry (std::deque-like stability through unique_ptr) */
struct RegistryTable {
    std::vector<std::unique_ptr<std::string>> names;
    std::vector<Registry> rows;
    std::mutex mu;
    RegistryTable() { for (const Registry& r : g_builtin) rows.push_back(r); }
};
RegistryTable& table() { static RegistryTable t; return t; }
int registry_size() { RegistryTable& t = table(); std::lock_guard<std::mutex> g(t.mu); return (int)t.rows.size(); }
bool registry_row(int i, Registry* out)
{
    RegistryTable& t = table(); std::lock_guard<std::mutex> g(t.mu);
    if (i < 0 || i >= (int)t.rows.size()) return false;
    *out = t.rows[i];
    return true;
}
/** copies the row out: the vector may grow under a concurrent registration */
bool find_problem(const char* name, Registry* out)
{
    if (!name) return false;
    RegistryTable& t = table(); std::lock_guard<std::mutex> g(t.mu);
    for (const Registry& r : t.rows) if (std::strcmp(r.name, name) == 0) { *out = r; return true; }
    return false;
}

bool have_device() { return rt_device_count() > 0; }

/** scoped set of device buffers filled from / drained to host arrays */
struct Staging {
    stream_t s = nullptr;
    std::vector<void*> bufs;
    bool ok = true;
    ~Staging() { for (void* p : bufs) rt_free(p); }
    template <class T> T* in(const T* host, size_t count)
    {
        if (!host) return nullptr;
        T* p = (T*)rt_alloc(count * sizeof(T));
        if (!p) { ok = false; return nullptr; }
        bufs.push_back(p);
        ok = rt_h2d(p, host, count * sizeof(T), s) && ok;
        return p;
    }
    template <class T> T* out(const T* host_flag, size_t count)
    {
        if (!host_flag) return nullptr;
        T* p = (T*)rt_alloc(count * sizeof(T));
        if (!p) { ok = false; return nullptr; }
        bufs.push_back(p);
        return p;
    }
    template <class T> void back(T* host, const T* dev, size_t count) { if (host && dev) ok = rt_d2h(host, dev, count * sizeof(T), s) && ok; }
};

constexpr size_t SMEM_CTA_MAX = 227 * 1024;   // B200: opt-in dynamic shared memory per CTA

/** launches the persistent boxADMM kernel over `batch` instances; `queue` is a zeroed device counter */
template <int R, bool IN_SMEM, bool FAST = false> bool launch_qp_r(size_t fac_doubles, size_t vec_bytes, stream_t s, const pmb_qp_settings_t& st, const QpBatch& qb,
                                                int batch, int* queue, DevBuf<double>& scratch)
{
    using Body = QpBody<R, IN_SMEM, FAST>;
    const size_t base = Cta::SCRATCH_DOUBLES * sizeof(double) + vec_bytes;
    const size_t smem = IN_SMEM ? base + fac_doubles * sizeof(double) : base;
    int grid = resident_ctas<Body, pmb_qp_settings_t, QpBatch, FactorStore, int, int*>(smem, st, qb, FactorStore{}, 0, (int*)nullptr);
    if (grid <= 0) { last_error_string() = "qp_box_admm: kernel does not fit on the device"; return false; }
    if (!IN_SMEM) grid = grid > 2 * 148 ? 2 * 148 : grid;      // keep the global factor slots L2 resident
    if (grid > batch) grid = batch;
    FactorStore fs{nullptr, fac_doubles, rt_sm_count()};
    if (!IN_SMEM) { if (!scratch.resize((size_t)grid * fac_doubles)) return false; fs.global = scratch.p; }
    return rt_launch<Body>(grid, smem, s, st, qb, fs, batch, queue);
}

/** process-wide default of the arithmetic mode (pmb_set_default_arithmetic): used by pmb_qp_solve and by new SQP handles */
std::atomic<int> g_default_arithmetic{PMB_ARITH_EXACT};

bool launch_qp(stream_t s, const pmb_qp_settings_t& st, const QpBatch& qb, int batch, int* queue, DevBuf<double>& scratch)
{
    const int n = qb.N + qb.M;
    const int R = (n + 31) / 32;
    if (g_default_arithmetic.load() == PMB_ARITH_FAST) {
        // fast arithmetic: tile workspace in shared memory (pmb_qp_fast.hpp); sizes whose workspace does not fit are refused
        const size_t fd = fast::workspace_doubles(n), vb = qp_vec_bytes(qb.N, qb.M);
        if (Cta::SCRATCH_DOUBLES * sizeof(double) + vb + fd * sizeof(double) > SMEM_CTA_MAX) {
            last_error_string() = "qp_solve: fast arithmetic needs the tile workspace in shared memory (N + M <= ~220)"; return false;
        }
        switch (R) {
#define PMB_QP_FAST(r) case r: return launch_qp_r<r, true, true>(fd, vb, s, st, qb, batch, queue, scratch);
        PMB_QP_FAST(1) PMB_QP_FAST(2) PMB_QP_FAST(3) PMB_QP_FAST(4) PMB_QP_FAST(5) PMB_QP_FAST(6) PMB_QP_FAST(7)
#undef PMB_QP_FAST
        default: break;
        }
        last_error_string() = "qp_solve: fast arithmetic is not instantiated for this size";
        return false;
    }
    const size_t fd = qp_factor_doubles(qb.N, qb.M), vb = qp_vec_bytes(qb.N, qb.M);
    if (Cta::SCRATCH_DOUBLES * sizeof(double) + vb > SMEM_CTA_MAX) { last_error_string() = "QP vectors do not fit in shared memory"; return false; }
    const bool in_smem = Cta::SCRATCH_DOUBLES * sizeof(double) + vb + fd * sizeof(double) <= SMEM_CTA_MAX;
    switch (R) {
    // the factor of every QP with n <= 192 fits in shared memory; n in (192, 224] may or may not; beyond that it never does
#define PMB_QP_SMEM(r) case r: return launch_qp_r<r, true>(fd, vb, s, st, qb, batch, queue, scratch);
#define PMB_QP_GLOB(r) case r: return launch_qp_r<r, false>(fd, vb, s, st, qb, batch, queue, scratch);
    PMB_QP_SMEM(1) PMB_QP_SMEM(2) PMB_QP_SMEM(3) PMB_QP_SMEM(4) PMB_QP_SMEM(5) PMB_QP_SMEM(6)
    case 7: return in_smem ? launch_qp_r<7, true>(fd, vb, s, st, qb, batch, queue, scratch) : launch_qp_r<7, false>(fd, vb, s, st, qb, batch, queue, scratch);
    PMB_QP_GLOB(8) PMB_QP_GLOB(9) PMB_QP_GLOB(10) PMB_QP_GLOB(11) PMB_QP_GLOB(12)
#undef PMB_QP_SMEM
#undef PMB_QP_GLOB
    default: break;
    }
    last_error_string() = "QP dimension N+M > 384 not instantiated";
    return false;
}

template <int R, bool IN_SMEM> bool launch_admm_r(size_t fac_doubles, size_t vec_bytes, stream_t s, const pmb_qp_settings_t& st, const AdmmBatch& qb,
                                                  int batch, int* queue, DevBuf<double>& scratch)
{
    using Body = QpAdmmBody<R, IN_SMEM>;
    const size_t base = Cta::SCRATCH_DOUBLES * sizeof(double) + vec_bytes;
    const size_t smem = IN_SMEM ? base + fac_doubles * sizeof(double) : base;
    int grid = resident_ctas<Body, pmb_qp_settings_t, AdmmBatch, FactorStore, int, int*>(smem, st, qb, FactorStore{}, 0, (int*)nullptr);
    if (grid <= 0) { last_error_string() = "qp_solve_admm: kernel does not fit on the device"; return false; }
    if (!IN_SMEM) grid = grid > 2 * 148 ? 2 * 148 : grid;
    if (grid > batch) grid = batch;
    FactorStore fs{nullptr, fac_doubles, rt_sm_count()};
    if (!IN_SMEM) { if (!scratch.resize((size_t)grid * fac_doubles)) return false; fs.global = scratch.p; }
    return rt_launch<Body>(grid, smem, s, st, qb, fs, batch, queue);
}

bool launch_admm(stream_t s, const pmb_qp_settings_t& st, const AdmmBatch& qb, int batch, int* queue, DevBuf<double>& scratch)
{
    const int n = 2 * qb.N + qb.M;
    const int R = (n + 31) / 32;
    const size_t fd = (size_t)n * (n + 1) / 2, vb = admm_vec_bytes(qb.N, qb.M);
    const bool in_smem = Cta::SCRATCH_DOUBLES * sizeof(double) + vb + fd * sizeof(double) <= SMEM_CTA_MAX;
    switch (R) {
#define PMB_ADMM_SMEM(r) case r: return launch_admm_r<r, true>(fd, vb, s, st, qb, batch, queue, scratch);
    PMB_ADMM_SMEM(1) PMB_ADMM_SMEM(2) PMB_ADMM_SMEM(3) PMB_ADMM_SMEM(4) PMB_ADMM_SMEM(5) PMB_ADMM_SMEM(6)
    case 7: return in_smem ? launch_admm_r<7, true>(fd, vb, s, st, qb, batch, queue, scratch) : launch_admm_r<7, false>(fd, vb, s, st, qb, batch, queue, scratch);
    case 8: return launch_admm_r<8, false>(fd, vb, s, st, qb, batch, queue, scratch);
#undef PMB_ADMM_SMEM
    default: break;
    }
    last_error_string() = "qp_solve_admm: KKT dimension 2N + M > 256 is not instantiated";
    return false;
}

void sqp_defaults(pmb_sqp_settings_t* s)
{ s->tau = 0.5; s->eta = 0.25; s->rho = 0.5; s->eps_prim = 1e-3; s->eps_dual = 1e-3; s->max_iter = 100; s->line_search_max_iter = 100; }
void qp_defaults(pmb_qp_settings_t* s)
{
    s->eps_rel = 1e-3; s->eps_abs = 1e-3; s->max_iter = 1000; s->warm_start = 0; s->reuse_pattern = 0; s->verbose = 0;
    s->rho = 1e-1; s->sigma = 1e-6; s->alpha = 1.0; s->check_termination = 25; s->adaptive_rho = 0; s->adaptive_rho_tolerance = 5;
    s->adaptive_rho_interval = 25; s->_pad = 0;
}

} // namespace
} // namespace pmb

struct pmb_ocp { pmb::IProblem* impl = nullptr; bool owned = true; };

struct pmb_sqp {
    pmb_ocp ocp;
    int batch = 0, device = 0;
    pmb_sqp_settings_t settings;
    pmb_qp_settings_t qp_settings;
    int opt_exact_hessian = 0, opt_gershgorin = 0;   // pmb_sqp_set_hessian_options
    int opt_block_bfgs = 0;                          // pmb_sqp_set_hessian_update
    int opt_precond = PMB_PRECOND_IDENTITY;          // pmb_sqp_set_preconditioner
    int opt_line_search = PMB_LS_L1_MERIT;           // pmb_sqp_set_line_search
    int filter_depth = 10;
    double filter_beta = 1e-5;
    pmb::DevBuf<double> ruiz, filter;                // allocated when the option is switched on
    int qp_solver = PMB_QP_BOX_ADMM;                 // pmb_sqp_set_qp_solver
    int grid_admm = 0;
    pmb::DevBuf<double> factor_scratch_admm, Ae;
    int arithmetic = PMB_ARITH_EXACT;                // pmb_sqp_set_arithmetic
    int schedule = PMB_SCHEDULE_LPT_HISTORY;         // pmb_sqp_set_schedule
    bool have_history = false;                       // info holds the iteration counts of a completed solve of this handle
    pmb::DevBuf<int> order;
    int grid_fast = 0;
    bool trace_on = false;                           // pmb_sqp_set_trace
    pmb::stream_t own_stream = nullptr, stream = nullptr;
    pmb::event_t ev0 = nullptr, ev1 = nullptr;
    pmb::DevBuf<double> x_guess, lam_guess;   // device copies of the initial guess (pmb_sqp_reset_guess)
    pmb::DevBuf<double> x, lam, lam_k, H, A, h, al, au, lx, ux, lbx, ubx, lbg, ubg, d, lag_grad, step_prev, p, plam, stats, tr_alpha;
    pmb::DevBuf<pmb_sqp_info_t> info;
    pmb::DevBuf<pmb_qp_info_t> qp_info;
    pmb::DevBuf<int> qp_nfac, tr_qp_iter, tr_bfgs, tr_ls, tr_qp_factor, queue;
    pmb::DevBuf<double> factor_scratch, factor_scratch_fast;
    int grid = 0;
    bool factor_in_smem = true;
    int trace_rows = 0;
    double last_ms = 0;
    long long last_launches = 0;
    bool profiling = false;
    pmb::event_t kev0 = nullptr, kev1 = nullptr;
    pmb::DevBuf<unsigned long long> phase;   // per-phase SM cycles of the fused kernel (profiling)
    double last_kernel_ms = 0;
    unsigned long long phase_host[16] = {0};
    double k_ms[3] = {0, 0, 0};
    long long k_launches[3] = {0, 0, 0};
    ~pmb_sqp()
    {
        pmb::rt_set_device(device);
        pmb::rt_event_destroy(ev0); pmb::rt_event_destroy(ev1);
        pmb::rt_event_destroy(kev0); pmb::rt_event_destroy(kev1);
        pmb::rt_stream_destroy(own_stream);
        delete ocp.impl;
    }
};

namespace pmb {
extern "C" {

const char* pmb_version(void)
{
#ifdef PMB_EMU
    return "polympc-b200 0.1 (warp-emulator build: TEST INFRASTRUCTURE)";
#else
    return "polympc-b200 0.1 (sm_100a)";
#endif
}
const char* pmb_last_error(void) { return last_error_string().c_str(); }
int pmb_device_count(void) { return rt_device_count(); }
int pmb_problem_count(void) { return registry_size(); }
const char* pmb_problem_name(int i) { Registry r; return registry_row(i, &r) ? r.name : nullptr; }
int pmb_register_problem(const char* name, void* (*factory)(void))
{
    if (!name || !*name || !factory) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "pmb_register_problem: null name or factory");
    RegistryTable& t = table();
    std::lock_guard<std::mutex> g(t.mu);
    for (const Registry& r : t.rows) if (std::strcmp(r.name, name) == 0) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "pmb_register_problem: name already registered");
    t.names.emplace_back(new std::string(name));
    t.rows.push_back(Registry{t.names.back()->c_str(), reinterpret_cast<IProblem* (*)()>(factory)});
    return PMB_OK;
}
int pmb_problem_dims(const char* name, pmb_dims_t* out)
{
    Registry reg; const Registry* r = find_problem(name, &reg) ? &reg : nullptr;
    if (!r) PMB_FAIL(PMB_ERR_UNKNOWN_PROBLEM, "unknown problem");
    if (!out) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null output");
    std::unique_ptr<IProblem> p(r->make());
    *out = p->dims;
    return PMB_OK;
}
void pmb_qp_default_settings(pmb_qp_settings_t* s) { if (s) qp_defaults(s); }
void pmb_sqp_default_settings(pmb_sqp_settings_t* s) { if (s) sqp_defaults(s); }
void pmb_sqp_default_qp_settings(pmb_qp_settings_t* s)
{
    if (!s) return;
    qp_defaults(s);   // then the SQPBase constructor overrides (sqp_base.hpp:83-90)
    s->warm_start = 0; s->check_termination = 10; s->eps_abs = 1e-4; s->eps_rel = 1e-4; s->max_iter = 100;
    s->adaptive_rho = 1; s->adaptive_rho_interval = 50; s->alpha = 1.0;
}

struct DmEvalBody {
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr const char* NAME = "dm_eval";
    static constexpr size_t EMU_STACK_BYTES = 256u << 10;
    PMB_DEV static void run(const Warp& w, int blk, unsigned char*, int fn, int n, const double* x, const double* y, double* out)
    {
        const int e = blk * THREADS + w.tid();
        if (e < n) out[e] = dm::dm_dispatch(fn, x[e], y ? y[e] : 0.0);
    }
};

int pmb_dm_eval(int fn, int n, const double* x, const double* y, double* out)
{
    if (fn < 0 || fn >= PMB_DM_COUNT || n < 0 || !x || !out) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "dm_eval: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (n == 0) return PMB_OK;
    Staging st;
    const double* dx = st.in(x, (size_t)n); const double* dy = st.in(y, (size_t)n);
    double* dout = st.out(out, (size_t)n);
    if (!st.ok) return PMB_ERR_CUDA;
    if (!rt_launch<DmEvalBody>((n + DmEvalBody::THREADS - 1) / DmEvalBody::THREADS, 0, st.s, fn, n, dx, dy, dout)) return PMB_ERR_CUDA;
    st.back(out, (const double*)dout, (size_t)n);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_cheb_tables(int P, double* nodes, double* D, double* w)
{
    if (P < 2 || P > 64 || !nodes || !D || !w) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "cheb_tables: bad argument");
    cheb_tables(P, nodes, D, w);
    return PMB_OK;
}

// ---- ContinuousOCP ---------------------------------------------------------------------------------------------
pmb_ocp_t* pmb_ocp_create(const char* name, int device)
{
    Registry reg; const Registry* r = find_problem(name, &reg) ? &reg : nullptr;
    if (!r) { last_error_string() = "unknown problem"; return nullptr; }
    pmb_ocp_t* h = new pmb_ocp_t();
    h->impl = r->make();
    h->impl->device = device;
    return h;
}
void pmb_ocp_destroy(pmb_ocp_t* h) { if (h && h->owned) { delete h->impl; delete h; } }   // a handle borrowed from pmb_sqp_problem() is left alone
int pmb_ocp_dims(const pmb_ocp_t* h, pmb_dims_t* out) { if (!h || !out) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); *out = h->impl->dims; return PMB_OK; }
int pmb_ocp_set_params(pmb_ocp_t* h, const double* v, int n)
{ if (!h || !v || n != h->impl->dims.NPARAM) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_params: bad argument"); h->impl->set_params(v); return PMB_OK; }
int pmb_ocp_get_params(const pmb_ocp_t* h, double* v, int n)
{ if (!h || !v || n != h->impl->dims.NPARAM) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "get_params: bad argument"); h->impl->get_params(v); return PMB_OK; }
int pmb_ocp_set_time_limits(pmb_ocp_t* h, double t0, double tf) { if (!h) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); h->impl->set_time_limits(t0, tf); return PMB_OK; }
int pmb_ocp_time_nodes(const pmb_ocp_t* h, double* t) { if (!h || !t) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); h->impl->time_nodes(t); return PMB_OK; }

static int ocp_eval_host(pmb_ocp_t* h, int mode, int batch, const double* var, const double* d, const double* lam, double* cost, double* c,
                         double* g, double* jac, double* grad, double* hess, double* lag_grad)
{
    if (!h || batch < 0 || !var) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ocp: bad argument");
    const pmb_dims_t& D = h->impl->dims;
    if (D.ND > 0 && !d) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ocp: static parameters required");
    // outputs every kernel of this mode writes unconditionally (only `cost` is optional)
    {
        const bool need_c = mode == OCP_EQ || mode == OCP_EQ_LIN || mode == OCP_LAG_GRAD || mode == OCP_LAG_GRAD_HESS;
        const bool need_jac = mode == OCP_EQ_LIN || mode == OCP_LAG_GRAD || mode == OCP_LAG_GRAD_HESS;
        const bool need_grad = mode == OCP_COST_GRAD || mode == OCP_COST_GRAD_HESS || mode == OCP_LAG_GRAD || mode == OCP_LAG_GRAD_HESS;
        const bool need_hess = mode == OCP_COST_GRAD_HESS || mode == OCP_LAG_GRAD_HESS;
        const bool need_lg = mode == OCP_LAG_GRAD || mode == OCP_LAG_GRAD_HESS;
        if ((need_c && !c) || (need_jac && !jac) || (need_grad && !grad) || (need_hess && !hess) || (need_lg && !lag_grad) ||
            (mode == OCP_INEQ && D.NG > 0 && !g) || (mode == OCP_COST && !cost))
            PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ocp: a required output pointer is NULL");
    }
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    if (!rt_set_device(h->impl->device)) return PMB_ERR_CUDA;
    const size_t B = batch, N = D.N, M = D.M, NE = (size_t)D.NX * D.NN, NI = (size_t)D.NG * D.NN;
    const bool full = (mode == OCP_LAG_GRAD || mode == OCP_LAG_GRAD_HESS);
    const size_t crow = full ? M : NE;
    Staging st;
    OcpIo io{};
    io.var = st.in(var, B * N);
    io.d = D.ND > 0 ? st.in(d, B * D.ND) : nullptr;
    io.lam = st.in(lam, B * D.DUAL);
    io.cost = st.out(cost, B);
    io.c = st.out(c, B * crow);
    io.g = NI > 0 ? st.out(g, B * NI) : nullptr;
    io.jac = st.out(jac, B * crow * N);
    io.grad = st.out(grad, B * N);
    io.hess = st.out(hess, B * N * N);
    io.lag_grad = st.out(lag_grad, B * N);
    if (!st.ok) return PMB_ERR_CUDA;
    if (mode == OCP_INEQ && NI == 0) return PMB_OK;
    if (!h->impl->launch_eval(mode, batch, io, st.s)) return PMB_ERR_CUDA;
    st.back(cost, io.cost, B); st.back(c, io.c, B * crow); if (NI > 0) st.back(g, io.g, B * NI);
    st.back(jac, io.jac, B * crow * N); st.back(grad, io.grad, B * N); st.back(hess, io.hess, B * N * N); st.back(lag_grad, io.lag_grad, B * N);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_ocp_cost(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost)
{ return ocp_eval_host(h, OCP_COST, batch, var, d, nullptr, cost, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); }
int pmb_ocp_equalities(pmb_ocp_t* h, int batch, const double* var, const double* d, double* c)
{ return ocp_eval_host(h, OCP_EQ, batch, var, d, nullptr, nullptr, c, nullptr, nullptr, nullptr, nullptr, nullptr); }
int pmb_ocp_inequalities(pmb_ocp_t* h, int batch, const double* var, const double* d, double* g)
{ return ocp_eval_host(h, OCP_INEQ, batch, var, d, nullptr, nullptr, nullptr, g, nullptr, nullptr, nullptr, nullptr); }
int pmb_ocp_equalities_linearised(pmb_ocp_t* h, int batch, const double* var, const double* d, double* c, double* jac)
{ return ocp_eval_host(h, OCP_EQ_LIN, batch, var, d, nullptr, nullptr, c, nullptr, jac, nullptr, nullptr, nullptr); }
int pmb_ocp_cost_gradient(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost, double* grad)
{ return ocp_eval_host(h, OCP_COST_GRAD, batch, var, d, nullptr, cost, nullptr, nullptr, nullptr, grad, nullptr, nullptr); }
int pmb_ocp_cost_gradient_hessian(pmb_ocp_t* h, int batch, const double* var, const double* d, double* cost, double* grad, double* hess)
{ return ocp_eval_host(h, OCP_COST_GRAD_HESS, batch, var, d, nullptr, cost, nullptr, nullptr, nullptr, grad, hess, nullptr); }
int pmb_ocp_lagrangian_gradient(pmb_ocp_t* h, int batch, const double* var, const double* d, const double* lam, double* cost,
                                double* lag_grad, double* cost_grad, double* g, double* jac)
{
    if (!lam) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "lagrangian_gradient: lam required");
    return ocp_eval_host(h, OCP_LAG_GRAD, batch, var, d, lam, cost, g, nullptr, jac, cost_grad, nullptr, lag_grad);
}
int pmb_ocp_lagrangian_gradient_hessian(pmb_ocp_t* h, int batch, const double* var, const double* d, const double* lam, double* cost,
                                        double* lag_grad, double* lag_hess, double* cost_grad, double* g, double* jac)
{
    if (!lam) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "lagrangian_gradient_hessian: lam required");
    return ocp_eval_host(h, OCP_LAG_GRAD_HESS, batch, var, d, lam, cost, g, nullptr, jac, cost_grad, lag_hess, lag_grad);
}

// ---- QP / KKT / BFGS operators ------------------------------------------------------------------------------------
int pmb_qp_solve(int N, int M, int batch, const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                 const double* xlb, const double* xub, const double* x_guess, const double* y_guess, const pmb_qp_settings_t* settings,
                 double* x, double* y, pmb_qp_info_t* info, double* z, double* q, int* perm, int* ctype, int* n_factor)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || (M > 0 && (!A || !Alb || !Aub)) || !xlb || !xub || !settings || !x || !y || !info)
        PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "qp_solve: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch, n = (size_t)N + M;
    Staging st;
    QpBatch qb{};
    qb.N = N; qb.M = M;
    qb.H = st.in(H, B * N * N); qb.h = st.in(h, B * N); qb.A = st.in(A, B * M * N); qb.Alb = st.in(Alb, B * M); qb.Aub = st.in(Aub, B * M);
    qb.xlb = st.in(xlb, B * N); qb.xub = st.in(xub, B * N); qb.xg = st.in(x_guess, B * N); qb.yg = st.in(y_guess, B * n);
    qb.x = st.out(x, B * N); qb.y = st.out(y, B * n); qb.info = st.out(info, B); qb.z = st.out(z, B * M); qb.q = st.out(q, B * N);
    qb.perm = st.out(perm, B * n); qb.ctype = st.out(ctype, B * n); qb.nfac = st.out(n_factor, B);
    DevBuf<int> queue;
    DevBuf<double> scratch;
    if (!st.ok || !queue.resize(1) || !rt_memset(queue.p, 0, sizeof(int), st.s)) return PMB_ERR_CUDA;
    if (!launch_qp(st.s, *settings, qb, batch, queue.p, scratch)) return PMB_ERR_CUDA;
    st.back(x, qb.x, B * N); st.back(y, qb.y, B * n); st.back(info, qb.info, B); st.back(z, qb.z, B * M); st.back(q, qb.q, B * N);
    st.back(perm, qb.perm, B * n); st.back(ctype, qb.ctype, B * n); st.back(n_factor, qb.nfac, B);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_qp_solve_admm(int N, int M, int batch, const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                      const double* xlb, const double* xub, const double* x_guess, const double* y_guess, const pmb_qp_settings_t* settings,
                      double* x, double* y, pmb_qp_info_t* info, double* z, int* perm, int* ctype, int* n_factor)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || (M > 0 && (!A || !Alb || !Aub)) || !xlb || !xub || !settings || !x || !y || !info)
        PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "qp_solve_admm: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch, Me = (size_t)N + M, n = (size_t)N + Me;
    // construct_A (admm.hpp:215-222): Ae = [A; I], column-major (M + N) x N per instance
    std::vector<double> Ae(B * Me * N, 0.0);
    for (size_t b = 0; b < B; ++b)
        for (size_t j = 0; j < (size_t)N; ++j) {
            double* col = Ae.data() + (b * N + j) * Me;
            for (size_t i = 0; i < (size_t)M; ++i) col[i] = A[(b * N + j) * M + i];
            col[M + j] = 1.0;
        }
    Staging st;
    AdmmBatch qb{};
    qb.N = N; qb.M = M;
    qb.H = st.in(H, B * N * N); qb.h = st.in(h, B * N); qb.Ae = st.in((const double*)Ae.data(), B * Me * N);
    qb.Alb = M ? st.in(Alb, B * M) : nullptr; qb.Aub = M ? st.in(Aub, B * M) : nullptr;
    qb.xlb = st.in(xlb, B * N); qb.xub = st.in(xub, B * N); qb.xg = st.in(x_guess, B * N); qb.yg = st.in(y_guess, B * Me);
    qb.x = st.out(x, B * N); qb.y = st.out(y, B * Me); qb.info = st.out(info, B); qb.z = st.out(z, B * Me);
    qb.perm = st.out(perm, B * n); qb.ctype = st.out(ctype, B * Me); qb.nfac = st.out(n_factor, B);
    DevBuf<int> queue;
    DevBuf<double> scratch;
    if (!st.ok || !queue.resize(1) || !rt_memset(queue.p, 0, sizeof(int), st.s)) return PMB_ERR_CUDA;
    if (!launch_admm(st.s, *settings, qb, batch, queue.p, scratch)) return PMB_ERR_CUDA;
    st.back(x, qb.x, B * N); st.back(y, qb.y, B * Me); st.back(info, qb.info, B); st.back(z, qb.z, B * Me);
    st.back(perm, qb.perm, B * n); st.back(ctype, qb.ctype, B * Me); st.back(n_factor, qb.nfac, B);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_kkt_assemble(int N, int M, int batch, const double* H, const double* A, const double* rho_box, const double* rho_inv, double sigma,
                     double* K)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || (M > 0 && (!A || !rho_inv)) || !rho_box || !K) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "kkt_assemble: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch, n = (size_t)N + M;
    Staging st;
    const double* dH = st.in(H, B * N * N); const double* dA = st.in(A, B * M * N);
    const double* drb = st.in(rho_box, B * N); const double* dri = st.in(rho_inv, B * M);
    double* dK = st.out(K, B * n * n);
    if (!st.ok) return PMB_ERR_CUDA;
    if (!rt_launch<KktDenseBody>(batch, 0, st.s, N, M, dH, dA, drb, dri, sigma, dK)) return PMB_ERR_CUDA;
    st.back(K, dK, B * n * n);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_kkt_assemble_dev(int N, int M, int batch, const double* H, const double* A, const double* rho_box, const double* rho_inv,
                         double sigma, double* K, void* cuda_stream)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !A || !rho_box || !rho_inv || !K) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "kkt_assemble_dev: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    return rt_launch<KktDenseBody>(batch, 0, (stream_t)cuda_stream, N, M, H, A, rho_box, rho_inv, sigma, K) ? PMB_OK : PMB_ERR_CUDA;
}

int pmb_bfgs_update(int N, int batch, double* Bm, const double* s, const double* y, int* branch)
{
    if (N <= 0 || batch < 0 || !Bm || !s || !y) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "bfgs_update: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch;
    Staging st;
    double* dB = st.in((const double*)Bm, B * N * N);
    const double* ds = st.in(s, B * N); const double* dy = st.in(y, B * N);
    int* dbr = st.out(branch, B);
    if (!st.ok) return PMB_ERR_CUDA;
    if (!rt_launch<BfgsBody>(batch, BfgsBody::smem_bytes(N), st.s, N, dB, ds, dy, dbr)) return PMB_ERR_CUDA;
    st.back(Bm, (const double*)dB, B * N * N); st.back(branch, (const int*)dbr, B);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_ruiz_equilibrate(int N, int M, int batch, int variant, double* H, double* h, double* A, double* Al, double* Au, double* l, double* u,
                         double* D, double* E, double* c)
{
    if (N <= 0 || M < 0 || batch < 0 || !H || !h || !A || !Al || !Au || !l || !u || !D || !E || !c) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ruiz_equilibrate: bad argument");
    if (variant != PMB_PRECOND_RUIZ_DENSE && variant != PMB_PRECOND_RUIZ_SPARSE) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ruiz_equilibrate: variant must be PMB_PRECOND_RUIZ_DENSE or _SPARSE");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch, S = (size_t)N + M + 1;
    Staging st;
    RuizArgs a{};
    a.H = st.in((const double*)H, B * N * N); a.h = st.in((const double*)h, B * N); a.A = M ? st.in((const double*)A, B * M * N) : st.out((const double*)A, 1);       // M == 0: never dereferenced, but not null
    a.Al = M ? st.in((const double*)Al, B * M) : st.out((const double*)Al, 1); a.Au = M ? st.in((const double*)Au, B * M) : st.out((const double*)Au, 1);
    a.l = st.in((const double*)l, B * N); a.u = st.in((const double*)u, B * N);
    a.st = st.out((const double*)D, B * S);
    if (!st.ok) return PMB_ERR_CUDA;
    if (!rt_launch<RuizComputeBody>(batch, RuizComputeBody::smem_bytes(N, M), st.s, N, M, variant, a)) return PMB_ERR_CUDA;
    std::vector<double> hs(B * S);
    st.back(H, (const double*)a.H, B * N * N); st.back(h, (const double*)a.h, B * N); st.back(A, (const double*)a.A, B * M * N);
    st.back(Al, (const double*)a.Al, B * M); st.back(Au, (const double*)a.Au, B * M); st.back(l, (const double*)a.l, B * N); st.back(u, (const double*)a.u, B * N);
    st.back(hs.data(), (const double*)a.st, B * S);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    for (size_t b = 0; b < B; ++b) {
        for (int k = 0; k < N; ++k) D[b * N + k] = hs[b * S + k];
        for (int k = 0; k < M; ++k) E[b * M + k] = hs[b * S + N + k];
        c[b] = hs[b * S + N + M];
    }
    return PMB_OK;
}

int pmb_ruiz_unscale(int N, int M, int batch, const double* D, const double* E, const double* c, double* H, double* h, double* A, double* Al,
                     double* Au, double* l, double* u, double* x, double* y)
{
    if (N <= 0 || M < 0 || batch < 0 || !D || !E || !c || !H || !h || !A || !Al || !Au || !l || !u) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ruiz_unscale: bad argument");
    if ((x == nullptr) != (y == nullptr)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "ruiz_unscale: give both x and y or neither");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    const size_t B = batch, S = (size_t)N + M + 1;
    std::vector<double> hs(B * S);
    for (size_t b = 0; b < B; ++b) {
        for (int k = 0; k < N; ++k) hs[b * S + k] = D[b * N + k];
        for (int k = 0; k < M; ++k) hs[b * S + N + k] = E[b * M + k];
        hs[b * S + N + M] = c[b];
    }
    Staging st;
    RuizArgs a{};
    a.H = st.in((const double*)H, B * N * N); a.h = st.in((const double*)h, B * N); a.A = M ? st.in((const double*)A, B * M * N) : st.out((const double*)A, 1);       // M == 0: never dereferenced, but not null
    a.Al = M ? st.in((const double*)Al, B * M) : st.out((const double*)Al, 1); a.Au = M ? st.in((const double*)Au, B * M) : st.out((const double*)Au, 1);
    a.l = st.in((const double*)l, B * N); a.u = st.in((const double*)u, B * N);
    a.st = st.in((const double*)hs.data(), B * S);
    a.x = st.in((const double*)x, B * N); a.y = st.in((const double*)y, B * ((size_t)N + M));
    if (!st.ok) return PMB_ERR_CUDA;
    if (!rt_launch<RuizUnscaleBody>(batch, RuizUnscaleBody::SMEM, st.s, N, M, a)) return PMB_ERR_CUDA;
    st.back(H, (const double*)a.H, B * N * N); st.back(h, (const double*)a.h, B * N); st.back(A, (const double*)a.A, B * M * N);
    st.back(Al, (const double*)a.Al, B * M); st.back(Au, (const double*)a.Au, B * M); st.back(l, (const double*)a.l, B * N); st.back(u, (const double*)a.u, B * N);
    st.back(x, (const double*)a.x, B * N); st.back(y, (const double*)a.y, B * ((size_t)N + M));
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

int pmb_ocp_block_bfgs_update(pmb_ocp_t* h, int batch, double* Bm, const double* s, const double* y, int* branch)
{
    if (!h || batch < 0 || !Bm || !s || !y) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "block_bfgs_update: bad argument");
    if (!have_device()) PMB_FAIL(PMB_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
    if (batch == 0) return PMB_OK;
    if (!rt_set_device(h->impl->device)) return PMB_ERR_CUDA;
    const size_t B = batch, N = h->impl->dims.N;
    Staging st;
    double* dB = st.in((const double*)Bm, B * N * N);
    const double* ds = st.in(s, B * N); const double* dy = st.in(y, B * N);
    int* dbr = st.out(branch, B);
    if (!st.ok) return PMB_ERR_CUDA;
    if (!h->impl->launch_block_bfgs(batch, dB, ds, dy, dbr, st.s)) return PMB_ERR_CUDA;
    st.back(Bm, (const double*)dB, B * N * N); st.back(branch, (const int*)dbr, B);
    if (!st.ok || !rt_sync(st.s)) return PMB_ERR_CUDA;
    return PMB_OK;
}

// ---- SQP -------------------------------------------------------------------------------------------------------
pmb_sqp_t* pmb_sqp_create(const char* name, int batch, int device)
{
    Registry reg; const Registry* r = find_problem(name, &reg) ? &reg : nullptr;
    if (!r || batch <= 0) { last_error_string() = "sqp_create: unknown problem or bad batch"; return nullptr; }
    if (!have_device()) { last_error_string() = "no CUDA device: the engine has no CPU fallback"; return nullptr; }
    if (device < 0 || device >= rt_device_count() || !rt_set_device(device)) { last_error_string() = "sqp_create: bad device"; return nullptr; }
    std::unique_ptr<pmb_sqp_t> s(new pmb_sqp_t());
    s->ocp.impl = r->make(); s->ocp.impl->device = device; s->ocp.owned = false;
    s->batch = batch; s->device = device;
    sqp_defaults(&s->settings);
    pmb_sqp_default_qp_settings(&s->qp_settings);
    const pmb_dims_t& D = s->ocp.impl->dims;
    const size_t B = batch, N = D.N, M = D.M, DU = D.DUAL, NI = (size_t)D.NG * D.NN, ND = D.ND;
    bool ok = s->x.resize(B * N) && s->lam.resize(B * DU) && s->x_guess.resize(B * N) && s->lam_guess.resize(B * DU) && s->lam_k.resize(B * DU) && s->H.resize(B * N * N) && s->A.resize(B * M * N) &&
              s->h.resize(B * N) && s->al.resize(B * M) && s->au.resize(B * M) && s->lx.resize(B * N) && s->ux.resize(B * N) &&
              s->lbx.resize(B * N) && s->ubx.resize(B * N) && s->lbg.resize(B * NI + 1) && s->ubg.resize(B * NI + 1) && s->d.resize(B * ND + 1) &&
              s->lag_grad.resize(B * N) && s->step_prev.resize(B * N) && s->p.resize(B * N) && s->plam.resize(B * DU) && s->stats.resize(B * 4) &&
              s->info.resize(B) && s->qp_info.resize(B) && s->qp_nfac.resize(B) && s->queue.resize(1);
    ok = ok && rt_stream_create(&s->own_stream) && rt_event_create(&s->ev0) && rt_event_create(&s->ev1) && rt_event_create(&s->kev0) &&
         rt_event_create(&s->kev1);
    if (!ok) return nullptr;
    s->stream = s->own_stream;
    {   // placement of the LDL^T factor and the persistent grid
        const IProblem& P = *s->ocp.impl;
        s->factor_in_smem = P.factor_in_smem();
        if (P.solve_smem_bytes() > SMEM_CTA_MAX) { last_error_string() = "sqp_create: problem too large for shared memory"; return nullptr; }
        int grid = P.solve_resident_ctas();
        if (grid <= 0) { last_error_string() = "sqp_create: sqp_solve kernel does not fit on the device"; return nullptr; }
        if (!s->factor_in_smem) grid = grid > 2 * 148 ? 2 * 148 : grid;
        if (grid > batch) grid = batch;
        s->grid = grid;
        if (!s->factor_in_smem && !s->factor_scratch.resize((size_t)grid * P.factor_doubles())) return nullptr;
    }
    // SQPBase constructor state (sqp_base.hpp:72-99): x = 0, lam = 0, bounds +-inf
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> lo(B * N, -INF), hi(B * N, INF);
    ok = rt_memset(s->x.p, 0, s->x.bytes(), s->stream) && rt_memset(s->lam.p, 0, s->lam.bytes(), s->stream) &&
         rt_memset(s->x_guess.p, 0, s->x_guess.bytes(), s->stream) && rt_memset(s->lam_guess.p, 0, s->lam_guess.bytes(), s->stream) &&
         rt_memset(s->d.p, 0, s->d.bytes(), s->stream) && rt_memset(s->stats.p, 0, s->stats.bytes(), s->stream) &&
         rt_memset(s->info.p, 0, s->info.bytes(), s->stream) &&
         rt_h2d(s->lbx.p, lo.data(), B * N * sizeof(double), s->stream) && rt_h2d(s->ubx.p, hi.data(), B * N * sizeof(double), s->stream);
    if (NI > 0) ok = ok && rt_h2d(s->lbg.p, lo.data(), B * NI * sizeof(double), s->stream) && rt_h2d(s->ubg.p, hi.data(), B * NI * sizeof(double), s->stream);
    ok = ok && rt_sync(s->stream);
    if (!ok) return nullptr;
    if (g_default_arithmetic.load() == PMB_ARITH_FAST) pmb_sqp_set_arithmetic(s.get(), PMB_ARITH_FAST);
    return s.release();
}
void pmb_sqp_destroy(pmb_sqp_t* s) { delete s; }
pmb_ocp_t* pmb_sqp_problem(pmb_sqp_t* s) { return s ? &s->ocp : nullptr; }
int pmb_sqp_batch(const pmb_sqp_t* s) { return s ? s->batch : (int)PMB_ERR_BAD_ARGUMENT; }
int pmb_sqp_set_settings(pmb_sqp_t* s, const pmb_sqp_settings_t* st) { if (!s || !st) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); s->settings = *st; return PMB_OK; }
int pmb_sqp_get_settings(const pmb_sqp_t* s, pmb_sqp_settings_t* st) { if (!s || !st) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); *st = s->settings; return PMB_OK; }
int pmb_sqp_set_qp_settings(pmb_sqp_t* s, const pmb_qp_settings_t* st) { if (!s || !st) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); s->qp_settings = *st; return PMB_OK; }
int pmb_sqp_get_qp_settings(const pmb_sqp_t* s, pmb_qp_settings_t* st) { if (!s || !st) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); *st = s->qp_settings; return PMB_OK; }
int pmb_sqp_set_hessian_options(pmb_sqp_t* s, int exact_every_iteration, int gershgorin_regularisation)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    s->opt_exact_hessian = exact_every_iteration != 0; s->opt_gershgorin = gershgorin_regularisation != 0;
    return PMB_OK;
}

int pmb_sqp_set_hessian_update(pmb_sqp_t* s, int mode)
{
    if (!s || (mode != PMB_HESSIAN_BFGS_DENSE && mode != PMB_HESSIAN_BFGS_BLOCK)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_hessian_update: bad argument");
    s->opt_block_bfgs = mode == PMB_HESSIAN_BFGS_BLOCK;
    return PMB_OK;
}
int pmb_sqp_set_preconditioner(pmb_sqp_t* s, int kind)
{
    if (!s || kind < PMB_PRECOND_IDENTITY || kind > PMB_PRECOND_RUIZ_SPARSE) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_preconditioner: bad argument");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const pmb_dims_t& d = s->ocp.impl->dims;
    if (kind != PMB_PRECOND_IDENTITY && !s->ruiz.resize((size_t)s->batch * (d.N + d.M + 1))) return PMB_ERR_CUDA;
    s->opt_precond = kind;
    return PMB_OK;
}

int pmb_sqp_set_line_search(pmb_sqp_t* s, int kind, double beta, int depth)
{
    if (!s || (kind != PMB_LS_L1_MERIT && kind != PMB_LS_FILTER)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_line_search: bad argument");
    if (kind == PMB_LS_FILTER && (depth < 1 || depth > PMB_FILTER_CAP || !(beta == beta)))
        PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_line_search: filter_max_depth must be in 1..PMB_FILTER_CAP and filter_beta a number");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    if (kind == PMB_LS_FILTER) {
        const size_t n = (size_t)s->batch * PMB_FILTER_DOUBLES;
        if (!s->filter.resize(n) || !rt_memset(s->filter.p, 0, n * sizeof(double), s->stream)) return PMB_ERR_CUDA;   // size 0.0 = empty
        s->filter_beta = beta; s->filter_depth = depth;
    }
    s->opt_line_search = kind;
    return PMB_OK;
}

int pmb_sqp_set_filter(pmb_sqp_t* s, const double* state, int stride)
{
    if (!s || !state || (stride != 0 && stride != PMB_FILTER_DOUBLES)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_filter: bad argument");
    if (s->opt_line_search != PMB_LS_FILTER) PMB_FAIL(PMB_ERR_UNSUPPORTED, "set_filter: the filter line search is not selected (pmb_sqp_set_line_search)");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const size_t B = s->batch;
    std::vector<double> tmp;
    const double* src = state;
    for (size_t b = 0; b < (stride ? B : 1); ++b) {
        const double n = state[b * PMB_FILTER_DOUBLES];
        if (!(n >= 0 && n <= PMB_FILTER_CAP && n == (double)(int)n)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_filter: size out of range");
    }
    if (stride == 0) {
        tmp.resize(B * PMB_FILTER_DOUBLES);
        for (size_t b = 0; b < B; ++b) std::memcpy(tmp.data() + b * PMB_FILTER_DOUBLES, state, PMB_FILTER_DOUBLES * sizeof(double));
        src = tmp.data();
    }
    return (rt_h2d(s->filter.p, src, B * PMB_FILTER_DOUBLES * sizeof(double), s->stream) && rt_sync(s->stream)) ? PMB_OK : PMB_ERR_CUDA;
}

int pmb_sqp_get_filter(const pmb_sqp_t* s, double* state)
{
    if (!s || !state) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    if (s->opt_line_search != PMB_LS_FILTER) PMB_FAIL(PMB_ERR_UNSUPPORTED, "get_filter: the filter line search is not selected (pmb_sqp_set_line_search)");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const size_t B = s->batch;
    if (!(rt_d2h(state, s->filter.p, B * PMB_FILTER_DOUBLES * sizeof(double), s->stream) && rt_sync(s->stream))) return PMB_ERR_CUDA;
    for (size_t b = 0; b < B; ++b) {            // entries beyond `size` are unspecified on the device: report zeros
        double* f = state + b * PMB_FILTER_DOUBLES;
        for (int k = (int)f[0]; k < PMB_FILTER_CAP; ++k) { f[1 + k] = 0.0; f[1 + PMB_FILTER_CAP + k] = 0.0; }
    }
    return PMB_OK;
}

int pmb_set_default_arithmetic(int mode)
{
    if (mode != PMB_ARITH_EXACT && mode != PMB_ARITH_FAST) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_default_arithmetic: bad mode");
    g_default_arithmetic.store(mode);
    return PMB_OK;
}
int pmb_get_default_arithmetic(void) { return g_default_arithmetic.load(); }
int pmb_sqp_set_arithmetic(pmb_sqp_t* s, int mode)
{
    if (!s || (mode != PMB_ARITH_EXACT && mode != PMB_ARITH_FAST)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_arithmetic: bad argument");
    if (mode == PMB_ARITH_FAST) {
        const IProblem& P = *s->ocp.impl;
        if (P.fast_smem_bytes() > SMEM_CTA_MAX) PMB_FAIL(PMB_ERR_UNSUPPORTED, "set_arithmetic: problem too large for shared memory");
        if (s->grid_fast == 0) {
            if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
            int grid = P.fast_resident_ctas();
            if (grid <= 0) PMB_FAIL(PMB_ERR_CUDA, "set_arithmetic: the fast sqp_solve kernel does not fit on the device");
            if (!P.fast_in_smem()) grid = grid > 2 * 148 ? 2 * 148 : grid;       // keep the global tile workspaces L2 resident
            grid = grid > s->batch ? s->batch : grid;
            if (!P.fast_in_smem() && !s->factor_scratch_fast.resize((size_t)grid * P.fast_factor_doubles())) return PMB_ERR_CUDA;
            s->grid_fast = grid;
        }
    }
    s->arithmetic = mode;
    return PMB_OK;
}
int pmb_sqp_set_qp_solver(pmb_sqp_t* s, int kind)
{
    if (!s || (kind != PMB_QP_BOX_ADMM && kind != PMB_QP_OSQP_ADMM)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_qp_solver: bad argument");
    if (kind == PMB_QP_OSQP_ADMM) {
        const IProblem& P = *s->ocp.impl;
        if (!P.admm_supported()) PMB_FAIL(PMB_ERR_UNSUPPORTED, "set_qp_solver: the OSQP-style ADMM is instantiated for 2N + M <= 256");
        if (P.admm_smem_bytes() > SMEM_CTA_MAX) PMB_FAIL(PMB_ERR_UNSUPPORTED, "set_qp_solver: problem too large for shared memory");
        if (s->grid_admm == 0) {
            if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
            int grid = P.admm_resident_ctas();
            if (grid <= 0) PMB_FAIL(PMB_ERR_CUDA, "set_qp_solver: the sqp_solve_osqp_admm kernel does not fit on the device");
            if (!P.admm_in_smem()) grid = grid > 2 * 148 ? 2 * 148 : grid;
            grid = grid > s->batch ? s->batch : grid;
            if (!P.admm_in_smem() && !s->factor_scratch_admm.resize((size_t)grid * P.admm_factor_doubles())) return PMB_ERR_CUDA;
            if (!s->Ae.resize((size_t)s->batch * (P.dims.N + P.dims.M) * P.dims.N)) return PMB_ERR_CUDA;
            s->grid_admm = grid;
        }
    }
    s->qp_solver = kind;
    return PMB_OK;
}
int pmb_sqp_set_schedule(pmb_sqp_t* s, int schedule)
{
    if (!s || (schedule != PMB_SCHEDULE_FIFO && schedule != PMB_SCHEDULE_LPT_HISTORY)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_schedule: bad argument");
    s->schedule = schedule;
    return PMB_OK;
}
int pmb_sqp_get_arithmetic(const pmb_sqp_t* s) { return s ? s->arithmetic : (int)PMB_ERR_BAD_ARGUMENT; }
int pmb_sqp_set_trace(pmb_sqp_t* s, int on) { if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); s->trace_on = on != 0; return PMB_OK; }

static int sqp_set_vec(pmb_sqp_t* s, double* dst, const double* v, int stride, size_t len)
{
    if (len == 0) return PMB_OK;
    if (!v || (stride != 0 && (size_t)stride != len)) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "sqp setter: bad pointer or stride");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const size_t B = s->batch;
    bool ok;
    if (stride == 0) {
        std::vector<double> tmp(B * len);
        for (size_t b = 0; b < B; ++b) std::memcpy(tmp.data() + b * len, v, len * sizeof(double));
        ok = rt_h2d(dst, tmp.data(), B * len * sizeof(double), s->stream) && rt_sync(s->stream);
    } else {
        ok = rt_h2d(dst, v, B * len * sizeof(double), s->stream) && rt_sync(s->stream);
    }
    return ok ? PMB_OK : PMB_ERR_CUDA;
}
int pmb_sqp_set_bounds_x(pmb_sqp_t* s, const double* lb, const double* ub, int stride)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    const size_t N = s->ocp.impl->dims.N;
    const int r = sqp_set_vec(s, s->lbx.p, lb, stride, N);
    return r ? r : sqp_set_vec(s, s->ubx.p, ub, stride, N);
}
int pmb_sqp_set_bounds_g(pmb_sqp_t* s, const double* lb, const double* ub, int stride)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    const size_t n = (size_t)s->ocp.impl->dims.NG * s->ocp.impl->dims.NN;
    const int r = sqp_set_vec(s, s->lbg.p, lb, stride, n);
    return r ? r : sqp_set_vec(s, s->ubg.p, ub, stride, n);
}
int pmb_sqp_set_parameters(pmb_sqp_t* s, const double* d, int stride)
{ if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); return sqp_set_vec(s, s->d.p, d, stride, s->ocp.impl->dims.ND); }
int pmb_sqp_set_primal(pmb_sqp_t* s, const double* x, int stride)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    const int r = sqp_set_vec(s, s->x.p, x, stride, s->ocp.impl->dims.N);
    if (r) return r;
    return (rt_d2d(s->x_guess.p, s->x.p, (size_t)s->batch * s->ocp.impl->dims.N * sizeof(double), s->stream) && rt_sync(s->stream)) ? PMB_OK : PMB_ERR_CUDA;
}
int pmb_sqp_set_dual(pmb_sqp_t* s, const double* l, int stride)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    const int r = sqp_set_vec(s, s->lam.p, l, stride, s->ocp.impl->dims.DUAL);
    if (r) return r;
    return (rt_d2d(s->lam_guess.p, s->lam.p, (size_t)s->batch * s->ocp.impl->dims.DUAL * sizeof(double), s->stream) && rt_sync(s->stream)) ? PMB_OK : PMB_ERR_CUDA;
}
int pmb_sqp_reset_guess(pmb_sqp_t* s)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const pmb_dims_t& D = s->ocp.impl->dims;
    const bool ok = rt_d2d(s->x.p, s->x_guess.p, (size_t)s->batch * D.N * sizeof(double), s->stream) &&
                    rt_d2d(s->lam.p, s->lam_guess.p, (size_t)s->batch * D.DUAL * sizeof(double), s->stream);
    return ok ? PMB_OK : PMB_ERR_CUDA;
}

/** strided 2-D copy helper for the initial-condition rows: dst[b*N + off + i] = src[b*NX + i] */
struct ScatterRowsBody {
    static constexpr int THREADS = 128;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr const char* NAME = "scatter_rows";
    static constexpr size_t EMU_STACK_BYTES = 128u << 10;
    PMB_DEV static void run(const Warp& w, int blk, unsigned char*, int batch, int len, int ld, int off, const double* src, double* dst)
    {
        const int e = blk * THREADS + w.tid();
        if (e < batch * len) { const int b = e / len, i = e - b * len; dst[(size_t)b * ld + off + i] = src[e]; }
    }
};

int pmb_sqp_set_initial_conditions(pmb_sqp_t* s, const double* x0_lb, const double* x0_ub)
{
    if (!s || !x0_lb || !x0_ub) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "set_initial_conditions: null");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const pmb_dims_t& D = s->ocp.impl->dims;
    const int B = s->batch, NX = D.NX, off = D.NX * D.NN - D.NX;
    DevBuf<double> tmp;
    if (!tmp.resize((size_t)2 * B * NX)) return PMB_ERR_CUDA;
    bool ok = rt_h2d(tmp.p, x0_lb, (size_t)B * NX * sizeof(double), s->stream) &&
              rt_h2d(tmp.p + (size_t)B * NX, x0_ub, (size_t)B * NX * sizeof(double), s->stream);
    const int grid = (B * NX + ScatterRowsBody::THREADS - 1) / ScatterRowsBody::THREADS;
    ok = ok && rt_launch<ScatterRowsBody>(grid, 0, s->stream, B, NX, D.N, off, (const double*)tmp.p, s->lbx.p);
    ok = ok && rt_launch<ScatterRowsBody>(grid, 0, s->stream, B, NX, D.N, off, (const double*)(tmp.p + (size_t)B * NX), s->ubx.p);
    ok = ok && rt_sync(s->stream);
    return ok ? PMB_OK : PMB_ERR_CUDA;
}

int pmb_sqp_solve(pmb_sqp_t* s)
{
    const int rc = pmb_sqp_solve_async(s);
    return rc != PMB_OK ? rc : pmb_sqp_wait(s);
}

int pmb_sqp_solve_async(pmb_sqp_t* s)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    if (s->qp_solver == PMB_QP_OSQP_ADMM && s->arithmetic == PMB_ARITH_FAST)
        PMB_FAIL(PMB_ERR_UNSUPPORTED, "sqp_solve: the OSQP-style ADMM runs in exact arithmetic only (pmb_sqp_set_arithmetic)");
    const int B = s->batch;
    const int rows = !s->trace_on ? 0 : (s->settings.max_iter > 0 ? s->settings.max_iter : 1);
    if (rows != s->trace_rows) {
        const size_t T = (size_t)B * rows;
        if (!(s->tr_qp_iter.resize(T) && s->tr_bfgs.resize(T) && s->tr_ls.resize(T) && s->tr_qp_factor.resize(T) && s->tr_alpha.resize(T))) return PMB_ERR_CUDA;
        s->trace_rows = rows;
    }
    stream_t st = s->stream;
    bool ok = rt_event_record(s->ev0, st);
    long long launches = 0;
    ok = ok && rt_memset(s->queue.p, 0, sizeof(int), st);
    if (rows > 0) {
        const size_t T = (size_t)B * rows;
        ok = ok && rt_memset(s->tr_qp_iter.p, 0xFF, T * sizeof(int), st) && rt_memset(s->tr_bfgs.p, 0xFF, T * sizeof(int), st) &&
             rt_memset(s->tr_ls.p, 0xFF, T * sizeof(int), st) && rt_memset(s->tr_qp_factor.p, 0xFF, T * sizeof(int), st) &&
             rt_memset(s->tr_alpha.p, 0xFF, T * sizeof(double), st);
    }
    SqpWs ws{};
    ws.x = s->x.p; ws.lam = s->lam.p; ws.lam_k = s->lam_k.p; ws.H = s->H.p; ws.A = s->A.p; ws.h = s->h.p; ws.al = s->al.p; ws.au = s->au.p;
    ws.lx = s->lx.p; ws.ux = s->ux.p; ws.lbx = s->lbx.p; ws.ubx = s->ubx.p; ws.lbg = s->lbg.p; ws.ubg = s->ubg.p; ws.d = s->d.p;
    ws.lag_grad = s->lag_grad.p; ws.step_prev = s->step_prev.p; ws.p = s->p.p; ws.plam = s->plam.p; ws.stats = s->stats.p;
    ws.info = s->info.p; ws.qp_info = s->qp_info.p; ws.qp_nfac = s->qp_nfac.p;
    if (rows > 0) { ws.tr_qp_iter = s->tr_qp_iter.p; ws.tr_bfgs = s->tr_bfgs.p; ws.tr_ls = s->tr_ls.p; ws.tr_qp_factor = s->tr_qp_factor.p; ws.tr_alpha = s->tr_alpha.p; }
    ws.trace_rows = rows;
    ws.opt_exact_hessian = s->opt_exact_hessian; ws.opt_gershgorin = s->opt_gershgorin; ws.opt_block_bfgs = s->opt_block_bfgs;
    ws.opt_precond = s->opt_precond; ws.opt_line_search = s->opt_line_search; ws.filter_depth = s->filter_depth; ws.filter_beta = s->filter_beta;
    ws.ruiz = s->ruiz.p; ws.filter = s->filter.p; ws.Ae = s->Ae.p;
    ws.order = nullptr;
    if (s->schedule == PMB_SCHEDULE_LPT_HISTORY && s->have_history && B > s->grid) {
        // longest-processing-time-first from the previous solve's iteration counts (still in s->info at this point of the stream)
        ok = ok && s->order.resize(B) && rt_launch<LptOrderBody>(1, LptOrderBody::SMEM, st, B, (const pmb_sqp_info_t*)s->info.p, s->order.p);
        ws.order = s->order.p;
        ++launches;
    }
    ws.phase = nullptr;
    if (s->profiling) {
        ok = ok && s->phase.resize(16) && rt_memset(s->phase.p, 0, 16 * sizeof(unsigned long long), st);
        ws.phase = s->phase.p;
    }
    // one persistent launch: CTAs draw instances from the queue and run their whole SQP loop on the device
    ok = ok && rt_event_record(s->kev0, st);
    if (s->qp_solver == PMB_QP_OSQP_ADMM) ok = ok && s->ocp.impl->launch_solve_admm(s->grid_admm, ws, s->settings, s->qp_settings, s->factor_scratch_admm.p, B, s->queue.p, st);
    else if (s->arithmetic == PMB_ARITH_FAST) ok = ok && s->ocp.impl->launch_solve_fast(s->grid_fast, ws, s->settings, s->qp_settings, s->factor_scratch_fast.p, B, s->queue.p, st);
    else ok = ok && s->ocp.impl->launch_solve(s->grid, ws, s->settings, s->qp_settings, s->factor_scratch.p, B, s->queue.p, st);
    ok = ok && rt_event_record(s->kev1, st);
    ++launches;
    ok = ok && rt_event_record(s->ev1, st);
    if (!ok) return PMB_ERR_CUDA;
    s->last_launches = launches;
    s->have_history = true;                 // stream order: the next solve's lpt_order kernel runs after this solve
    return PMB_OK;
}

int pmb_sqp_wait(pmb_sqp_t* s)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    stream_t st = s->stream;
    if (!rt_sync(st)) return PMB_ERR_CUDA;
    if (s->last_launches == 0) return PMB_OK;          // nothing was enqueued yet
    s->last_ms = rt_event_ms(s->ev0, s->ev1);
    s->last_kernel_ms = rt_event_ms(s->kev0, s->kev1);
    for (int k = 0; k < 3; ++k) { s->k_ms[k] = 0; s->k_launches[k] = 0; }
    if (s->profiling) {
        unsigned long long* ph = s->phase_host;
        if (!(rt_d2h(ph, s->phase.p, 16 * sizeof(unsigned long long), st) && rt_sync(st))) return PMB_ERR_CUDA;
        const double tot = (double)ph[0] + (double)ph[1] + (double)ph[2];
        for (int k = 0; k < 3; ++k) { s->k_ms[k] = tot > 0 ? s->last_kernel_ms * (double)ph[k] / tot : 0.0; s->k_launches[k] = (long long)ph[3]; }
    }
    return PMB_OK;
}

static int sqp_get(const pmb_sqp_t* s, void* host, const void* dev, size_t bytes)
{
    if (!s || !host) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    return (rt_d2h(host, dev, bytes, s->stream) && rt_sync(s->stream)) ? PMB_OK : PMB_ERR_CUDA;
}
int pmb_sqp_get_primal(const pmb_sqp_t* s, double* x) { return sqp_get(s, x, s ? s->x.p : nullptr, s ? (size_t)s->batch * s->ocp.impl->dims.N * sizeof(double) : 0); }
int pmb_sqp_get_dual(const pmb_sqp_t* s, double* l) { return sqp_get(s, l, s ? s->lam.p : nullptr, s ? (size_t)s->batch * s->ocp.impl->dims.DUAL * sizeof(double) : 0); }
int pmb_sqp_get_info(const pmb_sqp_t* s, pmb_sqp_info_t* info) { return sqp_get(s, info, s ? s->info.p : nullptr, s ? (size_t)s->batch * sizeof(pmb_sqp_info_t) : 0); }
int pmb_sqp_get_stats(const pmb_sqp_t* s, double* st) { return sqp_get(s, st, s ? s->stats.p : nullptr, s ? (size_t)s->batch * 4 * sizeof(double) : 0); }
int pmb_sqp_get_trace(const pmb_sqp_t* s, int rows, int* qi, double* al, int* bf, int* ls, int* qf)
{
    if (!s || rows <= 0) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "get_trace: bad argument");
    if (s->trace_rows == 0) PMB_FAIL(PMB_ERR_UNSUPPORTED, "get_trace: traces were off during the last solve (pmb_sqp_set_trace)");
    if (!rt_set_device(s->device)) return PMB_ERR_CUDA;
    const size_t B = s->batch, T = s->trace_rows;
    std::vector<int> ti(B * T); std::vector<double> td(B * T);
    auto pull_i = [&](int* out, const int* dev) -> bool {
        if (!out) return true;
        if (T > 0 && !(rt_d2h(ti.data(), dev, B * T * sizeof(int), s->stream) && rt_sync(s->stream))) return false;
        for (size_t b = 0; b < B; ++b) for (int r = 0; r < rows; ++r) out[b * rows + r] = (size_t)r < T ? ti[b * T + r] : -1;
        return true;
    };
    bool ok = pull_i(qi, s->tr_qp_iter.p) && pull_i(bf, s->tr_bfgs.p) && pull_i(ls, s->tr_ls.p) && pull_i(qf, s->tr_qp_factor.p);
    if (ok && al) {
        if (T > 0) ok = rt_d2h(td.data(), s->tr_alpha.p, B * T * sizeof(double), s->stream) && rt_sync(s->stream);
        for (size_t b = 0; b < B; ++b) for (int r = 0; r < rows; ++r) al[b * rows + r] = (size_t)r < T ? td[b * T + r] : std::nan("");
    }
    return ok ? PMB_OK : PMB_ERR_CUDA;
}
double pmb_sqp_last_solve_ms(const pmb_sqp_t* s) { return s ? s->last_ms : 0.0; }
double pmb_sqp_last_kernel_ms(const pmb_sqp_t* s) { return s ? s->last_kernel_ms : 0.0; }
long long pmb_sqp_last_solve_launches(const pmb_sqp_t* s) { return s ? s->last_launches : 0; }
int pmb_sqp_set_profiling(pmb_sqp_t* s, int on) { if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null"); s->profiling = on != 0; return PMB_OK; }
int pmb_sqp_get_kernel_times(const pmb_sqp_t* s, double* ms, long long* launches)
{
    if (!s || !ms || !launches) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    for (int k = 0; k < 3; ++k) { ms[k] = s->k_ms[k]; launches[k] = s->k_launches[k]; }
    return PMB_OK;
}
int pmb_sqp_get_phase_cycles(const pmb_sqp_t* s, unsigned long long* cycles16)
{
    if (!s || !cycles16) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    for (int k = 0; k < 16; ++k) cycles16[k] = s->phase_host[k];
    return PMB_OK;
}
int pmb_sqp_set_stream(pmb_sqp_t* s, void* cuda_stream)
{
    if (!s) PMB_FAIL(PMB_ERR_BAD_ARGUMENT, "null");
    s->stream = cuda_stream ? (stream_t)cuda_stream : s->own_stream;
    return PMB_OK;
}

} // extern "C"
} // namespace pmb
