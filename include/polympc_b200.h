/* polympc_b200.h — C ABI of the B200-native batched SQP engine (drop-in boundary for PolyMPC's collocated-NLP hot path).
 *
 * The reference (PREDICT-EPFL/polympc) is a header-only C++ template library with no FFI of its own; its plug-in seams
 * are C++ concepts.  Each entry point below replaces one of those seams for a *batch* of independent instances and cites
 * the reference interface it stands for (paths relative to the reference root).  Plain pointers and sizes only; all
 * arrays are caller-owned HOST memory unless the name ends in `_dev`; matrices are column-major (Eigen default);
 * batched arrays are instance-major (instance b at offset b*len).  Every function returns 0 on success and a negative
 * pmb_error_t otherwise; per-instance solver outcomes are reported through the info arrays, never through the return code
 * (the reference reports them through status enums as well: src/solvers/sqp_base.hpp:49-61, src/solvers/qp_base.hpp:55-72).
 *
 * The same signatures with the prefix `orc_` are exported by the CPU oracle (oracle/, test infrastructure only).
 */
#ifndef POLYMPC_B200_H
#define POLYMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pmb_error {
    PMB_OK = 0,
    PMB_ERR_UNKNOWN_PROBLEM = -1,   /* name not in the registry */
    PMB_ERR_BAD_ARGUMENT = -2,      /* null pointer, non-positive batch, size not instantiated ... */
    PMB_ERR_CUDA = -3,              /* a CUDA runtime call failed; see pmb_last_error() */
    PMB_ERR_NO_DEVICE = -4,         /* no usable CUDA device: the engine has no CPU fallback */
    PMB_ERR_UNSUPPORTED = -5
} pmb_error_t;

/* ---- problem sizes: enums of ContinuousOCP (src/control/continuous_ocp.hpp:69-98) -------------------------------- */
typedef struct pmb_dims {
    int NX, NU, NP, ND, NG;   /* polympc_traits<OCP> (continuous_ocp.hpp:23-33) */
    int P, S, NN;             /* POLY_ORDER, NUM_SEGMENTS, NUM_NODES = P*S+1 */
    int N;                    /* VAR_SIZE  = NX*NN + NU*NN + NP */
    int M;                    /* NUM_EQ + NUM_INEQ = NX*NN + NG*NN */
    int DUAL;                 /* DUAL_SIZE = M + N */
    int NPARAM;               /* number of doubles in the flattened data members of the problem class (Q, R, xs ...) */
} pmb_dims_t;

/* ---- sqp_settings_t / sqp_info_t (src/solvers/sqp_base.hpp:24-61) ----------------------------------------------- */
typedef struct pmb_sqp_settings {
    double tau, eta, rho, eps_prim, eps_dual;
    int max_iter, line_search_max_iter;
} pmb_sqp_settings_t;

typedef enum pmb_sqp_status { PMB_SQP_SOLVED = 0, PMB_SQP_MAX_ITER_EXCEEDED = 1, PMB_SQP_INVALID_SETTINGS = 2 } pmb_sqp_status_t;

typedef struct pmb_sqp_info {
    int iter;             /* number of QPs solved (starts at 1) */
    int qp_solver_iter;   /* accumulated ADMM iterations */
    int status;           /* pmb_sqp_status_t */
} pmb_sqp_info_t;

/* ---- qp_solver_settings_t / qp_solver_info_t / status_t (src/solvers/qp_base.hpp:17-72), ADMM-related fields ----- */
typedef struct pmb_qp_settings {
    double eps_rel, eps_abs;
    int max_iter, warm_start, reuse_pattern, verbose;
    double rho, sigma, alpha;
    int check_termination, adaptive_rho;
    double adaptive_rho_tolerance;
    int adaptive_rho_interval;
    int _pad;
} pmb_qp_settings_t;

typedef enum pmb_qp_status {
    PMB_QP_SOLVED = 0, PMB_QP_MAX_ITER_EXCEEDED = 1, PMB_QP_UNSOLVED = 2, PMB_QP_UNINITIALIZED = 3,
    PMB_QP_INFEASIBLE = 4, PMB_QP_INCONSISTENT = 5
} pmb_qp_status_t;

/* QPBase::constraint_type (qp_base.hpp:132-136) */
typedef enum pmb_constraint_type { PMB_INEQUALITY_CONSTRAINT = 0, PMB_EQUALITY_CONSTRAINT = 1, PMB_LOOSE_BOUNDS = 2 } pmb_constraint_type_t;

typedef struct pmb_qp_info {
    int status, iter, rho_updates, _pad;
    double rho_estimate, res_prim, res_dual;
} pmb_qp_info_t;

/* ---- library / registry ----------------------------------------------------------------------------------------- */
const char* pmb_version(void);
const char* pmb_last_error(void);
int pmb_device_count(void);
int pmb_problem_count(void);
const char* pmb_problem_name(int index);
int pmb_problem_dims(const char* name, pmb_dims_t* out);
/* Adds the kernels of a problem class compiled OUTSIDE the library (a user translation unit built by nvcc against
 * include/polympc_compat/ — the counterpart of instantiating SQPBase<Derived, Problem, QP> with a new Problem in the
 * header-only reference, sqp_base.hpp:64-70).  `factory` returns a new pmb::IProblem (pmb_registry.hpp) as void*.  The name
 * then works with pmb_ocp_create / pmb_sqp_create.  Re-registering an existing name fails with PMB_ERR_BAD_ARGUMENT. */
int pmb_register_problem(const char* name, void* (*factory)(void));

void pmb_qp_default_settings(pmb_qp_settings_t* s);      /* qp_base.hpp:17-53 defaults */
void pmb_sqp_default_settings(pmb_sqp_settings_t* s);    /* sqp_base.hpp:24-34 defaults */
void pmb_sqp_default_qp_settings(pmb_qp_settings_t* s);  /* defaults after the SQPBase ctor overrides, sqp_base.hpp:83-90 */

/* Deterministic fp64 elementary functions used by every problem functor and by the Chebyshev tables (csrc/pmb_detmath.h):
 * the same IEEE-only algorithms on host and device, so that decisions taken on their results are reproducible.  They stand
 * in for the std:: calls of the reference's AutoDiffScalar chain rules (src/autodiff/AutoDiffScalar.h:592-684).
 * out[i] = fn(x[i]) or fn(x[i], y[i]) evaluated ON THE DEVICE; y may be NULL for unary functions. */
typedef enum pmb_dm_fn { PMB_DM_SIN = 0, PMB_DM_COS, PMB_DM_TAN, PMB_DM_EXP, PMB_DM_LOG, PMB_DM_ATAN2, PMB_DM_ASIN, PMB_DM_ACOS,
                         PMB_DM_SINH, PMB_DM_COSH, PMB_DM_TANH, PMB_DM_POW, PMB_DM_SQRT, PMB_DM_COUNT } pmb_dm_fn_t;
int pmb_dm_eval(int fn, int n, const double* x, const double* y, double* out);

/* a1: Chebyshev<P, GAUSS_LOBATTO>::compute_nodes / compute_diff_matrix / compute_int_weights
 *     (src/polynomials/ebyshev.hpp:111-117, 198-214, 120-159).  nodes[P+1], D[(P+1)^2] column-major, w[P+1]. */
int pmb_cheb_tables(int P, double* nodes, double* D, double* w);

/* ---- ContinuousOCP: the "Problem" concept SQPBase consumes (continuous_ocp.hpp:430-647) -------------------------- */
typedef struct pmb_ocp pmb_ocp_t;
pmb_ocp_t* pmb_ocp_create(const char* problem_name, int device);
void pmb_ocp_destroy(pmb_ocp_t* ocp);
int pmb_ocp_dims(const pmb_ocp_t* ocp, pmb_dims_t* out);
int pmb_ocp_set_params(pmb_ocp_t* ocp, const double* values, int count);   /* data members of the problem class */
int pmb_ocp_get_params(const pmb_ocp_t* ocp, double* values, int count);
int pmb_ocp_set_time_limits(pmb_ocp_t* ocp, double t0, double tf);         /* continuous_ocp.hpp:147-159 */
int pmb_ocp_time_nodes(const pmb_ocp_t* ocp, double* time_nodes);          /* NN values, descending */

/* var[batch*N], d[batch*ND] (may be NULL when ND == 0), lam[batch*DUAL] */
int pmb_ocp_cost(pmb_ocp_t* ocp, int batch, const double* var, const double* d, double* cost);                    /* :1180-1207 */
int pmb_ocp_equalities(pmb_ocp_t* ocp, int batch, const double* var, const double* d, double* c);                 /* :738-766  */
int pmb_ocp_inequalities(pmb_ocp_t* ocp, int batch, const double* var, const double* d, double* g);               /* :769-782  */
int pmb_ocp_equalities_linearised(pmb_ocp_t* ocp, int batch, const double* var, const double* d,
                                  double* c, double* jac /* NUM_EQ x N */);                                        /* :794-878  */
int pmb_ocp_cost_gradient(pmb_ocp_t* ocp, int batch, const double* var, const double* d,
                          double* cost, double* grad);                                                             /* :1209-1249 */
int pmb_ocp_cost_gradient_hessian(pmb_ocp_t* ocp, int batch, const double* var, const double* d,
                                  double* cost, double* grad, double* hess /* N x N */);                           /* :1253-1367 */
int pmb_ocp_lagrangian_gradient(pmb_ocp_t* ocp, int batch, const double* var, const double* d, const double* lam,
                                double* cost, double* lag_grad, double* cost_grad,
                                double* g /* M */, double* jac /* M x N */);                                        /* :1957-1975 */
int pmb_ocp_lagrangian_gradient_hessian(pmb_ocp_t* ocp, int batch, const double* var, const double* d, const double* lam,
                                        double* cost, double* lag_grad, double* lag_hess /* N x N */, double* cost_grad,
                                        double* g, double* jac);                                                    /* :2097-2174 */

/* ContinuousOCP<..., SPARSE>::hessian_update_impl (continuous_ocp.hpp:2303-2431): block BFGS on the stored pattern of the
 * Lagrangian Hessian (per-node (x_k,u_k) blocks, parameter rows/columns).  B[batch*N*N] in/out (entries outside the pattern
 * are neither read nor written), s,y[batch*N]; branch[batch] (may be NULL): 0 plain, 1 damped. */
int pmb_ocp_block_bfgs_update(pmb_ocp_t* ocp, int batch, double* B, const double* s, const double* y, int* branch);

/* ---- QPBase<boxADMM<N,M,double,DENSE,LDLT,Lower>>::solve (qp_base.hpp:161-175, box_admm.hpp:81-205) --------------- */
/* H[batch*N*N], h[batch*N], A[batch*M*N], Alb/Aub[batch*M], xlb/xub[batch*N]; x_guess/y_guess may be NULL (cold start,
 * the 7-argument form).  Outputs: x[batch*N] = primal_solution(), y[batch*(M+N)] = dual_solution() = [y_A ; y_box].
 * Optional outputs (may be NULL): z[batch*M], q[batch*N] (ADMM splitting variables; the active set is
 * {i: z_i==Alb_i or z_i==Aub_i} U {j: q_j==xlb_j or q_j==xub_j}), perm[batch*(N+M)] (LDLT pivot permutation of the first
 * factorisation), ctype[batch*(M+N)] = [constr_type ; box_constr_type], n_factor[batch] (number of factorisations). */
int pmb_qp_solve(int N, int M, int batch,
                 const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                 const double* xlb, const double* xub, const double* x_guess, const double* y_guess,
                 const pmb_qp_settings_t* settings,
                 double* x, double* y, pmb_qp_info_t* info,
                 double* z, double* q, int* perm, int* ctype, int* n_factor);
/* ---- QPBase<ADMM<N,M,double,DENSE,LDLT,Lower>>::solve (qp_base.hpp:161-175, admm.hpp:112-213) ----------------------- */
/* The OSQP-style splitting: the box constraints are appended to A as identity rows (Ae = [A; I]), one auxiliary / multiplier
 * vector of size M + N, KKT system of size 2N + M.  Same inputs and settings as pmb_qp_solve.  Outputs: x[batch*N],
 * y[batch*(M+N)] = [y_A ; y_box]; optional (may be NULL): z[batch*(M+N)] = m_z (the active set is {i: z_i == bound_i}),
 * perm[batch*(2N+M)], ctype[batch*(M+N)], n_factor[batch].  Exact arithmetic only (bit-identical to the CPU oracle).
 * The same solver runs inside the fused SQP loop when selected with pmb_sqp_set_qp_solver(PMB_QP_OSQP_ADMM). */
int pmb_qp_solve_admm(int N, int M, int batch,
                      const double* H, const double* h, const double* A, const double* Alb, const double* Aub,
                      const double* xlb, const double* xub, const double* x_guess, const double* y_guess,
                      const pmb_qp_settings_t* settings,
                      double* x, double* y, pmb_qp_info_t* info,
                      double* z, int* perm, int* ctype, int* n_factor);

/* a17: boxADMM::construct_kkt_matrix, dense (box_admm.hpp:207-223).  K[batch*(N+M)^2] column-major; like the reference
 * only the lower triangle and the diagonal blocks are written (upper-right block is zero). */
int pmb_kkt_assemble(int N, int M, int batch, const double* H, const double* A, const double* rho_box,
                     const double* rho_inv, double sigma, double* K);
/* same operator on DEVICE pointers, asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream): no copies, no
 * synchronisation — the form a device-resident caller (or a bandwidth measurement) uses */
int pmb_kkt_assemble_dev(int N, int M, int batch, const double* H_dev, const double* A_dev, const double* rho_box_dev,
                         const double* rho_inv_dev, double sigma, double* K_dev, void* cuda_stream);

/* a12: BFGS_update (src/solvers/bfgs.hpp:23-52). B[batch*N*N] in/out, s,y[batch*N]; branch[batch]: 0 plain, 1 damped, 2 skipped */
int pmb_bfgs_update(int N, int batch, double* B, const double* s, const double* y, int* branch);

/* ---- SQPBase<..., ContinuousOCP<.., DENSE>, boxADMM, IdentityPreconditioner>::solve (sqp_base.hpp:568-696) -------- */
typedef struct pmb_sqp pmb_sqp_t;
pmb_sqp_t* pmb_sqp_create(const char* problem_name, int batch, int device);
void pmb_sqp_destroy(pmb_sqp_t* s);
pmb_ocp_t* pmb_sqp_problem(pmb_sqp_t* s);                                   /* SQPBase::get_problem(); BORROWED: owned by s, pmb_ocp_destroy ignores it */
int pmb_sqp_batch(const pmb_sqp_t* s);
int pmb_sqp_set_settings(pmb_sqp_t* s, const pmb_sqp_settings_t* st);       /* SQPBase::settings()    */
int pmb_sqp_get_settings(const pmb_sqp_t* s, pmb_sqp_settings_t* st);
int pmb_sqp_set_qp_settings(pmb_sqp_t* s, const pmb_qp_settings_t* st);     /* SQPBase::qp_settings() */
int pmb_sqp_get_qp_settings(const pmb_sqp_t* s, pmb_qp_settings_t* st);
/* The fixed menu of SQPBase CRTP hooks (sqp_base.hpp:198-350) a fused device loop can honour — the two overrides the
 * reference's own solvers install (tests/control/minimal_time_test.cpp:90-135):
 *   exact_every_iteration     update_linearisation_dense_impl := linearisation_dense_impl (exact Lagrangian Hessian at every
 *                             SQP iteration instead of the damped BFGS update, sqp_base.hpp:489-504);
 *   gershgorin_regularisation hessian_regularisation_dense_impl := for every column i with H_ii - r_i <= 0,
 *                             r_i = sum_{j != i} |H_ji|:  H_ii += (r_i - H_ii) + 0.01  (applied after each exact Hessian).
 * Both default to 0 = the reference defaults. */
int pmb_sqp_set_hessian_options(pmb_sqp_t* s, int exact_every_iteration, int gershgorin_regularisation);
/* hessian_update_impl (sqp_base.hpp:263-268), the quasi-Newton update used when the Hessian is not exact:
 *   PMB_HESSIAN_BFGS_DENSE  BFGS_update (src/solvers/bfgs.hpp:23-52) on the dense H — the SQPBase default;
 *   PMB_HESSIAN_BFGS_BLOCK  ContinuousOCP<..., SPARSE>::hessian_update_impl (src/control/continuous_ocp.hpp:2303-2431): the
 *                           "sparsity preserving block BFGS" on the per-node (x_k,u_k) blocks and the parameter rows/columns —
 *                           what every reference control test installs with
 *                           `this->problem.hessian_update_impl(hessian, x_step, grad_step)` on a SPARSE problem
 *                           (tests/control/mpc_wrapper_test.cpp:101-104, cstr_control_test.cpp:128-131 ...). */
typedef enum pmb_hessian_update { PMB_HESSIAN_BFGS_DENSE = 0, PMB_HESSIAN_BFGS_BLOCK = 1 } pmb_hessian_update_t;
int pmb_sqp_set_hessian_update(pmb_sqp_t* s, int mode);
/* Preconditioner template argument of SQPBase (sqp_base.hpp:605-611, 662-667): m_preconditioner.compute(H, h, A, al, au, lx, ux)
 * before the QP, unscale(p, p_lambda) and unscale(H, h, A, ...) after it, at every SQP iteration.
 *   PMB_PRECOND_IDENTITY     IdentityPreconditioner (qp_preconditioners.hpp:28-110) — the default;
 *   PMB_PRECOND_RUIZ_DENSE   RuizEquilibration<..., DENSE>::compute (qp_preconditioners.hpp:151-220): <= 4 passes of row/column
 *                            infinity-norm scaling + cost scaling gamma, zero guard = machine epsilon;
 *   PMB_PRECOND_RUIZ_SPARSE  RuizEquilibration<..., SPARSE>::compute (:236-300): the same passes with the zero guard 1e-4, no
 *                            guard on |h|_inf and A scaled as e_i * (A_ij * d_j) — what a SPARSE problem class selects
 *                            (tests/control/valet_parking_mpc_test.cpp:172). */
typedef enum pmb_preconditioner { PMB_PRECOND_IDENTITY = 0, PMB_PRECOND_RUIZ_DENSE = 1, PMB_PRECOND_RUIZ_SPARSE = 2 } pmb_preconditioner_t;
int pmb_sqp_set_preconditioner(pmb_sqp_t* s, int kind);
/* step_size_selection_impl (sqp_base.hpp:378-419):
 *   PMB_LS_L1_MERIT  the default backtracking on the l1 merit function;
 *   PMB_LS_FILTER    the filter line search the reference installs in tests/control/valet_parking_mpc_test.cpp:110-155 with
 *                    LSFilter (src/solvers/line_search.hpp:30-98): a trial point is taken when no filter entry (f_k, v_k)
 *                    has f_k - beta v_k <= cost and v_k - beta v_k <= violation; accepted points enter the filter (entries
 *                    they dominate leave it; at max_depth the oldest leaves).  The filter belongs to the solver object: it
 *                    survives from one solve to the next (pmb_sqp_set_filter / pmb_sqp_get_filter move it, this call empties it).
 * filter_max_depth in 1..PMB_FILTER_CAP. */
typedef enum pmb_line_search { PMB_LS_L1_MERIT = 0, PMB_LS_FILTER = 1 } pmb_line_search_t;
#define PMB_FILTER_CAP 16
#define PMB_FILTER_DOUBLES (1 + 2 * PMB_FILTER_CAP)   /* per instance: [size, cost[CAP], violation[CAP]], entry 0 = newest */
int pmb_sqp_set_line_search(pmb_sqp_t* s, int kind, double filter_beta, int filter_max_depth);
int pmb_sqp_set_filter(pmb_sqp_t* s, const double* state, int stride);   /* stride 0 broadcasts one state to the batch */
int pmb_sqp_get_filter(const pmb_sqp_t* s, double* state);
/* RuizEquilibration on host buffers, stand-alone (stage-wise parity): scales batch x {H[N*N], h[N], A[M*N], Al[M], Au[M], l[N],
 * u[N]} in place and returns D[N], E[M], c[1] per instance; pmb_ruiz_unscale applies RuizEquilibration::unscale to the QP data
 * and, when x / y are given, to a primal [N] / dual [M+N] solution (qp_preconditioners.hpp:364-404).  variant = pmb_preconditioner_t. */
int pmb_ruiz_equilibrate(int N, int M, int batch, int variant, double* H, double* h, double* A, double* Al, double* Au, double* l, double* u,
                         double* D, double* E, double* c);
int pmb_ruiz_unscale(int N, int M, int batch, const double* D, const double* E, const double* c, double* H, double* h, double* A, double* Al,
                     double* Au, double* l, double* u, double* x, double* y);
/* QPSolver template argument of SQPBase (sqp_base.hpp:64-70): the QP solver of the SQP loop.
 *   PMB_QP_BOX_ADMM   boxADMM<> (src/solvers/box_admm.hpp) — the reference's default; both arithmetics;
 *   PMB_QP_OSQP_ADMM  ADMM<> (src/solvers/admm.hpp:112-213), the OSQP-style splitting with the box rows appended to A (KKT
 *                     systems of size 2N + M); exact arithmetic only, instantiated for 2N + M <= 256 (PMB_ERR_UNSUPPORTED
 *                     beyond, and when the handle is in fast arithmetic at solve time). */
typedef enum pmb_qp_solver { PMB_QP_BOX_ADMM = 0, PMB_QP_OSQP_ADMM = 1 } pmb_qp_solver_t;
int pmb_sqp_set_qp_solver(pmb_sqp_t* s, int kind);
/* Arithmetic of the KKT linear algebra inside boxADMM (csrc/pmb_qp.hpp vs csrc/pmb_qp_fast.hpp).  Both run the same
 * algorithm with the same pivot permutation (Eigen::LDLT's diagonal rule):
 *   PMB_ARITH_EXACT  every fp64 operation in the order of the CPU oracle: results are bit-identical to it (default);
 *   PMB_ARITH_FAST   tile-blocked LDL^T on the fp64 tensor cores, explicit inverse of the unit-lower factor, triangular solves as
 *                    streaming mat-vecs: results agree with the oracle to rounding (<= 1e-10 rel-inf per SQP iteration; how
 *                    that propagates over a whole solve is measured in bench.py's `parity` object).
 * pmb_set_default_arithmetic sets the process-wide default (used by pmb_qp_solve and inherited by new SQP handles);
 * pmb_sqp_set_arithmetic switches one handle; PMB_ERR_UNSUPPORTED when the tile workspace does not fit in shared memory. */
typedef enum pmb_arithmetic { PMB_ARITH_EXACT = 0, PMB_ARITH_FAST = 1 } pmb_arithmetic_t;
int pmb_set_default_arithmetic(int mode);
int pmb_get_default_arithmetic(void);
int pmb_sqp_set_arithmetic(pmb_sqp_t* s, int mode);
int pmb_sqp_get_arithmetic(const pmb_sqp_t* s);
/* Order in which the persistent kernel's CTAs take instances off the work queue.  Results never depend on it.
 *   PMB_SCHEDULE_FIFO         instance index order;
 *   PMB_SCHEDULE_LPT_HISTORY  (default) from the second solve of a handle on: descending SQP iteration count of the PREVIOUS
 *                             solve (longest processing time first) — a fleet that is re-solved every control period keeps its
 *                             hard instances, so their long chains start at t = 0 instead of forming the tail of the batch. */
typedef enum pmb_schedule { PMB_SCHEDULE_FIFO = 0, PMB_SCHEDULE_LPT_HISTORY = 1 } pmb_schedule_t;
int pmb_sqp_set_schedule(pmb_sqp_t* s, int schedule);
/* per-iteration decision traces (pmb_sqp_get_trace) are recorded only when switched on before the solve (default off:
 * they cost batch x max_iter rows of device memory and five memsets per solve) */
int pmb_sqp_set_trace(pmb_sqp_t* s, int on);
/* stride == 0: one vector broadcast to every instance; stride == len: one vector per instance */
int pmb_sqp_set_bounds_x(pmb_sqp_t* s, const double* lbx, const double* ubx, int stride);   /* lower/upper_bound_x() */
int pmb_sqp_set_bounds_g(pmb_sqp_t* s, const double* lbg, const double* ubg, int stride);   /* lower/upper_bound_g() */
int pmb_sqp_set_parameters(pmb_sqp_t* s, const double* d, int stride);                      /* parameters()          */
int pmb_sqp_set_primal(pmb_sqp_t* s, const double* x, int stride);                          /* primal_solution() =   */
int pmb_sqp_set_dual(pmb_sqp_t* s, const double* lam, int stride);                          /* dual_solution() =     */
/* MPC::initial_conditions(x0) (src/control/mpc_wrapper.hpp:89-99): box equality on the LAST NX entries of the X block.
 * x0_lb/x0_ub[batch*NX] */
int pmb_sqp_set_initial_conditions(pmb_sqp_t* s, const double* x0_lb, const double* x0_ub);
/* MPC::x_guess()/u_guess()/lam_guess() kept on the device: restores the iterates to the values last given to
 * pmb_sqp_set_primal / pmb_sqp_set_dual (zeros after create, like the SQPBase ctor) without a host copy */
int pmb_sqp_reset_guess(pmb_sqp_t* s);
int pmb_sqp_solve(pmb_sqp_t* s);                                            /* SQPBase::solve(), whole batch */
/* The same, split: _async enqueues the whole batch solve on the handle's stream and returns; _wait blocks until it is done
 * (and makes pmb_sqp_last_solve_ms valid).  Two handles on two streams keep two batches in flight, which hides the straggler
 * tail of the persistent kernel.  Getters synchronise on the handle's stream by themselves. */
int pmb_sqp_solve_async(pmb_sqp_t* s);
int pmb_sqp_wait(pmb_sqp_t* s);
int pmb_sqp_get_primal(const pmb_sqp_t* s, double* x);                      /* [batch*N]    */
int pmb_sqp_get_dual(const pmb_sqp_t* s, double* lam);                      /* [batch*DUAL] */
int pmb_sqp_get_info(const pmb_sqp_t* s, pmb_sqp_info_t* info);             /* [batch]      */
/* stats[batch*4] = {cost(), primal_norm(), dual_norm(), constr_violation()} (sqp_base.hpp:192-195) */
int pmb_sqp_get_stats(const pmb_sqp_t* s, double* stats);
/* decision trace, row-major [batch][rows], rows <= max_iter; entries of iterations not executed are -1 / NaN.
 * qp_iter: ADMM trips of the QP; alpha: accepted step; bfgs: -1 exact Hessian, 0 plain, 1 damped, 2 skipped;
 * ls_trials: merit evaluations; qp_factor: LDLT factorisations in the QP. Any pointer may be NULL. */
int pmb_sqp_get_trace(const pmb_sqp_t* s, int rows, int* qp_iter, double* alpha, int* bfgs, int* ls_trials, int* qp_factor);
/* device time of the last pmb_sqp_solve in milliseconds (CUDA events on the engine's stream) and launches it issued */
double pmb_sqp_last_solve_ms(const pmb_sqp_t* s);
/* CUDA-event time of the fused sqp_solve launch alone inside the last completed solve (no queue reset, ordering kernel or copies). */
double pmb_sqp_last_kernel_ms(const pmb_sqp_t* s);
long long pmb_sqp_last_solve_launches(const pmb_sqp_t* s);
/* phase breakdown of the fused sqp_solve kernel: with profiling on, every CTA accumulates the SM cycles it spends in the
 * three phases of an SQP iteration; ms[3] = device time of the kernel (CUDA events on the engine's stream) attributed to
 * {linearise + BFGS, boxADMM QP, line search + step} in proportion to those cycles, launches[3] = number of SQP
 * iterations executed (all instances) */
int pmb_sqp_set_profiling(pmb_sqp_t* s, int on);
int pmb_sqp_get_kernel_times(const pmb_sqp_t* s, double* ms, long long* launches);
/* raw counters behind it, cycles16[16]: SM cycles (thread 0 of every CTA, summed) in {linearise, QP, step}, [3] = SQP
 * iterations, [4..9] = QP split {pivot order, gather, factorisation, triangular solves, ADMM updates, residual checks},
 * [10] = ADMM trips; the rest is reserved */
int pmb_sqp_get_phase_cycles(const pmb_sqp_t* s, unsigned long long* cycles16);
/* use an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the engine's own stream */
int pmb_sqp_set_stream(pmb_sqp_t* s, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* POLYMPC_B200_H */
