/*
 * vboc_b200.h -- C-ABI of libvboc_b200.so: the batched optimal-control engine that replaces the
 * acados solver objects behind the reference's per-system OCP classes.
 *
 * Every entry point takes plain pointers and sizes (no torch / C++ types) and returns 0 on success
 * or a negative error code; vboc_last_error() gives the message.  Host buffers stay owned by the
 * caller and are not retained past the call.  Per-problem results use the acados status integers
 * (0 success, 1 NaN, 2 max iter, 3 min step, 4 QP failure) that the reference's callers test
 * (VBOC/triplependulum_vboc.py:112, AL/triplependulum_class_al.py:164-169).
 *
 * What each function stands in for (file:line relative to the reference tree):
 *   vboc_create / vboc_destroy   AcadosOcpSolver(self.ocp, ...) construction
 *                                VBOC/triplependulum_class_vboc.py:153, VBOC/doublependulum_class_vboc.py:181,
 *                                AL/triplependulum_class_al.py:222, AL/doublependulum_class_al.py:486,
 *                                AL/pendulum_class_al.py:300
 *   vboc_set_opts                self.ocp.solver_options.* VBOC/triplependulum_class_vboc.py:129-141
 *   vboc_solve_batch             OCP_solve(...) = reset + per-stage set/constraints_set + solve() + get()
 *                                VBOC/triplependulum_class_vboc.py:155-191, triplependulum_testdata.py:47-75;
 *                                compute_problem(...) AL/triplependulum_class_al.py:148-169 (mode RTI);
 *                                one call replaces Pool.map over the problems
 *                                (VBOC/triplependulum_vboc.py:399-402)
 *   vboc_upload / vboc_solve_resident / vboc_download
 *                                the same three steps split so that inputs may stay resident in HBM
 *   vboc_datagen_*               data_generation(v) as a per-problem state machine on the device
 *                                VBOC/triplependulum_vboc.py:19-370, VBOC/doublependulum_vboc.py:19-402
 *   vboc_stream_*                the same solves as a ticket queue (see below): the drivers' per-problem loops
 *                                VBOC/triplependulum_vboc.py:107-136, 232-289, triplependulum_testdata.py:77-125
 *   vboc_mlp_create / _forward   model_dir(...) / sigmoid(model(...)) + entropy: my_nn.py:4-34,
 *                                VBOC/triplependulum_vboc.py:604-620, AL/triplependulum_al.py:253-264
 *   vboc_sim_step                SYMtriplependulumINIT.acados_integrator set/solve/get
 *                                VBOC/triplependulum_class_vboc.py:194-239, VBOC/triplependulum_vboc.py:348-352
 */
#ifndef VBOC_B200_H
#define VBOC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VBOC_FAMILY_VBOC 0 /* VBOC/ *_class_vboc.py: linear cost, dt state, stage-0 direction constraint */
#define VBOC_FAMILY_AL 1   /* AL/ *_class_al.py: LINEAR_LS cost on velocities, x0 fixed             */
#define VBOC_FAMILY_MPC 2  /* VBOC/Safe MPC/ *_class_fixedveldir.py, triplependulum_class_vboc.py: LINEAR_LS tracking cost,
                              x0 fixed, the learned viability margin h(x) >= 0 as a hard terminal constraint (vboc_set_mpc)
                              or softened at every stage (vboc_set_mpc_rows) */

#define VBOC_MODE_SQP 0 /* nlp_solver_type "SQP"     */
#define VBOC_MODE_RTI 1 /* nlp_solver_type "SQP_RTI" */

#define VBOC_SUCCESS 0
#define VBOC_FAILURE 1
#define VBOC_MAXITER 2
#define VBOC_MINSTEP 3
#define VBOC_QP_FAILURE 4

#define VBOC_ERR_ARG (-1)         /* bad argument / handle                                      */
#define VBOC_ERR_UNSUPPORTED (-2) /* problem data outside what the reference classes ever build */
#define VBOC_ERR_CUDA (-3)        /* CUDA runtime error (no device, launch failure, OOM)        */

typedef struct vboc_solver vboc_solver;

/* Solver options: the acados / HPIPM option names the reference sets or inherits. */
typedef struct {
    double tol_stat, tol_eq, tol_ineq, tol_comp; /* nlp_solver_tol_*          */
    int max_iter;                                /* nlp_solver_max_iter       */
    double levenberg_marquardt;
    double alpha_min, alpha_reduction; /* MERIT_BACKTRACKING               */
    int globalization;                 /* 1 merit back-tracking, 0 full step */
    double qp_tol_stat, qp_tol_eq, qp_tol_ineq, qp_tol_comp;
    int qp_iter_max;
    double qp_mu0, qp_alpha_min, qp_reg_prim, qp_lam_min, qp_t_min, qp_tau_min;
} vboc_opts;

/* Per-problem result record. */
typedef struct {
    int status;    /* acados status                          */
    int sqp_iter;  /* QPs solved                             */
    int qp_iter;   /* IPM iterations, total                  */
    int ls_evals;  /* merit evaluations at trial points      */
    int qp_status; /* last QP: 0 ok, 1 max iter, 2 min step, 3 NaN */
    int pad_;
    double cost; /* get_cost() at the returned iterate       */
    double res_stat, res_eq, res_ineq, res_comp;
} vboc_stats;

/* Reference defaults for a family (VBOC classes: VBOC/triplependulum_class_vboc.py:129-141; AL
 * classes: acados defaults). */
void vboc_default_opts(int family, vboc_opts *o);

/* n_dof 1..3, family VBOC_FAMILY_*, batch_capacity = largest batch of one call, N_max = largest
 * horizon, device = CUDA ordinal.  Allocates all device memory the solver will ever use. */
int vboc_create(int n_dof, int family, int batch_capacity, int N_max, int device, vboc_solver **out);
void vboc_destroy(vboc_solver *s);
int vboc_set_opts(vboc_solver *s, const vboc_opts *o);
/* Launch on this CUDA stream (cudaStream_t passed as void*); default: the legacy default stream. */
int vboc_set_stream(vboc_solver *s, void *cuda_stream);

/*
 * Batched OCP_solve / compute_problem.  All arrays are HOST, row-major, reference-shaped, with a
 * leading batch dimension:
 *   N        [batch]                horizon of each problem (1..N_max)
 *   x_guess  [batch][N_max+1][nx]   nx = 2n+1 (VBOC: q, v, dt) or 2n (AL); rows 0..N[b] used
 *   u_guess  [batch][N_max][nu]
 *   p        [batch][n+1]           VBOC cost weights (w, wt); NULL for AL
 *   lbx0/ubx0, lbx/ubx (stages 1..N-1), lbxN/ubxN  [batch][nx];   lbu/ubu [batch][nu]
 *   C0       [batch][n][nx] stage-0 general constraint 0 <= C0 x <= 0 or NULL; only the reference's
 *            projector C0 = [0 | I - d d' | 0] is accepted (VBOC/triplependulum_class_vboc.py:174-178)
 *   Tf       AL horizon in seconds (step Tf/N); ignored for VBOC (dt is the pinned state)
 * Outputs: x [batch][N_max+1][nx], u [batch][N_max][nu], stats [batch].
 * Restrictions checked here (VBOC_ERR_UNSUPPORTED): the dt state pinned to one value at every stage
 * and in the guess; terminal equalities on none or exactly all velocities.
 */
int vboc_solve_batch(vboc_solver *s, int mode, int batch, const int *N, const double *x_guess,
                     const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                     const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                     const double *lbu, const double *ubu, const double *C0, double Tf, double *x,
                     double *u, vboc_stats *stats);

/* The same, split: upload validates + copies the problem data to the device; solve_resident runs
 * the solver on what is resident (may be called repeatedly: every call starts from the uploaded
 * guess); download copies solutions and stats back. */
int vboc_upload(vboc_solver *s, int batch, const int *N, const double *x_guess,
                const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                const double *lbu, const double *ubu, const double *C0, double Tf);
int vboc_solve_resident(vboc_solver *s, int mode);
/* Launch without waiting / wait for everything queued on the solver's stream.  Two solvers on two
 * streams let the tail of one batch (a few long-running problems) overlap the next batch. */
int vboc_solve_resident_async(vboc_solver *s, int mode);
int vboc_sync(vboc_solver *s);
int vboc_download(vboc_solver *s, double *x, double *u, vboc_stats *stats);
/*
 * KKT multipliers at the returned iterate, so that acados' exit test (ocp_nlp_sqp: res_stat < tol_stat,
 * res_eq / res_ineq / res_comp < tol; options at VBOC/triplependulum_class_vboc.py:129-141) can be recomputed
 * from the results without trusting the solver (tools/certify.py).  Call vboc_export_multipliers(s, 1)
 * before the solve (allocates the device arrays), then after it
 *   pi  [batch][N_max][2n]        multipliers of the shooting equalities  Phi_h(x_k, u_k) - x_{k+1} = 0
 *   lam [batch][N_max+1][3n][2]   multipliers (lower, upper) of the bounds on z_k = [u_k; q_k; v_k]
 * in the engine's reduced coordinates (the pinned dt state is eliminated, so its bound multipliers do not
 * appear; the multipliers of the eliminated equalities -- fixed initial components, (I - d d') v_0 = 0,
 * v_N = const -- are free in sign and follow from stationarity).  The Lagrangian is
 *   cost + sum_k pi_k'(Phi(x_k, u_k) - x_{k+1}) + sum lam_u (z - ub) + lam_l (lb - z).
 * Warp kernel only (VBOC_ERR_UNSUPPORTED for the 1-DOF free-dt problem).
 */
int vboc_export_multipliers(vboc_solver *s, int on);
int vboc_download_multipliers(vboc_solver *s, double *pi, double *lam);
/*
 * MPC family (vboc_create(n, VBOC_FAMILY_MPC, ...), n = 2 or 3): the OCP of the reference's Safe-MPC classes with the
 * learned viability margin INSIDE the optimisation (SURVEY 8(f)4;
 * VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:110-258,
 * VBOC/Safe MPC/parallel/doublependulum_class_fixedveldir.py:240-264):
 *   min  sum_k Ts/2 |y_k - y_ref|^2_W + 1/2 |x_N - y_ref_e|^2_W_e     y = [x; u], LINEAR_LS, Gauss-Newton + LM
 *   s.t. x_{k+1} = Phi_Ts(x_k, u_k), x_0 = given (lbx0 == ubx0), box bounds, lh <= h(x_N) <= uh,
 *        h(x) = out(x) (100 - safety_margin)/100 - max(|v|, 1e-3),  out = the 2n-H-H-1 MLP on [(q - mean)/std, v/|v|]
 *        without its output ReLU (nn_decisionfunction).
 * vboc_set_mpc: the network (PyTorch nn.Linear layout, float32, one output), its normalisation, the constraint's
 * bounds and the diagonals of cost.W (3n, order [x; u]) and cost.W_e (2n); vboc_set_mpc_reference: y_ref [batch][3n]
 * and y_ref_e [batch][2n] of the problems of the next vboc_upload / vboc_solve_batch (cost_set(i, 'y_ref', ...),
 * :208-213).  vboc_solve_batch then takes x_guess [batch][N_max+1][2n], p = C0 = NULL, Tf = horizon in seconds,
 * mode VBOC_MODE_RTI (the classes' default nlp_solver_type) or VBOC_MODE_SQP.  vboc_download_mpc_multipliers: the
 * (lower, upper) multipliers of the terminal constraint at the returned iterate, lamg [batch][2].
 */
int vboc_set_mpc(vboc_solver *s, int hidden, const float *W1, const float *b1, const float *W2, const float *b2,
                 const float *W3, const float *b3, double mean, double stdv, double safety_margin, double lh, double uh,
                 const double *W, const double *W_e);
int vboc_set_mpc_reference(vboc_solver *s, int batch, const double *yref, const double *yref_e);
int vboc_download_mpc_multipliers(vboc_solver *s, double *lamg);
/*
 * Soft rows: the `parallel`, `receiding_hard_constraints` and `soft_traj_constraints` variants of the Safe-MPC classes
 * (VBOC/Safe MPC/parallel/doublependulum_class_fixedveldir.py:175-199) impose the margin at EVERY stage and soften it,
 *     con_h_expr = con_h_expr_e = h,  lh = 0, uh = 1e6,  idxsh = idxsh_e = [0]:
 *     lh <= h(x_k) + sl_k,  h(x_k) - su_k <= uh,  sl_k, su_k >= 0,  cost += 1/2 Zl_k sl_k^2 + zl_k sl_k + 1/2 Zu_k su_k^2 + zu_k su_k,
 * with the penalties set per stage at run time (cost_set(i, "Zl", ...), VBOC/Safe MPC/parallel/2dof_sym.py:44-50,
 * receiding_hard_constraints/2dof_sym.py:53-57, soft_traj_constraints/2dof_sym.py:110-111).
 * vboc_set_mpc_rows: Z [batch][N_max+1][4] = (Zl, Zu, zl, zu) per problem and stage (stage k of a problem with horizon
 * N <= N_max uses rows 0..N) for the problems of the next vboc_upload / vboc_solve_batch; batch = 0 switches back to the
 * single hard terminal row.  The slacks are variables of the NLP (zero at reset, stepped with the line search) and are
 * eliminated from the Newton systems of the IPM, so a row costs a rank-one term in its stage's Riccati step.
 * vboc_download_mpc_rows: rows [batch][N_max+1][6] = (lam_l, lam_u, lam_sl, lam_su, sl, su) at the returned iterate.
 */
int vboc_set_mpc_rows(vboc_solver *s, int batch, const double *Z);
/* vel_norm = |x[vstart:]| in the margin function (default vstart = n_dof: the velocities).  The triple-pendulum classes
 * write `norm_2(x[2:])` like the double-pendulum ones (VBOC/Safe MPC/triplependulum_class_vboc.py:217, 282), which for
 * n_dof = 3 includes theta_3; vstart = 2 reproduces their constraint function exactly.  Call after vboc_set_mpc. */
int vboc_set_mpc_velnorm_start(vboc_solver *s, int vstart);
int vboc_download_mpc_rows(vboc_solver *s, double *rows);
/*
 * Cartesian path constraint (VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:147-160, driven by
 * VBOC/Cartesian constraints/vboc_multiprocessing.py): the VBOC OCP of the double pendulum with the nonlinear constraint
 *     lh <= h(q_k) <= uh,  h(q) = (l1 sin q1 + l2 sin q2 - xc)^2 + (l1 cos q1 + l2 cos q2 - yc)^2   (con_h_expr, k = 0..N-1)
 * -- the end effector stays outside the circle of radius sqrt(lh) around (xc, yc); the reference uses xc = 0,
 * yc = -l1 - l2/2, lh = (l2/4)^2, uh = 1e6.  vboc_set_cartesian(s, 1, ...) on a VBOC-family handle with n = 2 switches the
 * handle's solves to the kernel that carries one hard general row per stage (linearised every SQP iteration, its barrier
 * term a rank-one update of the stage Hessian in the Riccati step, its violation in the merit function);
 * vboc_set_cartesian(s, 0, ...) switches back.  vboc_download_mpc_rows then returns the row multipliers in columns 0:2.
 */
int vboc_set_cartesian(vboc_solver *s, int on, double xc, double yc, double lh, double uh);
/*
 * AL family: `compute_problem_nnguess` (AL/triplependulum_class_al.py:171-201) with the guess network evaluated INSIDE the
 * solve kernel: a 2n-H-H-(N 2n) MLP (my_nn.py NeuralNetCLS, PyTorch nn.Linear layout, float32) predicts the state
 * trajectory of stages 1..N from the initial state, out = model((x0 - mean) / std) * std + mean; stage 0 takes x0.  After
 * vboc_set_guess_network the x_guess argument of vboc_solve_batch / vboc_upload is ignored (u_guess is still read);
 * hidden = 0 switches back to host-supplied guesses.  vboc_download_guess: the guesses the kernel computed,
 * [batch][N_max+1][2n] (FP64 arithmetic on the FP32 weights; PyTorch's FP32 forward differs by ~1e-6 relative).
 */
int vboc_set_guess_network(vboc_solver *s, int hidden, int n_out, const float *W1, const float *b1, const float *W2,
                           const float *b2, const float *W3, const float *b3, double mean, double stdv);
int vboc_download_guess(vboc_solver *s, double *x_guess);
/* Device time of the last vboc_solve_resident kernel in milliseconds (CUDA events on the solver's
 * stream); negative if none. */
double vboc_last_kernel_ms(vboc_solver *s);

/*
 * Streaming engine.  Replaces `Pool.map(data_generation, ...)` / `Pool.map(testing, ...)`
 * (VBOC/triplependulum_vboc.py:399-402, triplependulum_testdata.py:141-142) when every problem is a CHAIN of
 * dependent OCP_solve calls of very different lengths (extension loop VBOC/triplependulum_vboc.py:107-136,
 * sub-OCP chains :232-289): problems are submitted whenever a worker has one ready and are collected one by
 * one as they finish, so no solve ever waits for another worker's solve.
 *   vboc_stream_submit    same arrays as vboc_solve_batch with `count` rows; writes one ticket per problem.
 *                         count must not exceed vboc_stream_free_slots().  Returns right after the launch.
 *   vboc_stream_poll      non-blocking: up to `max` tickets whose solve has finished; returns how many.
 *   vboc_stream_fetch     copies the result of a finished ticket (rows 0..N of x, 0..N-1 of u) and frees it.
 *   vboc_stream_sim_step  vboc_sim_step on the engine's own stream (does not wait for solves in flight).
 * A free dt state (1-DOF VBOC driver) is not served here (VBOC_ERR_UNSUPPORTED): use vboc_solve_batch.
 * One handle per host thread.
 */
typedef struct vboc_stream vboc_stream;
int vboc_stream_create(int n_dof, int family, int capacity, int N_max, int device, vboc_stream **out);
void vboc_stream_destroy(vboc_stream *s);
int vboc_stream_set_opts(vboc_stream *s, const vboc_opts *o);
int vboc_stream_free_slots(vboc_stream *s);
int vboc_stream_pending(vboc_stream *s);
int vboc_stream_submit(vboc_stream *s, int mode, int count, const int *N, const double *x_guess,
                       const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                       const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                       const double *lbu, const double *ubu, const double *C0, double Tf, int *tickets);
int vboc_stream_poll(vboc_stream *s, int max, int *tickets);
int vboc_stream_fetch(vboc_stream *s, int ticket, double *x, double *u, vboc_stats *stats);
int vboc_stream_sim_step(vboc_stream *s, int count, const double *x, const double *u, double T, double *x_next);

/*
 * Device-resident data generation: `Pool.map(data_generation, range(count))` (VBOC/triplependulum_vboc.py:399-405,
 * VBOC/doublependulum_vboc.py:431-437) as ONE kernel -- one warp runs the whole per-problem state machine of
 * `data_generation(v)` (:19-370): the extreme trajectory with horizon extension and restarts, the walk along it with
 * its sub-OCP chains and the simulated unviable twin, the row filter.  No host round trip between the solves.
 * The random draws of the reference's worker are inputs (the host draws them from a seeded per-problem stream):
 *   joint_sel [count]            the joint that starts at a position limit
 *   p         [count][n+1]       cost direction (unit vector, then 0)
 *   lb0, ub0  [count][2n+1]      bounds of the initial state (fixed positions lb == ub, velocities +-v_max, dt)
 *   retry     [count][10][n+1]   per restart: perturbation added to the direction (n) and to the free positions (1)
 * Outputs: rows [rows_capacity][2n]: the saved states [q, v] of all problems back to back in problem order (problem b
 * owns stats[b].n_rows of them; at most VBOC_DG_ROWS_MAX each), *total_rows their number, and the per-problem
 * counters.  n_dof 2 or 3; N0 = initial horizon (100), dt = pinned time step (1e-2), tol (1e-3).
 */
#define VBOC_DG_ROWS_MAX 258
typedef struct {
    int status;    /* 0 rows returned, 1 no extreme trajectory found (the generator returns None), 2 row buffer overflow */
    int n_rows, solves, converged, sim_steps, sqp_iter, qp_iter;
    int t_done_us; /* completion time of the problem, microseconds after the kernel start (device clock) */
} vboc_dg_stats;
typedef struct vboc_datagen vboc_datagen;
int vboc_datagen_create(int n_dof, int capacity, int device, vboc_datagen **out);
void vboc_datagen_destroy(vboc_datagen *s);
int vboc_datagen_set_opts(vboc_datagen *s, const vboc_opts *o);
int vboc_datagen_run(vboc_datagen *s, int count, int N0, double dt, double tol, const int *joint_sel, const double *p,
                     const double *lb0, const double *ub0, const double *retry, double *rows, long long rows_capacity,
                     long long *total_rows, vboc_dg_stats *stats);
double vboc_datagen_last_kernel_ms(vboc_datagen *s);
/*
 * The test-data worker on the device: `Pool.map(testing, range(count))` (triplependulum_testdata.py:9-125, :141-142;
 * doublependulum_testdata.py:10-121) as one kernel on the same handle -- per problem the extension loop (horizon + 1
 * while the rounded cost still decreases) with restarts after failed solves.  Inputs are the worker's random draws:
 * ran [count][n] (un-normalised direction), q_init [count][n], retry [count][60][2n] (per restart the perturbations of
 * ran and of q_init).  rows [count][2n]: the boundary state of problem b if stats[b].status == 0.  Deliberate deviation:
 * at most max_solves (<= 60) solves per problem instead of the reference's `while True`.
 */
int vboc_testdata_run(vboc_datagen *s, int count, int N0, double dt, int max_solves, const double *ran,
                      const double *q_init, const double *retry, double *rows, vboc_dg_stats *stats);

/* One classical RK4 step of the unscaled 2n-state model: x_next = Phi_T(x, u).  HOST arrays
 * x [batch][2n], u [batch][n], x_next [batch][2n]. */
int vboc_sim_step(int n_dof, int device, int batch, const double *x, const double *u, double T,
                  double *x_next);

/*
 * Inference of the reference's MLPs (my_nn.py:4-34; hidden 100/300/500, n_in = 2n, n_out 1 or 2) with the
 * drivers' input normalisation and label / margin / entropy epilogues fused in.  Weights are HOST float
 * arrays in PyTorch nn.Linear layout ([out_features][in_features]); final_relu = 1 for NeuralNetDIR.
 * vboc_mlp_forward: x HOST [batch][n_in] raw states, out [batch][n_out];
 *   mode 0: x already normalised;
 *   mode 1 (VBOC/triplependulum_vboc.py:604-620): in = [(q-mean)/std, v/|v|]; label = |v| > phi ? 0 : 1;
 *          aux = phi*(100-safety_margin)/100 - |v|;
 *   mode 2 (AL/triplependulum_al.py:253-264): in = (x-mean)/std; aux = entropy of the renormalised
 *          sigmoid(logits).   aux / label may be NULL.
 */
typedef struct vboc_mlp vboc_mlp;
int vboc_mlp_create(int device, int n_in, int hidden, int n_out, int final_relu, const float *W1,
                    const float *b1, const float *W2, const float *b2, const float *W3, const float *b3,
                    vboc_mlp **out);
void vboc_mlp_destroy(vboc_mlp *m);
int vboc_mlp_forward(vboc_mlp *m, int batch, const float *x, int mode, double mean, double stdv,
                     double safety_margin, float *out, float *aux, int *label);
/* Device time of the last vboc_mlp_forward kernel in milliseconds (CUDA events around the launch). */
double vboc_mlp_last_kernel_ms(vboc_mlp *m);

/*
 * Resident unlabeled pool of the AL drivers (AL/triplependulum_al.py:100-123, 241-293): the pool (15^6 = 1.14e7 states
 * for the 3-DOF system) is uploaded ONCE and stays in HBM across the rounds of the loop; per round
 *   vboc_pool_score            entropy of sigmoid(model((x - mean) / std)) of every pool row (:253-264), scores stay
 *                              on the device;
 *   vboc_pool_select           the k most uncertain rows (`np.argpartition(etp, -k)[-k:]`, :267-270) by a radix select
 *                              on the device; idx [k] comes back sorted largest index first as the drivers sort it;
 *                              x [k][n_in] / score [k] (optional) are the selected rows / their entropies in that order;
 *   vboc_pool_remove_selected  `np.delete(X_prova, idx)` (:281): stable compaction of the pool on the device.
 * Only the k indices / rows per round cross the host link.  Ties at the k-th score are broken arbitrarily (as
 * argpartition does).  vboc_pool_download / _download_scores read the pool back (tests).
 */
typedef struct vboc_pool vboc_pool;
int vboc_pool_create(int device, int n_in, long long capacity, vboc_pool **out);
void vboc_pool_destroy(vboc_pool *p);
int vboc_pool_upload(vboc_pool *p, long long count, const float *x);
long long vboc_pool_size(vboc_pool *p);
int vboc_pool_score(vboc_pool *p, vboc_mlp *m, double mean, double stdv);
int vboc_pool_select(vboc_pool *p, int k, long long *idx, float *x, float *score);
int vboc_pool_remove_selected(vboc_pool *p);
int vboc_pool_download(vboc_pool *p, float *x);
int vboc_pool_download_scores(vboc_pool *p, float *score);
double vboc_pool_last_score_ms(vboc_pool *p);

/* Measured FP64 FMA peak of the device in TFLOP/s (dependent-free DFMA chains on every SM): the
 * roofline denominator bench.py reports against (MEASURED_PEAKS.json has no FP64 entry). */
int vboc_fp64_peak(int device, double *tflops);

const char *vboc_last_error(void);
const char *vboc_version(void);

#ifdef __cplusplus
}
#endif
#endif
