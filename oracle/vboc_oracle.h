/*
 * vboc_oracle.h -- CPU oracle for the VBOC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (vboc_b200/) never links or calls it.
 *
 * PARITY UNPINNED for the solver part: the arithmetic of the reference's OCP solves lives in
 * acados / HPIPM / BLASFEO / CasADi, none of which is vendored in /root/reference, pinned to a
 * version, or installable in the build container, and the reference holds no test or golden
 * vector.  What IS pinned: the dynamics (and their Jacobians / RK4 map) against the reference's own
 * expression text, through tests/golden/dynamics_golden.npz (tools/make_golden.py).
 * The solver is a restatement of the published algorithms (acados ocp_nlp_sqp / ocp_nlp_sqp_rti
 * with MERIT_BACKTRACKING, HPIPM's Mehrotra predictor-corrector Riccati IPM in BALANCE mode)
 * driven by the options the reference sets; every guessed detail is marked [restated] in the .c.
 */
#ifndef VBOC_ORACLE_H
#define VBOC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FAMILY_VBOC 0 /* time-scaled dynamics, dt state, linear cost (VBOC/ *_class_vboc.py) */
#define ORC_FAMILY_AL 1   /* plain dynamics, LINEAR_LS cost, x0 fixed (AL/ *_class_al.py)        */

#define ORC_MODE_SQP 0 /* acados nlp_solver_type "SQP"     */
#define ORC_MODE_RTI 1 /* acados nlp_solver_type "SQP_RTI" */

/* acados status integers (acados/utils/types.h [restated]) */
#define ORC_SUCCESS 0
#define ORC_FAILURE 1 /* NaN in the iterate */
#define ORC_MAXITER 2
#define ORC_MINSTEP 3
#define ORC_QP_FAILURE 4

typedef struct {
    /* NLP (acados) */
    double tol_stat, tol_eq, tol_ineq, tol_comp; /* nlp_solver_tol_*                     */
    int max_iter;                                /* nlp_solver_max_iter                  */
    double levenberg_marquardt;                  /* added to the Hessian diagonal        */
    double alpha_min, alpha_reduction;           /* MERIT_BACKTRACKING                   */
    int globalization;                           /* 1 = merit backtracking, 0 = full step */
    /* QP (HPIPM) */
    double qp_tol_stat, qp_tol_eq, qp_tol_ineq, qp_tol_comp;
    int qp_iter_max;
    double qp_mu0, qp_alpha_min, qp_reg_prim, qp_lam_min, qp_t_min, qp_tau_min;
    /* formulation switches (exact reformulations) */
    int eliminate_dt; /* 1: drop the dt state when it is pinned (lb==ub==guess) at every stage */
} orc_opts;

/* Reference defaults: VBOC classes (VBOC/triplependulum_class_vboc.py:129-141) / AL classes
 * (AL/triplependulum_class_al.py, acados defaults). */
void orc_default_opts(int family, orc_opts *o);

/*
 * One OCP, reference-shaped (row-major) data:
 *   n        DOF 1..3;  family ORC_FAMILY_*
 *   N        shooting intervals (<= 256)
 *   nx       = 2n+1 (VBOC: q, v, dt) or 2n (AL);  nu = n
 *   x_guess  (N+1) x nx,  u_guess  N x nu
 *   p        VBOC: n+1 cost weights (w, wt); AL: ignored (may be NULL)
 *   lbx0/ubx0, lbx/ubx (stages 1..N-1), lbxN/ubxN : nx each;  lbu/ubu : nu
 *   C0       ng x nx row-major stage-0 general constraint 0 <= C0 x <= 0, ng in {0, n}; NULL if ng==0
 *   Tf       AL: horizon in seconds (step h = Tf/N).  VBOC: ignored (step 1 on the dt-scaled model)
 * Outputs (caller allocated): x (N+1) x nx, u N x nu, pi N x nx, stats.
 * Returns the acados-style status.
 */
typedef struct {
    int status;
    int sqp_iter;  /* SQP iterations performed (QPs solved) */
    int qp_iter;   /* total IPM iterations                   */
    int ls_evals;  /* merit evaluations at trial points      */
    int qp_status; /* last HPIPM-style status 0 ok,1 maxiter,2 minstep,3 nan */
    double cost;   /* acados get_cost(): sum of stage costs at the returned iterate */
    double res_stat, res_eq, res_ineq, res_comp;
} orc_stats;

int orc_solve(int n, int family, int mode, int N, const double *x_guess, const double *u_guess,
              const double *p, const double *lbx0, const double *ubx0, const double *lbx,
              const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
              const double *ubu, const double *C0, int ng, double Tf, const orc_opts *opts,
              double *x, double *u, double *pi, orc_stats *stats);

/* Batched version of orc_solve over `batch` independent OCPs, OpenMP over problems.
 * All arrays gain a leading batch dimension; Nmax is the row stride of the guess / solution arrays
 * ((Nmax+1) x nx and Nmax x nu per problem); N[b] is the horizon of problem b. */
int orc_solve_batch(int n, int family, int mode, int batch, int Nmax, const int *N,
                    const double *x_guess, const double *u_guess, const double *p,
                    const double *lbx0, const double *ubx0, const double *lbx, const double *ubx,
                    const double *lbxN, const double *ubxN, const double *lbu, const double *ubu,
                    const double *C0, int ng, double Tf, const orc_opts *opts, double *x, double *u,
                    orc_stats *stats, int nthreads);

/* Dynamics pieces, exposed for the golden-vector tests.
 * orc_f: xdot = f(x,u) of the reference model (VBOC: dt-scaled incl. the zero dt row).
 * orc_f_jac: jx (nx x nx) and ju (nx x nu), row-major.
 * orc_rk4: one classical RK4 step (VBOC: step 1 on the scaled model; AL: step h) and, if A/B are
 * non-NULL, its exact Jacobians A = d(phi)/dx (nx x nx), B = d(phi)/du (nx x nu). */
void orc_f(int n, int family, const double *x, const double *u, double *xdot);
void orc_f_jac(int n, int family, const double *x, const double *u, double *jx, double *ju);
void orc_rk4(int n, int family, const double *x, const double *u, double h, double *xn, double *A,
             double *B);

/* Solve one structured QP with the oracle's IPM (for the K4 tests): data in the internal layout is
 * awkward to build from Python, so this solves the QP of the FIRST SQP iteration of the given OCP and
 * returns the step (dx (N+1) x nx_int, du N x nu) plus the HPIPM-style status. */
int orc_first_qp(int n, int family, int N, const double *x_guess, const double *u_guess,
                 const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                 const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                 const double *ubu, const double *C0, int ng, double Tf, const orc_opts *opts,
                 double *A_out, double *B_out, double *b_out, double *dx, double *du, int *qp_iter);

#ifdef __cplusplus
}
#endif
#endif
