/*
 * vboc_oracle.c -- CPU oracle for the VBOC hot path.  TEST INFRASTRUCTURE ONLY (see vboc_oracle.h).
 *
 * PARITY UNPINNED for the solver: acados/HPIPM are not available; this restates their published
 * algorithms.  Every detail taken from memory of upstream acados/HPIPM rather than from a file in
 * /root/reference is tagged [restated].  Dynamics ARE pinned (tests/golden/dynamics_golden.npz).
 *
 * What each block follows (file:line relative to /root/reference):
 *   dynamics            VBOC/pendulum_class_vboc.py:35-39, VBOC/doublependulum_class_vboc.py:40-91,
 *                       VBOC/triplependulum_class_vboc.py:47-58 (and the AL twins
 *                       AL/pendulum_class_al.py:40-44, AL/doublependulum_class_al.py:48-98,
 *                       AL/triplependulum_class_al.py:55-65).  Written here in manipulator form
 *                       M(q) a + c(q,v) + G(q) = u, which reproduces the reference's expanded
 *                       closed forms to rounding (checked against the golden vectors).
 *   OCP data / options  VBOC/triplependulum_class_vboc.py:71-141 and :155-191 (OCP_solve);
 *                       AL/triplependulum_class_al.py:82-169 and :204-222.
 *   integrator          acados sim_erk, 4 stages, 1 step, forward sensitivities [restated]; the
 *                       simulator VBOC/triplependulum_class_vboc.py:233-239.
 *   NLP                 acados ocp_nlp_sqp.c / ocp_nlp_sqp_rti.c / ocp_nlp_common.c [restated].
 *   QP                  HPIPM d_ocp_qp_ipm (BALANCE mode, no condensing) [restated].
 */
#include "vboc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* CPU-baseline build (bench.py --impl reference): -DORC_FIXED_N=<n> fixes the dimensions of the dt-eliminated
 * problem at compile time so that the small dense loops unroll and vectorise; prepare() refuses anything else.
 * The default build reads them from the problem (all systems, both families, free dt). */
#ifdef ORC_FIXED_N
#define DIM_NX(P) (2 * ORC_FIXED_N)
#define DIM_NU(P) (ORC_FIXED_N)
#define DIM_NZ(P) (3 * ORC_FIXED_N)
#else
#define DIM_NX(P) ((P)->nx)
#define DIM_NU(P) ((P)->nu)
#define DIM_NZ(P) ((P)->nz)
#endif
#define NXI 7  /* max internal nx (3 q + 3 v + dt) */
#define NUI 3
#define NZI 10 /* nu + nx */
#define NGI 3
#define ND 10  /* max derivative directions for the dual numbers */

/* ------------------------------------------------------------------------------------------ */
/* Model constants (reference class files, constructor blocks)                                  */
/* ------------------------------------------------------------------------------------------ */
static const double P1_M = 0.5, P1_G = 9.81, P1_D = 0.3, P1_B = 0.01; /* pendulum_class_vboc.py:14-17 */
static const double PN_M[3] = {0.4, 0.4, 0.4};                         /* triplependulum_class_vboc.py:15-17 */
static const double PN_L[3] = {0.8, 0.8, 0.8};                         /* :19-21 */
static const double PN_G = 9.81;

/* ------------------------------------------------------------------------------------------ */
/* Dual numbers: value + nd directional derivatives (forward-mode AD).                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double v;
    double d[ND];
} dual;

static int g_nd_dummy;
#define FORD for (int i_ = 0; i_ < nd; ++i_)

static dual d_const(double c, int nd) {
    dual r;
    r.v = c;
    FORD r.d[i_] = 0.0;
    (void)g_nd_dummy;
    return r;
}
static dual d_add(dual a, dual b, int nd) {
    dual r;
    r.v = a.v + b.v;
    FORD r.d[i_] = a.d[i_] + b.d[i_];
    return r;
}
static dual d_sub(dual a, dual b, int nd) {
    dual r;
    r.v = a.v - b.v;
    FORD r.d[i_] = a.d[i_] - b.d[i_];
    return r;
}
static dual d_mul(dual a, dual b, int nd) {
    dual r;
    r.v = a.v * b.v;
    FORD r.d[i_] = a.d[i_] * b.v + a.v * b.d[i_];
    return r;
}
static dual d_scale(dual a, double s, int nd) {
    dual r;
    r.v = a.v * s;
    FORD r.d[i_] = a.d[i_] * s;
    return r;
}
static dual d_div(dual a, dual b, int nd) {
    dual r;
    double inv = 1.0 / b.v;
    r.v = a.v * inv;
    FORD r.d[i_] = (a.d[i_] - r.v * b.d[i_]) * inv;
    return r;
}
static dual d_sin(dual a, int nd) {
    dual r;
    double c = cos(a.v);
    r.v = sin(a.v);
    FORD r.d[i_] = c * a.d[i_];
    return r;
}
static dual d_cos(dual a, int nd) {
    dual r;
    double s = -sin(a.v);
    r.v = cos(a.v);
    FORD r.d[i_] = s * a.d[i_];
    return r;
}

/* Joint accelerations a(q, v, u) of the n-link model on dual numbers. */
static void accel(int n, const dual *q, const dual *v, const dual *u, dual *a, int nd) {
    if (n == 1) {
        /* (m g d sin(theta) + F - b dtheta) / (d d m) */
        dual num = d_add(d_scale(d_sin(q[0], nd), P1_M * P1_G * P1_D, nd),
                         d_sub(u[0], d_scale(v[0], P1_B, nd), nd), nd);
        a[0] = d_scale(num, 1.0 / (P1_D * P1_D * P1_M), nd);
        return;
    }
    /* Point masses m_i at the tips of massless links l_i, ABSOLUTE link angles, theta = pi hanging:
     *   M_ij = mu_ij l_i l_j cos(q_i - q_j),   mu_ij = sum_{k >= max(i,j)} m_k
     *   c_i  = sum_j mu_ij l_i l_j sin(q_i - q_j) v_j^2
     *   G_i  = mu_ii g l_i sin(q_i)
     *   M a  = u - c - G                                                                          */
    dual M[3][3], r[3];
    for (int i = 0; i < n; ++i) {
        dual ci = d_const(0.0, nd);
        for (int j = 0; j < n; ++j) {
            double mu = 0.0;
            for (int k = (i > j ? i : j); k < n; ++k) mu += PN_M[k];
            double cf = mu * PN_L[i] * PN_L[j];
            dual dq = d_sub(q[i], q[j], nd);
            M[i][j] = d_scale(d_cos(dq, nd), cf, nd);
            ci = d_add(ci, d_scale(d_mul(d_sin(dq, nd), d_mul(v[j], v[j], nd), nd), cf, nd), nd);
        }
        double mui = 0.0;
        for (int k = i; k < n; ++k) mui += PN_M[k];
        dual Gi = d_scale(d_sin(q[i], nd), mui * PN_G * PN_L[i], nd);
        r[i] = d_sub(d_sub(u[i], ci, nd), Gi, nd);
    }
    /* Gaussian elimination without pivoting (M is SPD). */
    for (int k = 0; k < n; ++k) {
        for (int i = k + 1; i < n; ++i) {
            dual f = d_div(M[i][k], M[k][k], nd);
            for (int j = k; j < n; ++j) M[i][j] = d_sub(M[i][j], d_mul(f, M[k][j], nd), nd);
            r[i] = d_sub(r[i], d_mul(f, r[k], nd), nd);
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        dual s = r[i];
        for (int j = i + 1; j < n; ++j) s = d_sub(s, d_mul(M[i][j], a[j], nd), nd);
        a[i] = d_div(s, M[i][i], nd);
    }
}

/* xdot = f(x,u).  dts: x carries dt as its last component and f is scaled by it (VBOC models);
 * otherwise plain [v; a]. */
static void f_dual(int n, int dts, const dual *x, const dual *u, dual *xd, int nd) {
    dual a[3];
    accel(n, x, x + n, u, a, nd);
    if (dts) {
        dual dt = x[2 * n];
        for (int i = 0; i < n; ++i) xd[i] = d_mul(dt, x[n + i], nd);
        for (int i = 0; i < n; ++i) xd[n + i] = d_mul(dt, a[i], nd);
        xd[2 * n] = d_const(0.0, nd);
    } else {
        for (int i = 0; i < n; ++i) xd[i] = x[n + i];
        for (int i = 0; i < n; ++i) xd[n + i] = a[i];
    }
}

/* Classical RK4, one step of length h (acados sim_erk: 4 stages, 1 step [restated]).  With nd > 0
 * the duals carry d/d(x,u): the exact derivative of the discrete map, which is what ERK forward
 * sensitivities integrate. */
static void rk4_dual(int n, int dts, const dual *x, const dual *u, double h, dual *xn, int nd) {
    int nx = 2 * n + dts;
    dual k1[NXI], k2[NXI], k3[NXI], k4[NXI], xt[NXI];
    f_dual(n, dts, x, u, k1, nd);
    for (int i = 0; i < nx; ++i) xt[i] = d_add(x[i], d_scale(k1[i], 0.5 * h, nd), nd);
    f_dual(n, dts, xt, u, k2, nd);
    for (int i = 0; i < nx; ++i) xt[i] = d_add(x[i], d_scale(k2[i], 0.5 * h, nd), nd);
    f_dual(n, dts, xt, u, k3, nd);
    for (int i = 0; i < nx; ++i) xt[i] = d_add(x[i], d_scale(k3[i], h, nd), nd);
    f_dual(n, dts, xt, u, k4, nd);
    for (int i = 0; i < nx; ++i) {
        dual s = d_add(d_add(k1[i], d_scale(k2[i], 2.0, nd), nd),
                       d_add(d_scale(k3[i], 2.0, nd), k4[i], nd), nd);
        xn[i] = d_add(x[i], d_scale(s, h / 6.0, nd), nd);
    }
}

/* phi = RK4(x,u); A = dphi/dx (nx x nx), B = dphi/du (nx x nu); row-major; A/B may be NULL. */
static void integrate(int n, int dts, double h, const double *x, const double *u, double *phi,
                      double *A, double *B) {
    int nx = 2 * n + dts, nu = n;
    int nd = A ? nx + nu : 0;
    dual xd[NXI], ud[NUI], xn[NXI];
    for (int i = 0; i < nx; ++i) {
        xd[i] = d_const(x[i], nd);
        if (nd) xd[i].d[i] = 1.0;
    }
    for (int i = 0; i < nu; ++i) {
        ud[i] = d_const(u[i], nd);
        if (nd) ud[i].d[nx + i] = 1.0;
    }
    rk4_dual(n, dts, xd, ud, h, xn, nd);
    for (int i = 0; i < nx; ++i) {
        phi[i] = xn[i].v;
        if (nd) {
            for (int j = 0; j < nx; ++j) A[i * nx + j] = xn[i].d[j];
            for (int j = 0; j < nu; ++j) B[i * nu + j] = xn[i].d[nx + j];
        }
    }
}

void orc_f(int n, int family, const double *x, const double *u, double *xdot) {
    int dts = family == ORC_FAMILY_VBOC, nx = 2 * n + dts;
    dual xd[NXI], ud[NUI], o[NXI];
    for (int i = 0; i < nx; ++i) xd[i] = d_const(x[i], 0);
    for (int i = 0; i < n; ++i) ud[i] = d_const(u[i], 0);
    f_dual(n, dts, xd, ud, o, 0);
    for (int i = 0; i < nx; ++i) xdot[i] = o[i].v;
}

void orc_f_jac(int n, int family, const double *x, const double *u, double *jx, double *ju) {
    int dts = family == ORC_FAMILY_VBOC, nx = 2 * n + dts, nu = n, nd = nx + nu;
    dual xd[NXI], ud[NUI], o[NXI];
    for (int i = 0; i < nx; ++i) {
        xd[i] = d_const(x[i], nd);
        xd[i].d[i] = 1.0;
    }
    for (int i = 0; i < nu; ++i) {
        ud[i] = d_const(u[i], nd);
        ud[i].d[nx + i] = 1.0;
    }
    f_dual(n, dts, xd, ud, o, nd);
    for (int i = 0; i < nx; ++i) {
        for (int j = 0; j < nx; ++j) jx[i * nx + j] = o[i].d[j];
        for (int j = 0; j < nu; ++j) ju[i * nu + j] = o[i].d[nx + j];
    }
}

void orc_rk4(int n, int family, const double *x, const double *u, double h, double *xn, double *A,
             double *B) {
    int dts = family == ORC_FAMILY_VBOC;
    integrate(n, dts, dts ? 1.0 : h, x, u, xn, A, B);
}
/* ------------------------------------------------------------------------------------------ */
/* Internal OCP                                                                                 */
/*                                                                                              */
/* Exact reformulations applied before the QP (DESIGN.md "equalities"):                         */
/*   - a dt state pinned at every stage is dropped (opts.eliminate_dt);                         */
/*   - stage-0 equalities (components with lb == ub, and the projector constraint               */
/*     (I - d d') v_0 = 0 of OCP_solve, VBOC/triplependulum_class_vboc.py:174-178) are written */
/*     x_0 = c0 + Z0 y with orthonormal Z0 and solved in the reduced variable y;                */
/*   - terminal equalities (lbx_e == ubx_e on every velocity, :183-184 with q_fin bounds) are  */
/*     eliminated through the last control: G du_{N-1} = -(e_N + beta_v + Gx dx_{N-1}).         */
/* HPIPM instead sees all of these as two-sided inequalities with lb == ub [restated]; the QP   */
/* is strictly convex so both give the same primal solution, and the equality multipliers are   */
/* recovered from stationarity.                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n, family, N;
    int dts;        /* dt is a state (VBOC with free dt, or eliminate_dt == 0) */
    int nx, nu, nz; /* internal dims; z = [u; x] (HPIPM ordering)              */
    double h;        /* RK4 step: 1.0 when dts (scaled model), pinned dt or Tf/N otherwise */
    double w[3], wt; /* VBOC linear cost: w . v_0 + wt * sum_{k<N} dt_k                     */
    double dt_elim;  /* value of the eliminated dt (for cost reporting), 0 for AL          */
    double lb[3][NZI], ub[3][NZI]; /* z ordering, stage classes 0, 1..N-1, N */
    int fixed0[NXI];
    int hasdir;
    double d[3];
    int ny0;
    double Z0[NXI][NXI]; /* nx x ny0 */
    double c0[NXI];
    int fixedN[NXI], nfN, nqN, ivN[NXI], iqN[NXI];
    double cN[NXI];
    /* terminal equalities the last control cannot absorb (e.g. theta_N and dtheta_N fixed with one control,
     * VBOC/pendulum_class_vboc.py:123-124) are handled by BORDERING: the Riccati system is solved for the rhs
     * and for one unit terminal gradient per fixed component; the terminal multipliers follow from a small
     * dense system.  elimN = 1: eliminated through the last control; nb, bidx: bordered components. */
    int elimN, nb, bidx[NXI];
} iocp;

static inline int stage_class(const iocp *P, int k) { return k == 0 ? 0 : (k == P->N ? 2 : 1); }

/* is z-component i of stage k an inequality handled by the IPM? */
static inline int bnd_active(const iocp *P, int k, int i) {
    if (k == P->N && i < P->nu) return 0;
    if (k == 0 && i >= P->nu && P->fixed0[i - P->nu]) return 0;
    if (k == P->N && i >= P->nu && P->fixedN[i - P->nu]) return 0;
    return 1;
}

/* e = (I - Z0 Z0')(x - c0): violation of the stage-0 equalities at the state x */
static void eq0_violation(const iocp *P, const double *x, double *e) {
    int nx = DIM_NX(P);
    double y[NXI];
    for (int c = 0; c < P->ny0; ++c) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) s += P->Z0[i][c] * (x[i] - P->c0[i]);
        y[c] = s;
    }
    for (int i = 0; i < nx; ++i) {
        double s = x[i] - P->c0[i];
        for (int c = 0; c < P->ny0; ++c) s -= P->Z0[i][c] * y[c];
        e[i] = s;
    }
}
/* v <- Z0 Z0' v */
static void proj0(const iocp *P, double *v) {
    int nx = DIM_NX(P);
    double y[NXI], o[NXI];
    for (int c = 0; c < P->ny0; ++c) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) s += P->Z0[i][c] * v[i];
        y[c] = s;
    }
    for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int c = 0; c < P->ny0; ++c) s += P->Z0[i][c] * y[c];
        o[i] = s;
    }
    for (int i = 0; i < nx; ++i) v[i] = o[i];
}

typedef struct {
    int N, nx, nu, nz;
    /* NLP iterate */
    double *X, *U, *PI; /* (N+1)nx, N nu, N nx                                         */
    double *LAM;        /* (N+1) * 2nz : [lower nz | upper nz] in z ordering            */
    /* linearisation */
    double *A, *B, *bd; /* N*nx*nx, N*nx*nu, N*nx  (bd = phi(x_k,u_k) - x_{k+1})        */
    double *g, *hd;     /* (N+1)*nz cost gradient / Hessian diagonal (incl. LM)        */
    /* QP solution */
    double *DZ, *PIQ, *LAMQ, *TQ; /* (N+1)nz, N nx, (N+1)2nz, (N+1)2nz                    */
    double NU0Q[NXI], NUNQ[NXI];  /* multipliers of the eliminated equalities           */
    /* Riccati work */
    double *L;       /* N * nz*nz lower Cholesky factors of M_k                          */
    double *pv, *yv; /* (N+1)nx, N nu : Riccati vectors                                  */
    double hhN[NXI], rN[NXI];
    double Mlast[NZI][NZI], mlast[NZI], Kfb[NUI][NXI], Ginv[NUI][NUI], k0[NUI];
    double Lz[NXI][NXI];
    /* IPM work */
    double *rg, *rb, *rd, *rm; /* residuals: (N+1)nz, N nx, (N+1)2nz, (N+1)2nz          */
    double *dv, *dpi, *dlam, *dt;
    double *rmb, *hheff, *rr, *lbd, *ubd;
    double *dvb, *dpib, *zr, *zb; /* border basis solves: NXI x (N+1)nz, NXI x N nx; zero rhs vectors */
    double e0[NXI], eN[NXI];
    /* merit */
    double *wdyn, *wb; /* N nx, (N+1) 2nz */
    double w0[NXI], wN[NXI];
    double *Xt, *Ut; /* trial point */
} work;

static work *work_alloc(int N, int nx, int nu) {
    work *W = (work *)calloc(1, sizeof(work));
    int nz = nx + nu, S = N + 1;
    W->N = N, W->nx = nx, W->nu = nu, W->nz = nz;
#define AL_(ptr, cnt) W->ptr = (double *)calloc((size_t)(cnt), sizeof(double))
    AL_(X, S * nx), AL_(U, S * nu), AL_(PI, S * nx), AL_(LAM, S * 2 * nz);
    AL_(A, S * nx * nx), AL_(B, S * nx * nu), AL_(bd, S * nx), AL_(g, S * nz), AL_(hd, S * nz);
    AL_(DZ, S * nz), AL_(PIQ, S * nx), AL_(LAMQ, S * 2 * nz), AL_(TQ, S * 2 * nz);
    AL_(L, S * nz * nz), AL_(pv, S * nx), AL_(yv, S * nu);
    AL_(rg, S * nz), AL_(rb, S * nx), AL_(rd, S * 2 * nz), AL_(rm, S * 2 * nz);
    AL_(dv, S * nz), AL_(dpi, S * nx), AL_(dlam, S * 2 * nz), AL_(dt, S * 2 * nz);
    AL_(rmb, S * 2 * nz), AL_(hheff, S * nz), AL_(rr, S * nz), AL_(lbd, S * nz), AL_(ubd, S * nz);
    AL_(wdyn, S * nx), AL_(wb, S * 2 * nz), AL_(Xt, S * nx), AL_(Ut, S * nu);
    AL_(dvb, NXI * S * nz), AL_(dpib, NXI * S * nx), AL_(zr, S * nz), AL_(zb, S * nx);
#undef AL_
    return W;
}
static void work_free(work *W) {
    double **ps[] = {&W->X,   &W->U,   &W->PI,   &W->LAM, &W->A,   &W->B,     &W->bd,  &W->g,
                     &W->hd,  &W->DZ,  &W->PIQ,  &W->LAMQ, &W->TQ, &W->L,     &W->pv,  &W->yv,
                     &W->rg,  &W->rb,  &W->rd,   &W->rm,  &W->dv,  &W->dpi,   &W->dlam, &W->dt,
                     &W->rmb, &W->hheff, &W->rr, &W->lbd, &W->ubd, &W->wdyn,  &W->wb,  &W->Xt,
                     &W->Ut,  &W->dvb, &W->dpib, &W->zr, &W->zb};
    for (size_t i = 0; i < sizeof(ps) / sizeof(ps[0]); ++i) free(*ps[i]);
    free(W);
}

/* ------------------------------------------------------------------------------------------ */
/* Stage cost (acados cost modules EXTERNAL / LINEAR_LS scaled by the time step [restated])     */
/* ------------------------------------------------------------------------------------------ */
static double total_cost(const iocp *P, const double *X, const double *U) {
    int n = P->n, nx = DIM_NX(P), N = P->N;
    (void)U;
    double c = 0.0;
    if (P->family == ORC_FAMILY_VBOC) {
        /* stage 0: w.v + wt dt; stages 1..N-1: wt dt (triplependulum_class_vboc.py:85-86) */
        for (int i = 0; i < n; ++i) c += P->w[i] * X[n + i];
        if (P->dts)
            for (int k = 0; k < N; ++k) c += P->wt * X[k * nx + 2 * n];
        else
            c += P->wt * P->dt_elim * N;
    } else {
        /* 0.5 Ts y'Wy, W = 2 diag(0,1,0) ; terminal 0.5 y'W_e y (triplependulum_class_al.py:98-115) */
        for (int k = 0; k <= N; ++k) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += X[k * nx + n + i] * X[k * nx + n + i];
            c += (k < N ? P->h : 1.0) * s;
        }
    }
    return c;
}

static void cost_grad_hess(const iocp *P, const orc_opts *o, work *W) {
    int n = P->n, nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    for (int k = 0; k <= N; ++k) {
        double *g = W->g + k * nz, *hd = W->hd + k * nz;
        for (int i = 0; i < nz; ++i) g[i] = 0.0, hd[i] = o->levenberg_marquardt;
        if (P->family == ORC_FAMILY_VBOC) {
            if (k == 0)
                for (int i = 0; i < n; ++i) g[nu + n + i] = P->w[i];
            if (P->dts && k < N) g[nu + 2 * n] = P->wt;
        } else {
            double sc = (k < N ? P->h : 1.0);
            for (int i = 0; i < n; ++i) {
                g[nu + n + i] = 2.0 * sc * W->X[k * nx + n + i];
                hd[nu + n + i] += 2.0 * sc;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Linearisation: multiple shooting over all intervals                                          */
/* ------------------------------------------------------------------------------------------ */
static void linearize(const iocp *P, work *W) {
    int nx = DIM_NX(P), nu = DIM_NU(P), N = P->N;
    for (int k = 0; k < N; ++k) {
        double phi[NXI];
        integrate(P->n, P->dts, P->h, W->X + k * nx, W->U + k * nu, phi, W->A + k * nx * nx,
                  W->B + k * nx * nu);
        for (int i = 0; i < nx; ++i) W->bd[k * nx + i] = phi[i] - W->X[(k + 1) * nx + i];
    }
}

#define BAE(A, B, i, j) ((j) < nu ? (B)[(i) * nu + (j)] : (A)[(i) * nx + (j) - nu])

/* ------------------------------------------------------------------------------------------ */
/* NLP residuals (acados ocp_nlp_res_compute [restated]): inf-norms of                          */
/*   stat : gradient of the Lagrangian,  eq : shooting gaps,                                     */
/*   ineq : constraint violation,        comp : multiplier * constraint function                */
/* The eliminated equalities enter `ineq` through their violation; their multipliers absorb the  */
/* matching gradient components (stage 0: only the Z0-reduced gradient counts).                  */
/* ------------------------------------------------------------------------------------------ */
static void nlp_residuals(const iocp *P, work *W, double *rs, double *re, double *ri, double *rc) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    double s = 0, e = 0, in = 0, c = 0;
    for (int k = 0; k <= N; ++k) {
        int sc = stage_class(P, k);
        const double *lam = W->LAM + k * 2 * nz;
        double r[NZI];
        for (int i = 0; i < nz; ++i) r[i] = W->g[k * nz + i];
        if (k < N) {
            const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *pi = W->PI + k * nx;
            for (int j = 0; j < nz; ++j)
                for (int i = 0; i < nx; ++i) r[j] += BAE(A, B, i, j) * pi[i];
        }
        if (k > 0)
            for (int j = 0; j < nx; ++j) r[nu + j] -= W->PI[(k - 1) * nx + j];
        for (int i = (k == N ? nu : 0); i < nz; ++i) {
            if (!bnd_active(P, k, i)) continue;
            double z = i < nu ? W->U[k * nu + i] : W->X[k * nx + i - nu];
            double fl = P->lb[sc][i] - z, fu = z - P->ub[sc][i];
            r[i] += lam[nz + i] - lam[i];
            if (fl > in) in = fl;
            if (fu > in) in = fu;
            if (fabs(lam[i] * fl) > c) c = fabs(lam[i] * fl);
            if (fabs(lam[nz + i] * fu) > c) c = fabs(lam[nz + i] * fu);
        }
        if (k == 0) {
            double ev[NXI];
            proj0(P, r + nu);
            eq0_violation(P, W->X, ev);
            for (int i = 0; i < nx; ++i)
                if (fabs(ev[i]) > in) in = fabs(ev[i]);
        }
        if (k == N)
            for (int i = 0; i < nx; ++i)
                if (P->fixedN[i]) {
                    r[nu + i] = 0.0;
                    double v = fabs(W->X[N * nx + i] - P->cN[i]);
                    if (v > in) in = v;
                }
        for (int i = (k == N ? nu : 0); i < nz; ++i)
            if (fabs(r[i]) > s || r[i] != r[i]) s = fabs(r[i]);
        if (k < N)
            for (int i = 0; i < nx; ++i) {
                double v = fabs(W->bd[k * nx + i]);
                if (v > e || v != v) e = v;
            }
    }
    *rs = s, *re = e, *ri = in, *rc = c;
}

/* ------------------------------------------------------------------------------------------ */
/* Riccati factorisation / solves (HPIPM square-root backward recursion [restated]).            */
/*                                                                                              */
/* LQ sub-problem in the step dz_k = [du_k; dx_k]:                                               */
/*   min sum_k 1/2 dz'diag(hh_k)dz + r_k'dz                                                      */
/*   s.t. dx_{k+1} = A_k dx_k + B_k du_k + beta_k,                                               */
/*        dx_0 = -e0 + Z0 dy,     dx_N[fixed] = -eN                                              */
/* Value function 1/2 dx'P_k dx + p_k'dx with P_k = Lxx_k Lxx_k'.                               */
/* ------------------------------------------------------------------------------------------ */

/* in-place lower Cholesky of the leading n x n block of M (row-major, ld).  Non-positive pivots
 * zero the column, like BLASFEO's dpotrf_l [restated]. */
static void chol_lower(double *M, int n, int ld) {
    for (int j = 0; j < n; ++j) {
        double d = M[j * ld + j];
        for (int k = 0; k < j; ++k) d -= M[j * ld + k] * M[j * ld + k];
        double inv = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
        M[j * ld + j] = d > 0.0 ? sqrt(d) : 0.0;
        for (int i = j + 1; i < n; ++i) {
            double s = M[i * ld + j];
            for (int k = 0; k < j; ++k) s -= M[i * ld + k] * M[j * ld + k];
            M[i * ld + j] = s * inv;
        }
        for (int i = 0; i < j; ++i) M[i * ld + j] = 0.0;
    }
}

/* inverse of a small matrix by Gauss-Jordan with partial pivoting; returns 0 if singular */
static int small_inverse(int n, double G[NUI][NUI], double Gi[NUI][NUI]) {
    double a[NUI][2 * NUI];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) a[i][j] = G[i][j], a[i][n + j] = (i == j);
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int i = c + 1; i < n; ++i)
            if (fabs(a[i][c]) > fabs(a[p][c])) p = i;
        if (a[p][c] == 0.0 || a[p][c] != a[p][c]) return 0;
        if (p != c)
            for (int j = 0; j < 2 * n; ++j) {
                double t = a[c][j];
                a[c][j] = a[p][j], a[p][j] = t;
            }
        double inv = 1.0 / a[c][c];
        for (int j = 0; j < 2 * n; ++j) a[c][j] *= inv;
        for (int i = 0; i < n; ++i)
            if (i != c) {
                double f = a[i][c];
                for (int j = 0; j < 2 * n; ++j) a[i][j] -= f * a[c][j];
            }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Gi[i][j] = a[i][n + j];
    return 1;
}

/* hh: (N+1)*nz effective Hessian diagonal (hd + Gamma_l + Gamma_u). */
static int riccati_factor(const iocp *P, work *W, const double *hh) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    for (int i = 0; i < nx; ++i) W->hhN[i] = hh[N * nz + nu + i];
    for (int k = N - 1; k >= 0; --k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu;
        double *M = W->L + k * nz * nz;
        if (k == N - 1) {
            /* terminal value function is diagonal on the free components */
            double (*Mf)[NZI] = W->Mlast;
            for (int i = 0; i < nz; ++i)
                for (int j = 0; j < nz; ++j) Mf[i][j] = (i == j) ? hh[k * nz + i] : 0.0;
            for (int c = 0; c < nx; ++c) {
                if (P->elimN && P->fixedN[c]) continue;
                for (int i = 0; i < nz; ++i)
                    for (int j = 0; j < nz; ++j)
                        Mf[i][j] += W->hhN[c] * BAE(A, B, c, i) * BAE(A, B, c, j);
            }
            if (P->elimN) {
                double G[NUI][NUI];
                for (int a = 0; a < nu; ++a)
                    for (int b = 0; b < nu; ++b) G[a][b] = B[P->ivN[a] * nu + b];
                if (!small_inverse(nu, G, W->Ginv)) return 0;
                for (int a = 0; a < nu; ++a)
                    for (int j = 0; j < nx; ++j) {
                        double s = 0.0;
                        for (int b = 0; b < nu; ++b) s -= W->Ginv[a][b] * A[P->ivN[b] * nx + j];
                        W->Kfb[a][j] = s;
                    }
                /* P = T' Mf T, T = [K; I] */
                double MT[NZI][NXI];
                for (int i = 0; i < nz; ++i)
                    for (int j = 0; j < nx; ++j) {
                        double s = Mf[i][nu + j];
                        for (int a = 0; a < nu; ++a) s += Mf[i][a] * W->Kfb[a][j];
                        MT[i][j] = s;
                    }
                memset(M, 0, sizeof(double) * nz * nz);
                for (int i = 0; i < nx; ++i)
                    for (int j = 0; j <= i; ++j) {
                        double s = MT[nu + i][j];
                        for (int a = 0; a < nu; ++a) s += W->Kfb[a][i] * MT[a][j];
                        M[(nu + i) * nz + nu + j] = s;
                    }
                chol_lower(M + nu * nz + nu, nx, nz);
                continue;
            }
            for (int i = 0; i < nz; ++i)
                for (int j = 0; j <= i; ++j) M[i * nz + j] = Mf[i][j];
            chol_lower(M, nz, nz);
            continue;
        }
        double Wm[NXI][NZI]; /* Lxx+' [B A] */
        const double *Ln = W->L + (k + 1) * nz * nz;
        for (int j = 0; j < nz; ++j)
            for (int i = 0; i < nx; ++i) {
                double s = 0.0;
                for (int m = i; m < nx; ++m) s += Ln[(nu + m) * nz + nu + i] * BAE(A, B, m, j);
                Wm[i][j] = s;
            }
        for (int i = 0; i < nz; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = (i == j) ? hh[k * nz + i] : 0.0;
                for (int m = 0; m < nx; ++m) s += Wm[m][i] * Wm[m][j];
                M[i * nz + j] = s;
            }
        chol_lower(M, nz, nz);
    }
    /* stage-0 reduction: Pz = Z0' P0 Z0 = (Lxx0' Z0)'(Lxx0' Z0) */
    {
        const double *L0 = W->L;
        int ny = P->ny0;
        double T0[NXI][NXI];
        for (int i = 0; i < nx; ++i)
            for (int c = 0; c < ny; ++c) {
                double s = 0.0;
                for (int m = i; m < nx; ++m) s += L0[(nu + m) * nz + nu + i] * P->Z0[m][c];
                T0[i][c] = s;
            }
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b <= a; ++b) {
                double s = 0.0;
                for (int i = 0; i < nx; ++i) s += T0[i][a] * T0[i][b];
                W->Lz[a][b] = s;
            }
        chol_lower(&W->Lz[0][0], ny, NXI);
    }
    return 1;
}

/* y = P_k x with P_k = Lxx_k Lxx_k' (factor stored in stage k) */
static void P_times(const work *W, int k, const double *x, double *y) {
    int nx = DIM_NX(W), nu = DIM_NU(W), nz = DIM_NZ(W);
    const double *L = W->L + k * nz * nz;
    double t1[NXI];
    for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int m = i; m < nx; ++m) s += L[(nu + m) * nz + nu + i] * x[m];
        t1[i] = s;
    }
    for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int m = 0; m <= i; ++m) s += L[(nu + i) * nz + nu + m] * t1[m];
        y[i] = s;
    }
}

/* Solve with rhs r ((N+1)*nz gradient), beta (N*nx) and the equality residuals e0 (nx), eN (nx);
 * outputs dv ((N+1)*nz) and dpi (N*nx). */
static void riccati_solve(const iocp *P, work *W, const double *r, const double *beta,
                          const double *e0, const double *eN, double *dv, double *dpi) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    double *pv = W->pv, *yv = W->yv;
    for (int i = 0; i < nx; ++i) pv[N * nx + i] = W->rN[i] = r[N * nz + nu + i];
    for (int k = N - 1; k >= 0; --k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *L = W->L + k * nz * nz;
        double t2[NXI], m[NZI];
        if (k == N - 1) {
            for (int i = 0; i < nx; ++i)
                t2[i] = (P->elimN && P->fixedN[i]) ? 0.0 : W->hhN[i] * beta[k * nx + i] + pv[(k + 1) * nx + i];
        } else {
            P_times(W, k + 1, beta + k * nx, t2);
            for (int i = 0; i < nx; ++i) t2[i] += pv[(k + 1) * nx + i];
        }
        for (int j = 0; j < nz; ++j) {
            double s = r[k * nz + j];
            for (int i = 0; i < nx; ++i) s += BAE(A, B, i, j) * t2[i];
            m[j] = s;
        }
        if (k == N - 1 && P->elimN) {
            /* du = K dx + k0 with G k0 = -(eN + beta_v) */
            double tmp[NZI];
            for (int a = 0; a < nu; ++a) {
                double s = 0.0;
                for (int b = 0; b < nu; ++b)
                    s -= W->Ginv[a][b] * (eN[P->ivN[b]] + beta[k * nx + P->ivN[b]]);
                W->k0[a] = s;
            }
            for (int i = 0; i < nz; ++i) {
                double s = m[i];
                for (int a = 0; a < nu; ++a) s += W->Mlast[i][a] * W->k0[a];
                tmp[i] = s, W->mlast[i] = m[i];
            }
            for (int j = 0; j < nx; ++j) {
                double s = tmp[nu + j];
                for (int a = 0; a < nu; ++a) s += W->Kfb[a][j] * tmp[a];
                pv[k * nx + j] = s;
            }
            continue;
        }
        /* y = Luu^-1 m_u */
        for (int i = 0; i < nu; ++i) {
            double s = m[i];
            for (int j = 0; j < i; ++j) s -= L[i * nz + j] * yv[k * nu + j];
            yv[k * nu + i] = L[i * nz + i] > 0.0 ? s / L[i * nz + i] : 0.0;
        }
        for (int i = 0; i < nx; ++i) {
            double s = m[nu + i];
            for (int j = 0; j < nu; ++j) s -= L[(nu + i) * nz + j] * yv[k * nu + j];
            pv[k * nx + i] = s;
        }
    }
    /* stage 0: dx0 = -e0 + Z0 dy,  (Z0'P0 Z0) dy = -Z0'(p0 - P0 e0) */
    {
        int ny = P->ny0;
        double Pe[NXI], rhs[NXI], y[NXI], dy[NXI];
        P_times(W, 0, e0, Pe);
        for (int c = 0; c < ny; ++c) {
            double s = 0.0;
            for (int i = 0; i < nx; ++i) s -= P->Z0[i][c] * (pv[i] - Pe[i]);
            rhs[c] = s;
        }
        for (int i = 0; i < ny; ++i) {
            double s = rhs[i];
            for (int j = 0; j < i; ++j) s -= W->Lz[i][j] * y[j];
            y[i] = W->Lz[i][i] > 0.0 ? s / W->Lz[i][i] : 0.0;
        }
        for (int i = ny - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < ny; ++j) s -= W->Lz[j][i] * dy[j];
            dy[i] = W->Lz[i][i] > 0.0 ? s / W->Lz[i][i] : 0.0;
        }
        for (int i = 0; i < nx; ++i) {
            double s = -e0[i];
            for (int c = 0; c < ny; ++c) s += P->Z0[i][c] * dy[c];
            dv[nu + i] = s;
        }
    }
    for (int k = 0; k < N; ++k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *L = W->L + k * nz * nz;
        double *z = dv + k * nz, *zn = dv + (k + 1) * nz;
        if (k == N - 1 && P->elimN) {
            for (int a = 0; a < nu; ++a) {
                double s = W->k0[a];
                for (int j = 0; j < nx; ++j) s += W->Kfb[a][j] * z[nu + j];
                z[a] = s;
            }
        } else {
            /* u = -Luu^-T (y + Lxu' x) */
            double t[NUI];
            for (int i = 0; i < nu; ++i) {
                double s = yv[k * nu + i];
                for (int j = 0; j < nx; ++j) s += L[(nu + j) * nz + i] * z[nu + j];
                t[i] = s;
            }
            for (int i = nu - 1; i >= 0; --i) {
                double s = t[i];
                for (int j = i + 1; j < nu; ++j) s -= L[j * nz + i] * (-z[j]);
                z[i] = L[i * nz + i] > 0.0 ? -s / L[i * nz + i] : 0.0;
            }
        }
        for (int i = 0; i < nx; ++i) {
            double s = beta[k * nx + i];
            for (int j = 0; j < nz; ++j) s += BAE(A, B, i, j) * z[j];
            zn[nu + i] = s;
        }
        if (k == N - 1) {
            for (int i = 0; i < nu; ++i) zn[i] = 0.0;
            for (int i = 0; i < nx; ++i)
                dpi[k * nx + i] = (P->elimN && P->fixedN[i]) ? 0.0 : W->hhN[i] * zn[nu + i] + pv[(k + 1) * nx + i];
            if (P->elimN) {
                /* multiplier of the eliminated rows from the u-stationarity of the last stage */
                double tu[NUI];
                for (int a = 0; a < nu; ++a) {
                    double s = W->mlast[a];
                    for (int j = 0; j < nz; ++j) s += W->Mlast[a][j] * z[j];
                    tu[a] = s;
                }
                for (int b = 0; b < nu; ++b) {
                    double s = 0.0;
                    for (int a = 0; a < nu; ++a) s -= W->Ginv[a][b] * tu[a];
                    dpi[k * nx + P->ivN[b]] = s;
                }
            }
        } else {
            P_times(W, k + 1, zn + nu, dpi + k * nx);
            for (int i = 0; i < nx; ++i) dpi[k * nx + i] += pv[(k + 1) * nx + i];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* IPM (HPIPM d_ocp_qp_ipm_solve, BALANCE mode, pred_corr = cond_pred_corr = 1, cold start,     */
/* no iterative refinement / LQ fall-back [restated]).                                          */
/* Returns 0 success, 1 max iter, 2 min step, 3 NaN.                                            */
/* ------------------------------------------------------------------------------------------ */
static double qp_residuals(const iocp *P, work *W, double *ng_, double *nb_, double *nd_,
                           double *nm_) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    const double *lbd = W->lbd, *ubd = W->ubd;
    double rg = 0, rb = 0, rd = 0, rm = 0, mu = 0;
    int nc = 0;
    for (int k = 0; k <= N; ++k) {
        const double *v = W->DZ + k * nz, *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
        double *r = W->rg + k * nz;
        for (int i = 0; i < nz; ++i) r[i] = W->hd[k * nz + i] * v[i] + W->g[k * nz + i];
        if (k < N) {
            const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *pi = W->PIQ + k * nx;
            for (int j = 0; j < nz; ++j)
                for (int i = 0; i < nx; ++i) r[j] += BAE(A, B, i, j) * pi[i];
            const double *vn = W->DZ + (k + 1) * nz;
            for (int i = 0; i < nx; ++i) {
                double s = W->bd[k * nx + i] - vn[nu + i];
                for (int j = 0; j < nz; ++j) s += BAE(A, B, i, j) * v[j];
                W->rb[k * nx + i] = s;
                if (fabs(s) > rb || s != s) rb = fabs(s);
            }
        }
        if (k > 0)
            for (int j = 0; j < nx; ++j) r[nu + j] -= W->PIQ[(k - 1) * nx + j];
        for (int i = 0; i < nz; ++i) {
            if (!bnd_active(P, k, i)) {
                W->rd[k * 2 * nz + i] = W->rd[k * 2 * nz + nz + i] = 0.0;
                W->rm[k * 2 * nz + i] = W->rm[k * 2 * nz + nz + i] = 0.0;
                if (k == N) r[i] = 0.0; /* terminal controls do not exist; fixed x_N absorbed */
                continue;
            }
            r[i] += lam[nz + i] - lam[i];
            double dl = lbd[k * nz + i] - v[i] + t[i];
            double du = v[i] - ubd[k * nz + i] + t[nz + i];
            W->rd[k * 2 * nz + i] = dl, W->rd[k * 2 * nz + nz + i] = du;
            double ml = lam[i] * t[i], mu_ = lam[nz + i] * t[nz + i];
            W->rm[k * 2 * nz + i] = ml, W->rm[k * 2 * nz + nz + i] = mu_;
            if (fabs(dl) > rd || dl != dl) rd = fabs(dl);
            if (fabs(du) > rd || du != du) rd = fabs(du);
            if (fabs(ml) > rm || ml != ml) rm = fabs(ml);
            if (fabs(mu_) > rm || mu_ != mu_) rm = fabs(mu_);
            mu += ml + mu_;
            nc += 2;
        }
        if (k == 0) proj0(P, r + nu);
        for (int i = 0; i < nz; ++i)
            if (fabs(r[i]) > rg || r[i] != r[i]) rg = fabs(r[i]);
    }
    /* residuals of the eliminated equalities at the current QP iterate */
    {
        double x0[NXI];
        for (int i = 0; i < nx; ++i) x0[i] = W->X[i] + W->DZ[nu + i];
        eq0_violation(P, x0, W->e0);
        for (int i = 0; i < nx; ++i) {
            W->eN[i] = P->fixedN[i] ? W->X[N * nx + i] + W->DZ[N * nz + nu + i] - P->cN[i] : 0.0;
            if (fabs(W->e0[i]) > rb) rb = fabs(W->e0[i]);
            if (fabs(W->eN[i]) > rb) rb = fabs(W->eN[i]);
        }
    }
    *ng_ = rg, *nb_ = rb, *nd_ = rd, *nm_ = rm;
    return nc ? mu / nc : 0.0;
}

/* Build Gamma/gamma from (lam, t, res_d, res_m), factorise if asked, solve, then recover dlam, dt
 * and the maximum step alpha. */
static double ipm_step(const iocp *P, const orc_opts *o, work *W, int factor, const double *rm,
                       int *ok) {
    int nz = DIM_NZ(P), N = P->N;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            double h = W->hd[k * nz + i] + o->qp_reg_prim, r = W->rg[k * nz + i];
            if (bnd_active(P, k, i)) {
                const double *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
                double Gl = lam[i] / t[i], Gu = lam[nz + i] / t[nz + i];
                double gl = (rm[k * 2 * nz + i] - lam[i] * W->rd[k * 2 * nz + i]) / t[i];
                double gu =
                    (rm[k * 2 * nz + nz + i] - lam[nz + i] * W->rd[k * 2 * nz + nz + i]) / t[nz + i];
                h += Gl + Gu;
                r += gl - gu;
            }
            W->hheff[k * nz + i] = h;
            W->rr[k * nz + i] = r;
        }
    if (factor && !riccati_factor(P, W, W->hheff)) {
        *ok = 0;
        return 0.0;
    }
    riccati_solve(P, W, W->rr, W->rb, W->e0, W->eN, W->dv, W->dpi);
    if (P->nb) {
        int nx = DIM_NX(P), nu = DIM_NU(P), nb = P->nb, S = N + 1;
        double ze[NXI] = {0};
        if (factor) /* basis solves: unit terminal gradient on one fixed component, everything else zero */
            for (int j = 0; j < nb; ++j) {
                memset(W->zr, 0, sizeof(double) * S * nz);
                W->zr[N * nz + nu + P->bidx[j]] = 1.0;
                riccati_solve(P, W, W->zr, W->zb, ze, ze, W->dvb + (size_t)j * S * nz, W->dpib + (size_t)j * S * nx);
            }
        double Sm[NXI][NXI + 1], nuv[NXI];
        for (int i = 0; i < nb; ++i) {
            for (int j = 0; j < nb; ++j) Sm[i][j] = W->dvb[(size_t)j * S * nz + N * nz + nu + P->bidx[i]];
            Sm[i][nb] = -(W->eN[P->bidx[i]] + W->dv[N * nz + nu + P->bidx[i]]);
        }
        for (int c = 0; c < nb; ++c) {
            int pr = c;
            for (int i = c + 1; i < nb; ++i)
                if (fabs(Sm[i][c]) > fabs(Sm[pr][c])) pr = i;
            for (int j = 0; j <= nb; ++j) {
                double t = Sm[c][j];
                Sm[c][j] = Sm[pr][j], Sm[pr][j] = t;
            }
            for (int i = c + 1; i < nb; ++i) {
                double f = Sm[i][c] / Sm[c][c];
                for (int j = c; j <= nb; ++j) Sm[i][j] -= f * Sm[c][j];
            }
        }
        for (int i = nb - 1; i >= 0; --i) {
            double a = Sm[i][nb];
            for (int j = i + 1; j < nb; ++j) a -= Sm[i][j] * nuv[j];
            nuv[i] = a / Sm[i][i];
        }
        for (int j = 0; j < nb; ++j) {
            for (int c = 0; c < S * nz; ++c) W->dv[c] += nuv[j] * W->dvb[(size_t)j * S * nz + c];
            for (int c = 0; c < N * nx; ++c) W->dpi[c] += nuv[j] * W->dpib[(size_t)j * S * nx + c];
        }
    }
    double alpha = 1.0;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            if (!bnd_active(P, k, i)) continue;
            const double *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
            double dvv = W->dv[k * nz + i];
            for (int s = 0; s < 2; ++s) {
                int c = k * 2 * nz + s * nz + i;
                double dtt = (s ? -dvv : dvv) - W->rd[c];
                double dl = -(rm[c] + lam[s * nz + i] * dtt) / t[s * nz + i];
                W->dt[c] = dtt, W->dlam[c] = dl;
                if (dtt < 0.0 && -t[s * nz + i] / dtt < alpha) alpha = -t[s * nz + i] / dtt;
                if (dl < 0.0 && -lam[s * nz + i] / dl < alpha) alpha = -lam[s * nz + i] / dl;
            }
        }
    return alpha;
}

static double mu_aff(const iocp *P, work *W, double alpha) {
    int nz = DIM_NZ(P), N = P->N, nc = 0;
    double mu = 0;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            if (!bnd_active(P, k, i)) continue;
            for (int s = 0; s < 2; ++s) {
                int c = k * 2 * nz + s * nz + i;
                mu += (W->LAMQ[c] + alpha * W->dlam[c]) * (W->TQ[c] + alpha * W->dt[c]);
                nc++;
            }
        }
    return nc ? mu / nc : 0.0;
}

static int ipm_solve(const iocp *P, const orc_opts *o, work *W, int *iters) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    double *lbd = W->lbd, *ubd = W->ubd;
    const double thr0 = 0.1; /* HPIPM d_ocp_qp_init_var cold start threshold [restated] */
    /* bounds of the step: lb - z <= dz <= ub - z */
    for (int k = 0; k <= N; ++k) {
        int sc = stage_class(P, k);
        for (int i = 0; i < nz; ++i) {
            double z = i < nu ? (k < N ? W->U[k * nu + i] : 0.0) : W->X[k * nx + i - nu];
            lbd[k * nz + i] = P->lb[sc][i] - z;
            ubd[k * nz + i] = P->ub[sc][i] - z;
        }
    }
    /* cold start */
    memset(W->DZ, 0, sizeof(double) * (N + 1) * nz);
    memset(W->PIQ, 0, sizeof(double) * N * nx);
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            double *v = W->DZ + k * nz + i, *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
            if (!bnd_active(P, k, i)) {
                lam[i] = lam[nz + i] = t[i] = t[nz + i] = 0.0;
                if (k == N && i >= nu) *v = lbd[k * nz + i]; /* pinned terminal component */
                continue;
            }
            double tl = *v - lbd[k * nz + i], tu = ubd[k * nz + i] - *v;
            if (tl < thr0) {
                if (tu < thr0) {
                    *v = 0.5 * (lbd[k * nz + i] + ubd[k * nz + i]);
                    tl = tu = thr0;
                } else {
                    tl = thr0;
                    *v = lbd[k * nz + i] + thr0;
                }
            } else if (tu < thr0) {
                tu = thr0;
                *v = ubd[k * nz + i] - thr0;
            }
            t[i] = tl, t[nz + i] = tu;
            lam[i] = o->qp_mu0 / tl, lam[nz + i] = o->qp_mu0 / tu;
        }
    {
        /* put the stage-0 step on the equality manifold */
        double x0[NXI], e[NXI];
        for (int i = 0; i < nx; ++i) x0[i] = W->X[i] + W->DZ[nu + i];
        eq0_violation(P, x0, e);
        for (int i = 0; i < nx; ++i) W->DZ[nu + i] -= e[i];
    }
    double rg, rb, rd, rm, alpha = 1.0;
    double mu = qp_residuals(P, W, &rg, &rb, &rd, &rm);
    int kk = 0, ok = 1;
    for (; kk < o->qp_iter_max && alpha > o->qp_alpha_min &&
           (rg > o->qp_tol_stat || rb > o->qp_tol_eq || rd > o->qp_tol_ineq || rm > o->qp_tol_comp);
         ++kk) {
        /* affine (predictor) direction: res_m = lam * t */
        memcpy(W->rmb, W->rm, sizeof(double) * (N + 1) * 2 * nz);
        double a_aff = ipm_step(P, o, W, 1, W->rm, &ok);
        if (!ok) break;
        double m_aff = mu_aff(P, W, a_aff);
        double sigma = m_aff / mu;
        sigma = sigma * sigma * sigma;
        double sm = sigma * mu;
        if (sm < o->qp_tau_min) sm = o->qp_tau_min;
        /* centering + corrector: res_m = lam*t + dt_aff*dlam_aff - sigma*mu */
        for (int c = 0; c < (N + 1) * 2 * nz; ++c)
            W->rm[c] = W->TQ[c] != 0.0 ? W->rmb[c] + W->dt[c] * W->dlam[c] - sm : 0.0;
        alpha = ipm_step(P, o, W, 0, W->rm, &ok);
        /* conditional predictor-corrector: if the corrected step is much worse than the affine
         * one, fall back to the centering direction only (HPIPM cond_pred_corr [restated]). */
        double m_cor = mu_aff(P, W, alpha);
        if (m_cor > 2.0 * m_aff) {
            for (int c = 0; c < (N + 1) * 2 * nz; ++c)
                W->rm[c] = W->TQ[c] != 0.0 ? W->rmb[c] - sm : 0.0;
            alpha = ipm_step(P, o, W, 0, W->rm, &ok);
        }
        /* update (HPIPM d_update_var_qp: step shortened away from the boundary [restated]) */
        double as = alpha;
        if (as < 1.0) as = as * ((1.0 - as) * 0.99 + as * 0.9999);
        for (int c = 0; c < (N + 1) * nz; ++c) W->DZ[c] += as * W->dv[c];
        for (int c = 0; c < N * nx; ++c) W->PIQ[c] += as * W->dpi[c];
        for (int k = 0; k <= N; ++k)
            for (int i = 0; i < nz; ++i) {
                if (!bnd_active(P, k, i)) continue;
                for (int s = 0; s < 2; ++s) {
                    int c = k * 2 * nz + s * nz + i;
                    W->LAMQ[c] += as * W->dlam[c];
                    W->TQ[c] += as * W->dt[c];
                    if (W->LAMQ[c] < o->qp_lam_min) W->LAMQ[c] = o->qp_lam_min;
                    if (W->TQ[c] < o->qp_t_min) W->TQ[c] = o->qp_t_min;
                }
            }
        mu = qp_residuals(P, W, &rg, &rb, &rd, &rm);
        if (getenv("ORC_DEBUG"))
            fprintf(stderr,
                    "  ipm %3d a_aff %.3e alpha %.3e sigma %.2e mu %.3e rg %.2e rb %.2e rd %.2e rm "
                    "%.2e\n",
                    kk, a_aff, alpha, sigma, mu, rg, rb, rd, rm);
    }
    *iters = kk;
    /* multipliers of the eliminated equalities from stationarity (what HPIPM would return for the
     * lb == ub pairs, up to the split between the two sides [restated]) */
    {
        double r[NXI], rp[NXI];
        for (int i = 0; i < nx; ++i) {
            double s = W->hd[nu + i] * W->DZ[nu + i] + W->g[nu + i];
            for (int m = 0; m < nx; ++m) s += W->A[m * nx + i] * W->PIQ[m];
            if (bnd_active(P, 0, nu + i)) s += W->LAMQ[nz + nu + i] - W->LAMQ[nu + i];
            r[i] = rp[i] = s;
        }
        proj0(P, rp);
        for (int i = 0; i < nx; ++i) W->NU0Q[i] = r[i] - rp[i];
        for (int i = 0; i < nx; ++i)
            W->NUNQ[i] = P->fixedN[i] ? W->PIQ[(N - 1) * nx + i] - W->g[N * nz + nu + i] -
                                            W->hd[N * nz + nu + i] * W->DZ[N * nz + nu + i]
                                      : 0.0;
    }
    if (!ok || mu != mu || rg != rg || rb != rb || rd != rd) return 3;
    if (rg > o->qp_tol_stat || rb > o->qp_tol_eq || rd > o->qp_tol_ineq || rm > o->qp_tol_comp) {
        if (kk >= o->qp_iter_max) return 1;
        return 2;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Merit function (acados ocp_nlp_evaluate_merit_fun [restated]):                                */
/*   cost + sum w_dyn |gap| + sum w_ineq max(0, violation) (+ the eliminated equalities)         */
/* ------------------------------------------------------------------------------------------ */
static double merit(const iocp *P, work *W, const double *X, const double *U) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    double m = total_cost(P, X, U);
    for (int k = 0; k < N; ++k) {
        double phi[NXI];
        integrate(P->n, P->dts, P->h, X + k * nx, U + k * nu, phi, NULL, NULL);
        for (int i = 0; i < nx; ++i) m += W->wdyn[k * nx + i] * fabs(phi[i] - X[(k + 1) * nx + i]);
    }
    for (int k = 0; k <= N; ++k) {
        int sc = stage_class(P, k);
        for (int i = (k == N ? nu : 0); i < nz; ++i) {
            if (!bnd_active(P, k, i)) continue;
            double z = i < nu ? U[k * nu + i] : X[k * nx + i - nu];
            double fl = P->lb[sc][i] - z, fu = z - P->ub[sc][i];
            if (fl > 0) m += W->wb[k * 2 * nz + i] * fl;
            if (fu > 0) m += W->wb[k * 2 * nz + nz + i] * fu;
        }
    }
    double e[NXI];
    eq0_violation(P, X, e);
    for (int i = 0; i < nx; ++i) m += W->w0[i] * fabs(e[i]);
    for (int i = 0; i < nx; ++i)
        if (P->fixedN[i]) m += W->wN[i] * fabs(X[N * nx + i] - P->cN[i]);
    return m;
}

static double line_search(const iocp *P, const orc_opts *o, work *W, int sqp_iter, int *evals) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    /* merit weights from the QP multipliers: first iteration w = |mult|, afterwards
     * w = max(|mult|, (w + |mult|)/2)  (acados ocp_nlp_line_search [restated]) */
#define WUPD(w, a) (w) = sqp_iter == 0 ? (a) : fmax((a), 0.5 * ((w) + (a)))
    for (int c = 0; c < N * nx; ++c) WUPD(W->wdyn[c], fabs(W->PIQ[c]));
    for (int c = 0; c < (N + 1) * 2 * nz; ++c) WUPD(W->wb[c], fabs(W->LAMQ[c]));
    for (int i = 0; i < nx; ++i) WUPD(W->w0[i], fabs(W->NU0Q[i]));
    for (int i = 0; i < nx; ++i) WUPD(W->wN[i], fabs(W->NUNQ[i]));
#undef WUPD
    double m0 = merit(P, W, W->X, W->U);
    double alpha = 1.0;
    for (;;) {
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < nx; ++i)
                W->Xt[k * nx + i] = W->X[k * nx + i] + alpha * W->DZ[k * nz + nu + i];
            if (k < N)
                for (int i = 0; i < nu; ++i)
                    W->Ut[k * nu + i] = W->U[k * nu + i] + alpha * W->DZ[k * nz + i];
        }
        double m1 = merit(P, W, W->Xt, W->Ut);
        (*evals)++;
        if (m1 < m0) break;
        if (alpha * o->alpha_reduction < o->alpha_min) break; /* smallest step is taken anyway */
        alpha *= o->alpha_reduction;
    }
    return alpha;
}

/* ------------------------------------------------------------------------------------------ */
/* SQP / RTI driver (acados ocp_nlp_sqp / ocp_nlp_sqp_rti [restated])                            */
/* ------------------------------------------------------------------------------------------ */
static int sqp(const iocp *P, const orc_opts *o, int mode, work *W, orc_stats *st) {
    int nx = DIM_NX(P), nu = DIM_NU(P), nz = DIM_NZ(P), N = P->N;
    memset(st, 0, sizeof(*st));
    int status = ORC_MAXITER;
    int it = 0;
    int maxit = mode == ORC_MODE_RTI ? 1 : o->max_iter;
    for (;; ++it) {
        linearize(P, W);
        cost_grad_hess(P, o, W);
        nlp_residuals(P, W, &st->res_stat, &st->res_eq, &st->res_ineq, &st->res_comp);
        if (getenv("ORC_DEBUG"))
            fprintf(stderr, "sqp %3d cost %.6f res %.2e %.2e %.2e %.2e\n", it,
                    total_cost(P, W->X, W->U), st->res_stat, st->res_eq, st->res_ineq, st->res_comp);
        if (mode == ORC_MODE_SQP || it > 0) {
            if (st->res_stat != st->res_stat || st->res_eq != st->res_eq) {
                status = ORC_FAILURE;
                break;
            }
            if (mode == ORC_MODE_SQP && st->res_stat < o->tol_stat && st->res_eq < o->tol_eq &&
                st->res_ineq < o->tol_ineq && st->res_comp < o->tol_comp) {
                status = ORC_SUCCESS;
                break;
            }
        }
        if (it >= maxit) {
            status = mode == ORC_MODE_RTI ? ORC_SUCCESS : ORC_MAXITER;
            break;
        }
        int qit = 0;
        int qs = ipm_solve(P, o, W, &qit);
        st->qp_iter += qit;
        st->qp_status = qs;
        st->sqp_iter = it + 1;
        if (qs != 0 && qs != 1) { /* min step / NaN are fatal, max iter is tolerated */
            status = ORC_QP_FAILURE;
            break;
        }
        double alpha = 1.0;
        if (mode == ORC_MODE_SQP && o->globalization) alpha = line_search(P, o, W, it, &st->ls_evals);
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < nx; ++i) W->X[k * nx + i] += alpha * W->DZ[k * nz + nu + i];
            if (k < N)
                for (int i = 0; i < nu; ++i) W->U[k * nu + i] += alpha * W->DZ[k * nz + i];
        }
        for (int c = 0; c < N * nx; ++c) W->PI[c] = (1.0 - alpha) * W->PI[c] + alpha * W->PIQ[c];
        for (int c = 0; c < (N + 1) * 2 * nz; ++c)
            W->LAM[c] = (1.0 - alpha) * W->LAM[c] + alpha * W->LAMQ[c];
    }
    st->status = status;
    st->cost = total_cost(P, W->X, W->U);
    return status;
}

/* ------------------------------------------------------------------------------------------ */
/* Reference-shaped entry points                                                                */
/* ------------------------------------------------------------------------------------------ */
void orc_default_opts(int family, orc_opts *o) {
    memset(o, 0, sizeof(*o));
    o->tol_eq = o->tol_ineq = o->tol_comp = 1e-6; /* acados defaults [restated] */
    o->alpha_min = 0.05, o->alpha_reduction = 0.7; /* acados defaults [restated] */
    o->qp_tol_stat = 1e-6, o->qp_tol_eq = o->qp_tol_ineq = o->qp_tol_comp = 1e-8; /* HPIPM BALANCE */
    o->qp_mu0 = 1e1, o->qp_alpha_min = 1e-12, o->qp_reg_prim = 1e-15;
    o->qp_lam_min = 1e-16, o->qp_t_min = 1e-16, o->qp_tau_min = 1e-16;
    o->eliminate_dt = 1;
    if (family == ORC_FAMILY_VBOC) {
        /* VBOC/triplependulum_class_vboc.py:129-141 */
        o->tol_stat = 1e-3;
        o->qp_tol_stat = 1e-3;
        o->qp_iter_max = 100;
        o->max_iter = 1000;
        o->globalization = 1;
        o->alpha_reduction = 0.3;
        o->alpha_min = 1e-2;
        o->levenberg_marquardt = 1e-5;
    } else {
        /* AL classes set nothing: acados defaults, SQP_RTI, GAUSS_NEWTON, qp iter 50 [restated] */
        o->tol_stat = 1e-6;
        o->qp_iter_max = 50;
        o->max_iter = 100;
        o->globalization = 0;
        o->levenberg_marquardt = 0.0;
    }
}

typedef struct {
    iocp P;
    int nx_ref;
} prep;

/* Translate reference-shaped data into the internal OCP and load the guess into W.
 * Returns NULL for data outside what the reference ever builds (DESIGN.md "boundary"). */
static work *prepare(prep *pp, int n, int family, int N, const double *xg, const double *ug,
                     const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                     const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                     const double *ubu, const double *C0, int ng, double Tf, const orc_opts *o) {
    iocp *P = &pp->P;
    memset(P, 0, sizeof(*P));
    int nxr = 2 * n + (family == ORC_FAMILY_VBOC);
    pp->nx_ref = nxr;
    P->n = n, P->family = family, P->N = N, P->nu = n;
    if (family == ORC_FAMILY_VBOC) {
        /* dt can be dropped iff pinned to one value at every stage and the guess agrees */
        int pinned = o->eliminate_dt;
        double dtv = lbx0[2 * n];
        if (ubx0[2 * n] != dtv || lbx[2 * n] != dtv || ubx[2 * n] != dtv || lbxN[2 * n] != dtv ||
            ubxN[2 * n] != dtv)
            pinned = 0;
        for (int k = 0; k <= N && pinned; ++k)
            if (xg[k * nxr + 2 * n] != dtv) pinned = 0;
        P->dts = !pinned;
        P->h = pinned ? dtv : 1.0;
        P->dt_elim = pinned ? dtv : 0.0;
        for (int i = 0; i < n; ++i) P->w[i] = p[i];
        P->wt = p[n];
    } else {
        P->dts = 0;
        P->h = Tf / N;
    }
    P->nx = 2 * n + P->dts;
    P->nz = P->nx + P->nu;
#ifdef ORC_FIXED_N
    if (n != ORC_FIXED_N || P->dts) return NULL; /* this build serves one dimension set only */
#endif
    int nx = P->nx, nu = P->nu;
    const double *lbs[3] = {lbx0, lbx, lbxN}, *ubs[3] = {ubx0, ubx, ubxN};
    for (int s = 0; s < 3; ++s) {
        for (int i = 0; i < nu; ++i) P->lb[s][i] = lbu[i], P->ub[s][i] = ubu[i];
        for (int i = 0; i < nx; ++i) P->lb[s][nu + i] = lbs[s][i], P->ub[s][nu + i] = ubs[s][i];
    }
    /* stage-0 equalities */
    for (int i = 0; i < nx; ++i) {
        P->fixed0[i] = lbx0[i] == ubx0[i];
        P->c0[i] = P->fixed0[i] ? lbx0[i] : 0.0;
    }
    if (ng) {
        /* only the reference's projector C = [0 | I - d d' | 0], lg = ug = 0 is supported */
        if (ng != n || !C0) return NULL;
        double Md[3][3];
        int jm = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) Md[i][j] = (i == j) - C0[i * nxr + n + j];
        for (int i = 1; i < n; ++i)
            if (Md[i][i] > Md[jm][jm]) jm = i;
        if (!(Md[jm][jm] > 0.0)) return NULL;
        double dj = sqrt(Md[jm][jm]), nrm = 0.0;
        for (int i = 0; i < n; ++i) P->d[i] = Md[i][jm] / dj, nrm += P->d[i] * P->d[i];
        nrm = sqrt(nrm);
        for (int i = 0; i < n; ++i) P->d[i] /= nrm;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < nxr; ++j) {
                double want = (j >= n && j < 2 * n) ? (i == j - n) - P->d[i] * P->d[j - n] : 0.0;
                if (fabs(C0[i * nxr + j] - want) > 1e-9) return NULL;
            }
        for (int i = 0; i < n; ++i)
            if (P->fixed0[n + i]) return NULL;
        P->hasdir = 1;
    }
    P->ny0 = 0;
    for (int i = 0; i < nx; ++i) {
        if (P->fixed0[i]) continue;
        if (P->hasdir && i >= n && i < 2 * n) continue;
        P->Z0[i][P->ny0++] = 1.0;
    }
    if (P->hasdir) {
        for (int i = 0; i < n; ++i) P->Z0[n + i][P->ny0] = P->d[i];
        P->ny0++;
    }
    /* terminal equalities: none, or exactly the n velocities */
    for (int i = 0; i < nx; ++i) {
        P->fixedN[i] = lbxN[i] == ubxN[i];
        P->cN[i] = P->fixedN[i] ? lbxN[i] : 0.0;
        if (P->fixedN[i])
            P->ivN[P->nfN++] = i;
        else
            P->iqN[P->nqN++] = i;
    }
    {
        int velonly = P->nfN == nu;
        for (int i = 0; i < nx; ++i)
            if (P->fixedN[i] && !(i >= n && i < 2 * n)) velonly = 0;
        P->elimN = P->nfN > 0 && velonly;
        P->nb = 0;
        if (!P->elimN)
            for (int i = 0; i < nx; ++i)
                if (P->fixedN[i]) P->bidx[P->nb++] = i;
    }
    work *W = work_alloc(N, nx, nu);
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nx; ++i) W->X[k * nx + i] = xg[k * nxr + i];
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) W->U[k * nu + i] = ug[k * nu + i];
    return W;
}

int orc_solve(int n, int family, int mode, int N, const double *x_guess, const double *u_guess,
              const double *p, const double *lbx0, const double *ubx0, const double *lbx,
              const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
              const double *ubu, const double *C0, int ng, double Tf, const orc_opts *opts,
              double *x, double *u, double *pi, orc_stats *stats) {
    prep pp;
    if (N < 1 || N > 256 || n < 1 || n > 3) return -1;
    work *W = prepare(&pp, n, family, N, x_guess, u_guess, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu,
                      ubu, C0, ng, Tf, opts);
    if (!W) {
        memset(stats, 0, sizeof(*stats));
        stats->status = -2;
        return -2;
    }
    iocp *P = &pp.P;
    int st = sqp(P, opts, mode, W, stats);
    int nx = P->nx, nu = P->nu, nxr = pp.nx_ref;
    for (int k = 0; k <= N; ++k) {
        for (int i = 0; i < nx; ++i) x[k * nxr + i] = W->X[k * nx + i];
        if (nxr > nx) x[k * nxr + nx] = P->dt_elim;
    }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) u[k * nu + i] = W->U[k * nu + i];
    if (pi)
        for (int k = 0; k < N; ++k) {
            for (int i = 0; i < nx; ++i) pi[k * nxr + i] = W->PI[k * nx + i];
            if (nxr > nx) pi[k * nxr + nx] = 0.0;
        }
    work_free(W);
    return st;
}

int orc_solve_batch(int n, int family, int mode, int batch, int Nmax, const int *N,
                    const double *x_guess, const double *u_guess, const double *p,
                    const double *lbx0, const double *ubx0, const double *lbx, const double *ubx,
                    const double *lbxN, const double *ubxN, const double *lbu, const double *ubu,
                    const double *C0, int ng, double Tf, const orc_opts *opts, double *x, double *u,
                    orc_stats *stats, int nthreads) {
    int nxr = 2 * n + (family == ORC_FAMILY_VBOC), nu = n, np = n + 1;
    (void)nthreads;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < batch; ++b) {
        size_t ox = (size_t)b * (Nmax + 1) * nxr, ou = (size_t)b * Nmax * nu;
        orc_solve(n, family, mode, N[b], x_guess + ox, u_guess + ou, p ? p + (size_t)b * np : NULL,
                  lbx0 + (size_t)b * nxr, ubx0 + (size_t)b * nxr, lbx + (size_t)b * nxr,
                  ubx + (size_t)b * nxr, lbxN + (size_t)b * nxr, ubxN + (size_t)b * nxr,
                  lbu + (size_t)b * nu, ubu + (size_t)b * nu,
                  C0 ? C0 + (size_t)b * ng * nxr : NULL, C0 ? ng : 0, Tf, opts, x + ox, u + ou, NULL,
                  stats + b);
    }
    return 0;
}

int orc_first_qp(int n, int family, int N, const double *x_guess, const double *u_guess,
                 const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                 const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                 const double *ubu, const double *C0, int ng, double Tf, const orc_opts *opts,
                 double *A_out, double *B_out, double *b_out, double *dx, double *du, int *qp_iter) {
    prep pp;
    work *W = prepare(&pp, n, family, N, x_guess, u_guess, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu,
                      ubu, C0, ng, Tf, opts);
    if (!W) return -2;
    iocp *P = &pp.P;
    int nx = P->nx, nu = P->nu, nz = P->nz;
    linearize(P, W);
    cost_grad_hess(P, opts, W);
    int st = ipm_solve(P, opts, W, qp_iter);
    memcpy(A_out, W->A, sizeof(double) * N * nx * nx);
    memcpy(B_out, W->B, sizeof(double) * N * nx * nu);
    memcpy(b_out, W->bd, sizeof(double) * N * nx);
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nx; ++i) dx[k * nx + i] = W->DZ[k * nz + nu + i];
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) du[k * nu + i] = W->DZ[k * nz + i];
    work_free(W);
    return st;
}
