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
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n, family, N;
    int dts;        /* dt is a state (VBOC with free dt, or eliminate_dt == 0) */
    int nx, nu, nz; /* internal dims; z = [u; x] (HPIPM ordering)              */
    int ng;
    double h;        /* RK4 step: 1.0 when dts (scaled model), pinned dt or Tf/N otherwise */
    double w[3], wt; /* VBOC linear cost: w . v_0 + wt * sum_{k<N} dt_k                     */
    double dt_elim;  /* value of the eliminated dt (for cost reporting), 0 for AL          */
    /* bounds in z ordering for the three stage classes 0, 1..N-1, N */
    double lb[3][NZI], ub[3][NZI];
    int fixed0[NXI]; /* stage-0 state components with lb == ub: removed from the QP
                        (exact; acados does the same for constraints.x0 via idxbxe_0 [restated]) */
    double C0[NGI][NXI];
} iocp;

static inline int stage_class(const iocp *P, int k) { return k == 0 ? 0 : (k == P->N ? 2 : 1); }

/* is z-component i of stage k an inequality handled by the IPM? */
static inline int bnd_active(const iocp *P, int k, int i) {
    if (k == P->N && i < P->nu) return 0;
    if (k == 0 && i >= P->nu && P->fixed0[i - P->nu]) return 0;
    return 1;
}

typedef struct {
    int N, nx, nu, nz, ng;
    /* NLP iterate */
    double *X, *U, *PI;  /* (N+1)nx, N nu, N nx                                         */
    double *LAM;         /* (N+1) * 2nz : [lower nz | upper nz] in z ordering            */
    double LAMG[2 * NGI];
    /* linearisation */
    double *A, *B, *bd;  /* N*nx*nx, N*nx*nu, N*nx  (bd = phi(x_k,u_k) - x_{k+1})        */
    double *g, *hd;      /* (N+1)*nz cost gradient / Hessian diagonal (incl. LM)        */
    /* QP solution */
    double *DZ, *PIQ, *LAMQ, *TQ; /* (N+1)nz, N nx, (N+1)2nz, (N+1)2nz                    */
    double LAMGQ[2 * NGI], TGQ[2 * NGI];
    /* IPM work */
    double *L;                  /* N * nz*nz lower Cholesky factors of M_k                 */
    double *LN;                 /* nx: sqrt of the terminal diagonal                       */
    double *pv, *yv;            /* (N+1)nx, N nu : Riccati vectors                         */
    double *rg, *rb, *rd, *rm;  /* residuals: (N+1)nz, N nx, (N+1)2nz, (N+1)2nz           */
    double *dv, *dpi, *dlam, *dt; /* step                                                 */
    double *Gam, *gam;          /* (N+1)2nz                                                */
    double *rmb;                /* backup of res_m                                         */
    double rdg[2 * NGI], rmg[2 * NGI], dlamg[2 * NGI], dtg[2 * NGI], Gamg[2 * NGI], gamg[2 * NGI],
        rmbg[2 * NGI];
    /* merit */
    double *wdyn, *wb; /* N nx, (N+1) 2nz */
    double wg[2 * NGI];
    double *Xt, *Ut; /* trial point */
} work;

static work *work_alloc(int N, int nx, int nu, int ng) {
    work *W = (work *)calloc(1, sizeof(work));
    int nz = nx + nu, S = N + 1;
    W->N = N, W->nx = nx, W->nu = nu, W->nz = nz, W->ng = ng;
#define AL_(ptr, cnt) W->ptr = (double *)calloc((size_t)(cnt), sizeof(double))
    AL_(X, S * nx), AL_(U, S * nu), AL_(PI, S * nx), AL_(LAM, S * 2 * nz);
    AL_(A, S * nx * nx), AL_(B, S * nx * nu), AL_(bd, S * nx), AL_(g, S * nz), AL_(hd, S * nz);
    AL_(DZ, S * nz), AL_(PIQ, S * nx), AL_(LAMQ, S * 2 * nz), AL_(TQ, S * 2 * nz);
    AL_(L, S * nz * nz), AL_(LN, nx), AL_(pv, S * nx), AL_(yv, S * nu);
    AL_(rg, S * nz), AL_(rb, S * nx), AL_(rd, S * 2 * nz), AL_(rm, S * 2 * nz);
    AL_(dv, S * nz), AL_(dpi, S * nx), AL_(dlam, S * 2 * nz), AL_(dt, S * 2 * nz);
    AL_(Gam, S * 2 * nz), AL_(gam, S * 2 * nz), AL_(rmb, S * 2 * nz);
    AL_(wdyn, S * nx), AL_(wb, S * 2 * nz), AL_(Xt, S * nx), AL_(Ut, S * nu);
#undef AL_
    return W;
}
static void work_free(work *W) {
    double **ps[] = {&W->X,  &W->U,   &W->PI,  &W->LAM, &W->A,    &W->B,  &W->bd, &W->g,   &W->hd,
                     &W->DZ, &W->PIQ, &W->LAMQ, &W->TQ, &W->L,    &W->LN, &W->pv, &W->yv,  &W->rg,
                     &W->rb, &W->rd,  &W->rm,  &W->dv,  &W->dpi,  &W->dlam, &W->dt, &W->Gam, &W->gam,
                     &W->rmb, &W->wdyn, &W->wb, &W->Xt, &W->Ut};
    for (size_t i = 0; i < sizeof(ps) / sizeof(ps[0]); ++i) free(*ps[i]);
    free(W);
}

/* ------------------------------------------------------------------------------------------ */
/* Stage cost (acados cost modules EXTERNAL / LINEAR_LS scaled by the time step [restated])     */
/* ------------------------------------------------------------------------------------------ */
static double total_cost(const iocp *P, const double *X, const double *U) {
    int n = P->n, nx = P->nx, N = P->N;
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
    int n = P->n, nx = P->nx, nu = P->nu, nz = P->nz, N = P->N;
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
    int nx = P->nx, nu = P->nu, N = P->N;
    for (int k = 0; k < N; ++k) {
        double phi[NXI];
        integrate(P->n, P->dts, P->h, W->X + k * nx, W->U + k * nu, phi, W->A + k * nx * nx,
                  W->B + k * nx * nu);
        for (int i = 0; i < nx; ++i) W->bd[k * nx + i] = phi[i] - W->X[(k + 1) * nx + i];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* NLP residuals (acados ocp_nlp_res_compute [restated]): inf-norms of                          */
/*   stat : gradient of the Lagrangian,  eq : shooting gaps,                                     */
/*   ineq : constraint violation,        comp : multiplier * constraint function                */
/* ------------------------------------------------------------------------------------------ */
static void nlp_residuals(const iocp *P, work *W, double *rs, double *re, double *ri, double *rc) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    double s = 0, e = 0, in = 0, c = 0;
    for (int k = 0; k <= N; ++k) {
        int sc = stage_class(P, k);
        const double *lam = W->LAM + k * 2 * nz;
        double r[NZI];
        for (int i = 0; i < nz; ++i) r[i] = W->g[k * nz + i];
        if (k < N) {
            const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *pi = W->PI + k * nx;
            for (int j = 0; j < nu; ++j)
                for (int i = 0; i < nx; ++i) r[j] += B[i * nu + j] * pi[i];
            for (int j = 0; j < nx; ++j)
                for (int i = 0; i < nx; ++i) r[nu + j] += A[i * nx + j] * pi[i];
        }
        if (k > 0)
            for (int j = 0; j < nx; ++j) r[nu + j] -= W->PI[(k - 1) * nx + j];
        if (k == 0 && ng) {
            for (int a = 0; a < ng; ++a)
                for (int j = 0; j < nx; ++j) r[nu + j] += P->C0[a][j] * (W->LAMG[ng + a] - W->LAMG[a]);
        }
        for (int i = (k == N ? nu : 0); i < nz; ++i) {
            double z = i < nu ? W->U[k * nu + i] : W->X[k * nx + i - nu];
            double fl = P->lb[sc][i] - z, fu = z - P->ub[sc][i];
            r[i] += lam[nz + i] - lam[i];
            if (fabs(r[i]) > s || r[i] != r[i]) s = fabs(r[i]);
            if (fl > in) in = fl;
            if (fu > in) in = fu;
            if (fabs(lam[i] * fl) > c) c = fabs(lam[i] * fl);
            if (fabs(lam[nz + i] * fu) > c) c = fabs(lam[nz + i] * fu);
        }
        if (k < N)
            for (int i = 0; i < nx; ++i) {
                double v = fabs(W->bd[k * nx + i]);
                if (v > e || v != v) e = v;
            }
    }
    for (int a = 0; a < ng; ++a) {
        double v = 0;
        for (int j = 0; j < nx; ++j) v += P->C0[a][j] * W->X[j];
        if (-v > in) in = -v; /* lg - Cx, lg = 0 */
        if (v > in) in = v;   /* Cx - ug, ug = 0 */
        if (fabs(W->LAMG[a] * v) > c) c = fabs(W->LAMG[a] * v);
        if (fabs(W->LAMG[ng + a] * v) > c) c = fabs(W->LAMG[ng + a] * v);
    }
    *rs = s, *re = e, *ri = in, *rc = c;
}

/* ------------------------------------------------------------------------------------------ */
/* Riccati factorisation / solves (HPIPM square-root backward recursion [restated]).            */
/*                                                                                              */
/* LQ sub-problem in the step dz_k = [du_k; dx_k]:                                               */
/*   min sum_k 1/2 dz'diag(hh_k)dz + r_k'dz  (+ stage-0 general-constraint block)                */
/*   s.t. dx_{k+1} = A_k dx_k + B_k du_k + beta_k,   dx_0 components in fixed0 are 0            */
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

/* hh: (N+1)*nz effective Hessian diagonal (hd + Gamma_l + Gamma_u); gg: 2*ng Gamma of the stage-0
 * general constraint. */
static void riccati_factor(const iocp *P, work *W, const double *hh, const double *gg) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    for (int i = 0; i < nx; ++i) W->LN[i] = sqrt(hh[N * nz + nu + i]);
    for (int k = N - 1; k >= 0; --k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu;
        double Wm[NXI][NZI]; /* Lp' [B A] */
        for (int j = 0; j < nz; ++j)
            for (int i = 0; i < nx; ++i) {
                double s = 0.0;
                if (k == N - 1) {
                    s = W->LN[i] * (j < nu ? B[i * nu + j] : A[i * nx + j - nu]);
                } else {
                    const double *Ln = W->L + (k + 1) * nz * nz;
                    for (int m = i; m < nx; ++m)
                        s += Ln[(nu + m) * nz + nu + i] * (j < nu ? B[m * nu + j] : A[m * nx + j - nu]);
                }
                Wm[i][j] = s;
            }
        double *M = W->L + k * nz * nz;
        for (int i = 0; i < nz; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = (i == j) ? hh[k * nz + i] : 0.0;
                for (int m = 0; m < nx; ++m) s += Wm[m][i] * Wm[m][j];
                M[i * nz + j] = s;
            }
        if (k == 0) {
            for (int a = 0; a < ng; ++a) {
                double G = gg[a] + gg[ng + a];
                for (int i = 0; i < nx; ++i)
                    for (int j = 0; j <= i; ++j)
                        M[(nu + i) * nz + nu + j] += G * P->C0[a][i] * P->C0[a][j];
            }
            for (int f = 0; f < nx; ++f)
                if (P->fixed0[f]) {
                    int r = nu + f;
                    for (int j = 0; j < r; ++j) M[r * nz + j] = 0.0;
                    for (int i = r + 1; i < nz; ++i) M[i * nz + r] = 0.0;
                    M[r * nz + r] = 1.0;
                }
        }
        chol_lower(M, nz, nz);
    }
}

/* Solve with rhs r ((N+1)*nz gradient) and beta (N*nx); outputs dv ((N+1)*nz) and dpi (N*nx). */
static void riccati_solve(const iocp *P, work *W, const double *r, const double *beta, double *dv,
                          double *dpi) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N;
    double *pv = W->pv, *yv = W->yv;
    for (int i = 0; i < nx; ++i) pv[N * nx + i] = r[N * nz + nu + i];
    for (int k = N - 1; k >= 0; --k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *L = W->L + k * nz * nz;
        double t1[NXI], t2[NXI], m[NZI];
        /* t2 = P+ beta + p+ */
        if (k == N - 1) {
            for (int i = 0; i < nx; ++i)
                t2[i] = W->LN[i] * W->LN[i] * beta[k * nx + i] + pv[(k + 1) * nx + i];
        } else {
            const double *Ln = W->L + (k + 1) * nz * nz;
            for (int i = 0; i < nx; ++i) {
                double s = 0.0;
                for (int mm = i; mm < nx; ++mm) s += Ln[(nu + mm) * nz + nu + i] * beta[k * nx + mm];
                t1[i] = s;
            }
            for (int i = 0; i < nx; ++i) {
                double s = pv[(k + 1) * nx + i];
                for (int mm = 0; mm <= i; ++mm) s += Ln[(nu + i) * nz + nu + mm] * t1[mm];
                t2[i] = s;
            }
        }
        for (int j = 0; j < nz; ++j) {
            double s = r[k * nz + j];
            for (int i = 0; i < nx; ++i) s += (j < nu ? B[i * nu + j] : A[i * nx + j - nu]) * t2[i];
            m[j] = s;
        }
        if (k == 0)
            for (int f = 0; f < nx; ++f)
                if (P->fixed0[f]) m[nu + f] = 0.0;
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
    /* x0 = -Lxx^-T Lxx^-1 p0 */
    {
        const double *L = W->L;
        double y[NXI], x0[NXI];
        for (int i = 0; i < nx; ++i) {
            double s = pv[i];
            for (int j = 0; j < i; ++j) s -= L[(nu + i) * nz + nu + j] * y[j];
            double d = L[(nu + i) * nz + nu + i];
            y[i] = d > 0.0 ? s / d : 0.0;
        }
        for (int i = nx - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < nx; ++j) s -= L[(nu + j) * nz + nu + i] * x0[j];
            double d = L[(nu + i) * nz + nu + i];
            x0[i] = d > 0.0 ? s / d : 0.0;
        }
        for (int i = 0; i < nx; ++i) dv[nu + i] = -x0[i];
    }
    for (int k = 0; k < N; ++k) {
        const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *L = W->L + k * nz * nz;
        double *z = dv + k * nz, *zn = dv + (k + 1) * nz;
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
        for (int i = 0; i < nx; ++i) {
            double s = beta[k * nx + i];
            for (int j = 0; j < nx; ++j) s += A[i * nx + j] * z[nu + j];
            for (int j = 0; j < nu; ++j) s += B[i * nu + j] * z[j];
            zn[nu + i] = s;
        }
        if (k == N - 1)
            for (int i = 0; i < nu; ++i) zn[i] = 0.0;
        /* dpi_k = P_{k+1} dx_{k+1} + p_{k+1} */
        if (k == N - 1) {
            for (int i = 0; i < nx; ++i)
                dpi[k * nx + i] = W->LN[i] * W->LN[i] * zn[nu + i] + pv[(k + 1) * nx + i];
        } else {
            const double *Ln = W->L + (k + 1) * nz * nz;
            double t1[NXI];
            for (int i = 0; i < nx; ++i) {
                double s = 0.0;
                for (int mm = i; mm < nx; ++mm) s += Ln[(nu + mm) * nz + nu + i] * zn[nu + mm];
                t1[i] = s;
            }
            for (int i = 0; i < nx; ++i) {
                double s = pv[(k + 1) * nx + i];
                for (int mm = 0; mm <= i; ++mm) s += Ln[(nu + i) * nz + nu + mm] * t1[mm];
                dpi[k * nx + i] = s;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* IPM (HPIPM d_ocp_qp_ipm_solve, BALANCE mode, pred_corr = cond_pred_corr = 1, cold start,     */
/* no iterative refinement / LQ fall-back [restated]).                                          */
/* Returns 0 success, 1 max iter, 2 min step, 3 NaN.                                            */
/* ------------------------------------------------------------------------------------------ */
static double qp_residuals(const iocp *P, work *W, const double *lbd, const double *ubd,
                           const double *gd, double *ng_, double *nb_, double *nd_, double *nm_) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    double rg = 0, rb = 0, rd = 0, rm = 0, mu = 0;
    int nc = 0;
    for (int k = 0; k <= N; ++k) {
        const double *v = W->DZ + k * nz, *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
        double *r = W->rg + k * nz;
        for (int i = 0; i < nz; ++i) r[i] = W->hd[k * nz + i] * v[i] + W->g[k * nz + i];
        if (k < N) {
            const double *A = W->A + k * nx * nx, *B = W->B + k * nx * nu, *pi = W->PIQ + k * nx;
            for (int j = 0; j < nu; ++j)
                for (int i = 0; i < nx; ++i) r[j] += B[i * nu + j] * pi[i];
            for (int j = 0; j < nx; ++j)
                for (int i = 0; i < nx; ++i) r[nu + j] += A[i * nx + j] * pi[i];
            const double *vn = W->DZ + (k + 1) * nz;
            for (int i = 0; i < nx; ++i) {
                double s = W->bd[k * nx + i] - vn[nu + i];
                for (int j = 0; j < nx; ++j) s += A[i * nx + j] * v[nu + j];
                for (int j = 0; j < nu; ++j) s += B[i * nu + j] * v[j];
                W->rb[k * nx + i] = s;
                if (fabs(s) > rb || s != s) rb = fabs(s);
            }
        }
        if (k > 0)
            for (int j = 0; j < nx; ++j) r[nu + j] -= W->PIQ[(k - 1) * nx + j];
        if (k == 0 && ng)
            for (int a = 0; a < ng; ++a)
                for (int j = 0; j < nx; ++j)
                    r[nu + j] += P->C0[a][j] * (W->LAMGQ[ng + a] - W->LAMGQ[a]);
        for (int i = 0; i < nz; ++i) {
            if (!bnd_active(P, k, i)) {
                W->rd[k * 2 * nz + i] = W->rd[k * 2 * nz + nz + i] = 0.0;
                W->rm[k * 2 * nz + i] = W->rm[k * 2 * nz + nz + i] = 0.0;
                if (k == 0 && i >= nu && P->fixed0[i - nu]) r[i] = 0.0;
                if (k == N && i < nu) r[i] = 0.0;
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
        for (int i = 0; i < nz; ++i)
            if (fabs(r[i]) > rg || r[i] != r[i]) rg = fabs(r[i]);
    }
    for (int a = 0; a < ng; ++a) {
        double s = 0;
        for (int j = 0; j < nx; ++j) s += P->C0[a][j] * W->DZ[nu + j];
        double dl = gd[a] - s + W->TGQ[a], du = s - gd[ng + a] + W->TGQ[ng + a];
        W->rdg[a] = dl, W->rdg[ng + a] = du;
        W->rmg[a] = W->LAMGQ[a] * W->TGQ[a], W->rmg[ng + a] = W->LAMGQ[ng + a] * W->TGQ[ng + a];
        if (fabs(dl) > rd || dl != dl) rd = fabs(dl);
        if (fabs(du) > rd || du != du) rd = fabs(du);
        if (fabs(W->rmg[a]) > rm) rm = fabs(W->rmg[a]);
        if (fabs(W->rmg[ng + a]) > rm) rm = fabs(W->rmg[ng + a]);
        mu += W->rmg[a] + W->rmg[ng + a];
        nc += 2;
    }
    *ng_ = rg, *nb_ = rb, *nd_ = rd, *nm_ = rm;
    return nc ? mu / nc : 0.0;
}

/* Build Gamma/gamma from (lam, t, res_d, res_m), factorise if asked, solve, then recover dlam, dt
 * and the maximum step alpha. */
static __thread double g_reg_prim = 0.0;
static double ipm_step(const iocp *P, work *W, int factor, const double *rm, const double *rmg) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    double *hh = W->Gam; /* reuse: Gam holds Gamma per constraint; effective Hessian goes to gam+.. */
    static __thread double hheff[(256 + 1) * NZI];
    static __thread double rr[(256 + 1) * NZI];
    (void)hh;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            double h = W->hd[k * nz + i] + g_reg_prim, r = W->rg[k * nz + i];
            if (bnd_active(P, k, i)) {
                const double *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
                double Gl = lam[i] / t[i], Gu = lam[nz + i] / t[nz + i];
                double gl = (rm[k * 2 * nz + i] - lam[i] * W->rd[k * 2 * nz + i]) / t[i];
                double gu = (rm[k * 2 * nz + nz + i] - lam[nz + i] * W->rd[k * 2 * nz + nz + i]) / t[nz + i];
                h += Gl + Gu;
                r += gl - gu;
            }
            hheff[k * nz + i] = h;
            rr[k * nz + i] = r;
        }
    for (int a = 0; a < ng; ++a) {
        W->Gamg[a] = W->LAMGQ[a] / W->TGQ[a];
        W->Gamg[ng + a] = W->LAMGQ[ng + a] / W->TGQ[ng + a];
        W->gamg[a] = (rmg[a] - W->LAMGQ[a] * W->rdg[a]) / W->TGQ[a];
        W->gamg[ng + a] = (rmg[ng + a] - W->LAMGQ[ng + a] * W->rdg[ng + a]) / W->TGQ[ng + a];
        for (int j = 0; j < nx; ++j) rr[nu + j] += P->C0[a][j] * (W->gamg[a] - W->gamg[ng + a]);
    }
    if (factor) riccati_factor(P, W, hheff, W->Gamg);
    riccati_solve(P, W, rr, W->rb, W->dv, W->dpi);
    /* the Riccati solves the step equations with rhs = -(residual): dv is the minimiser of
     * 1/2 dv'H dv + rr'dv, i.e. already the Newton step. */
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
    if (ng) {
        double cs[NGI];
        for (int a = 0; a < ng; ++a) {
            cs[a] = 0;
            for (int j = 0; j < nx; ++j) cs[a] += P->C0[a][j] * W->dv[nu + j];
        }
        for (int a = 0; a < 2 * ng; ++a) {
            double dtt = (a < ng ? cs[a] : -cs[a - ng]) - W->rdg[a];
            double dl = -(rmg[a] + W->LAMGQ[a] * dtt) / W->TGQ[a];
            W->dtg[a] = dtt, W->dlamg[a] = dl;
            if (dtt < 0.0 && -W->TGQ[a] / dtt < alpha) alpha = -W->TGQ[a] / dtt;
            if (dl < 0.0 && -W->LAMGQ[a] / dl < alpha) alpha = -W->LAMGQ[a] / dl;
        }
    }
    return alpha;
}

static double mu_aff(const iocp *P, work *W, double alpha) {
    int nz = P->nz, N = P->N, ng = P->ng, nc = 0;
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
    for (int a = 0; a < 2 * ng; ++a) {
        mu += (W->LAMGQ[a] + alpha * W->dlamg[a]) * (W->TGQ[a] + alpha * W->dtg[a]);
        nc++;
    }
    return nc ? mu / nc : 0.0;
}

static int ipm_solve(const iocp *P, const orc_opts *o, work *W, int *iters) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    static __thread double lbd[(256 + 1) * NZI], ubd[(256 + 1) * NZI];
    double gd[2 * NGI];
    g_reg_prim = o->qp_reg_prim;
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
    for (int a = 0; a < ng; ++a) {
        double s = 0;
        for (int j = 0; j < nx; ++j) s += P->C0[a][j] * W->X[j];
        gd[a] = -s, gd[ng + a] = -s; /* lg - C x, ug - C x with lg = ug = 0 */
    }
    /* cold start */
    memset(W->DZ, 0, sizeof(double) * (N + 1) * nz);
    memset(W->PIQ, 0, sizeof(double) * N * nx);
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nz; ++i) {
            double *v = W->DZ + k * nz + i, *lam = W->LAMQ + k * 2 * nz, *t = W->TQ + k * 2 * nz;
            if (k == 0 && i >= nu && P->fixed0[i - nu]) {
                *v = lbd[i]; /* eliminated component: the step is known */
                lam[i] = lam[nz + i] = t[i] = t[nz + i] = 0.0;
                continue;
            }
            if (!bnd_active(P, k, i)) {
                lam[i] = lam[nz + i] = t[i] = t[nz + i] = 0.0;
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
    for (int a = 0; a < ng; ++a) {
        double s = 0;
        for (int j = 0; j < nx; ++j) s += P->C0[a][j] * W->DZ[nu + j];
        double tl = s - gd[a], tu = gd[ng + a] - s;
        W->TGQ[a] = tl > thr0 ? tl : thr0;
        W->TGQ[ng + a] = tu > thr0 ? tu : thr0;
        W->LAMGQ[a] = o->qp_mu0 / W->TGQ[a];
        W->LAMGQ[ng + a] = o->qp_mu0 / W->TGQ[ng + a];
    }
    double rg, rb, rd, rm, alpha = 1.0;
    double mu = qp_residuals(P, W, lbd, ubd, gd, &rg, &rb, &rd, &rm);
    int kk = 0;
    for (; kk < o->qp_iter_max && alpha > o->qp_alpha_min &&
           (rg > o->qp_tol_stat || rb > o->qp_tol_eq || rd > o->qp_tol_ineq || rm > o->qp_tol_comp);
         ++kk) {
        /* affine (predictor) direction: res_m = lam * t */
        memcpy(W->rmb, W->rm, sizeof(double) * (N + 1) * 2 * nz);
        memcpy(W->rmbg, W->rmg, sizeof(W->rmg));
        double a_aff = ipm_step(P, W, 1, W->rm, W->rmg);
        double m_aff = mu_aff(P, W, a_aff);
        double sigma = m_aff / mu;
        sigma = sigma * sigma * sigma;
        double sm = sigma * mu;
        if (sm < o->qp_tau_min) sm = o->qp_tau_min;
        /* centering + corrector: res_m = lam*t + dt_aff*dlam_aff - sigma*mu */
        for (int c = 0; c < (N + 1) * 2 * nz; ++c)
            W->rm[c] = W->rmb[c] != 0.0 || W->TQ[c] != 0.0 ? W->rmb[c] + W->dt[c] * W->dlam[c] - sm : 0.0;
        for (int a = 0; a < 2 * ng; ++a) W->rmg[a] = W->rmbg[a] + W->dtg[a] * W->dlamg[a] - sm;
        alpha = ipm_step(P, W, 0, W->rm, W->rmg);
        /* conditional predictor-corrector: if the corrected step is much worse than the affine
         * one, fall back to the centering direction only (HPIPM cond_pred_corr [restated]). */
        double m_cor = mu_aff(P, W, alpha);
        if (m_cor > 2.0 * m_aff) {
            for (int c = 0; c < (N + 1) * 2 * nz; ++c)
                W->rm[c] = W->rmb[c] != 0.0 || W->TQ[c] != 0.0 ? W->rmb[c] - sm : 0.0;
            for (int a = 0; a < 2 * ng; ++a) W->rmg[a] = W->rmbg[a] - sm;
            alpha = ipm_step(P, W, 0, W->rm, W->rmg);
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
        for (int a = 0; a < 2 * ng; ++a) {
            W->LAMGQ[a] += as * W->dlamg[a];
            W->TGQ[a] += as * W->dtg[a];
            if (W->LAMGQ[a] < o->qp_lam_min) W->LAMGQ[a] = o->qp_lam_min;
            if (W->TGQ[a] < o->qp_t_min) W->TGQ[a] = o->qp_t_min;
        }
        mu = qp_residuals(P, W, lbd, ubd, gd, &rg, &rb, &rd, &rm);
        if (getenv("ORC_DEBUG"))
            fprintf(stderr, "  ipm %3d a_aff %.3e alpha %.3e sigma %.2e mu %.3e rg %.2e rb %.2e rd %.2e rm %.2e\n", kk,
                    a_aff, alpha, sigma, mu, rg, rb, rd, rm);
    }
    *iters = kk;
    /* multipliers of the eliminated stage-0 components from stationarity (what HPIPM's
     * restore_eq_dof does for removed equality bounds [restated]) */
    for (int f = 0; f < nx; ++f)
        if (P->fixed0[f]) {
            int i = nu + f;
            double r = W->hd[i] * W->DZ[i] + W->g[i];
            const double *A = W->A;
            for (int m = 0; m < nx; ++m) r += A[m * nx + f] * W->PIQ[m];
            for (int a = 0; a < ng; ++a) r += P->C0[a][f] * (W->LAMGQ[ng + a] - W->LAMGQ[a]);
            W->LAMQ[i] = r > 0 ? r : 0.0;
            W->LAMQ[nz + i] = r < 0 ? -r : 0.0;
        }
    if (mu != mu || rg != rg || rb != rb || rd != rd) return 3;
    if (kk >= o->qp_iter_max &&
        (rg > o->qp_tol_stat || rb > o->qp_tol_eq || rd > o->qp_tol_ineq || rm > o->qp_tol_comp))
        return 1;
    if (alpha <= o->qp_alpha_min) return 2;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Merit function (acados ocp_nlp_evaluate_merit_fun [restated]):                                */
/*   cost + sum w_dyn |gap| + sum w_ineq max(0, violation)                                       */
/* ------------------------------------------------------------------------------------------ */
static double merit(const iocp *P, work *W, const double *X, const double *U) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    double m = total_cost(P, X, U);
    for (int k = 0; k < N; ++k) {
        double phi[NXI];
        integrate(P->n, P->dts, P->h, X + k * nx, U + k * nu, phi, NULL, NULL);
        for (int i = 0; i < nx; ++i) m += W->wdyn[k * nx + i] * fabs(phi[i] - X[(k + 1) * nx + i]);
    }
    for (int k = 0; k <= N; ++k) {
        int sc = stage_class(P, k);
        for (int i = (k == N ? nu : 0); i < nz; ++i) {
            double z = i < nu ? U[k * nu + i] : X[k * nx + i - nu];
            double fl = P->lb[sc][i] - z, fu = z - P->ub[sc][i];
            if (fl > 0) m += W->wb[k * 2 * nz + i] * fl;
            if (fu > 0) m += W->wb[k * 2 * nz + nz + i] * fu;
        }
    }
    for (int a = 0; a < ng; ++a) {
        double v = 0;
        for (int j = 0; j < nx; ++j) v += P->C0[a][j] * X[j];
        if (-v > 0) m += W->wg[a] * (-v);
        if (v > 0) m += W->wg[ng + a] * v;
    }
    return m;
}

static double line_search(const iocp *P, const orc_opts *o, work *W, int sqp_iter, int *evals) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    /* merit weights from the QP multipliers: first iteration w = |mult|, afterwards
     * w = max(|mult|, (w + |mult|)/2)  (acados ocp_nlp_line_search [restated]) */
    for (int c = 0; c < N * nx; ++c) {
        double a = fabs(W->PIQ[c]);
        W->wdyn[c] = sqp_iter == 0 ? a : fmax(a, 0.5 * (W->wdyn[c] + a));
    }
    for (int c = 0; c < (N + 1) * 2 * nz; ++c) {
        double a = fabs(W->LAMQ[c]);
        W->wb[c] = sqp_iter == 0 ? a : fmax(a, 0.5 * (W->wb[c] + a));
    }
    for (int a = 0; a < 2 * ng; ++a) {
        double v = fabs(W->LAMGQ[a]);
        W->wg[a] = sqp_iter == 0 ? v : fmax(v, 0.5 * (W->wg[a] + v));
    }
    double m0 = merit(P, W, W->X, W->U);
    double alpha = 1.0;
    for (; alpha * o->alpha_reduction > o->alpha_min;) {
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
        alpha *= o->alpha_reduction;
    }
    return alpha;
}

/* ------------------------------------------------------------------------------------------ */
/* SQP / RTI driver (acados ocp_nlp_sqp / ocp_nlp_sqp_rti [restated])                            */
/* ------------------------------------------------------------------------------------------ */
static int sqp(const iocp *P, const orc_opts *o, int mode, work *W, orc_stats *st) {
    int nx = P->nx, nu = P->nu, nz = P->nz, N = P->N, ng = P->ng;
    memset(st, 0, sizeof(*st));
    int status = ORC_MAXITER;
    int it = 0;
    int maxit = mode == ORC_MODE_RTI ? 1 : o->max_iter;
    for (;; ++it) {
        linearize(P, W);
        cost_grad_hess(P, o, W);
        nlp_residuals(P, W, &st->res_stat, &st->res_eq, &st->res_ineq, &st->res_comp);
        if (mode == ORC_MODE_SQP) {
            if (st->res_stat != st->res_stat || st->res_eq != st->res_eq) {
                status = ORC_FAILURE;
                break;
            }
            if (st->res_stat < o->tol_stat && st->res_eq < o->tol_eq && st->res_ineq < o->tol_ineq &&
                st->res_comp < o->tol_comp) {
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
        for (int a = 0; a < 2 * ng; ++a) W->LAMG[a] = (1.0 - alpha) * W->LAMG[a] + alpha * W->LAMGQ[a];
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
    o->qp_mu0 = 1e1, o->qp_alpha_min = 1e-12, o->qp_reg_prim = 1e-13;
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

/* Translate reference-shaped data into the internal OCP and load the guess into W. */
static work *prepare(prep *pp, int n, int family, int N, const double *xg, const double *ug,
                     const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                     const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                     const double *ubu, const double *C0, int ng, double Tf, const orc_opts *o) {
    iocp *P = &pp->P;
    memset(P, 0, sizeof(*P));
    int nxr = 2 * n + (family == ORC_FAMILY_VBOC);
    pp->nx_ref = nxr;
    P->n = n, P->family = family, P->N = N, P->nu = n, P->ng = ng;
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
    int nx = P->nx, nu = P->nu;
    const double *lbs[3] = {lbx0, lbx, lbxN}, *ubs[3] = {ubx0, ubx, ubxN};
    for (int s = 0; s < 3; ++s) {
        for (int i = 0; i < nu; ++i) P->lb[s][i] = lbu[i], P->ub[s][i] = ubu[i];
        for (int i = 0; i < nx; ++i) P->lb[s][nu + i] = lbs[s][i], P->ub[s][nu + i] = ubs[s][i];
    }
    for (int i = 0; i < nx; ++i) P->fixed0[i] = lbx0[i] == ubx0[i];
    for (int a = 0; a < ng; ++a)
        for (int j = 0; j < nx; ++j) P->C0[a][j] = C0[a * nxr + j];
    work *W = work_alloc(N, nx, nu, ng);
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
