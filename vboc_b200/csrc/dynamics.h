// dynamics.h -- n-link pendulum dynamics, forward-mode tangent and the ERK4 step.
//
// Replaces the CasADi-generated C (f_expl, forward VDE) and acados' sim_erk for the models of
//   VBOC/pendulum_class_vboc.py:35-39        (n = 1: m g d sin(q) + F - b v over d^2 m)
//   VBOC/doublependulum_class_vboc.py:40-91  (n = 2)
//   VBOC/triplependulum_class_vboc.py:47-58  (n = 3)
// and their AL twins (AL/*_class_al.py).  The reference writes the accelerations as expanded closed
// forms; here they are evaluated in manipulator form  M(q) a = u - c(q, v) - G(q)  for point masses
// m_i at the tips of massless links l_i in ABSOLUTE link angles (theta = pi hanging):
//     M_ij = mu_ij l_i l_j cos(q_i - q_j),  mu_ij = sum_{k >= max(i,j)} m_k
//     c_i  = sum_j mu_ij l_i l_j sin(q_i - q_j) v_j^2,     G_i = mu_ii g l_i sin(q_i)
// which equals the reference expressions up to rounding (tests/golden/dynamics_golden.npz).
//
// The scalar type T is `double` (plain evaluation: merit function, simulator) or `Dual1` (value +
// one tangent: each lane of the linearisation kernel carries one column of [B A]).
#pragma once
#include "warp_spmd.h"

namespace vboc {

struct Dual1 {
    double v, d;
};
VB_HD Dual1 operator+(Dual1 a, Dual1 b) { return {a.v + b.v, a.d + b.d}; }
VB_HD Dual1 operator-(Dual1 a, Dual1 b) { return {a.v - b.v, a.d - b.d}; }
VB_HD Dual1 operator-(Dual1 a) { return {-a.v, -a.d}; }
VB_HD Dual1 operator*(Dual1 a, Dual1 b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
VB_HD Dual1 operator*(double s, Dual1 a) { return {s * a.v, s * a.d}; }

VB_HD double recip(double a) { return 1.0 / a; }
VB_HD Dual1 recip(Dual1 a) {
    double r = 1.0 / a.v;
    return {r, -a.d * r * r};
}
VB_HD void sincos_t(double a, double &s, double &c) {
#if defined(__CUDA_ARCH__)
    sincos(a, &s, &c);
#else
    s = sin(a), c = cos(a);
#endif
}
VB_HD void sincos_t(Dual1 a, Dual1 &s, Dual1 &c) {
    double sv, cv;
    sincos_t(a.v, sv, cv);
    s = {sv, cv * a.d};
    c = {cv, -sv * a.d};
}
template <class T> VB_HD T constant(double c);
template <> VB_HD double constant<double>(double c) { return c; }
template <> VB_HD Dual1 constant<Dual1>(double c) { return {c, 0.0}; }
VB_HD double value_of(double a) { return a; }
VB_HD double value_of(Dual1 a) { return a.v; }
VB_HD double tangent_of(double) { return 0.0; }
VB_HD double tangent_of(Dual1 a) { return a.d; }

// model constants (constructor blocks of the reference classes)
struct Pend1 {  // VBOC/pendulum_class_vboc.py:14-17
    static constexpr double m = 0.5, g = 9.81, d = 0.3, b = 0.01;
};
struct PendN {  // VBOC/triplependulum_class_vboc.py:15-21, VBOC/doublependulum_class_vboc.py:14-18
    static constexpr double m = 0.4, l = 0.8, g = 9.81;
};

// joint accelerations a(q, v, u)
template <int NQ, class T>
VB_HD void accel(const T *q, const T *v, const T *u, T *a) {
    if constexpr (NQ == 1) {
        T s, c;
        sincos_t(q[0], s, c);
        T num = (Pend1::m * Pend1::g * Pend1::d) * s + (u[0] - Pend1::b * v[0]);
        a[0] = (1.0 / (Pend1::d * Pend1::d * Pend1::m)) * num;
    } else {
        constexpr double ll = PendN::l * PendN::l;
        T M[NQ][NQ], r[NQ], v2[NQ], sq[NQ], cq[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            sincos_t(q[i], sq[i], cq[i]);
            v2[i] = v[i] * v[i];
            r[i] = u[i] - (PendN::m * (NQ - i) * PendN::g * PendN::l) * sq[i];
            M[i][i] = constant<T>(PendN::m * (NQ - i) * ll);
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
#pragma unroll
            for (int j = i + 1; j < NQ; ++j) {
                // sin / cos of the angle difference from the addition theorems: n sincos evaluations per
                // model call instead of n (n + 1) / 2 (each is ~80 instructions of kernel code)
                T s = sq[i] * cq[j] - cq[i] * sq[j];
                T c = cq[i] * cq[j] + sq[i] * sq[j];
                const double cf = PendN::m * (NQ - j) * ll;  // mu_ij = sum_{k >= j} m_k for j > i
                M[i][j] = cf * c;
                M[j][i] = M[i][j];
                r[i] = r[i] - cf * (s * v2[j]);
                r[j] = r[j] + cf * (s * v2[i]);
            }
        }
        // Gaussian elimination without pivoting (M is SPD)
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            T inv = recip(M[k][k]);
#pragma unroll
            for (int i = k + 1; i < NQ; ++i) {
                T f = M[i][k] * inv;
#pragma unroll
                for (int j = k + 1; j < NQ; ++j) M[i][j] = M[i][j] - f * M[k][j];
                r[i] = r[i] - f * r[k];
            }
            M[k][k] = inv;  // keep the reciprocal pivot
        }
#pragma unroll
        for (int i = NQ - 1; i >= 0; --i) {
            T s = r[i];
#pragma unroll
            for (int j = i + 1; j < NQ; ++j) s = s - M[i][j] * a[j];
            a[i] = s * M[i][i];
        }
    }
}

// One classical RK4 step of length h of xdot = [v; a(q, v, u)], x = [q; v] (acados sim_erk with
// 4 stages, 1 step; the VBOC models' dt-scaled dynamics over a unit step with dt pinned is the
// same map with h = dt).
template <int NQ, class T>
VB_HD void rk4_step(const T *x, const T *u, double h, T *xn) {
    constexpr int NX = 2 * NQ;
    T k[NX], xt[NX], acc[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xt[i] = x[i], acc[i] = constant<T>(0.0);
    // the four stages share one copy of the model code (instruction-cache footprint)
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        accel<NQ, T>(xt, xt + NQ, u, k + NQ);
#pragma unroll
        for (int i = 0; i < NQ; ++i) k[i] = xt[NQ + i];
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;  // k1 + 2 k2 + 2 k3 + k4
        const double adv = st == 2 ? h : 0.5 * h;             // x + h/2 k1, x + h/2 k2, x + h k3
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            acc[i] = acc[i] + wgt * k[i];
            xt[i] = x[i] + adv * k[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = x[i] + (h / 6.0) * acc[i];
}


// ---------------------------------------------------------------------------------------------------
// Analytic Jacobian of the accelerations (one thread evaluates a whole interval in the lane-per-OCP
// kernel, where carrying nz forward tangents through the model would not fit the register file).
//   M a = r  =>  da/dtheta = M^-1 (dr/dtheta - (dM/dtheta) a)
// Ja is [NQ][3 NQ] in column order (q, v, u).
template <int NQ>
VB_HD void accel_jac(const double *q, const double *v, const double *u, double *a, double (*Ja)[3 * NQ]) {
    if constexpr (NQ == 1) {
        double s, c;
        sincos_t(q[0], s, c);
        const double inv = 1.0 / (Pend1::d * Pend1::d * Pend1::m);
        a[0] = inv * ((Pend1::m * Pend1::g * Pend1::d) * s + (u[0] - Pend1::b * v[0]));
        Ja[0][0] = inv * (Pend1::m * Pend1::g * Pend1::d) * c;
        Ja[0][1] = -inv * Pend1::b;
        Ja[0][2] = inv;
    } else {
        constexpr double ll = PendN::l * PendN::l;
        double M[NQ][NQ], S[NQ][NQ], K[NQ][NQ], r[NQ], cq[NQ], v2[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            double si;
            sincos_t(q[i], si, cq[i]);
            v2[i] = v[i] * v[i];
            r[i] = u[i] - (PendN::m * (NQ - i) * PendN::g * PendN::l) * si;
            M[i][i] = PendN::m * (NQ - i) * ll;
            S[i][i] = 0.0, K[i][i] = 1.0;
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i)
#pragma unroll
            for (int j = i + 1; j < NQ; ++j) {
                double s, c;
                sincos_t(q[i] - q[j], s, c);
                const double cf = PendN::m * (NQ - j) * ll;
                S[i][j] = cf * s, S[j][i] = -cf * s;  // c_ij sin(q_i - q_j)
                K[i][j] = cf * c, K[j][i] = cf * c;   // c_ij cos(q_i - q_j)
                M[i][j] = cf * c, M[j][i] = cf * c;
                r[i] -= S[i][j] * v2[j];
                r[j] -= S[j][i] * v2[i];
            }
        // LU of M without pivoting (SPD); reciprocal pivots kept on the diagonal
        double Lf[NQ][NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i)
#pragma unroll
            for (int j = 0; j < NQ; ++j) Lf[i][j] = M[i][j];
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            Lf[k][k] = 1.0 / Lf[k][k];
#pragma unroll
            for (int i = k + 1; i < NQ; ++i) {
                Lf[i][k] *= Lf[k][k];
#pragma unroll
                for (int j = k + 1; j < NQ; ++j) Lf[i][j] -= Lf[i][k] * Lf[k][j];
            }
        }
        auto solve = [&](double *b) {  // in place: b <- M^-1 b
#pragma unroll
            for (int i = 1; i < NQ; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) b[i] -= Lf[i][j] * b[j];
#pragma unroll
            for (int i = NQ - 1; i >= 0; --i) {
#pragma unroll
                for (int j = i + 1; j < NQ; ++j) b[i] -= Lf[i][j] * b[j];
                b[i] *= Lf[i][i];
            }
        };
#pragma unroll
        for (int i = 0; i < NQ; ++i) a[i] = r[i];
        solve(a);
        // columns of D = dr/dtheta - (dM/dtheta) a, then Ja = M^-1 D
#pragma unroll
        for (int m = 0; m < NQ; ++m) {
            double dq[NQ], dv[NQ], du[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
                if (i == m) {
                    double t = -(PendN::m * (NQ - i) * PendN::g * PendN::l) * cq[i];
#pragma unroll
                    for (int j = 0; j < NQ; ++j)
                        if (j != i) t -= K[i][j] * v2[j] - S[i][j] * a[j];
                    dq[i] = t;
                    dv[i] = 0.0;
                } else {
                    dq[i] = K[i][m] * v2[m] - S[i][m] * a[m];
                    dv[i] = -2.0 * S[i][m] * v[m];
                }
                du[i] = (i == m) ? 1.0 : 0.0;
            }
            solve(dq), solve(dv), solve(du);
#pragma unroll
            for (int i = 0; i < NQ; ++i) Ja[i][m] = dq[i], Ja[i][NQ + m] = dv[i], Ja[i][2 * NQ + m] = du[i];
        }
    }
}

// One RK4 step with forward sensitivities: xn = Phi_h(x, u) and Phi = d xn / d z, z = [u; x]
// ([2 NQ][3 NQ], column order (u, q, v)) -- what acados' ERK integrator returns as [B A].
template <int NQ>
VB_HD void rk4_sens(const double *x, const double *u, double h, double *xn, double (*Phi)[3 * NQ]) {
    constexpr int NX = 2 * NQ, NZ = 3 * NQ;
    double xt[NX], acc[NX], St[NX][NZ];  // stage point and its sensitivity d xt / d z
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        xt[i] = x[i], acc[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NZ; ++j) St[i][j] = (j == NQ + i) ? 1.0 : 0.0, Phi[i][j] = 0.0;
    }
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        double a[NQ], Ja[NQ][3 * NQ], Kx[NX], Ks[NX][NZ];
        accel_jac<NQ>(xt, xt + NQ, u, a, Ja);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            Kx[i] = xt[NQ + i], Kx[NQ + i] = a[i];
#pragma unroll
            for (int j = 0; j < NZ; ++j) {
                Ks[i][j] = St[NQ + i][j];  // d v / d z
                double t = (j < NQ) ? Ja[i][2 * NQ + j] : 0.0;  // direct dependence on u
#pragma unroll
                for (int m = 0; m < NX; ++m) t += Ja[i][m] * St[m][j];
                Ks[NQ + i][j] = t;
            }
        }
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0, adv = st == 2 ? h : 0.5 * h;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            acc[i] += wgt * Kx[i];
            xt[i] = x[i] + adv * Kx[i];
#pragma unroll
            for (int j = 0; j < NZ; ++j) {
                Phi[i][j] += wgt * Ks[i][j];
                St[i][j] = ((j == NQ + i) ? 1.0 : 0.0) + adv * Ks[i][j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        xn[i] = x[i] + (h / 6.0) * acc[i];
#pragma unroll
        for (int j = 0; j < NZ; ++j) Phi[i][j] = ((j == NQ + i) ? 1.0 : 0.0) + (h / 6.0) * Phi[i][j];
    }
}

// The VBOC models with a FREE dt state (VBOC/pendulum_class_vboc.py:35-39): x = [q; v; dt],
// xdot = dt [v; a(q, v, u); 0], integrated over a unit step.  xn and Phi = d xn / d [u; x]
// ([2 NQ + 1][3 NQ + 1], column order (u, q, v, dt)).
template <int NQ>
VB_HD void rk4_sens_dts(const double *x, const double *u, double *xn, double (*Phi)[3 * NQ + 1]) {
    constexpr int NX = 2 * NQ + 1, NZ = 3 * NQ + 1;
    const double h = 1.0;
    double xt[NX], acc[NX], St[NX][NZ];
    for (int i = 0; i < NX; ++i) {
        xt[i] = x[i], acc[i] = 0.0;
        for (int j = 0; j < NZ; ++j) St[i][j] = (j == NQ + i) ? 1.0 : 0.0, Phi[i][j] = 0.0;
    }
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        double a[NQ], Ja[NQ][3 * NQ], Kx[NX], Ks[NX][NZ];
        accel_jac<NQ>(xt, xt + NQ, u, a, Ja);
        const double dt = xt[2 * NQ];
        for (int i = 0; i < NQ; ++i) {
            Kx[i] = dt * xt[NQ + i], Kx[NQ + i] = dt * a[i];
            for (int j = 0; j < NZ; ++j) {
                // d(dt v_i)/dz = dt dv_i/dz + v_i ddt/dz ; d(dt a_i)/dz = dt da_i/dz + a_i ddt/dz
                double dv = St[NQ + i][j], da = (j < NQ) ? Ja[i][2 * NQ + j] : 0.0;
                for (int m = 0; m < 2 * NQ; ++m) da += Ja[i][m] * St[m][j];
                const double ddt = St[2 * NQ][j];
                Ks[i][j] = dt * dv + xt[NQ + i] * ddt;
                Ks[NQ + i][j] = dt * da + a[i] * ddt;
            }
        }
        Kx[2 * NQ] = 0.0;
        for (int j = 0; j < NZ; ++j) Ks[2 * NQ][j] = 0.0;
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0, adv = st == 2 ? h : 0.5 * h;
        for (int i = 0; i < NX; ++i) {
            acc[i] += wgt * Kx[i];
            xt[i] = x[i] + adv * Kx[i];
            for (int j = 0; j < NZ; ++j) {
                Phi[i][j] += wgt * Ks[i][j];
                St[i][j] = ((j == NQ + i) ? 1.0 : 0.0) + adv * Ks[i][j];
            }
        }
    }
    for (int i = 0; i < NX; ++i) {
        xn[i] = x[i] + (h / 6.0) * acc[i];
        for (int j = 0; j < NZ; ++j) Phi[i][j] = ((j == NQ + i) ? 1.0 : 0.0) + (h / 6.0) * Phi[i][j];
    }
}

// value only (merit function)
template <int NQ>
VB_HD void rk4_step_dts(const double *x, const double *u, double *xn) {
    constexpr int NX = 2 * NQ + 1;
    double xt[NX], acc[NX], k[NX];
    for (int i = 0; i < NX; ++i) xt[i] = x[i], acc[i] = 0.0;
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        double a[NQ];
        accel<NQ, double>(xt, xt + NQ, u, a);
        for (int i = 0; i < NQ; ++i) k[i] = xt[2 * NQ] * xt[NQ + i], k[NQ + i] = xt[2 * NQ] * a[i];
        k[2 * NQ] = 0.0;
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0, adv = st == 2 ? 1.0 : 0.5;
        for (int i = 0; i < NX; ++i) acc[i] += wgt * k[i], xt[i] = x[i] + adv * k[i];
    }
    for (int i = 0; i < NX; ++i) xn[i] = x[i] + acc[i] / 6.0;
}

}  // namespace vboc
