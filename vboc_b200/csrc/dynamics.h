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
        T M[NQ][NQ], r[NQ], v2[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            T si, ci;
            sincos_t(q[i], si, ci);
            v2[i] = v[i] * v[i];
            r[i] = u[i] - (PendN::m * (NQ - i) * PendN::g * PendN::l) * si;
            M[i][i] = constant<T>(PendN::m * (NQ - i) * ll);
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
#pragma unroll
            for (int j = i + 1; j < NQ; ++j) {
                T s, c;
                sincos_t(q[i] - q[j], s, c);
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

}  // namespace vboc
