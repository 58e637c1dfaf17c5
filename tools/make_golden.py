"""Generate tests/golden/dynamics_golden.npz from the reference's own dynamics text.

Run in the BUILD container only (needs /root/reference; the GPU box does not have it):

    python tools/make_golden.py

For each of the six model classes on the hot path
(`VBOC/{pendulum,doublependulum,triplependulum}_class_vboc.py`,
 `AL/{pendulum,doublependulum,triplependulum}_class_al.py`) the reference file is imported
unmodified under sympy stand-ins for casadi/acados (tools/ref_import.py); `model.f_expl_expr`
is evaluated at seeded random points and differentiated symbolically.  The result pins
 * f(x,u)                      
 * [df/dx df/du]               (what CasADi's forward VDE would integrate)
 * one classical RK4 step of the reference f with h = 1e-2 / step 1.0 on the dt-scaled
   model (the map acados' ERK integrator computes; A4/A8)
against which `tests/test_oracle_dynamics.py` checks the oracle (CPU) and
`tests/test_gpu_*.py` check the CUDA kernels.
"""
import os
import sys

import numpy as np
import sympy as sp

sys.path.insert(0, os.path.dirname(__file__))
import ref_import  # noqa: E402

K = 24  # points per model


def rk4(F, x, u, h):
    k1 = F(x, u)
    k2 = F(x + 0.5 * h * k1, u)
    k3 = F(x + 0.5 * h * k2, u)
    k4 = F(x + h * k3, u)
    return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    exprs = ref_import.dynamics_exprs()
    for name, (f, xs, us) in exprs.items():
        nx, nu = len(xs), len(us)
        n = nu
        scaled = name.startswith("vboc")
        Jx = f.jacobian(xs)
        Ju = f.jacobian(us)
        Ff = sp.lambdify(xs + us, f, "numpy")
        FJx = sp.lambdify(xs + us, Jx, "numpy")
        FJu = sp.lambdify(xs + us, Ju, "numpy")
        X = np.empty((K, nx))
        U = np.empty((K, nu))
        X[:, :n] = rng.uniform(3 * np.pi / 4, 5 * np.pi / 4, (K, n))
        X[:, n:2 * n] = rng.uniform(-10, 10, (K, n))
        if scaled:
            X[:, 2 * n] = rng.uniform(1e-3, 1e-2, K)
            X[0, 2 * n] = 1e-2
        umax = 3.0 if n == 1 else 10.0
        U[:] = rng.uniform(-umax, umax, (K, nu))
        # a few points outside the joint box (the dynamics are defined everywhere)
        X[1, :n] = rng.uniform(-np.pi, np.pi, n)
        fv = np.empty((K, nx))
        jx = np.empty((K, nx, nx))
        ju = np.empty((K, nx, nu))
        xn = np.empty((K, nx))
        for i in range(K):
            args = list(X[i]) + list(U[i])
            fv[i] = np.array(Ff(*args), dtype=float).ravel()
            jx[i] = np.array(FJx(*args), dtype=float)
            ju[i] = np.array(FJu(*args), dtype=float)
            Fw = lambda x, u: np.array(Ff(*x, *u), dtype=float).ravel()  # noqa: E731
            h = 1.0 if scaled else 1e-2
            xn[i] = rk4(Fw, X[i], U[i], h)
        out[f"{name}_x"] = X
        out[f"{name}_u"] = U
        out[f"{name}_f"] = fv
        out[f"{name}_jx"] = jx
        out[f"{name}_ju"] = ju
        out[f"{name}_rk4"] = xn
        print(name, "ok", fv.shape)
    # K1 known answer: gravity-compensation torques of the 2-DOF guess,
    # reference VBOC/doublependulum_vboc.py:84, give zero acceleration at rest.
    dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dynamics_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.abspath(dst))


if __name__ == "__main__":
    main()
