"""Known-answer checks of the oracle's solver (the reference holds no test or golden vector for the
solver part, SURVEY 8(c): PARITY UNPINNED; these are the self-made K4/K5 checks).

 * a converged VBOC solution satisfies the NLP constraints it was asked to satisfy, checked with
   nothing but the (golden-pinned) RK4 map;
 * on a short horizon the optimal cost agrees with scipy's SLSQP on the same NLP;
 * AL labels: status 0  <=>  the linearised QP is feasible, checked with scipy's HiGHS LP solver.
"""
import numpy as np
import pytest
from scipy.optimize import linprog, minimize

from vboc_b200 import problems as pr


def test_vboc_solution_is_feasible(oracle):
    n = 3
    bp = pr.sample_vboc(n, 6, seed=21)
    r = oracle.solve_batch(n, 0, 0, bp)
    assert (r["status"] == 0).all()
    for b in range(6):
        x, u, N = r["x"][b], r["u"][b], int(bp["N"][b])
        for k in range(N):
            xn = oracle.rk4(n, 0, x[k], u[k])
            assert np.abs(xn - x[k + 1]).max() < 1e-6                      # shooting gaps (tol_eq)
        assert (x[:, :n] > bp["lbx"][b, :n] - 1e-6).all() and (x[:, :n] < bp["ubx"][b, :n] + 1e-6).all()
        assert np.abs(x[:, n:2 * n]).max() < 10 + 1e-6 and np.abs(u).max() < 10 + 1e-6
        assert np.abs(x[N, n:2 * n]).max() < 1e-6                          # v_N = 0
        d = bp["p"][b, :n]
        v0 = x[0, n:2 * n]
        assert np.abs(v0 - d * (d @ v0)).max() < 1e-6                      # v_0 parallel to d
        fixed = bp["lbx0"][b, :n] == bp["ubx0"][b, :n]
        assert np.abs(x[0, :n] - bp["lbx0"][b, :n])[fixed].max() < 1e-6
        assert abs(r["cost"][b] - d @ v0) < 1e-9                            # get_cost() = w . v_0
        assert r["cost"][b] < 0                                            # v_0 points along -d * |v|


def test_short_horizon_cost_matches_slsqp(oracle):
    """n = 2, N = 6: same NLP through scipy SLSQP (decision variables: all states and controls)."""
    n, N, h = 2, 6, 1e-2
    bp = pr.sample_testdata(n, 3, seed=4, N=N)
    oo = oracle.default_opts(0)
    oo.tol_stat = 1e-8
    oo.qp_tol_stat = 1e-9
    r = oracle.solve_batch(n, 0, 0, bp, oo)
    nx, nu = 2 * n, n
    for b in range(3):
        if r["status"][b] != 0:
            continue
        d, q0 = bp["p"][b, :n], bp["lbx0"][b, :n]
        Qc = np.linalg.svd(d[None, :])[2][1:]  # rows spanning the complement of d (full-rank form of C0)

        def unpack(z):
            return z[:(N + 1) * nx].reshape(N + 1, nx), z[(N + 1) * nx:].reshape(N, nu)

        def gaps(z):
            X, U = unpack(z)
            g = [X[0, :n] - q0, X[N, n:]]
            v0 = X[0, n:]
            g.append(Qc @ v0)
            for k in range(N):
                xk = np.concatenate([X[k], [h]])
                g.append(oracle.rk4(n, 0, xk, U[k])[:nx] - X[k + 1])
            return np.concatenate(g)

        lb = np.concatenate([np.tile(np.concatenate([bp["lbx"][b, :n], [-10] * n]), N + 1), np.full(N * nu, -10.0)])
        ub = np.concatenate([np.tile(np.concatenate([bp["ubx"][b, :n], [10] * n]), N + 1), np.full(N * nu, 10.0)])
        z0 = np.concatenate([r["x"][b, :N + 1, :nx].ravel() * 0 + np.tile(np.concatenate([q0, [0] * n]), N + 1),
                             np.zeros(N * nu)])
        res = minimize(lambda z: d @ unpack(z)[0][0, n:], z0, method="SLSQP", bounds=list(zip(lb, ub)),
                       constraints=[{"type": "eq", "fun": gaps}], options={"maxiter": 500, "ftol": 1e-12})
        assert res.success, res.message
        assert abs(res.fun - r["cost"][b]) < 1e-5, (res.fun, r["cost"][b])


@pytest.mark.parametrize("n", [1, 2])
def test_al_labels_equal_lp_feasibility(oracle, n):
    """compute_problem returns 1 iff the RTI QP solved (AL/triplependulum_class_al.py:164-169); the QP
    is feasible iff the LP over the same linearised dynamics and boxes is."""
    N = 100
    bp = pr.sample_al(n, 24, seed=8, N=N)
    r = oracle.solve_batch(n, 1, 1, bp)
    nx, nu = 2 * n, n
    nvar = (N + 1) * nx + N * nu
    agree = 0
    for b in range(24):
        q = oracle.first_qp(n, 1, pr.take(bp, b))
        A, B, bb = q["A"], q["B"], q["b"]
        xg = bp["x_guess"][b]
        rows, rhs = [], []
        for k in range(N):
            for i in range(nx):
                row = np.zeros(nvar)
                row[(k + 1) * nx + i] = 1.0
                row[k * nx:(k + 1) * nx] -= A[k][i]
                row[(N + 1) * nx + k * nu:(N + 1) * nx + (k + 1) * nu] -= B[k][i]
                rows.append(row)
                rhs.append(bb[k][i])
        lo = np.concatenate([np.tile(bp["lbx"][b], N + 1) - xg.ravel(), np.tile(bp["lbu"][b], N)])
        hi = np.concatenate([np.tile(bp["ubx"][b], N + 1) - xg.ravel(), np.tile(bp["ubu"][b], N)])
        lo[:nx] = hi[:nx] = bp["lbx0"][b] - xg[0]                # x_0 fixed (the guess has v = 0)
        lo[N * nx + n:(N + 1) * nx] = hi[N * nx + n:(N + 1) * nx] = -xg[N, n:]  # v_N = 0
        res = linprog(np.zeros(nvar), A_eq=np.array(rows), b_eq=np.array(rhs), bounds=list(zip(lo, hi)),
                      method="highs")
        feasible = res.status == 0
        agree += int(feasible == (r["status"][b] == 0))
    assert agree >= 23, agree  # allow one borderline (feasible set of measure ~ solver tolerance)
