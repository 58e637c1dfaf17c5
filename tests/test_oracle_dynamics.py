"""The oracle's dynamics against the golden vectors generated from the reference's own expression
text (tools/make_golden.py, tests/golden/dynamics_golden.npz): f, [df/dx df/du] and one RK4 step for
the six model classes on the hot path."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dynamics_golden.npz"))
CASES = [("vboc", n) for n in (1, 2, 3)] + [("al", n) for n in (1, 2, 3)]


@pytest.mark.parametrize("fam,n", CASES)
def test_f_and_jacobians(oracle, fam, n):
    family = 0 if fam == "vboc" else 1
    key = f"{fam}{n}"
    X, U = G[key + "_x"], G[key + "_u"]
    for i in range(X.shape[0]):
        f = oracle.f(n, family, X[i], U[i])
        jx, ju = oracle.f_jac(n, family, X[i], U[i])
        assert np.allclose(f, G[key + "_f"][i], rtol=1e-11, atol=1e-11)
        assert np.allclose(jx, G[key + "_jx"][i], rtol=1e-10, atol=1e-10)
        assert np.allclose(ju, G[key + "_ju"][i], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("fam,n", CASES)
def test_rk4_step(oracle, fam, n):
    family = 0 if fam == "vboc" else 1
    key = f"{fam}{n}"
    X, U = G[key + "_x"], G[key + "_u"]
    for i in range(X.shape[0]):
        xn = oracle.rk4(n, family, X[i], U[i], 1e-2)
        assert np.allclose(xn, G[key + "_rk4"][i], rtol=1e-12, atol=1e-12)


def test_rk4_jacobians_match_finite_differences(oracle):
    X, U = G["al3_x"], G["al3_u"]
    x, u = X[0], U[0]
    _, A, B = oracle.rk4(3, 1, x, u, 1e-2, jac=True)
    eps = 1e-6
    for j in range(6):
        e = np.zeros(6); e[j] = eps
        fd = (oracle.rk4(3, 1, x + e, u, 1e-2) - oracle.rk4(3, 1, x - e, u, 1e-2)) / (2 * eps)
        assert np.allclose(A[:, j], fd, atol=1e-7)
    for j in range(3):
        e = np.zeros(3); e[j] = eps
        fd = (oracle.rk4(3, 1, x, u + e, 1e-2) - oracle.rk4(3, 1, x, u - e, 1e-2)) / (2 * eps)
        assert np.allclose(B[:, j], fd, atol=1e-7)


def test_gravity_compensation_equilibrium(oracle):
    """K1: u = g l (sum_{k>=i} m_k) sin(q_i), v = 0 => zero acceleration (VBOC/doublependulum_vboc.py:84)."""
    from vboc_b200.problems import Model
    rng = np.random.default_rng(1)
    for n in (2, 3):
        for _ in range(5):
            q = rng.uniform(2.4, 3.9, n)
            x = np.concatenate([q, np.zeros(n)])
            f = oracle.f(n, 1, x, Model(n).gravity_comp(q))
            assert np.abs(f).max() < 1e-12
