"""Host state machines of `testing(v)` / `data_generation(v)` (vboc_b200/drivers.py) with the oracle as the
solver backend: the control flow runs on the CPU here; tests/test_gpu_drivers.py runs the same generators
on the GPU engine and compares the rows."""
import numpy as np
import pytest

from vboc_b200 import drivers, problems as pr


class OracleBackend:
    def __init__(self, oracle, n):
        self.oracle, self.n = oracle, n

    def solve(self, bp, mode):
        return self.oracle.solve_batch(self.n, 0, mode, bp)

    def sim(self, n, X, U, T):
        return np.stack([self.oracle.rk4(n, 1, x, u, T) for x, u in zip(X, U)])


@pytest.fixture
def backend(oracle):
    def make(n):
        b = OracleBackend(oracle, n)
        return b, b.sim
    return make


def test_testing_batch_returns_boundary_points(backend):
    n = 2
    st = {}
    X = drivers.testing_batch(n, 6, seed=3, backend=backend(n), stats=st)
    mdl = pr.Model(n)
    assert X.shape[1] == 4 and X.shape[0] >= 5
    assert (X[:, :n] >= mdl.thetamin - 1e-9).all() and (X[:, :n] <= mdl.thetamax + 1e-9).all()
    assert (np.abs(X[:, n:]) <= mdl.dthetamax + 1e-6).all()
    assert (np.linalg.norm(X[:, n:], axis=1) > 0.1).all()  # a maximal velocity, not the rest state
    assert st["solves"] >= 2 * X.shape[0]  # at least one horizon extension each


def test_data_generation_rows(backend):
    n = 3
    st = {}
    X = drivers.data_generation_batch(n, 4, seed=11, backend=backend(n), stats=st)
    mdl = pr.Model(n)
    assert X.shape[1] == 6 and X.shape[0] >= 4
    assert st["problems_ok"] >= 3
    eps = 1e-2
    inner = X[1:]
    assert (np.abs(X[:, n:]) <= mdl.dthetamax + 1e-6).all()
    assert (X[:, :n] >= mdl.thetamin - 1e-9).all() and (X[:, :n] <= mdl.thetamax + 1e-9).all()
    assert inner.shape[0] > 0 and eps > 0


def test_workers_are_deterministic(backend):
    a = drivers.testing_batch(2, 3, seed=5, backend=backend(2))
    b = drivers.testing_batch(2, 3, seed=5, backend=backend(2))
    assert np.array_equal(a, b)


def test_pendulum_data_generation(backend):
    """configs[0]: the 1-DOF VBOC driver end to end on the oracle backend.  Known answer (K6): every saved
    state lies on the boundary of the viability kernel, i.e. full braking torque from it stops exactly at the
    position limit: checked by integrating the bang-bang arc with the golden-pinned RK4 map."""
    st = {}
    X = drivers.pendulum_data_generation(backend=backend(1), stats=st)
    mdl = pr.Model(1)
    assert X.shape[1] == 2 and X.shape[0] > 40 and st["solves"] >= 2
    assert (X[:, 0] >= mdl.thetamin - 1e-9).all() and (X[:, 0] <= mdl.thetamax + 1e-9).all()
    assert (np.abs(X[:, 1]) <= mdl.dthetamax + 1e-6).all()
    ob, _ = backend(1)
    checked = 0
    for q, v in X[::7]:
        if abs(abs(v) - mdl.dthetamax) < 1e-3 or abs(v) < 0.5:
            continue  # on the velocity limit (boundary of X, not of V) or nearly at rest
        x = np.array([q, v])
        u = np.array([mdl.umax if v < 0 else -mdl.umax])  # brake
        for _ in range(4000):
            xn = ob.oracle.rk4(1, 1, x, u, 1e-3)
            if xn[1] * v <= 0:
                break
            x = xn
        lim = mdl.thetamin if v < 0 else mdl.thetamax
        assert abs(x[0] - lim) < 5e-3, (q, v, x)
        checked += 1
    assert checked >= 3
