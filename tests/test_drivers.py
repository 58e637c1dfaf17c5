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


class FakeStream:
    """The ticket-queue protocol of `engine.StreamSolver` served by the oracle: finished tickets are handed
    out late and out of order, at most a few per poll, with a tiny slot pool -- the event loop must cope."""

    def __init__(self, oracle, n, slots=3):
        self.oracle, self.n, self.slots = oracle, n, slots
        self.results, self.order, self.next_ticket, self.polls = {}, [], 0, 0
        self.rng = np.random.default_rng(0)

    @property
    def free_slots(self):
        return self.slots - len(self.results)

    def submit(self, bp, mode=0):
        B = len(bp["N"])
        assert 1 <= B <= self.free_slots
        out = self.oracle.solve_batch(self.n, 0, mode, bp)
        tk = []
        for b in range(B):
            t = self.next_ticket
            self.next_ticket += 1
            N = int(bp["N"][b])
            self.results[t] = dict(status=int(out["status"][b]), cost=float(out["cost"][b]),
                                   x=out["x"][b, :N + 1].copy(), u=out["u"][b, :N].copy())
            self.order.append(t)
            tk.append(t)
        return np.array(tk, dtype=np.int32)

    def poll(self):
        self.polls += 1
        if self.polls % 2 or not self.order:
            return []
        self.rng.shuffle(self.order)
        k = int(self.rng.integers(1, 3))
        out, self.order = self.order[:k], self.order[k:]
        return out

    def fetch(self, t):
        return self.results.pop(t)

    def sim_step(self, X, U, T):
        return np.stack([self.oracle.rk4(self.n, 1, x, u, T) for x, u in zip(X, U)])


def test_stream_event_loop_equals_rounds(oracle, backend):
    """`run_workers_stream` (tickets, out-of-order completion, slot back-pressure) returns exactly what the
    round-synchronous `run_workers` returns for the same workers."""
    n = 2
    st_r, st_s = {}, {}
    a = drivers.data_generation_batch(n, 5, seed=21, backend=backend(n), stats=st_r)
    b = drivers.data_generation_stream(n, 5, seed=21, ssol=FakeStream(oracle, n), stats=st_s)
    assert np.array_equal(a, b)
    assert st_r["solves"] == st_s["solves"] and st_r["converged"] == st_s["converged"]
    c = drivers.testing_batch(n, 4, seed=2, backend=backend(n))
    d = drivers.testing_stream(n, 4, seed=2, ssol=FakeStream(oracle, n, slots=2))
    assert np.array_equal(c, d)


def test_generic_vboc_worker_ends_on_a_failed_solve():
    """`VBOC/vboc.py:162-213`: the generic driver has its restart block commented out -- a failed solve ends the problem --
    while the per-system drivers restart with a perturbed direction and positions (`VBOC/triplependulum_vboc.py:138-174`)."""
    from types import SimpleNamespace
    n = 2
    failed = SimpleNamespace(status=4, cost=0.0, x=None, u=None)
    g = drivers.data_generation_worker(n, drivers._rng(1, 0), restarts=False)
    first = next(g)
    with pytest.raises(StopIteration) as stop:
        g.send(failed)
    assert stop.value.value is None
    g = drivers.data_generation_worker(n, drivers._rng(1, 0))
    first2 = next(g)
    again = g.send(failed)                  # the per-system worker asks for another solve ...
    assert np.array_equal(first.p, first2.p) and again.N == first2.N
    assert not np.array_equal(again.p, first2.p) and abs(np.linalg.norm(again.p[:n]) - 1.0) < 1e-12   # ... of a perturbed problem
