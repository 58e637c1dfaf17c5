"""The batched drivers on the GPU engine against the same generators on the oracle."""
import numpy as np
import pytest

from vboc_b200 import drivers

pytestmark = pytest.mark.gpu


class OracleBackend:
    def __init__(self, oracle, n):
        self.oracle, self.n = oracle, n

    def solve(self, bp, mode):
        return self.oracle.solve_batch(self.n, 0, mode, bp)

    def sim(self, n, X, U, T):
        return np.stack([self.oracle.rk4(n, 1, x, u, T) for x, u in zip(X, U)])


@pytest.mark.parametrize("n", [2, 3])
def test_testing_batch_matches_oracle_run(oracle, n):
    ob = OracleBackend(oracle, n)
    ref = drivers.testing_batch(n, 8, seed=2, backend=(ob, ob.sim))
    out = drivers.testing_batch(n, 8, seed=2)
    assert ref.shape == out.shape
    assert np.abs(ref - out).max() < 1e-5


def test_data_generation_matches_oracle_run(oracle):
    n = 3
    ob = OracleBackend(oracle, n)
    s1, s2 = {}, {}
    ref = drivers.data_generation_batch(n, 6, seed=4, backend=(ob, ob.sim), stats=s1)
    out = drivers.data_generation_batch(n, 6, seed=4, stats=s2)
    assert s1["solves"] == s2["solves"] and s1["problems_ok"] == s2["problems_ok"]
    assert ref.shape == out.shape
    assert np.abs(ref - out).max() < 1e-5
