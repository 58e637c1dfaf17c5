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


def test_al_label_batch(oracle, tmp_path):
    from vboc_b200 import problems as pr
    n = 3
    bp = pr.sample_al(n, 96, seed=12)
    X = bp["x0"]
    labels, traj = drivers.al_label_batch(n, X)
    mdl = pr.Model(n)
    outside = np.any(np.abs(X[:, n:]) > mdl.dthetamax, axis=1)
    assert (labels[outside] == 0).all() and outside.any()
    ref = oracle.solve_batch(n, 1, 1, pr.al_problems(n, X[~outside]))
    want = np.where(ref["status"] == 0, 1, np.where(ref["status"] == 4, 0, 2))
    assert (labels[~outside] == want).all()
    assert np.isfinite(traj[labels == 1]).all() and np.isnan(traj[labels == 0]).all()
    p = drivers.save_testdata(n, X[:4], str(tmp_path))
    assert np.load(p).shape == (4, 2 * n)


def test_al_query_selects_most_uncertain():
    import torch
    import torch.nn as nn
    from vboc_b200 import nn as vnn
    torch.manual_seed(0)
    m = nn.Module()
    m.linear_relu_stack = nn.Sequential(nn.Linear(4, 300), nn.ReLU(), nn.Linear(300, 300), nn.ReLU(), nn.Linear(300, 2))
    net = vnn.MLP.from_torch(m)
    pool = np.random.default_rng(1).uniform(-3, 3, (5000, 4)).astype(np.float32)
    idx, etp = drivers.al_query(net, pool, 0.0, 1.0, 50)
    assert len(idx) == 50 and idx == sorted(idx, reverse=True)
    assert etp[idx].min() >= np.sort(etp)[-50] - 1e-7
    net.close()


def test_pendulum_data_generation_matches_oracle_run(oracle):
    ob = OracleBackend(oracle, 1)
    ref = drivers.pendulum_data_generation(backend=(ob, ob.sim))
    out = drivers.pendulum_data_generation()
    assert ref.shape == out.shape and np.abs(ref - out).max() < 1e-5
