"""Host logic of the active-learning loop (vboc_b200/al_loop.py) with a synthetic labeller: windows, pool
bookkeeping, stopping rule, the guess-network plumbing.  The GPU run of the same loop is in tests/test_gpu_al_loop.py."""
import numpy as np
import torch

from vboc_b200 import al_loop


class Net(torch.nn.Module):
    def __init__(self, n_in, hidden, n_out):
        super().__init__()
        self.linear_relu_stack = torch.nn.Sequential(torch.nn.Linear(n_in, hidden), torch.nn.ReLU(),
                                                     torch.nn.Linear(hidden, hidden), torch.nn.ReLU(),
                                                     torch.nn.Linear(hidden, n_out))

    def forward(self, x):
        return self.linear_relu_stack(x)


def test_active_learning_bookkeeping():
    torch.manual_seed(0)
    n, N, nx = 1, 4, 2
    rng = np.random.default_rng(0)
    pool = rng.uniform(-1, 1, size=(200, nx))
    mean, std = torch.tensor(0.0), torch.tensor(1.0)
    calls = []

    def label_fn(X, xg):  # viable iff |v| < 0.5; trajectory = the state repeated
        calls.append((len(X), None if xg is None else xg.shape))
        lab = (np.abs(X[:, 1]) < 0.5).astype(np.int64)
        traj = np.repeat(X[:, None, :], N + 1, axis=1)
        if xg is not None:
            assert np.allclose(xg[:, 0], X)
        return lab, traj

    def query_fn(model, pool, B):  # "entropy" = closeness to the decision boundary
        etp = 1.0 - np.abs(np.abs(pool[:, 1]) - 0.5)
        idx = np.argpartition(etp, -B)[-B:].tolist()
        idx.sort(reverse=True)
        return idx, etp

    model, guess = Net(nx, 8, 2), Net(nx, 8, N * nx)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    optg = torch.optim.Adam(guess.parameters(), lr=1e-2)
    fit_cls = lambda m, Xi: al_loop.fit_minibatch(m, opt, torch.nn.BCEWithLogitsLoss(), Xi[:, :nx], Xi[:, nx:], mean, std,
                                                  n_minibatch=16, it_max=5)
    fit_guess = lambda m, Xt: al_loop.fit_minibatch(m, optg, torch.nn.MSELoss(), Xt[:, :nx], Xt[:, nx:], mean, std,
                                                    n_minibatch=16, it_max=5, normalize_targets=True)
    hist = []
    Xi, Xt, rest = al_loop.active_learning(n, pool, 40, 20, model, guess, mean, std, fit_cls, fit_guess, etp_stop=0.0,
                                           max_rounds=3, N=N, label_fn=label_fn, query_fn=query_fn, history=hist)
    assert len(hist) == 3 and [h["labelled"] for h in hist] == [20, 20, 20]
    assert len(rest) == 200 - 40 - 60
    assert Xi.shape == (40, nx + 2)  # sliding window: 40 initial rows, 20 out / 20 in per round
    assert set(map(tuple, Xi[:, nx:])) <= {(0.0, 1.0), (1.0, 0.0)}
    assert Xt.shape[1] == (N + 1) * nx  # the reference's X_traj layout: the flattened trajectory, stage 0 first
    assert calls[0] == (40, None) and calls[1][1] == (20, N + 1, nx)  # rounds use the guess network's trajectory
    # the most uncertain states were taken: what is left is farther from the boundary than what was labelled last
    assert np.abs(np.abs(rest[:, 1]) - 0.5).min() >= np.abs(np.abs(Xi[-20:, 1]) - 0.5).max() - 1e-12


def test_predict_guess_shapes():
    guess = Net(4, 8, 3 * 4)
    X = np.random.default_rng(1).normal(size=(5, 4))
    xg = al_loop.predict_guess(guess, X, torch.tensor(0.5), torch.tensor(2.0), 3, 4)
    assert xg.shape == (5, 4, 4) and np.array_equal(xg[:, 0], X)


def test_label2_samples_produce_no_rows():
    X = np.arange(12, dtype=float).reshape(6, 2)
    labels = np.array([1, 0, 2, 1, 2, 0])
    traj = np.repeat(X[:, None, :], 3, axis=1)
    it, tr, dropped = al_loop._rows(X, labels, traj)
    assert it.shape == (4, 4) and tr.shape == (2, 6) and dropped == 2
    assert np.array_equal(it[:, :2], X[[0, 1, 3, 5]]) and np.array_equal(it[:, 2:], [[0, 1], [1, 0], [0, 1], [1, 0]])
    assert np.array_equal(tr[:, :2], X[[0, 3]])
