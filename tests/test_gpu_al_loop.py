"""The active-learning loop end to end on the GPU engine: entropy + top-B on the tensor-core MLP kernel, labelling by
batched SQP_RTI solves started from the guess network's trajectories (2-DOF, two rounds)."""
import numpy as np
import pytest
import torch

from vboc_b200 import al_loop, problems as pr

pytestmark = pytest.mark.gpu


class Net(torch.nn.Module):
    def __init__(self, n_in, hidden, n_out):
        super().__init__()
        self.linear_relu_stack = torch.nn.Sequential(torch.nn.Linear(n_in, hidden), torch.nn.ReLU(),
                                                     torch.nn.Linear(hidden, hidden), torch.nn.ReLU(),
                                                     torch.nn.Linear(hidden, n_out))

    def forward(self, x):
        return self.linear_relu_stack(x)


def test_active_learning_two_rounds_on_gpu():
    torch.manual_seed(0)
    n, N, nx = 2, 100, 4
    pool = pr.sample_al(n, 1200, seed=9)["x0"]
    mean, std = torch.tensor(float(pool.mean())), torch.tensor(float(pool.std()))
    model, guess = Net(nx, 64, 2), Net(nx, 64, N * nx)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    optg = torch.optim.Adam(guess.parameters(), lr=1e-3)
    fit_cls = lambda m, Xi: al_loop.fit_minibatch(m, opt, torch.nn.BCEWithLogitsLoss(), Xi[:, :nx], Xi[:, nx:], mean, std,
                                                  n_minibatch=128, it_max=200)
    fit_guess = lambda m, Xt: al_loop.fit_minibatch(m, optg, torch.nn.MSELoss(), Xt[:, :nx], Xt[:, nx:], mean, std,
                                                    n_minibatch=128, it_max=100, normalize_targets=True)
    hist = []
    Xi, Xt, rest = al_loop.active_learning(n, pool, 400, 200, model, guess, mean, std, fit_cls, fit_guess,
                                           etp_stop=0.0, max_rounds=2, N=N, history=hist)
    assert len(hist) == 2 and len(rest) == 1200 - 400 - 400
    assert Xi.shape == (400, nx + 2) and Xt.shape[1] == (N + 1) * nx
    assert all(0 < h["viable"] < h["labelled"] for h in hist)  # the queried states straddle the boundary
    # viable rows carry a trajectory that starts at the state and ends at rest
    tr = Xt.reshape(len(Xt), N + 1, nx)
    assert np.abs(tr[:, -1, n:]).max() < 1e-6
    mdl = pr.Model(n)
    assert (tr[:, :, :n] >= mdl.thetamin - 1e-6).all() and (tr[:, :, :n] <= mdl.thetamax + 1e-6).all()
