"""A12: the fused MLP inference kernel against plain PyTorch FP32 (the reference's own code path:
my_nn.py + the normalisation / label / entropy lines of the drivers)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
from scipy.stats import entropy as sp_entropy

pytestmark = pytest.mark.gpu


def _net(n_in, hidden, n_out, final_relu, seed):
    torch.manual_seed(seed)
    layers = [nn.Linear(n_in, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, n_out)]
    if final_relu:
        layers.append(nn.ReLU())
    m = nn.Module()
    m.linear_relu_stack = nn.Sequential(*layers)
    m.forward = lambda x: m.linear_relu_stack(x)
    return m


@pytest.mark.parametrize("n,hidden", [(1, 100), (2, 300), (3, 500)])
def test_viability_filter_matches_torch(n, hidden):
    from vboc_b200 import nn as vnn
    model = _net(2 * n, hidden, 1, True, seed=n)
    rng = np.random.default_rng(n)
    B = 5000 + 7
    X = np.concatenate([rng.uniform(2.36, 3.93, (B, n)), rng.uniform(-10, 10, (B, n))], axis=1).astype(np.float32)
    mean, std = 3.14, 0.45
    # reference lines VBOC/triplependulum_vboc.py:604-620
    inp = X.copy()
    vel = np.linalg.norm(inp[:, n:], axis=1)
    inp[:, :n] = (inp[:, :n] - np.float32(mean)) / np.float32(std)
    inp[:, n:] = inp[:, n:] / vel[:, None]
    with torch.no_grad():
        ref = model.linear_relu_stack(torch.from_numpy(inp)).numpy()[:, 0]
    lab_ref = np.where(vel > ref, 0, 1)
    net = vnn.MLP.from_torch(model)
    phi, lab, margin = net.viability(X, mean, std, safety_margin=2.0)
    assert np.abs(phi - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
    assert (lab == lab_ref).mean() >= 0.999
    assert np.abs(margin - (ref * 0.98 - vel)).max() < 1e-3
    net.close()


def test_entropy_query_matches_torch_and_scipy():
    from vboc_b200 import nn as vnn
    n, hidden = 3, 500
    model = _net(2 * n, hidden, 2, False, seed=7)
    rng = np.random.default_rng(0)
    B = 4096 + 33
    X = np.concatenate([rng.uniform(2.36, 3.93, (B, n)), rng.uniform(-10.5, 10.5, (B, n))], axis=1).astype(np.float32)
    mean, std = float(X.mean()), float(X.std())
    with torch.no_grad():
        logits = model.linear_relu_stack((torch.from_numpy(X) - mean) / std)
        prob = torch.sigmoid(logits).numpy()
    etp_ref = sp_entropy(prob, axis=1)
    net = vnn.MLP.from_torch(model)
    out, etp = net.entropy(X, mean, std)
    assert np.abs(out - logits.numpy()).max() < 1e-4
    assert np.abs(etp - etp_ref).max() < 1e-4
    top = vnn.select_max_entropy(etp, 64)
    top_ref = np.argpartition(etp_ref, -64)[-64:]
    assert len(set(top) & set(top_ref.tolist())) >= 62
    net.close()
