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


@pytest.mark.parametrize("n,hidden,n_out", [(1, 100, 2), (2, 300, 2), (3, 500, 2), (3, 500, 1)])
def test_pipelined_kernel_equals_serial_and_cuda_core_kernels(monkeypatch, n, hidden, n_out):
    """Three implementations of the same op: the pipelined tcgen05 kernel (product), the serial tcgen05 kernel of round
    1 and the FP32 CUDA-core kernel.  Batches that are not a multiple of the tile and span several persistent-CTA
    rounds (more tiles than SMs)."""
    from vboc_b200 import nn as vnn
    model = _net(2 * n, hidden, n_out, n_out == 1, seed=10 + n)
    rng = np.random.default_rng(n)
    B = 128 * 400 + 77
    X = np.concatenate([rng.uniform(2.36, 3.93, (B, n)), rng.uniform(-10.5, 10.5, (B, n))], axis=1).astype(np.float32)
    mean, std = 3.1, 4.2
    outs = {}
    for name, env in (("pipe", {}), ("serial", {"VBOC_MLP_SERIAL": "1"}), ("cuda", {"VBOC_MLP_CUDA_CORES": "1"})):
        monkeypatch.delenv("VBOC_MLP_SERIAL", raising=False)
        monkeypatch.delenv("VBOC_MLP_CUDA_CORES", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        net = vnn.MLP.from_torch(model)
        outs[name] = net.entropy(X, mean, std)
        net.close()
    with torch.no_grad():
        ref = model.linear_relu_stack((torch.from_numpy(X) - mean) / std).numpy()
    scale = max(1.0, np.abs(ref).max())
    for name in outs:
        assert np.abs(outs[name][0] - ref).max() < 1e-4 * scale, name
    assert np.abs(outs["pipe"][0] - outs["serial"][0]).max() < 2e-5 * scale
    assert np.abs(outs["pipe"][1] - outs["serial"][1]).max() < 2e-5


def test_resident_pool_query_equals_numpy():
    """vboc_pool_*: scores of the resident pool = the host-path entropies; device top-k = the numpy selection (as a
    set, above the k-th score; ties arbitrary); device removal = np.delete; two rounds."""
    from vboc_b200 import nn as vnn
    n, hidden = 3, 500
    model = _net(2 * n, hidden, 2, False, seed=3)
    rng = np.random.default_rng(5)
    P, k = 300_000 + 123, 4096
    X = np.concatenate([rng.uniform(2.36, 3.93, (P, n)), rng.uniform(-10.5, 10.5, (P, n))], axis=1).astype(np.float32)
    mean, std = float(X.mean()), float(X.std())
    net = vnn.MLP.from_torch(model)
    rp = vnn.ResidentPool(X)
    host = X.copy()
    for rnd in range(2):
        rp.score(net, mean, std)
        _, etp = net.entropy(host, mean, std)
        assert np.array_equal(rp.scores(), etp)              # same kernel, same rows -> bit-identical
        idx, rows, sc = rp.select(k)
        assert len(idx) == k and (np.diff(idx) < 0).all()    # unique, largest index first
        kth = np.partition(etp, -k)[-k]
        assert (etp[idx] >= kth).all() and set(np.where(etp > kth)[0].tolist()) <= set(idx.tolist())
        assert np.array_equal(rows, host[idx]) and np.array_equal(sc, etp[idx])
        rp.remove_selected()
        host = np.delete(host, idx, axis=0)
        assert len(rp) == len(host) and np.array_equal(rp.rows(), host)
    rp.close()
    net.close()


def test_al_loop_with_resident_pool_matches_host_pool():
    """The AL loop with the pool resident on the device returns what the host-pool loop returns (2-DOF, 2 rounds)."""
    from vboc_b200 import al_loop, problems as pr
    from vboc_b200.shim.my_nn import NeuralNetCLS
    n, N, nx = 2, 100, 4
    pool = pr.sample_al(n, 1500, seed=9)["x0"]
    mean, std = torch.tensor(float(pool.mean())), torch.tensor(float(pool.std()))
    res = []
    for resident in (False, True):
        torch.manual_seed(0)
        model, guess = NeuralNetCLS(nx, 100, 2), NeuralNetCLS(nx, 100, N * nx)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        optg = torch.optim.Adam(guess.parameters(), lr=1e-3)
        fit_cls = lambda m, Xi: al_loop.fit_minibatch(m, opt, torch.nn.BCEWithLogitsLoss(), Xi[:, :nx], Xi[:, nx:], mean, std,
                                                      n_minibatch=128, it_max=100)
        fit_guess = lambda m, Xt: al_loop.fit_minibatch(m, optg, torch.nn.MSELoss(), Xt[:, :nx], Xt[:, nx:], mean, std,
                                                        n_minibatch=128, it_max=50, normalize_targets=True)
        hist = []
        res.append(al_loop.active_learning(n, pool, 400, 200, model, guess, mean, std, fit_cls, fit_guess, etp_stop=0.0,
                                           max_rounds=2, N=N, history=hist, resident=resident) + (hist,))
    (Xi0, Xt0, rest0, h0), (Xi1, Xt1, rest1, h1) = res
    assert [h["labelled"] for h in h0] == [h["labelled"] for h in h1] == [200, 200]
    assert Xi0.shape == Xi1.shape and rest0.shape == rest1.shape == (1500 - 400 - 400, nx)
    # same seeds, same scores: the same states are queried (the float32 pool copy only feeds the network)
    assert np.array_equal(np.sort(rest0, axis=0), np.sort(rest1, axis=0))
    assert np.array_equal(Xi0, Xi1)
