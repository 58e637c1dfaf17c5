"""On-disk formats: what vboc_b200.io writes is read back with the EXACT loader lines of
triplependulum_comparison.py:28-57 (np.load / torch.load / load_state_dict on the reference's file names) and
gives the same network outputs."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from vboc_b200 import io as vio  # noqa: E402
from vboc_b200.shim.my_nn import NeuralNetCLS, NeuralNetDIR  # noqa: E402


def test_comparison_script_loader_lines(tmp_path, monkeypatch):
    rng = np.random.default_rng(0)
    n = 3
    X_save = np.hstack([rng.uniform(2.4, 3.9, (200, n)), rng.uniform(-10, 10, (200, n))])
    X_test = X_save[:50]
    X_iter = np.hstack([X_save, np.eye(2)[rng.integers(0, 2, 200)]])
    torch.manual_seed(0)
    net_dir, net_al = NeuralNetDIR(6, 500, 1), NeuralNetCLS(6, 500, 2)
    mean, std = vio.position_stats(X_save, n)
    assert isinstance(mean, float) and abs(mean - X_save[:, :n].astype(np.float32).mean()) < 1e-5
    assert abs(std - X_save[:, :n].astype(np.float32).std(ddof=1)) < 1e-5      # torch.std is unbiased
    vio.save_testdata(n, X_test, str(tmp_path))
    vio.save_run(n, "vboc", str(tmp_path / "VBOC"), data=X_save, model=net_dir, mean=mean, std=std,
                 times=[1.0, 2.5], rmse=[0.9, 0.4])
    vio.save_run(n, "al", str(tmp_path / "AL"), data=X_iter, model=net_al, mean=mean, std=std, times=[3.0], rmse=[0.7])
    monkeypatch.chdir(tmp_path)
    device = torch.device("cpu")
    # ---- triplependulum_comparison.py:28-41, 50-55, verbatim in structure
    X_test_l = np.load('data3_test.npy')
    model_dir = NeuralNetDIR(6, 500, 1).to(device)
    model_dir.load_state_dict(torch.load('VBOC/model_3dof_vboc'))
    data_reverse = np.load('VBOC/data_3dof_vboc.npy')
    mean_dir = torch.load('VBOC/mean_3dof_vboc')
    std_dir = torch.load('VBOC/std_3dof_vboc')
    model_al = NeuralNetCLS(6, 500, 2).to(device)
    model_al.load_state_dict(torch.load('AL/model_3dof_al'))
    mean_al = torch.load('AL/mean_3dof_al')
    std_al = torch.load('AL/std_3dof_al')
    data_al = np.load('AL/data_3dof_al.npy')
    times_al, rmse_al = np.load('AL/times_3dof_al.npy'), np.load('AL/rmse_3dof_al.npy')
    times_vboc, rmse_vboc = np.load('VBOC/times_3dof_vboc.npy'), np.load('VBOC/rmse_3dof_vboc.npy')
    # ----
    assert np.array_equal(X_test_l, X_test) and np.array_equal(data_reverse, X_save) and np.array_equal(data_al, X_iter)
    assert mean_dir == mean and std_dir == std and mean_al == mean and std_al == std
    assert times_vboc.tolist() == [1.0, 2.5] and rmse_vboc.tolist() == [0.9, 0.4]
    assert times_al.tolist() == [3.0] and rmse_al.tolist() == [0.7]
    x = torch.randn(8, 6)
    with torch.no_grad():
        assert torch.equal(model_dir(x), net_dir(x)) and torch.equal(model_al(x), net_al(x))
        assert (model_dir(x) >= 0).all()
    assert list(net_dir.state_dict()) == [f"linear_relu_stack.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias")]
    # the training rows the drivers build from X_save (VBOC/triplependulum_vboc.py:425-437)
    rows = vio.vboc_training_rows(X_save, n, mean, std)
    assert rows.shape == (200, 7)
    assert np.allclose(np.linalg.norm(rows[:, 3:6], axis=1), 1.0) and np.allclose(rows[:, 6], np.linalg.norm(X_save[:, 3:], axis=1))
    got = vio.load_run(n, "vboc", "VBOC", model=NeuralNetDIR(6, 500, 1))
    assert np.array_equal(got["data"], X_save) and got["rmse"].tolist() == [0.9, 0.4]


def test_reference_quirk_names(tmp_path):
    net = NeuralNetDIR(4, 300, 1)
    vio.save_run(2, "vboc", str(tmp_path), model=net, quirks=True)
    assert os.path.exists(tmp_path / "model_2dof_vboc") and os.path.exists(tmp_path / "model_2dof_vboc.npy")
    vio.save_run(3, "al", str(tmp_path), model=NeuralNetCLS(6, 500, 2), quirks=True)
    assert os.path.exists(tmp_path / "model_3dof_al") and os.path.exists(tmp_path / "model_3dof")
