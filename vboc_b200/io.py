"""On-disk formats of the reference drivers, so that the untouched `*_comparison.py` scripts read
what this engine produced.  File names and container types are the reference's:

  data{n}_test.npy                  X_test rows [q, v]            {pendulum,doublependulum,triplependulum}_testdata.py:72 / :141 / :145
  data_{n}dof_vboc.npy              X_save rows [q, v]            VBOC/triplependulum_vboc.py:583, VBOC/doublependulum_vboc.py:611, VBOC/vboc.py:458
  mean_{n}dof_vboc, std_{n}dof_vboc torch.save(python float)      VBOC/triplependulum_vboc.py:422-423, VBOC/vboc.py:476-477
  model_{n}dof_vboc                 torch.save(state_dict)        VBOC/triplependulum_vboc.py:584, VBOC/vboc.py:645
  times_/rmse_{n}dof_vboc.npy       1-D arrays, one entry per epoch checkpoint   VBOC/triplependulum_vboc.py:575-576
  data_{n}dof_al.npy                X_iter rows [x, one-hot]      AL/triplependulum_al.py:432
  mean_/std_{n}dof_al, model_{n}dof_al, times_/rmse_{n}dof_al.npy              AL/triplependulum_al.py:128-129, 429-431

Read back by triplependulum_comparison.py:28-57 (`np.load`, `torch.load`, `load_state_dict`).

Reference quirks that are NOT mirrored (the names the comparison scripts load are written instead):
  * VBOC/doublependulum_vboc.py:612 saves the 2-DOF state_dict as 'model_2dof_vboc.npy' while
    doublependulum_comparison.py loads 'VBOC/model_2dof_vboc';
  * AL/triplependulum_al.py:431 saves 'model_3dof' while triplependulum_comparison.py:39 loads 'AL/model_3dof_al'.
`quirks=True` writes those names as well.
"""
import os

import numpy as np


def _p(directory, name):
    os.makedirs(directory, exist_ok=True)
    return os.path.join(directory, name)


def save_testdata(n, X_test, directory="."):
    path = _p(directory, f"data{n}_test.npy")
    np.save(path, np.asarray(X_test, dtype=float))
    return path


def position_stats(X, n):
    """mean / std of all joint positions as the drivers compute them (torch.mean / torch.std -- the UNBIASED
    estimator -- of the float32 tensor of X[:, :n], VBOC/triplependulum_vboc.py:421): python floats."""
    import torch
    t = torch.tensor(np.asarray(X)[:, :n].tolist())
    return torch.mean(t).item(), torch.std(t).item()


def save_run(n, kind, directory=".", data=None, model=None, mean=None, std=None, times=None, rmse=None, quirks=False):
    """Write whatever of a finished VBOC / AL run is given (`kind` = 'vboc' or 'al') under the reference's names.
    Returns {what: path}."""
    import torch
    assert kind in ("vboc", "al")
    tag = f"{n}dof_{kind}"
    out = {}
    if data is not None:
        out["data"] = _p(directory, f"data_{tag}.npy")
        np.save(out["data"], np.asarray(data, dtype=float))
    if mean is not None:
        out["mean"] = _p(directory, f"mean_{tag}")
        torch.save(float(mean), out["mean"])
    if std is not None:
        out["std"] = _p(directory, f"std_{tag}")
        torch.save(float(std), out["std"])
    if model is not None:
        sd = model.state_dict() if hasattr(model, "state_dict") else model
        out["model"] = _p(directory, f"model_{tag}")
        torch.save(sd, out["model"])
        if quirks and kind == "vboc" and n == 2:
            torch.save(sd, _p(directory, "model_2dof_vboc.npy"))
        if quirks and kind == "al" and n == 3:
            torch.save(sd, _p(directory, "model_3dof"))
    if times is not None:
        out["times"] = _p(directory, f"times_{tag}.npy")
        np.save(out["times"], np.asarray(times, dtype=float))
    if rmse is not None:
        out["rmse"] = _p(directory, f"rmse_{tag}.npy")
        np.save(out["rmse"], np.asarray(rmse, dtype=float))
    return out


def vboc_training_rows(X_save, n, mean, std):
    """`X_train_dir` of the VBOC drivers (VBOC/triplependulum_vboc.py:425-437): [normalised positions, velocity
    direction, velocity norm]; rows with zero velocity keep a zero direction."""
    X = np.asarray(X_save, dtype=float)
    nrm = np.linalg.norm(X[:, n:2 * n], axis=1)
    out = np.zeros((X.shape[0], 2 * n + 1))
    out[:, :n] = (X[:, :n] - mean) / std
    nz = nrm != 0
    out[nz, n:2 * n] = X[nz, n:2 * n] / nrm[nz, None]
    out[:, 2 * n] = nrm
    return out


def load_run(n, kind, directory=".", model=None):
    """The loader lines of triplependulum_comparison.py:31-41 as a function (used by the tests; a user of the
    comparison scripts does not need it)."""
    import torch
    tag = f"{n}dof_{kind}"
    out = dict(data=np.load(os.path.join(directory, f"data_{tag}.npy")),
               mean=torch.load(os.path.join(directory, f"mean_{tag}")),
               std=torch.load(os.path.join(directory, f"std_{tag}")))
    if model is not None:
        model.load_state_dict(torch.load(os.path.join(directory, f"model_{tag}")))
    for k in ("times", "rmse"):
        p = os.path.join(directory, f"{k}_{tag}.npy")
        if os.path.exists(p):
            out[k] = np.load(p)
    return out
