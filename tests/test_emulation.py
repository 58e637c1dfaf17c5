"""The warp solver's lane program (vboc_b200/csrc/ocp_warp.h) compiled for the host by tools/emu
(32 lanes of every lane region run as a loop) against the oracle: same statuses and iteration counts,
trajectories to 1e-6.  This is the CPU-side check of the kernel SOURCE; the GPU parity tests check the
compiled kernel."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))


@pytest.fixture(scope="module")
def emu():
    import emu as e
    e.build()
    return e


def _opts(emu, oo):
    o = emu.Opts()
    for f, _ in emu.Opts._fields_:
        setattr(o, f, getattr(oo, f))
    return o


@pytest.mark.parametrize("n,fam,mode", [(3, 0, 0), (2, 0, 0), (1, 0, 0), (3, 1, 1), (2, 1, 1), (1, 1, 1)])
def test_lane_program_matches_oracle(oracle, emu, n, fam, mode):
    if fam == 0:
        bp = pr.sample_vboc(n, 6, seed=1) if n > 1 else pr.sample_testdata(1, 6, seed=1)
    else:
        bp = pr.sample_al(n, 24, seed=2)
    oo = oracle.default_opts(fam)
    ref = oracle.solve_batch(n, fam, mode, bp, oo)
    out = emu.solve_batch(n, fam, mode, bp, _opts(emu, oo))
    assert (ref["status"] == out["status"]).all()
    assert (ref["sqp_iter"] == out["sqp_iter"]).all() and (ref["qp_iter"] == out["qp_iter"]).all()
    ok = out["status"] == 0
    assert ok.any()
    assert np.abs(ref["x"] - out["x"])[ok].max() < 1e-6


def test_free_dt_lane_program_matches_oracle(oracle, emu):
    """1-DOF VBOC with the dt state kept (VBOC/pendulum_vboc.py): lane-per-OCP program with bordering of the
    two terminal equalities against the oracle's square-root Riccati + bordering."""
    bp = pr.pendulum_free_dt_problems(10, seed=3)
    oo = oracle.default_opts(0)
    ref = oracle.solve_batch(1, 0, 0, bp, oo)
    out = emu.solve_batch(1, 0, 0, bp, _opts(emu, oo), "lane_dts")
    assert (ref["status"] == out["status"]).all() and (out["status"] == 0).sum() >= 8
    assert (ref["sqp_iter"] == out["sqp_iter"]).all()
    ok = out["status"] == 0
    assert np.abs(ref["x"] - out["x"])[ok].max() < 1e-6
    # the optimal dt is interior (time costs) and constant along the horizon
    x = out["x"][ok]
    assert (x[:, 0, 2] > 1e-4).all() and (x[:, 0, 2] < 1e-2).all()
    assert np.abs(np.diff(x[:, :, 2], axis=1)).max() < 1e-9


def test_guess_network_inside_the_solver(oracle, emu):
    """`compute_problem_nnguess` with the guess network evaluated by the solver itself (vboc_set_guess_network): the
    computed guess equals the float64 forward of the network, and the RTI result equals the oracle's from that guess."""
    n, N, H = 2, 100, 48
    nx = 2 * n
    rng = np.random.default_rng(0)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    net = dict(W1=f32(rng.normal(size=(H, nx)) / 2), b1=f32(rng.normal(size=H) * 0.1), W2=f32(rng.normal(size=(H, H)) / 7),
               b2=f32(rng.normal(size=H) * 0.1), W3=f32(rng.normal(size=(N * nx, H)) * 0.02), b3=f32(rng.normal(size=N * nx) * 0.02),
               mean=3.0, std=2.5)
    bp = pr.sample_al(n, 24, seed=6)
    out = emu.solve_batch(n, 1, 1, bp, _opts(emu, oracle.default_opts(1)), guess_net=net)
    x0 = bp["lbx0"][:, :nx]
    a = np.maximum(((x0 - net["mean"]) / net["std"]) @ net["W1"].T + net["b1"], 0)
    a = np.maximum(a @ net["W2"].T + net["b2"], 0)
    want = (a @ net["W3"].T + net["b3"]) * net["std"] + net["mean"]
    assert np.abs(out["x_guess"][:, 1:] - want.reshape(24, N, nx)).max() < 1e-12
    assert np.array_equal(out["x_guess"][:, 0], x0)
    ref = oracle.solve_batch(n, 1, 1, pr.al_problems(n, x0, x_guess=out["x_guess"]))
    assert (ref["status"] == out["status"]).all() and (ref["qp_iter"] == out["qp_iter"]).all()
    ok = out["status"] == 0
    assert ok.any() and np.abs(ref["x"] - out["x"])[ok].max() < 1e-6
