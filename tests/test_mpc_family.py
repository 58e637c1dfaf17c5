"""SURVEY 8(f)4 on the CPU: the MPC family of the warp solver (tracking cost + the learned viability margin as a
nonlinear terminal constraint, `nn_margin.h`) compiled for the host by tools/emu, certified WITHOUT an oracle
(tools/certify.py, numpy): the step of one SQP_RTI iteration satisfies the dense KKT conditions of the QP linearised
at the guess -- including the constraint row, whose gradient the kernel forms by reverse mode through the network and
numpy by the complex step -- and a converged SQP run satisfies the NLP's KKT conditions."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
import certify  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    import emu as e
    e.build()
    return e


def make_net(n, H, seed, b3):
    rng = np.random.default_rng(seed)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    return dict(W1=f32(rng.normal(size=(H, 2 * n)) / np.sqrt(2 * n)), b1=f32(rng.normal(size=H) * 0.1),
                W2=f32(rng.normal(size=(H, H)) / np.sqrt(H)), b2=f32(rng.normal(size=H) * 0.1),
                W3=f32(rng.normal(size=H) / np.sqrt(H)), b3=b3, mean=np.pi, std=0.45, scale=1.0)


def mpc_opts(emu, qp_tol=1e-8, tol=1e-2):
    o = emu.Opts()
    o.tol_stat = o.tol_eq = o.tol_ineq = o.tol_comp = tol
    o.max_iter, o.levenberg_marquardt = 1000, 1.0
    o.alpha_min, o.alpha_reduction, o.globalization = 1e-2, 0.3, 1
    o.qp_tol_stat = o.qp_tol_eq = o.qp_tol_ineq = o.qp_tol_comp = qp_tol
    o.qp_iter_max, o.qp_mu0, o.qp_alpha_min, o.qp_reg_prim = 100, 1e1, 1e-12, 1e-15
    o.qp_lam_min = o.qp_t_min = o.qp_tau_min = 1e-16
    return o


@pytest.mark.parametrize("n,H", [(2, 64), (3, 96)])
def test_rti_step_satisfies_dense_kkt_with_the_margin_row(emu, n, H):
    net = make_net(n, H, n, 4.0)
    bp = pr.sample_mpc(n, 24, seed=3)
    out = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), multipliers=True)
    assert set(np.unique(out["status"]).tolist()) == {0, 4}
    active = 0
    for b in range(24):
        if out["status"][b] != 0:
            continue   # infeasible QP: a margin too negative to repair within 10 ms, or a state about to leave the box
        r = certify.mpc_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["lamg"][b], 1.0,
                            first_qp_at_guess=True)
        sc = max(1.0, float(out["lamg"][b].max()))   # the QP's stationarity test is relative to the row multiplier
        assert max(r["res_stat"] / sc, r["res_eq"], r["res_ineq"]) < 1e-7 and r["res_comp"] < 1e-7 and r["lam_min"] >= 0.0, (b, r)
        active += out["lamg"][b, 0] > 1e-3
    assert active >= 2                   # the terminal constraint binds on some problems (row and multiplier exercised)


def test_sqp_run_satisfies_nlp_kkt(emu):
    n = 2
    net = make_net(n, 64, 0, 4.0)
    bp = pr.sample_mpc(n, 12, seed=3)
    out = emu.solve_mpc(n, 0, bp, net, mpc_opts(emu, tol=1e-2), multipliers=True)
    ok = out["status"] == 0
    assert ok.sum() >= 6 and ((out["status"] == 4) | ok).all()
    for b in np.where(ok)[0]:
        r = certify.mpc_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["lamg"][b], 1.0)
        assert r["res_stat"] < 1e-2 and r["res_eq"] < 1e-2 and r["res_ineq"] < 1e-2 and r["res_comp"] < 1e-2
        # the engine's own residuals are the recomputed ones
        assert abs(r["res_stat"] - out["res"][b, 0]) < 1e-9 and abs(r["res_ineq"] - out["res"][b, 2]) < 1e-9
        assert r["h"] >= -1e-2            # the returned terminal state is inside the learned set (to tolerance)


def test_margin_function_matches_the_reference_expression():
    """certify.nn_margin restates nn_decisionfunction; the shim's numeric twin is an independent transcription."""
    sys.path.insert(0, os.path.join(ROOT, "vboc_b200", "shim", "SafeMPC"))
    from vboc_b200.shim.SafeMPC.doublependulum_class_fixedveldir import OCPdoublependulumINIT
    net = make_net(2, 32, 5, 3.0)
    params = [net[k] if k != "W3" else net[k][None, :] for k in ("W1", "b1", "W2", "b2", "W3")] + [np.array([net["b3"]])]
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = np.concatenate([rng.uniform(2.4, 3.9, 2), rng.uniform(-8, 8, 2)])
        want = OCPdoublependulumINIT.nn_decisionfunction(None, params, net["mean"], net["std"], 0.0, x)
        assert abs(certify.nn_margin(net, x, 2) - want) < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# soft rows: the margin at every stage with slacks (parallel / receiding_hard_constraints / soft_traj_constraints)
def _row_penalties(kind, B, N):
    Z = np.zeros((B, N + 1, 4))
    if kind == "soft_traj":          # VBOC/Safe MPC/soft_traj_constraints/2dof_sym.py:110-111
        Z[:, :, 0] = 1e6
    elif kind == "parallel":         # VBOC/Safe MPC/parallel/2dof_sym.py:44-50: 1e9 at one stage p, vacuous rows elsewhere
        for b in range(B):
            Z[b, N - b % 6, 0] = 1e9
    elif kind == "receding":         # VBOC/Safe MPC/receiding_hard_constraints/2dof_sym.py:53-57
        Z[:, :, 0] = 10 ** ((1 - 0.5) * 6)
        for b in range(B):
            Z[b, N - b % 6, 0] = 1e12
    elif kind == "generic":          # every penalty field in use (the engine's interface is general)
        Z[:, :, 0], Z[:, :, 1], Z[:, :, 2], Z[:, :, 3] = 1e2, 3.0, 0.5, 0.1
    return Z


def _rel(r, rowm):
    """KKT residuals relative to the size of the numbers they are differences of (multipliers reach Zl * sl ~ 1e9)."""
    s = max(1.0, float(np.abs(rowm[:, :4]).max()))
    return max(r["res_stat"], r["res_comp"]) / s, max(r["res_eq"], r["res_ineq"])


@pytest.mark.parametrize("kind", ["soft_traj", "parallel", "receding", "generic"])
def test_soft_rows_rti_step_satisfies_dense_kkt(emu, kind):
    n, B, N = 2, 24, 10
    net = make_net(n, 64, n, 4.0)
    bp = pr.sample_mpc(n, B, seed=3)
    Z = _row_penalties(kind, B, N)
    out = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), multipliers=True, rowZ=Z)
    ok = np.where(out["status"] == 0)[0]
    # with slacks the QP is feasible wherever the box constraints alone are: more problems solve than with the hard row
    hard = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), multipliers=True)
    assert len(ok) >= (hard["status"] == 0).sum() and len(ok) >= 20
    slack_used = 0
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b],
                                 1.0, first_qp_at_guess=True)
        rel, feas = _rel(r, out["rowm"][b])
        assert rel < 2e-6 and feas < 2e-6 and r["lam_min"] >= 0.0, (b, r)
        slack_used += r["sl"].max() > 1e-3
    assert slack_used >= 3               # the margin is violated (and paid for) on some problems


def test_soft_rows_sqp_run_satisfies_nlp_kkt(emu):
    n, B, N = 2, 12, 10
    net = make_net(n, 64, 0, 4.0)
    bp = pr.sample_mpc(n, B, seed=3)
    Z = _row_penalties("soft_traj", B, N)
    out = emu.solve_mpc(n, 0, bp, net, mpc_opts(emu, tol=1e-2), multipliers=True, rowZ=Z)
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) >= 8
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b], 1.0)
        assert max(r["res_stat"], r["res_eq"], r["res_ineq"], r["res_comp"]) < 1e-2, (b, r)
        s = max(1.0, float(np.abs(out["rowm"][b, :, :4]).max()))
        assert abs(r["res_stat"] - out["res"][b, 0]) < 1e-9 * s and abs(r["res_ineq"] - out["res"][b, 2]) < 1e-9


def test_soft_rows_limits(emu):
    """Zl = 0 everywhere: every row is vacuous, the solution is that of the problem without the margin constraint.
    Zl = 1e9 at the terminal stage only: the solution approaches that of the hard terminal row."""
    n, B, N = 2, 16, 10
    net = make_net(n, 64, n, 4.0)
    bp = pr.sample_mpc(n, B, seed=3)
    free = dict(bp)
    free["lh"] = -1e9                                    # hard row that can never bind
    ref = emu.solve_mpc(n, 1, free, net, mpc_opts(emu))
    out = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), rowZ=np.zeros((B, N + 1, 4)))
    ok = (ref["status"] == 0) & (out["status"] == 0)
    assert ok.sum() >= 14
    assert np.abs(out["x"] - ref["x"])[ok].max() < 1e-6 and np.abs(out["u"] - ref["u"])[ok].max() < 1e-5
    hard = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu))
    Z = np.zeros((B, N + 1, 4))
    Z[:, N, 0] = 1e9
    stiff = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), rowZ=Z)
    ok = (hard["status"] == 0) & (stiff["status"] == 0)
    assert ok.sum() >= 10
    assert np.abs(stiff["x"] - hard["x"])[ok].max() < 1e-5 and np.abs(stiff["u"] - hard["u"])[ok].max() < 1e-3


def test_triple_pendulum_velnorm_over_x2(emu):
    """The triple-pendulum Safe-MPC classes write vel_norm = norm_2(x[2:]) (VBOC/Safe MPC/triplependulum_class_vboc.py:217,
    282), i.e. including theta_3: with vstart = 2 the engine's row is that function's linearisation (dense KKT with the
    complex-step gradient of the numpy restatement), and the shim's numeric twin is the same expression."""
    n, B, N = 3, 16, 10
    net = make_net(n, 96, 1, 4.0)
    net["vstart"] = 2
    bp = pr.sample_mpc(n, B, seed=4)
    Z = _row_penalties("soft_traj", B, N)
    out = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), multipliers=True, rowZ=Z)
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) >= 12
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b],
                                 1.0, first_qp_at_guess=True)
        rel, feas = _rel(r, out["rowm"][b])
        assert rel < 2e-6 and feas < 2e-6, (b, r)
    wrong = dict(net, vstart=3)          # the certificate distinguishes the two expressions where the row is active
    b = ok[np.argmax(out["rowm"][ok, :, 0].max(axis=1))]
    assert out["rowm"][b, :, 0].max() > 1.0
    r = certify.mpc_rows_kkt(n, bp, wrong, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b], 1.0,
                             first_qp_at_guess=True)
    assert max(_rel(r, out["rowm"][b])) > 1e-4
    from vboc_b200.shim.SafeMPC.triplependulum_class_vboc import OCPtriplependulumSoftTraj
    params = [net[k] if k != "W3" else net[k][None, :] for k in ("W1", "b1", "W2", "b2", "W3")] + [np.array([net["b3"]])]
    ocp = OCPtriplependulumSoftTraj("SQP_RTI", 5e-3, 0.18, params, net["mean"], net["std"], 5.0, True)
    assert ocp.N == 36 and ocp.ocp.dims.N == 36 and ocp.SOFT_ROWS and callable(ocp.OCP_solve)
    rng = np.random.default_rng(0)
    for _ in range(10):
        x = np.concatenate([rng.uniform(2.4, 3.9, 3), rng.uniform(-8, 8, 3)])
        want = certify.nn_margin(dict(net, scale=0.95), x, 3)
        assert abs(ocp.nn_decisionfunction_conservative(params, net["mean"], net["std"], 5.0, x) - want) < 1e-12


def test_soft_rows_with_shorter_horizons_in_a_larger_buffer(emu):
    """N < N_max: the penalties of problem b, stage k are Z[b, k] with the N_max + 1 row stride of the C-ABI
    (vboc_set_mpc_rows), rows beyond N are ignored; the certificate runs on the first N + 1 stages."""
    n, B, Nmax = 2, 12, 10
    net = make_net(n, 64, n, 4.0)
    bp = pr.sample_mpc(n, B, seed=5, N=Nmax)
    bp["N"] = np.array([6, 10, 8] * 4, dtype=np.int32)
    Z = np.zeros((B, Nmax + 1, 4))
    Z[:, :, 0] = 1e6
    for b in range(B):
        Z[b, bp["N"][b] + 1:, 0] = np.nan          # must never be read
    out = emu.solve_mpc(n, 1, bp, net, mpc_opts(emu), multipliers=True, rowZ=np.nan_to_num(Z, nan=-7.0))
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) >= 10
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b],
                                 np.nan_to_num(Z[b], nan=0.0), 1.0, first_qp_at_guess=True)
        rel, feas = _rel(r, out["rowm"][b, :bp["N"][b] + 1])
        assert rel < 2e-6 and feas < 2e-6, (b, r)
