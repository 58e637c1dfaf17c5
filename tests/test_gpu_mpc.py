"""SURVEY 8(f)4 on the GPU: the MPC family through the C-ABI (`vboc_create(n, VBOC_FAMILY_MPC, ..)`, `vboc_set_mpc`,
`vboc_set_mpc_reference`, `vboc_solve_batch`) -- the Safe-MPC OCP with the learned viability margin as a nonlinear
terminal constraint -- certified with numpy only (tools/certify.py), and the drop-in class of
`VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py` called as its driver calls it."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import certify  # noqa: E402
from test_mpc_family import make_net  # noqa: E402

pytestmark = pytest.mark.gpu


# QP tolerance 1e-8 (HPIPM's default): with a strongly active row (multiplier ~1e4, slack ~1e-14) the Newton systems are
# conditioned ~1e17 and 1e-9 on the stationarity residual is at the rounding floor of FP64 -- an IPM asked for it wanders
# after reaching ~3e-9 (the iteration limit is then hit and, as in acados, tolerated).
def _solve(n, bp, net, mode, tol=1e-2, qp_tol=1e-8):
    from vboc_b200 import engine
    B = len(bp["N"])
    sol = engine.BatchSolver(n, "mpc", B, int(bp["x_guess"].shape[1] - 1))
    o = engine.default_opts("mpc")
    o.tol_stat = o.tol_eq = o.tol_ineq = o.tol_comp = tol
    o.qp_tol_stat = o.qp_tol_eq = o.qp_tol_ineq = o.qp_tol_comp = qp_tol
    sol.set_opts(o)
    w = dict(net)
    w["W3"], w["b3"] = net["W3"][None, :], np.array([net["b3"]])
    sol.set_mpc(w, net["mean"], net["std"], 100.0 * (1.0 - net["scale"]), bp["W"], bp["W_e"], lh=bp["lh"], uh=bp["uh"])
    sol.set_mpc_reference(bp["yref_acados"], bp["yrefN"])
    sol.export_multipliers(True)
    out = sol.solve(bp, mode)
    out["pi"], out["lam"] = sol.multipliers()
    out["lamg"] = sol.mpc_multipliers()
    sol.close()
    return out


@pytest.mark.parametrize("n,H", [(2, 300), (3, 500)])
def test_rti_step_is_the_qp_solution(n, H):
    """The reference's network sizes (4-300-300-1, 6-500-500-1); 256 problems; every solved QP satisfies its dense KKT
    conditions (constraint row included), safety margin 2 %."""
    net = make_net(n, H, n, 4.0)
    net["scale"] = 0.98
    bp = pr.sample_mpc(n, 256, seed=3)
    out = _solve(n, bp, net, 1)
    assert set(np.unique(out["status"]).tolist()) == {0, 4}
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) > 100
    active = 0
    for b in ok:
        r = certify.mpc_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["lamg"][b], 1.0,
                            first_qp_at_guess=True)
        sc = max(1.0, float(out["lamg"][b].max()))   # the QP's stationarity test is relative to the row multiplier
        assert max(r["res_stat"] / sc, r["res_eq"], r["res_ineq"], r["res_comp"]) < 1e-7 and r["lam_min"] >= 0.0, (b, r)
        active += out["lamg"][b, 0] > 1e-3
    assert active >= 10


def test_sqp_run_is_a_kkt_point():
    n = 2
    net = make_net(n, 300, 7, 4.0)
    bp = pr.sample_mpc(n, 128, seed=11)
    out = _solve(n, bp, net, 0, tol=1e-2)
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) > 60 and np.isin(out["status"], (0, 2, 4)).all()
    for b in ok:
        r = certify.mpc_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["lamg"][b], 1.0)
        assert max(r["res_stat"], r["res_eq"], r["res_ineq"], r["res_comp"]) < 1e-2, (b, r)
        assert abs(r["res_stat"] - out["res_stat"][b]) < 1e-8 and abs(r["res_ineq"] - out["res_ineq"][b]) < 1e-8


def test_safe_mpc_class_like_the_driver():
    """VBOC/Safe MPC/hard_terminal_constraints/2dof_sym.py: OCP_solve(x0, q_ref, guesses) -> status, get(0, 'u'), shifted
    warm start, a-posteriori check of the margin at the terminal state."""
    import torch
    from vboc_b200.shim.my_nn import NeuralNetDIR
    from vboc_b200.shim.SafeMPC.doublependulum_class_fixedveldir import OCPdoublependulumINIT, SYMdoublependulumINIT
    torch.manual_seed(0)
    model = NeuralNetDIR(4, 300, 1)
    with torch.no_grad():
        model.linear_relu_stack[4].bias.fill_(6.0)       # predicted max velocity norm ~ 6 everywhere
    params = list(model.parameters())
    mean, std = 3.14, 0.45
    ocp = OCPdoublependulumINIT(True, params, mean, std, 2.0)
    sim = SYMdoublependulumINIT(True)
    x = np.array([3.0, 3.3, 1.0, -0.5])
    q_ref = np.array([np.pi, np.pi])
    xg = np.full((ocp.N + 1, 4), x)
    ug = np.full((ocp.N, 2), [ocp.g * ocp.l1 * (ocp.m1 + ocp.m2) * np.sin(x[0]), ocp.g * ocp.l2 * ocp.m2 * np.sin(x[1])])
    for step in range(5):
        status = ocp.OCP_solve(x, q_ref, xg, ug)
        assert status == 0
        xN = ocp.ocp_solver.get(ocp.N, "x")
        assert ocp.nn_decisionfunction(params, mean, std, 2.0, xN) >= -1e-6
        u0 = ocp.ocp_solver.get(0, "u")
        assert np.abs(u0).max() <= ocp.Cmax + 1e-9
        for i in range(ocp.N - 1):
            xg[i], ug[i] = ocp.ocp_solver.get(i + 1, "x"), ocp.ocp_solver.get(i + 1, "u")
        xg[ocp.N - 1] = xg[ocp.N] = ocp.ocp_solver.get(ocp.N, "x")
        sim.acados_integrator.set("u", u0)
        sim.acados_integrator.set("x", x)
        sim.acados_integrator.solve()
        x = sim.acados_integrator.get("x")
    # a state far outside the learned set: the terminal constraint cannot be met within the horizon -> QP failure
    x_bad = np.array([3.0, 3.3, 9.0, 9.0])
    assert ocp.OCP_solve(x_bad, q_ref, np.full((ocp.N + 1, 4), x_bad), ug) == 4


# ---------------------------------------------------------------------------------------------------------------
# soft rows (vboc_set_mpc_rows): the parallel / receiding_hard_constraints / soft_traj_constraints variants
def _solve_rows(n, bp, net, mode, Z, tol=1e-2, qp_tol=1e-8):
    from vboc_b200 import engine
    B = len(bp["N"])
    sol = engine.BatchSolver(n, "mpc", B, int(bp["x_guess"].shape[1] - 1))
    o = engine.default_opts("mpc")
    o.tol_stat = o.tol_eq = o.tol_ineq = o.tol_comp = tol
    o.qp_tol_stat = o.qp_tol_eq = o.qp_tol_ineq = o.qp_tol_comp = qp_tol
    sol.set_opts(o)
    w = dict(net)
    w["W3"], w["b3"] = net["W3"][None, :], np.array([net["b3"]])
    sol.set_mpc(w, net["mean"], net["std"], 100.0 * (1.0 - net["scale"]), bp["W"], bp["W_e"], lh=bp["lh"], uh=bp["uh"])
    sol.set_mpc_reference(bp["yref_acados"], bp["yrefN"])
    sol.set_mpc_rows(Z)
    sol.export_multipliers(True)
    out = sol.solve(bp, mode)
    out["pi"], out["lam"] = sol.multipliers()
    out["rowm"] = sol.mpc_rows()
    sol.close()
    return out


@pytest.mark.parametrize("n,H,kind", [(2, 300, "soft_traj"), (2, 300, "parallel"), (2, 300, "receding"), (3, 500, "soft_traj"),
                                      (3, 500, "generic")])
def test_soft_rows_rti_step_is_the_qp_solution(n, H, kind):
    """The margin row at every stage with slacks: every solved QP of 256 problems satisfies the dense KKT conditions of the
    QP WITH its slack variables (numpy, tools/certify.py::mpc_rows_kkt), relative to the size of the multipliers."""
    from test_mpc_family import _rel, _row_penalties
    net = make_net(n, H, n, 4.0)
    net["scale"] = 0.98
    B, N = 256, 10
    bp = pr.sample_mpc(n, B, seed=3)
    Z = _row_penalties(kind, B, N)
    out = _solve_rows(n, bp, net, 1, Z)
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) > 200 and np.isin(out["status"], (0, 4)).all()
    slack_used = 0
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b],
                                 1.0, first_qp_at_guess=True)
        rel, feas = _rel(r, out["rowm"][b])
        assert rel < 2e-6 and feas < 2e-6 and r["lam_min"] >= 0.0, (b, r)
        slack_used += r["sl"].max() > 1e-3
    assert slack_used >= 10


def test_soft_rows_sqp_run_is_a_kkt_point():
    from test_mpc_family import _row_penalties
    n, B, N = 2, 128, 10
    net = make_net(n, 300, 7, 4.0)
    bp = pr.sample_mpc(n, B, seed=11)
    Z = _row_penalties("soft_traj", B, N)
    out = _solve_rows(n, bp, net, 0, Z, tol=1e-2)
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) > 80 and np.isin(out["status"], (0, 2, 4)).all()
    for b in ok:
        r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b], Z[b], 1.0)
        assert max(r["res_stat"], r["res_eq"], r["res_ineq"], r["res_comp"]) < 1e-2, (b, r)
        s = max(1.0, float(np.abs(out["rowm"][b, :, :4]).max()))
        assert abs(r["res_stat"] - out["res_stat"][b]) < 1e-8 * s and abs(r["res_ineq"] - out["res_ineq"][b]) < 1e-8


def test_parallel_safe_mpc_class_like_the_driver():
    """VBOC/Safe MPC/parallel/2dof_sym.py:31-75: per MPC step the slack penalty Zl = 1e9 is put on ONE stage p, tried from
    p = N downwards until a solve succeeds; the margin is then checked a posteriori at that stage."""
    import torch
    from vboc_b200.shim.my_nn import NeuralNetDIR
    from vboc_b200.shim.SafeMPC.parallel.doublependulum_class_fixedveldir import OCPdoublependulumINIT, SYMdoublependulumINIT
    torch.manual_seed(0)
    model = NeuralNetDIR(4, 300, 1)
    with torch.no_grad():
        model.linear_relu_stack[4].bias.fill_(6.0)
    params = list(model.parameters())
    mean, std, safety = 3.14, 0.45, 2.0
    ocp = OCPdoublependulumINIT(True, params, mean, std, safety)
    sim = SYMdoublependulumINIT(True)
    x = np.array([3.0, 3.3, 1.0, -0.5])
    q_ref = np.array([np.pi, np.pi])
    xg = np.full((ocp.N + 1, 4), x)
    ug = np.full((ocp.N, 2), [ocp.g * ocp.l1 * (ocp.m1 + ocp.m2) * np.sin(x[0]), ocp.g * ocp.l2 * ocp.m2 * np.sin(x[1])])
    for step in range(4):
        for p in reversed(range(ocp.N - 2, ocp.N + 1)):
            for i in range(ocp.N + 1):
                ocp.ocp_solver.cost_set(i, "Zl", (1e9 if i == p else 0.0) * np.ones((1,)))
            status = ocp.OCP_solve(x, q_ref, xg, ug)
            if status == 0:
                break
        assert status == 0
        xp = ocp.ocp_solver.get(p, "x")
        assert ocp.nn_decisionfunction(params, mean, std, safety, xp) >= -1e-5      # the penalised stage is inside the set
        assert ocp.ocp_solver.get(p, "sl")[0] < 1e-5
        u0 = ocp.ocp_solver.get(0, "u")
        assert np.abs(u0).max() <= ocp.Cmax + 1e-9
        for i in range(ocp.N - 1):
            xg[i], ug[i] = ocp.ocp_solver.get(i + 1, "x"), ocp.ocp_solver.get(i + 1, "u")
        xg[ocp.N - 1] = xg[ocp.N] = ocp.ocp_solver.get(ocp.N, "x")
        sim.acados_integrator.set("u", u0)
        sim.acados_integrator.set("x", x)
        sim.acados_integrator.solve()
        x = sim.acados_integrator.get("x")
    # far outside the learned set the softened problem still solves (the hard-row class returns 4 here) and pays a slack
    x_bad = np.array([3.0, 3.3, 9.0, 9.0])
    for i in range(ocp.N + 1):
        ocp.ocp_solver.cost_set(i, "Zl", 1e6 * np.ones((1,)))
    assert ocp.OCP_solve(x_bad, q_ref, np.full((ocp.N + 1, 4), x_bad), ug) == 0
    assert max(ocp.ocp_solver.get(i, "sl")[0] for i in range(ocp.N + 1)) > 1e-2


def test_triple_pendulum_safe_mpc_classes_like_the_driver():
    """VBOC/Safe MPC/soft_traj_constraints/3dof_sym.py:17-70, 102-116 and hard_terminal_constraints/3dof_sym.py: the
    triple-pendulum classes (time_step 5e-3, tot_time 0.18 -> N = 36), Zl = 0 on the running stages and 1e6 on the
    terminal one, OCP_solve(x0, guesses) in closed loop with the simulator; the tracking-only and hard-terminal classes
    on the same state."""
    import torch
    from vboc_b200.shim.my_nn import NeuralNetDIR
    from vboc_b200.shim.SafeMPC.triplependulum_class_vboc import (OCPtriplependulumHardTerm, OCPtriplependulumSoftTraj,
                                                                   OCPtriplependulumSTD, SYMtriplependulum)
    torch.manual_seed(0)
    model = NeuralNetDIR(6, 500, 1)
    with torch.no_grad():
        model.linear_relu_stack[4].bias.fill_(8.0)
    params = list(model.parameters())
    mean, std, safety = 3.14, 0.45, 5.0
    time_step, tot_time = 5e-3, 0.18
    ocp = OCPtriplependulumSoftTraj("SQP_RTI", time_step, tot_time, params, mean, std, safety, True)
    sim = SYMtriplependulum(time_step, tot_time, True)
    N = ocp.ocp.dims.N
    assert N == 36
    for i in range(N):
        ocp.ocp_solver.cost_set(i, "Zl", 0 * np.ones((1,)))
    ocp.ocp_solver.cost_set(N, "Zl", 1e6 * np.ones((1,)))
    x = np.array([3.0, 3.3, 3.2, 0.5, -0.5, 0.2])
    xg = np.full((N + 1, 6), x)
    ug = np.zeros((N, 3))
    for step in range(4):
        assert ocp.OCP_solve(x, xg, ug) == 0
        xN = ocp.ocp_solver.get(N, "x")
        # the terminal row: inside the set, or the violation is the (penalised) slack
        h = ocp.nn_decisionfunction_conservative(params, mean, std, safety, xN)
        assert h + ocp.ocp_solver.get(N, "sl")[0] >= -1e-4
        u0 = ocp.ocp_solver.get(0, "u")
        assert np.abs(u0).max() <= ocp.Cmax + 1e-9
        for i in range(N - 1):
            xg[i], ug[i] = ocp.ocp_solver.get(i + 1, "x"), ocp.ocp_solver.get(i + 1, "u")
        xg[N - 1] = xg[N] = ocp.ocp_solver.get(N, "x")
        sim.acados_integrator.set("u", u0)
        sim.acados_integrator.set("x", x)
        sim.acados_integrator.solve()
        x = sim.acados_integrator.get("x")
    x0 = np.array([3.0, 3.3, 3.2, 0.5, -0.5, 0.2])
    std_ocp = OCPtriplependulumSTD("SQP_RTI", time_step, tot_time, True)
    hard = OCPtriplependulumHardTerm("SQP_RTI", time_step, tot_time, params, mean, std, True)
    g = np.full((N + 1, 6), x0)
    assert std_ocp.OCP_solve(x0, g, np.zeros((N, 3))) == 0 and hard.OCP_solve(x0, g, np.zeros((N, 3))) == 0
    assert hard.nn_decisionfunction(params, mean, std, hard.ocp_solver.get(N, "x")) >= -1e-6
    # the margin is far from binding here (bias 8 against |x[2:]| ~ 3.3): both classes return the tracking solution
    assert np.abs(hard.ocp_solver.get(0, "u") - std_ocp.ocp_solver.get(0, "u")).max() < 1e-6
