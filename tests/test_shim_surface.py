"""The shim modules expose the reference's module / class / attribute names without touching the GPU
(`*_comparison.py` construct an OCP object only to read the bounds, triplependulum_comparison.py:14-21)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(sub, mod):
    path = os.path.join(ROOT, "vboc_b200", "shim", sub)
    sys.path.insert(0, path)
    try:
        sys.modules.pop(mod, None)
        return importlib.import_module(mod)
    finally:
        sys.path.remove(path)


@pytest.mark.parametrize("mod,cls,sym,n", [("doublependulum_class_vboc", "OCPdoublependulumINIT", "SYMdoublependulumINIT", 2),
                                           ("triplependulum_class_vboc", "OCPtriplependulumINIT", "SYMtriplependulumINIT", 3)])
def test_vboc_classes(mod, cls, sym, n):
    m = _load("VBOC", mod)
    ocp, sim = getattr(m, cls)(), getattr(m, sym)()
    assert ocp.N == 100 and ocp.Cmax == 10. and ocp.dthetamax == 10.
    assert np.isclose(ocp.thetamax, np.pi / 4 + np.pi) and np.isclose(ocp.thetamin, -np.pi / 4 + np.pi)
    assert ocp.ocp.dims.nx == 2 * n + 1 and ocp.ocp.dims.nu == n
    assert ocp.ocp.solver_options.nlp_solver_tol_stat == 1e-3
    for name in ("reset", "set", "constraints_set", "solve", "get", "get_cost", "set_new_time_steps",
                 "update_qp_solver_cond_N"):
        assert callable(getattr(ocp.ocp_solver, name))
    for name in ("set", "solve", "get"):
        assert callable(getattr(sim.acados_integrator, name))
    assert callable(ocp.OCP_solve)
    ocp.ocp_solver.set_new_time_steps(np.full((101,), 1.))
    assert ocp.ocp_solver.N == 101


@pytest.mark.parametrize("mod,cls,n", [("pendulum_class_al", "OCPpendulumINIT", 1),
                                       ("doublependulum_class_al", "OCPdoublependulumINIT", 2),
                                       ("triplependulum_class_al", "OCPtriplependulumINIT", 3)])
def test_al_classes(mod, cls, n):
    m = _load("AL", mod)
    ocp = getattr(m, cls)()
    assert (ocp.N, ocp.nx, ocp.nu) == (100, 2 * n, n)
    assert callable(ocp.compute_problem) and callable(ocp.compute_problem_nnguess) and callable(ocp.set_bounds)
    s = ocp.ocp_solver
    assert (s.lbx[100, n:] == 0).all() and (s.ubx[100, n:] == 0).all()  # terminal zero velocity


def test_pendulum_vboc_class_surface():
    m = _load("VBOC", "pendulum_class_vboc")
    ocp = m.OCPpendulum()
    assert ocp.N == 50 and ocp.Fmax == 3 and ocp.ocp.dims.nx == 3
    assert callable(ocp.OCP_solve) and callable(ocp.ocp_solver.solve)


def test_safe_mpc_class_surface():
    """VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:110-276 and its driver 2dof_sym.py:
    constructor arguments, dimensions, cost / constraint data the driver reads, callable surface."""
    m = _load("SafeMPC", "doublependulum_class_fixedveldir")
    rng = np.random.default_rng(0)
    params = [rng.normal(size=s).astype(np.float32) for s in ((30, 4), (30,), (30, 30), (30,), (1, 30), (1,))]
    ocp = m.OCPdoublependulumINIT(True, params, 3.14, 0.45, 2.0)
    sim = m.SYMdoublependulumINIT(True)
    assert ocp.N == 10 and ocp.Tf == 0.01 and (ocp.nx, ocp.nu) == (4, 2) and ocp.Cmax == 10.
    assert ocp.ocp.dims.nx == 4 and ocp.ocp.dims.nu == 2 and ocp.ocp.solver_options.nlp_solver_type == "SQP_RTI"
    assert np.allclose(ocp.ocp.cost.yref, [np.pi, np.pi, 0, 0, 0, 0])
    for name in ("reset", "set", "cost_set", "constraints_set", "solve", "get"):
        assert callable(getattr(ocp.ocp_solver, name))
    assert callable(ocp.OCP_solve) and callable(sim.acados_integrator.solve) and sim.acados_integrator.T == 1e-3
    # the numeric twin of the constraint function: at rest the margin is the network's output minus the 1e-3 floor
    x = np.array([3.0, 3.2, 0.0, 0.0])
    h = ocp.nn_decisionfunction(params, 3.14, 0.45, 2.0, x)
    a = np.maximum(params[0].astype(float) @ np.array([(3.0 - 3.14) / 0.45, (3.2 - 3.14) / 0.45, 0, 0]) + params[1], 0)
    a = np.maximum(params[2].astype(float) @ a + params[3], 0)
    assert abs(h - (float((params[4].astype(float) @ a + params[5])[0]) - 1e-3)) < 1e-12


@pytest.mark.parametrize("variant", ["parallel", "receiding_hard_constraints", "soft_traj_constraints"])
def test_soft_row_safe_mpc_class_surface(variant):
    """VBOC/Safe MPC/{parallel, receiding_hard_constraints, soft_traj_constraints}/doublependulum_class_fixedveldir.py:
    the same surface plus cost_set(i, "Zl", ..) (2dof_sym.py of each variant) and the margin WITH the safety factor."""
    m = importlib.import_module(f"vboc_b200.shim.SafeMPC.{variant}.doublependulum_class_fixedveldir")
    rng = np.random.default_rng(0)
    params = [rng.normal(size=s).astype(np.float32) for s in ((30, 4), (30,), (30, 30), (30,), (1, 30), (1,))]
    ocp = m.OCPdoublependulumINIT(True, params, 3.14, 0.45, 2.0)
    assert ocp.SOFT_ROWS and ocp.N == 10 and callable(ocp.OCP_solve) and callable(m.SYMdoublependulumINIT)
    for i in range(ocp.N + 1):
        ocp.ocp_solver.cost_set(i, "Zl", 1e9 * np.ones((1,)) if i == 7 else np.zeros((1,)))
    assert ocp.ocp_solver.Zl[7] == 1e9 and ocp.ocp_solver.Zl.sum() == 1e9
    x = np.array([3.0, 3.2, 0.0, 0.0])
    a = np.maximum(params[0].astype(float) @ np.array([(3.0 - 3.14) / 0.45, (3.2 - 3.14) / 0.45, 0, 0]) + params[1], 0)
    a = np.maximum(params[2].astype(float) @ a + params[3], 0)
    want = float((params[4].astype(float) @ a + params[5])[0]) * 0.98 - 1e-3
    assert abs(ocp.nn_decisionfunction(params, 3.14, 0.45, 2.0, x) - want) < 1e-12


def test_my_nn_mirror():
    torch = pytest.importorskip("torch")
    from vboc_b200.shim.my_nn import NeuralNetCLS, NeuralNetDIR
    x = torch.randn(5, 6)
    assert NeuralNetCLS(6, 500, 2)(x).shape == (5, 2) and (NeuralNetDIR(6, 500, 1)(x) >= 0).all()
