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
