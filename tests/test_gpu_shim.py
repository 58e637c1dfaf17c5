"""The drop-in classes called the way the reference's drivers call them, against the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(sub, mod):
    path = os.path.join(ROOT, "vboc_b200", "shim", sub)
    sys.path.insert(0, path)
    try:
        sys.modules.pop(mod, None)
        return importlib.import_module(mod)
    finally:
        sys.path.remove(path)


def test_ocp_solve_like_data_generation(oracle):
    """VBOC/triplependulum_vboc.py:86-129: guess, OCP_solve, get_cost, get(i, 'x'), then one extension step."""
    m = _load("VBOC", "triplependulum_class_vboc")
    ocp = m.OCPtriplependulumINIT()
    bp = pr.sample_vboc(3, 3, seed=9)
    for b in range(3):
        one = pr.take(bp, b)
        N = 100
        ocp.N = N
        ocp.ocp_solver.set_new_time_steps(np.full((N,), 1.))
        ocp.ocp_solver.update_qp_solver_cond_N(N)
        status = ocp.OCP_solve(one["x_guess"][:N], one["u_guess"], one["p"], one["lbx"], one["ubx"], one["lbu"],
                               one["ubu"], one["lbx0"], one["ubx0"], one["lbxN"], one["ubxN"])
        ref = oracle.solve(3, 0, 0, one)
        assert status == ref["status"]
        if status != 0:
            continue
        x_sol = np.array([ocp.ocp_solver.get(i, "x") for i in range(N + 1)])
        u_sol = np.array([ocp.ocp_solver.get(i, "u") for i in range(N)])
        assert np.abs(x_sol - ref["x"]).max() < 1e-5 and np.abs(u_sol - ref["u"]).max() < 1e-4
        assert abs(ocp.ocp_solver.get_cost() - ref["cost"]) < 1e-6
        # horizon extension with warm start (:119-136)
        x_g = np.vstack([x_sol, x_sol[-1]])
        u_g = np.vstack([u_sol, np.zeros(3)])
        ocp.N = N + 1
        ocp.ocp_solver.set_new_time_steps(np.full((N + 1,), 1.))
        st2 = ocp.OCP_solve(x_g, u_g, one["p"], one["lbx"], one["ubx"], one["lbu"], one["ubu"], one["lbx0"],
                            one["ubx0"], one["lbxN"], one["ubxN"])
        one2 = dict(one)
        one2["x_guess"], one2["u_guess"] = pr.expand_guess(x_g, u_g, N + 1)
        ref2 = oracle.solve(3, 0, 0, one2)
        assert st2 == ref2["status"]
        if st2 == 0:
            assert abs(ocp.ocp_solver.get_cost() - ref2["cost"]) < 1e-6
            assert ocp.ocp_solver.get_cost() <= ref["cost"] + 1e-6  # a longer horizon cannot do worse


def test_sim_integrator_like_the_driver(oracle):
    """VBOC/triplependulum_vboc.py:348-352"""
    m = _load("VBOC", "triplependulum_class_vboc")
    sim = m.SYMtriplependulumINIT()
    x = np.array([3.0, 3.2, 2.9, 1.0, -2.0, 0.5])
    u = np.array([1.0, -3.0, 2.0])
    sim.acados_integrator.set("u", u)
    sim.acados_integrator.set("x", x)
    sim.acados_integrator.set("T", 1e-2)
    assert sim.acados_integrator.solve() == 0
    assert np.abs(sim.acados_integrator.get("x") - oracle.rk4(3, 1, x, u, 1e-2)).max() < 1e-12


@pytest.mark.parametrize("mod,cls,n", [("pendulum_class_al", "OCPpendulumINIT", 1),
                                       ("doublependulum_class_al", "OCPdoublependulumINIT", 2),
                                       ("triplependulum_class_al", "OCPtriplependulumINIT", 3)])
def test_compute_problem_labels(oracle, mod, cls, n):
    """AL/triplependulum_al.py:24-41 `testing`"""
    ocp = getattr(_load("AL", mod), cls)()
    bp = pr.sample_al(n, 12, seed=6)
    ref = oracle.solve_batch(n, 1, 1, bp)
    for b in range(12):
        lab = ocp.compute_problem(bp["x0"][b, :n], bp["x0"][b, n:])
        want = 1 if ref["status"][b] == 0 else (0 if ref["status"][b] == 4 else 2)
        assert lab == want
        if lab == 1:
            traj = np.array([ocp.ocp_solver.get(i, "x") for i in range(ocp.N + 1)])
            assert np.abs(traj - ref["x"][b]).max() < 1e-6


def test_pendulum_testdata_like_the_script(oracle):
    """pendulum_testdata.py:7-53: pinned dt, p = [+-1, 0], one solve, N = 50; and the free-dt OCP_solve of
    VBOC/pendulum_vboc.py is refused, not mis-solved."""
    from vboc_b200._lib import VbocError
    ocp = _load("VBOC", "pendulum_class_vboc").OCPpendulum()
    N, dt = ocp.N, 1e-2
    q_min, q_max, v_max = ocp.thetamin, ocp.thetamax, ocp.dthetamax
    rng = np.random.default_rng(3)
    for trial in range(4):
        p = np.array([rng.choice([-1.0, 1.0]), 0.0])
        q_init = q_min + rng.random() * (q_max - q_min)
        xg = np.full((N, 3), np.array([q_init, 0.0, dt]))
        s = ocp.ocp_solver
        s.reset()
        for i in range(N):
            s.set(i, "x", xg[i])
            s.set(i, "p", p)
            s.constraints_set(i, "lbx", np.array([q_min, -v_max, dt]))
            s.constraints_set(i, "ubx", np.array([q_max, v_max, dt]))
        s.constraints_set(0, "lbx", np.array([q_init, -v_max, dt]))
        s.constraints_set(0, "ubx", np.array([q_init, v_max, dt]))
        s.constraints_set(N, "lbx", np.array([q_min, 0.0, dt]))
        s.constraints_set(N, "ubx", np.array([q_max, 0.0, dt]))
        s.set(N, "x", xg[-1])
        s.set(N, "p", p)
        status = s.solve()
        prob = dict(x_guess=np.vstack([xg, xg[-1:]]), u_guess=np.zeros((N, 1)), p=p,
                    lbx0=np.array([q_init, -v_max, dt]), ubx0=np.array([q_init, v_max, dt]),
                    lbx=np.array([q_min, -v_max, dt]), ubx=np.array([q_max, v_max, dt]),
                    lbxN=np.array([q_min, 0.0, dt]), ubxN=np.array([q_max, 0.0, dt]),
                    lbu=np.array([-3.0]), ubu=np.array([3.0]), C0=None)
        ref = oracle.solve(1, 0, 0, prob)
        assert status == ref["status"]
        if status == 0:
            assert np.abs(s.get(0, "x") - ref["x"][0]).max() < 1e-5


def test_pendulum_vboc_free_dt_like_the_driver():
    """VBOC/pendulum_vboc.py:54-140: the two extreme trajectories with dt a free state in [0, 1e-2] and a
    unit weight on time; checked against the NLP itself (constraints, bang-bang structure) and against the
    semi-analytic boundary: from (q_max, v) with full braking torque the pendulum must stop exactly at q_min."""
    ocp = _load("VBOC", "pendulum_class_vboc").OCPpendulum()
    N = ocp.N
    q_min, q_max, v_max = ocp.thetamin, ocp.thetamax, ocp.dthetamax
    for v_sel in (-v_max, v_max):
        if v_sel < 0:
            q_init, q_fin, lb, ub, cd = q_max, q_min, np.array([q_min, -v_max, 0.]), np.array([q_max, 0., 1e-2]), 1.
        else:
            q_init, q_fin, lb, ub, cd = q_min, q_max, np.array([q_min, 0., 0.]), np.array([q_max, v_max, 1e-2]), -1.
        xg = np.empty((N + 1, 3))
        xg[:, 0], xg[:, 1], xg[:, 2] = np.linspace(q_init, q_fin, N + 1), v_sel, 1e-2
        status = ocp.OCP_solve(xg, np.zeros((N, 1)), cd, lb, ub, q_init, q_fin)
        assert status == 0
        x = np.array([ocp.ocp_solver.get(i, "x") for i in range(N + 1)])
        u = np.array([ocp.ocp_solver.get(i, "u") for i in range(N)])
        assert abs(x[0, 0] - q_init) < 1e-8 and abs(x[N, 0] - q_fin) < 1e-6 and abs(x[N, 1]) < 1e-6
        assert np.abs(u).max() <= 3.0 + 1e-9 and (x[:, 2] >= -1e-12).all() and (x[:, 2] <= 1e-2 + 1e-12).all()
        assert np.abs(np.diff(x[:, 2])).max() < 1e-9          # dt is constant along the horizon (dt' = 0)
        assert abs(abs(x[0, 1]) - v_max) < 1e-6               # the velocity limit is reachable from the far limit
        assert abs(ocp.ocp_solver.get_cost() - (cd * x[0, 1] + x[:N, 2].sum())) < 1e-9


@pytest.mark.parametrize("mod,cls,n", [("doublependulum_class_al", "OCPdoublependulumINIT", 2),
                                       ("triplependulum_class_al", "OCPtriplependulumINIT", 3)])
def test_compute_problem_nnguess_labels(oracle, mod, cls, n):
    """AL/triplependulum_class_al.py:171-201 called as AL/triplependulum_al.py:45-62 does: the guess network's
    (de-normalised) output becomes the state guess of stages 1..N; label and trajectory against the oracle started
    from the same guess."""
    import torch
    from vboc_b200.shim.my_nn import NeuralNetCLS
    torch.manual_seed(n)
    ocp = getattr(_load("AL", mod), cls)()
    N, nx = ocp.N, 2 * n
    model = NeuralNetCLS(nx, 64, N * nx)
    with torch.no_grad():           # small weights: the predicted trajectory stays near the (normalised) state mean
        for prm in model.parameters():
            prm.mul_(0.05)
    bp = pr.sample_al(n, 10, seed=8)
    X = bp["x0"][np.all(np.abs(bp["x0"][:, n:]) <= ocp.dthetamax, axis=1)]
    mean, std = torch.tensor(float(X.mean())), torch.tensor(float(X.std()))
    for x0 in X:
        lab = ocp.compute_problem_nnguess(x0[:n], x0[n:], model, mean, std)
        with torch.no_grad():
            out = (model((torch.Tensor([x0.tolist()]) - mean) / std) * std + mean).numpy().reshape(N, nx)
        xg = np.vstack([x0[None], out.astype(np.float64)])[None]
        ref = oracle.solve_batch(n, 1, 1, pr.al_problems(n, x0[None], x_guess=xg))
        want = 1 if ref["status"][0] == 0 else (0 if ref["status"][0] == 4 else 2)
        assert lab == want
        if lab == 1:
            traj = np.array([ocp.ocp_solver.get(i, "x") for i in range(N + 1)])
            assert np.abs(traj - ref["x"][0]).max() < 1e-6
