"""Mirror of the reference's `VBOC/pendulum_class_vboc.py:8-130` (1-DOF).

Both uses work: the pinned-`dt` one of `pendulum_testdata.py:7-53` (`lbx = ubx = dt_sym` on the third state at
every stage, `p = [+-1, 0]`; the dt state is eliminated) and `OCP_solve` as `VBOC/pendulum_vboc.py:91` calls it,
where `dt` is a free state in [0, 1e-2] with cost weight 1 (lane-per-OCP kernel with the dt state kept and
the two terminal equalities handled by bordering, vboc_b200/csrc/ocp_lane.h)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200.shim._acados_like import NS, OcpSolverShim  # noqa: E402


class OCPpendulum:
    def __init__(self):
        self.m, self.g, self.d, self.b = 0.5, 9.81, 0.3, 0.01
        self.N = 50
        self.Fmax = 3
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.0
        self.ocp = NS(dims=NS(N=self.N, nx=3, nu=1, np=2),
                      solver_options=NS(nlp_solver_type="SQP", nlp_solver_tol_stat=1e-3, qp_solver_tol_stat=1e-3,
                                        qp_solver_iter_max=100, nlp_solver_max_iter=1000,
                                        globalization="MERIT_BACKTRACKING", alpha_reduction=0.3, alpha_min=1e-2,
                                        levenberg_marquardt=1e-5, tf=self.N))
        lbx = np.array([self.thetamin, -self.dthetamax, 0.])
        ubx = np.array([self.thetamax, self.dthetamax, 1e-2])
        self.ocp_solver = OcpSolverShim(1, "vboc", self.N, lbx, ubx, [-self.Fmax], [self.Fmax], lbx, ubx, lbx, ubx, "SQP")
        for i in range(self.N + 1):
            self.ocp_solver.set(i, 'p', np.array([0., 1.]))

    def OCP_solve(self, x_sol_guess, u_sol_guess, cost_dir, q_lb, q_ub, q_init, q_fin):
        s = self.ocp_solver
        s.reset()
        for i in range(self.N):
            s.set(i, "x", np.array(x_sol_guess[i]))
            s.set(i, "u", np.array(u_sol_guess[i]))
            s.set(i, 'p', np.array([cost_dir, 1.]))
            s.constraints_set(i, "lbx", q_lb)
            s.constraints_set(i, "ubx", q_ub)
        s.constraints_set(0, "lbx", np.array([q_init, -self.dthetamax, 0.]))
        s.constraints_set(0, "ubx", np.array([q_init, self.dthetamax, 1e-2]))
        s.constraints_set(self.N, "lbx", np.array([q_fin, 0., 0.]))
        s.constraints_set(self.N, "ubx", np.array([q_fin, 0., 1e-2]))
        s.set(self.N, "x", np.array(x_sol_guess[self.N]))
        s.set(self.N, 'p', np.array([cost_dir, 1.]))
        return s.solve()
