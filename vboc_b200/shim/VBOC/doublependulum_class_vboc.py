"""Mirror of the reference's `VBOC/doublependulum_class_vboc.py:8-181, 222-304`: same class names, attributes and methods, solved by
the CUDA engine instead of acados (see vboc_b200/shim/__init__.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200.shim._acados_like import NS, OcpSolverShim, SimSolverShim  # noqa: E402

N_DOF = 2


class OCPdoublependulum:
    def __init__(self):
        self.m1 = self.m2 = 0.4
        self.l1 = self.l2 = 0.8
        self.g = 9.81
        self.N = 100
        self.Cmax = 10.
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.
        n = N_DOF
        self.ocp = NS(dims=NS(N=self.N, nx=5, nu=n, np=n + 1),
                      solver_options=NS(nlp_solver_type="SQP", nlp_solver_tol_stat=1e-3, qp_solver_tol_stat=1e-3,
                                        qp_solver_iter_max=100, nlp_solver_max_iter=1000,
                                        globalization="MERIT_BACKTRACKING", alpha_reduction=0.3, alpha_min=1e-2,
                                        levenberg_marquardt=1e-5, tf=self.N))


class OCPdoublependulumINIT(OCPdoublependulum):
    def __init__(self):
        super().__init__()
        n = N_DOF
        lbx = np.array([self.thetamin] * n + [-self.dthetamax] * n + [0.])
        ubx = np.array([self.thetamax] * n + [self.dthetamax] * n + [1e-2])
        lbu, ubu = np.full(n, -self.Cmax), np.full(n, self.Cmax)
        self.ocp_solver = OcpSolverShim(n, "vboc", self.N, lbx, ubx, lbu, ubu, lbx, ubx, lbx, ubx, "SQP")

    def OCP_solve(self, x_sol_guess, u_sol_guess, p, q_lb, q_ub, u_lb, u_ub, q_init_lb, q_init_ub, q_fin_lb, q_fin_ub):
        n = N_DOF
        s = self.ocp_solver
        if self.N != s.N:  # the drivers mutate ocp.N and call set_new_time_steps before OCP_solve
            s.set_new_time_steps(np.full((self.N,), 1.))
        s.reset()
        for i in range(self.N):
            s.set(i, 'x', x_sol_guess[i])
            s.set(i, 'u', u_sol_guess[i])
            s.set(i, 'p', p)
            s.constraints_set(i, 'lbx', q_lb)
            s.constraints_set(i, 'ubx', q_ub)
            s.constraints_set(i, 'lbu', u_lb)
            s.constraints_set(i, 'ubu', u_ub)
            s.constraints_set(i, 'C', np.zeros((n, 2 * n + 1)))
            s.constraints_set(i, 'D', np.zeros((n, n)))
            s.constraints_set(i, 'lg', np.zeros((n)))
            s.constraints_set(i, 'ug', np.zeros((n)))
        C = np.zeros((n, 2 * n + 1))
        d = np.array([np.asarray(p)[:n].tolist()])
        C[:, n:2 * n] = np.identity(n) - np.matmul(np.transpose(d), d)
        s.constraints_set(0, "C", C, api='new')
        s.constraints_set(0, "lbx", q_init_lb)
        s.constraints_set(0, "ubx", q_init_ub)
        s.constraints_set(self.N, "lbx", q_fin_lb)
        s.constraints_set(self.N, "ubx", q_fin_ub)
        s.set(self.N, 'x', x_sol_guess[-1])
        s.set(self.N, 'p', p)
        return s.solve()


class SYMdoublependulum:
    def __init__(self):
        self.m1 = self.m2 = 0.4
        self.l1 = self.l2 = 0.8
        self.g = 9.81


class SYMdoublependulumINIT(SYMdoublependulum):
    def __init__(self):
        super().__init__()
        self.acados_integrator = SimSolverShim(N_DOF, 1e-2)
