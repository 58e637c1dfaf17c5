"""Mirror of the reference's `AL/triplependulum_class_al.py:11-222`: same class names, attributes and methods, solved by the CUDA engine
instead of acados (SQP_RTI: one linearisation + one QP; label 1 iff status 0, 0 iff status 4)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200.shim._acados_like import NS, OcpSolverShim  # noqa: E402

N_DOF = 3


class OCPtriplependulum:
    def __init__(self):
        self.m1 = self.m2 = self.m3 = 0.4
        self.l1 = self.l2 = self.l3 = 0.8
        self.g = 9.81
        self.Cmax = 10
        self.Tf = 1.
        self.N = int(100 * self.Tf)
        self.nx, self.nu = 2 * N_DOF, N_DOF
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.
        self.ocp = NS(dims=NS(N=self.N, nx=self.nx, nu=self.nu), solver_options=NS(nlp_solver_type="SQP_RTI", tf=self.Tf))

    def _label(self, status):
        return 1 if status == 0 else (0 if status == 4 else 2)

    def compute_problem(self, q0, v0):
        n, s = N_DOF, self.ocp_solver
        s.reset()
        x0 = np.concatenate([np.atleast_1d(q0).astype(float)[:n], np.atleast_1d(v0).astype(float)[:n]])
        s.constraints_set(0, "lbx", x0)
        s.constraints_set(0, "ubx", x0)
        x_guess = np.concatenate([x0[:n], np.zeros(n)])
        for i in range(self.N + 1):
            s.set(i, "x", x_guess)
        return self._label(s.solve())

    def compute_problem_nnguess(self, q0, v0, model, mean, std):
        import torch
        n, s = N_DOF, self.ocp_solver
        s.reset()
        x0 = np.concatenate([np.atleast_1d(q0).astype(float)[:n], np.atleast_1d(v0).astype(float)[:n]])
        s.constraints_set(0, "lbx", x0)
        s.constraints_set(0, "ubx", x0)
        with torch.no_grad():
            inp = torch.Tensor([x0.tolist()])
            out = (model((inp - mean) / std) * std + mean).numpy()
        out = np.reshape(out, (self.N, self.nx))
        s.set(0, "x", x0)
        for i in range(self.N):
            s.set(i + 1, "x", out[i])
        return self._label(s.solve())

    def set_bounds(self, q_max, q_min):
        n, s = N_DOF, self.ocp_solver
        self.thetamax, self.thetamin = q_max, q_min
        for i in range(1, self.N + 1):
            s.lbx[i, :n], s.ubx[i, :n] = q_min, q_max
        s._lbx_e[:n], s._ubx_e[:n] = q_min, q_max


class OCPtriplependulumINIT(OCPtriplependulum):
    def __init__(self):
        super().__init__()
        n = N_DOF
        lbx = np.array([self.thetamin] * n + [-self.dthetamax] * n)
        ubx = np.array([self.thetamax] * n + [self.dthetamax] * n)
        lbx_e = np.array([self.thetamin] * n + [0.0] * n)  # zero final velocity
        ubx_e = np.array([self.thetamax] * n + [0.0] * n)
        x0 = np.array([self.thetamin] * n + [0.0] * n)
        lbu, ubu = np.full(n, -float(self.Cmax)), np.full(n, float(self.Cmax))
        self.ocp_solver = OcpSolverShim(n, "al", self.N, lbx, ubx, lbu, ubu, lbx_e, ubx_e, x0, x0, "SQP_RTI", Tf=self.Tf)
