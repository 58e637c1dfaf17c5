"""Mirror of the reference's `VBOC/Safe MPC/triplependulum_class_vboc.py:8-300`: the Safe-MPC OCPs of the TRIPLE pendulum
-- `OCPtriplependulumSTD` (tracking only), `OCPtriplependulumHardTerm` (the learned margin as a hard terminal constraint,
:197-239) and `OCPtriplependulumSoftTraj` (the margin at every stage, softened, penalties by `cost_set(i, "Zl", ..)`,
:242-300; driven by `VBOC/Safe MPC/soft_traj_constraints/3dof_sym.py:112-116`) -- solved by the CUDA engine's MPC family
(SURVEY 8(f)4).  Same class names, constructor arguments and `OCP_solve(x0, x_sol_guess, u_sol_guess)`.

The reference's constraint function uses `vel_norm = norm_2(x[2:])` for this 6-state system as well (:217, :282), which
includes theta_3; the engine reproduces exactly that expression (`vboc_set_mpc_velnorm_start(2)`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200 import engine  # noqa: E402
from vboc_b200.shim._acados_like import NS, SimSolverShim  # noqa: E402
from vboc_b200.shim.SafeMPC.doublependulum_class_fixedveldir import _MpcSolverShim, _np_params  # noqa: E402

N_DOF = 3


class MODELtriplependulum:
    def __init__(self, time_step, tot_time):
        self.m1 = self.m2 = self.m3 = 0.4
        self.l1 = self.l2 = self.l3 = 0.8
        self.g = 9.81
        self.time_step = time_step
        self.tot_time = tot_time


class SYMtriplependulum(MODELtriplependulum):
    def __init__(self, time_step, tot_time, regenerate):
        super().__init__(time_step, tot_time)
        self.acados_integrator = SimSolverShim(N_DOF, T=time_step)


class OCPtriplependulum(MODELtriplependulum):
    SOFT_ROWS = False

    def __init__(self, nlp_solver_type, time_step, tot_time):
        super().__init__(time_step, tot_time)
        self.Tf = tot_time
        self.N = int(tot_time / time_step)
        self.nx, self.nu = 6, 3
        self.ny, self.ny_e = self.nx + self.nu, self.nx
        Q = np.array([1e-4, 1e4, 1e-4, 1e-4, 1e-4, 1e-4])
        R = np.array([1e-4, 1e-4, 1e-4])
        self._W, self._W_e = np.concatenate([Q, R]), Q.copy()
        self.Cmax = 10.
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.
        self.Cmax_limits = np.full(3, self.Cmax)
        self.Cmin_limits = -self.Cmax_limits
        self.Xmax_limits = np.array([self.thetamax] * 3 + [self.dthetamax] * 3)
        self.Xmin_limits = np.array([self.thetamin] * 3 + [-self.dthetamax] * 3)
        # options (:155-161): acados' default tolerances (no `tol` here), Levenberg-Marquardt 1e-2
        self.opts = engine.default_opts("mpc")
        self.opts.tol_stat = self.opts.tol_eq = self.opts.tol_ineq = self.opts.tol_comp = 1e-6
        self.opts.qp_tol_stat = 1e-6
        self.opts.qp_tol_eq = self.opts.qp_tol_ineq = self.opts.qp_tol_comp = 1e-8
        self.opts.levenberg_marquardt = 1e-2
        yref = np.array([np.pi, self.thetamax - 0.05, np.pi, 0., 0., 0., 0., 0., 0.])
        self.ocp = NS(dims=NS(N=self.N, nx=self.nx, nu=self.nu), cost=NS(yref=yref, yref_e=yref[:6].copy()),
                      solver_options=NS(nlp_solver_type=nlp_solver_type, tf=tot_time, levenberg_marquardt=1e-2))
        # tracking only unless a subclass installs the margin: a one-unit zero network and a bound that never binds
        self._params = [np.zeros((1, 6)), np.zeros(1), np.zeros((1, 1)), np.zeros(1), np.zeros((1, 1)), np.zeros(1)]
        self._mean, self._std, self._margin_pct, self._lh, self._vstart = 0.0, 1.0, 0.0, -1e12, 2

    def _make_solver(self):
        self.ocp_solver = _MpcSolverShim(self)

    def OCP_solve(self, x0, x_sol_guess, u_sol_guess):
        s = self.ocp_solver
        s.reset()
        s.constraints_set(0, "lbx", x0)
        s.constraints_set(0, "ubx", x0)
        for i in range(self.ocp.dims.N):
            s.set(i, 'x', x_sol_guess[i])
            s.set(i, 'u', u_sol_guess[i])
        s.set(self.ocp.dims.N, 'x', x_sol_guess[self.ocp.dims.N])
        return s.solve()

    def _margin(self, params, mean, std, scale, x):
        """Numeric twin of nn_decisionfunction[_conservative] (:215-239, :280-300), vel_norm over x[2:] as written there."""
        W1, b1, W2, b2, W3, b3 = _np_params(params)
        x = np.asarray(x, dtype=float).ravel()
        vel_norm = max(np.linalg.norm(x[2:]), 1e-3)
        out = (x - np.array([mean] * 3 + [0.] * 3)) / np.array([std] * 3 + [vel_norm] * 3)
        out = np.maximum(W1 @ out + b1, 0.)
        out = np.maximum(W2 @ out + b2, 0.)
        return float(np.ravel(W3 @ out + b3)[0] * scale - vel_norm)


class OCPtriplependulumSTD(OCPtriplependulum):
    def __init__(self, nlp_solver_type, time_step, tot_time, regenerate):
        super().__init__(nlp_solver_type, time_step, tot_time)
        self._make_solver()


class OCPtriplependulumHardTerm(OCPtriplependulum):
    def __init__(self, nlp_solver_type, time_step, tot_time, nn_params, mean, std, regenerate):
        super().__init__(nlp_solver_type, time_step, tot_time)
        self._params, self._mean, self._std, self._lh = _np_params(nn_params), float(mean), float(std), 0.0
        self._make_solver()

    def nn_decisionfunction(self, params, mean, std, x):
        return self._margin(params, mean, std, 1.0, x)


class OCPtriplependulumSoftTraj(OCPtriplependulum):
    SOFT_ROWS = True

    def __init__(self, nlp_solver_type, time_step, tot_time, nn_params, mean, std, safety_margin, regenerate):
        super().__init__(nlp_solver_type, time_step, tot_time)
        self._params, self._mean, self._std, self._lh = _np_params(nn_params), float(mean), float(std), 0.0
        self._margin_pct = float(safety_margin)
        self._make_solver()

    def nn_decisionfunction_conservative(self, params, mean, std, safety_margin, x):
        return self._margin(params, mean, std, (100 - safety_margin) / 100, x)
