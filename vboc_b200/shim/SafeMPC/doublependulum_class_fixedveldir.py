"""Mirror of the reference's `VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:7-276`: the
Safe-MPC OCP of the double pendulum with the learned viability margin as a hard nonlinear TERMINAL constraint
(`con_h_expr_e = nn_decisionfunction(...)`, `lh_e = 0`, `uh_e = 1e6`, :176-179), solved by the CUDA engine's MPC family
(`vboc_set_mpc`) instead of acados.  Same class names, constructor arguments, attributes and methods:
`OCPdoublependulumINIT(regenerate, nn_params, mean, std, safety_margin)`, `OCP_solve(x0, q_ref, x_sol_guess,
u_sol_guess)`, `ocp_solver.get / cost_set / constraints_set`, `nn_decisionfunction`, `SYMdoublependulumINIT`.
`nlp_solver_type` is not set by the reference class (the line is commented out, :183), so acados' default SQP_RTI
applies: one linearisation + one QP per call."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200 import engine  # noqa: E402
from vboc_b200._lib import MODE_RTI, MODE_SQP  # noqa: E402
from vboc_b200.shim._acados_like import NS, SimSolverShim  # noqa: E402

N_DOF = 2


def _np_params(nn_params):
    """list(model.parameters()) (torch) or arrays -> [W1, b1, W2, b2, W3, b3] as float64 numpy."""
    out = []
    for p in nn_params:
        a = p.detach().cpu().numpy() if hasattr(p, "detach") else np.asarray(p)
        out.append(np.asarray(a, dtype=np.float32).astype(np.float64))
    return out


class OCPdoublependulum:
    def __init__(self):
        self.m1 = self.m2 = 0.4
        self.l1 = self.l2 = 0.8
        self.g = 9.81


class _MpcSolverShim:
    """The slice of AcadosOcpSolver the Safe-MPC driver uses (VBOC/Safe MPC/hard_terminal_constraints/2dof_sym.py):
    reset, set, cost_set(i, 'y_ref' | 'W'), constraints_set(0, 'lbx' | 'ubx'), solve, get."""

    def __init__(self, owner):
        self.o = owner
        N, nx, nu = owner.N, owner.nx, owner.nu
        self.x, self.u = np.zeros((N + 1, nx)), np.zeros((N, nu))
        self.yref = np.tile(owner.ocp.cost.yref, (N, 1))
        self.yref_e = owner.ocp.cost.yref_e.copy()
        self.x0 = np.zeros(nx)
        self.Zl = np.zeros(N + 1)    # cost_set(i, "Zl", ...) of the soft-row variants (zl = zu = Zu = 0 in the reference)
        self._sol = self._out = self._rows = None

    def reset(self):
        self.x[:] = 0.0
        self.u[:] = 0.0

    def set(self, stage, field, value):
        (self.x if field == "x" else self.u)[stage] = np.asarray(value, dtype=float).ravel()

    def cost_set(self, stage, field, value):
        v = np.asarray(value, dtype=float)
        if field == "y_ref":
            if stage == self.o.N:
                self.yref_e = v.ravel().copy()
            else:
                self.yref[stage] = v.ravel()
        elif field == "W":
            d = np.diag(v) if v.ndim == 2 else v
            if v.ndim == 2 and np.any(v != np.diag(d)):
                raise NotImplementedError("only diagonal weights (the reference's Q, R)")
            if stage == self.o.N:
                self.o._W_e = d.copy()
            else:
                self.o._W = d.copy()
            self._sol_dirty = True
        elif field == "Zl" and self.o.SOFT_ROWS:
            self.Zl[stage] = float(np.ravel(v)[0])
        else:
            raise NotImplementedError(f"cost_set(.., {field!r}, ..) is not part of this path")

    def constraints_set(self, stage, field, value):
        if stage != 0 or field not in ("lbx", "ubx"):
            raise NotImplementedError("only the initial state is set at run time (OCP_solve :204-205)")
        self.x0 = np.asarray(value, dtype=float).ravel().copy()

    def solve(self):
        o = self.o
        if np.any(self.yref != self.yref[0]):
            raise NotImplementedError("one y_ref for all stages (OCP_solve sets the same one everywhere, :208-213)")
        if self._sol is None or getattr(self, "_sol_dirty", False):
            if self._sol is None:
                self._sol = engine.BatchSolver(o.nu, "mpc", 1, o.N)
            self._sol.set_mpc(dict(zip(("W1", "b1", "W2", "b2", "W3", "b3"), o._params)), o._mean, o._std, o._margin_pct,
                              o._W, o._W_e, lh=getattr(o, "_lh", 0.0), uh=1e6, vstart=getattr(o, "_vstart", None))
            self._sol_dirty = False
        self._sol.set_opts(o.opts)
        self._sol.set_mpc_reference(self.yref[0], self.yref_e)
        if o.SOFT_ROWS:
            self._sol.set_mpc_rows(self.Zl[None, :])
        one = lambda a: np.ascontiguousarray(np.asarray(a, dtype=float)[None])
        nq = o.nu
        lo = np.concatenate([[o.thetamin] * nq, [-o.dthetamax] * nq])
        hi = np.concatenate([[o.thetamax] * nq, [o.dthetamax] * nq])
        bp = dict(n=nq, family="mpc", N=np.array([o.N], dtype=np.int32), Tf=o.Tf, x_guess=one(self.x), u_guess=one(self.u),
                  p=None, C0=None, lbx0=one(self.x0), ubx0=one(self.x0), lbx=one(lo), ubx=one(hi), lbxN=one(lo),
                  ubxN=one(hi), lbu=one([-o.Cmax] * nq), ubu=one([o.Cmax] * nq))
        self._out = self._sol.solve(bp, MODE_RTI if o.ocp.solver_options.nlp_solver_type == "SQP_RTI" else MODE_SQP)
        self.x[:] = self._out["x"][0, :o.N + 1]
        self.u[:] = self._out["u"][0, :o.N]
        self._rows = self._sol.mpc_rows()[0] if o.SOFT_ROWS else None
        return int(self._out["status"][0])

    def get(self, stage, field):
        if field in ("sl", "su"):      # slacks of the softened margin row
            return self._rows[stage, 4 + (field == "su")].reshape(1).copy()
        return (self.x if field == "x" else self.u)[stage].copy()

    def get_cost(self):
        return float(self._out["cost"][0])


class OCPdoublependulumINIT(OCPdoublependulum):
    # False: hard terminal row (this module).  True (the subclasses under parallel/, receiding_hard_constraints/,
    # soft_traj_constraints/): the row at every stage with slacks, penalties by cost_set(i, "Zl", ..)
    SOFT_ROWS = False

    def __init__(self, regenerate, nn_params, mean, std, safety_margin):
        super().__init__()
        self.Tf = 0.01
        self.N = int(1000 * self.Tf)
        self.nx, self.nu = 4, 2
        self.ny, self.ny_e = self.nx + self.nu, self.nx
        Q = np.array([1e4, 1e4, 1e-4, 1e-4])
        R = np.array([1e-4, 1e-4])
        self._W, self._W_e = np.concatenate([Q, R]), Q.copy()       # diagonals of cost.W ([x; u]) and cost.W_e
        self.Cmax = 10.
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.
        self._params = _np_params(nn_params)
        self._mean, self._std = float(mean), float(std)
        # the hard-terminal-constraint class passes safety_margin but its constraint function does not use it
        # (`return out - vel_norm`, :258); the parallel class applies it (`out*(100-safety_margin)/100 - vel_norm`)
        self._margin_pct = float(safety_margin) if self.SOFT_ROWS else 0.0
        self.safety_margin = safety_margin
        self.opts = engine.default_opts("mpc")
        self.ocp = NS(dims=NS(N=self.N, nx=self.nx, nu=self.nu),
                      cost=NS(yref=np.array([np.pi, np.pi, 0., 0., 0., 0.]), yref_e=np.array([np.pi, np.pi, 0., 0.])),
                      solver_options=NS(nlp_solver_type="SQP_RTI", tf=self.Tf, tol=1e-2, levenberg_marquardt=1.))
        self.ocp_solver = _MpcSolverShim(self)

    def OCP_solve(self, x0, q_ref, x_sol_guess, u_sol_guess):
        s = self.ocp_solver
        s.reset()
        s.constraints_set(0, "lbx", x0)
        s.constraints_set(0, "ubx", x0)
        for i in range(self.N):
            s.set(i, 'x', x_sol_guess[i])
            s.set(i, 'u', u_sol_guess[i])
            s.cost_set(i, 'y_ref', np.array([q_ref[0], q_ref[1], 0., 0., 0., 0.]))
        s.set(self.N, 'x', x_sol_guess[self.N])
        s.cost_set(self.N, 'y_ref', np.array([q_ref[0], q_ref[1], 0., 0.]))
        return s.solve()

    def nn_decisionfunction(self, params, mean, std, safety_margin, x):
        """Numeric twin of the CasADi expression (:232-258) for the driver's a-posteriori checks."""
        W1, b1, W2, b2, W3, b3 = _np_params(params)
        x = np.asarray(x, dtype=float).ravel()
        vel_norm = max(np.linalg.norm(x[2:]), 1e-3)
        out = (x - np.array([mean, mean, 0., 0.])) / np.array([std, std, vel_norm, vel_norm])
        out = np.maximum(W1 @ out + b1, 0.)
        out = np.maximum(W2 @ out + b2, 0.)
        out = W3 @ out + b3
        if self is not None and getattr(self, "SOFT_ROWS", False):   # the soft-row classes apply the margin (:264)
            return float(np.ravel(out)[0] * (100 - safety_margin) / 100 - vel_norm)
        return float(np.ravel(out)[0] - vel_norm)


class SYMdoublependulumINIT(OCPdoublependulum):
    def __init__(self, regenerate):
        super().__init__()
        self.acados_integrator = SimSolverShim(N_DOF, T=1e-3)
