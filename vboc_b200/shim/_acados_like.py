"""acados-shaped facades over the C-ABI: the subset of `AcadosOcpSolver` / `AcadosSimSolver` the
reference's drivers call (`reset, set, constraints_set, solve, get, get_cost, set_new_time_steps,
update_qp_solver_cond_N`; `acados_integrator.set / solve / get`), VBOC/triplependulum_vboc.py:24-25,
110-129, 181-184, 348-352; triplependulum_testdata.py:47-75."""
import numpy as np

from .. import engine
from .._lib import MODE_RTI, MODE_SQP

N_MAX = 128


class NS:
    """attribute bag standing in for AcadosOcp (`ocp.dims.nx`, `ocp.solver_options.nlp_solver_tol_stat`, ...)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class OcpSolverShim:
    def __init__(self, n, family, N, lbx, ubx, lbu, ubu, lbx_e, ubx_e, lbx_0, ubx_0, mode, Tf=1.0, device=0):
        self.n, self.family, self.N, self.mode, self.Tf, self.device = n, family, int(N), mode, Tf, device
        self.nx = 2 * n + (family == "vboc")
        self.nu = n
        S = N_MAX + 1
        self.x = np.zeros((S, self.nx))
        self.u = np.zeros((S, self.nu))
        self.p = np.zeros((S, n + 1))
        self.lbx = np.tile(np.asarray(lbx, float), (S, 1))
        self.ubx = np.tile(np.asarray(ubx, float), (S, 1))
        self.lbu = np.tile(np.asarray(lbu, float), (S, 1))
        self.ubu = np.tile(np.asarray(ubu, float), (S, 1))
        self.lbx[0], self.ubx[0] = lbx_0, ubx_0
        self._lbx_e, self._ubx_e = np.asarray(lbx_e, float), np.asarray(ubx_e, float)
        self.lbx[self.N], self.ubx[self.N] = self._lbx_e, self._ubx_e
        self.C = np.zeros((S, n, self.nx))
        self.opts = engine.default_opts(family)
        self._sol = None
        self._out = None
        self.cartesian = None   # dict(xc, yc, radius, uh): the Cartesian path constraint (VBOC family, n = 2)

    # -- acados API ----------------------------------------------------------------------------
    def reset(self):
        self.x[:] = 0.0
        self.u[:] = 0.0
        self._out = None

    def set(self, stage, field, value):
        v = np.asarray(value, dtype=float).ravel()
        if field == "x":
            self.x[stage] = v
        elif field == "u":
            self.u[stage] = v
        elif field == "p":
            self.p[stage] = v
        else:
            raise NotImplementedError(f"set(.., {field!r}, ..) is not part of the hot path")

    def constraints_set(self, stage, field, value, api="warn"):
        v = np.asarray(value, dtype=float)
        if field in ("lbx", "ubx", "lbu", "ubu"):
            getattr(self, field)[stage] = v.ravel()
        elif field == "C":
            self.C[stage] = v.reshape(self.n, self.nx)
        elif field in ("D", "lg", "ug"):
            if np.any(v != 0):
                raise NotImplementedError("only D = 0, lg = ug = 0 (the reference's stage-0 direction constraint)")
        else:
            raise NotImplementedError(f"constraints_set(.., {field!r}, ..) is not part of the hot path")

    def set_new_time_steps(self, steps):
        N = len(steps)
        if N > N_MAX:
            raise ValueError(f"horizon {N} exceeds N_MAX = {N_MAX}")
        if N != self.N:
            # acados re-creates the solver with default terminal bounds at the new last stage
            self.lbx[N], self.ubx[N] = self._lbx_e, self._ubx_e
        self.N = N

    def update_qp_solver_cond_N(self, N):
        pass  # no condensing here either

    def _problem(self):
        N = self.N
        if np.any(self.C[1:N] != 0):
            raise NotImplementedError("general constraints are supported at stage 0 only")
        one = lambda a: np.ascontiguousarray(a[None])
        bp = dict(n=self.n, family=self.family, N=np.array([N], dtype=np.int32), Tf=self.Tf,
                  x_guess=one(self.x[:N_MAX + 1]), u_guess=one(self.u[:N_MAX]),
                  lbx0=one(self.lbx[0]), ubx0=one(self.ubx[0]), lbx=one(self.lbx[1]), ubx=one(self.ubx[1]),
                  lbxN=one(self.lbx[N]), ubxN=one(self.ubx[N]), lbu=one(self.lbu[0]), ubu=one(self.ubu[0]),
                  p=None, C0=None)
        if N > 1 and (np.any(self.lbx[1:N] != self.lbx[1]) or np.any(self.ubx[1:N] != self.ubx[1])
                      or np.any(self.lbu[:N] != self.lbu[0]) or np.any(self.ubu[:N] != self.ubu[0])):
            raise NotImplementedError("stage-dependent path bounds are not part of the hot path")
        if self.family == "vboc":
            bp["p"] = one(self.p[0])
            if np.any(self.C[0] != 0):
                bp["C0"] = one(self.C[0])
        return bp

    def solve(self):
        if self._sol is None:
            self._sol = engine.BatchSolver(self.n, self.family, 1, N_MAX, device=self.device)
            if self.cartesian is not None:
                self._sol.set_cartesian(**self.cartesian)
        self._sol.set_opts(self.opts)
        self._out = self._sol.solve(self._problem(), MODE_RTI if self.mode == "SQP_RTI" else MODE_SQP)
        N = self.N
        self.x[:N + 1] = self._out["x"][0, :N + 1]
        self.u[:N] = self._out["u"][0, :N]
        return int(self._out["status"][0])

    def get(self, stage, field):
        if field == "x":
            return self.x[stage].copy()
        if field == "u":
            return self.u[stage].copy()
        raise NotImplementedError(f"get(.., {field!r}) is not part of the hot path")

    def get_cost(self):
        if self._out is None:
            raise RuntimeError("get_cost() before solve()")
        return float(self._out["cost"][0])

    def get_stats(self, field):
        k = {"sqp_iter": "sqp_iter", "qp_iter": "qp_iter"}.get(field)
        if k is None or self._out is None:
            raise NotImplementedError(field)
        return int(self._out[k][0])


class SimSolverShim:
    """`sim.acados_integrator`: one RK4 step of the unscaled model, T = 1e-2 by default."""

    def __init__(self, n, T=1e-2, device=0):
        self.n, self.T, self.device = n, T, device
        self._x = np.zeros(2 * n)
        self._u = np.zeros(n)

    def set(self, field, value):
        if field == "x":
            self._x = np.asarray(value, dtype=float).ravel().copy()
        elif field == "u":
            self._u = np.asarray(value, dtype=float).ravel().copy()
        elif field == "T":
            self.T = float(value)
        else:
            raise NotImplementedError(field)

    def solve(self):
        self._x = engine.sim_step(self.n, self._x[None], self._u[None], self.T, self.device)[0]
        return 0

    def get(self, field):
        if field != "x":
            raise NotImplementedError(field)
        return self._x.copy()
