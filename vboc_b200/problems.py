"""Seeded, batched restatement of the reference's per-problem sampling.

The reference draws every OCP's data with the unseeded `random` module inside forked workers
(SURVEY §9), so no run is repeatable.  Here the same distributions are drawn from an explicit seed
as *arrays* (one row per problem) so that the CUDA engine, the CPU oracle and -- when it is
available -- acados all consume identical inputs.

Follows (file:line relative to /root/reference):
  sample_vboc      VBOC/triplependulum_vboc.py:33-103, VBOC/doublependulum_vboc.py:33-99
  sample_testdata  triplependulum_testdata.py:15-38, doublependulum_testdata.py:16-37,
                   pendulum_testdata.py:11-27
  sample_al        AL/triplependulum_al.py:115-123 (uniform pool; velocities 5 % beyond the box)

All arrays are reference-shaped: VBOC states are [q(n), v(n), dt], AL states [q(n), v(n)].
"""
import math

import numpy as np


class Model:
    """Constants of the reference models (constructor blocks of the *_class_*.py files)."""

    def __init__(self, n):
        self.n = n
        self.thetamax = np.pi / 4 + np.pi
        self.thetamin = -np.pi / 4 + np.pi
        self.dthetamax = 10.0
        if n == 1:
            # VBOC/pendulum_class_vboc.py:14-17, :64-67
            self.m, self.g, self.d, self.b = 0.5, 9.81, 0.3, 0.01
            self.Fmax = 3.0
            self.umax = 3.0
            self.N0 = 50
        else:
            # VBOC/doublependulum_class_vboc.py:14-18, :114-117; VBOC/triplependulum_class_vboc.py:15-21, :90-93
            self.m1 = self.m2 = self.m3 = 0.4
            self.l1 = self.l2 = self.l3 = 0.8
            self.g = 9.81
            self.Cmax = 10.0
            self.umax = 10.0
            self.N0 = 100

    def gravity_comp(self, q):
        """Torques holding the arm still at positions q (reference VBOC/doublependulum_vboc.py:84;
        generalised to n links: u_i = g l_i (sum_{k>=i} m_k) sin q_i)."""
        q = np.asarray(q, dtype=float)
        n = self.n
        if n == 1:
            return -self.m * self.g * self.d * np.sin(q)
        mu = np.array([0.4 * (n - i) for i in range(n)])
        return self.g * 0.8 * mu * np.sin(q)


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def expand_guess(x_rows, u_rows, N):
    """Reference `OCP_solve` semantics (VBOC/triplependulum_class_vboc.py:161-186): stages 0..N-1 take
    guess rows 0..N-1, stage N takes the LAST row of the state guess (the guesses the drivers pass have
    N or N+1 rows)."""
    x_rows = np.asarray(x_rows, dtype=float)
    u_rows = np.asarray(u_rows, dtype=float)
    x = np.empty((N + 1, x_rows.shape[1]))
    x[:N] = x_rows[:N]
    x[N] = x_rows[-1]
    return x, np.array(u_rows[:N], dtype=float)


def sample_vboc(n, batch, seed, N=None, dt_sym=1e-2, tol=1e-3):
    """Batched restatement of the sampling block of `data_generation` (n = 2 or 3).

    Returns a dict of arrays (leading dim = batch) with the reference-shaped OCP data and the
    bookkeeping the state machine needs (joint_sel, vel_sel).  x_guess is already expanded to N+1 rows.
    """
    assert n in (2, 3)
    mdl = Model(n)
    N = N or mdl.N0
    eps = 10 * tol
    q_min, q_max, v_max, tau_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax, mdl.umax
    v_min = -v_max
    rng = _rng(seed)
    R = rng.random((batch, 4 * n + 2))  # fixed column count => batch-size independent streams
    c = 0
    joint_sel = np.minimum((R[:, c] * n).astype(int), n - 1); c += 1
    vel_sel = np.where(R[:, c] < 0.5, -1.0, 1.0); c += 1
    ran = np.empty((batch, n))
    ran[:, 0] = vel_sel * R[:, c]; c += 1
    for j in range(1, n):
        sgn = np.where(R[:, c] < 0.5, -1.0, 1.0); c += 1
        ran[:, j] = sgn * R[:, c]; c += 1
    ran /= np.linalg.norm(ran, axis=1, keepdims=True)
    # place ran1 on the selected joint, keep the others in order (triplependulum_vboc.py:49-54)
    p = np.zeros((batch, n + 1))
    for b_sel in range(n):
        m = joint_sel == b_sel
        others = [j for j in range(n) if j != b_sel]
        p[m, b_sel] = ran[m, 0]
        for t, j in enumerate(others):
            p[m, j] = ran[m, 1 + t]
    q_init = np.empty((batch, n))
    for j in range(n):
        q = q_min + R[:, c] * (q_max - q_min); c += 1
        q = np.where(q > q_max - eps, q - eps, q)
        q = np.where(q < q_min + eps, q + eps, q)
        q_init[:, j] = q
    q_init_sel = np.where(vel_sel < 0, q_min, q_max)
    q_fin_sel = np.where(vel_sel < 0, q_max, q_min)
    rows = np.arange(batch)

    nx = 2 * n + 1
    lbx0 = np.empty((batch, nx)); ubx0 = np.empty((batch, nx))
    lbx0[:, :n] = q_init; ubx0[:, :n] = q_init
    lbx0[:, n:2 * n] = v_min; ubx0[:, n:2 * n] = v_max
    lbx0[:, 2 * n] = dt_sym; ubx0[:, 2 * n] = dt_sym
    fixed = np.where(vel_sel < 0, q_min + eps, q_max - eps)
    lbx0[rows, joint_sel] = fixed; ubx0[rows, joint_sel] = fixed

    # guess: N rows, linspace(0,1,N), then stage N repeats the last row
    tau = np.linspace(0, 1, N)
    xg = np.zeros((batch, N + 1, nx))
    xg[:, :N, :n] = q_init[:, None, :]
    xg[:, :N, 2 * n] = dt_sym
    qs = (1 - tau)[None, :] * q_init_sel[:, None] + tau[None, :] * q_fin_sel[:, None]
    vs = 2 * (1 - tau)[None, :] * (q_fin_sel - q_init_sel)[:, None]
    for b_sel in range(n):
        m = joint_sel == b_sel
        xg[m, :N, b_sel] = qs[m]
        xg[m, :N, n + b_sel] = vs[m]
    xg[:, N] = xg[:, N - 1]
    ug = np.zeros((batch, N, n))
    if n == 2:
        # gravity-compensation torques along the guess (doublependulum_vboc.py:84)
        ug[:, :, 0] = mdl.g * mdl.l1 * (mdl.m1 + mdl.m2) * np.sin(xg[:, :N, 0])
        ug[:, :, 1] = mdl.g * mdl.l2 * mdl.m2 * np.sin(xg[:, :N, 1])

    def rep(v):
        return np.tile(np.asarray(v, dtype=float), (batch, 1))

    out = dict(
        n=n, family="vboc", N=np.full(batch, N, dtype=np.int32), x_guess=xg, u_guess=ug, p=p,
        lbx0=lbx0, ubx0=ubx0,
        lbx=rep([q_min] * n + [v_min] * n + [dt_sym]), ubx=rep([q_max] * n + [v_max] * n + [dt_sym]),
        lbxN=rep([q_min] * n + [0.0] * n + [dt_sym]), ubxN=rep([q_max] * n + [0.0] * n + [dt_sym]),
        lbu=rep([-tau_max] * n), ubu=rep([tau_max] * n),
        joint_sel=joint_sel, vel_sel=vel_sel, q_init_sel=q_init_sel, q_fin_sel=q_fin_sel,
    )
    out["C0"] = stage0_projector(p[:, :n], nx)
    return out


def stage0_projector(d, nx):
    """C[:, n:2n] = I - d d^T  (VBOC/triplependulum_class_vboc.py:174-178): forces v_0 parallel to d."""
    d = np.asarray(d, dtype=float)
    batch, n = d.shape
    C0 = np.zeros((batch, n, nx))
    C0[:, :, n:2 * n] = np.eye(n)[None] - d[:, :, None] * d[:, None, :]
    return C0


def sample_testdata(n, batch, seed, N=None, dt_sym=1e-2):
    """Batched restatement of `testing(v)` sampling: random position, random direction, constant guess."""
    mdl = Model(n)
    N = N or mdl.N0
    q_min, q_max, v_max, tau_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax, mdl.umax
    v_min = -v_max
    rng = _rng(seed)
    R = rng.random((batch, 3 * n))
    nx = 2 * n + 1
    if n == 1:
        # pendulum_testdata.py:14-27: p = [+-1, 0]
        p = np.zeros((batch, 2))
        p[:, 0] = np.where(R[:, 0] < 0.5, -1.0, 1.0)
        q_init = (q_min + R[:, 2] * (q_max - q_min))[:, None]
    else:
        ran = np.where(R[:, :n] < 0.5, -1.0, 1.0) * R[:, n:2 * n]
        ran /= np.linalg.norm(ran, axis=1, keepdims=True)
        p = np.zeros((batch, n + 1))
        p[:, :n] = ran
        q_init = q_min + R[:, 2 * n:3 * n] * (q_max - q_min)
    lbx0 = np.empty((batch, nx)); ubx0 = np.empty((batch, nx))
    lbx0[:, :n] = q_init; ubx0[:, :n] = q_init
    lbx0[:, n:2 * n] = v_min; ubx0[:, n:2 * n] = v_max
    lbx0[:, 2 * n] = dt_sym; ubx0[:, 2 * n] = dt_sym
    xg = np.zeros((batch, N + 1, nx))
    xg[:, :, :n] = q_init[:, None, :]
    xg[:, :, 2 * n] = dt_sym
    ug = np.zeros((batch, N, n))
    if n == 2:
        ug[:, :, 0] = (mdl.g * mdl.l1 * (mdl.m1 + mdl.m2) * np.sin(q_init[:, 0]))[:, None]
        ug[:, :, 1] = (mdl.g * mdl.l2 * mdl.m2 * np.sin(q_init[:, 1]))[:, None]

    def rep(v):
        return np.tile(np.asarray(v, dtype=float), (batch, 1))

    out = dict(
        n=n, family="vboc", N=np.full(batch, N, dtype=np.int32), x_guess=xg, u_guess=ug, p=p,
        lbx0=lbx0, ubx0=ubx0,
        lbx=rep([q_min] * n + [v_min] * n + [dt_sym]), ubx=rep([q_max] * n + [v_max] * n + [dt_sym]),
        lbxN=rep([q_min] * n + [0.0] * n + [dt_sym]), ubxN=rep([q_max] * n + [0.0] * n + [dt_sym]),
        lbu=rep([-tau_max] * n), ubu=rep([tau_max] * n),
    )
    out["C0"] = stage0_projector(p[:, :n], nx) if n > 1 else None
    return out


def sample_al(n, batch, seed, N=100, Tf=1.0):
    """AL feasibility problems (`compute_problem`, AL/triplependulum_class_al.py:148-169): x0 uniform in the
    position box x velocities 5 % beyond the velocity box (AL/triplependulum_al.py:115-123); guess
    x_k = [q0, 0], u = 0."""
    mdl = Model(n)
    q_min, q_max, v_max, tau_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax, mdl.umax
    rng = _rng(seed)
    R = rng.random((batch, 2 * n))
    x0 = np.empty((batch, 2 * n))
    x0[:, :n] = q_min + R[:, :n] * (q_max - q_min)
    vlim = v_max * 1.05
    x0[:, n:] = -vlim + R[:, n:] * 2 * vlim
    return al_problems(n, x0, N=N, Tf=Tf)


def al_problems(n, x0, N=100, Tf=1.0, x_guess=None):
    """Reference-shaped AL OCP data for given initial states x0 (batch, 2n)."""
    mdl = Model(n)
    x0 = np.asarray(x0, dtype=float)
    batch = x0.shape[0]
    q_min, q_max, v_max, tau_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax, mdl.umax
    nx = 2 * n
    if x_guess is None:
        xg = np.zeros((batch, N + 1, nx))
        xg[:, :, :n] = x0[:, None, :n]
    else:
        xg = np.array(x_guess, dtype=float)

    def rep(v):
        return np.tile(np.asarray(v, dtype=float), (batch, 1))

    return dict(
        n=n, family="al", N=np.full(batch, N, dtype=np.int32), Tf=Tf, x_guess=xg,
        u_guess=np.zeros((batch, N, n)), p=None, lbx0=x0.copy(), ubx0=x0.copy(),
        lbx=rep([q_min] * n + [-v_max] * n), ubx=rep([q_max] * n + [v_max] * n),
        lbxN=rep([q_min] * n + [0.0] * n), ubxN=rep([q_max] * n + [0.0] * n),
        lbu=rep([-tau_max] * n), ubu=rep([tau_max] * n), C0=None, x0=x0,
    )


def take(bp, i):
    """Single-problem view (dict) of batched problem data."""
    N = int(bp["N"][i])
    out = dict(n=bp["n"], family=bp["family"], Tf=bp.get("Tf", 1.0))
    out["x_guess"] = np.ascontiguousarray(bp["x_guess"][i, :N + 1])
    out["u_guess"] = np.ascontiguousarray(bp["u_guess"][i, :N])
    for k in ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu", "C0"):
        out[k] = None if bp.get(k) is None else np.ascontiguousarray(bp[k][i])
    return out


def from_ocp_solve_args(n, N, x_sol_guess, u_sol_guess, p, q_lb, q_ub, u_lb, u_ub, q_init_lb, q_init_ub,
                        q_fin_lb, q_fin_ub):
    """Batch-of-one problem data from the arguments of the reference's `OCP_solve`
    (VBOC/triplependulum_class_vboc.py:155-191)."""
    xg, ug = expand_guess(x_sol_guess, u_sol_guess, N)
    one = lambda a: np.ascontiguousarray(np.asarray(a, dtype=float)[None])
    p = np.asarray(p, dtype=float)
    return dict(n=n, family="vboc", N=np.array([N], dtype=np.int32), x_guess=one(xg), u_guess=one(ug), p=one(p),
                lbx0=one(q_init_lb), ubx0=one(q_init_ub), lbx=one(q_lb), ubx=one(q_ub), lbxN=one(q_fin_lb),
                ubxN=one(q_fin_ub), lbu=one(u_lb), ubu=one(u_ub), C0=stage0_projector(p[None, :n], 2 * n + 1))


def pendulum_free_dt_problems(batch, seed, N=50):
    """1-DOF VBOC OCPs with the dt state FREE in [0, 1e-2] and a unit weight on time, in the shape
    `OCP_solve` of VBOC/pendulum_class_vboc.py:107-130 builds them: theta_0 = q_init fixed, theta_N = q_fin and
    dtheta_N = 0 fixed, the velocity sign restricted by the path bounds (VBOC/pendulum_vboc.py:63-87).
    Problem 0 / 1 are the driver's two extreme trajectories (q_max -> q_min and q_min -> q_max); the others
    start and end at seeded random positions in between (the sub-OCPs of the trajectory walk, :171-181)."""
    mdl = Model(1)
    q_min, q_max, v_max, dt = mdl.thetamin, mdl.thetamax, mdl.dthetamax, 1e-2
    rng = _rng(seed)
    R = rng.random((batch, 3))
    xg = np.zeros((batch, N + 1, 3))
    p = np.zeros((batch, 2))
    lbx = np.zeros((batch, 3)); ubx = np.zeros((batch, 3))
    lbx0 = np.zeros((batch, 3)); ubx0 = np.zeros((batch, 3))
    lbxN = np.zeros((batch, 3)); ubxN = np.zeros((batch, 3))
    for b in range(batch):
        down = (b % 2 == 0)  # v_sel = v_min: move from the upper to the lower position limit
        if b < 2:
            q_init, q_fin = (q_max, q_min) if down else (q_min, q_max)
        else:
            a, c = sorted(q_min + R[b, :2] * (q_max - q_min))
            c = max(c, a + 0.05)
            q_init, q_fin = (c, a) if down else (a, c)
        v_sel = -v_max if down else v_max
        lbx[b] = [q_min, -v_max if down else 0.0, 0.0]
        ubx[b] = [q_max, 0.0 if down else v_max, dt]
        p[b] = [1.0 if down else -1.0, 1.0]
        xg[b, :, 0] = np.linspace(q_init, q_fin, N + 1)
        xg[b, :, 1] = v_sel
        xg[b, :, 2] = dt
        lbx0[b] = [q_init, -v_max, 0.0]; ubx0[b] = [q_init, v_max, dt]
        lbxN[b] = [q_fin, 0.0, 0.0]; ubxN[b] = [q_fin, 0.0, dt]
    return dict(n=1, family="vboc", N=np.full(batch, N, dtype=np.int32), x_guess=xg, u_guess=np.zeros((batch, N, 1)),
                p=p, lbx0=lbx0, ubx0=ubx0, lbx=lbx, ubx=ubx, lbxN=lbxN, ubxN=ubxN,
                lbu=np.full((batch, 1), -mdl.umax), ubu=np.full((batch, 1), mdl.umax), C0=None)


# ------------------------------------------------------------------------------------------------
# MPC family: the Safe-MPC OCP with the learned viability margin as a terminal constraint
def mpc_weights(n):
    """Diagonal LINEAR_LS weights of VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:123-137
    (Q = diag(1e4 on positions, 1e-4 on velocities), R = 1e-4) in the engine's z = [u; x] ordering; terminal W_e = Q."""
    Q = np.array([1e4] * n + [1e-4] * n)
    R = np.full(n, 1e-4)
    return np.concatenate([R, Q]), Q.copy()


def sample_mpc(n, batch, seed, N=10, Tf=0.01, v_scale=6.0):
    """Initial states and guesses of the Safe-MPC simulation (VBOC/Safe MPC/parallel/2dof_sym.py:17-28): x0 inside the
    box, reference q_ref = pi (hanging), guess = x0 at every stage with gravity-compensation torques."""
    mdl = Model(n)
    rng = _rng(seed)
    q0 = mdl.thetamin + rng.random((batch, n)) * (mdl.thetamax - mdl.thetamin)
    d = rng.normal(size=(batch, n))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v0 = d * (rng.random((batch, 1)) * v_scale)
    x0 = np.concatenate([q0, v0], axis=1)
    xg = np.repeat(x0[:, None, :], N + 1, axis=1)
    ug = np.repeat(np.stack([mdl.gravity_comp(q) for q in q0])[:, None, :], N, axis=1)
    Wz, WzN = mpc_weights(n)
    yref = np.tile(np.concatenate([np.zeros(n), np.full(n, np.pi), np.zeros(n)]), (batch, 1))   # [u_ref; q_ref; v_ref]
    yrefN = np.tile(np.concatenate([np.full(n, np.pi), np.zeros(n)]), (batch, 1))
    lo = np.concatenate([np.full(n, mdl.thetamin), np.full(n, -mdl.dthetamax)])
    hi = np.concatenate([np.full(n, mdl.thetamax), np.full(n, mdl.dthetamax)])
    rep = lambda v: np.tile(np.asarray(v, dtype=float), (batch, 1))
    W_acados = np.concatenate([WzN, Wz[:n]])      # cost.W diagonal in acados' y = [x; u] order
    yref_acados = np.concatenate([yref[:, n:], yref[:, :n]], axis=1)
    return dict(n=n, family="mpc", N=np.full(batch, N, dtype=np.int32), Tf=Tf, x_guess=xg, u_guess=ug, x0=x0, p=None, C0=None,
                lbx0=x0.copy(), ubx0=x0.copy(), lbx=rep(lo), ubx=rep(hi), lbxN=rep(lo), ubxN=rep(hi),
                lbu=rep([-mdl.umax] * n), ubu=rep([mdl.umax] * n), Wz=Wz, WzN=WzN, yref=yref, yrefN=yrefN,
                W=W_acados, W_e=WzN.copy(), yref_acados=yref_acados, lh=0.0, uh=1e6)
