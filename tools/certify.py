"""Solver-independent certificates for the results of the CUDA engine -- numpy / scipy only.

Nothing here imports the oracle or the engine's own solver code: the dynamics are re-stated in numpy (checked
against the golden vectors derived from the reference's model files, tests/test_certify.py), Jacobians come
from the complex step, and the checks are DEFINITIONS, not re-solves:

  kkt_residuals   acados' SQP exit test (ocp_nlp_sqp: res_stat < nlp_solver_tol_stat, res_eq / res_ineq /
                  res_comp < 1e-6, options VBOC/triplependulum_class_vboc.py:129-141) recomputed from the returned
                  iterate (x, u) and the exported multipliers (pi, lam: vboc_download_multipliers).  A status-0
                  result whose residuals pass here IS a point acados' exit test accepts, whoever computed it.
  al_lp_feasible  the AL label (AL/triplependulum_class_al.py:148-169: 1 iff the single SQP_RTI QP solves) is the
                  feasibility of the linearised stopping problem; checked as an LP with HiGHS
                  (scipy.optimize.linprog).
  qp_kkt          K4: KKT conditions of ONE linearised QP at the step the engine returns (RTI step of the VBOC
                  family), dense algebra.

    python tools/certify.py [--c4 4096] [--c23 2048] [--c5 2048] > profiles/r2_certify.md     (on the GPU box)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------ dynamics
def accel(n, q, v, u):
    """Joint accelerations of the reference's n-link pendulums, manipulator form M(q) a = u - c(q, v) - G(q)
    (point masses m at the tips of massless links l, absolute angles; constants of
    VBOC/pendulum_class_vboc.py:14-17, VBOC/triplependulum_class_vboc.py:15-21).  Leading dimensions are
    broadcast; complex arguments are allowed (complex-step differentiation)."""
    if n == 1:
        m, g, d, b = 0.5, 9.81, 0.3, 0.01
        return (m * g * d * np.sin(q) + u - b * v) / (d * d * m)
    m, l, g = 0.4, 0.8, 9.81
    dtype = np.result_type(q, v, u)
    M = np.zeros(q.shape[:-1] + (n, n), dtype=dtype)
    r = np.array(u, dtype=dtype, copy=True)
    for i in range(n):
        r[..., i] -= m * (n - i) * g * l * np.sin(q[..., i])
        for j in range(n):
            mu = m * (n - max(i, j)) * l * l
            dq = q[..., i] - q[..., j]
            M[..., i, j] = mu * np.cos(dq)
            if i != j:
                r[..., i] -= mu * np.sin(dq) * v[..., j] ** 2
    return np.linalg.solve(M, r[..., None])[..., 0]


def rk4(n, x, u, h):
    """One classical RK4 step of xdot = [v; a(q, v, u)] (acados sim_erk, 4 stages, 1 step)."""
    def f(x_):
        return np.concatenate([x_[..., n:], accel(n, x_[..., :n], x_[..., n:], u)], axis=-1)
    h = np.asarray(h)[..., None] if np.ndim(h) else h
    k1 = f(x)
    k2 = f(x + 0.5 * h * k1)
    k3 = f(x + 0.5 * h * k2)
    k4 = f(x + h * k3)
    return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def rk4_jac(n, x, u, h):
    """(x_next, A = d x_next / d x, B = d x_next / d u) by the complex step (exact to rounding)."""
    nx = 2 * n
    eps = 1e-30
    xn = rk4(n, x, u, h)
    A = np.empty(x.shape[:-1] + (nx, nx))
    B = np.empty(x.shape[:-1] + (nx, n))
    for j in range(nx):
        xc = x.astype(complex)
        xc[..., j] += 1j * eps
        A[..., :, j] = rk4(n, xc, u.astype(complex), h).imag / eps
    for j in range(n):
        uc = u.astype(complex)
        uc[..., j] += 1j * eps
        B[..., :, j] = rk4(n, x.astype(complex), uc, h).imag / eps
    return xn, A, B


# ------------------------------------------------------------------------------------------------ KKT exit test
def cartesian_h(q, xc, yc, l1=0.8, l2=0.8):
    """Squared distance of the double pendulum's end effector from (xc, yc) and its gradient wrt (q1, q2):
    VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:154-156.  q (..., 2)."""
    ex = l1 * np.sin(q[..., 0]) + l2 * np.sin(q[..., 1]) - xc
    ey = l1 * np.cos(q[..., 0]) + l2 * np.cos(q[..., 1]) - yc
    g = np.stack([2 * l1 * (ex * np.cos(q[..., 0]) - ey * np.sin(q[..., 0])),
                  2 * l2 * (ex * np.cos(q[..., 1]) - ey * np.sin(q[..., 1]))], axis=-1)
    return ex ** 2 + ey ** 2, g


def kkt_residuals(n, bp, x, u, pi, lam, cart=None):
    """acados' four SQP residuals of the VBOC-family NLP at (x, u, pi, lam), per problem.

    bp: the batched problem dict (reference-shaped, vboc_b200.problems); x (B, Nmax+1, 2n+1), u (B, Nmax, n),
    pi (B, Nmax, 2n), lam (B, Nmax+1, 3n, 2).  The pinned dt state is dropped (its bounds lb == ub absorb any
    gradient).  Multipliers of the other equalities the reference writes as lb == ub pairs / projector rows
    (fixed initial positions, (I - d d') v_0 = 0, v_N = 0) are free in sign, so the stationarity residual is
    measured on the complement of their normals.  Returns dict of (B,) arrays res_stat, res_eq, res_ineq,
    res_comp, lam_min.
    cart = dict(xc, yc, lh, uh, rowm (B, Nmax+1, 6)): the Cartesian path constraint lh <= h(q_k) <= uh at stages 0..N-1
    (n = 2) with its multipliers rowm[..., 0:2] = (lower, upper), Lagrangian terms lam_l (lh - h) + lam_u (h - uh)."""
    B = x.shape[0]
    nx, nz = 2 * n, 3 * n
    Nv = np.asarray(bp["N"])
    out = {k: np.zeros(B) for k in ("res_stat", "res_eq", "res_ineq", "res_comp", "lam_min")}
    for N in np.unique(Nv):
        sel = np.where(Nv == N)[0]
        r = _kkt_fixed_horizon(n, int(N), {k: (np.asarray(v)[sel] if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v)
                                            for k, v in bp.items()}, x[sel], u[sel], pi[sel], lam[sel],
                               None if cart is None else dict(cart, rowm=cart["rowm"][sel]))
        for k in out:
            out[k][sel] = r[k]
    return out


def _kkt_fixed_horizon(n, N, bp, x, u, pi, lam, cart=None):
    nx, nz = 2 * n, 3 * n
    B = x.shape[0]
    h = np.asarray(bp["lbx0"])[:, 2 * n]                      # the pinned dt
    X, U = x[:, :N + 1, :nx], u[:, :N]
    PI, LAM = pi[:, :N], lam[:, :N + 1]
    hh = np.broadcast_to(h[:, None], (B, N))
    xn, A, Bm = rk4_jac(n, X[:, :N], U, hh)
    res_eq = np.abs(xn - X[:, 1:]).reshape(B, -1).max(axis=1)
    # bounds per stage class in z = [u; q; v] ordering
    def zb(key_x, key_u):
        return np.concatenate([np.asarray(bp[key_u]), np.asarray(bp[key_x])[:, :nx]], axis=1)
    lb = np.stack([zb("lbx0", "lbu")] + [zb("lbx", "lbu")] * (N - 1) + [zb("lbxN", "lbu")], axis=1)
    ub = np.stack([zb("ubx0", "ubu")] + [zb("ubx", "ubu")] * (N - 1) + [zb("ubxN", "ubu")], axis=1)
    Z = np.concatenate([np.concatenate([U, np.zeros((B, 1, n))], axis=1), X], axis=2)   # (B, N+1, nz)
    fixed = lb == ub
    fixed[:, N, :n] = False
    exists = np.ones((B, N + 1, nz), dtype=bool)
    exists[:, N, :n] = False                                                            # no u_N
    ineq = exists & ~fixed
    fl, fu = lb - Z, Z - ub
    viol = np.where(ineq, np.maximum(np.maximum(fl, fu), 0.0), 0.0)
    viol = np.maximum(viol, np.where(exists & fixed, np.abs(Z - lb), 0.0))
    res_ineq = viol.reshape(B, -1).max(axis=1)
    ll, lu = LAM[..., 0], LAM[..., 1]
    comp = np.where(ineq, np.maximum(np.abs(ll * fl), np.abs(lu * fu)), 0.0)
    res_comp = comp.reshape(B, -1).max(axis=1)
    lam_min = np.where(ineq, np.minimum(ll, lu), 0.0).reshape(B, -1).min(axis=1)
    # stationarity
    g = np.zeros((B, N + 1, nz))
    g[:, 0, 2 * n:] = np.asarray(bp["p"])[:, :n]                                        # cost = p[:n] . v_0
    BA = np.concatenate([Bm, A], axis=3)                                                # (B, N, nx, nz)
    r = g + np.where(ineq, lu - ll, 0.0)
    r[:, :N] += np.einsum("bkij,bki->bkj", BA, PI)
    r[:, 1:, n:] -= PI
    if cart is not None:
        hv, hg = cartesian_h(X[:, :N, :n], cart["xc"], cart["yc"])                      # (B, N), (B, N, 2)
        l1, l2 = cart["rowm"][:, :N, 0], cart["rowm"][:, :N, 1]
        r[:, :N, n:2 * n] += hg * (l2 - l1)[..., None]
        rl, ru = cart["lh"] - hv, hv - cart["uh"]
        res_ineq = np.maximum(res_ineq, np.maximum(np.maximum(rl, ru), 0.0).max(axis=1))
        res_comp = np.maximum(res_comp, np.maximum(np.abs(l1 * rl), np.abs(l2 * ru)).max(axis=1))
        lam_min = np.minimum(lam_min, np.minimum(l1, l2).min(axis=1))
    r = np.where(exists & ~fixed, r, 0.0)      # fixed components: free multiplier; u_N does not exist
    # stage 0: v_0 = alpha d  ->  only the component of the velocity gradient along d counts
    if bp.get("C0") is not None:
        d = np.asarray(bp["p"])[:, :n]
        d = d / np.linalg.norm(d, axis=1, keepdims=True)
        rv = r[:, 0, 2 * n:]
        r[:, 0, 2 * n:] = d * np.einsum("bi,bi->b", d, rv)[:, None]
        v0 = X[:, 0, n:]
        res_ineq = np.maximum(res_ineq, np.abs(v0 - d * np.einsum("bi,bi->b", d, v0)[:, None]).max(axis=1))
    res_stat = np.abs(r).reshape(B, -1).max(axis=1)
    return dict(res_stat=res_stat, res_eq=res_eq, res_ineq=res_ineq, res_comp=res_comp, lam_min=lam_min)


def passes_exit_test(res, tol_stat=1e-3, tol=1e-6):
    return ((res["res_stat"] < tol_stat) & (res["res_eq"] < tol) & (res["res_ineq"] < tol) & (res["res_comp"] < tol)
            & (res["lam_min"] > -1e-12))


# ------------------------------------------------------------------------------------------------ 1-DOF, free dt
def free_dt_kkt(bp, b, x, u, act_tol=1e-5):
    """Certificate for the 1-DOF VBOC OCP with a FREE dt state (configs[0], VBOC/pendulum_class_vboc.py:60-124,
    VBOC/pendulum_vboc.py:63-130) WITHOUT multipliers from the engine (the lane kernel does not export them): they are
    recovered by bounded least squares from the returned iterate -- free-sign multipliers for the shooting equalities and
    the components fixed by lb == ub, non-negative ones for the bounds that are active to `act_tol` -- so that
    complementarity holds by construction and the remaining stationarity residual decides whether the point is a KKT
    point.  NLP: states (theta, dtheta, dt), control F; x+ = RK4 of dt * [dtheta; a; 0] over a unit step (one step of the
    unscaled model with h = dt, dt+ = dt); cost p[0] dtheta_0 + p[1] sum_{k<N} dt_k (EXTERNAL, :73-75).
    Returns dict(res_stat, res_eq, res_ineq, n_active, lam_min)."""
    from scipy.optimize import lsq_linear
    N = int(bp["N"][b])
    X, U, p = x[:N + 1, :3], u[:N, :1], np.asarray(bp["p"][b], dtype=float)

    def step(xk, uk):           # complex-step safe
        q, v, dt = xk[0:1], xk[1:2], xk[2]
        nxt = rk4(1, np.concatenate([q, v]), uk, dt)
        return np.concatenate([nxt, [dt]])

    nz = 4                      # z_k = [u_k; theta_k; dtheta_k; dt_k], no u_N
    eps = 1e-30
    Jk, gap = [], np.zeros((N, 3))
    for k in range(N):
        z = np.concatenate([U[k], X[k]])
        J = np.zeros((3, nz))
        for j in range(nz):
            zc = z.astype(complex)
            zc[j] += 1j * eps
            J[:, j] = step(zc[1:], zc[:1]).imag / eps
        Jk.append(J)
        gap[k] = step(X[k], U[k]).real - X[k + 1]
    res_eq = np.abs(gap).max()
    lb = np.stack([np.concatenate([bp["lbu"][b], bp["lbx0"][b]])] + [np.concatenate([bp["lbu"][b], bp["lbx"][b]])] * (N - 1)
                  + [np.concatenate([bp["lbu"][b], bp["lbxN"][b]])])
    ub = np.stack([np.concatenate([bp["ubu"][b], bp["ubx0"][b]])] + [np.concatenate([bp["ubu"][b], bp["ubx"][b]])] * (N - 1)
                  + [np.concatenate([bp["ubu"][b], bp["ubxN"][b]])])
    Z = np.concatenate([np.concatenate([U, np.zeros((1, 1))]), X], axis=1)
    exists = np.ones((N + 1, nz), dtype=bool)
    exists[N, 0] = False
    res_ineq = float(np.where(exists, np.maximum(np.maximum(lb - Z, Z - ub), 0.0), 0.0).max())
    g = np.zeros((N + 1, nz))
    g[0, 2] = p[0]
    g[:N, 3] += p[1]
    # unknowns: pi (N x 3, free), then one multiplier per (stage, component, side) that is fixed (free sign) or active (>= 0)
    cols, lo = [], []
    nvar = 3 * N
    idx = lambda k, i: k * nz + i
    rows = (N + 1) * nz
    A = np.zeros((rows, nvar))
    for k in range(N):
        for m in range(3):
            A[idx(k, 0):idx(k, 0) + nz, 3 * k + m] += Jk[k][m]          # + J_k' pi_k on z_k
            A[idx(k + 1, 1 + m), 3 * k + m] -= 1.0                       # - pi_k on x_{k+1}
    extra, extra_lo = [], []
    for k in range(N + 1):
        for i in range(nz):
            if not exists[k, i]:
                continue
            if lb[k, i] == ub[k, i]:
                c = np.zeros(rows); c[idx(k, i)] = 1.0
                extra.append(c); extra_lo.append(-np.inf)
            else:
                if Z[k, i] - lb[k, i] < act_tol:
                    c = np.zeros(rows); c[idx(k, i)] = -1.0              # lam_l (lb - z)
                    extra.append(c); extra_lo.append(0.0)
                if ub[k, i] - Z[k, i] < act_tol:
                    c = np.zeros(rows); c[idx(k, i)] = 1.0               # lam_u (z - ub)
                    extra.append(c); extra_lo.append(0.0)
    Afull = np.concatenate([A, np.stack(extra, axis=1)], axis=1) if extra else A
    keep = exists.reshape(-1)
    lo_b = np.concatenate([np.full(nvar, -np.inf), np.array(extra_lo)])
    sol = lsq_linear(Afull[keep], -g.reshape(-1)[keep], bounds=(lo_b, np.full(lo_b.shape, np.inf)), tol=1e-14, max_iter=2000)
    r = Afull[keep] @ sol.x + g.reshape(-1)[keep]
    lam = sol.x[nvar:][np.array(extra_lo) == 0.0] if extra else np.zeros(0)
    return dict(res_stat=float(np.abs(r).max()), res_eq=float(res_eq), res_ineq=res_ineq, n_active=int((np.array(extra_lo) == 0.0).sum()),
                lam_min=float(lam.min()) if lam.size else 0.0)


# ------------------------------------------------------------------------------------------------ AL labels
def al_lp_feasible(n, x0, N=100, Tf=1.0, x_guess=None, u_guess=None):
    """Feasibility of the QP the AL classes hand to HPIPM in their single SQP_RTI step
    (AL/triplependulum_class_al.py:148-169): dynamics linearised at the guess (x_k = [q0, 0], u_k = 0 unless a
    guess is given), x_0 fixed, box bounds, v_N = 0.  One LP (zero objective) per state through HiGHS.
    Returns (feasible: bool, status message)."""
    import scipy.sparse as sp
    from scipy.optimize import linprog
    from vboc_b200.problems import Model
    mdl = Model(n)
    nx, h = 2 * n, Tf / N
    xg = np.tile(np.concatenate([x0[:n], np.zeros(n)]), (N + 1, 1)) if x_guess is None else np.array(x_guess, dtype=float)
    ug = np.zeros((N, n)) if u_guess is None else np.asarray(u_guess, dtype=float)
    # acados linearises at the ITERATE, i.e. at the guess -- also at stage 0, where `compute_problem` sets the guess
    # [q0, 0] (AL/triplependulum_class_al.py:157-160) while the bounds lbx_0 = ubx_0 pin x_0 = [q0, v0]: the QP's
    # first step is dx_0 = x0 - guess_0, not zero
    dx0 = np.asarray(x0, dtype=float) - xg[0]
    xn, A, Bm = rk4_jac(n, xg[:N], ug, h)
    b = xn - xg[1:]                                             # gap:  dx_{k+1} = A dx_k + B du_k + b
    b[0] = b[0] + A[0] @ dx0
    # variables: dx_1..dx_N (nx each), du_0..du_{N-1} (n each); dx_0 is data
    nvx = N * nx
    rows, cols, vals, rhs = [], [], [], []
    for k in range(N):
        for i in range(nx):
            row = k * nx + i
            rows.append(row), cols.append(k * nx + i), vals.append(1.0)              # dx_{k+1}
            if k > 0:
                for j in range(nx):
                    rows.append(row), cols.append((k - 1) * nx + j), vals.append(-A[k, i, j])
            for j in range(n):
                rows.append(row), cols.append(nvx + k * n + j), vals.append(-Bm[k, i, j])
            rhs.append(b[k, i])
    Aeq = sp.csr_matrix((vals, (rows, cols)), shape=(N * nx, nvx + N * n))
    lo = np.concatenate([np.tile(np.concatenate([np.full(n, mdl.thetamin), np.full(n, -mdl.dthetamax)]), N) - xg[1:].ravel(),
                         np.full(N * n, -mdl.umax) - ug.ravel()])
    hi = np.concatenate([np.tile(np.concatenate([np.full(n, mdl.thetamax), np.full(n, mdl.dthetamax)]), N) - xg[1:].ravel(),
                         np.full(N * n, mdl.umax) - ug.ravel()])
    vN = slice((N - 1) * nx + n, N * nx)
    lo[vN] = hi[vN] = -xg[N, n:]                                # v_N = 0
    res = linprog(np.zeros(nvx + N * n), A_eq=Aeq, b_eq=np.array(rhs), bounds=np.stack([lo, hi], axis=1), method="highs")
    return res.status == 0, res.message


def _lp_worker(args):
    return al_lp_feasible(*args)[0]


def al_lp_labels(n, X, N=100, Tf=1.0, processes=None):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(processes or os.cpu_count()) as pool:
        return np.array(pool.map(_lp_worker, [(n, x, N, Tf) for x in X], chunksize=8))


# ------------------------------------------------------------------------------------------------ K4: one QP
def qp_kkt(n, bp, b, dx, du, pi_q, lam_q, lm=1e-5):
    """KKT residuals of the FIRST linearised QP of VBOC problem b of `bp` at the step (dx, du) with multipliers
    (pi_q, lam_q) -- what one SQP_RTI step of the engine returns (x - x_guess, u - u_guess, exported
    multipliers).  QP (acados EXACT Hessian with exact_hess_dyn = exact_hess_constr = 0 and the Levenberg-Marquardt term,
    VBOC/triplependulum_class_vboc.py:129-141):
        min  g'dz + lm/2 |dz|^2   s.t.  dx_{k+1} = A_k dx_k + B_k du_k + b_k,   lb - z <= dz <= ub - z,
        stage-0 equalities on x_0 + dx_0, terminal equalities on x_N + dx_N.
    The QP is strictly convex, so a point satisfying its KKT conditions is THE solution.  Returns a dict of
    scalar residuals (stationarity on the free subspace, dynamics, bounds, complementarity, min multiplier)."""
    N = int(bp["N"][b])
    sub = {k: (np.asarray(v)[b:b + 1] if isinstance(v, np.ndarray) and v.shape[:1] == (len(bp["N"]),) else v) for k, v in bp.items()}
    nx, nz = 2 * n, 3 * n
    xg, ug = np.asarray(bp["x_guess"])[b, :N + 1, :nx], np.asarray(bp["u_guess"])[b, :N]
    h = float(np.asarray(bp["lbx0"])[b, 2 * n])
    xn, A, Bm = rk4_jac(n, xg[:N], ug, h)
    gap = xn - xg[1:]
    res_b = np.abs(np.einsum("kij,kj->ki", A, dx[:N]) + np.einsum("kij,kj->ki", Bm, du) + gap - dx[1:]).max()
    Zg = np.concatenate([np.concatenate([ug, np.zeros((1, n))]), xg], axis=1)
    DZ = np.concatenate([np.concatenate([du, np.zeros((1, n))]), dx], axis=1)
    def zb(key_x, key_u):
        return np.concatenate([np.asarray(sub[key_u])[0], np.asarray(sub[key_x])[0, :nx]])
    lb = np.stack([zb("lbx0", "lbu")] + [zb("lbx", "lbu")] * (N - 1) + [zb("lbxN", "lbu")])
    ub = np.stack([zb("ubx0", "ubu")] + [zb("ubx", "ubu")] * (N - 1) + [zb("ubxN", "ubu")])
    exists = np.ones((N + 1, nz), dtype=bool)
    exists[N, :n] = False
    fixed = (lb == ub) & exists
    ineq = exists & ~fixed
    Zn = Zg + DZ
    fl, fu = lb - Zn, Zn - ub
    res_d = max(np.where(ineq, np.maximum(np.maximum(fl, fu), 0), 0).max(), np.where(fixed, np.abs(Zn - lb), 0).max())
    ll, lu = lam_q[:N + 1, :, 0], lam_q[:N + 1, :, 1]
    res_m = np.where(ineq, np.maximum(np.abs(ll * fl), np.abs(lu * fu)), 0).max()
    g = np.zeros((N + 1, nz))
    g[0, 2 * n:] = np.asarray(bp["p"])[b, :n]
    r = g + lm * DZ + np.where(ineq, lu - ll, 0.0)
    r[:N] += np.einsum("kij,ki->kj", np.concatenate([Bm, A], axis=2), pi_q[:N])
    r[1:, n:] -= pi_q[:N]
    r = np.where(ineq, r, 0.0)
    if bp.get("C0") is not None:
        d = np.asarray(bp["p"])[b, :n]
        d = d / np.linalg.norm(d)
        r[0, 2 * n:] = d * (d @ r[0, 2 * n:])
        v0 = Zn[0, 2 * n:]
        res_d = max(res_d, np.abs(v0 - d * (d @ v0)).max())
    return dict(res_g=np.abs(r).max(), res_b=res_b, res_d=res_d, res_m=res_m, lam_min=np.where(ineq, np.minimum(ll, lu), 0).min())


# ------------------------------------------------------------------------------------------------ MPC family (8(f)4)
def nn_margin(net, x, n):
    """`nn_decisionfunction` of the reference's Safe-MPC classes
    (VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:232-258) in numpy, complex-step safe:
    h(x) = scale * MLP([(q - mean) / std, v / vn]) - vn,  vn = max(|v|, 1e-3), no ReLU on the output."""
    q, v = x[..., :n], x[..., n:]
    # net["vstart"] (default n): vel_norm = norm_2(x[vstart:]).  The reference writes x[2:] for the double AND the triple
    # pendulum (VBOC/Safe MPC/triplependulum_class_vboc.py:217), which for n = 3 includes theta_3; vstart = 2 mirrors that.
    vs = x[..., int(net.get("vstart", n)):]
    vn = np.sqrt(np.sum(vs * vs, axis=-1))
    vn = np.where(vn.real > 1e-3, vn, 1e-3)
    a = np.concatenate([(q - net["mean"]) / net["std"], v / vn[..., None]], axis=-1)
    a = a @ net["W1"].T + net["b1"]
    a = np.where(a.real > 0, a, 0)
    a = a @ net["W2"].T + net["b2"]
    a = np.where(a.real > 0, a, 0)
    out = a @ net["W3"] + net["b3"]
    return out * net["scale"] - vn


def nn_margin_grad(net, x, n):
    eps = 1e-30
    g = np.empty(x.shape)
    for j in range(x.shape[-1]):
        xc = x.astype(complex)
        xc[..., j] += 1j * eps
        g[..., j] = nn_margin(net, xc, n).imag / eps
    return nn_margin(net, x, n).real, g


def mpc_kkt(n, bp, net, b, x, u, pi, lam, lamg, lm, first_qp_at_guess=False):
    """KKT residuals of MPC problem b of `bp` (problems.sample_mpc).

    first_qp_at_guess=False: the NLP's four acados residuals at the iterate (x, u) with multipliers (pi, lam, lamg) --
    the SQP exit test.  True: (x, u) = guess + the step of ONE SQP_RTI iteration; the residuals are then those of the
    QP linearised at the guess (dynamics, constraint row and Gauss-Newton cost + Levenberg-Marquardt term), whose
    strict convexity makes a KKT point THE solution.  Lagrangian: cost + pi'(Phi - x+) + lam_u (z - ub) + lam_l (lb - z)
    + lamg_u (h - uh) + lamg_l (lh - h)."""
    N = int(bp["N"][b])
    nx, nz, h = 2 * n, 3 * n, bp["Tf"] / int(bp["N"][b])
    X, U = x[:N + 1, :nx], u[:N]
    Wz, WzN, yref, yrefN = bp["Wz"], bp["WzN"], bp["yref"][b], bp["yrefN"][b]
    lb = np.concatenate([bp["lbu"][b], bp["lbx"][b]])
    ub = np.concatenate([bp["ubu"][b], bp["ubx"][b]])
    Z = np.concatenate([np.concatenate([U, np.zeros((1, n))]), X], axis=1)
    exists = np.ones((N + 1, nz), dtype=bool)
    exists[N, :n] = False
    fixed = np.zeros((N + 1, nz), dtype=bool)
    fixed[0, n:] = True                                                # x_0 = x0
    ineq = exists & ~fixed
    ll, lu = lam[:N + 1, :, 0], lam[:N + 1, :, 1]
    if first_qp_at_guess:
        Xg, Ug = bp["x_guess"][b, :N + 1], bp["u_guess"][b, :N]
        Zg = np.concatenate([np.concatenate([Ug, np.zeros((1, n))]), Xg], axis=1)
        xn, A, Bm = rk4_jac(n, Xg[:N], Ug, h)
        DZ = Z - Zg
        res_eq = np.abs(np.einsum("kij,kj->ki", A, DZ[:N, n:]) + np.einsum("kij,kj->ki", Bm, DZ[:N, :n]) + (xn - Xg[1:]) - DZ[1:, n:]).max()
        hv, gc = nn_margin_grad(net, Xg[N], n)
        hlin = hv + gc @ DZ[N, n:]                                     # the row of the QP
        scale = np.concatenate([np.full(N, h), [1.0]])[:, None]
        W = np.concatenate([np.tile(Wz, (N, 1)), np.concatenate([np.zeros(n), WzN])[None]])
        ref = np.concatenate([np.tile(yref, (N, 1)), np.concatenate([np.zeros(n), yrefN])[None]])
        grad = scale * W * (Zg - ref) + (scale * W + lm) * DZ          # Gauss-Newton + LM, at the QP solution
        x0_err = np.abs(Z[0, n:] - bp["x0"][b]).max()
    else:
        xn, A, Bm = rk4_jac(n, X[:N], U, h)
        res_eq = np.abs(xn - X[1:]).max()
        hv, gc = nn_margin_grad(net, X[N], n)
        hlin = hv
        scale = np.concatenate([np.full(N, h), [1.0]])[:, None]
        W = np.concatenate([np.tile(Wz, (N, 1)), np.concatenate([np.zeros(n), WzN])[None]])
        ref = np.concatenate([np.tile(yref, (N, 1)), np.concatenate([np.zeros(n), yrefN])[None]])
        grad = scale * W * (Z - ref)
        x0_err = np.abs(Z[0, n:] - bp["x0"][b]).max()
    fl, fu = lb - Z, Z - ub
    res_ineq = max(np.where(ineq, np.maximum(np.maximum(fl, fu), 0), 0).max(), x0_err, bp["lh"] - hlin, hlin - bp["uh"], 0.0)
    res_comp = max(np.where(ineq, np.maximum(np.abs(ll * fl), np.abs(lu * fu)), 0).max(),
                   abs(lamg[0] * (bp["lh"] - hlin)), abs(lamg[1] * (hlin - bp["uh"])))
    r = grad + np.where(ineq, lu - ll, 0.0)
    r[:N] += np.einsum("kij,ki->kj", np.concatenate([Bm, A], axis=2), pi[:N])
    r[1:, n:] -= pi[:N]
    r[N, n:] += gc * (lamg[1] - lamg[0])
    r = np.where(ineq, r, 0.0)
    return dict(res_stat=np.abs(r).max(), res_eq=res_eq, res_ineq=res_ineq, res_comp=res_comp,
                lam_min=min(np.where(ineq, np.minimum(ll, lu), 0).min(), lamg.min()), h=hlin)


def mpc_rows_kkt(n, bp, net, b, x, u, pi, lam, rowm, rowZ, lm, first_qp_at_guess=False):
    """mpc_kkt for the SOFT-ROW variants of the Safe-MPC classes (VBOC/Safe MPC/{parallel, receiding_hard_constraints,
    soft_traj_constraints}/doublependulum_class_fixedveldir.py:175-199): the margin row at every stage k = 0..N,
        lh <= h(x_k) + sl_k,   h(x_k) - su_k <= uh,   sl_k, su_k >= 0,   cost += 1/2 Zl sl^2 + zl sl + 1/2 Zu su^2 + zu su.
    rowm [N+1][6] = (lam_l, lam_u, lam_sl, lam_su, sl, su) as the engine exports them, rowZ [N+1][4] = (Zl, Zu, zl, zu).
    Lagrangian terms: lam_l (lh - h - sl) + lam_u (h - su - uh) - lam_sl sl - lam_su su."""
    N = int(bp["N"][b])
    nx, nz, h = 2 * n, 3 * n, bp["Tf"] / int(bp["N"][b])
    X, U = x[:N + 1, :nx], u[:N]
    Wz, WzN, yref, yrefN = bp["Wz"], bp["WzN"], bp["yref"][b], bp["yrefN"][b]
    lb = np.concatenate([bp["lbu"][b], bp["lbx"][b]])
    ub = np.concatenate([bp["ubu"][b], bp["ubx"][b]])
    Z = np.concatenate([np.concatenate([U, np.zeros((1, n))]), X], axis=1)
    exists = np.ones((N + 1, nz), dtype=bool)
    exists[N, :n] = False
    fixed = np.zeros((N + 1, nz), dtype=bool)
    fixed[0, n:] = True
    ineq = exists & ~fixed
    ll, lu = lam[:N + 1, :, 0], lam[:N + 1, :, 1]
    rowm, rowZ = rowm[:N + 1], rowZ[:N + 1]
    l1, l2, l3, l4, sl, su = (rowm[:, j] for j in range(6))
    scale = np.concatenate([np.full(N, h), [1.0]])[:, None]
    W = np.concatenate([np.tile(Wz, (N, 1)), np.concatenate([np.zeros(n), WzN])[None]])
    ref = np.concatenate([np.tile(yref, (N, 1)), np.concatenate([np.zeros(n), yrefN])[None]])
    if first_qp_at_guess:
        Xg, Ug = bp["x_guess"][b, :N + 1], bp["u_guess"][b, :N]
        Zg = np.concatenate([np.concatenate([Ug, np.zeros((1, n))]), Xg], axis=1)
        xn, A, Bm = rk4_jac(n, Xg[:N], Ug, h)
        DZ = Z - Zg
        res_eq = np.abs(np.einsum("kij,kj->ki", A, DZ[:N, n:]) + np.einsum("kij,kj->ki", Bm, DZ[:N, :n]) + (xn - Xg[1:]) - DZ[1:, n:]).max()
        hg = [nn_margin_grad(net, Xg[k], n) for k in range(N + 1)]
        gc = np.stack([g_[1] for g_ in hg])
        hlin = np.array([g_[0] for g_ in hg]) + np.einsum("ki,ki->k", gc, DZ[:, n:])
        grad = scale * W * (Zg - ref) + (scale * W + lm) * DZ
    else:
        xn, A, Bm = rk4_jac(n, X[:N], U, h)
        res_eq = np.abs(xn - X[1:]).max()
        hg = [nn_margin_grad(net, X[k], n) for k in range(N + 1)]
        gc = np.stack([g_[1] for g_ in hg])
        hlin = np.array([g_[0] for g_ in hg])
        grad = scale * W * (Z - ref)
    x0_err = np.abs(Z[0, n:] - bp["x0"][b]).max()
    fl, fu = lb - Z, Z - ub
    rl, ru = bp["lh"] - hlin - sl, hlin - su - bp["uh"]
    res_ineq = max(np.where(ineq, np.maximum(np.maximum(fl, fu), 0), 0).max(), x0_err, rl.max(), ru.max(), (-sl).max(), (-su).max(), 0.0)
    res_comp = max(np.where(ineq, np.maximum(np.abs(ll * fl), np.abs(lu * fu)), 0).max(), np.abs(l1 * rl).max(),
                   np.abs(l2 * ru).max(), np.abs(l3 * sl).max(), np.abs(l4 * su).max())
    r = grad + np.where(ineq, lu - ll, 0.0)
    r[:N] += np.einsum("kij,ki->kj", np.concatenate([Bm, A], axis=2), pi[:N])
    r[1:, n:] -= pi[:N]
    r[:, n:] += gc * (l2 - l1)[:, None]
    r = np.where(ineq, r, 0.0)
    rs = np.stack([rowZ[:, 0] * sl + rowZ[:, 2] - l1 - l3, rowZ[:, 1] * su + rowZ[:, 3] - l2 - l4])
    return dict(res_stat=max(np.abs(r).max(), np.abs(rs).max()), res_eq=res_eq, res_ineq=res_ineq, res_comp=res_comp,
                lam_min=min(np.where(ineq, np.minimum(ll, lu), 0).min(), rowm[:, :4].min()), h=hlin, sl=sl)


# ------------------------------------------------------------------------------------------------ report
def _solve_with_multipliers(n, family, bp, mode, opts=None):
    from vboc_b200 import engine
    B = len(bp["N"])
    sol = engine.BatchSolver(n, family, B, int(np.asarray(bp["x_guess"]).shape[1] - 1))
    if opts is not None:
        sol.set_opts(opts)
    sol.export_multipliers(True)
    out = sol.solve(bp, mode)
    out["pi"], out["lam"] = sol.multipliers()
    sol.close()
    return out


def certify_vboc(n, bp, tag, log):
    """Solve on the GPU, recompute the exit test for every status-0 result, classify the others."""
    out = _solve_with_multipliers(n, "vboc", bp, 0)
    res = kkt_residuals(n, bp, out["x"], out["u"], out["pi"], out["lam"])
    ok = out["status"] == 0
    cert = passes_exit_test(res)
    B = len(ok)
    log(f"## {tag}: {B} problems, {n}-DOF\n")
    log(f"* status counts: {dict(zip(*[a.tolist() for a in np.unique(out['status'], return_counts=True)]))}")
    log(f"* status 0 and exit test re-computed in numpy passes: **{int((ok & cert).sum())} / {int(ok.sum())}**"
        f" (violations: {int((ok & ~cert).sum())})")
    for k in ("res_stat", "res_eq", "res_ineq", "res_comp"):
        v = res[k][ok]
        if v.size:
            log(f"* {k} (numpy) over status-0 results: p50 {np.percentile(v, 50):.2e}, p99 {np.percentile(v, 99):.2e}, max {v.max():.2e};"
                f" max |numpy - engine| = {np.abs(res[k][ok] - out[k][ok]).max():.2e}")
    log(f"* smallest bound multiplier over status-0 results: {res['lam_min'][ok].min() if ok.any() else 0:.2e}")
    bad = np.where(~ok)[0]
    # status != 0 yet the returned point passes the exit test would be a wrongly reported failure
    log(f"* status != 0: {bad.size}; of these the returned point nevertheless passes the exit test: {int(cert[bad].sum())}\n")
    return out, res, bad


def existence_for_failures(n, bp, out, bad, log, max_list=40):
    """For every problem the reference-settings SQP did not converge on: does a KKT point exist at all?  The
    same problem is re-solved with a heavier Levenberg-Marquardt term (a different algorithm as far as the path
    is concerned) and the point it returns is certified with kkt_residuals -- an existence proof that does not
    depend on who found the point."""
    if bad.size == 0:
        return
    from vboc_b200 import engine
    sub = {k: (np.asarray(v)[bad] if isinstance(v, np.ndarray) and v.shape[:1] == (len(bp["N"]),) else v) for k, v in bp.items()}
    found = np.zeros(bad.size, dtype=bool)
    cost = np.full(bad.size, np.nan)
    how = np.array([""] * bad.size, dtype=object)
    for lm in (1e-4, 1e-3, 1e-2):
        o = engine.default_opts("vboc")
        o.levenberg_marquardt = lm
        alt = _solve_with_multipliers(n, "vboc", sub, 0, o)
        res = kkt_residuals(n, sub, alt["x"], alt["u"], alt["pi"], alt["lam"])
        good = (alt["status"] == 0) & passes_exit_test(res) & ~found
        found |= good
        cost[good] = alt["cost"][good]
        how[good] = f"LM {lm:g}"
    log(f"Existence of a KKT point for the {bad.size} problems that end with status != 0 at the reference settings "
        f"(re-solved with a larger Levenberg-Marquardt term, point certified in numpy): **{int(found.sum())} / {bad.size}** "
        "have a certified KKT point, i.e. a converged answer exists and acados may or may not reach it.\n")
    log("| problem | status (sqp iters) | res_stat at exit | cost at exit | certified KKT point found with | its cost |")
    log("|---|---|---|---|---|---|")
    for j, b in enumerate(bad[:max_list]):
        log(f"| {b} | {out['status'][b]} ({out['sqp_iter'][b]}) | {out['res_stat'][b]:.2e} | {out['cost'][b]:.6f} | "
            f"{how[j] or 'none'} | {cost[j]:.6f} |")
    log("")


def main():
    import argparse
    from vboc_b200 import problems as pr
    from vboc_b200._lib import MODE_RTI
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4", type=int, default=4096)
    ap.add_argument("--c23", type=int, default=2048)
    ap.add_argument("--c5", type=int, default=2048)
    a = ap.parse_args()
    log = print
    log("# Round 2 - solver-independent certificates of the CUDA engine's results (`tools/certify.py`)\n")
    log("numpy / scipy only: the exit test of acados' SQP is recomputed from the returned iterate and the exported "
        "multipliers with a numpy restatement of the dynamics (complex-step Jacobians); AL labels are compared with "
        "LP feasibility (HiGHS) of the linearised problem.  Neither the oracle nor the engine's solver code is used.\n")
    for tag, n, bp in (("C4 `triplependulum_vboc.py` sampling", 3, pr.sample_vboc(3, a.c4, seed=4242)),
                       ("C3 `doublependulum_vboc.py` sampling", 2, pr.sample_vboc(2, a.c23, seed=4243)),
                       ("C2 `doublependulum_testdata.py` sampling", 2, pr.sample_testdata(2, a.c23, seed=4244))):
        out, res, bad = certify_vboc(n, bp, tag, log)
        existence_for_failures(n, bp, out, bad, log)
    # C5: AL labels
    n = 3
    bp = pr.sample_al(n, a.c5, seed=4245)
    out = _solve_with_multipliers(n, "al", bp, MODE_RTI)
    lab = np.where(out["status"] == 0, 1, np.where(out["status"] == 4, 0, 2))
    X0 = np.asarray(bp["lbx0"])[:, :2 * n]
    feas = al_lp_labels(n, X0)
    decided = out["qp_status"] != 1
    agree = (lab == 1) == feas
    log(f"## C5 `triplependulum_al.py`: {a.c5} states, label vs LP feasibility (HiGHS) of the linearised problem\n")
    log(f"* engine: viable {int((lab == 1).sum())}, unviable {int((lab == 0).sum())}, other {int((lab == 2).sum())}; "
        f"LP feasible {int(feas.sum())}")
    log(f"* QP decided by the IPM (converged -> label 1, minimum step -> status 4 -> label 0): {int(decided.sum())} states, "
        f"agreement with LP feasibility **{agree[decided].mean() * 100:.3f} %** ({int((~agree[decided]).sum())} disagreements)")
    log(f"* QP NOT decided (the IPM used up its 50 iterations; acados tolerates the QP solver's max-iter status, so "
        f"`compute_problem` returns 1): {int((~decided).sum())} states, of which LP-feasible {int(feas[~decided].sum())}. "
        "These labels are the reference's semantics, not a feasibility statement.\n")
    if (~agree[decided]).any() or (~decided).any():
        log("| state | engine label (qp status, ipm iterations) | LP feasible | explanation |")
        log("|---|---|---|---|")
        for b in np.where(~agree | ~decided)[0]:
            why = ("iteration limit of the QP solver reached: label 1 by acados' tolerance of max-iter, feasibility undecided"
                   if not decided[b] else
                   "QP at the edge of feasibility: a feasible set of (near) zero width has no interior for the IPM")
            log(f"| {b} | {lab[b]} ({out['qp_status'][b]}, {out['qp_iter'][b]}) | {bool(feas[b])} | {why} |")
        log("")


if __name__ == "__main__":
    main()
