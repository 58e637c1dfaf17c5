"""ctypes wrapper around oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  The product package vboc_b200 never does.  PARITY UNPINNED for the solver part
(see vboc_oracle.h); the dynamics are pinned by tests/golden/dynamics_golden.npz.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAMILY_VBOC, FAMILY_AL = 0, 1
MODE_SQP, MODE_RTI = 0, 1


class OrcOpts(C.Structure):
    _fields_ = [
        ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
        ("max_iter", C.c_int), ("levenberg_marquardt", C.c_double),
        ("alpha_min", C.c_double), ("alpha_reduction", C.c_double), ("globalization", C.c_int),
        ("qp_tol_stat", C.c_double), ("qp_tol_eq", C.c_double), ("qp_tol_ineq", C.c_double), ("qp_tol_comp", C.c_double),
        ("qp_iter_max", C.c_int),
        ("qp_mu0", C.c_double), ("qp_alpha_min", C.c_double), ("qp_reg_prim", C.c_double),
        ("qp_lam_min", C.c_double), ("qp_t_min", C.c_double), ("qp_tau_min", C.c_double),
        ("eliminate_dt", C.c_int),
    ]


class OrcStats(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("sqp_iter", C.c_int), ("qp_iter", C.c_int), ("ls_evals", C.c_int),
        ("qp_status", C.c_int), ("cost", C.c_double),
        ("res_stat", C.c_double), ("res_eq", C.c_double), ("res_ineq", C.c_double), ("res_comp", C.c_double),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "vboc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_solve.restype = C.c_int
        _LIB.orc_solve_batch.restype = C.c_int
        _LIB.orc_first_qp.restype = C.c_int
    return _LIB


def use_native_baseline_build():
    """Switch this module to oracle/_native/liboracle_n3.so, built here and now with -O3 -march=native and the 3-DOF
    dimensions fixed at compile time (the CPU-baseline build; bench.py only).  Results follow the same algorithm
    but are not bit-identical to liboracle.so (FMA contraction), so the parity tests never use it."""
    global _LIB
    so = os.path.join(_HERE, "_native", "liboracle_n3.so")
    subprocess.check_call(["make", "-C", _HERE, "-B", "_native/liboracle_n3.so"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    _LIB = C.CDLL(so)
    _LIB.orc_solve.restype = C.c_int
    _LIB.orc_solve_batch.restype = C.c_int
    _LIB.orc_first_qp.restype = C.c_int
    return so


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def default_opts(family):
    o = OrcOpts()
    lib().orc_default_opts(C.c_int(family), C.byref(o))
    return o


def f(n, family, x, u):
    nx = 2 * n + (family == FAMILY_VBOC)
    out = np.empty(nx)
    lib().orc_f(n, family, _p(_c(x)), _p(_c(u)), _p(out))
    return out


def f_jac(n, family, x, u):
    nx = 2 * n + (family == FAMILY_VBOC)
    jx, ju = np.empty((nx, nx)), np.empty((nx, n))
    lib().orc_f_jac(n, family, _p(_c(x)), _p(_c(u)), _p(jx), _p(ju))
    return jx, ju


def rk4(n, family, x, u, h=1e-2, jac=False):
    nx = 2 * n + (family == FAMILY_VBOC)
    xn = np.empty(nx)
    if jac:
        A, B = np.empty((nx, nx)), np.empty((nx, n))
        lib().orc_rk4(n, family, _p(_c(x)), _p(_c(u)), C.c_double(h), _p(xn), _p(A), _p(B))
        return xn, A, B
    lib().orc_rk4(n, family, _p(_c(x)), _p(_c(u)), C.c_double(h), _p(xn), None, None)
    return xn


def solve(n, family, mode, prob, opts=None):
    """prob: dict with x_guess (N+1,nx), u_guess (N,nu), p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu,
    optional C0 (ng,nx), Tf.  Returns dict(status, x, u, pi, cost, stats...)."""
    opts = opts or default_opts(family)
    xg, ug = _c(prob["x_guess"]), _c(prob["u_guess"])
    N = ug.shape[0]
    nx = xg.shape[1]
    assert xg.shape[0] == N + 1
    x, u, pi = np.empty((N + 1, nx)), np.empty((N, n)), np.empty((N, nx))
    st = OrcStats()
    C0 = _c(prob.get("C0"))
    ng = 0 if C0 is None else C0.shape[0]
    keep = [_c(prob.get(k)) for k in ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu")]
    status = lib().orc_solve(n, family, mode, N, _p(xg), _p(ug), *[_p(a) for a in keep], _p(C0), ng,
                             C.c_double(prob.get("Tf", 1.0)), C.byref(opts), _p(x), _p(u), _p(pi), C.byref(st))
    return dict(status=status, x=x, u=u, pi=pi, cost=st.cost, sqp_iter=st.sqp_iter, qp_iter=st.qp_iter,
                ls_evals=st.ls_evals, qp_status=st.qp_status, res=(st.res_stat, st.res_eq, st.res_ineq, st.res_comp))


def solve_batch(n, family, mode, bp, opts=None, nthreads=0):
    """bp: batched dict: N (B,), x_guess (B,Nmax+1,nx), u_guess (B,Nmax,nu), p (B,n+1), bounds (B,nx)/(B,nu),
    C0 (B,ng,nx) or None."""
    opts = opts or default_opts(family)
    xg, ug = _c(bp["x_guess"]), _c(bp["u_guess"])
    B, Np1, nx = xg.shape
    Nmax = Np1 - 1
    Nv = np.ascontiguousarray(bp["N"], dtype=np.int32)
    x, u = np.zeros_like(xg), np.zeros_like(ug)
    stats = (OrcStats * B)()
    C0 = _c(bp.get("C0"))
    ng = 0 if C0 is None else C0.shape[1]
    keep = [_c(bp.get(k)) for k in ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu")]
    lib().orc_solve_batch(n, family, mode, B, Nmax, Nv.ctypes.data_as(C.POINTER(C.c_int)), _p(xg), _p(ug),
                          *[_p(a) for a in keep], _p(C0), ng, C.c_double(bp.get("Tf", 1.0)), C.byref(opts),
                          _p(x), _p(u), stats, nthreads)
    return dict(
        status=np.array([s.status for s in stats]), x=x, u=u,
        cost=np.array([s.cost for s in stats]), sqp_iter=np.array([s.sqp_iter for s in stats]),
        qp_iter=np.array([s.qp_iter for s in stats]), ls_evals=np.array([s.ls_evals for s in stats]),
        qp_status=np.array([s.qp_status for s in stats]),
        res=np.array([[s.res_stat, s.res_eq, s.res_ineq, s.res_comp] for s in stats]))


def first_qp(n, family, prob, opts=None):
    opts = opts or default_opts(family)
    xg, ug = _c(prob["x_guess"]), _c(prob["u_guess"])
    N = ug.shape[0]
    nxr = xg.shape[1]
    # internal nx: dt eliminated when pinned
    nx = 2 * n + (1 if (family == FAMILY_VBOC and not _dt_pinned(prob, n, opts)) else 0)
    A, B, b = np.empty((N, nx, nx)), np.empty((N, nx, n)), np.empty((N, nx))
    dx, du = np.empty((N + 1, nx)), np.empty((N, n))
    it = C.c_int(0)
    C0 = _c(prob.get("C0"))
    ng = 0 if C0 is None else C0.shape[0]
    keep = [_c(prob.get(k)) for k in ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu")]
    st = lib().orc_first_qp(n, family, N, _p(xg), _p(ug), *[_p(a) for a in keep], _p(C0), ng,
                            C.c_double(prob.get("Tf", 1.0)), C.byref(opts), _p(A), _p(B), _p(b), _p(dx), _p(du),
                            C.byref(it))
    assert nxr >= nx
    return dict(status=st, A=A, B=B, b=b, dx=dx, du=du, iters=it.value)


def _dt_pinned(prob, n, opts):
    if not opts.eliminate_dt:
        return False
    v = prob["lbx0"][2 * n]
    for k in ("ubx0", "lbx", "ubx", "lbxN", "ubxN"):
        if prob[k][2 * n] != v:
            return False
    return bool(np.all(np.asarray(prob["x_guess"])[:, 2 * n] == v))
