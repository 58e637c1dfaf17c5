"""ctypes wrapper of tools/emu/libemu.so (host emulation of the warp solver; dev/test harness only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Opts(C.Structure):
    _fields_ = [
        ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
        ("max_iter", C.c_int), ("levenberg_marquardt", C.c_double),
        ("alpha_min", C.c_double), ("alpha_reduction", C.c_double), ("globalization", C.c_int),
        ("qp_tol_stat", C.c_double), ("qp_tol_eq", C.c_double), ("qp_tol_ineq", C.c_double), ("qp_tol_comp", C.c_double),
        ("qp_iter_max", C.c_int),
        ("qp_mu0", C.c_double), ("qp_alpha_min", C.c_double), ("qp_reg_prim", C.c_double),
        ("qp_lam_min", C.c_double), ("qp_t_min", C.c_double), ("qp_tau_min", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("sqp_iter", C.c_int), ("qp_iter", C.c_int), ("ls_evals", C.c_int),
        ("qp_status", C.c_int), ("pad_", C.c_int), ("cost", C.c_double),
        ("res_stat", C.c_double), ("res_eq", C.c_double), ("res_ineq", C.c_double), ("res_comp", C.c_double),
    ]


def build():
    so = os.path.join(_HERE, "libemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-o", so, os.path.join(_HERE, "emu.cpp")])
    return so


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def solve_batch(n, family, mode, bp, opts, kernel="warp", multipliers=False, guess_net=None, cartesian=None):
    """guess_net (AL family): dict(W1, b1, W2, b2, W3, b3, mean, std) -> the guess network is evaluated by the solver
    itself (compute_problem_nnguess); the result then carries the computed guesses as `x_guess`.
    cartesian (VBOC family, n = 2): (xc, yc, lh, uh) -> the Cartesian path constraint at stages 0..N-1; the result then
    carries `rowm` (B, Nmax+1, 6), the row multipliers in columns 0:2."""
    lib = C.CDLL(build() if not os.path.exists(os.path.join(_HERE, "libemu.so")) else os.path.join(_HERE, "libemu.so"))
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    xg, ug = c(bp["x_guess"]), c(bp["u_guess"])
    B, Np1, nxr = xg.shape
    Nmax = Np1 - 1
    Nv = np.ascontiguousarray(bp["N"], dtype=np.int32)
    keep = [c(bp.get(k)) for k in ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu")]
    if family == 0:
        h = np.ascontiguousarray(bp["lbx0"][:, 2 * n])
        d = None if bp.get("C0") is None else c(bp["p"][:, :n] / np.linalg.norm(bp["p"][:, :n], axis=1, keepdims=True))
    else:
        h = np.full(B, bp.get("Tf", 1.0) / Nv, dtype=np.float64)
        d = None
    x, u = np.zeros_like(xg), np.zeros_like(ug)
    st = (Stats * B)()
    pi = lam = None
    if multipliers:
        assert kernel == "warp"
        pi, lam = np.zeros((B, Nmax, 2 * n)), np.zeros((B, Nmax + 1, 3 * n, 2))
        lib.emu_set_multiplier_out(_p(pi), _p(lam))
    rowm = None
    if cartesian is not None:
        assert kernel == "warp" and family == 0 and n == 2
        rowm = np.zeros((B, Nmax + 1, 6))
        lib.emu_set_cartesian(1, *[C.c_double(float(v)) for v in cartesian], _p(rowm))
    xg_out = None
    if guess_net is not None:
        assert kernel == "warp" and family == 1
        gw = [c(guess_net["W1"]), c(guess_net["b1"]), c(np.asarray(guess_net["W2"]).T), c(guess_net["b2"]),
              c(np.asarray(guess_net["W3"]).T), c(guess_net["b3"])]
        xg_out = np.zeros_like(xg)
        lib.emu_set_guess_net(1, gw[1].shape[0], gw[5].shape[0], *[_p(a) for a in gw], C.c_double(guess_net["mean"]),
                              C.c_double(guess_net["std"]), _p(xg_out))
    fn = {"warp": lib.emu_solve_batch, "lane": lib.emu_lane_solve_batch, "lane_dts": lib.emu_lane_dts_solve_batch}[kernel]
    fn(n, family, mode, B, Nmax, Nv.ctypes.data_as(C.POINTER(C.c_int)), _p(xg), _p(ug),
                        *[_p(a) for a in keep], _p(d), _p(h), C.byref(opts), _p(x), _p(u), st)
    if multipliers:
        lib.emu_set_multiplier_out(None, None)
    if cartesian is not None:
        lib.emu_set_cartesian(0, C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), None)
    if guess_net is not None:
        lib.emu_set_guess_net(0, 0, 0, None, None, None, None, None, None, C.c_double(0.0), C.c_double(1.0), None)
    f = lambda name: np.array([getattr(s_, name) for s_ in st])
    return dict(rowm=rowm, x_guess=xg_out, pi=pi, lam=lam, status=f("status"), x=x, u=u, cost=f("cost"), sqp_iter=f("sqp_iter"), qp_iter=f("qp_iter"),
                ls_evals=f("ls_evals"), qp_status=f("qp_status"),
                res=np.stack([f("res_stat"), f("res_eq"), f("res_ineq"), f("res_comp")], axis=1))


class DgStats(C.Structure):
    _fields_ = [("status", C.c_int), ("n_rows", C.c_int), ("solves", C.c_int), ("converged", C.c_int),
                ("sim_steps", C.c_int), ("sqp_iter", C.c_int), ("qp_iter", C.c_int), ("t_done_us", C.c_int)]


ROWS_MAX = 258


def datagen_run(n, inp, opts, N0=100, dt=1e-2, tol=1e-3):
    """The device state machine (vboc_b200/csrc/datagen_warp.h) on the host.  inp: drivers.dg_inputs(...).
    Returns (list of per-problem row arrays or None, stats list)."""
    lib = C.CDLL(os.path.join(_HERE, "libemu.so"))
    B = len(inp["joint_sel"])
    rows = np.zeros((B, ROWS_MAX, 2 * n))
    st = (DgStats * B)()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib.emu_datagen_run(n, B, N0, C.c_double(dt), C.c_double(tol), ip(inp["joint_sel"]), _p(inp["p"]), _p(inp["lb0"]),
                             _p(inp["ub0"]), _p(inp["retry"]), C.byref(opts), _p(rows), st)
    assert rc == 0
    out = [rows[b, :st[b].n_rows].copy() if st[b].status != 1 else None for b in range(B)]
    return out, [dict((f, getattr(s_, f)) for f, _ in DgStats._fields_) for s_ in st]


def solve_mpc(n, mode, bp, net, opts, multipliers=False, rowZ=None):
    """MPC family on the host emulation.  bp: problems.sample_mpc(...); net: dict(W1, b1, W2, b2, W3 (H,), b3, mean, std,
    scale) in float64.  rowZ (B, N+1, 4) = per-stage (Zl, Zu, zl, zu): the margin row at every stage, softened (the
    parallel / receding / soft_traj variants); None: the hard terminal row."""
    lib = C.CDLL(os.path.join(_HERE, "libemu.so"))
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    xg, ug = c(bp["x_guess"]), c(bp["u_guess"])
    B, Np1, nx = xg.shape
    Nmax = Np1 - 1
    Nv = np.ascontiguousarray(bp["N"], dtype=np.int32)
    x, u = np.zeros_like(xg), np.zeros_like(ug)
    st = (Stats * B)()
    lamg = np.zeros((B, 2))
    rowm = np.zeros((B, Nmax + 1, 6))
    if rowZ is not None:
        rowZ = c(rowZ)
        assert rowZ.shape == (B, Nmax + 1, 4)
    pi = lam = None
    if multipliers:
        pi, lam = np.zeros((B, Nmax, 2 * n)), np.zeros((B, Nmax + 1, 3 * n, 2))
        lib.emu_set_multiplier_out(_p(pi), _p(lam))
    keep = [c(bp["x0"])] + [c(bp[k][0]) for k in ("lbx", "ubx", "lbu", "ubu")] + [c(bp[k]) for k in ("Wz", "WzN", "yref", "yrefN")]
    w = [c(net[k]) for k in ("W1", "b1", "W2", "b2", "W3")]
    rc = lib.emu_solve_mpc(n, mode, B, Nmax, Nv.ctypes.data_as(C.POINTER(C.c_int)), _p(xg), _p(ug), *[_p(a) for a in keep],
                           C.c_double(bp["Tf"]), int(w[0].shape[0]), *[_p(a) for a in w], C.c_double(float(net["b3"])),
                           C.c_double(net["mean"]), C.c_double(net["std"]), C.c_double(net["scale"]), C.c_double(bp["lh"]),
                           C.c_double(bp["uh"]), C.byref(opts), _p(x), _p(u), st, _p(lamg), int(rowZ is not None),
                           _p(rowZ) if rowZ is not None else None, _p(rowm), int(net.get("vstart", -1)))
    assert rc == 0
    if multipliers:
        lib.emu_set_multiplier_out(None, None)
    f = lambda name: np.array([getattr(s_, name) for s_ in st])
    return dict(pi=pi, lam=lam, lamg=lamg, rowm=rowm, status=f("status"), x=x, u=u, cost=f("cost"), sqp_iter=f("sqp_iter"),
                qp_iter=f("qp_iter"), qp_status=f("qp_status"),
                res=np.stack([f("res_stat"), f("res_eq"), f("res_ineq"), f("res_comp")], axis=1))


def testdata_run(n, inp, opts, N0=100, dt=1e-2, max_solves=60):
    """The device test-data state machine (DataGen::run_testing) on the host.  Returns (rows (B, 2n), stats list)."""
    lib = C.CDLL(os.path.join(_HERE, "libemu.so"))
    B = len(inp["ran"])
    rows = np.zeros((B, 2 * n))
    st = (DgStats * B)()
    rc = lib.emu_testdata_run(n, B, N0, C.c_double(dt), max_solves, _p(np.ascontiguousarray(inp["ran"])),
                              _p(np.ascontiguousarray(inp["q_init"])), _p(np.ascontiguousarray(inp["retry"])), C.byref(opts),
                              _p(rows), st)
    assert rc == 0
    return rows, [dict((f, getattr(s_, f)) for f, _ in DgStats._fields_) for s_ in st]
