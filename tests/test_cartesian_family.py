"""SURVEY 8(f)4, nonlinear PATH constraint: the VBOC OCP of the double pendulum with the Cartesian constraint of
`VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:147-160` (the end effector stays outside a circle) as one
hard general row per stage 0..N-1 of the warp solver.  On the CPU the kernel source runs on the host (tools/emu) and is
certified WITHOUT an oracle: acados' SQP exit test recomputed in numpy from the returned iterate and multipliers
(tools/certify.py, row included)."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
import certify  # noqa: E402

CART = (0.0, -1.2, 0.04, 1e6)   # x_c = 0, y_c = -l1 - l2/2, lh = (l2/4)^2, uh (:150-158)


def check_cartesian(bp, base, out):
    """Shared by the emulation and the GPU test."""
    n = 2
    cart = dict(xc=CART[0], yc=CART[1], lh=CART[2], uh=CART[3], rowm=out["rowm"])
    res = certify.kkt_residuals(n, bp, out["x"], out["u"], out["pi"], out["lam"], cart=cart)
    ok = out["status"] == 0
    assert np.isin(out["status"], (0, 2, 4)).all() and ok.mean() > 0.6
    # status 0 is certified by definition, and the engine's residuals are the recomputed ones
    assert certify.passes_exit_test(res)[ok].all()
    eng = np.stack([out[k] for k in ("res_stat", "res_eq", "res_ineq", "res_comp")], axis=1) if "res" not in out else out["res"]
    assert np.abs(res["res_stat"] - eng[:, 0])[ok].max() < 1e-9 and np.abs(res["res_ineq"] - eng[:, 2])[ok].max() < 1e-9
    N = int(bp["N"][0])
    h, _ = certify.cartesian_h(out["x"][:, :N, :2], CART[0], CART[1])
    assert h[ok].min() > CART[2] - 1e-6                       # the returned trajectories stay outside the circle
    hb, _ = certify.cartesian_h(base["x"][:, :N, :2], CART[0], CART[1])
    crossing = (base["status"] == 0) & (hb.min(axis=1) < CART[2] - 1e-3)
    assert crossing.sum() >= 3                                # ... which the unconstrained optima of some problems cross
    active = ok & (out["rowm"][:, :, 0] > 1e-6).any(axis=1)
    assert active.sum() >= 3                                  # row and multiplier exercised
    # where the constraint never binds the solution is the unconstrained one
    free = ok & (base["status"] == 0) & ~active & (hb.min(axis=1) > CART[2] + 1e-3)
    assert free.sum() >= 10 and np.abs(out["cost"] - base["cost"])[free].max() < 1e-3   # both converged to tol_stat 1e-3
    # a start inside the circle cannot be repaired: QP failure, like an infeasible AL problem
    h0, _ = certify.cartesian_h(bp["lbx0"][:, :2], CART[0], CART[1])
    inside = (h0 < CART[2] - 1e-3) & (bp["lbx0"][:, :2] == bp["ubx0"][:, :2]).all(axis=1)
    assert (out["status"][inside] != 0).all()


def test_cartesian_rows_on_the_host_emulation(oracle):
    import emu
    emu.build()
    n, B = 2, 64
    bp = pr.sample_vboc(n, B, seed=5)
    oo = oracle.default_opts(0)
    o = emu.Opts()
    for f, _ in emu.Opts._fields_:
        setattr(o, f, getattr(oo, f))
    base = emu.solve_batch(n, 0, 0, bp, o)
    out = emu.solve_batch(n, 0, 0, bp, o, multipliers=True, cartesian=CART)
    check_cartesian(bp, base, out)


def test_cartesian_h_gradient():
    rng = np.random.default_rng(0)
    q = rng.uniform(2.3, 4.0, (20, 2))
    h, g = certify.cartesian_h(q, 0.0, -1.2)
    for j in range(2):
        e = np.zeros(2)
        e[j] = 1e-6
        fd = (certify.cartesian_h(q + e, 0.0, -1.2)[0] - certify.cartesian_h(q - e, 0.0, -1.2)[0]) / 2e-6
        assert np.abs(fd - g[:, j]).max() < 1e-8
    # l1 = l2 = 0.8: hanging straight down the end effector is 0.4 below the centre
    assert abs(certify.cartesian_h(np.array([np.pi, np.pi]), 0.0, -1.2)[0] - 0.16) < 1e-12


def test_cartesian_class_surface():
    from vboc_b200.shim.Cartesian.doublependulum_class_fixedveldir import OCPdoublependulumINIT, SYMdoublependulumINIT
    ocp = OCPdoublependulumINIT()
    assert ocp.N == 100 and ocp.radius == ocp.l2 / 4 and ocp.x_c == 0 and abs(ocp.y_c + 1.2) < 1e-12
    assert ocp.ocp_solver.cartesian["radius"] == 0.2 and callable(ocp.OCP_solve) and callable(SYMdoublependulumINIT)
