"""CPU checks of the solver-independent certificates (tools/certify.py): its numpy dynamics against the golden
vectors derived from the reference's own model files, its KKT exit test on the results of the kernel SOURCE run
through the host emulation (multipliers exported exactly as vboc_download_multipliers does), the K4 QP check and
the AL label <-> LP feasibility equivalence on a few states."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
import certify  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dynamics_golden.npz"))


@pytest.fixture(scope="module")
def emu():
    import emu as e
    e.build()
    return e


def _opts(emu, oo, **kw):
    o = emu.Opts()
    for f, _ in emu.Opts._fields_:
        setattr(o, f, getattr(oo, f))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


@pytest.mark.parametrize("n", [1, 2, 3])
def test_numpy_dynamics_match_reference_golden(n):
    x, u = GOLD[f"al{n}_x"], GOLD[f"al{n}_u"]
    f = np.concatenate([x[:, n:], certify.accel(n, x[:, :n], x[:, n:], u)], axis=1)
    assert np.abs(f - GOLD[f"al{n}_f"]).max() < 1e-10 * max(1.0, np.abs(GOLD[f"al{n}_f"]).max())
    assert np.abs(certify.rk4(n, x, u, 1e-2) - GOLD[f"al{n}_rk4"]).max() < 1e-11
    # complex-step Jacobian of the RK4 map against the golden continuous Jacobian in the limit h -> 0
    h = 1e-6
    _, A, B = certify.rk4_jac(n, x, u, h)
    assert np.abs((A - np.eye(2 * n)) / h - GOLD[f"al{n}_jx"]).max() < 1e-3 * max(1.0, np.abs(GOLD[f"al{n}_jx"]).max())
    assert np.abs(B / h - GOLD[f"al{n}_ju"]).max() < 1e-3 * max(1.0, np.abs(GOLD[f"al{n}_ju"]).max())
    # the dt-scaled VBOC model over a unit step is the same map with h = dt
    xv, uv = GOLD[f"vboc{n}_x"], GOLD[f"vboc{n}_u"]
    got = certify.rk4(n, xv[:, :2 * n], uv, xv[:, 2 * n])
    assert np.abs(got - GOLD[f"vboc{n}_rk4"][:, :2 * n]).max() < 1e-11


@pytest.mark.parametrize("n", [2, 3])
def test_exit_test_recomputed_from_exported_multipliers(oracle, emu, n):
    bp = pr.sample_vboc(n, 8, seed=21)
    out = emu.solve_batch(n, 0, 0, bp, _opts(emu, oracle.default_opts(0)), multipliers=True)
    res = certify.kkt_residuals(n, bp, out["x"], out["u"], out["pi"], out["lam"])
    ok = out["status"] == 0
    assert ok.sum() >= 6
    assert certify.passes_exit_test(res)[ok].all()
    # the numpy residuals are the ones the kernel tested (same definition, independent arithmetic)
    for j, k in enumerate(("res_stat", "res_eq", "res_ineq", "res_comp")):
        assert np.abs(res[k][ok] - out["res"][ok, j]).max() < 1e-7, k
    # a perturbed iterate must fail: the certificate is not vacuous
    x2 = out["x"].copy()
    x2[:, 50, 0] += 1e-3
    assert not certify.passes_exit_test(certify.kkt_residuals(n, bp, x2, out["u"], out["pi"], out["lam"]))[ok].any()
    pi2 = out["pi"].copy()
    pi2[:, 50, 0] += 1e-2
    bad = certify.kkt_residuals(n, bp, out["x"], out["u"], pi2, out["lam"])
    assert (bad["res_stat"][ok] > 1e-3).all()


@pytest.mark.parametrize("n", [2, 3])
def test_k4_first_qp_satisfies_dense_kkt(oracle, emu, n):
    """K4: the step of ONE linearised QP (an SQP_RTI step of the VBOC family, QP solved to 1e-9)
    satisfies the KKT conditions of that QP assembled densely in numpy -- including the stage-0 equalities the
    engine eliminates instead of handing them to the IPM as lb == ub pairs, and the terminal v_N = 0."""
    bp = pr.sample_vboc(n, 4, seed=5)
    o = _opts(emu, oracle.default_opts(0), qp_tol_stat=1e-9, qp_tol_eq=1e-9, qp_tol_ineq=1e-9, qp_tol_comp=1e-9)
    out = emu.solve_batch(n, 0, 1, bp, o, multipliers=True)
    assert (out["qp_status"] == 0).all()
    for b in range(4):
        N = int(bp["N"][b])
        dx = out["x"][b, :N + 1, :2 * n] - bp["x_guess"][b, :N + 1, :2 * n]
        du = out["u"][b, :N] - bp["u_guess"][b, :N]
        r = certify.qp_kkt(n, bp, b, dx, du, out["pi"][b], out["lam"][b])
        assert r["res_g"] < 1e-8 and r["res_b"] < 1e-8 and r["res_d"] < 1e-8 and r["res_m"] < 1e-8, r
        assert r["lam_min"] >= 0.0


def test_al_labels_equal_lp_feasibility(oracle, emu):
    n = 2
    bp = pr.sample_al(n, 24, seed=4)
    out = emu.solve_batch(n, 1, 1, bp, _opts(emu, oracle.default_opts(1)))
    lab = out["status"] == 0
    feas = np.array([certify.al_lp_feasible(n, bp["lbx0"][b, :2 * n])[0] for b in range(24)])
    assert lab.any() and (~lab).any()
    assert (lab == feas).all()


def test_free_dt_solutions_are_kkt_points_with_recovered_multipliers(oracle, emu):
    """configs[0] (1-DOF, dt a free state, VBOC/pendulum_vboc.py): the lane kernel exports no multipliers, so the
    certificate recovers them by bounded least squares from the returned iterate (free sign for the equalities, >= 0 for
    the bounds that are active) -- the point is a KKT point iff the remaining stationarity residual is below tolerance.
    Negative control: the same point is NOT a KKT point of the problem with the opposite velocity cost."""
    bp = pr.pendulum_free_dt_problems(10, seed=3)
    out = emu.solve_batch(1, 0, 0, bp, _opts(emu, oracle.default_opts(0)), "lane_dts")
    ok = np.where(out["status"] == 0)[0]
    assert len(ok) >= 8
    for b in ok:
        r = certify.free_dt_kkt(bp, b, out["x"][b], out["u"][b])
        assert r["res_stat"] < 1e-3 and r["res_eq"] < 1e-6 and r["res_ineq"] < 1e-6 and r["lam_min"] >= 0.0, (b, r)
        assert r["n_active"] >= 40                      # bang-bang: the force bound is active almost everywhere
    wrong = dict(bp)
    wrong["p"] = bp["p"].copy()
    wrong["p"][:, 0] *= -1.0
    for b in ok[:4]:
        assert certify.free_dt_kkt(wrong, b, out["x"][b], out["u"][b])["res_stat"] > 1e-2
