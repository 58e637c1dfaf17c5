"""Certificates at scale for the families with general rows: the learned margin as hard terminal / soft
per-stage constraint of the Safe-MPC OCPs, and the Cartesian path constraint of the VBOC OCP.  numpy only
(tools/certify.py); writes a markdown report.  Usage: python tools/certify_rows.py [out.md] [scale]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import certify  # noqa: E402
from vboc_b200 import engine, problems as pr  # noqa: E402
from test_mpc_family import _row_penalties, make_net  # noqa: E402


def solve_mpc(n, bp, net, mode, Z, qp_tol=1e-8, tol=1e-2):
    B = len(bp["N"])
    sol = engine.BatchSolver(n, "mpc", B, int(bp["x_guess"].shape[1] - 1))
    o = engine.default_opts("mpc")
    o.tol_stat = o.tol_eq = o.tol_ineq = o.tol_comp = tol
    o.qp_tol_stat = o.qp_tol_eq = o.qp_tol_ineq = o.qp_tol_comp = qp_tol
    sol.set_opts(o)
    w = dict(net)
    w["W3"], w["b3"] = net["W3"][None, :], np.array([net["b3"]])
    sol.set_mpc(w, net["mean"], net["std"], 100.0 * (1.0 - net["scale"]), bp["W"], bp["W_e"], lh=bp["lh"], uh=bp["uh"],
                vstart=net.get("vstart"))
    sol.set_mpc_reference(bp["yref_acados"], bp["yrefN"])
    sol.set_mpc_rows(Z)
    sol.export_multipliers(True)
    t0 = time.perf_counter()
    out = sol.solve(bp, mode)
    out["wall_s"] = time.perf_counter() - t0
    out["pi"], out["lam"] = sol.multipliers()
    out["rowm"], out["lamg"] = sol.mpc_rows(), sol.mpc_multipliers()
    sol.close()
    return out


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_certify_rows.md")
    scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    L = []
    log = L.append
    log("# Round 2 - certificates of the families with general rows (`tools/certify_rows.py`)\n")
    log("numpy only (`tools/certify.py`): every solved QP is checked against the dense KKT conditions of the QP linearised at "
        "the guess -- dynamics by complex-step RK4 Jacobians, the margin row by the complex-step gradient of a numpy restatement of "
        "`nn_decisionfunction`, slack variables included -- relative to the largest multiplier of the problem (the engine's QP "
        "stationarity test is relative to it, DESIGN.md section 7).  SQP_RTI, QP tolerance 1e-8, one B200.\n")
    log("| system, network | rows | problems | status 0 / 4 | QP iterations mean / max | slack > 1e-3 on | rows binding on | worst relative KKT residual | worst feasibility residual | solve wall |")
    log("|---|---|---|---|---|---|---|---|---|---|")
    for n, H, B in ((2, 300, 2048 * scale), (3, 500, 1024 * scale)):
        net = make_net(n, H, n, 4.0)
        net["scale"] = 0.98
        if n == 3:
            net["vstart"] = 2
        N = 10
        bp = pr.sample_mpc(n, B, seed=3)
        for kind in ("hard terminal", "soft_traj", "parallel", "receding", "generic"):
            Z = None if kind == "hard terminal" else _row_penalties(kind, B, N)
            out = solve_mpc(n, bp, net, 1, Z)
            ok = np.where(out["status"] == 0)[0]
            worst_rel = worst_feas = 0.0
            slack = bind = 0
            for b in ok:
                if Z is None:
                    r = certify.mpc_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["lamg"][b], 1.0,
                                        first_qp_at_guess=True)
                    sc = max(1.0, float(out["lamg"][b].max()))
                    bind += out["lamg"][b, 0] > 1e-3
                else:
                    r = certify.mpc_rows_kkt(n, bp, net, b, out["x"][b], out["u"][b], out["pi"][b], out["lam"][b], out["rowm"][b],
                                             Z[b], 1.0, first_qp_at_guess=True)
                    sc = max(1.0, float(np.abs(out["rowm"][b, :, :4]).max()))
                    slack += r["sl"].max() > 1e-3
                    bind += out["rowm"][b, :, 0].max() > 1e-3
                worst_rel = max(worst_rel, max(r["res_stat"], r["res_comp"]) / sc)
                worst_feas = max(worst_feas, r["res_eq"], r["res_ineq"])
                assert r["lam_min"] >= 0.0
            log(f"| {n}-DOF, {2 * n}-{H}-{H}-1{' (vel_norm over x[2:])' if n == 3 else ''} | {kind} | {B} | {len(ok)} / {int((out['status'] == 4).sum())} | "
                f"{out['qp_iter'].mean():.1f} / {out['qp_iter'].max()} | {slack if Z is not None else '-'} | {bind} | {worst_rel:.1e} | {worst_feas:.1e} | "
                f"{out['wall_s'] * 1e3:.0f} ms |")
            print(L[-1], flush=True)
    # Cartesian path constraint
    from test_cartesian_family import CART
    B = 8192 * scale
    bp = pr.sample_vboc(2, B, seed=5)
    sol = engine.BatchSolver(2, "vboc", B, 100)
    base = sol.solve(bp)
    sol.set_cartesian(CART[0], CART[1], np.sqrt(CART[2]), CART[3])
    sol.export_multipliers(True)
    t0 = time.perf_counter()
    out = sol.solve(bp)
    wall = time.perf_counter() - t0
    pi, lam = sol.multipliers()
    rowm = sol.mpc_rows()
    sol.close()
    res = certify.kkt_residuals(2, bp, out["x"], out["u"], pi, lam, cart=dict(xc=CART[0], yc=CART[1], lh=CART[2], uh=CART[3], rowm=rowm))
    ok = out["status"] == 0
    cert = certify.passes_exit_test(res)
    h, _ = certify.cartesian_h(out["x"][:, :100, :2], CART[0], CART[1])
    hb, _ = certify.cartesian_h(base["x"][:, :100, :2], CART[0], CART[1])
    h0, _ = certify.cartesian_h(bp["lbx0"][:, :2], CART[0], CART[1])
    log(f"\n## Cartesian path constraint (`vboc_set_cartesian`, VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:147-160): {B} 2-DOF VBOC problems, SQP\n")
    log(f"* status counts with the constraint: {dict(zip(*[a.tolist() for a in np.unique(out['status'], return_counts=True)]))}; without: "
        f"{dict(zip(*[a.tolist() for a in np.unique(base['status'], return_counts=True)]))}; solve wall {wall:.2f} s")
    log(f"* status 0 and acados' exit test recomputed in numpy WITH the row passes: **{int((ok & cert).sum())} / {int(ok.sum())}**; "
        f"status != 0 whose point nevertheless passes: {int((~ok & cert).sum())}")
    log(f"* max |numpy - engine| of the four residuals over status-0 results: "
        + ", ".join(f"{np.abs(res[k][ok] - out[k][ok]).max():.1e}" for k in ("res_stat", "res_eq", "res_ineq", "res_comp")))
    log(f"* smallest h - radius^2 over the returned trajectories: {float(h[ok].min() - CART[2]):.2e}; unconstrained optima that cross the "
        f"circle: {int(((base['status'] == 0) & (hb.min(axis=1) < CART[2] - 1e-6)).sum())}; problems with the row binding: "
        f"{int((ok & (rowm[:, :, 0] > 1e-6).any(axis=1)).sum())}")
    inside = (h0 < CART[2] - 1e-6)
    log(f"* problems whose fixed initial position lies inside the circle: {int(inside.sum())}, of these reported solved: {int((ok & inside).sum())}")
    open(path, "w").write("\n".join(L) + "\n")
    print("\n".join(L[-6:]))


if __name__ == "__main__":
    main()
