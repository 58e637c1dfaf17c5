"""SURVEY 8(f)4 on the GPU: the Cartesian path constraint through the C-ABI (`vboc_set_cartesian` on a VBOC-family handle,
`vboc_solve_batch`, `vboc_download_multipliers`, `vboc_download_mpc_rows`), certified with numpy only, and the drop-in
class of `VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py` called like `vboc_multiprocessing.py` calls it."""
import numpy as np
import pytest

from vboc_b200 import problems as pr
from test_cartesian_family import CART, check_cartesian

pytestmark = pytest.mark.gpu


def test_cartesian_rows_certified_at_scale():
    from vboc_b200 import engine
    n, B = 2, 1024
    bp = pr.sample_vboc(n, B, seed=5)
    sol = engine.BatchSolver(n, "vboc", B, 100)
    base = sol.solve(bp)
    sol.set_cartesian(CART[0], CART[1], np.sqrt(CART[2]), CART[3])
    sol.export_multipliers(True)
    out = sol.solve(bp)
    out["pi"], out["lam"] = sol.multipliers()
    out["rowm"] = sol.mpc_rows()
    again = sol.solve(bp)
    assert (again["x"] == out["x"]).all() and (again["status"] == out["status"]).all()   # bit-identical re-run
    sol.set_cartesian(on=False)
    back = sol.solve(bp)
    sol.close()
    assert (back["x"] == base["x"]).all()                     # switching the constraint off restores the VBOC kernel
    check_cartesian(bp, base, out)


def test_cartesian_class_like_the_driver():
    """VBOC/Cartesian constraints/vboc_multiprocessing.py:131-241 (`testing`): sample one problem, OCP_solve, read the
    boundary state; the constrained trajectory stays outside the circle."""
    import tools_path  # noqa: F401
    import certify
    from vboc_b200.shim.Cartesian.doublependulum_class_fixedveldir import OCPdoublependulumINIT
    ocp = OCPdoublependulumINIT()
    bp = pr.sample_vboc(2, 64, seed=5)
    solved = 0
    for b in range(12):
        N = int(bp["N"][b])
        status = ocp.OCP_solve(bp["x_guess"][b, :N + 1], bp["u_guess"][b, :N], bp["p"][b], bp["lbx"][b], bp["ubx"][b], bp["lbu"][b],
                               bp["ubu"][b], bp["lbx0"][b], bp["ubx0"][b], bp["lbxN"][b], bp["ubxN"][b])
        if status != 0:
            continue
        solved += 1
        X = np.stack([ocp.ocp_solver.get(k, "x") for k in range(N + 1)])
        h, _ = certify.cartesian_h(X[:N, :2], ocp.x_c, ocp.y_c)
        assert h.min() > ocp.radius ** 2 - 1e-6
        assert abs(ocp.ocp_solver.get_cost() - bp["p"][b, :2] @ X[0, 2:4] - bp["p"][b, 2] * 1e-2 * N) < 1e-9
    assert solved >= 6
