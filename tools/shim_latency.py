"""Latency of the drop-in classes used one problem at a time, the way an UNMODIFIED reference driver would call them
(VBOC/triplependulum_vboc.py:110-129, 347-352): one OCP_solve = one kernel launch on one warp.  INTEGRATION.md quotes
these numbers next to the batched entry points."""
import importlib, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from vboc_b200 import problems as pr
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vboc_b200", "shim", "VBOC"))
m = importlib.import_module("triplependulum_class_vboc")
ocp, sim = m.OCPtriplependulumINIT(), m.SYMtriplependulumINIT()
bp = pr.sample_vboc(3, 12, seed=3)
ts, its = [], []
for b in range(12):
    one = pr.take(bp, b)
    ocp.N = 100
    t = time.perf_counter()
    st = ocp.OCP_solve(one["x_guess"][:100], one["u_guess"], one["p"], one["lbx"], one["ubx"], one["lbu"], one["ubu"],
                       one["lbx0"], one["ubx0"], one["lbxN"], one["ubxN"])
    x = np.array([ocp.ocp_solver.get(i, "x") for i in range(101)])
    ts.append(time.perf_counter() - t)
print(f"OCP_solve + 101 get(): median {np.median(ts)*1e3:.0f} ms, min {min(ts)*1e3:.0f} ms, max {max(ts)*1e3:.0f} ms over 12 problems")
x0, u0 = np.array([3.0, 3.2, 2.9, 1.0, -2.0, 0.5]), np.array([1.0, -3.0, 2.0])
t = time.perf_counter()
for _ in range(200):
    sim.acados_integrator.set("u", u0); sim.acados_integrator.set("x", x0); sim.acados_integrator.set("T", 1e-2)
    sim.acados_integrator.solve(); x0 = sim.acados_integrator.get("x")
print(f"simulator set/solve/get: {(time.perf_counter() - t) / 200 * 1e6:.0f} us per RK4 step")
