"""Throughput of the other BASELINE.json configs on one GPU (the headline config C4 is bench.py):
C2 doublependulum_testdata, C3 doublependulum_vboc data generation, C5 triplependulum_al labelling + query."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from vboc_b200 import drivers, engine, problems as pr
from vboc_b200._lib import MODE_RTI

def timed(f):
    t = time.perf_counter(); r = f(); return r, time.perf_counter() - t

st = {}
X, dt = timed(lambda: drivers.testing_stream(2, 1024, seed=1, stats=st))
print(f"C2 doublependulum_testdata: 1024 test points, {st['solves']} solves ({st['converged']} converged) in {dt:.1f} s "
      f"= {st['converged'] / dt:.0f} converged solves/s, {X.shape[0] / dt:.0f} points/s")
st = {}
X, dt = timed(lambda: drivers.data_generation_stream(2, 2048, seed=1, stats=st))
print(f"C3 doublependulum_vboc data generation: 2048 problems -> {X.shape[0]} rows, {st['solves']} solves in {dt:.1f} s "
      f"= {st['converged'] / dt:.0f} converged solves/s ({st.get('solves_per_s_first_90pct')} while full)")
n, B = 3, 46656
bp = pr.sample_al(n, B, seed=3)
sol = engine.BatchSolver(n, "al", B, 100)
sol.solve(bp, MODE_RTI)
out, dt = timed(lambda: sol.solve(bp, MODE_RTI))
print(f"C5 triplependulum_al labelling: {B} states, one SQP_RTI each, {dt:.2f} s host-to-host = {B / dt:.0f} labels/s "
      f"(kernel {sol.last_kernel_ms:.0f} ms = {B / sol.last_kernel_ms * 1e3:.0f}/s; IPM iterations {int(out['qp_iter'].sum())}, "
      f"{out['qp_iter'].sum() / sol.last_kernel_ms * 1e3 / 1e6:.2f} M/s); viable {(out['status'] == 0).mean() * 100:.1f} %")
sol.close()
