"""Bounded run of the solve kernel for ncu: B problems, SQP iterations capped (same code path as the
benchmark, shorter)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from vboc_b200 import problems as pr, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bp = pr.sample_vboc(3, B, seed=1)
sol = engine.BatchSolver(3, 'vboc', B, 100)
o = engine.default_opts('vboc'); o.max_iter = cap
sol.set_opts(o)
sol.upload(bp)
for rep in range(2):
    ms = sol.solve_resident(0)
    print('B', B, 'cap', cap, 'kernel ms', ms, flush=True)
out = sol.download()
print('sqp', out['sqp_iter'].sum(), 'qp', out['qp_iter'].sum(), 'IPM iters/s', out['qp_iter'].sum() / ms * 1e3)
