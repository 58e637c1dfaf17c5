import sys, time
sys.path.insert(0,'.')
import numpy as np
from vboc_b200 import problems as pr, engine
for B in (2368, 16384):
    bp = pr.sample_vboc(3, B, seed=1)
    sol = engine.BatchSolver(3, 'vboc', B, 100)
    t=time.time(); sol.upload(bp); t_up=time.time()-t
    for rep in range(2):
        ms = sol.solve_resident(0)
        print('B',B,'kernel ms',ms,'OCP/s',B/ms*1e3, flush=True)
    t=time.time(); out = sol.download(); t_dn=time.time()-t
    print(' upload s',t_up,'download s',t_dn)
    print(' status', np.unique(out['status'],return_counts=True), 'sqp mean', out['sqp_iter'].mean(), 'qp/sqp', out['qp_iter'].sum()/out['sqp_iter'].sum(), 'ls/sqp', out['ls_evals'].sum()/out['sqp_iter'].sum())
    sol.close()
