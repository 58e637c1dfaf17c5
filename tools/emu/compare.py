import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools/emu')
import numpy as np
from oracle import oracle as orc
from vboc_b200 import problems as pr
import emu
B=int(sys.argv[1]) if len(sys.argv)>1 else 8
KERNEL=sys.argv[2] if len(sys.argv)>2 else 'warp'
def copy_opts(oo):
    e = emu.Opts()
    for f,_ in emu.Opts._fields_: setattr(e,f,getattr(oo,f))
    return e
for n,fam,mode in ((3,0,1),(3,0,0),(2,0,0),(3,1,1),(2,1,1),(1,1,1)):
    bp = pr.sample_vboc(n,B,seed=1) if fam==0 else pr.sample_al(n,B,seed=2)
    oo = orc.default_opts(fam)
    t=time.time(); r = orc.solve_batch(n, fam, mode, bp, oo, nthreads=8); t1=time.time()-t
    t=time.time(); e = emu.solve_batch(n, fam, mode, bp, copy_opts(oo), KERNEL); t2=time.time()-t
    print('n',n,'fam',fam,'mode',mode,'oracle',round(t1,2),'emu',round(t2,2))
    print(' status', r['status'][:12], e['status'][:12])
    print(' sqp', r['sqp_iter'][:12], e['sqp_iter'][:12])
    print(' qp ', r['qp_iter'][:12], e['qp_iter'][:12])
    ok = (r['status']==0)&(e['status']==0)
    print(' agree status', (r['status']==e['status']).mean(), 'max|dx| on ok', np.abs(r['x']-e['x'])[ok].max() if ok.any() else None, 'cost diff', np.abs(r['cost']-e['cost'])[ok].max() if ok.any() else None)
