"""2-GPU check of the sharded AL query on resident pools (torchrun --nproc-per-node 2 tools/al_sharded_check.py):
every rank holds a shard of the pool in ITS GPU's HBM; the union of the ranks' selections must be the global top-B
of the whole pool, the parts disjoint, the pools shrunk accordingly."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from vboc_b200 import drivers, nn as vnn, distributed as vd
from vboc_b200.shim.my_nn import NeuralNetCLS

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = NeuralNetCLS(6, 500, 2)
rng = np.random.default_rng(1)
P, B = 400_000, 5000
full = np.concatenate([rng.uniform(2.36, 3.93, (P, 3)), rng.uniform(-10.5, 10.5, (P, 3))], axis=1).astype(np.float32)
lo, hi = vd.shard_range(P, rank, world)
net = vnn.MLP.from_torch(model, device=local)
rp = vnn.ResidentPool(full[lo:hi], device=local)
idx, rows, emax = drivers.al_query_resident(rp, net, 3.0, 5.0, B, sharded=True)
_, etp_full = net.entropy(full, 3.0, 5.0)
sel_global = np.asarray(idx) + lo
counts = vd.all_gather_rows(np.array([[float(len(idx))]]))[:, 0].astype(int)
allsel = vd.all_gather_rows(sel_global.astype(np.float64)[:, None])[:, 0].astype(np.int64)
kth = np.partition(etp_full, -B)[-B]
ok = (len(allsel) == B and len(np.unique(allsel)) == B and (etp_full[allsel] >= kth).all()
      and set(np.where(etp_full > kth)[0].tolist()) <= set(allsel.tolist())
      and len(rp) == (hi - lo) - len(idx) and abs(emax - float(etp_full.max())) < 1e-12
      and np.array_equal(rows, full[sel_global]))
print(f"rank {rank}: selected {len(idx)} of {B} (per rank {counts.tolist()}), pool {hi - lo} -> {len(rp)}, ok = {ok}", flush=True)
rp.close(); net.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
