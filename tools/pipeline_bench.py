"""Pipeline throughput: the full `data_generation` of `P` problems (extensions, retries,
sub-OCP chains, twin simulation) through the batched drivers.  python tools/pipeline_bench.py [n] [P] [mode]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from vboc_b200 import drivers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mode = sys.argv[3] if len(sys.argv) > 3 else "rounds"
stats = {}
t0 = time.perf_counter()
if mode == "rounds":
    X = drivers.data_generation_batch(n, P, seed=5, stats=stats)
else:
    X = drivers.data_generation_stream(n, P, seed=5, stats=stats)
dt = time.perf_counter() - t0
print(f"mode {mode} n {n} problems {P} rows {X.shape[0]} wall {dt:.1f} s  {stats}")
print(f"  {stats.get('solves', 0) / dt:.1f} solves/s  {stats.get('converged', 0) / dt:.1f} converged solves/s  {P / dt:.2f} problems/s")
print("  checksum", float(np.round(X, 6).sum()))
