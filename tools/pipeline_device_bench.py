"""Pipeline throughput of the device-resident state machine: python tools/pipeline_device_bench.py [n] [P ...]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from vboc_b200 import drivers, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for P in [int(a) for a in sys.argv[2:]] or [1024]:
    dg = engine.DataGenerator(n, P)
    drivers.data_generation_device(n, min(P, 64), seed=3, dgen=dg)
    t0 = time.perf_counter(); inp = drivers.dg_inputs(n, P, seed=5); ti = time.perf_counter() - t0
    t0 = time.perf_counter(); rows, st = dg.run(inp); dt = time.perf_counter() - t0
    print(f"n {n} problems {P}: inputs {ti:.2f} s, run {dt:.2f} s (kernel {dg.last_kernel_ms / 1e3:.2f} s), rows {rows.shape[0]}, "
          f"solves {st['solves'].sum()}, converged {st['converged'].sum()}, sim steps {st['sim_steps'].sum()}, "
          f"failed problems {(st['status'] == 1).sum()} -> {st['converged'].sum() / dt:.0f} converged solves/s; problems done after "
          f"p50/p90/p99/max = {np.round(np.percentile(st['t_done_us'], [50, 90, 99, 100]) * 1e-6, 2).tolist()} s; "
          f"solves per problem max {st['solves'].max()}, failed solves per problem max {(st['solves'] - st['converged']).max()}", flush=True)
    dg.close()
