"""Throughput and cross-check of the two MLP inference kernels (tcgen05 3xTF32 vs FP32 CUDA cores)."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.nn as nn
from vboc_b200 import nn as vnn
n, H, B = 3, 500, 1 << 20
torch.manual_seed(0)
m = nn.Module(); m.linear_relu_stack = nn.Sequential(nn.Linear(2*n, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU(), nn.Linear(H, 2))
rng = np.random.default_rng(0)
X = np.concatenate([rng.uniform(2.36, 3.93, (B, n)), rng.uniform(-10.5, 10.5, (B, n))], axis=1).astype(np.float32)
with torch.no_grad():
    ref = m.linear_relu_stack((torch.from_numpy(X[:65536]) - 3.0) / 5.0).numpy()
res = {}
for name, env in (("tcgen05_3xtf32_pipelined", {}), ("tcgen05_3xtf32_serial_r1", {"VBOC_MLP_SERIAL": "1"}),
                  ("fp32_cuda_cores", {"VBOC_MLP_CUDA_CORES": "1"})):
    os.environ.pop("VBOC_MLP_SERIAL", None), os.environ.pop("VBOC_MLP_CUDA_CORES", None)
    os.environ.update(env)
    net = vnn.MLP.from_torch(m)
    net.entropy(X[:4096], 3.0, 5.0)
    t = time.perf_counter(); out, etp = net.entropy(X, 3.0, 5.0); dt = time.perf_counter() - t
    flops = 2.0 * B * (2*n*H + H*H + H*2)
    kms = net.last_kernel_ms
    print(f"{name}: {dt*1e3:.1f} ms for {B} rows incl. H2D/D2H and allocation; kernel alone {kms:.2f} ms = "
          f"{flops/kms/1e9:.1f} TFLOP/s algorithmic ({flops/B/1e6:.2f} MFLOP per row), max|out-torch| = {np.abs(out[:65536]-ref).max():.2e}")
    res[name] = out
    net.close()
print("max |pipelined - fp32| =", np.abs(res["tcgen05_3xtf32_pipelined"] - res["fp32_cuda_cores"]).max())
os.environ.pop("VBOC_MLP_SERIAL", None), os.environ.pop("VBOC_MLP_CUDA_CORES", None)
# the AL pool pass on a resident pool: 15^6 states (AL/triplependulum_al.py:100)
P = 15 ** 6
Xp = np.concatenate([rng.uniform(2.36, 3.93, (P, n)), rng.uniform(-10.5, 10.5, (P, n))], axis=1).astype(np.float32)
net = vnn.MLP.from_torch(m)
t = time.perf_counter(); rp = vnn.ResidentPool(Xp); tu = time.perf_counter() - t
rp.score(net, 3.0, 5.0)
t = time.perf_counter(); ms = rp.score(net, 3.0, 5.0); ts = time.perf_counter() - t
t = time.perf_counter(); idx, rows, sc = rp.select(6 ** 6); tsel = time.perf_counter() - t
t = time.perf_counter(); rp.remove_selected(); trem = time.perf_counter() - t
flops = 2.0 * P * (2*n*H + H*H + H*2)
print(f"resident pool {P} rows: upload once {tu*1e3:.0f} ms; score kernel {ms:.1f} ms = {flops/ms/1e9:.1f} TFLOP/s algorithmic "
      f"(host-to-host {ts*1e3:.1f} ms); top-{6**6} select {tsel*1e3:.1f} ms; remove {trem*1e3:.1f} ms; pool now {len(rp)}")
