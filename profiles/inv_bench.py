"""Time sla_inv_IpA (2 LU inversions + GEMM + element-wise) at (N, batch): python profiles/inv_bench.py N batch"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]); batch = int(sys.argv[2])
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, n, n))
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA, ws)
G = torch.empty_like(dA)
for _ in range(3):
    sla.inv_IpA_(G, F, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    sla.inv_IpA_(G, F, ws)
e1.record(); torch.cuda.synchronize()
print(f"inv_IpA N={n} batch={batch}: {e0.elapsed_time(e1)/10:.3f} ms/call")
X = dA.clone()
for _ in range(2):
    X.copy_(dA); sla.det_lu_(X, ws)
torch.cuda.synchronize()
tot = 0.0
for _ in range(5):
    X.copy_(dA); torch.cuda.synchronize()
    e0.record(); sla.det_lu_(X, ws); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"det_lu (factor only) N={n} batch={batch}: {tot/5:.3f} ms/call")
tot = 0.0
for _ in range(5):
    X.copy_(dA); torch.cuda.synchronize()
    e0.record(); sla.inv_lu_(X, ws); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"inv_lu (factor + unsplit sweeps + copy) N={n} batch={batch}: {tot/5:.3f} ms/call")
