"""Three ldr! and three inv_lu! launches on a batch of 64 x (64 x 64) matrices (for an ncu capture of the small-matrix kernels)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
rng = np.random.default_rng(0)
A = rng.standard_normal((64, 64, 64)) * 10.0 ** rng.uniform(-10, 10, size=(64, 1, 64))
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
for _ in range(2):
    sla.ldr_(F, dA, ws)
for _ in range(3):
    X = dA.clone()
    sla.inv_lu_(X, ws)
torch.cuda.synchronize()
print("ok", sla.last_ldr_algo())
