"""Per-phase cycles of lu_factor (needs a build with SLA_NVCC_EXTRA="-DSLA_LU_PROFILE"; the kernel prints the counters)."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
rng = np.random.default_rng(0)
A = rng.standard_normal((64, 256, 256))
dA = sla.to_device(A); ws = sla.ldr_workspace(dA)
X = dA.clone(); sla.det_lu_(X, ws); torch.cuda.synchronize()
