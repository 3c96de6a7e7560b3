"""Per-warp timeline of 4 columns (k = 100..103) of the register LDR kernel, CTA 0 (needs a build with
SLA_NVCC_EXTRA="-DSLA_LDR_TRACE"): python profiles/ldr_trace.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n, batch = 256, 8
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, n, n)) * (10.0 ** rng.uniform(-40, 40, size=(batch, 1, n)))
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
prof = torch.zeros(128 + 4 * 16 * 12 + 64, dtype=torch.int64, device="cuda")
lib = sla.load()
for _ in range(2):
    sla.ldr_(F, dA, ws)
torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(prof.data_ptr())
sla.ldr_(F, dA, ws); torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(None)
t = prof.cpu().numpy()[128:128 + 4 * 16 * 12].reshape(4, 16, 12)
names = ["wait done", "pick done", "dots done", "reduce done", "update done", "key ready", "after sync", "scan done",
         "norm reduce", "scalars", "copy issued"]
for kk in range(3):
    base = t[kk][t[kk, :, 0] > 0, 0].min()
    nxt = t[kk + 1][t[kk + 1, :, 0] > 0, 0].min()
    print(f"column {100 + kk}: next column starts at +{nxt - base} cycles")
    for ev in range(11):
        col = t[kk, :, ev]
        v = col[col > 0] - base
        if v.size:
            print(f"  {names[ev]:12s} min {v.min():6d}  median {int(np.median(v)):6d}  max {v.max():6d}   (warps reporting: {v.size})")
