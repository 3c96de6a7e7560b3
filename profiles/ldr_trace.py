"""Per-warp timeline of 4 columns (k = 100..103) of the register LDR kernel for the CTAs of cluster 0 (needs a build
with SLA_NVCC_EXTRA="-DSLA_LDR_TRACE"): python profiles/ldr_trace.py
Events (clock64 of lane 0 of each warp, relative to the earliest "wait done" of the CTA in that column):
0 wait done, 1 pick done, 2 dots done, 3 reduce done, 4 update done, 5 key stored, 6 after barrier, 7 scan done,
8 (winner) norm reduced, 9 scalars done, 10 copies issued."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n, batch = 256, 8
CS = 2 if os.environ.get("SLA_LDR_ALGO", "").startswith("h") else 4
rng = np.random.default_rng(0)
graded = "graded" in sys.argv   # DQMC-like input: column scales sorted descending
A = rng.standard_normal((batch, n, n)) * ((10.0 ** np.linspace(20, -20, n))[None, None, :] if graded
                                          else 10.0 ** rng.uniform(-40, 40, size=(batch, 1, n)))
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
prof = torch.zeros(128 + 4 * CS * 16 * 12 + 64, dtype=torch.int64, device="cuda")
lib = sla.load()
for _ in range(2):
    sla.ldr_(F, dA, ws)
torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(prof.data_ptr())
sla.ldr_(F, dA, ws); torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(None)
t = prof.cpu().numpy()[128:128 + 4 * CS * 16 * 12].reshape(4, CS, 16, 12)
names = ["wait done", "pick done", "dots done", "reduce done", "update done", "key stored", "after barrier", "scan done",
         "norm reduced", "scalars", "copies issued"]
for kk in range(3):
    print(f"column {100 + kk}")
    for r in range(CS):
        base = t[kk, r][t[kk, r, :, 0] > 0, 0].min()
        nxt = t[kk + 1, r][t[kk + 1, r, :, 0] > 0, 0].min()
        line = []
        for ev in range(11):
            col = t[kk, r, :, ev]
            v = col[col > 0] - base
            if v.size:
                line.append(f"{names[ev]} {int(np.median(v))}/{v.max()}" if v.size > 1 else f"{names[ev]} {v.max()}")
        print(f"  CTA {r}: next wait done +{nxt - base}: " + "; ".join(line))
