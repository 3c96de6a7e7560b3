"""Per-phase cycle counters of the cluster LDR kernel (cluster 0): python profiles/ldr_phase_cycles.py [N] [batch]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(0)
graded = len(sys.argv) > 3 and sys.argv[3] == "graded"   # DQMC-like: column scales sorted descending
A = rng.standard_normal((batch, n, n)) * ((10.0 ** np.linspace(20, -20, n))[None, None, :] if graded
                                          else 10.0 ** rng.uniform(-40, 40, size=(batch, 1, n)))
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
prof = torch.zeros(128, dtype=torch.int64, device="cuda")
lib = sla.load()
for it in range(3):
    sla.ldr_(F, dA, ws)
torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(prof.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sla.ldr_(F, dA, ws); e1.record(); torch.cuda.synchronize()
lib.sla_debug_set_ldr_prof(None)
p = prof.cpu().numpy().reshape(-1, 2, 8)
names = ["tail(publish)", "cluster barrier", "pick/finalise", "pass+finish", "(end)", "extract+init", "Q phase", "writeback"]
print(f"N={n} batch={batch} kernel {e0.elapsed_time(e1)*1e3:.0f} us")
sp = prof.cpu().numpy()[64:72]
if sp.sum():
    print("select_and_publish (sum over the cluster's CTAs, lane 0; per column per CTA): " + " | ".join(
        f"{nm} {v/n/4:.0f}" for nm, v in zip(["warp key reduce(w0)", "sync wait(w0)", "key scan(winner)", "col select+norm reduce", "scalar math", "remote stores"], sp)))
for r in range(p.shape[0]):
    if p[r].sum() == 0: continue
    for w in range(2):
        tot = p[r, w].sum()
        print(f"rank {r} warp {w}: total {tot} cyc | " + " | ".join(f"{nm} {v} ({v/n:.0f}/col)" for nm, v in zip(names, p[r, w])))
