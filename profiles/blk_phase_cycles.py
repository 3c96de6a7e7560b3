"""Per-phase cycle counters of the blocked LDR kernel (needs a build with -DSLA_BLK_PROFILE, see
profiles/scripts/r2_blk_prof.sh): python profiles/blk_phase_cycles.py [N] [batch] [decades]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
decades = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, n, n)) * (10.0 ** np.linspace(decades / 2, -decades / 2, n))[None, None, :]
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
p = prof.cpu().numpy()
print(f"N={n} batch={batch} decades={decades}: ldr! {e0.elapsed_time(e1)*1e3:.0f} us, algo {sla.api.last_ldr_algo()}")
main = ["wait candidates", "meta+GEMV", "reduce/f/row/down-date", "publish (non-owner)", "R row + reflector store", "trailing DMMA", "loop"]
own = ["keys+barrier+scan", "gather", "lazy update", "norm+dots+reduce", "reflector scalars", "stage+fence+issue"]
print("warp 0 of CTA 0, cycles per column: " + " | ".join(f"{nm} {p[i]/n:.0f}" for i, nm in enumerate(main)) + f" | total {p[:7].sum()/n:.0f}")
print("candidate builders (two CTAs; per column and CTA): " + " | ".join(f"{nm} {p[16+i]/n/2:.0f}" for i, nm in enumerate(own)) + f" | total {p[16:22].sum()/n/2:.0f}")
