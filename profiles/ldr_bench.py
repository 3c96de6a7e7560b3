"""Time sla_ldr for one (N, batch, dtype) with the LDR kernel chosen by SLA_LDR_ALGO (reg|cluster|streaming|unset)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]); batch = int(sys.argv[2]); cplx = len(sys.argv) > 3 and sys.argv[3] == "c"
rng = np.random.default_rng(0)
A = rng.standard_normal((min(batch, 64), n, n)) * (10.0 ** rng.uniform(-30, 30, size=(min(batch, 64), 1, n)))
if cplx:
    A = A + 1j * rng.standard_normal(A.shape)
dA = sla.to_device(A)
dA = dA.repeat((batch + dA.shape[0] - 1) // dA.shape[0], 1, 1)[:batch].contiguous()
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
for _ in range(2):
    sla.ldr_(F, dA, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    sla.ldr_(F, dA, ws)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"algo={os.environ.get('SLA_LDR_ALGO','auto'):9s} N={n} batch={batch} {'c128' if cplx else 'f64'}: {ms:8.3f} ms/launch  {ms/batch*1e3:8.1f} us/matrix")
