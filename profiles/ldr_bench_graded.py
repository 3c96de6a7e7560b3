"""Like ldr_bench.py but with the column scaling a DQMC chain produces: (random) * diag(d), d sorted DESCENDING over
`decades` orders of magnitude -- the input of every re-factorisation lmul!(B, F) performs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]); batch = int(sys.argv[2]); decades = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, n, n)) * (10.0 ** np.linspace(decades / 2, -decades / 2, n))[None, None, :]
dA = sla.to_device(A)
ws = sla.ldr_workspace(dA)
F = sla.ldr(dA)
for _ in range(2):
    sla.ldr_(F, dA, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    sla.ldr_(F, dA, ws)
e1.record(); torch.cuda.synchronize()
print(f"algo={os.environ.get('SLA_LDR_ALGO','auto'):9s} N={n} batch={batch} graded {decades:.0f} decades: {e0.elapsed_time(e1)/3:8.3f} ms/launch")
