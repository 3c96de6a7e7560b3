"""Time sla_gemm (shared A, per-matrix B) at (N, batch): python profiles/gemm_bench.py N batch [c]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]); batch = int(sys.argv[2]); cplx = "c" in sys.argv[3:]
shared = "s" in sys.argv[3:]   # one A for the whole batch (stride 0): the group-product shape
dt = torch.complex128 if cplx else torch.float64
A = torch.randn(batch, n, n, dtype=torch.float64, device="cuda").to(dt)
B = torch.randn(batch, n, n, dtype=torch.float64, device="cuda").to(dt)
C = torch.empty_like(B)
if shared:
    A = A[:1].contiguous()
for _ in range(3):
    sla.gemm(C, A, B)
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sla.gemm(C, A, B)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = (8.0 if cplx else 2.0) * n ** 3 * batch
print(f"tile={os.environ.get('SLA_GEMM_TILE','auto'):4s} sa={os.environ.get('SLA_GEMM_SA','1')} N={n} batch={batch} {'c128' if cplx else 'f64'}{' sharedA' if shared else ''}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.2f} TFLOP/s")
# cuBLAS comparison
Cb = torch.empty_like(B)
for _ in range(3):
    torch.bmm(A.expand(batch, n, n) if shared else A, B, out=Cb)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    torch.bmm(A.expand(batch, n, n) if shared else A, B, out=Cb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"cuBLAS bmm                         : {ms*1e3:9.1f} us  {fl/ms/1e9:7.2f} TFLOP/s")
