"""Cycles per phase of the small-matrix LDR kernel, version 2 (thread 0 of CTA 0; library built with -DSLA_SMALL_PROFILE and
selected through SLA_B200_LIB): python profiles/ldr_small_phases.py [N] [batch]

Build of the profiling library (from stablelinearalgebra.jl_b200/csrc, after the normal build has filled build/):
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -DSLA_SMALL_PROFILE -c sla_ldr_small.cu -o /tmp/small_prof.o
  nvcc -shared -o abl/libsla_b200_smallprof.so $(ls build/*.o | grep -v sla_ldr_small.o) /tmp/small_prof.o \
       -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
  SLA_B200_LIB=$PWD/abl/libsla_b200_smallprof.so SLA_LDR_SMALL=3 python ../../profiles/ldr_small_phases.py 64 1"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, n, n)) * 10.0 ** rng.uniform(-20, 20, size=(batch, 1, n))
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
p = prof.cpu().numpy()[:8]
ver = os.environ.get("SLA_LDR_SMALL", "3")
names = (["norms+store (next column)", "barrier 1", "own larfg | pivot | raw dot", "barrier 2", "scalars of the pivot, w, update", "(last barrier)",
          "R extraction", "Q formation+store"] if ver == "3" else
         ["norms+store", "barrier", "pivot (keys, 3 redux)", "pivot data + sqrt/rcp", "dot+shuffles+update", "(last barrier)",
          "R extraction", "Q formation+store"])
print(f"SLA_LDR_SMALL={ver} N={n} batch={batch} kernel={sla.last_ldr_algo()} {e0.elapsed_time(e1) * 1e3:.0f} us; cycles of thread 0, CTA 0: total {p.sum()}")
for nm, v in zip(names, p):
    print(f"  {nm:34s} {v:8d}  ({v / n:7.0f} per column)")
