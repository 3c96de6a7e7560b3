"""Time inv_lu! on small matrices in ONE process (SLA_LU_SMALL=0 restores the general LU kernel for n <= 64).  Back-to-back
launches on one stream, CUDA events around `reps` of them."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla

rng = np.random.default_rng(0)
for n, batch, cplx in [(64, 1, False), (64, 64, False), (64, 2048, False), (32, 1, False), (32, 2048, False), (64, 1, True), (64, 1024, True), (48, 1, False)]:
    nb = min(batch, 16)
    A = rng.standard_normal((nb, n, n)) + (1j * rng.standard_normal((nb, n, n)) if cplx else 0.0)
    dA = sla.to_device(A).repeat((batch + nb - 1) // nb, 1, 1)[:batch].contiguous()
    X = dA.clone()
    ws = sla.ldr_workspace(X)
    for _ in range(5):
        X.copy_(dA); sla.inv_lu_(X, ws)
    torch.cuda.synchronize()
    reps = 100 if batch <= 64 else 20
    # the copy that restores the input is timed separately and subtracted
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        X.copy_(dA)
    e[1].record()
    for _ in range(reps):
        X.copy_(dA); sla.inv_lu_(X, ws)
    e[2].record(); torch.cuda.synchronize()
    ms = (e[1].elapsed_time(e[2]) - e[0].elapsed_time(e[1])) / reps
    print(f"SLA_LU_SMALL={os.environ.get('SLA_LU_SMALL', '1')} N={n:3d} batch={batch:5d} {'c128' if cplx else 'f64 '}: "
          f"{ms * 1e3:9.1f} us/launch  {ms / batch * 1e3:8.2f} us/matrix", flush=True)
