"""Time ldr! on small matrices over several (n, batch, eltype) shapes in ONE process (the LDR kernel choice is read from
the environment once per process: SLA_LDR_SMALL=0 restores the pre-small-kernel dispatch).  Back-to-back launches on one
stream, CUDA events around `reps` of them."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla

shapes = [(64, 1, False), (64, 64, False), (64, 2048, False), (32, 1, False), (32, 2048, False), (16, 1, False),
          (64, 1, True), (64, 1024, True), (48, 1, False)]
rng = np.random.default_rng(0)
for n, batch, cplx in shapes:
    nb = min(batch, 16)
    A = rng.standard_normal((nb, n, n)) * (10.0 ** rng.uniform(-20, 20, size=(nb, 1, n)))
    if cplx:
        A = A + 1j * rng.standard_normal(A.shape) * np.abs(A)
    dA = sla.to_device(A)
    dA = dA.repeat((batch + nb - 1) // nb, 1, 1)[:batch].contiguous()
    ws = sla.ldr_workspace(dA)
    F = sla.ldr(dA)
    for _ in range(5):
        sla.ldr_(F, dA, ws)
    torch.cuda.synchronize()
    reps = 200 if batch <= 64 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sla.ldr_(F, dA, ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"SLA_LDR_SMALL={os.environ.get('SLA_LDR_SMALL', '1')} kernel={sla.last_ldr_algo():9s} N={n:3d} batch={batch:5d} "
          f"{'c128' if cplx else 'f64 '}: {ms * 1e3:9.1f} us/launch  {ms / batch * 1e3:8.2f} us/matrix", flush=True)
