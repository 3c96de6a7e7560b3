"""Correctness of one LDR kernel (SLA_LDR_ALGO in the environment) on random and graded input: reconstruction, unitarity of L,
|det R| = 1 and d against numpy's pivoted-QR-free reference (singular values are not d; d is compared with the hybrid kernel
by the caller running this twice)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
rng = np.random.default_rng(3)
ok = True
for n, batch, decades in ((256, 3, 0), (256, 64, 100), (200, 5, 40), (130, 2, 0), (255, 3, 30), (256, 70, 10)):
    A = rng.standard_normal((batch, n, n)) * (10.0 ** np.linspace(decades / 2, -decades / 2, n))[None, None, :]
    dA = sla.to_device(A)
    ws = sla.ldr_workspace(dA)
    try:
        F = sla.ldr(dA, ws)
        torch.cuda.synchronize()
    except Exception as e:
        print(n, batch, decades, "FAILED", e); ok = False; break
    L, d, R = sla.to_host(F.L), F.d.cpu().numpy(), sla.to_host(F.R)
    rec = np.abs(L @ (d[:, :, None] * R) - A).max(axis=(1, 2)) / np.abs(A).max(axis=(1, 2))
    orth = np.abs(np.swapaxes(L, 1, 2) @ L - np.eye(n)).max()
    ldR = np.abs(np.linalg.slogdet(R)[1]).max()
    srt = bool((np.diff(d, axis=1) <= 1e-12 * d[:, :-1]).all())
    algo = sla.api.last_ldr_algo() if hasattr(sla.api, "last_ldr_algo") else "?"
    print(f"n={n} batch={batch} decades={decades}: algo={algo} recon {rec.max():.2e} orth {orth:.2e} log|detR| {ldR:.2e} sorted {srt} d0 {d[0,0]:.6e} dN {d[0,-1]:.6e}")
    ok &= rec.max() < 1e-12 and orth < 1e-12 and ldR < 1e-9
print("BLK_CHECK", "OK" if ok else "BAD")
