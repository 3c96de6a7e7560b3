import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import sla_b200 as sla
for n, batch in ((256, 8), (200, 5), (130, 3), (255, 4)):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((batch, n, n)) * (10.0 ** rng.uniform(-20, 20, size=(batch, 1, n)))
    dA = sla.to_device(A)
    F = sla.ldr(dA); ws = sla.ldr_workspace(dA)
    sla.ldr_(F, dA, ws); torch.cuda.synchronize()
    L, d, R = sla.to_host(F.L), F.d.cpu().numpy(), sla.to_host(F.R)
    rec = L @ (d[:, :, None] * R)
    err = np.abs(rec - A).max(axis=(1, 2)) / np.abs(A).max(axis=(1, 2))
    orth = np.abs(np.einsum('bji,bjk->bik', L, L) - np.eye(n)).max()
    print(f"n={n} batch={batch}: recon rel err {err.max():.2e}  orth {orth:.2e}  d sorted desc: {bool((np.diff(d, axis=1) <= 1e-12 * d[:, :-1]).all())}")
