"""Small end-to-end run of the newest kernels (hybrid LDR: factor + blocked Q launches, LU register panel, chain C1),
sized for `compute-sanitizer --tool memcheck python profiles/sanitize_small.py` (the tool is closed on the shared B200 pool;
the script doubles as a quick reconstruction check)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
rng = np.random.default_rng(1)
for n, batch in ((200, 3), (256, 2)):
    A = rng.standard_normal((batch, n, n)) * (10.0 ** np.linspace(10, -10, n))[None, None, :]
    dA = sla.to_device(A)
    ws = sla.ldr_workspace(dA)
    F = sla.ldr(dA, ws)
    G = torch.empty_like(dA)
    ld, sg = sla.inv_IpA_(G, F, ws)
    torch.cuda.synchronize()
    L, d, R = sla.to_host(F.L), F.d.cpu().numpy(), sla.to_host(F.R)
    print(n, batch, "recon", np.abs(L @ (d[:, :, None] * R) - A).max() / np.abs(A).max(), "logdet", ld.cpu().numpy()[:2])
cfg = sla.synthetic.CONFIGS["C1"]
expK = sla.to_device(cfg.expK())
fields = torch.from_numpy(cfg.fields(range(2))).cuda()
G, logdet, sign, _ = sla.greens_chain(expK, fields, cfg.lam, cfg.n_stab, cfg.inverse)
torch.cuda.synchronize()
print("chain C1 logdet", logdet.cpu().numpy())
