"""Run-to-run bit identity of the fused drivers (a data race in the cluster kernels -- split barriers of the Q formation,
the LU inverse's concurrent forward sweep, the side streams of the chain driver -- would show up as differences):
python profiles/determinism_check.py [CONFIG] [REPS]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = sla.synthetic.CONFIGS[name]
chains = min(cfg.chains, 64)
expK = sla.to_device(cfg.expK())
fields = torch.from_numpy(cfg.fields(range(chains))).cuda()
ws = None
ref = None
bad = 0
for it in range(reps):
    G, logdet, sign, ws = sla.greens_chain(expK, fields, cfg.lam, cfg.n_stab, cfg.inverse, ws=ws)
    torch.cuda.synchronize()
    cur = (G.clone(), logdet.clone(), sign.clone())
    if ref is None:
        ref = cur
    else:
        same = all(torch.equal(a.view(torch.uint8) if False else a, b) for a, b in zip(cur, ref))
        bad += (not same)
print(f"{name}: {reps} evaluations of {chains} chains, {bad} differ from the first (bitwise)")
sys.exit(1 if bad else 0)
