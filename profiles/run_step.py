"""Minimal driver for profiling: `python profiles/run_step.py [CONFIG] [CHAINS] [STEPS]` runs STEPS
(default 2: one warm-up + one profiled) evaluations of the named workload through the C ABI."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sla_b200 as sla  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = sla.synthetic.CONFIGS[name]
chains = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.chains
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
expK = sla.to_device(cfg.expK())
fields = torch.from_numpy(cfg.fields(range(chains))).cuda()
ws = None
for _ in range(steps):
    G, logdet, sign, ws = sla.greens_chain(expK, fields, cfg.lam, cfg.n_stab, cfg.inverse, ws=ws)
torch.cuda.synchronize()
print("launches:", sla.load().sla_launch_count(), "logdet[0] =", float(logdet[0]))
