"""Observed oracle errors of the sweep stepper at C1 size, strict vs panel pivoting: python profiles/sweep_err.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import sla_b200 as S
import test_gpu_parity as T
from oracle import sweep as osweep
cfg = S.synthetic.CONFIGS["C1"]
chains = [0, 1, 2, 3]
expK, expKinv, fields, out = T.run_sweep(S, cfg, chains)
for i in range(len(chains)):
    ref = osweep.sweep(expK, expKinv, fields[i], cfg.lam, cfg.n_stab)
    print(os.environ.get("SLA_LDR_PIVOT", "strict"), "chain", chains[i], "G0 err %.2e  G err %.2e  wrap max gpu %.2e cpu %.2e" % (
        T.relerr(out["G0"][i], ref["G0"]), T.relerr(out["G"][i], ref["G"]), out["dG"][:, i].max(), ref["dG"].max()))
