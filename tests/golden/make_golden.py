"""Generate tests/golden/golden_small.npz.

The reference is Julia and cannot run here, and it ships no golden vectors, so
these fixtures come from the two sources that ARE pinned: (1) the closed-form answers of the reference's
own test setup (test/runtests.jl:7-55,59-100) and (2) the CPU oracle (oracle/, itself pinned by (1)) on
small seeded synthetic chains.  They guard against drift of the oracle/LAPACK build and give the GPU
tests a box-independent target.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sla_b200  # noqa: E402
from oracle import analytic as P  # noqa: E402
from oracle import chain as ochain  # noqa: E402

out = {}
# (1) closed form, real and complex hopping
for tag, t in (("real", 1.0), ("cplx", np.exp(0.3j))):
    eps, U, H = P.hamiltonian(4, t, 0.0)
    G0, s0, l0 = P.greens(0.0, 40.0, eps, U)
    Gh, sh, lh = P.greens(20.0, 40.0, eps, U)
    B = P.propagator(0.1, eps, U)[0]
    out[f"analytic_{tag}_Bbar"] = np.linalg.matrix_power(B, 10)
    out[f"analytic_{tag}_G0"], out[f"analytic_{tag}_logdet0"] = G0, l0
    out[f"analytic_{tag}_Gh"], out[f"analytic_{tag}_logdeth"] = Gh, lh
# (2) oracle on seeded synthetic Hubbard chains (fields come from the counter-based hash: nothing else to store)
syn = sla_b200.synthetic
for tag, cfg, inv in (("hub_real", syn.HubbardConfig("g", 4, 40, 10, 3), "IpA"),
                      ("hub_cplx", syn.HubbardConfig("g", 4, 40, 10, 3, dtype=np.complex128, phase=0.3), "IpUV")):
    expK, fields = cfg.expK(), cfg.fields(range(3))
    Gs, lds, sgs = [], [], []
    for c in range(3):
        G, ld, sg, _ = ochain.eval_greens(expK, fields[c], cfg.lam, cfg.n_stab, inv)
        Gs.append(G); lds.append(ld); sgs.append(sg)
    out[f"{tag}_G"], out[f"{tag}_logdet"], out[f"{tag}_sign"] = np.stack(Gs), np.array(lds), np.array(sgs)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
print("wrote", len(out), "arrays")
