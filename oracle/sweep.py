"""CPU oracle for one stabilised imaginary-time SWEEP of a DQMC chain: the caller pattern
of the reference's `inv_IpUV!` (src/exported_functions.jl:97-177) over stacks of LDR factorisations (the `ldrs`
containers of src/LDR.jl:197-257), as the SmoQy DQMC drivers use it:

    setup   F_left[g]  = B(beta, tau_g) = B_bar_ng ... B_bar_{g+1}   by  rmul!(F, B_bar)   overloaded_functions.jl:234-253
            G(0)       = inv_IpA!(G, F_left[0])                                            exported_functions.jl:25-71
    sweep   for every slice l:   G <- B_l G B_l^-1        (wrap; where the Monte-Carlo updates of a real run would sit)
            every n_stab slices: F_right <- lmul!(B_bar_g, F_right)                        overloaded_functions.jl:126-145
                                 G' = inv_IpUV!(G', F_right, F_left[g])  = (I + B(tau_g,0) B(beta,tau_g))^-1
                                 dG_g = max|G' - G|  (the wrap error every DQMC code monitors),  G <- G'

Fields are NOT updated (the linear algebra of a sweep is the hot path; acceptance/rejection is the caller's business),
so G(beta) must reproduce G(0) -- the size-independent property the GPU tests use at full size.

TEST INFRASTRUCTURE ONLY (see sla_oracle.py): imported by tests/ and bench.py's cpu legs, never by the product.
"""
import numpy as np

from . import sla_oracle as O
from .chain import group_products


def _identity_ldr(n, dtype):
    return O.ldr(np.zeros((n, n), dtype=dtype, order="F"))


def sweep(expK, expKinv, fields, lam, n_stab):
    """One upward sweep of one chain -> dict(G0, logdet0, sign0, G, logdet[ng], sign[ng], dG[ng]).

    expK / expKinv: exp(-dtau K) and its inverse; fields: int8 [L][N]; B_l = diag(exp(lam s_l)) expK."""
    L, N = fields.shape
    ng = L // n_stab
    Bbars = group_products(expK, fields, lam, n_stab)
    ws = O.ldr_workspace(Bbars[0])
    # ---- setup: the stack of left factorisations, built from the top of the chain downwards
    F = _identity_ldr(N, expK.dtype)
    left = [None] * (ng + 1)
    left[ng] = O.ldr(F)
    for g in range(ng - 1, -1, -1):
        O.rmul_(F, Bbars[g], ws)                         # F <- F * B_bar_{g+1}
        left[g] = O.ldr(F)
    G = np.zeros((N, N), dtype=expK.dtype, order="F")
    logdet0, sign0 = O.inv_IpA_(G, left[0], ws)
    out = {"G0": G.copy(order="F"), "logdet0": logdet0, "sign0": sign0,
           "logdet": np.zeros(ng), "sign": np.zeros(ng, dtype=expK.dtype), "dG": np.zeros(ng)}
    # ---- sweep
    right = _identity_ldr(N, expK.dtype)
    Gs = np.zeros_like(G)
    for g in range(ng):
        for l in range(g * n_stab, (g + 1) * n_stab):
            e = np.exp(lam * fields[l].astype(np.float64))
            G = (e[:, None] * (expK @ G)) @ expKinv / e[None, :]      # B_l G B_l^-1
        O.lmul_(Bbars[g], right, ws)
        out["logdet"][g], out["sign"][g] = O.inv_IpUV_(Gs, right, left[g + 1], ws)
        out["dG"][g] = np.abs(Gs - G).max()
        G = Gs.copy(order="F")
    out["G"] = G
    return out
