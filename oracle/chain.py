"""CPU oracle for one stabilised G(tau=0) evaluation (the driver shape of test/runtests.jl:93-100,
151-160 applied to Hubbard-like B matrices).

TEST INFRASTRUCTURE ONLY (see sla_oracle.py).
"""
import numpy as np

from . import sla_oracle as O


def group_products(expK, fields, lam, n_stab):
    """B_bar[g] = B_{g*n} ... B_{(g-1)n+1}, built by left-multiplying in slice order exactly like
    runtests.jl:96-97 (``mul!(A, B, B_bar); copyto!(B_bar, A)``), B_l = diag(exp(lam*s_l)) expK."""
    L, N = fields.shape
    out = []
    for g in range(L // n_stab):
        Bbar = np.eye(N, dtype=expK.dtype)
        for l in range(g * n_stab, (g + 1) * n_stab):
            B = np.exp(lam * fields[l].astype(np.float64))[:, None] * expK
            Bbar = B @ Bbar
        out.append(np.asfortranarray(Bbar))
    return out


def eval_greens(expK, fields, lam, n_stab, inverse="IpA"):
    """One evaluation -> (G, log|det G|, sign det G, F) for a single chain.

    inverse="IpA":  F <- I; lmul!(B_bar_g, F) for every group; inv_IpA!(G, F)   (runtests.jl:152-157)
    inverse="IpUV": first half of the groups into V, second half into U (both by lmul!),
                    then inv_IpUV!(G, U, V)  (G = (I + U V)^-1, U V = B_L ... B_1)."""
    N = expK.shape[0]
    Bbars = group_products(expK, fields, lam, n_stab)
    ws = O.ldr_workspace(Bbars[0])
    G = np.zeros((N, N), dtype=expK.dtype, order="F")
    if inverse == "IpA":
        F = O.ldr(Bbars[0])
        for Bbar in Bbars:
            O.lmul_(Bbar, F, ws)
        logdet, sgn = O.inv_IpA_(G, F, ws)
        return G, logdet, sgn, F
    half = len(Bbars) // 2
    V = O.ldr(Bbars[0])
    U = O.ldr(Bbars[0])
    for Bbar in Bbars[:half]:
        O.lmul_(Bbar, V, ws)
    for Bbar in Bbars[half:]:
        O.lmul_(Bbar, U, ws)
    logdet, sgn = O.inv_IpUV_(G, U, V, ws)
    return G, logdet, sgn, (U, V)
