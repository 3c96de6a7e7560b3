"""Closed-form known answers used by the reference's test-suite (test/runtests.jl:7-55).

TEST INFRASTRUCTURE ONLY (see sla_oracle.py).  Free fermions on a periodic Lx x Lx square lattice:
G(tau) = U diag(exp(-tau*eps)/(1+exp(-beta*eps))) U^H, B(tau) = U diag(exp(-tau*eps)) U^H.
LatticeUtilities (un-vendored test dependency, test/Project.toml:2) is only used upstream to
enumerate nearest-neighbour bonds; the neighbour table is built directly here.
"""
import numpy as np


def hamiltonian(Lx, t, mu):
    """runtests.jl:7-29 -> (eps, U, H)."""
    N = Lx * Lx
    dtype = np.complex128 if np.iscomplexobj(np.asarray(t)) else np.float64
    H = np.zeros((N, N), dtype=dtype)
    for dx, dy in ((1, 0), (0, 1)):
        for y in range(Lx):
            for x in range(Lx):
                i = x + Lx * y
                j = ((x + dx) % Lx) + Lx * ((y + dy) % Lx)
                H[j, i] = -t
                H[i, j] = np.conj(-t)
    H[np.arange(N), np.arange(N)] = -mu
    eps, U = np.linalg.eigh(H)
    return eps, U, H


def greens(tau, beta, eps, U):
    """runtests.jl:32-41 -> (G(tau), sign det, log|det|)."""
    g = np.exp(-tau * eps) / (1.0 + np.exp(-beta * eps))
    G = (U * g[None, :]) @ U.conj().T
    return G, 1.0, float(np.sum(np.log(g)))


def propagator(tau, eps, U):
    """runtests.jl:44-55 -> (B(tau), log det, sign)."""
    b = np.exp(-tau * eps)
    B = (U * b[None, :]) @ U.conj().T
    return B, float(-tau * np.sum(eps)), 1.0
