"""Synthetic Hubbard-like inputs for the stabilised Green's-function path (host side, numpy only).

Follows SURVEY.md section 8(d): periodic Lx x Lx square lattice, K[j,i] = -t, K[i,j] = -conj(t) on
nearest-neighbour bonds exactly as the reference test builds its Hamiltonian
(test/runtests.jl:7-29), expK = exp(-dtau*K) by eigendecomposition (runtests.jl:26,44-48),
Hubbard-Stratonovich fields s[c,l,i] in {-1,+1} from a counter-based splitmix64 hash so that the
CPU oracle and the GPU see identical fields without shipping them, and
B_{c,l} = diag(exp(lambda*s[c,l,:])) * expK with lambda = acosh(exp(dtau*U/2)).
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EED


def square_lattice_K(Lx: int, t=1.0, mu: float = 0.0) -> np.ndarray:
    """Hopping matrix of the periodic Lx x Lx square lattice (test/runtests.jl:7-24).

    Bond n runs from site i to j = i + x_hat (then i + y_hat); ``K[j,i] = -t``,
    ``K[i,j] = conj(-t)``, ``K[i,i] = -mu``.  Complex ``t`` gives a ComplexF64 matrix."""
    N = Lx * Lx
    dtype = np.complex128 if np.iscomplexobj(np.asarray(t)) else np.float64
    K = np.zeros((N, N), dtype=dtype)
    for dx, dy in ((1, 0), (0, 1)):
        for y in range(Lx):
            for x in range(Lx):
                i = x + Lx * y
                j = ((x + dx) % Lx) + Lx * ((y + dy) % Lx)
                K[j, i] = -t
                K[i, j] = np.conj(-t)
    K[np.arange(N), np.arange(N)] = -mu
    return K


def exp_minus_dtau_K(K: np.ndarray, dtau: float) -> np.ndarray:
    """exp(-dtau*K) for Hermitian K through its eigendecomposition (runtests.jl:44-48)."""
    eps, U = np.linalg.eigh(K)
    return (U * np.exp(-dtau * eps)[None, :]) @ U.conj().T


def hs_lambda(U: float, dtau: float) -> float:
    return float(np.arccosh(np.exp(dtau * U / 2.0)))


def hs_fields(chain_ids, L: int, N: int, seed: int = SEED) -> np.ndarray:
    """s[c, l, i] in {-1,+1} (int8) = top bit of splitmix64(seed, counter), with counter
    ((chain*L + l)*N + i).  ``chain_ids`` are *global* chain indices so that sharding the batch
    over ranks does not change any chain's fields."""
    c = np.asarray(chain_ids, dtype=np.uint64).reshape(-1, 1, 1)
    l = np.arange(L, dtype=np.uint64).reshape(1, -1, 1)
    i = np.arange(N, dtype=np.uint64).reshape(1, 1, -1)
    with np.errstate(over="ignore"):
        idx = (c * np.uint64(L) + l) * np.uint64(N) + i
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where((z >> np.uint64(63)) != 0, 1, -1).astype(np.int8)


class HubbardConfig:
    """One named workload (BASELINE.json ``configs``)."""

    def __init__(self, name, Lx, L, n_stab, chains, dtype=np.float64, U=4.0, dtau=0.1, mu=0.0,
                 phase=0.0, inverse="IpA"):
        self.name, self.Lx, self.N = name, Lx, Lx * Lx
        self.L, self.n_stab, self.chains = L, n_stab, chains
        self.dtype = np.dtype(dtype)
        self.U, self.dtau, self.mu, self.phase = U, dtau, mu, phase
        self.inverse = inverse
        assert L % n_stab == 0

    @property
    def t(self):
        return np.exp(1j * self.phase) if self.dtype.kind == "c" else 1.0

    def expK(self) -> np.ndarray:
        return np.asarray(exp_minus_dtau_K(square_lattice_K(self.Lx, self.t, self.mu), self.dtau),
                          dtype=self.dtype)

    @property
    def lam(self) -> float:
        return hs_lambda(self.U, self.dtau)

    def fields(self, chain_ids) -> np.ndarray:
        return hs_fields(chain_ids, self.L, self.N)

    def flops_per_eval(self) -> float:
        """Algorithmic FLOPs of one evaluation."""
        c = 4.0 if self.dtype.kind == "c" else 1.0
        inv = 6.0 if self.inverse == "IpA" else 38.0 / 3.0
        return c * self.N ** 3 * (2.0 * self.L + (14.0 / 3.0) * (self.L / self.n_stab) + inv)


CONFIGS = {
    "C1": HubbardConfig("C1", 8, 100, 10, 1),
    "C2": HubbardConfig("C2", 16, 200, 10, 64),
    "C3": HubbardConfig("C3", 16, 200, 10, 128, dtype=np.complex128, phase=0.3, inverse="IpUV"),
    "C4": HubbardConfig("C4", 32, 400, 10, 32),
    "C5": HubbardConfig("C5", 20, 200, 10, 4096),
}
