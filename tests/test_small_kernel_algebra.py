"""CPU restatements of the index algebra of the small-matrix kernels (csrc/sla_lu_small.cu, csrc/sla_ldr_small.cu): the
in-place Gauss-Jordan with implicit partial pivoting must return A^-1 through  A^-1[k, p_c] = slot_c[p_k], its sign through
the inversion count of k -> p_k, and the reflector application in its raw-dot form  v^H a = a_k + conj(scal) sum conj(x_r) a_r
must equal the textbook  H^H a = a - conj(tau) v (v^H a).  The kernels themselves are tested on the GPU
(tests/test_gpu_parity.py: test_inv_lu_small_kernel, test_ldr_small_kernel); this file pins the formulas they implement."""
import numpy as np
import pytest


def gauss_jordan_implicit(A):
    """Step k: pivot = largest |a_rk| among rows not used yet (first index on ties), column k is replaced by the column the
    inverse gains; returns (A^-1, log|det A|, sign det A) assembled the way the kernel's epilogue does."""
    n = A.shape[0]
    a = A.astype(A.dtype, copy=True)
    used = np.zeros(n, bool)
    prow = np.zeros(n, int)
    kofr = np.zeros(n, int)
    piv = np.zeros(n, A.dtype)
    for k in range(n):
        key = np.where(used, -2.0, np.abs(a[:, k].real) + np.abs(a[:, k].imag))
        pr = int(np.argmax(key))
        pv = a[pr, k]
        rc = 1.0 / pv
        col = a[:, k].copy()
        for c in range(n):
            if c == k:
                continue
            apc = a[pr, c]
            t = apc * rc
            a[:, c] -= col * t
            a[pr, c] = t
        a[:, k] = col * (-rc)
        a[pr, k] = rc
        used[pr] = True
        prow[k], kofr[pr], piv[k] = pr, k, pv
    X = np.zeros_like(a)
    for c in range(n):
        for r in range(n):
            X[kofr[r], prow[c]] = a[r, c]
    inversions = sum(1 for k in range(n) for j in range(k + 1, n) if prow[j] < prow[k])
    sign = np.prod(piv / np.abs(piv)) * (-1.0) ** inversions
    return X, float(np.sum(np.log(np.abs(piv)))), sign


@pytest.mark.parametrize("n", [1, 2, 3, 7, 33, 64])
@pytest.mark.parametrize("complex_", [False, True])
def test_gauss_jordan_implicit_pivoting(n, complex_):
    rng = np.random.default_rng(100 * n + complex_)
    A = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if complex_ else 0.0)
    X, logdet, sign = gauss_jordan_implicit(A)
    sref, lref = np.linalg.slogdet(A)
    assert np.max(np.abs(X - np.linalg.inv(A))) < 1e-12 * np.linalg.cond(A)
    assert abs(logdet - lref) < 1e-11 and abs(sign - sref) < 1e-12


def test_gauss_jordan_ties_and_permutations():
    rng = np.random.default_rng(5)
    P = np.eye(9)[rng.permutation(9)]
    for M in (np.eye(5), P, -3.0 * P, np.sign(rng.standard_normal((12, 12)))):
        X, logdet, sign = gauss_jordan_implicit(M)
        sref, lref = np.linalg.slogdet(M)
        assert np.max(np.abs(X - np.linalg.inv(M))) < 1e-12 * np.linalg.cond(M)
        assert abs(logdet - lref) < 1e-12 and sign == sref


@pytest.mark.parametrize("complex_", [False, True])
def test_reflector_raw_dot_form(complex_):
    """?larfg scalars (beta, tau, scal = 1 / (alpha - beta)) and H^H a in the form the small LDR kernel evaluates."""
    rng = np.random.default_rng(3 + complex_)
    n, k = 11, 4
    col = rng.standard_normal(n) + (1j * rng.standard_normal(n) if complex_ else 0.0)    # pivot column, rows >= k live
    a = rng.standard_normal(n) + (1j * rng.standard_normal(n) if complex_ else 0.0)      # another column
    alpha, x = col[k], col[k + 1:]
    xn2 = float(np.sum(np.abs(x) ** 2))
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = (beta - alpha.real) / beta - 1j * (alpha.imag / beta)      # real input: imaginary part 0
    scal = 1.0 / (alpha - beta)
    v = np.concatenate([[1.0], x * scal])
    want = a[k:] - np.conj(tau) * v * np.vdot(v, a[k:])                                  # H^H a, H = I - tau v v^H
    s = np.sum(np.conj(x) * a[k + 1:])                                                   # raw dot: needs no scalar
    w = np.conj(tau) * (a[k] + np.conj(scal) * s)
    got = a[k:].astype(complex)
    got[1:] -= x * (scal * w)
    got[0] -= w
    assert np.max(np.abs(got - want)) < 1e-13 * np.max(np.abs(want))
    # and H^H annihilates the pivot column below the diagonal, leaving beta on it
    hcol = col[k:] - np.conj(tau) * v * np.vdot(v, col[k:])
    assert abs(hcol[0] - beta) < 1e-13 * abs(beta) and np.max(np.abs(hcol[1:])) < 1e-13 * abs(beta)
