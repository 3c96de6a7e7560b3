"""CPU oracle: a restatement of StableLinearAlgebra.jl (v1.5.1) over the LAPACK routines it calls.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker / the CPU arm -- never on the CUDA path.

What it is
----------
The reference is pure Julia whose arithmetic is done by an un-vendored third-party library: LAPACK
and BLAS reached through Julia's ``LinearAlgebra`` stdlib (``src/StableLinearAlgebra.jl:13-17``;
version unpinned, ``Project.toml:9-10``).  Julia is not installed in this image, so the reference
cannot be executed here.  This file restates every reference function on the hot path *line by
line* over the same LAPACK entry points (``?geqp3``, ``?orgqr/?ungqr``, ``?getrf``, ``?getri``,
``?getrs``, ``?gemm``) as shipped in scipy's OpenBLAS (0.3.3x), for Float64 and ComplexF64.

How it is pinned
----------------
By the closed-form known answers of the reference's own test-suite (``test/runtests.jl:7-55,
151-283``): the free-fermion Green's function of the 4x4 periodic tight-binding model at beta=40.
``tests/test_oracle.py`` replays all 45 executed reference assertions (plus a complex-hopping
variant the reference does not test -- ComplexF64 is *unpinned by the reference*, pinned here only
by the same analytic construction).  No golden files exist upstream.

Conventions
-----------
Matrices are numpy arrays (any memory order; LAPACK calls get Fortran-ordered copies).  ``jpvt`` is
1-based as in Fortran/Julia (``qr.jl:14``); scipy's ``getrf`` returns 0-based ``piv`` which is
converted to 1-based ``ipiv`` to keep ``det_lu!``'s comparison ``i != ipiv[i]`` literal.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack as _lp

__all__ = [
    "QRWorkspace", "LUWorkspace", "LDR", "LDRWorkspace",
    "geqp3_", "orgqr_", "getrf_", "getri_", "getrs_", "det_lu_", "inv_lu_", "ldiv_lu_", "mul_P_",
    "ldr", "ldr_", "ldrs", "ldrs_", "ldr_workspace",
    "copyto_", "adjoint_", "lmul_", "rmul_", "mul_", "ldiv_", "rdiv_", "logabsdet",
    "inv_IpA_", "inv_IpUV_", "inv_UpV_", "inv_invUpV_",
]


def _is_complex(A):
    return np.iscomplexobj(A)


# --------------------------------------------------------------------------------------------
# L1: LAPACK wrappers  (src/qr.jl, src/lu.jl)
# --------------------------------------------------------------------------------------------
class QRWorkspace:
    """``QRWorkspace{T,E}`` -- src/qr.jl:9-15.  ``work`` length follows the workspace query at
    src/qr.jl:36-67, which for geqp3 returns ``2n + (n+1)*NB`` with NB=32."""

    def __init__(self, A):
        n = A.shape[0]
        assert A.shape == (n, n)
        self.n = n
        self.dtype = A.dtype
        self.lwork = 2 * n + (n + 1) * 32
        self.tau = np.zeros(n, dtype=A.dtype)
        self.jpvt = np.zeros(n, dtype=np.int64)


class LUWorkspace:
    """``LUWorkspace{T}`` -- src/lu.jl:10-14."""

    def __init__(self, A):
        n = A.shape[0]
        self.n = n
        self.dtype = A.dtype
        self.ipiv = np.zeros(n, dtype=np.int64)
        self.lwork = max(1, 64 * n)


def geqp3_(A, ws: QRWorkspace):
    """``LAPACK.geqp3!(A, ws)`` -- src/qr.jl:70-92.  jpvt is zeroed first (qr.jl:77) so every
    column is free; A is overwritten with the reflectors (below diag) and R' (on/above)."""
    f = _lp.zgeqp3 if _is_complex(A) else _lp.dgeqp3
    qr, jpvt, tau, _work, info = f(np.asfortranarray(A), lwork=ws.lwork, overwrite_a=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"geqp3 info={info}")
    A[...] = qr
    ws.tau[...] = tau
    ws.jpvt[...] = jpvt  # 1-based
    return None


def orgqr_(A, ws: QRWorkspace):
    """``LAPACK.orgqr!(A, ws)`` -- src/qr.jl:95-109 (``?orgqr`` real / ``?ungqr`` complex)."""
    f = _lp.zungqr if _is_complex(A) else _lp.dorgqr
    q, _work, info = f(np.asfortranarray(A), ws.tau, lwork=ws.lwork, overwrite_a=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"orgqr info={info}")
    A[...] = q
    return None


def getrf_(A, ws: LUWorkspace):
    """``LAPACK.getrf!(A, ws)`` -- src/lu.jl:63-77."""
    f = _lp.zgetrf if _is_complex(A) else _lp.dgetrf
    lu, piv, info = f(np.asfortranarray(A), overwrite_a=0)
    if info < 0:
        raise ValueError(f"getrf info={info}")
    if info > 0:
        raise np.linalg.LinAlgError(f"getrf: singular U({info},{info})")
    A[...] = lu
    ws.ipiv[...] = piv + 1  # to 1-based like Julia/Fortran
    return None


def getri_(A, ws: LUWorkspace):
    """``LAPACK.getri!(A, ws)`` -- src/lu.jl:80-93."""
    f = _lp.zgetri if _is_complex(A) else _lp.dgetri
    inv, info = f(np.asfortranarray(A), (ws.ipiv - 1).astype(np.int32), lwork=ws.lwork, overwrite_lu=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"getri info={info}")
    A[...] = inv
    return None


def getrs_(A, B, ws: LUWorkspace, trans="N"):
    """``LAPACK.getrs!(A, B, ws, trans)`` -- src/lu.jl:96-113."""
    f = _lp.zgetrs if _is_complex(A) else _lp.dgetrs
    t = {"N": 0, "T": 1, "C": 2}[trans]
    x, info = f(np.asfortranarray(A), (ws.ipiv - 1).astype(np.int32), np.asfortranarray(B), trans=t)
    if info != 0:
        raise np.linalg.LinAlgError(f"getrs info={info}")
    B[...] = x
    return None


def _sign(z):
    # Julia sign(z) = z/abs(z) for complex, +-1 for real (0 -> 0)
    if z == 0:
        return z
    return z / abs(z)


def det_lu_(A, ws: LUWorkspace):
    """``det_lu!(A, ws)`` -- src/lu.jl:123-140.  Returns (log|det A|, sign det A); A := LU."""
    getrf_(A, ws)
    logdetA = 0.0
    sgndetA = A.dtype.type(1)
    for i in range(A.shape[0]):
        logdetA += np.log(abs(A[i, i]))
        sgndetA *= _sign(A[i, i])
        if (i + 1) != ws.ipiv[i]:
            sgndetA = -sgndetA
    return float(np.real(logdetA)), sgndetA


def inv_lu_(A, ws: LUWorkspace):
    """``inv_lu!(A, ws)`` -- src/lu.jl:149-158.  A := A^-1; returns (log|det A^-1|, sign det A^-1)."""
    logdetA, sgndetA = det_lu_(A, ws)
    getri_(A, ws)
    return float(-logdetA), np.conj(sgndetA)


def ldiv_lu_(A, B, ws: LUWorkspace):
    """``ldiv_lu!(A, B, ws)`` -- src/lu.jl:167-176.  B := A^-1 B, A := LU."""
    logdetA, sgndetA = det_lu_(A, ws)
    getrs_(A, B, ws)
    return float(-logdetA), np.conj(sgndetA)


def mul_P_(A, B, p):
    """``mul_P!(A, B, p)``: A[:, p] = B  -- src/developer_functions.jl:20-26 (p 1-based)."""
    A[:, np.asarray(p) - 1] = B
    return None


# --------------------------------------------------------------------------------------------
# L2: LDR type + factorisation  (src/LDR.jl)
# --------------------------------------------------------------------------------------------
class LDR:
    """``LDR{T,E}`` -- src/LDR.jl:23-33: A = L * Diagonal(d) * R."""

    def __init__(self, L, d, R):
        self.L, self.d, self.R = L, d, R

    @property
    def dtype(self):  # eltype, overloaded_functions.jl:10
        return self.L.dtype

    @property
    def shape(self):  # size, overloaded_functions.jl:22
        return self.L.shape


class LDRWorkspace:
    """``LDRWorkspace{T,E}`` -- src/LDR.jl:51-70 (built by ``ldr_workspace``, LDR.jl:267-282)."""

    def __init__(self, A):
        n = A.shape[0]
        T = A.dtype
        E = np.float64
        self.M = np.zeros((n, n), dtype=T, order="F")
        self.Mp = np.zeros((n, n), dtype=T, order="F")   # M'
        self.Mpp = np.zeros((n, n), dtype=T, order="F")  # M''
        self.v = np.zeros(n, dtype=E)
        self.qr_ws = QRWorkspace(self.M)
        self.lu_ws = LUWorkspace(self.M)


def ldr_workspace(A):
    """``ldr_workspace(A)`` / ``ldr_workspace(F)`` -- src/LDR.jl:267-282."""
    if isinstance(A, LDR):
        A = A.L
    return LDRWorkspace(A)


def _ldr_identity_(F: LDR):
    """``ldr!(F, I)`` -- src/LDR.jl:133-140."""
    n = F.L.shape[0]
    F.L[...] = np.eye(n, dtype=F.L.dtype)
    F.d[...] = 1
    F.R[...] = np.eye(n, dtype=F.R.dtype)


def ldr(A, ws: LDRWorkspace | None = None):
    """``ldr(A)`` (identity-initialised, LDR.jl:79-90), ``ldr(A, ws)`` (factorise, LDR.jl:97-102),
    ``ldr(F)`` (copy, LDR.jl:108-114)."""
    if isinstance(A, LDR):
        return LDR(A.L.copy(order="F"), A.d.copy(), A.R.copy(order="F"))
    n = A.shape[0]
    T = A.dtype if np.issubdtype(A.dtype, np.inexact) else np.float64
    F = LDR(np.zeros((n, n), dtype=T, order="F"), np.zeros(n, dtype=np.float64),
            np.zeros((n, n), dtype=T, order="F"))
    _ldr_identity_(F)
    if ws is not None:
        ldr_(F, A, ws)
    return F


def ldr_(F: LDR, A=None, ws: LDRWorkspace | None = None):
    """``ldr!`` -- four methods, src/LDR.jl:122-188:
    ``ldr_(F, A, ws)`` copy A into F.L then factorise (LDR.jl:122-126);
    ``ldr_(F, "I")`` identity (133-140); ``ldr_(Fout, Fin)`` copy (147-153);
    ``ldr_(F, ws=ws)`` factorise F.L in place (160-188)."""
    if isinstance(A, LDRWorkspace) and ws is None:
        A, ws = None, A
    if isinstance(A, str) and A == "I":
        _ldr_identity_(F)
        return None
    if isinstance(A, LDR):
        F.L[...] = A.L
        F.d[...] = A.d
        F.R[...] = A.R
        return None
    if A is not None:
        F.L[...] = A
    qr_ws, M = ws.qr_ws, ws.M
    L, d, R = F.L, F.d, F.R
    # LDR.jl:166  column-pivoted QR in place
    geqp3_(L, qr_ws)
    # LDR.jl:169-170  M = triu(L)
    M[...] = np.triu(L)
    # LDR.jl:173-175  d = |diag|
    d[...] = np.abs(np.diagonal(M))
    # LDR.jl:178-179  M = D^-1 M
    M[...] = M / d[:, None]
    # LDR.jl:182  R[:, jpvt] = M
    mul_P_(R, M, qr_ws.jpvt)
    # LDR.jl:185  L = Q
    orgqr_(L, qr_ws)
    return None


def ldrs(A, N=None, ws: LDRWorkspace | None = None):
    """``ldrs`` -- src/LDR.jl:197-241.  ``ldrs(A, N)`` N identities; ``ldrs(A, N, ws)`` N copies of
    ldr(A); ``ldrs(A3, ws=ws)`` one factorisation per ``A3[:, :, i]``."""
    if isinstance(N, LDRWorkspace):
        N, ws = None, N
    if A.ndim == 3:
        return [ldr(A[:, :, i], ws) for i in range(A.shape[2])]
    if ws is None:
        return [ldr(A) for _ in range(N)]
    F = ldr(A, ws)
    return [F] + [ldr(F) for _ in range(1, N)]


def ldrs_(Fs, A3, ws: LDRWorkspace):
    """``ldrs!(Fs, A3, ws)`` -- src/LDR.jl:249-257."""
    for i in range(len(Fs)):
        ldr_(Fs[i], A3[:, :, i], ws)
    return None


# --------------------------------------------------------------------------------------------
# L3: overloaded functions  (src/overloaded_functions.jl)
# --------------------------------------------------------------------------------------------
def copyto_(U, V, ws: LDRWorkspace | None = None):
    """``copyto!`` -- overloaded_functions.jl:34-63.
    (Matrix, LDR, ws): U = L*D*R (34-45); (LDR, Matrix, ws): factorise (55);
    (LDR, "I"): identity (56); (LDR, LDR): copy (63)."""
    if isinstance(U, LDR):
        return ldr_(U, V, ws)
    M = ws.M
    M[...] = V.R
    M[...] = V.d[:, None] * M          # lmul!(D, M)
    U[...] = V.L @ M                   # mul!(U, L, M)
    return None


def adjoint_(At, A: LDR, ws: LDRWorkspace):
    """``adjoint!(A', A::LDR, ws)`` -- overloaded_functions.jl:75-86."""
    ws.Mp[...] = A.R.conj().T
    ws.M[...] = A.L.conj().T
    ws.M[...] = A.d[:, None] * ws.M
    At[...] = ws.Mp @ ws.M
    return None


def lmul_(U, V, ws: LDRWorkspace):
    """``lmul!`` -- overloaded_functions.jl:97-193.
    (LDR, Matrix): V := L D R V (97-106, not stabilised);
    (Matrix, LDR): V := U V stabilised (126-145);
    (LDR, LDR):    V := U V stabilised (166-193)."""
    if isinstance(U, LDR) and not isinstance(V, LDR):
        ws.M[...] = U.R @ V
        ws.M[...] = U.d[:, None] * ws.M
        V[...] = U.L @ ws.M
        return None
    if not isinstance(U, LDR):
        Rv = ws.Mp
        Rv[...] = V.R                               # :129-130
        ws.M[...] = U @ V.L                         # :133
        V.L[...] = ws.M * V.d[None, :]              # :134-135
        ldr_(V, ws=ws)                              # :138
        ws.M[...] = V.R @ Rv                        # :141
        V.R[...] = ws.M                             # :142
        return None
    Rv = ws.Mp
    Rv[...] = V.R                                   # :169-170
    ws.M[...] = U.R @ V.L                           # :173
    ws.M[...] = ws.M * V.d[None, :]                 # :178
    V.L[...] = U.d[:, None] * ws.M                  # :179
    ldr_(V, ws=ws)                                  # :182
    ws.M[...] = U.L @ V.L                           # :185
    V.L[...] = ws.M
    ws.M[...] = V.R @ Rv                            # :189
    V.R[...] = ws.M
    return None


def rmul_(U, V, ws: LDRWorkspace):
    """``rmul!`` -- overloaded_functions.jl:205-301.
    (Matrix, LDR): U := U L D R (205-214); (LDR, Matrix): U := U V stabilised (234-253);
    (LDR, LDR): U := U V stabilised (274-301)."""
    if not isinstance(U, LDR):
        ws.M[...] = U @ V.L
        ws.M[...] = ws.M * V.d[None, :]
        U[...] = ws.M @ V.R
        return None
    Lu = ws.Mp
    Lu[...] = U.L
    if not isinstance(V, LDR):
        U.L[...] = U.R @ V                          # :241
        U.L[...] = U.d[:, None] * U.L               # :243
        ldr_(U, ws=ws)                              # :246
        ws.M[...] = Lu @ U.L                        # :249
        U.L[...] = ws.M
        return None
    ws.M[...] = U.R @ V.L                           # :281
    ws.M[...] = ws.M * V.d[None, :]                 # :286
    U.L[...] = U.d[:, None] * ws.M                  # :287
    ldr_(U, ws=ws)                                  # :290
    ws.M[...] = Lu @ U.L                            # :293
    U.L[...] = ws.M
    ws.M[...] = U.R @ V.R                           # :297
    U.R[...] = ws.M
    return None


def mul_(H, U, V, ws: LDRWorkspace):
    """``mul!`` x5 -- overloaded_functions.jl:314-376 (= copyto! + lmul!/rmul!)."""
    if not isinstance(H, LDR):
        if isinstance(U, LDR):              # :314-318  H = U*V, V matrix
            H[...] = V
            return lmul_(U, H, ws)
        H[...] = U                          # :326-330  H = U*V, V LDR
        return rmul_(H, V, ws)
    if isinstance(U, LDR) and not isinstance(V, LDR):   # :351-355
        ldr_(H, U)
        return rmul_(H, V, ws)
    ldr_(H, V)                              # :338-342, :364-368
    return lmul_(U, H, ws)


def ldiv_(*args):
    """``ldiv!`` x6 -- overloaded_functions.jl:388-538 ("next" row)."""
    if len(args) == 4:
        H, U, V, ws = args
        if isinstance(H, LDR):
            ldr_(H, V)
        else:
            H[...] = V
        return ldiv_(U, H, ws)
    U, V, ws = args
    if isinstance(U, LDR) and not isinstance(V, LDR):       # :388-400
        Lu = ws.M
        Lu[...] = U.L
        ldiv_lu_(Lu, V, ws.lu_ws)
        V[...] = V / U.d[:, None]
        Ru = ws.M
        Ru[...] = U.R
        ldiv_lu_(Ru, V, ws.lu_ws)
        return None
    if isinstance(U, LDR):                                  # :437-464
        ws.M[...] = U.L.conj().T @ V.L
        V.L[...] = ws.M
        Rv = ws.Mp
        Rv[...] = V.R
        V.L[...] = V.L * V.d[None, :]
        V.L[...] = V.L / U.d[:, None]
        Ru = ws.M
        Ru[...] = U.R
        ldiv_lu_(Ru, V.L, ws.lu_ws)
        ldr_(V, ws=ws)
        ws.M[...] = V.R @ Rv
        V.R[...] = ws.M
        return None
    Rv = ws.Mp                                              # :502-521
    Rv[...] = V.R
    V.L[...] = V.L * V.d[None, :]
    ws.M[...] = U
    ldiv_lu_(ws.M, V.L, ws.lu_ws)
    ldr_(V, ws=ws)
    ws.M[...] = V.R @ Rv
    V.R[...] = ws.M
    return None


def rdiv_(*args):
    """``rdiv!`` x6 -- overloaded_functions.jl:549-709 ("next" row)."""
    if len(args) == 4:
        H, U, V, ws = args
        if isinstance(H, LDR):
            ldr_(H, U)
        else:
            H[...] = U
        return rdiv_(H, V, ws)
    U, V, ws = args
    n = ws.M.shape[0]
    if not isinstance(U, LDR):                              # :549-565
        ws.M[...] = np.eye(n, dtype=ws.M.dtype)
        Lv = ws.Mp
        Lv[...] = V.L
        ldiv_lu_(Lv, ws.M, ws.lu_ws)
        ws.M[...] = ws.M / V.d[:, None]
        Rv = ws.Mp
        Rv[...] = V.R
        ldiv_lu_(Rv, ws.M, ws.lu_ws)
        ws.Mp[...] = U @ ws.M
        U[...] = ws.Mp
        return None
    if isinstance(V, LDR):                                  # :603-637
        Rvi = ws.Mp
        Rvi[...] = V.R
        inv_lu_(Rvi, ws.lu_ws)
        ws.M[...] = U.R @ Rvi
        Lu = ws.Mp
        Lu[...] = U.L
        U.L[...] = U.d[:, None] * ws.M
        U.L[...] = U.L / V.d[None, :]
        ldr_(U, ws=ws)
        ws.M[...] = Lu @ U.L
        U.L[...] = ws.M
        ws.M[...] = U.R @ V.L.conj().T
        U.R[...] = ws.M
        return None
    Lu = ws.Mp                                              # :673-693
    Lu[...] = U.L
    ws.M[...] = V
    inv_lu_(ws.M, ws.lu_ws)
    U.L[...] = U.R @ ws.M
    U.L[...] = U.d[:, None] * U.L
    ldr_(U, ws=ws)
    ws.M[...] = Lu @ U.L
    U.L[...] = ws.M
    return None


def logabsdet(A: LDR, ws: LDRWorkspace):
    """``logabsdet(A::LDR, ws)`` -- overloaded_functions.jl:721-733."""
    ws.M[...] = A.L
    logdetL, sgndetL = det_lu_(ws.M, ws.lu_ws)
    logdetD, sgndetD = float(np.sum(np.log(A.d))), A.L.dtype.type(1)
    ws.M[...] = A.R
    logdetR, sgndetR = det_lu_(ws.M, ws.lu_ws)
    return logdetL + logdetD + logdetR, sgndetL * sgndetD * sgndetR


# --------------------------------------------------------------------------------------------
# L3: exported stable inverses  (src/exported_functions.jl)
# --------------------------------------------------------------------------------------------
def inv_IpA_(G, A: LDR, ws: LDRWorkspace):
    """``inv_IpA!(G, A, ws)`` -- exported_functions.jl:25-71.  G := (I + A)^-1."""
    La, da, Ra = A.L, A.d, A.R
    Rai = ws.Mp
    Rai[...] = Ra                                           # :32-33
    _logdetRai, sgndetRai = inv_lu_(Rai, ws.lu_ws)          # :34
    ws.v[...] = np.minimum(da, 1.0)                         # :37-38
    ws.M[...] = La * ws.v[None, :]                          # :41-42
    ws.v[...] = np.maximum(da, 1.0)                         # :45-46
    logdetDp, sgndetDp = float(np.sum(np.log(ws.v))), 1.0   # :49-50
    Rai[...] = Rai / ws.v[None, :]                          # :53-54
    ws.M[...] = ws.M + Rai                                  # :57
    logdetMi, sgndetMi = inv_lu_(ws.M, ws.lu_ws)            # :60-61
    G[...] = Rai @ ws.M                                     # :64
    sgndetG = sgndetRai * np.conj(sgndetDp) * sgndetMi      # :67
    logdetG = -logdetDp + logdetMi                          # :68
    return float(np.real(logdetG)), sgndetG


def inv_IpUV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_IpUV!(G, U, V, ws)`` -- exported_functions.jl:97-177.  G := (I + U V)^-1."""
    Lu, du, Ru = U.L, U.d, U.R
    Lv, dv, Rv = V.L, V.d, V.R
    ws.M[...] = Lu                                          # :107
    _logdetLu, sgndetLu = det_lu_(ws.M, ws.lu_ws)           # :108
    Rvi = ws.Mp
    Rvi[...] = Rv                                           # :111-112
    _l, sgndetRvi = inv_lu_(Rvi, ws.lu_ws)                  # :113
    dvp = np.maximum(dv, 1.0)                               # :116-118
    logdetDvp, sgndetDvp = float(np.sum(np.log(dvp))), 1.0  # :121
    Rvi[...] = Rvi / dvp[None, :]                           # :124
    dup = np.maximum(du, 1.0)                               # :128-130
    logdetDup, sgndetDup = float(np.sum(np.log(dup))), 1.0  # :133
    ws.M[...] = Lu.conj().T                                 # :136
    ws.M[...] = ws.M / dup[:, None]                         # :137
    dum = np.minimum(du, 1.0)                               # :141-143
    G[...] = Ru @ Lv                                        # :146
    G[...] = dum[:, None] * G                               # :147
    dvm = np.minimum(dv, 1.0)                               # :150-152
    G[...] = G * dvm[None, :]                               # :155
    ws.Mpp[...] = ws.M @ Rvi                                # :158
    G[...] = G + ws.Mpp                                     # :161-162
    logdetMi, sgndetMi = inv_lu_(G, ws.lu_ws)               # :165-166
    ws.Mpp[...] = G @ ws.M                                  # :169
    G[...] = Rvi @ ws.Mpp                                   # :170
    sgndetG = sgndetRvi * np.conj(sgndetDvp) * sgndetMi * np.conj(sgndetDup) * np.conj(sgndetLu)
    logdetG = -logdetDvp + logdetMi - logdetDup             # :173-174
    return float(np.real(logdetG)), sgndetG


def inv_UpV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_UpV!(G, U, V, ws)`` -- exported_functions.jl:211-289.  G := (U + V)^-1."""
    Lu, du, Ru = U.L, U.d, U.R
    Lv, dv, Rv = V.L, V.d, V.R
    ws.M[...] = Lu                                          # :221
    _logdetLu, sgndetLu = det_lu_(ws.M, ws.lu_ws)           # :222
    Rvi = ws.Mp
    Rvi[...] = Rv                                           # :225-226
    _l, sgndetRvi = inv_lu_(Rvi, ws.lu_ws)                  # :227
    dvp = np.maximum(dv, 1.0)
    logdetDvp, sgndetDvp = float(np.sum(np.log(dvp))), 1.0  # :235
    Rvi[...] = Rvi / dvp[None, :]                           # :238
    dup = np.maximum(du, 1.0)
    logdetDup, sgndetDup = float(np.sum(np.log(dup))), 1.0  # :247
    ws.M[...] = Lu.conj().T                                 # :250
    ws.M[...] = ws.M / dup[:, None]                         # :251
    dum = np.minimum(du, 1.0)
    G[...] = Ru @ Rvi                                       # :260
    G[...] = dum[:, None] * G                               # :261
    dvm = np.minimum(dv, 1.0)
    ws.Mpp[...] = ws.M @ Lv                                 # :269
    ws.Mpp[...] = ws.Mpp * dvm[None, :]                     # :270
    G[...] = G + ws.Mpp                                     # :273-274
    logdetMi, sgndetMi = inv_lu_(G, ws.lu_ws)               # :277-278
    ws.Mpp[...] = Rvi @ G                                   # :281
    G[...] = ws.Mpp @ ws.M                                  # :282
    sgndetG = sgndetRvi * np.conj(sgndetDvp) * sgndetMi * np.conj(sgndetDup) * np.conj(sgndetLu)
    logdetG = -logdetDvp + logdetMi - logdetDup             # :285-286
    return float(np.real(logdetG)), sgndetG


def inv_invUpV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_invUpV!(G, U, V, ws)`` -- exported_functions.jl:324-402.  G := (U^-1 + V)^-1.
    U and V may be the same object (test/runtests.jl:280): inputs are read-only."""
    Lu, du, Ru = U.L, U.d, U.R
    Lv, dv, Rv = V.L, V.d, V.R
    ws.Mp[...] = Ru                                         # :334
    _logdetRu, sgndetRu = det_lu_(ws.Mp, ws.lu_ws)          # :335
    dum = np.minimum(du, 1.0)                               # :338-340
    logdetDum, sgndetDum = float(np.sum(np.log(dum))), 1.0  # :343
    DumRu = ws.Mp
    DumRu[...] = Ru                                         # :346-347
    DumRu[...] = dum[:, None] * DumRu                       # :348
    Rvi = ws.Mpp
    Rvi[...] = Rv                                           # :351-352
    _l, sgndetRvi = inv_lu_(Rvi, ws.lu_ws)                  # :353
    dvp = np.maximum(dv, 1.0)
    logdetDvp, sgndetDvp = float(np.sum(np.log(dvp))), 1.0  # :361
    Rvi[...] = Rvi / dvp[None, :]                           # :365
    dvm = np.minimum(dv, 1.0)
    G[...] = DumRu @ Lv                                     # :373
    G[...] = G * dvm[None, :]                               # :374
    dup = np.maximum(du, 1.0)
    ws.M[...] = Lu.conj().T @ Rvi                           # :383
    ws.M[...] = ws.M / dup[:, None]                         # :384
    G[...] = G + ws.M                                       # :387
    logdetMi, sgndetMi = inv_lu_(G, ws.lu_ws)               # :390-391
    ws.M[...] = G @ DumRu                                   # :394
    G[...] = Rvi @ ws.M                                     # :395
    sgndetG = sgndetRvi * np.conj(sgndetDvp) * sgndetMi * sgndetDum * sgndetRu   # :398
    logdetG = -logdetDvp + logdetMi + logdetDum             # :399
    return float(np.real(logdetG)), sgndetG
