"""Host-side mirror of StableLinearAlgebra.jl's public API over the C ABI (include/sla_b200.h).

Same names and argument meaning as the reference (``!`` spelled ``_``): ``LDR``, ``ldr``, ``ldr_``,
``ldrs``, ``ldrs_``, ``ldr_workspace``, ``lmul_``, ``rmul_``, ``mul_``, ``copyto_``, ``adjoint_``,
``logabsdet``, ``det_lu_``, ``inv_lu_``, ``inv_IpA_``, ``inv_IpUV_``, ``inv_UpV_``, ``inv_invUpV_``
(src/StableLinearAlgebra.jl:19-32) -- but every object is a *batch* of independent matrices living on
the GPU (the batched form of ``ldrs!(Fs, A::AbstractArray{T,3}, ws)``, src/LDR.jl:249-257).

Storage convention: a batch of matrices is a torch CUDA tensor of shape ``(batch, n, n)`` whose memory
is the Julia ``CuArray{T,3}`` of size ``(n, n, batch)``, i.e. ``t[b, j, i] == A_b[i, j]`` (column
major).  ``to_device`` / ``to_host`` convert from/to ordinary ``(batch, n, n)`` numpy arrays.
PyTorch is used only for device memory and streams; all arithmetic is in libsla_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = [
    "LDR", "LDRWorkspace", "to_device", "to_host", "ldr", "ldr_", "ldrs", "ldrs_", "ldr_workspace",
    "copyto_", "adjoint_", "lmul_", "rmul_", "mul_", "logabsdet", "det_lu_", "inv_lu_",
    "ldiv_", "rdiv_", "ldiv_lu_",
    "inv_IpA_", "inv_IpUV_", "inv_UpV_", "inv_invUpV_", "gemm", "greens_chain", "check_info",
    "greens_chain_workspace_bytes", "greens_chain_info", "sweep", "sweep_info", "last_ldr_algo", "ldr_algo_launches", "walker_sums", "SlaError",
]


class SlaError(RuntimeError):
    pass


_CODES = {torch.float64: _lib.SLA_F64, torch.complex128: _lib.SLA_C64, torch.float32: _lib.SLA_F32,
          torch.complex64: _lib.SLA_C32}


def _code_of(dtype: torch.dtype) -> int:
    """Element-type code of the C ABI.  Float32 / ComplexF32 are storage types (promoted to FP64 inside the library;
    ``d`` and ``log|det|`` stay Float64, see include/sla_b200.h)."""
    try:
        return _CODES[dtype]
    except KeyError:
        raise TypeError(f"unsupported element type {dtype} (Float64 / ComplexF64 / Float32 / ComplexF32)") from None


def _dtype_code(t: torch.Tensor) -> int:
    return _code_of(t.dtype)


def _check(code: int, what: str):
    if code == 0:
        return
    if code == _lib.SLA_ENOTIMPL:
        raise SlaError(f"{what}: the requested kernel does not exist for this element type / size (SLA_ENOTIMPL)")
    if code == _lib.SLA_EDEVICE:
        raise SlaError(f"{what}: the handle belongs to another device than the current one (SLA_EDEVICE)")
    if code < 0:
        raise SlaError(f"{what}: invalid argument #{-code}")
    raise SlaError(f"{what}: CUDA error {code - 1000}")


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise SlaError("device tensor expected (the library has no CPU path)")
    if not t.is_contiguous():
        raise SlaError("contiguous tensor expected")
    return C.c_void_p(t.data_ptr())


class _Handle:
    """One library handle per (device, stream); created lazily."""
    _cache: dict = {}

    @classmethod
    def current(cls):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise SlaError("no CUDA device: libsla_b200 has no CPU fallback")
        dev = torch.cuda.current_device()
        stream = torch.cuda.current_stream().cuda_stream
        key = (dev, stream)
        h = cls._cache.get(key)
        if h is None:
            hp = C.c_void_p()
            _check(lib.sla_create(C.byref(hp), C.c_void_p(stream)), "sla_create")
            h = cls._cache[key] = hp
        return lib, h


def to_device(A, device="cuda") -> torch.Tensor:
    """numpy ``(n,n)`` or ``(batch,n,n)`` (ordinary row/col indexing) -> column-major device batch."""
    A = np.asarray(A)
    if A.ndim == 2:
        A = A[None]
    if A.dtype in (np.float32, np.complex64):
        dt = A.dtype                       # Float32 / ComplexF32 storage types are kept
    else:
        dt = np.complex128 if np.iscomplexobj(A) else np.float64
    At = np.ascontiguousarray(np.swapaxes(A.astype(dt, copy=False), 1, 2))
    return torch.from_numpy(At).to(device)


def to_host(t: torch.Tensor) -> np.ndarray:
    """column-major device batch -> numpy ``(batch,n,n)`` with ordinary indexing."""
    return np.swapaxes(t.detach().cpu().numpy(), 1, 2).copy()


class LDR:
    """Batched ``LDR{T,E}`` (src/LDR.jl:23-33): ``A_b = L_b * Diagonal(d_b) * R_b`` for every b."""

    def __init__(self, L: torch.Tensor, d: torch.Tensor, R: torch.Tensor):
        assert L.ndim == 3 and L.shape[1] == L.shape[2] and R.shape == L.shape and d.shape == L.shape[:2]
        assert d.dtype == torch.float64
        self.L, self.d, self.R = L, d, R

    @property
    def dtype(self):           # eltype, overloaded_functions.jl:10
        return self.L.dtype

    @property
    def batch(self):
        return self.L.shape[0]

    @property
    def n(self):
        return self.L.shape[1]

    @property
    def shape(self):           # size, overloaded_functions.jl:22
        return (self.n, self.n)


class LDRWorkspace:
    """Batched ``LDRWorkspace{T,E}`` (src/LDR.jl:51-70): one caller-owned scratch blob."""

    def __init__(self, n: int, batch: int, dtype: torch.dtype, device="cuda"):
        lib = _lib.load()
        self.n, self.batch, self.dtype = n, batch, dtype
        self.code = _code_of(dtype)
        self.nbytes = lib.sla_workspace_bytes(self.code, n, batch)
        if self.nbytes == 0:
            raise SlaError(f"unsupported size n={n} (1 <= n <= {lib.sla_max_n(self.code)} for this element type) or "
                           f"batch={batch}")
        self.blob = torch.empty(self.nbytes, dtype=torch.uint8, device=device)

    def info(self) -> torch.Tensor:
        """Per-matrix LAPACK-style info of the last LU-based call (0 = ok, k>0: U(k,k) == 0)."""
        lib = _lib.load()
        p = lib.sla_workspace_info(C.c_void_p(self.blob.data_ptr()), self.code, self.n, self.batch)
        off = p - self.blob.data_ptr()
        return self.blob[off:off + 4 * self.batch].view(torch.int32)


def _raise_info(info):
    bad = np.nonzero(info)[0]
    if bad.size:
        k = int(info[bad[0]])
        if k == _lib.SLA_INFO_RANGE:
            raise np.linalg.LinAlgError(f"ldr!: a column norm of batch entry {bad[0]} is outside the range in which "
                                        "squared norms are representable (about 1e-145 .. 1e150)")
        raise np.linalg.LinAlgError(f"singular U({k},{k}) in batch entry {bad[0]}")


def check_info(ws: LDRWorkspace):
    """Lazy ``chklapackerror`` (src/lu.jl:75,91): raise if any matrix hit an exactly-zero (or non-finite) pivot in ANY
    of the LU factorisations of the last call on this workspace, or if ``ldr!`` met a column norm out of range."""
    _raise_info(ws.info().cpu().numpy())


def greens_chain_info(ws: torch.Tensor, dtype: torch.dtype, n: int, batch: int, L: int, n_stab: int) -> torch.Tensor:
    """The per-chain info codes inside a ``greens_chain`` workspace blob (same meaning as ``LDRWorkspace.info``)."""
    lib = _lib.load()
    code = _code_of(dtype)
    p = lib.sla_greens_chain_workspace_info(C.c_void_p(ws.data_ptr()), code, n, batch, L, n_stab)
    off = p - ws.data_ptr()
    return ws[off:off + 4 * batch].view(torch.int32)


def last_ldr_algo() -> str:
    """Name of the LDR factorisation kernel the dispatcher used for the last ``ldr!`` on the current (device, stream)
    handle: one of ``_lib.LDR_ALGOS`` (diagnostics; the parity tests assert which kernel they exercised)."""
    lib, h = _Handle.current()
    return _lib.LDR_ALGOS[lib.sla_last_ldr_algo(h)]


def ldr_algo_launches() -> dict:
    lib, h = _Handle.current()
    return {name: int(lib.sla_ldr_algo_launches(h, code)) for code, name in _lib.LDR_ALGOS.items() if code}


def ldr_workspace(A) -> LDRWorkspace:
    """``ldr_workspace(A)`` / ``ldr_workspace(F)`` (src/LDR.jl:267-282)."""
    L = A.L if isinstance(A, LDR) else A
    return LDRWorkspace(L.shape[1], L.shape[0], L.dtype, L.device)


def _chk_ws(ws: LDRWorkspace, F):
    if not isinstance(ws, LDRWorkspace):
        raise SlaError("an LDRWorkspace is required (ldr_workspace(A))")
    L = F.L if isinstance(F, LDR) else F
    if ws.n != L.shape[1] or ws.batch != L.shape[0] or ws.dtype != L.dtype:
        raise SlaError("workspace does not match (n, batch, eltype)")
    if ws.blob.device != L.device:
        raise SlaError("workspace lives on another device")


def _chk_mat(M, F, what, shared_ok=True):
    """A matrix operand of an LDR call: same (n, n), eltype and device; batch equal to F's (or 1 = shared)."""
    L = F.L if isinstance(F, LDR) else F
    if M.ndim != 3 or M.shape[1:] != L.shape[1:]:
        raise SlaError(f"{what}: expected a column-major batch of {L.shape[1]}x{L.shape[2]} matrices, got {tuple(M.shape)}")
    if M.dtype != L.dtype:
        raise SlaError(f"{what}: element type {M.dtype} does not match {L.dtype}")
    if M.device != L.device:
        raise SlaError(f"{what}: operand lives on another device")
    if M.shape[0] != L.shape[0] and not (shared_ok and M.shape[0] == 1):
        raise SlaError(f"{what}: batch {M.shape[0]} does not match {L.shape[0]}" + (" (or 1 = shared)" if shared_ok else ""))


def _chk_ldr(U: LDR, V: LDR, what):
    if U.L.shape != V.L.shape or U.dtype != V.dtype or U.L.device != V.L.device:
        raise SlaError(f"{what}: the two factorisations differ in (n, batch), element type or device")


def _identity_(F: LDR):
    lib, h = _Handle.current()
    _check(lib.sla_set_identity(h, _dtype_code(F.L), F.n, F.batch, _ptr(F.L), _ptr(F.d), _ptr(F.R)),
           "sla_set_identity")


def ldr(A, ws: LDRWorkspace | None = None) -> LDR:
    """``ldr(A)`` identity-initialised (LDR.jl:79-90); ``ldr(A, ws)`` factorise (97-102);
    ``ldr(F)`` copy (108-114).  ``A``: column-major device batch (see module docstring)."""
    if isinstance(A, LDR):
        return LDR(A.L.clone(), A.d.clone(), A.R.clone())
    F = LDR(torch.empty_like(A), torch.empty(A.shape[:2], dtype=torch.float64, device=A.device), torch.empty_like(A))
    if ws is None:
        _identity_(F)
    else:
        ldr_(F, A, ws)
    return F


def ldr_(F: LDR, A=None, ws: LDRWorkspace | None = None):
    """``ldr!``: (F, A, ws) factorise A (LDR.jl:122-126); (F, "I") identity (133-140); (Fout, Fin)
    copy (147-153); (F, ws=ws) factorise F.L in place (160-188)."""
    if isinstance(A, LDRWorkspace) and ws is None:
        A, ws = None, A
    lib, h = _Handle.current()
    code = _dtype_code(F.L)
    if isinstance(A, str) and A == "I":
        return _identity_(F)
    if isinstance(A, LDR):
        _check(lib.sla_copy(h, code, F.n, F.batch, _ptr(F.L), _ptr(F.d), _ptr(F.R), _ptr(A.L), _ptr(A.d), _ptr(A.R)),
               "sla_copy")
        return None
    _chk_ws(ws, F)
    if A is not None:
        _chk_mat(A, F, "ldr!(F, A, ws)")
    if A is None:
        _check(lib.sla_ldr(h, code, F.n, F.batch, _ptr(F.L), _ptr(F.d), _ptr(F.R), _ptr(ws.blob)), "sla_ldr")
    else:
        stride = 0 if A.shape[0] == 1 and F.batch > 1 else F.n * F.n
        _check(lib.sla_ldr_from(h, code, F.n, F.batch, _ptr(A), stride, _ptr(F.L), _ptr(F.d), _ptr(F.R),
                                _ptr(ws.blob)), "sla_ldr_from")
    return None


def ldrs(A, N=None, ws: LDRWorkspace | None = None) -> LDR:
    """``ldrs`` (LDR.jl:197-241).  A batch *is* the vector of factorisations here:
    ``ldrs(A1, N)`` -> N identities; ``ldrs(A1, N, ws)`` -> N copies of ldr(A1); ``ldrs(A3, ws=ws)`` ->
    one factorisation per matrix of the batch."""
    if isinstance(N, LDRWorkspace):
        N, ws = None, N
    if N is None:
        return ldr(A, ws)
    A1 = A[:1]
    big = A1.expand(N, -1, -1).contiguous()
    if ws is None:
        return ldr(big)
    return ldr(big, ws)


def ldrs_(Fs: LDR, A3, ws: LDRWorkspace):
    """``ldrs!(Fs, A3, ws)`` (LDR.jl:249-257)."""
    return ldr_(Fs, A3, ws)


def copyto_(U, V, ws: LDRWorkspace | None = None):
    """``copyto!``: (Matrix, LDR, ws) U = L D R (overloaded_functions.jl:34-45); (LDR, Matrix, ws)
    factorise (55); (LDR, "I") (56); (LDR, LDR) (63)."""
    if isinstance(U, LDR):
        return ldr_(U, V, ws)
    lib, h = _Handle.current()
    _chk_ws(ws, V)
    _chk_mat(U, V, "copyto!(U, V::LDR, ws)", shared_ok=False)
    _check(lib.sla_ldr_to_mat(h, _dtype_code(U), V.n, V.batch, _ptr(U), _ptr(V.L), _ptr(V.d), _ptr(V.R),
                              _ptr(ws.blob)), "sla_ldr_to_mat")


def adjoint_(At, A: LDR, ws: LDRWorkspace):
    """``adjoint!(A', A::LDR, ws)`` (overloaded_functions.jl:75-86)."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    _chk_mat(At, A, "adjoint!", shared_ok=False)
    _check(lib.sla_adjoint(h, _dtype_code(At), A.n, A.batch, _ptr(At), _ptr(A.L), _ptr(A.d), _ptr(A.R),
                           _ptr(ws.blob)), "sla_adjoint")


def _mat_stride(M, F):
    return 0 if (M.shape[0] == 1 and F.batch > 1) else F.n * F.n


def lmul_(U, V, ws: LDRWorkspace):
    """``lmul!``: (LDR, Matrix) V := L D R V (overloaded_functions.jl:97-106); (Matrix, LDR) stabilised
    V := U V (126-145); (LDR, LDR) stabilised (166-193).  A one-matrix batch U is shared by all of V."""
    lib, h = _Handle.current()
    if isinstance(U, LDR) and not isinstance(V, LDR):
        _chk_ws(ws, U)
        _chk_mat(V, U, "lmul!(U::LDR, V)", shared_ok=False)
        return _check(lib.sla_lmul_ldr_mat(h, _dtype_code(V), U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V),
                                           _ptr(ws.blob)), "sla_lmul_ldr_mat")
    _chk_ws(ws, V)
    code = _dtype_code(V.L)
    if not isinstance(U, LDR):
        _chk_mat(U, V, "lmul!(U, V::LDR)")
        return _check(lib.sla_lmul_mat_ldr(h, code, V.n, V.batch, _ptr(U), _mat_stride(U, V), _ptr(V.L), _ptr(V.d),
                                           _ptr(V.R), _ptr(ws.blob)), "sla_lmul_mat_ldr")
    _chk_ldr(U, V, "lmul!(U::LDR, V::LDR)")
    return _check(lib.sla_lmul_ldr_ldr(h, code, V.n, V.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V.L), _ptr(V.d),
                                       _ptr(V.R), _ptr(ws.blob)), "sla_lmul_ldr_ldr")


def rmul_(U, V, ws: LDRWorkspace):
    """``rmul!``: (Matrix, LDR) U := U L D R (overloaded_functions.jl:205-214); (LDR, Matrix) stabilised
    U := U V (234-253); (LDR, LDR) stabilised (274-301)."""
    lib, h = _Handle.current()
    if not isinstance(U, LDR):
        _chk_ws(ws, V)
        _chk_mat(U, V, "rmul!(U, V::LDR)", shared_ok=False)
        return _check(lib.sla_rmul_mat_ldr(h, _dtype_code(U), V.n, V.batch, _ptr(U), _ptr(V.L), _ptr(V.d), _ptr(V.R),
                                           _ptr(ws.blob)), "sla_rmul_mat_ldr")
    _chk_ws(ws, U)
    code = _dtype_code(U.L)
    if not isinstance(V, LDR):
        _chk_mat(V, U, "rmul!(U::LDR, V)")
        return _check(lib.sla_rmul_ldr_mat(h, code, U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V),
                                           _mat_stride(V, U), _ptr(ws.blob)), "sla_rmul_ldr_mat")
    _chk_ldr(U, V, "rmul!(U::LDR, V::LDR)")
    return _check(lib.sla_rmul_ldr_ldr(h, code, U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V.L), _ptr(V.d),
                                       _ptr(V.R), _ptr(ws.blob)), "sla_rmul_ldr_ldr")


def mul_(H, U, V, ws: LDRWorkspace):
    """``mul!`` x5 (overloaded_functions.jl:314-376) = ``copyto!`` + ``lmul!``/``rmul!``."""
    if not isinstance(H, LDR):
        if isinstance(U, LDR):
            H.copy_(V)
            return lmul_(U, H, ws)
        H.copy_(U)
        return rmul_(H, V, ws)
    if isinstance(U, LDR) and not isinstance(V, LDR):
        ldr_(H, U)
        return rmul_(H, V, ws)
    ldr_(H, V)
    return lmul_(U, H, ws)


def ldiv_(*args):
    """``ldiv!`` x6 (overloaded_functions.jl:388-538): (U::LDR, V::Matrix) V := U^-1 V (388-400);
    (U::LDR, V::LDR) stabilised (437-464); (U::Matrix, V::LDR) stabilised (502-521); the four-argument
    forms ``ldiv_(H, U, V, ws)`` copy V into H first (408, 474, 531)."""
    if len(args) == 4:
        H, U, V, ws = args
        if isinstance(H, LDR):
            ldr_(H, V)
        else:
            H.copy_(V)
        return ldiv_(U, H, ws)
    U, V, ws = args
    lib, h = _Handle.current()
    if isinstance(U, LDR) and not isinstance(V, LDR):
        _chk_ws(ws, U)
        _chk_mat(V, U, "ldiv!(U::LDR, V)", shared_ok=False)
        return _check(lib.sla_ldiv_ldr_mat(h, _dtype_code(V), U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V),
                                           _ptr(ws.blob)), "sla_ldiv_ldr_mat")
    _chk_ws(ws, V)
    code = _dtype_code(V.L)
    if isinstance(U, LDR):
        _chk_ldr(U, V, "ldiv!(U::LDR, V::LDR)")
        return _check(lib.sla_ldiv_ldr_ldr(h, code, V.n, V.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V.L), _ptr(V.d),
                                           _ptr(V.R), _ptr(ws.blob)), "sla_ldiv_ldr_ldr")
    _chk_mat(U, V, "ldiv!(U, V::LDR)")
    return _check(lib.sla_ldiv_mat_ldr(h, code, V.n, V.batch, _ptr(U), _mat_stride(U, V), _ptr(V.L), _ptr(V.d), _ptr(V.R),
                                       _ptr(ws.blob)), "sla_ldiv_mat_ldr")


def rdiv_(*args):
    """``rdiv!`` x6 (overloaded_functions.jl:549-709): (U::Matrix, V::LDR) U := U V^-1 (549-565);
    (U::LDR, V::LDR) stabilised (603-637); (U::LDR, V::Matrix) stabilised (673-693); the four-argument
    forms ``rdiv_(H, U, V, ws)`` copy U into H first (573, 645, 703)."""
    if len(args) == 4:
        H, U, V, ws = args
        if isinstance(H, LDR):
            ldr_(H, U)
        else:
            H.copy_(U)
        return rdiv_(H, V, ws)
    U, V, ws = args
    lib, h = _Handle.current()
    if not isinstance(U, LDR):
        _chk_ws(ws, V)
        _chk_mat(U, V, "rdiv!(U, V::LDR)", shared_ok=False)
        return _check(lib.sla_rdiv_mat_ldr(h, _dtype_code(U), V.n, V.batch, _ptr(U), _ptr(V.L), _ptr(V.d), _ptr(V.R),
                                           _ptr(ws.blob)), "sla_rdiv_mat_ldr")
    _chk_ws(ws, U)
    code = _dtype_code(U.L)
    if isinstance(V, LDR):
        _chk_ldr(U, V, "rdiv!(U::LDR, V::LDR)")
        return _check(lib.sla_rdiv_ldr_ldr(h, code, U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V.L), _ptr(V.d),
                                           _ptr(V.R), _ptr(ws.blob)), "sla_rdiv_ldr_ldr")
    _chk_mat(V, U, "rdiv!(U::LDR, V)")
    return _check(lib.sla_rdiv_ldr_mat(h, code, U.n, U.batch, _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V), _mat_stride(V, U),
                                       _ptr(ws.blob)), "sla_rdiv_ldr_mat")


def _scalars(F_or_A, batch, device):
    dt = F_or_A.dtype
    return (torch.empty(batch, dtype=torch.float64, device=device), torch.empty(batch, dtype=dt, device=device))


def det_lu_(A, ws: LDRWorkspace):
    """``det_lu!(A, ws)`` (src/lu.jl:123-140): A := LU; returns (log|det A|, sign det A) per matrix."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    logdet, sgn = _scalars(A, A.shape[0], A.device)
    _check(lib.sla_det_lu(h, _dtype_code(A), A.shape[1], A.shape[0], _ptr(A), _ptr(logdet), _ptr(sgn),
                          _ptr(ws.blob)), "sla_det_lu")
    return logdet, sgn


def inv_lu_(A, ws: LDRWorkspace):
    """``inv_lu!(A, ws)`` (src/lu.jl:149-158): A := A^-1; returns (log|det A^-1|, sign det A^-1)."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    logdet, sgn = _scalars(A, A.shape[0], A.device)
    _check(lib.sla_inv_lu(h, _dtype_code(A), A.shape[1], A.shape[0], _ptr(A), _ptr(logdet), _ptr(sgn),
                          _ptr(ws.blob)), "sla_inv_lu")
    return logdet, sgn


def ldiv_lu_(A, B, ws: LDRWorkspace):
    """``ldiv_lu!(A, B, ws)`` (src/lu.jl:167-176): B := A^-1 B, A := LU; returns (log|det A^-1|, sign det A^-1)."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    _chk_mat(B, A, "ldiv_lu!(A, B, ws)", shared_ok=False)
    logdet, sgn = _scalars(A, A.shape[0], A.device)
    _check(lib.sla_ldiv_lu(h, _dtype_code(A), A.shape[1], A.shape[0], _ptr(A), _ptr(B), _ptr(logdet), _ptr(sgn),
                           _ptr(ws.blob)), "sla_ldiv_lu")
    return logdet, sgn


def logabsdet(A: LDR, ws: LDRWorkspace):
    """``logabsdet(A::LDR, ws)`` (overloaded_functions.jl:721-733)."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    logdet, sgn = _scalars(A, A.batch, A.L.device)
    _check(lib.sla_logabsdet(h, _dtype_code(A.L), A.n, A.batch, _ptr(A.L), _ptr(A.d), _ptr(A.R), _ptr(logdet),
                             _ptr(sgn), _ptr(ws.blob)), "sla_logabsdet")
    return logdet, sgn


def inv_IpA_(G, A: LDR, ws: LDRWorkspace):
    """``inv_IpA!(G, A, ws)`` (exported_functions.jl:25-71): G := (I + A)^-1 -> (log|det G|, sign)."""
    lib, h = _Handle.current()
    _chk_ws(ws, A)
    _chk_mat(G, A, "inv_IpA!(G, A, ws)", shared_ok=False)
    logdet, sgn = _scalars(A, A.batch, A.L.device)
    _check(lib.sla_inv_IpA(h, _dtype_code(G), A.n, A.batch, _ptr(G), _ptr(A.L), _ptr(A.d), _ptr(A.R), _ptr(logdet),
                           _ptr(sgn), _ptr(ws.blob)), "sla_inv_IpA")
    return logdet, sgn


def _uv(name, G, U: LDR, V: LDR, ws: LDRWorkspace):
    lib, h = _Handle.current()
    _chk_ws(ws, U)
    _chk_ws(ws, V)
    _chk_ldr(U, V, name)
    _chk_mat(G, U, name, shared_ok=False)
    logdet, sgn = _scalars(U, U.batch, U.L.device)
    fn = getattr(lib, name)
    _check(fn(h, _dtype_code(G), U.n, U.batch, _ptr(G), _ptr(U.L), _ptr(U.d), _ptr(U.R), _ptr(V.L), _ptr(V.d),
              _ptr(V.R), _ptr(logdet), _ptr(sgn), _ptr(ws.blob)), name)
    return logdet, sgn


def inv_IpUV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_IpUV!(G, U, V, ws)`` (exported_functions.jl:97-177): G := (I + U V)^-1."""
    return _uv("sla_inv_IpUV", G, U, V, ws)


def inv_UpV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_UpV!(G, U, V, ws)`` (exported_functions.jl:211-289): G := (U + V)^-1."""
    return _uv("sla_inv_UpV", G, U, V, ws)


def inv_invUpV_(G, U: LDR, V: LDR, ws: LDRWorkspace):
    """``inv_invUpV!(G, U, V, ws)`` (exported_functions.jl:324-402): G := (U^-1 + V)^-1."""
    return _uv("sla_inv_invUpV", G, U, V, ws)


def gemm(C_, A, B, opa="N", opb="N", accumulate=False):
    """Plain batched ``mul!(C, A, B)`` on column-major device batches (A or B may be a 1-batch, shared)."""
    lib, h = _Handle.current()
    batch, n = C_.shape[0], C_.shape[1]
    sA = 0 if (A.shape[0] == 1 and batch > 1) else n * n
    sB = 0 if (B.shape[0] == 1 and batch > 1) else n * n
    ops = {"N": _lib.SLA_OP_N, "C": _lib.SLA_OP_C}
    _check(lib.sla_gemm(h, _dtype_code(C_), ops[opa], ops[opb], n, n, n, _ptr(A), n, sA, _ptr(B), n, sB, _ptr(C_), n,
                        n * n, batch, int(bool(accumulate))), "sla_gemm")


def walker_sums(logdet: torch.Tensor, sign: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``[sum logdet, sum Re sign, sum Im sign, batch]`` on the device (one kernel, fixed summation order): the local
    part of the cross-walker all-reduce of SURVEY.md section 8(e)."""
    lib, h = _Handle.current()
    if out is None:
        out = torch.empty(4, dtype=torch.float64, device=logdet.device)
    _check(lib.sla_walker_sums(h, _dtype_code(sign), logdet.numel(), _ptr(logdet), _ptr(sign), _ptr(out)),
           "sla_walker_sums")
    return out


def greens_chain_workspace_bytes(dtype: torch.dtype, n: int, batch: int, L: int, n_stab: int) -> int:
    lib = _lib.load()
    code = _code_of(dtype)
    return lib.sla_greens_chain_workspace_bytes(code, n, batch, L, n_stab)


def greens_chain(expK, fields, lam: float, n_stab: int, inverse="IpA", G=None, logdet=None, sign=None, out_F=None,
                 ws=None):
    """One stabilised G(tau=0) evaluation per chain (fused driver, ``sla_greens_chain``).

    expK: column-major device matrix (1, n, n) or (n, n); fields: int8 device tensor (batch, L, n).
    Returns (G, logdet, sign[, ws]) -- all device tensors; G is a column-major batch."""
    lib, h = _Handle.current()
    if expK.ndim == 2:
        expK = expK[None]
    batch, L, n = fields.shape
    assert fields.dtype == torch.int8 and expK.shape[1] == n
    dt = expK.dtype
    code = _dtype_code(expK)
    dev = expK.device
    if G is None:
        G = torch.empty((batch, n, n), dtype=dt, device=dev)
    if logdet is None:
        logdet = torch.empty(batch, dtype=torch.float64, device=dev)
    if sign is None:
        sign = torch.empty(batch, dtype=dt, device=dev)
    need = lib.sla_greens_chain_workspace_bytes(code, n, batch, L, n_stab)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
    inv = {"IpA": _lib.SLA_INV_IPA, "IpUV": _lib.SLA_INV_IPUV}[inverse]
    Lo = do = Ro = None
    if out_F is not None:
        Lo, do, Ro = out_F.L, out_F.d, out_F.R
    _check(lib.sla_greens_chain(h, code, n, batch, L, n_stab, _ptr(expK), _ptr(fields), float(lam), inv, _ptr(G),
                                _ptr(logdet), _ptr(sign), _ptr(Lo), _ptr(do), _ptr(Ro), _ptr(ws), ws.numel()),
           "sla_greens_chain")
    return G, logdet, sign, ws


def sweep(expK, expKinv, fields, lam: float, n_stab: int, G=None, G0=None, logdet=None, sign=None, dG=None, ws=None):
    """One stabilised upward sweep per chain (fused stepper ``sla_sweep``): the stack of left
    factorisations by ``rmul!``, G(0) by ``inv_IpA!``, then per slice the wrap G <- B_l G B_l^-1 and every n_stab
    slices ``lmul!`` + ``inv_IpUV!(G, F_right, F_left[g])`` (exported_functions.jl:97-177).

    expK / expKinv: column-major device matrices (n, n) or (1, n, n); fields: int8 (batch, L, n).
    Returns dict(G, G0, logdet[ng+1, batch], sign[ng+1, batch], dG[ng, batch], ws)."""
    lib, h = _Handle.current()
    if expK.ndim == 2:
        expK = expK[None]
    if expKinv.ndim == 2:
        expKinv = expKinv[None]
    batch, L, n = fields.shape
    if fields.dtype != torch.int8 or expK.shape[1:] != (n, n) or expKinv.shape != expK.shape or expKinv.dtype != expK.dtype:
        raise ValueError("sweep: expK / expKinv must be (n, n) of one element type and fields int8 (batch, L, n)")
    if n_stab <= 0 or L % n_stab:
        raise ValueError("sweep: L must be a multiple of n_stab")
    ng = L // n_stab
    dt, dev, code = expK.dtype, expK.device, _dtype_code(expK)
    G = torch.empty((batch, n, n), dtype=dt, device=dev) if G is None else G
    G0 = torch.empty((batch, n, n), dtype=dt, device=dev) if G0 is None else G0
    logdet = torch.empty((ng + 1, batch), dtype=torch.float64, device=dev) if logdet is None else logdet
    sign = torch.empty((ng + 1, batch), dtype=dt, device=dev) if sign is None else sign
    dG = torch.empty((ng, batch), dtype=torch.float64, device=dev) if dG is None else dG
    need = lib.sla_sweep_workspace_bytes(code, n, batch, L, n_stab)
    if need == 0:
        raise SlaError(f"sweep: unsupported size n={n}")
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
    _check(lib.sla_sweep(h, code, n, batch, L, n_stab, _ptr(expK), _ptr(expKinv), _ptr(fields), float(lam), _ptr(G0),
                         _ptr(G), _ptr(logdet), _ptr(sign), _ptr(dG), _ptr(ws), ws.numel()), "sla_sweep")
    return {"G": G, "G0": G0, "logdet": logdet, "sign": sign, "dG": dG, "ws": ws}


def sweep_info(ws: torch.Tensor, dtype: torch.dtype, n: int, batch: int, L: int, n_stab: int) -> torch.Tensor:
    """The per-chain info codes inside a ``sweep`` workspace blob (same meaning as ``LDRWorkspace.info``)."""
    lib = _lib.load()
    code = _code_of(dtype)
    p = lib.sla_sweep_workspace_info(C.c_void_p(ws.data_ptr()), code, n, batch, L, n_stab)
    off = p - ws.data_ptr()
    return ws[off:off + 4 * batch].view(torch.int32)
