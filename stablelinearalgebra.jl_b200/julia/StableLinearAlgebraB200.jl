# StableLinearAlgebraB200.jl -- Julia host shim over libsla_b200.so (include/sla_b200.h).
#
# Same exported names as StableLinearAlgebra.jl v1.5.1 (src/StableLinearAlgebra.jl:19-32), but `LDR` and
# `LDRWorkspace` hold CuArrays and represent a BATCH of independent factorisations (the batched form of
# `ldrs!(Fs, A::AbstractArray{T,3}, ws)`, src/LDR.jl:249-257): every method body is ONE ccall.
#
# NOTE: Julia is not installed in the build image, so this file is shipped as source and has not been
# executed there; the identical C symbols are exercised from Python (tests/test_gpu_parity.py) and, with exactly the
# argument types of the ccall tuples below, from a plain-C client (tests/abi_c/abi_check.c).
module StableLinearAlgebraB200

using CUDA
using LinearAlgebra
import LinearAlgebra: adjoint!, lmul!, rmul!, mul!, ldiv!, rdiv!, logabsdet
import Base: eltype, size, copyto!

export LDR, LDRWorkspace, ldr, ldr!, ldrs, ldrs!, ldr_workspace
export inv_IpA!, inv_IpUV!, inv_UpV!, inv_invUpV!, inv_lu!, det_lu!, ldiv_lu!, greens_chain!, sweep!

const libsla = get(ENV, "LIBSLA_B200", "libsla_b200.so")
const SLA_F64, SLA_C64 = Cint(0), Cint(1)
const BlasT = Union{Float64,ComplexF64}
dtype(::Type{Float64}) = SLA_F64
dtype(::Type{ComplexF64}) = SLA_C64

mutable struct Handle
    ptr::Ptr{Cvoid}
end
function Handle()
    r = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall((:sla_create, libsla), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}), r, CUDA.stream().handle), "sla_create")
    h = Handle(r[])
    finalizer(x -> ccall((:sla_destroy, libsla), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
# CUDA.jl gives every task its own stream: one library handle per task (task_local_storage), re-bound to the task's
# CURRENT stream before every call, so that library kernels are ordered with the surrounding CuArray operations.
function handle()
    h = get!(() -> Handle(), task_local_storage(), :sla_b200_handle)::Handle
    chk(ccall((:sla_set_stream, libsla), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, CUDA.stream().handle), "sla_set_stream")
    return h.ptr
end

function chk(code::Cint, what)
    code == 0 && return nothing
    code < 0 && throw(ArgumentError("$what: invalid argument #$(-code)"))
    error("$what: CUDA error $(code - 1000)")
end

"Batched LDR factorisation (src/LDR.jl:23-33): L, R are n x n x batch, d is n x batch."
struct LDR{T<:BlasT,E<:AbstractFloat} <: Factorization{T}
    L::CuArray{T,3}
    d::CuArray{E,2}
    R::CuArray{T,3}
end
eltype(::LDR{T}) where {T} = T
size(F::LDR, dim...) = size(F.L, dim...)
nb(F::LDR) = (Cint(size(F.L, 1)), Cint(size(F.L, 3)))

"Caller-owned scratch blob (LDRWorkspace, src/LDR.jl:51-70), sized by sla_workspace_bytes."
struct LDRWorkspace{T<:BlasT,E<:AbstractFloat}
    blob::CuArray{UInt8,1}
    n::Int
    batch::Int
end
function ldr_workspace(A::CuArray{T,3}) where {T<:BlasT}
    n, batch = size(A, 1), size(A, 3)
    bytes = ccall((:sla_workspace_bytes, libsla), Csize_t, (Cint, Cint, Cint), dtype(T), n, batch)
    return LDRWorkspace{T,real(T)}(CUDA.zeros(UInt8, bytes), n, batch)
end
ldr_workspace(F::LDR) = ldr_workspace(F.L)

# lazy chklapackerror (src/lu.jl:75,91): info lives inside the workspace blob
function check_info(ws::LDRWorkspace{T}) where {T}
    p = ccall((:sla_workspace_info, libsla), CuPtr{Cint}, (CuPtr{Cvoid}, Cint, Cint, Cint), ws.blob, dtype(T), ws.n, ws.batch)
    info = Array(unsafe_wrap(CuArray, p, ws.batch))
    k = findfirst(!=(0), info)
    k === nothing || throw(LinearAlgebra.SingularException(Int(info[k])))
end

# ---- ldr / ldr! (src/LDR.jl:79-188) -------------------------------------------------------------
function ldr(A::CuArray{T,3}) where {T<:BlasT}
    F = LDR{T,real(T)}(similar(A), CuArray{real(T)}(undef, size(A, 1), size(A, 3)), similar(A))
    ldr!(F, I)
    return F
end
ldr(A::CuArray{T,3}, ws::LDRWorkspace{T}) where {T} = (F = ldr(A); ldr!(F, A, ws); F)
ldr(F::LDR, ignore...) = LDR(copy(F.L), copy(F.d), copy(F.R))
function ldr!(F::LDR{T}, ::UniformScaling, ignore...) where {T}
    n, b = nb(F)
    chk(ccall((:sla_set_identity, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}),
              handle(), dtype(T), n, b, F.L, F.d, F.R), "sla_set_identity")
end
function ldr!(Fo::LDR{T}, Fi::LDR{T}, ignore...) where {T}
    n, b = nb(Fo)
    chk(ccall((:sla_copy, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}), handle(), dtype(T), n, b, Fo.L, Fo.d, Fo.R, Fi.L, Fi.d, Fi.R), "sla_copy")
end
function ldr!(F::LDR{T}, ws::LDRWorkspace{T}) where {T}
    n, b = nb(F)
    chk(ccall((:sla_ldr, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}),
              handle(), dtype(T), n, b, F.L, F.d, F.R, ws.blob), "sla_ldr")
end
function ldr!(F::LDR{T}, A::CuArray{T}, ws::LDRWorkspace{T}) where {T}
    n, b = nb(F)
    strideA = ndims(A) == 2 ? 0 : n * n          # a single matrix is shared by the whole batch
    chk(ccall((:sla_ldr_from, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, Clonglong, CuPtr{Cvoid},
              CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, A, strideA, F.L, F.d, F.R, ws.blob), "sla_ldr_from")
end
ldrs(A::CuArray{T,3}, ws::LDRWorkspace{T}) where {T} = ldr(A, ws)       # a batch IS the vector of LDRs
ldrs!(Fs::LDR{T}, A::CuArray{T,3}, ws::LDRWorkspace{T}) where {T} = ldr!(Fs, A, ws)

# ---- copyto! / adjoint! (src/overloaded_functions.jl:34-86) -------------------------------------
copyto!(U::LDR{T}, V::CuArray{T}, ws::LDRWorkspace{T}) where {T} = ldr!(U, V, ws)
copyto!(U::LDR, I::UniformScaling, ignore...) = ldr!(U, I)
copyto!(U::LDR{T}, V::LDR{T}, ignore...) where {T} = ldr!(U, V)
for (jl, sym) in ((:copyto!, :sla_ldr_to_mat), (:adjoint!, :sla_adjoint))
    @eval function $jl(U::CuArray{T,3}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}
        n, b = nb(V)
        chk(ccall(($(QuoteNode(sym)), libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid},
                  CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U, V.L, V.d, V.R, ws.blob), $(string(sym)))
    end
end

# ---- lmul! / rmul! / mul! (src/overloaded_functions.jl:97-376) ----------------------------------
mstride(M::CuArray, n) = ndims(M) == 2 ? 0 : n * n
function lmul!(U::CuArray{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}       # :126-145
    n, b = nb(V)
    chk(ccall((:sla_lmul_mat_ldr, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, Clonglong, CuPtr{Cvoid},
              CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U, mstride(U, n), V.L, V.d, V.R, ws.blob), "sla_lmul_mat_ldr")
end
function rmul!(U::LDR{T}, V::CuArray{T}, ws::LDRWorkspace{T}) where {T}       # :234-253
    n, b = nb(U)
    chk(ccall((:sla_rmul_ldr_mat, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cvoid}, Clonglong, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V, mstride(V, n), ws.blob), "sla_rmul_ldr_mat")
end
for (jl, sym) in ((:lmul!, :sla_lmul_ldr_ldr), (:rmul!, :sla_rmul_ldr_ldr))  # :166-193, :274-301
    @eval function $jl(U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}
        U === V && throw(ArgumentError("U and V must be distinct"))
        n, b = nb(U)
        chk(ccall(($(QuoteNode(sym)), libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
                  CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V.L, V.d, V.R, ws.blob), $(string(sym)))
    end
end
function lmul!(U::LDR{T}, V::CuArray{T,3}, ws::LDRWorkspace{T}) where {T}     # :97-106
    n, b = nb(U)
    chk(ccall((:sla_lmul_ldr_mat, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V, ws.blob), "sla_lmul_ldr_mat")
end
function rmul!(U::CuArray{T,3}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}     # :205-214
    n, b = nb(V)
    chk(ccall((:sla_rmul_mat_ldr, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble},
              CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U, V.L, V.d, V.R, ws.blob), "sla_rmul_mat_ldr")
end
mul!(H::CuArray{T,3}, U::LDR{T}, V::CuArray{T,3}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); lmul!(U, H, ws))   # :314-318
mul!(H::CuArray{T,3}, U::CuArray{T,3}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, U); rmul!(H, V, ws))   # :326-330
mul!(H::LDR{T}, U::CuArray{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); lmul!(U, H, ws))           # :338-342
mul!(H::LDR{T}, U::LDR{T}, V::CuArray{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, U); rmul!(H, V, ws))           # :351-355
mul!(H::LDR{T}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); lmul!(U, H, ws))               # :364-368

# ---- ldiv! / rdiv! (src/overloaded_functions.jl:388-709) ---------------------------------------
function ldiv!(U::LDR{T}, V::CuArray{T,3}, ws::LDRWorkspace{T}) where {T}     # :388-400
    n, b = nb(U)
    chk(ccall((:sla_ldiv_ldr_mat, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V, ws.blob), "sla_ldiv_ldr_mat")
end
function rdiv!(U::CuArray{T,3}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}     # :549-565
    n, b = nb(V)
    chk(ccall((:sla_rdiv_mat_ldr, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble},
              CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U, V.L, V.d, V.R, ws.blob), "sla_rdiv_mat_ldr")
end
for (jl, sym) in ((:ldiv!, :sla_ldiv_ldr_ldr), (:rdiv!, :sla_rdiv_ldr_ldr))  # :437-464, :603-637
    @eval function $jl(U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}
        U === V && throw(ArgumentError("U and V must be distinct"))
        n, b = nb(U)
        chk(ccall(($(QuoteNode(sym)), libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
                  CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V.L, V.d, V.R, ws.blob), $(string(sym)))
    end
end
function ldiv!(U::CuArray{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}       # :502-521
    n, b = nb(V)
    chk(ccall((:sla_ldiv_mat_ldr, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, Clonglong, CuPtr{Cvoid},
              CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, U, mstride(U, n), V.L, V.d, V.R, ws.blob), "sla_ldiv_mat_ldr")
end
function rdiv!(U::LDR{T}, V::CuArray{T}, ws::LDRWorkspace{T}) where {T}       # :673-693
    n, b = nb(U)
    chk(ccall((:sla_rdiv_ldr_mat, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cvoid}, Clonglong, CuPtr{Cvoid}), handle(), dtype(T), n, b, U.L, U.d, U.R, V, mstride(V, n), ws.blob), "sla_rdiv_ldr_mat")
end
ldiv!(H::CuArray{T,3}, U::LDR{T}, V::CuArray{T,3}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); ldiv!(U, H, ws))   # :408-412
ldiv!(H::LDR{T}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); ldiv!(U, H, ws))               # :474-478
ldiv!(H::LDR{T}, U::CuArray{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, V); ldiv!(U, H, ws))           # :531-535
rdiv!(H::CuArray{T,3}, U::CuArray{T,3}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, U); rdiv!(H, V, ws))   # :573-577
rdiv!(H::LDR{T}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, U); rdiv!(H, V, ws))               # :645-649
rdiv!(H::LDR{T}, U::LDR{T}, V::CuArray{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(H, U); rdiv!(H, V, ws))           # :703-707
function ldiv_lu!(A::CuArray{T,3}, B::CuArray{T,3}, ws::LDRWorkspace{T}) where {T}                                     # src/lu.jl:167-176
    n, b = Cint(size(A, 1)), Cint(size(A, 3))
    ld, sg = scalars(T, b)
    chk(ccall((:sla_ldiv_lu, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble},
              CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, A, B, ld, sg, ws.blob), "sla_ldiv_lu")
    return ld, sg
end

# ---- determinants / inverses; each returns (logdet::CuVector{E}, sign::CuVector{T}) per matrix ----
scalars(::Type{T}, b) where {T} = (CuArray{real(T)}(undef, b), CuArray{T}(undef, b))
for (jl, sym) in ((:det_lu!, :sla_det_lu), (:inv_lu!, :sla_inv_lu))            # src/lu.jl:123-158
    @eval function $jl(A::CuArray{T,3}, ws::LDRWorkspace{T}) where {T}
        n, b = Cint(size(A, 1)), Cint(size(A, 3))
        ld, sg = scalars(T, b)
        chk(ccall(($(QuoteNode(sym)), libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
                  CuPtr{Cvoid}), handle(), dtype(T), n, b, A, ld, sg, ws.blob), $(string(sym)))
        return ld, sg
    end
end
function logabsdet(A::LDR{T}, ws::LDRWorkspace{T}) where {T}                   # overloaded_functions.jl:721-733
    n, b = nb(A)
    ld, sg = scalars(T, b)
    chk(ccall((:sla_logabsdet, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid},
              CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, A.L, A.d, A.R, ld, sg, ws.blob), "sla_logabsdet")
    return ld, sg
end
function inv_IpA!(G::CuArray{T,3}, A::LDR{T}, ws::LDRWorkspace{T}) where {T}   # exported_functions.jl:25-71
    n, b = nb(A)
    ld, sg = scalars(T, b)
    chk(ccall((:sla_inv_IpA, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble},
              CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}), handle(), dtype(T), n, b, G, A.L, A.d, A.R, ld, sg, ws.blob), "sla_inv_IpA")
    return ld, sg
end
for (jl, sym) in ((:inv_IpUV!, :sla_inv_IpUV), (:inv_UpV!, :sla_inv_UpV), (:inv_invUpV!, :sla_inv_invUpV))   # :97-402
    @eval function $jl(G::CuArray{T,3}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T}
        n, b = nb(U)
        ld, sg = scalars(T, b)
        chk(ccall(($(QuoteNode(sym)), libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble},
                  CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}),
                  handle(), dtype(T), n, b, G, U.L, U.d, U.R, V.L, V.d, V.R, ld, sg, ws.blob), $(string(sym)))
        return ld, sg
    end
end

# ---- source compatibility with the reference's single-matrix signatures -----------------------------------------
# `ldr(A::CuMatrix)` is a batch of one (L, R of size (n, n, 1)); the inverses / determinants called with a CuMatrix
# return the reference's `Tuple{E,T}` of host scalars (src/exported_functions.jl:25, src/lu.jl:123,149) instead of
# device vectors.
as3(A::CuArray{T,2}) where {T} = reshape(A, size(A, 1), size(A, 2), 1)
first_scalar(ld, sg) = (Array(ld)[1], Array(sg)[1])
ldr_workspace(A::CuArray{T,2}) where {T<:BlasT} = ldr_workspace(as3(A))
ldr(A::CuArray{T,2}) where {T<:BlasT} = ldr(as3(A))
ldr(A::CuArray{T,2}, ws::LDRWorkspace{T}) where {T<:BlasT} = ldr(as3(A), ws)
copyto!(U::CuArray{T,2}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (copyto!(as3(U), V, ws); U)
adjoint!(U::CuArray{T,2}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (adjoint!(as3(U), V, ws); U)
lmul!(U::LDR{T}, V::CuArray{T,2}, ws::LDRWorkspace{T}) where {T} = (lmul!(U, as3(V), ws); V)
rmul!(U::CuArray{T,2}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (rmul!(as3(U), V, ws); U)
ldiv!(U::LDR{T}, V::CuArray{T,2}, ws::LDRWorkspace{T}) where {T} = (ldiv!(U, as3(V), ws); V)
rdiv!(U::CuArray{T,2}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = (rdiv!(as3(U), V, ws); U)
det_lu!(A::CuArray{T,2}, ws::LDRWorkspace{T}) where {T} = first_scalar(det_lu!(as3(A), ws)...)
inv_lu!(A::CuArray{T,2}, ws::LDRWorkspace{T}) where {T} = first_scalar(inv_lu!(as3(A), ws)...)
ldiv_lu!(A::CuArray{T,2}, B::CuArray{T,2}, ws::LDRWorkspace{T}) where {T} = first_scalar(ldiv_lu!(as3(A), as3(B), ws)...)
inv_IpA!(G::CuArray{T,2}, A::LDR{T}, ws::LDRWorkspace{T}) where {T} = first_scalar(inv_IpA!(as3(G), A, ws)...)
inv_IpUV!(G::CuArray{T,2}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = first_scalar(inv_IpUV!(as3(G), U, V, ws)...)
inv_UpV!(G::CuArray{T,2}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = first_scalar(inv_UpV!(as3(G), U, V, ws)...)
inv_invUpV!(G::CuArray{T,2}, U::LDR{T}, V::LDR{T}, ws::LDRWorkspace{T}) where {T} = first_scalar(inv_invUpV!(as3(G), U, V, ws)...)

# ---- fused drivers: their workspace blobs are cached per (entry point, eltype, sizes) and task -------------------------
function driver_blob(key, bytes)
    cache = get!(() -> Dict{Any,CuArray{UInt8,1}}(), task_local_storage(), :sla_b200_blobs)::Dict{Any,CuArray{UInt8,1}}
    blob = get(cache, key, nothing)
    if blob === nothing || length(blob) < bytes
        blob = CUDA.zeros(UInt8, bytes)
        cache[key] = blob
    end
    return blob
end

"Fused driver: one stabilised G(0) evaluation per chain (sla_greens_chain); fields is Int8 (n, L, batch)."
function greens_chain!(G::CuArray{T,3}, expK::CuArray{T,2}, fields::CuArray{Int8,3}, lam::Real, n_stab::Integer;
                       inverse::Symbol = :IpA) where {T<:BlasT}
    n, L, b = size(fields)
    bytes = ccall((:sla_greens_chain_workspace_bytes, libsla), Csize_t, (Cint, Cint, Cint, Cint, Cint), dtype(T), n, b, L, n_stab)
    bytes == 0 && throw(ArgumentError("greens_chain!: unsupported sizes (n = $n, L = $L, n_stab = $n_stab)"))
    ws = driver_blob((:chain, T, n, b, L, n_stab), bytes)
    ld, sg = scalars(T, b)
    chk(ccall((:sla_greens_chain, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Int8}, Cdouble,
              Cint, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t),
              handle(), dtype(T), n, b, L, n_stab, expK, fields, Float64(lam), inverse === :IpA ? 0 : 1, G, ld, sg,
              CU_NULL, CU_NULL, CU_NULL, ws, bytes), "sla_greens_chain")
    return ld, sg
end

"""
Fused sweep stepper (sla_sweep): stack of left factorisations by `rmul!`, G(0) by `inv_IpA!`, then per slice
G <- B_l G B_l^-1 and every n_stab slices `lmul!` + `inv_IpUV!(G, F_right, F_left[g])`.  Returns
(logdet (batch, ng+1), sign (batch, ng+1), dG (batch, ng)); G receives G(beta), G0 (optional) G(0).
"""
function sweep!(G::CuArray{T,3}, expK::CuArray{T,2}, expKinv::CuArray{T,2}, fields::CuArray{Int8,3}, lam::Real,
                n_stab::Integer; G0::Union{Nothing,CuArray{T,3}} = nothing) where {T<:BlasT}
    n, L, b = size(fields)
    ng = L ÷ n_stab
    bytes = ccall((:sla_sweep_workspace_bytes, libsla), Csize_t, (Cint, Cint, Cint, Cint, Cint), dtype(T), n, b, L, n_stab)
    bytes == 0 && throw(ArgumentError("sweep!: unsupported sizes (n = $n, L = $L, n_stab = $n_stab)"))
    ws = driver_blob((:sweep, T, n, b, L, n_stab), bytes)
    ld = CuArray{Float64}(undef, b, ng + 1)
    sg = CuArray{T}(undef, b, ng + 1)
    dG = CuArray{Float64}(undef, b, ng)
    chk(ccall((:sla_sweep, libsla), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Int8},
              Cdouble, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cvoid}, Csize_t),
              handle(), dtype(T), n, b, L, n_stab, expK, expKinv, fields, Float64(lam), G0 === nothing ? CU_NULL : G0, G, ld, sg,
              dG, ws, bytes), "sla_sweep")
    return ld, sg, dG
end

end # module
