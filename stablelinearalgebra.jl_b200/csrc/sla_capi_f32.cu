// Float32 / ComplexF32 element types (the reference generates its LAPACK wrappers for all four
// element types, src/qr.jl:29-32, src/lu.jl:26-29).  Here SLA_F32 / SLA_C32 are STORAGE types: matrices and signs are
// float / (float, float) in the caller's buffers, every entry point promotes its operands to FP64 inside the workspace
// blob, runs the FP64 kernels and demotes the results -- on B200 the FP64 tensor pipe is the fast path of this library,
// and a chain's D spans far more than the 1e+-38 a Float32 can hold, which is also why d and log|det| stay Float64
// for these types (LDR{Float32,Float64} in the reference's own type parameters, src/LDR.jl:23).
// Results are therefore at least as accurate as the reference's single-precision LAPACK path.
#include "../../include/sla_b200.h"
#include "sla_capi_f32.h"
#include "sla_kernels.h"

using namespace sla;

namespace {

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// bump allocator over the promote region of a workspace blob + the list of results to demote at the end
struct Prom {
    cudaStream_t s;
    int hi;              // SLA_F64 / SLA_C64
    int reals;           // reals per element: 1 or 2
    char* base;
    size_t off = 0, cap;
    cudaError_t err = cudaSuccess;
    struct Out { float* lo; const double* hi; long long count; } outs[8];
    int nouts = 0;

    Prom(cudaStream_t s_, int dtype, void* region, size_t cap_) : s(s_), hi(dtype - 2), reals(dtype == SLA_C32 ? 2 : 1), base((char*)region), cap(cap_) {}
    void* take(long long elems) {
        const size_t bytes = al256((size_t)elems * reals * sizeof(double));
        if (off + bytes > cap) { err = cudaErrorMemoryAllocation; return nullptr; }
        void* p = base + off;
        off += bytes;
        return p;
    }
    void* in(const void* lo, long long elems) {
        if (!lo || !base) return nullptr;
        void* p = take(elems);
        if (p && err == cudaSuccess) err = convert_f32_f64((double*)p, (const float*)lo, elems * reals, s);
        return p;
    }
    void* out(void* lo, long long elems) {
        if (!lo || !base) return nullptr;
        void* p = take(elems);
        if (p && nouts < 8) outs[nouts++] = {(float*)lo, (const double*)p, elems * reals};
        return p;
    }
    void* io(void* lo, long long elems) {
        void* p = out(lo, elems);
        if (p && err == cudaSuccess) err = convert_f32_f64((double*)p, (const float*)lo, elems * reals, s);
        return p;
    }
    int done(int rc) {
        if (err != cudaSuccess) return 1000 + (int)err;
        if (rc != 0) return rc;
        for (int i = 0; i < nouts; ++i) {
            cudaError_t e = convert_f64_f32(outs[i].lo, outs[i].hi, outs[i].count, s);
            if (e != cudaSuccess) return 1000 + (int)e;
        }
        return 0;
    }
};

constexpr int kPromMats = 8;   // most matrices any op-level entry point touches (inv_IpUV!: G, UL, UR, VL, VR)

size_t prom_bytes(int hi, int n, int batch) {
    const size_t el = (hi == SLA_C64 ? 2 : 1) * sizeof(double);
    return kPromMats * al256((size_t)n * n * batch * el) + 2 * al256((size_t)batch * el);
}

}  // namespace

namespace sla {

size_t f32_workspace_bytes(int dtype, int n, int batch) {
    const size_t w = sla_workspace_bytes(dtype - 2, n, batch);
    return w ? al256(w) + prom_bytes(dtype - 2, n, batch) : 0;
}
static size_t driver_prom_bytes(int hi, int n, int batch, int nsign, int nshared, int nbatched) {
    const size_t el = (hi == SLA_C64 ? 2 : 1) * sizeof(double);
    return nshared * al256((size_t)n * n * el) + nbatched * al256((size_t)n * n * batch * el) + al256((size_t)nsign * batch * el);
}
size_t f32_chain_workspace_bytes(int dtype, int n, int batch, int L, int n_stab) {
    const size_t w = sla_greens_chain_workspace_bytes(dtype - 2, n, batch, L, n_stab);
    return w ? al256(w) + driver_prom_bytes(dtype - 2, n, batch, 1, 1, 3) : 0;
}
size_t f32_sweep_workspace_bytes(int dtype, int n, int batch, int L, int n_stab) {
    const size_t w = sla_sweep_workspace_bytes(dtype - 2, n, batch, L, n_stab);
    return w ? al256(w) + driver_prom_bytes(dtype - 2, n, batch, L / n_stab + 1, 2, 2) : 0;
}

}  // namespace sla

// every wrapper: promote, call the FP64 entry point on the same handle/stream, demote
#define PROM_BEGIN(ws_expr)                                                                        \
    if (!h) return -1;                                                                             \
    if (n <= 0) return -3;                                                                         \
    if (batch < 0) return -4;                                                                      \
    if (batch == 0) return 0;                                                                      \
    if (!ws) return -99;                                                                           \
    const long long nn = (long long)n * n, NB = nn * batch;                                        \
    (void)NB;                                                                                      \
    const int hi = dtype - 2;                                                                      \
    const size_t wsz = al256(sla_workspace_bytes(hi, n, batch));                                   \
    if (wsz == 0) return SLA_ENOTIMPL;                                                             \
    Prom P(sla_handle_stream(h), dtype, (char*)ws + wsz, prom_bytes(hi, n, batch))

extern "C" {

int f32_ldr(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, void* ws) {
    PROM_BEGIN();
    return P.done(sla_ldr(h, hi, n, batch, P.io(L, NB), d, P.out(R, NB), ws));
}
int f32_ldr_from(sla_handle_t h, int dtype, int n, int batch, const void* A, long long strideA, void* L, double* d, void* R, void* ws) {
    PROM_BEGIN();
    return P.done(sla_ldr_from(h, hi, n, batch, P.in(A, strideA ? NB : nn), strideA, P.out(L, NB), d, P.out(R, NB), ws));
}
int f32_ldr_to_mat(sla_handle_t h, int dtype, int n, int batch, void* A, const void* L, const double* d, const void* R, void* ws, int adjoint) {
    PROM_BEGIN();
    void *A2 = P.out(A, NB), *L2 = P.in(L, NB), *R2 = P.in(R, NB);
    return P.done(adjoint ? sla_adjoint(h, hi, n, batch, A2, L2, d, R2, ws) : sla_ldr_to_mat(h, hi, n, batch, A2, L2, d, R2, ws));
}
// which: 0 lmul_mat_ldr, 1 ldiv_mat_ldr  (matrix U, then the LDR that is updated)
int f32_mat_ldr(sla_handle_t h, int dtype, int n, int batch, const void* U, long long strideU, void* L, double* d, void* R, void* ws, int which) {
    PROM_BEGIN();
    void *U2 = P.in(U, strideU ? NB : nn), *L2 = P.io(L, NB), *R2 = P.io(R, NB);
    return P.done(which == 0 ? sla_lmul_mat_ldr(h, hi, n, batch, U2, strideU, L2, d, R2, ws)
                             : sla_ldiv_mat_ldr(h, hi, n, batch, U2, strideU, L2, d, R2, ws));
}
// which: 0 rmul_ldr_mat, 1 rdiv_ldr_mat  (the LDR that is updated, then matrix V)
int f32_ldr_mat(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R, const void* V, long long strideV, void* ws, int which) {
    PROM_BEGIN();
    void *L2 = P.io(L, NB), *R2 = P.io(R, NB), *V2 = P.in(V, strideV ? NB : nn);
    return P.done(which == 0 ? sla_rmul_ldr_mat(h, hi, n, batch, L2, d, R2, V2, strideV, ws)
                             : sla_rdiv_ldr_mat(h, hi, n, batch, L2, d, R2, V2, strideV, ws));
}
// which: 0 lmul_ldr_ldr, 1 ldiv_ldr_ldr (U read-only, V updated); 2 rmul_ldr_ldr, 3 rdiv_ldr_ldr (U updated, V read-only)
int f32_ldr_ldr(sla_handle_t h, int dtype, int n, int batch, void* UL, double* Ud, void* UR, void* VL, double* Vd, void* VR, void* ws, int which) {
    PROM_BEGIN();
    const bool upd_v = which < 2;
    void *UL2 = upd_v ? P.in(UL, NB) : P.io(UL, NB), *UR2 = upd_v ? P.in(UR, NB) : P.io(UR, NB);
    void *VL2 = upd_v ? P.io(VL, NB) : P.in(VL, NB), *VR2 = upd_v ? P.io(VR, NB) : P.in(VR, NB);
    int rc;
    switch (which) {
        case 0: rc = sla_lmul_ldr_ldr(h, hi, n, batch, UL2, Ud, UR2, VL2, Vd, VR2, ws); break;
        case 1: rc = sla_ldiv_ldr_ldr(h, hi, n, batch, UL2, Ud, UR2, VL2, Vd, VR2, ws); break;
        case 2: rc = sla_rmul_ldr_ldr(h, hi, n, batch, UL2, Ud, UR2, VL2, Vd, VR2, ws); break;
        default: rc = sla_rdiv_ldr_ldr(h, hi, n, batch, UL2, Ud, UR2, VL2, Vd, VR2, ws); break;
    }
    return P.done(rc);
}
// dense products / quotients with an LDR: which: 0 lmul_ldr_mat (V <- F V), 1 ldiv_ldr_mat, 2 rmul_mat_ldr (U <- U F), 3 rdiv_mat_ldr
int f32_dense(sla_handle_t h, int dtype, int n, int batch, void* M, const void* L, const double* d, const void* R, void* ws, int which) {
    PROM_BEGIN();
    void *M2 = P.io(M, NB), *L2 = P.in(L, NB), *R2 = P.in(R, NB);
    int rc;
    switch (which) {
        case 0: rc = sla_lmul_ldr_mat(h, hi, n, batch, L2, d, R2, M2, ws); break;
        case 1: rc = sla_ldiv_ldr_mat(h, hi, n, batch, L2, d, R2, M2, ws); break;
        case 2: rc = sla_rmul_mat_ldr(h, hi, n, batch, M2, L2, d, R2, ws); break;
        default: rc = sla_rdiv_mat_ldr(h, hi, n, batch, M2, L2, d, R2, ws); break;
    }
    return P.done(rc);
}
// which: 0 det_lu, 1 inv_lu, 2 ldiv_lu (B != null)
int f32_lu(sla_handle_t h, int dtype, int n, int batch, void* A, void* B, double* logdet, void* sign, void* ws, int which) {
    PROM_BEGIN();
    void *A2 = P.io(A, NB), *B2 = which == 2 ? P.io(B, NB) : nullptr, *s2 = P.out(sign, batch);
    int rc;
    switch (which) {
        case 0: rc = sla_det_lu(h, hi, n, batch, A2, logdet, s2, ws); break;
        case 1: rc = sla_inv_lu(h, hi, n, batch, A2, logdet, s2, ws); break;
        default: rc = sla_ldiv_lu(h, hi, n, batch, A2, B2, logdet, s2, ws); break;
    }
    return P.done(rc);
}
int f32_logabsdet(sla_handle_t h, int dtype, int n, int batch, const void* L, const double* d, const void* R, double* logdet, void* sign, void* ws) {
    PROM_BEGIN();
    return P.done(sla_logabsdet(h, hi, n, batch, P.in(L, NB), d, P.in(R, NB), logdet, P.out(sign, batch), ws));
}
int f32_inv_IpA(sla_handle_t h, int dtype, int n, int batch, void* G, const void* L, const double* d, const void* R, double* logdet, void* sign, void* ws) {
    PROM_BEGIN();
    return P.done(sla_inv_IpA(h, hi, n, batch, P.out(G, NB), P.in(L, NB), d, P.in(R, NB), logdet, P.out(sign, batch), ws));
}
// which: 0 inv_IpUV, 1 inv_UpV, 2 inv_invUpV
int f32_inv_uv(sla_handle_t h, int dtype, int n, int batch, void* G, const void* UL, const double* Ud, const void* UR, const void* VL,
               const double* Vd, const void* VR, double* logdet, void* sign, void* ws, int which) {
    PROM_BEGIN();
    void *G2 = P.out(G, NB), *UL2 = P.in(UL, NB), *UR2 = P.in(UR, NB);
    // U and V may be the same object (test/runtests.jl:280): promote it once so that the aliasing survives
    void *VL2 = (VL == UL) ? UL2 : P.in(VL, NB), *VR2 = (VR == UR) ? UR2 : P.in(VR, NB);
    void* s2 = P.out(sign, batch);
    int rc = which == 0 ? sla_inv_IpUV(h, hi, n, batch, G2, UL2, Ud, UR2, VL2, Vd, VR2, logdet, s2, ws)
           : which == 1 ? sla_inv_UpV(h, hi, n, batch, G2, UL2, Ud, UR2, VL2, Vd, VR2, logdet, s2, ws)
                        : sla_inv_invUpV(h, hi, n, batch, G2, UL2, Ud, UR2, VL2, Vd, VR2, logdet, s2, ws);
    return P.done(rc);
}
int f32_set_identity(sla_handle_t h, int dtype, int n, int batch, void* L, double* d, void* R) {
    // no workspace in this signature: the two identity matrices are written in the storage type directly
    cudaStream_t s = sla_handle_stream(h);
    const long long nn = (long long)n * n;
    cudaError_t e = fill_batched(d, 1.0, (long long)n * batch, s);
    if (e == cudaSuccess) e = set_identity_f32((float*)L, nn, n, batch, dtype == SLA_C32, s);
    if (e == cudaSuccess) e = set_identity_f32((float*)R, nn, n, batch, dtype == SLA_C32, s);
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}

int f32_copy(sla_handle_t h, int dtype, int n, int batch, void* Lo, double* dout, void* Ro, const void* Li, const double* din, const void* Ri) {
    if (!h) return -1;
    if (n <= 0) return -3;
    if (batch < 0) return -4;
    if (batch == 0) return 0;
    if (!Lo) return -5; if (!dout) return -6; if (!Ro) return -7; if (!Li) return -8; if (!din) return -9; if (!Ri) return -10;
    cudaStream_t s = sla_handle_stream(h);
    const size_t mb = (size_t)n * n * batch * (dtype == SLA_C32 ? 8 : 4);
    cudaError_t e = cudaSuccess;
    if (Lo != Li) e = cudaMemcpyAsync(Lo, Li, mb, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && dout != din) e = cudaMemcpyAsync(dout, din, (size_t)n * batch * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && Ro != Ri) e = cudaMemcpyAsync(Ro, Ri, mb, cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}

int f32_greens_chain(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const signed char* fields,
                     double lam, int inverse, void* G, double* logdet, void* sign, void* Lout, double* dout, void* Rout, void* ws, size_t ws_bytes) {
    if (!h) return -1;
    if (n <= 0) return -3;
    if (batch < 0) return -4;
    if (batch == 0) return 0;
    if (!ws) return -17;
    const int hi = dtype - 2;
    const size_t inner = sla_greens_chain_workspace_bytes(hi, n, batch, L, n_stab);
    if (inner == 0) return (L <= 0) ? -5 : ((n_stab <= 0 || L % n_stab) ? -6 : SLA_ENOTIMPL);
    if (ws_bytes < f32_chain_workspace_bytes(dtype, n, batch, L, n_stab)) return -17;
    const long long nn = (long long)n * n, NB = nn * batch;
    Prom P(sla_handle_stream(h), dtype, (char*)ws + al256(inner), driver_prom_bytes(hi, n, batch, 1, 1, 3));
    void *K2 = P.in(expK, nn), *G2 = P.out(G, NB), *s2 = P.out(sign, batch), *L2 = P.out(Lout, NB), *R2 = P.out(Rout, NB);
    return P.done(sla_greens_chain(h, hi, n, batch, L, n_stab, K2, fields, lam, inverse, G2, logdet, s2, L2, dout, R2, ws, inner));
}

int f32_sweep(sla_handle_t h, int dtype, int n, int batch, int L, int n_stab, const void* expK, const void* expKinv,
              const signed char* fields, double lam, void* G0, void* G, double* logdet, void* sign, double* dG, void* ws, size_t ws_bytes) {
    if (!h) return -1;
    if (n <= 0) return -3;
    if (batch < 0) return -4;
    if (batch == 0) return 0;
    if (!ws) return -16;
    const int hi = dtype - 2;
    const size_t inner = sla_sweep_workspace_bytes(hi, n, batch, L, n_stab);
    if (inner == 0) return (L <= 0) ? -5 : ((n_stab <= 0 || L % n_stab) ? -6 : SLA_ENOTIMPL);
    if (ws_bytes < f32_sweep_workspace_bytes(dtype, n, batch, L, n_stab)) return -17;
    const int ng = L / n_stab;
    const long long nn = (long long)n * n, NB = nn * batch;
    Prom P(sla_handle_stream(h), dtype, (char*)ws + al256(inner), driver_prom_bytes(hi, n, batch, ng + 1, 2, 2));
    void *K2 = P.in(expK, nn), *Ki2 = P.in(expKinv, nn), *G02 = P.out(G0, NB), *G2 = P.out(G, NB);
    void* s2 = P.out(sign, (long long)(ng + 1) * batch);
    return P.done(sla_sweep(h, hi, n, batch, L, n_stab, K2, Ki2, fields, lam, G02, G2, logdet, s2, dG, ws, inner));
}

}  // extern "C"
