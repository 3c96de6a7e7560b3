// Internal (C++) interface between the kernel translation units and the C-ABI layer.
#pragma once
#include "sla_gemm.cuh"

namespace sla {

template <typename T> cudaError_t gemm_batched(int opa, int opb, const GemmArgs<T>& p, cudaStream_t s);
// persistent A-stationary TMA kernel for a shared A (strideA = 0), real, N/N (sla_gemm_sa.cu); cudaErrorNotSupported = not this shape
cudaError_t gemm_shared_a_batched(const GemmArgs<double>& p, cudaStream_t s);

// ---- LDR factorisation (sla_ldr.cu) -------------------------------------------------------
template <typename T> struct LdrArgs {
    T* L; long long strideL; int ldl;        // in: matrix to factor, out: Q (unitary)
    const T* Ain; long long strideAin; int ldain;  // optional separate input (copied into L while the norms are read)
    double* d; long long stride_d;           // out: |diag R'|
    T* R; long long strideR; int ldr;        // out: D^-1 R' P^T
    T* Twork; long long strideT;             // scratch: ldr_twork_elems(n) elements per matrix
    T* Qwork; long long strideQ;             // scratch n*n per matrix, multi-CTA kernel only (sla_ldr_mc.cu); may be null
    long long* jpvt_out;                     // optional: 1-based pivots (n per matrix), may be null
    long long* prof;                         // optional (diagnostics): per-phase cycle counters of cluster 0, may be null
    int* info;                               // optional per-matrix info: raised from 0 to LDR_INFO_RANGE (never cleared here)
    int n; int batch;
};
// Column norms are carried SQUARED (no sqrt on the per-column critical path), unlike ?nrm2 which is safe to 1e+-308:
// a column whose squared norm overflows, is NaN, or underflows although the column is not zero is reported through
// info (include/sla_b200.h: SLA_INFO_RANGE) instead of being mis-pivoted silently.
constexpr int LDR_INFO_RANGE = -1;
__device__ __forceinline__ bool ldr_norm2_out_of_range(double ssq, bool column_nonzero) {
    return !(ssq < 1e300) || (column_nonzero && ssq < 1e-290);
}
__device__ __forceinline__ void ldr_raise_range(int* info, int b) {
    if (info) atomicCAS(info + b, 0, LDR_INFO_RANGE);
}
template <typename T> cudaError_t ldr_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> size_t ldr_twork_elems(int n);
// on-chip (thread-block cluster + DSMEM) version, sla_ldr_cluster.cu; cluster size 0 = matrix too large
template <typename T> cudaError_t ldr_cluster_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> int ldr_cluster_size(int n);
// register-resident cluster version for n <= 256, sla_ldr_reg.cu; cluster size 0 = not supported
template <typename T> cudaError_t ldr_reg_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> int ldr_reg_cluster_size(int n);
// hybrid register + shared-memory shape of the same kernel (clusters of 2; real, 128 < n <= 256); 0 = unsupported
template <typename T> cudaError_t ldr_hyb_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> int ldr_hyb_cluster_size(int n);
void ldr_n256_resident(int* reg_clusters, int* hyb_clusters);
// blocked panel + tensor-core trailing update (sla_ldr_blk.cu; real, 128 < n <= 256, clusters of 2); 0 = unsupported.
// Factorises and forms Q (sla_orgq.cu); needs p.Twork with strideT >= n.
template <typename T> cudaError_t ldr_blk_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> int ldr_blk_cluster_size(int n);
// multi-CTA L2-streaming kernel for large n and small batches (sla_ldr_mc.cu): a cluster of cs CTAs per matrix, needs p.Qwork
template <typename T> cudaError_t ldr_mc_batched(const LdrArgs<T>& p, int cs, cudaStream_t s);
template <typename T> bool ldr_mc_supported(int n);
// panel-restricted pivoting (sla_ldr_pp.cu): NOT the ?geqp3 pivot sequence -- the opt-in mode SLA_LDR_PIVOT=panel
template <typename T> cudaError_t ldr_pp_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> bool ldr_pp_supported(int n);
// small matrices (sla_ldr_small.cu; n <= 64): one CTA per matrix, matrix in registers, one barrier per column
template <typename T> cudaError_t ldr_small_batched(const LdrArgs<T>& p, cudaStream_t s);
template <typename T> bool ldr_small_supported(int n);
// Q = H_0...H_{n-1} I on the FP64 tensor cores after a factorisation-only launch (real, n <= 256; taus in p.Twork)
cudaError_t orgq_dmma_batched(const LdrArgs<double>& p, cudaStream_t s);

// ---- LU / inverse with log|det| and sign (sla_lu.cu) ----------------------------------------
template <typename T> struct LuArgs {
    T* A; long long strideA; int lda;        // in: matrix; out: LU factors (invert=0) or A^-1 (invert=1, copy_back=1)
    T* X; long long strideX; int ldx;        // invert=1: receives A^-1 (scratch n*n per matrix)
    T* rhs; long long strideRhs; int ldrhs;  // invert=1 and rhs != null: rhs <- A^-1 rhs (ldiv_lu!, src/lu.jl:167-176)
    double* logdet; T* sign;                 // per matrix; det A (invert=0) or det A^-1 (invert=1); may be null
    long long* ipiv_out;                     // optional 1-based pivots
    int* info;                               // optional per-matrix: 0 ok, k>0 = U(k,k) exactly zero
    int invert, copy_back;
    int n; int batch;
    int pipe;                                // set by lu_batched: clusters overlap the forward sweep with the factorisation
};
template <typename T> cudaError_t lu_batched(const LuArgs<T>& p, cudaStream_t s);
template <typename T> bool lu_supported(int n);   // the panels of an n x n matrix fit the shared memory of one CTA
// n <= 64, inverse form: one CTA per matrix, Gauss-Jordan in registers (sla_lu_small.cu)
template <typename T> bool lu_small_applies(const LuArgs<T>& p);
template <typename T> cudaError_t lu_small_inverse_batched(const LuArgs<T>& p, cudaStream_t s);

// ---- element-wise / small kernels (sla_ops.cu) ---------------------------------------------
// dst = rowscale(rs) * op(src) * colscale(cs); src may be shared (stride 0); op in {OP_N, OP_C}
template <typename T>
cudaError_t scale_batched(T* dst, long long sdst, int lddst, const T* src, long long ssrc, int ldsrc, int op,
                          const double* rs, long long srs, int rs_mode, const double* cs, long long scs,
                          int cs_mode, int n, int batch, cudaStream_t s);
template <typename T>
cudaError_t set_identity_batched(T* A, long long sA, int lda, int n, int batch, cudaStream_t s);
cudaError_t fill_batched(double* d, double value, long long count, cudaStream_t s);
// M = L*diag(min(d,1)) + Rinv*diag(1/max(d,1));  Rinv <- Rinv*diag(1/max(d,1))   (exported_functions.jl:37-57)
template <typename T>
cudaError_t ipa_combine_batched(T* M, T* Rinv, const T* L, const double* d, int n, int batch, cudaStream_t s);
// out[b] = sum_i log(f(d[b,i])),  f by ScaleMode (NONE: d, MUL_MAX1: max(d,1), MUL_MIN1: min(d,1))
cudaError_t logsum_batched(double* out, const double* d, long long sd, int mode, int n, int batch, cudaStream_t s);
// logdet[b] = sum_k coef[k]*terms[k][b];  sign[b] = prod_k (conj_k ? conj : id)(signs[k][b])
template <typename T> struct CombineArgs {
    const double* lterm[4]; double lcoef[4]; int nl;
    const T* sterm[4]; int sconj[4]; int ns;
    double* logdet; T* sign; int batch;
};
template <typename T> cudaError_t combine_batched(const CombineArgs<T>& a, cudaStream_t s);
// out4 = [sum logdet, sum Re sign, sum Im sign, batch] (local part of the cross-walker all-reduce)
template <typename T> cudaError_t walker_sums_batched(double* out4, const double* logdet, const T* sign, int batch, cudaStream_t s);
// out[b] = max |A[b] - B[b]| over the nn elements of matrix b (NaN if any difference is NaN)
template <typename T> cudaError_t maxabsdiff_batched(double* out, const T* A, const T* B, long long nn, long long stride, int batch, cudaStream_t s);
// Float32 storage types: element-wise promotion / demotion (count in REALS) and identity matrices in the storage type
cudaError_t convert_f32_f64(double* dst, const float* src, long long count, cudaStream_t s);
cudaError_t convert_f64_f32(float* dst, const double* src, long long count, cudaStream_t s);
cudaError_t set_identity_f32(float* A, long long nn, int n, int batch, bool complex_, cudaStream_t s);
// out[i] = multiplier realising ScaleMode `mode` with d[i] (feeds the GEMM epilogue, which only multiplies)
cudaError_t scale_prep_batched(double* out, const double* d, int mode, long long count, cudaStream_t s);
// ev[c,l,i] = exp(lam * s[c,l,i])
cudaError_t hs_exp_batched(double* ev, const signed char* fields, double lam, long long count, cudaStream_t s);

}  // namespace sla
